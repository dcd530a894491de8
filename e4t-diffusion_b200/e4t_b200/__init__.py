"""e4t_b200 — B200-native (sm_100a) kernels behind the mkshing/e4t-diffusion module API.

`e4t_b200._lib` binds the C-ABI shared library; `e4t_b200.ops` are raw kernel wrappers;
`e4t_b200.functional` holds the torch.autograd.Function adapters used by the `e4t.*` module mirror.
"""
from . import _lib  # noqa: F401

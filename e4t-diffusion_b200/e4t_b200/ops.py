"""Raw (non-autograd) wrappers over the C-ABI entry points declared in include/e4t_b200.h.

Every function takes CUDA tensors, launches hand-written sm_100a kernels on torch's current stream and
returns torch tensors that merely own the output memory.
"""
import ctypes

import torch

from . import _lib
from ._lib import c_float, c_int, c_ll, c_void_p, ptr, stream

BF16 = torch.bfloat16
F32 = torch.float32


def _fp(t):
    return ptr(t)


def gemm(A, B, *, a_mn=False, b_mn=False, out=None, out_dtype=BF16, bias=None, rowgroup=None,
         rows_per_group=1, residual=None, alpha=1.0, splits=1, accumulate=False, force_bn=0):
    """out[b] = alpha * op(A[b]) @ op(B[b])^T (+bias +rowgroup +residual).

    K-major operands are (.., rows, K) with K contiguous; MN-major operands are (.., K, rows) with rows
    contiguous (i.e. the transposed storage).  A 2-D operand is shared across the batch.
    accumulate=True -> fp32 atomic accumulation into `out` (required for split-K).
    """
    assert A.dtype == BF16 and B.dtype == BF16
    assert A.stride(-1) == 1 and B.stride(-1) == 1
    batch = 1
    if A.dim() == 3:
        batch = A.shape[0]
    if B.dim() == 3:
        batch = max(batch, B.shape[0])
    if a_mn:
        K, M = A.shape[-2], A.shape[-1]
    else:
        M, K = A.shape[-2], A.shape[-1]
    if b_mn:
        Kb, N = B.shape[-2], B.shape[-1]
    else:
        N, Kb = B.shape[-2], B.shape[-1]
    assert K == Kb, (A.shape, B.shape, a_mn, b_mn)
    a_bs = A.stride(0) if A.dim() == 3 and batch > 1 else 0
    b_bs = B.stride(0) if B.dim() == 3 and batch > 1 else 0
    if accumulate:
        assert out is not None and out.dtype == F32
        out_mode = 2
    else:
        if out is None:
            shape = (batch, M, N) if (A.dim() == 3 or B.dim() == 3) else (M, N)
            out = torch.empty(shape, device=A.device, dtype=out_dtype)
        out_mode = 0 if out.dtype == BF16 else 1
    assert out.stride(-1) == 1
    o_bs = out.stride(0) if out.dim() == 3 and batch > 1 else 0
    r_bs = 0
    ldr = 0
    if residual is not None:
        assert residual.dtype == BF16 and residual.stride(-1) == 1
        ldr = residual.stride(-2)
        r_bs = residual.stride(0) if residual.dim() == 3 and batch > 1 else 0
    if bias is not None:
        assert bias.dtype == F32 and bias.is_contiguous()
    if rowgroup is not None:
        assert rowgroup.dtype == F32 and rowgroup.is_contiguous()
    _lib.call("e4t_gemm_bf16", ptr(A), ptr(B), ptr(out), c_int(M), c_int(N), c_int(K), c_int(batch),
              c_int(int(a_mn)), c_int(int(b_mn)), c_ll(A.stride(-2)), c_ll(B.stride(-2)), c_ll(a_bs), c_ll(b_bs),
              c_int(out_mode), c_ll(out.stride(-2)), c_ll(o_bs), ptr(bias), ptr(rowgroup), c_int(rows_per_group),
              ptr(residual), c_ll(ldr), c_ll(r_bs), c_float(alpha), c_int(splits), c_int(force_bn), stream())
    return out


def conv3x3(x, w9, *, bias=None, rowgroup=None, residual=None, out_dtype=BF16, force_bn=0):
    """3x3 stride-1 pad-1 convolution on NHWC bf16.  x: (B,H,W,Cin); w9: (9,Cout,Cin) bf16, tap = ky*3+kx."""
    assert x.dtype == BF16 and w9.dtype == BF16 and x.is_contiguous() and w9.is_contiguous()
    Bn, H, W, Cin = x.shape
    Cout = w9.shape[1]
    assert w9.shape == (9, Cout, Cin)
    out = torch.empty((Bn, H, W, Cout), device=x.device, dtype=out_dtype)
    if residual is not None:
        assert residual.dtype == BF16 and residual.is_contiguous() and residual.shape == out.shape
    _lib.call("e4t_conv3x3_bf16", ptr(x), ptr(w9), ptr(out), c_int(Bn), c_int(H), c_int(W), c_int(Cin), c_int(Cout),
              c_int(0 if out_dtype == BF16 else 1), ptr(bias), ptr(rowgroup), ptr(residual), c_int(force_bn),
              stream())
    return out

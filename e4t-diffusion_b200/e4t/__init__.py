"""`e4t` — drop-in mirror of the mkshing/e4t-diffusion module API (e4t.weightoffsets, e4t.models.*, e4t.encoder,
e4t.utils) whose device compute runs on the hand-written sm_100a kernels of e4t_b200 through a thin C-ABI.
Same import paths, constructor signatures, attribute names and state-dict keys as the reference, so
pretrain_e4t.py's training loop (lines 595-654) drops in unchanged."""

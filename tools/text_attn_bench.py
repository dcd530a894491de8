"""CLIP-L text tower attention (B=16, 12 heads x 64, 77 tokens, causal): the scalar small-attention kernels against the
tensor-core kernels at the same shape (which have no causal mask: timing reference only).   python tools/text_attn_bench.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "e4t-diffusion_b200")]
import torch  # noqa: E402

from e4t_b200 import ops  # noqa: E402
from gemm_probe import timeit  # noqa: E402


def main():
    B, H, N, dh = 16, 12, 77, 64
    C = H * dh
    q, k, v, do = ((torch.randn(B, N, C, device="cuda") * 0.5).to(torch.bfloat16) for _ in range(4))
    o, lse = ops.attn_small_fwd(q, k, v, H, None, True)
    d = torch.empty(B, N, 3 * C, device="cuda", dtype=torch.bfloat16)
    t_f = timeit(lambda: ops.attn_small_fwd(q, k, v, H, None, True), iters=50)
    t_b = timeit(lambda: ops.attn_small_bwd(q, k, v, o, do, lse, H, None, True, dq=d[..., :C], dk=d[..., C:2 * C], dv=d[..., 2 * C:]), iters=50)
    o2, lse2 = ops.attn_fwd(q, k, v, H)
    t_f2 = timeit(lambda: ops.attn_fwd(q, k, v, H), iters=50)
    t_b2 = timeit(lambda: ops.attn_bwd(q, k, v, o2, do, lse2, H), iters=50)
    print(f"text attention 16x12x77x64: small fwd {t_f:.1f} us  bwd {t_b:.1f} us | tensor-core (no mask) fwd {t_f2:.1f} us  bwd {t_b2:.1f} us")


if __name__ == "__main__":
    main()

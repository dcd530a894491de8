"""Localise errors of the double-buffered-S forward (E4T_ATTN_FWD2=d) against the default kernel: error per 128-query tile
and per 16-column group of the head dimension, over a few (N, M, dh).   python tools/diag_fwd3.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "e4t-diffusion_b200")]
import torch  # noqa: E402

from e4t_b200 import ops  # noqa: E402


def main():
    torch.manual_seed(0)
    for (B, H, N, M, dh, peak) in [(1, 2, 256, 192, 64, 1.0), (1, 2, 256, 288, 64, 1.0), (1, 2, 256, 384, 64, 1.0),
                                   (1, 2, 256, 960, 64, 1.0), (1, 2, 256, 1000, 64, 1.0), (1, 2, 256, 960, 64, 6.0),
                                   (1, 2, 256, 960, 40, 6.0), (1, 2, 256, 960, 40, 1.0), (1, 2, 384, 1000, 64, 6.0),
                                   (1, 2, 512, 4096, 40, 1.0), (1, 2, 512, 4096, 40, 6.0), (1, 2, 256, 960, 32, 6.0)]:
        C = H * dh
        q, k, v = ((torch.randn(B, n, C, device="cuda") * 0.5).to(torch.bfloat16) for n in (N, M, M))
        q = q * peak
        os.environ["E4T_ATTN_FWD2"] = "s"
        o0, l0 = ops.attn_fwd(q, k, v, H)
        os.environ["E4T_ATTN_FWD2"] = "d"
        o1, l1 = ops.attn_fwd(q, k, v, H)
        torch.cuda.synchronize()
        d = (o1.float() - o0.float()).view(B, N, H, dh)
        ref = o0.float().view(B, N, H, dh)
        tiles = [f"{(d[:, i:i + 128].norm() / (ref[:, i:i + 128].norm() + 1e-9)).item():.3f}" for i in range(0, N, 128)]
        cols = [f"{(d[..., c:c + 16].norm() / (ref[..., c:c + 16].norm() + 1e-9)).item():.3f}" for c in range(0, dh, 16)]
        print(f"N={N} M={M} dh={dh} peak={peak}: rel {(d.norm() / ref.norm()).item():.4f}  dlse {(l1 - l0).abs().max().item():.4f}"
              f" | per 128-row tile {tiles} | per 16-col group {cols}", flush=True)
    os.environ.pop("E4T_ATTN_FWD2", None)


if __name__ == "__main__":
    main()

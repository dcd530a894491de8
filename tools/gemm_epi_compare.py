"""Plain-epilogue variants of the GEMM kernel on the step's no-bias signatures (E4T_GEMM_EPI_PLAIN = 0 general loop,
1 plain loop with 128-thread slabs, 2 plain loop with per-warp 32x32 TMA stores).   python tools/gemm_epi_compare.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "e4t-diffusion_b200")]
import torch  # noqa: E402

from e4t_b200 import ops  # noqa: E402
from gemm_probe import timeit  # noqa: E402


def main():
    bf = torch.bfloat16
    shapes = [(65536, 960, 320), (65536, 320, 320), (16384, 1920, 640), (16384, 640, 640), (4096, 3840, 1280),
              (4096, 1280, 1280), (16384, 640, 2560), (65536, 320, 1280), (1232, 640, 768), (4112, 3840, 1280)]
    for M, N, K in shapes:
        A = torch.randn(M, K, device="cuda").to(bf)
        Bm = (torch.randn(N, K, device="cuda") * 0.05).to(bf)
        row = []
        for mode in ("0", "1", "2"):
            os.environ["E4T_GEMM_EPI_PLAIN"] = mode
            t = timeit(lambda: ops.gemm(A, Bm), iters=20)
            row.append(f"epi{mode} {t:7.1f}us {2 * M * N * K / t / 1e6:7.1f} TF/s")
        os.environ.pop("E4T_GEMM_EPI_PLAIN", None)
        print(f"gemm {M}x{N}x{K}: " + " | ".join(row), flush=True)


if __name__ == "__main__":
    main()

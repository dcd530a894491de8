"""Residual-epilogue GEMMs of the step (to_out / ff-out with bias + residual) against the same shapes without the residual.
    python tools/gemm_res_bench.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "e4t-diffusion_b200")]
import torch  # noqa: E402

from e4t_b200 import ops  # noqa: E402
from gemm_probe import timeit  # noqa: E402


def main():
    bf = torch.bfloat16
    for M, N, K in [(65536, 320, 320), (16384, 640, 640), (4096, 1280, 1280), (65536, 320, 1280), (4112, 1280, 5120),
                    (4112, 1280, 1280)]:
        A = torch.randn(M, K, device="cuda").to(bf)
        W = (torch.randn(N, K, device="cuda") * 0.05).to(bf)
        bias = torch.randn(N, device="cuda")
        R = torch.randn(M, N, device="cuda").to(bf)
        y = ops.gemm(A, W, bias=bias, residual=R)
        ref = (A.float() @ W.float().t() + bias + R.float())
        err = ((y.float() - ref).norm() / ref.norm()).item()
        t0 = timeit(lambda: ops.gemm(A, W), iters=20)
        t1 = timeit(lambda: ops.gemm(A, W, bias=bias), iters=20)
        t2 = timeit(lambda: ops.gemm(A, W, bias=bias, residual=R), iters=20)
        gb = (M * K + 2 * M * N) * 2 / 1e9
        print(f"gemm {M}x{N}x{K}: plain {t0:6.1f} us | +bias {t1:6.1f} us | +bias+residual {t2:6.1f} us "
              f"({gb / t2 * 1e6:.0f} GB/s of A+R+C traffic)  rel err {err:.2e}", flush=True)


if __name__ == "__main__":
    main()

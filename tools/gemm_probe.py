"""Where does the GEMM mainloop time go?  Times one conv / GEMM signature under the engine's probe switches
(E4T_GEMM_DEBUG: 1 = no epilogue, 8 = no TMA operand loads, 16 = no MMAs) at several tile widths.  Results are garbage by
construction; only the timings matter.   python tools/gemm_probe.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "e4t-diffusion_b200")]
import torch  # noqa: E402

from e4t_b200 import ops  # noqa: E402


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3   # us


def main():
    bf = torch.bfloat16
    x = torch.randn(16, 32, 32, 640, device="cuda").to(bf)
    w9 = (torch.randn(9, 640, 640, device="cuda") * 0.02).to(bf)
    A = torch.randn(65536, 320, device="cuda").to(bf)
    Bm = (torch.randn(960, 320, device="cuda") * 0.05).to(bf)
    A2 = torch.randn(16384, 2560, device="cuda").to(bf)
    B2 = (torch.randn(640, 2560, device="cuda") * 0.05).to(bf)
    cases = [("conv 640->640 @32 (K=5760)", lambda bn: ops.conv3x3(x, w9, force_bn=bn), 2 * 16384 * 640 * 5760),
             ("gemm 65536x960x320", lambda bn: ops.gemm(A, Bm, force_bn=bn), 2 * 65536 * 960 * 320),
             ("gemm 16384x640x2560", lambda bn: ops.gemm(A2, B2, force_bn=bn), 2 * 16384 * 640 * 2560)]
    for name, fn, fl in cases:
        for bn in (64, 128, 192, 256):
            row = []
            for dbg, tag in ((0, "full"), (1, "no-epi"), (8, "no-tma"), (16, "no-mma"), (9, "no-tma,no-epi"), (24, "barriers only")):
                os.environ["E4T_GEMM_DEBUG"] = str(dbg)
                t = timeit(lambda: fn(bn))
                row.append(f"{tag} {t:7.1f}us")
            os.environ["E4T_GEMM_DEBUG"] = "0"
            t = timeit(lambda: fn(bn))
            print(f"{name:28s} BN={bn:3d}  {fl / t / 1e6:7.1f} TF/s | " + " | ".join(row), flush=True)


if __name__ == "__main__":
    main()

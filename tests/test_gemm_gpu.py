"""GPU parity of the tcgen05 GEMM engine (C-ABI e4t_gemm_bf16 / e4t_conv3x3_bf16) against an fp32 restatement.

Inputs are bf16-representable, so the only differences from the fp32 reference are accumulation order and the
final rounding: tolerance 2e-3 relative (of the output RMS) for fp32 outputs, one bf16 ulp (2^-8) for bf16 outputs.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False


def _rel(a, b):
    a = a.float(); b = b.float()
    return ((a - b).pow(2).mean().sqrt() / b.pow(2).mean().sqrt().clamp_min(1e-12)).item()


def _mk(shape, g, scale=1.0):
    return (torch.randn(shape, generator=g, device="cuda") * scale).to(torch.bfloat16)


@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 320, 320), (1000, 640, 1232), (4096, 1280, 768), (77, 320, 768)])
def test_gemm_majors(a_mn, b_mn, M, N, K):
    from e4t_b200 import ops
    if (a_mn and M % 8) or (b_mn and N % 8):
        pytest.skip("MN-major operands need 16-byte aligned rows")
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    A = _mk((M, K), g); B = _mk((N, K), g)
    ref = A.float() @ B.float().t()
    Ain = A.t().contiguous() if a_mn else A
    Bin = B.t().contiguous() if b_mn else B
    out = ops.gemm(Ain, Bin, a_mn=a_mn, b_mn=b_mn, out_dtype=torch.float32)
    torch.cuda.synchronize()
    assert _rel(out, ref) < 2e-3, (_rel(out, ref))
    out16 = ops.gemm(Ain, Bin, a_mn=a_mn, b_mn=b_mn)
    assert _rel(out16, ref) < 4e-3


@pytest.mark.parametrize("bn", [64, 128, 160, 256])
def test_gemm_tile_widths_and_epilogue(bn):
    from e4t_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(bn)
    M, N, K = 2048, 640, 640
    A = _mk((M, K), g); B = _mk((N, K), g)
    bias = torch.randn(N, generator=g, device="cuda")
    rg = torch.randn(M // 256, N, generator=g, device="cuda")
    res = _mk((M, N), g)
    ref = 0.5 * (A.float() @ B.float().t()) + bias + rg.repeat_interleave(256, 0) + res.float()
    out = ops.gemm(A, B, bias=bias, rowgroup=rg, rows_per_group=256, residual=res, alpha=0.5,
                   out_dtype=torch.float32, force_bn=bn)
    assert _rel(out, ref) < 2e-3


def test_gemm_batched_and_splitk():
    from e4t_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(5)
    Bt, M, N, K = 3, 300, 200, 520
    A = _mk((Bt, M, K), g); B = _mk((Bt, N, K), g)
    ref = torch.einsum("bmk,bnk->bmn", A.float(), B.float())
    out = ops.gemm(A, B, out_dtype=torch.float32)
    assert _rel(out, ref) < 2e-3
    # shared B
    out = ops.gemm(A, B[0], out_dtype=torch.float32)
    assert _rel(out, torch.einsum("bmk,nk->bmn", A.float(), B[0].float())) < 2e-3
    # split-K weight-gradient shape: dW[C,R] = dY[m,C]^T X[m,R], both MN-major, accumulate twice
    m, C, R = 8192, 320, 768
    dY = _mk((m, C), g); X = _mk((m, R), g)
    acc = torch.zeros(C, R, device="cuda")
    ops.gemm(dY, X, a_mn=True, b_mn=True, out=acc, accumulate=True, splits=16)
    ops.gemm(dY, X, a_mn=True, b_mn=True, out=acc, accumulate=True, splits=7)
    ref = 2 * (dY.float().t() @ X.float())
    assert _rel(acc, ref) < 2e-3


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(2, 64, 64, 320, 320), (2, 32, 32, 640, 320), (3, 16, 16, 1280, 640),
                                            (4, 8, 8, 1280, 1280), (3, 8, 8, 64, 96), (1, 64, 64, 64, 32)])
def test_conv3x3(B, H, W, Cin, Cout):
    from e4t_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(B + H + Cin)
    x = _mk((B, H, W, Cin), g)
    w = _mk((Cout, Cin, 3, 3), g, scale=0.05)
    bias = torch.randn(Cout, generator=g, device="cuda")
    temb = torch.randn(B, Cout, generator=g, device="cuda")
    res = _mk((B, H, W, Cout), g)
    ref = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), w.float(), bias, padding=1)
    ref = ref + temb[:, :, None, None] + res.float().permute(0, 3, 1, 2)
    w9 = w.permute(2, 3, 0, 1).reshape(9, Cout, Cin).contiguous()
    out = ops.conv3x3(x, w9, bias=bias, rowgroup=temb, residual=res, out_dtype=torch.float32)
    assert _rel(out.permute(0, 3, 1, 2), ref) < 2e-3

"""Parity of the runtime-selectable kernel variants against each other (all run on the GPU every round).

Round 2 made the round-1 opt-in variants the defaults after a GPU parity + timing run
(profiles/r02_gpurun49_optin_sweep.log): P / P^T through TMEM, dQ through a TMA reduce-add, thread-per-row delta,
plain GEMM epilogue, and added the two-tile forward (attention_fwd2.cu).  The library re-reads its E4T_* switches on
every call, so the LEGACY code paths (switch = "0") are compared with the defaults inside one process; every default is
separately held to an fp32 torch reference in tests/test_kernels_gpu.py.
"""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

SWITCHES = ("E4T_ATTN_DQ_TMA", "E4T_ATTN_DELTA2", "E4T_ATTN_PT_TMEM", "E4T_ATTN_FWD_PT", "E4T_GEMM_EPI_PLAIN",
            "E4T_ATTN_CG", "E4T_ATTN_FWD2", "E4T_ATTN_BWD_REORD")


@pytest.fixture(autouse=True)
def _clean_env():
    for k in SWITCHES:
        os.environ.pop(k, None)
    yield
    for k in SWITCHES:
        os.environ.pop(k, None)


def _mk(shape, g, s=0.5):
    return (torch.randn(*shape, device="cuda", generator=g) * s).to(torch.bfloat16)


def _rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


@pytest.mark.parametrize("env", [{"E4T_ATTN_DQ_TMA": "0"}, {"E4T_ATTN_DELTA2": "0"}, {"E4T_ATTN_PT_TMEM": "0"},
                                 {"E4T_ATTN_BWD_REORD": "0"},
                                 {"E4T_ATTN_PT_TMEM": "0", "E4T_ATTN_DQ_TMA": "0", "E4T_ATTN_DELTA2": "0"}])
@pytest.mark.parametrize("B,H,N,M,dh", [(2, 8, 256, 256, 40), (2, 8, 1024, 77, 40), (1, 8, 300, 200, 40),
                                        (1, 4, 384, 128, 64), (1, 8, 1024, 1024, 80), (1, 8, 4096, 4096, 40)])
def test_attention_backward_legacy_paths_match_default(env, B, H, N, M, dh):
    from e4t_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(N + M + dh)
    C = H * dh
    q, k, v, do = _mk((B, N, C), g), _mk((B, M, C), g), _mk((B, M, C), g), _mk((B, N, C), g)
    o, lse = ops.attn_fwd(q, k, v, H)
    ref = ops.attn_bwd(q, k, v, o, do, lse, H)
    torch.cuda.synchronize()
    os.environ.update(env)
    got = ops.attn_bwd(q, k, v, o, do, lse, H)
    torch.cuda.synchronize()
    for name, a, b in zip(("dq", "dk", "dv"), got, ref):
        assert _rel(a, b) < 2e-3, (name, env, _rel(a, b))     # same math, different accumulation order for dQ


def _attn_ref(q, k, v, H):
    B, N, C = q.shape
    dh = C // H
    qh, kh, vh = (t.float().view(B, -1, H, dh).transpose(1, 2) for t in (q, k, v))
    s = (qh @ kh.transpose(-1, -2)) * dh ** -0.5
    return (s.softmax(-1) @ vh).transpose(1, 2).reshape(B, N, C), torch.logsumexp(s, -1)


@pytest.mark.parametrize("fwd2", ["0", "d", "dp0", "dp3", "s", "sn", "sf", "sp2"])
@pytest.mark.parametrize("B,H,N,M,dh", [(2, 8, 256, 256, 40), (1, 8, 300, 200, 80), (2, 16, 257, 257, 80),
                                        (1, 8, 1024, 1024, 80), (2, 4, 384, 1000, 64), (1, 4, 256, 128, 128),
                                        (1, 8, 4096, 4096, 40)])
def test_attention_forward_variants_vs_fp32_reference(fwd2, B, H, N, M, dh):
    """Forward variants against an fp32 torch softmax(QK^T)V: "0" the single-tile kernel, "d" the double-buffered-S kernel
    (the default where it applies: dh <= 64, M >= 192) with its exp2 pipe splits, "s" the single-buffered two-tile kernel
    (default elsewhere) with / without the exp-phase token, in its 4 x 64 shape and with a polynomial share.
    A strongly peaked row (scores spread over > 2^8 in the exp2 domain) exercises the threshold rescale, and the
    polynomial exp2 is exercised far below its clamp."""
    from e4t_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(N + M + dh)
    C = H * dh
    q, k, v = _mk((B, N, C), g), _mk((B, M, C), g), _mk((B, M, C), g)
    q[:, : N // 2] *= 6.0          # peaked rows: large score range, the running max keeps moving
    oref, lse_ref = _attn_ref(q, k, v, H)
    os.environ["E4T_ATTN_FWD2"] = fwd2
    o, lse = ops.attn_fwd(q, k, v, H)
    torch.cuda.synchronize()
    assert _rel(o, oref) < 6e-3, (fwd2, _rel(o, oref))
    assert (lse - lse_ref).abs().max().item() < 2e-2


@pytest.mark.parametrize("B,H,N,M,dh", [(2, 8, 1024, 77, 40), (2, 8, 256, 256, 160), (2, 4, 64, 77, 32)])
def test_attention_forward_p_in_smem_matches_default(B, H, N, M, dh):
    from e4t_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(N + M + dh)
    C = H * dh
    q, k, v = _mk((B, N, C), g), _mk((B, M, C), g), _mk((B, M, C), g)
    o0, lse0 = ops.attn_fwd(q, k, v, H)
    os.environ["E4T_ATTN_FWD_PT"] = "0"
    o1, lse1 = ops.attn_fwd(q, k, v, H)
    torch.cuda.synchronize()
    assert torch.equal(lse0, lse1)
    assert _rel(o1, o0) < 1e-3, _rel(o1, o0)


@pytest.mark.parametrize("B,H,N,M,dh", [(2, 8, 1024, 77, 40), (1, 8, 4096, 77, 40), (2, 4, 300, 128, 64)])
def test_attention_forward_single_key_block_one_vs_two_ctas_per_sm(B, H, N, M, dh):
    """Cross-attention shapes (one key block) run two CTAs per SM by default; the one-CTA launch must give the same result."""
    from e4t_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(N + M + dh)
    C = H * dh
    q, k, v = _mk((B, N, C), g), _mk((B, M, C), g), _mk((B, M, C), g)
    o0, lse0 = ops.attn_fwd(q, k, v, H)
    oref, lse_ref = _attn_ref(q, k, v, H)
    os.environ["E4T_ATTN_CG"] = "4,4,4,1"
    o1, lse1 = ops.attn_fwd(q, k, v, H)
    torch.cuda.synchronize()
    assert _rel(o0, oref) < 6e-3 and (lse0 - lse_ref).abs().max().item() < 2e-2
    assert _rel(o1, o0) < 2e-3 and (lse1 - lse0).abs().max().item() < 1e-4


@pytest.mark.parametrize("M,N,K,b_mn", [(4096, 960, 320, False), (8192, 320, 320, False), (1000, 328, 192, False),
                                        (4096, 320, 960, True), (256, 64, 64, False)])
@pytest.mark.parametrize("with_bias", [False, True])
def test_gemm_plain_epilogue_is_bit_identical(M, N, K, b_mn, with_bias):
    """The lean epilogue loop (no addend, or a bias only) against the general loop on the same inputs."""
    from e4t_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = _mk((M, K), g, 0.2)
    Bm = _mk((K, N) if b_mn else (N, K), g, 0.2)
    bias = torch.randn(N, device="cuda", generator=g) if with_bias else None
    y1 = ops.gemm(A, Bm, b_mn=b_mn, bias=bias)
    os.environ["E4T_GEMM_EPI_PLAIN"] = "0"
    y0 = ops.gemm(A, Bm, b_mn=b_mn, bias=bias)
    torch.cuda.synchronize()
    assert torch.equal(y0, y1)

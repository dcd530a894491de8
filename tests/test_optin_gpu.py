"""Parity of the opt-in kernel variants against the default kernels.

These variants are OFF by default and have not all been run on a GPU yet, so the module is skipped unless
E4T_TEST_OPTIN=1 is set:   E4T_TEST_OPTIN=1 python -m pytest tests/test_optin_gpu.py -m gpu -q
The library re-reads its E4T_* switches on every call, so the variants can be toggled inside one process.
"""
import os

import pytest
import torch

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("E4T_TEST_OPTIN") != "1", reason="opt-in variants: set E4T_TEST_OPTIN=1")]

SWITCHES = ("E4T_ATTN_PP", "E4T_ATTN_DQ_TMA", "E4T_ATTN_DELTA2", "E4T_ATTN_PT_TMEM", "E4T_ATTN_FWD_PT", "E4T_GEMM_EPI_PLAIN",
            "E4T_ATTN_CG")


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


@pytest.mark.parametrize("env", [{"E4T_ATTN_PP": "1"}, {"E4T_ATTN_PP": "2"}, {"E4T_ATTN_PP": "3"}, {"E4T_ATTN_PP": "4"},
                                 {"E4T_ATTN_DQ_TMA": "1"}, {"E4T_ATTN_DELTA2": "1"},
                                 {"E4T_ATTN_PT_TMEM": "1"}, {"E4T_ATTN_PT_TMEM": "1", "E4T_ATTN_DQ_TMA": "1"},
                                 {"E4T_ATTN_DQ_TMA": "1", "E4T_ATTN_DELTA2": "1"}])
@pytest.mark.parametrize("B,H,N,M,dh", [(2, 8, 256, 256, 40), (2, 8, 1024, 77, 40), (2, 4, 256, 256, 16),
                                        (1, 8, 300, 200, 40), (1, 4, 384, 128, 64), (1, 8, 1024, 1024, 80),
                                        (1, 8, 4096, 4096, 40)])
def test_attention_backward_variants_match_default(env, B, H, N, M, dh):
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


@pytest.mark.parametrize("cg", [None, "4,4,4,0", "2,4,4,0"])
@pytest.mark.parametrize("B,H,N,M,dh", [(2, 8, 256, 256, 40), (2, 8, 1024, 77, 40), (1, 8, 300, 200, 80),
                                        (2, 8, 256, 256, 160), (2, 4, 64, 77, 32), (1, 8, 4096, 4096, 40)])
def test_attention_forward_p_in_tmem_matches_default(cg, B, H, N, M, dh):
    from e4t_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(N + M + dh)
    C = H * dh
    q, k, v = _mk((B, N, C), g), _mk((B, M, C), g), _mk((B, M, C), g)
    if cg:
        os.environ["E4T_ATTN_CG"] = cg
    o0, lse0 = ops.attn_fwd(q, k, v, H)
    os.environ["E4T_ATTN_FWD_PT"] = "1"
    o1, lse1 = ops.attn_fwd(q, k, v, H)
    torch.cuda.synchronize()
    assert torch.equal(lse0, lse1)
    assert _rel(o1, o0) < 1e-3, _rel(o1, o0)      # same P values, same MMAs: expected bit-identical


@pytest.mark.parametrize("M,N,K,b_mn", [(4096, 960, 320, False), (8192, 320, 320, False), (4096, 2560, 320, False),
                                        (1000, 328, 192, False), (4096, 320, 960, True), (256, 64, 64, False)])
def test_gemm_plain_epilogue_is_bit_identical(M, N, K, b_mn):
    from e4t_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = _mk((M, K), g, 0.2)
    Bm = _mk((K, N) if b_mn else (N, K), g, 0.2)
    y0 = ops.gemm(A, Bm, b_mn=b_mn)
    os.environ["E4T_GEMM_EPI_PLAIN"] = "1"
    y1 = ops.gemm(A, Bm, b_mn=b_mn)
    torch.cuda.synchronize()
    assert torch.equal(y0, y1)

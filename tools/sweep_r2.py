"""One-call tuning sweep (meant as the FIRST gpurun call of a round: ~60-90 s of GPU time).

    python tools/sweep_r2.py [gemm] [attn] [norm]          # default: all three; writes gpurun_out/sweep_r2.json

gemm : records every e4t_gemm_bf16 / e4t_conv3x3_bf16 signature of one eager pre-training step (B=16) with its call
       count, then times each signature alone at the cost model's tile width and at every forced BN
       -> where the step's GEMM time goes by shape, and which shapes the tile-width model gets wrong.
attn : level-0/1/2 self- and cross-attention shapes under the runtime-selectable variants
       (E4T_ATTN_CG forward/backward column groups and 2-CTA/SM, E4T_ATTN_PP ping-pong, E4T_ATTN_DQ_TMA),
       each checked against the default kernel's output before it is timed.
norm : GroupNorm forward/backward at the UNet's shapes under E4T_GN_ROWS / E4T_GN_THREADS.

Everything is timed with CUDA events after warm-up; inputs of consecutive iterations rotate over enough buffers to
exceed the 126 MB L2 where the tensors are small.  Nothing here is a bench value; it ranks variants.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "e4t-diffusion_b200")]
OUT = os.path.join(ROOT, "gpurun_out")

import torch  # noqa: E402

from e4t_b200 import ops  # noqa: E402

dev = "cuda"
BF16, F32 = torch.bfloat16, torch.float32


def bf(*s, scale=1.0):
    return (torch.randn(*s, device=dev) * scale).to(BF16)


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


# ------------------------------------------------------------------------------------------------------------------
# GEMM / conv signatures of one step
# ------------------------------------------------------------------------------------------------------------------
def record_step_signatures(B=16):
    import bench
    from e4t_b200.engine import PretrainStep
    sigs = {}
    real_gemm, real_conv = ops.gemm, ops.conv3x3

    def gemm(A, Bm, *, a_mn=False, b_mn=False, out=None, out_dtype=BF16, bias=None, rowgroup=None, rows_per_group=1,
             residual=None, alpha=1.0, splits=1, accumulate=False, force_bn=0):
        batch = max(A.shape[0] if A.dim() == 3 else 1, Bm.shape[0] if Bm.dim() == 3 else 1)
        K, M = (A.shape[-2], A.shape[-1]) if a_mn else (A.shape[-1], A.shape[-2])
        N = Bm.shape[-1] if b_mn else Bm.shape[-2]
        odt = "f32acc" if accumulate else ("bf16" if (out.dtype if out is not None else out_dtype) == BF16 else "f32")
        key = ("gemm", M, N, K, batch, int(a_mn), int(b_mn), odt, int(bias is not None), int(rowgroup is not None),
               int(residual is not None), splits, A.dim(), Bm.dim())
        sigs[key] = sigs.get(key, 0) + 1
        return real_gemm(A, Bm, a_mn=a_mn, b_mn=b_mn, out=out, out_dtype=out_dtype, bias=bias, rowgroup=rowgroup,
                         rows_per_group=rows_per_group, residual=residual, alpha=alpha, splits=splits,
                         accumulate=accumulate, force_bn=force_bn)

    def conv3x3(x, w9, *, bias=None, rowgroup=None, residual=None, out_dtype=BF16, force_bn=0):
        Bn, H, W, Cin = x.shape
        key = ("conv", Bn, H, W, Cin, w9.shape[1], "bf16" if out_dtype == BF16 else "f32", int(bias is not None),
               int(rowgroup is not None), int(residual is not None))
        sigs[key] = sigs.get(key, 0) + 1
        return real_conv(x, w9, bias=bias, rowgroup=rowgroup, residual=residual, out_dtype=out_dtype,
                         force_bn=force_bn)

    unet, enc, text = bench.build_models(torch.device(dev))
    step = PretrainStep(unet, enc, text, placeholder_token_id=49408, class_token_id=320, lr=1.6e-5,
                        weight_dtype=BF16)
    batch = bench.to_device(bench.host_batch(B, 42, pinned=False), dev)
    step(batch)                       # warm-up: operand caches, first-call attributes
    torch.cuda.synchronize()
    ops.gemm, ops.conv3x3 = gemm, conv3x3
    try:
        step(batch)
        torch.cuda.synchronize()
    finally:
        ops.gemm, ops.conv3x3 = real_gemm, real_conv
    del step, unet, enc, text, batch
    torch.cuda.empty_cache()
    return sigs


def time_gemm_signature(key, bns):
    kind = key[0]
    res = {}
    if kind == "gemm":
        _, M, N, K, batch, a_mn, b_mn, odt, has_b, has_rg, has_res, splits, adim, bdim = key
        lead_a = (batch,) if adim == 3 else ()
        lead_b = (batch,) if bdim == 3 else ()
        A = bf(*lead_a, *((K, M) if a_mn else (M, K)), scale=0.1)
        Bm = bf(*lead_b, *((K, N) if b_mn else (N, K)), scale=0.1)
        oshape = ((batch, M, N) if (adim == 3 or bdim == 3) else (M, N))
        out = torch.zeros(oshape, device=dev, dtype=F32) if odt != "bf16" else None
        bias = torch.randn(N, device=dev) if has_b else None
        rg = torch.randn(max(M // 4096, 1), N, device=dev) if has_rg else None
        resid = bf(*oshape) if has_res else None
        kw = dict(a_mn=bool(a_mn), b_mn=bool(b_mn), bias=bias, residual=resid, splits=splits)
        if rg is not None:
            kw.update(rowgroup=rg, rows_per_group=max(M // rg.shape[0], 1))
        if odt == "f32acc":
            kw.update(out=out, accumulate=True)
        elif odt == "f32":
            kw.update(out=out)
        for bn in bns:
            if bn and (bn > 256 or (b_mn and bn % 64)):
                continue
            try:
                res[bn] = timeit(lambda: ops.gemm(A, Bm, force_bn=bn, **kw))
            except Exception as ex:      # a forced width the kernel rejects for this shape
                res[bn] = f"err: {str(ex)[:80]}"
        if odt == "bf16" and not (has_b or has_rg or has_res):
            # opt-in plain epilogue loop (E4T_GEMM_EPI_PLAIN=1): must be bit-identical to the general loop
            try:
                y0 = ops.gemm(A, Bm, **kw)
                os.environ["E4T_GEMM_EPI_PLAIN"] = "1"
                y1 = ops.gemm(A, Bm, **kw)
                torch.cuda.synchronize()
                res["plain"] = timeit(lambda: ops.gemm(A, Bm, **kw))
                res["plain_equal"] = bool(torch.equal(y0, y1))
            except Exception as ex:
                res["plain"] = f"err: {str(ex)[:80]}"
            finally:
                os.environ.pop("E4T_GEMM_EPI_PLAIN", None)
        flops = 2.0 * M * N * K * batch
    else:
        _, Bn, H, W, Cin, Cout, odt, has_b, has_rg, has_res = key
        x = bf(Bn, H, W, Cin, scale=0.5)
        w9 = bf(9, Cout, Cin, scale=0.05)
        bias = torch.randn(Cout, device=dev) if has_b else None
        rg = torch.randn(Bn, Cout, device=dev) if has_rg else None
        resid = bf(Bn, H, W, Cout) if has_res else None
        for bn in bns:
            try:
                res[bn] = timeit(lambda: ops.conv3x3(x, w9, bias=bias, rowgroup=rg, residual=resid,
                                                     out_dtype=BF16 if odt == "bf16" else F32, force_bn=bn))
            except Exception as ex:
                res[bn] = f"err: {str(ex)[:80]}"
        flops = 2.0 * Bn * H * W * 9 * Cin * Cout
    return res, flops


def sweep_gemm():
    t0 = time.time()
    sigs = record_step_signatures()
    rows = []
    bns = [0, 64, 96, 128, 160, 192, 224, 256]
    for key, cnt in sigs.items():
        res, flops = time_gemm_signature(key, bns)
        t_def = res.get(0)
        good = {bn: t for bn, t in res.items() if isinstance(t, float) and isinstance(bn, int)}
        best_bn = min(good, key=good.get) if good else None
        rows.append({"sig": list(key), "calls": cnt, "ms_default": t_def, "tflops_default":
                     (flops / t_def / 1e9) if isinstance(t_def, float) else None, "best_bn": best_bn,
                     "ms_best": good.get(best_bn), "by_bn": {str(k): v for k, v in res.items()},
                     "ms_plain_epilogue": res.get("plain"), "plain_equal": res.get("plain_equal"),
                     "step_ms_default": cnt * t_def if isinstance(t_def, float) else None,
                     "step_ms_best": cnt * good[best_bn] if good else None})
    rows.sort(key=lambda r: -(r["step_ms_default"] or 0))
    tot = sum(r["step_ms_default"] or 0 for r in rows)
    totb = sum(r["step_ms_best"] or 0 for r in rows)
    print(f"[gemm] {len(rows)} signatures, sum(calls x default) = {tot:.2f} ms/step, with best forced BN {totb:.2f} ms "
          f"({time.time()-t0:.0f}s)")
    for r in rows[:70]:
        print(f"  {r['calls']:3d} x {r['ms_default']*1e3 if r['ms_default'] else 0:7.1f} us = {r['step_ms_default'] or 0:6.2f} ms  "
              f"{r['tflops_default'] or 0:6.0f} TF/s  best bn {r['best_bn']} {((r['ms_best'] or 0)*1e3):7.1f} us  "
              f"plain-epi {r['ms_plain_epilogue']} eq={r['plain_equal']}  {r['sig']}")
    return {"rows": rows, "sum_default_ms": tot, "sum_best_ms": totb}


# ------------------------------------------------------------------------------------------------------------------
# attention variants
# ------------------------------------------------------------------------------------------------------------------
def set_env(**kw):
    for k, v in kw.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = str(v)


def sweep_attn(B=16):
    out = []
    shapes = [(4096, 4096, 320), (1024, 1024, 640), (256, 256, 1280), (4096, 77, 320), (1024, 77, 640)]
    fwd_variants = {"default": {}, "cg4_occ1": {"E4T_ATTN_CG": "4,4,4,0"}, "cg2_occ1": {"E4T_ATTN_CG": "2,4,4,0"},
                    "pt": {"E4T_ATTN_FWD_PT": 1}, "cg4_occ1+pt": {"E4T_ATTN_CG": "4,4,4,0", "E4T_ATTN_FWD_PT": 1}}
    bwd_variants = {"default": {}, "pp1": {"E4T_ATTN_PP": 1}, "pp2": {"E4T_ATTN_PP": 2}, "pp4": {"E4T_ATTN_PP": 4},
                    "dq_tma": {"E4T_ATTN_DQ_TMA": 1}, "fused_cg2": {"E4T_ATTN_CG": "4,2,4,1"},
                    "delta2": {"E4T_ATTN_DELTA2": 1}, "pt_tmem": {"E4T_ATTN_PT_TMEM": 1},
                    "pt+dq": {"E4T_ATTN_PT_TMEM": 1, "E4T_ATTN_DQ_TMA": 1},
                    "pt+dq+delta2": {"E4T_ATTN_PT_TMEM": 1, "E4T_ATTN_DQ_TMA": 1, "E4T_ATTN_DELTA2": 1}}
    for (N, M, C) in shapes:
        q, k, v, do = bf(B, N, C, scale=0.5), bf(B, M, C, scale=0.5), bf(B, M, C, scale=0.5), bf(B, N, C)
        set_env(E4T_ATTN_CG=None, E4T_ATTN_PP=None, E4T_ATTN_DQ_TMA=None, E4T_ATTN_DELTA2=None, E4T_ATTN_PT_TMEM=None)
        o0, lse0 = ops.attn_fwd(q, k, v, 8)
        g0 = ops.attn_bwd(q, k, v, o0, do, lse0, 8)
        torch.cuda.synchronize()
        row = {"N": N, "M": M, "dh": C // 8, "fwd": {}, "bwd": {}}
        for name, env in fwd_variants.items():
            set_env(E4T_ATTN_CG=None, E4T_ATTN_FWD_PT=None)
            set_env(**env)
            try:
                o, lse = ops.attn_fwd(q, k, v, 8)
                torch.cuda.synchronize()
                row["fwd"][name] = {"err": rel(o, o0), "ms": timeit(lambda: ops.attn_fwd(q, k, v, 8))}
            except Exception as ex:
                row["fwd"][name] = {"error": str(ex)[:120]}
        set_env(E4T_ATTN_CG=None, E4T_ATTN_FWD_PT=None)
        for name, env in bwd_variants.items():
            set_env(E4T_ATTN_CG=None, E4T_ATTN_PP=None, E4T_ATTN_DQ_TMA=None, E4T_ATTN_DELTA2=None, E4T_ATTN_PT_TMEM=None)
            set_env(**env)
            try:
                g = ops.attn_bwd(q, k, v, o0, do, lse0, 8)
                torch.cuda.synchronize()
                errs = [rel(a, b) for a, b in zip(g, g0)]
                row["bwd"][name] = {"err": max(errs), "ms": timeit(lambda: ops.attn_bwd(q, k, v, o0, do, lse0, 8))}
            except Exception as ex:
                row["bwd"][name] = {"error": str(ex)[:120]}
        set_env(E4T_ATTN_CG=None, E4T_ATTN_PP=None, E4T_ATTN_DQ_TMA=None, E4T_ATTN_DELTA2=None, E4T_ATTN_PT_TMEM=None)
        print(f"[attn] N={N} M={M} dh={C//8}: fwd " +
              " ".join(f"{n}={d.get('ms', float('nan')):.3f}" for n, d in row["fwd"].items()) + " | bwd " +
              " ".join(f"{n}={d.get('ms', float('nan')):.3f}(e{d.get('err', 0):.0e})" for n, d in row["bwd"].items()))
        out.append(row)
    return out


# ------------------------------------------------------------------------------------------------------------------
# GroupNorm launch configurations
# ------------------------------------------------------------------------------------------------------------------
def sweep_norm(B=16):
    out = []
    shapes = [(4096, 320), (4096, 640), (4096, 960), (1024, 640), (1024, 1280), (1024, 1920), (256, 1280), (256, 2560)]
    for (HW, C) in shapes:
        nbuf = max(2, int(300e6 // (B * HW * C * 2)) + 1)          # rotate past L2
        xs = [bf(B, HW, C) for _ in range(min(nbuf, 6))]
        dys = [bf(B, HW, C) for _ in range(len(xs))]
        gamma, beta = torch.randn(C, device=dev), torch.randn(C, device=dev)
        set_env(E4T_GN_ROWS=None, E4T_GN_THREADS=None)
        y0, st0 = ops.groupnorm_fwd(xs[0], gamma, beta, 32, 1e-5, True)
        row = {"HW": HW, "C": C, "cfg": {}}
        for rows in (None, 16, 32, 64, 128, 256):
            for thr in (None, 128, 320):
                set_env(E4T_GN_ROWS=rows, E4T_GN_THREADS=thr)
                i = [0]

                def fwd():
                    i[0] = (i[0] + 1) % len(xs)
                    return ops.groupnorm_fwd(xs[i[0]], gamma, beta, 32, 1e-5, True)

                def bwd():
                    i[0] = (i[0] + 1) % len(xs)
                    return ops.groupnorm_bwd(xs[i[0]], dys[i[0]], gamma, beta, st0, 32, 1e-5, True)
                try:
                    y, _ = ops.groupnorm_fwd(xs[0], gamma, beta, 32, 1e-5, True)
                    torch.cuda.synchronize()
                    row["cfg"][f"rows={rows},thr={thr}"] = {"err": rel(y, y0), "fwd_ms": timeit(fwd),
                                                            "bwd_ms": timeit(bwd)}
                except Exception as ex:
                    row["cfg"][f"rows={rows},thr={thr}"] = {"error": str(ex)[:120]}
        set_env(E4T_GN_ROWS=None, E4T_GN_THREADS=None)
        d = row["cfg"]["rows=None,thr=None"]
        ok = {k: v for k, v in row["cfg"].items() if "fwd_ms" in v and v["err"] < 1e-2}
        bf_ = min(ok, key=lambda k: ok[k]["fwd_ms"])
        bb_ = min(ok, key=lambda k: ok[k]["bwd_ms"])
        gb = B * HW * C * 2 / 1e9
        print(f"[norm] HW={HW} C={C}: default fwd {d['fwd_ms']*1e3:.1f} us ({3*gb/d['fwd_ms']:.0f} GB/s) bwd "
              f"{d['bwd_ms']*1e3:.1f} us ({5*gb/d['bwd_ms']:.0f} GB/s) | best fwd {bf_} {ok[bf_]['fwd_ms']*1e3:.1f} us, "
              f"best bwd {bb_} {ok[bb_]['bwd_ms']*1e3:.1f} us")
        out.append(row)
        del xs, dys
    return out


if __name__ == "__main__":
    what = sys.argv[1:] or ["attn", "norm", "gemm"]
    os.makedirs(OUT, exist_ok=True)
    res = {}
    for w in what:
        try:
            res[w] = {"attn": sweep_attn, "norm": sweep_norm, "gemm": sweep_gemm}[w]()
        except Exception as ex:     # keep the other sections
            import traceback
            traceback.print_exc()
            res[w] = {"error": repr(ex)}
        with open(os.path.join(OUT, "sweep_r2.json"), "w") as f:
            json.dump(res, f, indent=1, default=str)

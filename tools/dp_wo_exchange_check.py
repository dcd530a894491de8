"""2-GPU check of the WeightOffsets factor exchange (engine.PretrainStep under data parallelism): two eager steps with
per-rank batches; rank 0 writes the trained-parameter arena's WeightOffsets slice, the head slice checksum and the losses.
Run once with E4T_WO_FACTOR_EXCHANGE=1 and once with =0 (slice all-reduce) and compare the files:
    torchrun --nproc-per-node 2 tools/dp_wo_exchange_check.py out.pt ;  python tools/dp_wo_exchange_check.py --compare a.pt b.pt
NOTE (round-2 call 31): comparing PARAMETERS after AdamW steps is ill-conditioned — the first Adam update is lr * sign(g),
so run-to-run rounding noise of near-zero gradients (fp32 atomics) moves entries by 2 lr: the encoder-head sample, which
the exchange mode does not touch, differed by the same 3e-3 as the WeightOffsets slice (step-1 losses equal to 1e-7).  The
exact check of the exchange is tests/test_e2e_gpu.py::test_wo_bank_two_phase_backward_is_linear_in_the_reductions."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "e4t-diffusion_b200")]
import torch  # noqa: E402


def compare(a, b):
    x, y = torch.load(a), torch.load(b)
    rel = ((x["wo"] - y["wo"]).norm() / y["wo"].norm()).item()
    relh = ((x["head"] - y["head"]).norm() / y["head"].norm()).item()
    print(f"WO slice after 2 steps: rel diff {rel:.3e} (max abs {(x['wo'] - y['wo']).abs().max().item():.3e}); "
          f"head sample rel diff {relh:.3e}; losses {x['loss']} vs {y['loss']}; flags {x['flag']} vs {y['flag']}")
    ok = abs(x["loss"][0] - y["loss"][0]) < 1e-4 * abs(y["loss"][0]) and rel < 3 * max(relh, 1e-6)
    print("consistent (WeightOffsets slice differs no more than the untouched head slice)" if ok else "MISMATCH")
    return 0 if ok else 1


def main():
    if sys.argv[1] == "--compare":
        sys.exit(compare(sys.argv[2], sys.argv[3]))
    import datetime
    import torch.distributed as dist
    import bench
    from e4t_b200.engine import PretrainStep
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=120))
    unet, enc, text = bench.build_models(dev)
    step = PretrainStep(unet, enc, text, placeholder_token_id=49408, class_token_id=320, lr=1e-3, weight_dtype=torch.bfloat16)
    losses = []
    for i in range(2):
        b = bench.to_device(bench.host_batch(4, 1000 * rank + i), dev)
        losses.append(round(step(b)["loss"].item(), 6))
    torch.cuda.synchronize()
    if rank == 0:
        e = step._early_end
        torch.save({"wo": step.opt.arena[e:].detach().float().cpu(), "head": step.opt.arena[:e:997].detach().float().cpu(),
                    "loss": losses, "flag": step._wo_factor_exchange}, sys.argv[1])
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

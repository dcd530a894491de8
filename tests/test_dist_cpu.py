"""world_size-2 gloo tests (CPU) of the N>1 host logic: the flat-arena gradient exchange used by FlatAdamW,
per-rank data sharding, and the reference arm's rank gating."""
import json
import os
import subprocess
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "e4t-diffusion_b200")]
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from e4t_b200.engine import all_reduce_sum_, shard_seed
    g = torch.arange(1000, dtype=torch.float32) * (rank + 1)
    scale = all_reduce_sum_(g)
    ok = torch.allclose(g * scale, torch.arange(1000, dtype=torch.float32) * 1.5) and scale == 0.5
    # averaged-gradient AdamW == what DDP + torch.optim.AdamW would do on the mean gradient
    p = torch.ones(1000); ref = torch.ones(1000, requires_grad=True)
    opt = torch.optim.AdamW([ref], lr=1e-2)
    ref.grad = torch.arange(1000, dtype=torch.float32) * 1.5
    opt.step()
    q.put((rank, bool(ok), shard_seed(42, rank), float(ref.detach().sum())))
    dist.barrier()
    dist.destroy_process_group()


def test_flat_gradient_allreduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[1] for r in res] == [True, True]
    assert res[0][2] != res[1][2]                    # ranks draw different images
    assert res[0][3] == res[1][3]                    # identical replicas after the averaged step


def test_reference_arm_rank_gating():
    """Under torchrun only rank 0 runs/prints the CPU reference arm; other ranks exit 0 silently."""
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "0"], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_bench_refuses_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)

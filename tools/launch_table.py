"""Per-kernel table of ONE steady-state step from an `ncu --metrics gpu__time_duration.sum --csv` launch list of
`bench.py --profile-one-step` (the capture also holds model construction and the warm-up step: the step is cut out as
the launches between the last two optimiser kernels).   python tools/launch_table.py gpurun_out/launches.csv out.md"""
import collections
import csv
import re
import sys


def main():
    src, dst = sys.argv[1], sys.argv[2]
    with open(src) as f:
        lines = [l for l in f if not l.startswith("==")]
    r = csv.reader(lines)
    hdr = next(r)
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    rows = []
    for x in r:
        try:
            v = float(x[iv].replace(",", ""))
        except (ValueError, IndexError):
            continue
        u = x[iu]
        v *= {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}.get(u, 1e-3)
        rows.append((x[ik], v))
    ad = [i for i, (k, _) in enumerate(rows) if "adamw" in k]
    # AdamW runs as consecutive launches at the end of a step: the step = after the previous group .. end of the last group
    last_end = ad[-1]
    prev_end = max(i for i in ad if i < last_end - 8)
    step = rows[prev_end + 1:last_end + 1]
    tot = sum(v for _, v in step)
    agg = collections.defaultdict(lambda: [0, 0.0])
    for k, v in step:
        k = re.sub(r"\(.*", "", k)
        agg[k][0] += 1
        agg[k][1] += v
    ours = sum(v for k, (n, v) in agg.items() if not k.startswith("void at::") and "cutlass" not in k and "nvjet" not in k
               and "cudnn" not in k and "nccl" not in k)
    with open(dst, "w") as o:
        o.write("# ncu launch list of one steady-state pre-training step (B=16), eager launches (the bench replays a CUDA graph)\n\n")
        o.write("command: `ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file <csv> python bench.py "
                "--profile-one-step --warmup 1 --no-cpu-baseline`; table by `tools/launch_table.py` (launches between the last two "
                "optimiser kernels).  Cold-cache, serialised durations: compare SHARES.\n\n")
        o.write(f"{len(step)} launches, sum {tot / 1e3:.2f} ms; e4t_b200 kernels {ours / 1e3:.2f} ms ({100 * ours / tot:.1f} %), "
                f"torch glue / library {(tot - ours) / 1e3:.2f} ms.\n\n| ms | share | launches | kernel |\n|---:|---:|---:|---|\n")
        for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            if v < 0.04e3 and n < 20:
                continue
            o.write(f"| {v / 1e3:.2f} | {100 * v / tot:.1f}% | {n} | `{k[:90]}` |\n")
    print(open(dst).read()[:3500])


if __name__ == "__main__":
    main()

"""Summarise one kernel of an .ncu-rep (ncu --set full) as the small JSON that bench.py reads for `roofline.traffic`.
    python tools/ncu_to_json.py gpurun_out/<report>.ncu-rep profiles/r02_attn_fwd_ncu.json "<source note>" [batch] ["shape"]
bench.py only uses the traffic figure when `batch` equals the batch it runs at.
"""
import csv
import json
import subprocess
import sys


def main():
    rep, out, note = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ""
    batch = int(sys.argv[4]) if len(sys.argv) > 4 else None
    shape = sys.argv[5] if len(sys.argv) > 5 else None
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    h, units, v = rows[0], rows[1], rows[2]
    col = {k: i for i, k in enumerate(h)}

    def val(k, scale_to=None):
        x = float(v[col[k]].replace(",", ""))
        u = units[col[k]]
        if scale_to == "bytes":
            x *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        if scale_to == "ms":
            x *= {"ns": 1e-6, "us": 1e-3, "ms": 1, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1, "s": 1e3, "second": 1e3}[u]
        return x

    d = {"kernel": v[col["Kernel Name"]][:64], "batch": batch, "shape": shape,
         "dram_bytes_read": int(val("dram__bytes_read.sum", "bytes")),
         "dram_bytes_write": int(val("dram__bytes_write.sum", "bytes")),
         "duration_ms_under_ncu": val("gpu__time_duration.sum", "ms"),
         "tensor_pipe_pct": val("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
         "xu_pipe_pct": val("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
         "issue_active_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active"),
         "registers": int(val("launch__registers_per_thread")),
         "grid": v[col["Grid Size"]] if "Grid Size" in col else None,
         "source": note}
    json.dump(d, open(out, "w"), indent=1)
    print(json.dumps(d, indent=1))


if __name__ == "__main__":
    main()

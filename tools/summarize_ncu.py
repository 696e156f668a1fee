"""Summarise ncu captures into small tracked files under profiles/.

    python tools/summarize_ncu.py launches <csv> <out.md>
    python tools/summarize_ncu.py kernel <ncu-rep> <out.json> [kernel-substring]
"""
import csv
import json
import subprocess
import sys
from collections import OrderedDict


def launches(path, out):
    rows = [r for r in csv.reader(open(path, errors="ignore")) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = OrderedDict()
    for r in rows:
        if r is hdr or len(r) <= vi:
            continue
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        name = r[ki].split("(")[0]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    total = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write("| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
        for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{name[:90]}` | {n} | {ns / 1e3:.1f} | {100 * ns / total:.1f} % |\n")
        f.write(f"\ntotal {total / 1e3:.1f} us over {sum(a[0] for a in agg.values())} launches "
                "(ncu per-launch times: cold cache, serialised -- compare shares)\n")


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "sm__cycles_elapsed.max"]


def kernel(rep, out, sub=None):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        if sub and sub not in name:
            continue
        d = {"kernel": name[:120]}
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                try:
                    d[w] = float(r[i].replace(",", ""))
                except ValueError:
                    d[w] = r[i]
                d[w + "__unit"] = units[i]
        def to_bytes(k):
            v, u = d.get(k), d.get(k + "__unit", "")
            if v is None:
                return None
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        rd, wr = to_bytes("dram__bytes_read.sum"), to_bytes("dram__bytes_write.sum")
        if rd is not None and wr is not None:
            d["dram_bytes_per_launch"] = rd + wr
        res.append(d)
    json.dump(res[0] if len(res) == 1 else res, open(out, "w"), indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        kernel(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)

"""Turn ncu output into the markdown summaries kept under profiles/.

    python tools/ncu_summary.py rep gpurun_out/kernels_r1.ncu-rep > profiles/r1_ncu_kernels.md
    python tools/ncu_summary.py launches gpurun_out/launches_r1b.csv > profiles/r1_launch_breakdown.md
"""
import collections
import csv
import io
import re
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "time us", 1e-3),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %", 1),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %", 1),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %", 1),
    ("lts__t_sector_hit_rate.pct", "L2 hit %", 1),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts", 1),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %", 1),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %", 1),
    ("launch__registers_per_thread", "regs", 1),
    ("launch__shared_mem_per_block_dynamic", "dyn smem B", 1),
    ("launch__grid_size", "grid", 1),
    ("sass__inst_executed_local_loads", "local loads", 1),
]


def short(name):
    name = re.sub(r"b200::<unnamed>::|b200::\(anonymous namespace\)::|void ", "", name)
    return re.sub(r"\(CUtensorMap_st.*", "", name)[:70]


def rep(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    units = rows[1] if len(rows) > 1 else [""] * len(hdr)
    idx = {h: i for i, h in enumerate(hdr)}
    to_us = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}
    print("# ncu --set full captures (clock-control none)\n")
    print("| kernel | " + " | ".join(m[1] for m in METRICS) + " |")
    print("|---|" + "---|" * len(METRICS))
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        cells = []
        for key, _, scale in METRICS:
            i = idx.get(key)
            v = r[i] if i is not None else ""
            try:
                f = float(v.replace(",", ""))
                f = f * to_us.get(units[i], 1.0) if key == "gpu__time_duration.sum" else f * scale
                cells.append(f"{f:.1f}" if abs(f) < 1e6 else f"{f:.3g}")
            except ValueError:
                cells.append(v)
        print(f"| `{short(r[idx['Kernel Name']])}` | " + " | ".join(cells) + " |")


def launches(path):
    tot, cnt = collections.Counter(), collections.Counter()
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(io.StringIO("".join(lines)))
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = short(r["Kernel Name"])
        name = re.sub(r"<.*", "", name)
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(unit, 1e-3)
        tot[name] += v
        cnt[name] += 1
    total = sum(tot.values())
    print("# Kernel time breakdown (ncu launch list, gpu__time_duration.sum)\n")
    print(f"total kernel time {total / 1e3:.1f} ms over {sum(cnt.values())} launches\n")
    print("| kernel | launches | total ms | share | avg us |")
    print("|---|---|---|---|---|")
    for k, v in tot.most_common(40):
        print(f"| `{k}` | {cnt[k]} | {v / 1e3:.2f} | {100 * v / total:.1f}% | {v / cnt[k]:.1f} |")


if __name__ == "__main__":
    {"rep": rep, "launches": launches}[sys.argv[1]](sys.argv[2])

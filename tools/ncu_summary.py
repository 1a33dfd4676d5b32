"""Turns ncu output into the small text summaries committed under profiles/ (developer tool).

    ncu_summary.py launches <launches.csv> [out.md]     per-kernel totals and shares of a launch list
    ncu_summary.py full <file.ncu-rep> [out.md]         key counters of every launch in a --set full report
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEY = [
    ("gpu__time_duration.sum", "time"),
    ("launch__grid_size", "grid"),
    ("launch__registers_per_thread", "regs"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm throughput %"),
    ("sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", "DMMA pipe % (active)"),
    ("TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor pipe % (elapsed, realtime)"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 ALU pipe % (active)"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU pipe % (active)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "smem ld bank conflicts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ldgsts.sum", "smem ldgsts bank conflicts"),
    ("smsp__inst_executed.sum", "warp instructions"),
]


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"^void ", "", name)
    return name[:60]


def launches(path, out):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 14 and r[0].isdigit()]
    tot = collections.OrderedDict()
    for r in rows:
        k = short(r[4])
        val = float(r[14].replace(",", ""))
        unit = r[13]
        us = val / 1e3 if unit in ("ns", "nsecond") else (val if unit in ("us", "usecond") else val * 1e3)
        t = tot.setdefault(k, [0, 0.0])
        t[0] += 1
        t[1] += us
    total = sum(t[1] for t in tot.values())
    print(f"# launch list: {path}\n", file=out)
    print(f"{len(rows)} launches, {total / 1e3:.1f} ms of device time (cold-cache, serialised: compare SHARES)\n", file=out)
    print("| kernel | launches | total ms | share |", file=out)
    print("|---|---:|---:|---:|", file=out)
    for k, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {n} | {us / 1e3:.3f} | {100 * us / total:.1f} % |", file=out)


def full(path, out):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    print(f"# ncu --set full: {path}\n", file=out)
    print("| metric | unit | " + " | ".join(f"{short(r[ki])} #{i}" for i, r in enumerate(data)) + " |", file=out)
    print("|---|---|" + "---:|" * len(data), file=out)
    for m, label in KEY:
        if m not in hdr:
            continue
        i = hdr.index(m)
        print(f"| {label} (`{m.split('.')[-3] if m.count('.') > 2 else m}`) | {units[i]} | " + " | ".join(r[i] for r in data) + " |", file=out)


if __name__ == "__main__":
    mode, path = sys.argv[1], sys.argv[2]
    out = open(sys.argv[3], "w") if len(sys.argv) > 3 else sys.stdout
    (launches if mode == "launches" else full)(path, out)

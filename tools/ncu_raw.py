"""Prints selected metrics of an .ncu-rep (developer tool): ncu_raw.py file.ncu-rep [regex]"""
import csv, io, re, subprocess, sys
rep = sys.argv[1]
pat = re.compile(sys.argv[2] if len(sys.argv) > 2 else
                 r"gpu__time_duration.sum|sm__cycles_elapsed.max|smsp__inst_executed.sum$|warps_active.avg.pct|issue_stalled.*per_warp_active.pct$|registers_per_thread|dram__bytes_(read|write).sum$|pipe_fp64.*pct|inst_executed_pipe_fp64|sm__throughput.avg.pct|dram__throughput.avg.pct|issue_active.avg.pct|l1tex__data_bank_conflicts|shared_ld_bank|sm__pipe_fp64_cycles_active")
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, data = rows[0], rows[1], rows[2:]
ki = hdr.index("Kernel Name")
for r in data:
    print("==", r[ki][:90], r[hdr.index("Grid Size")], r[hdr.index("Block Size")])
for i, h in enumerate(hdr):
    if pat.search(h):
        print(f"{h:100s} {units[i]:12s}", [r[i] for r in data])

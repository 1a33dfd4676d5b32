"""Per-source-line sample totals of one kernel in an .ncu-rep (developer tool).
usage: ncu_lines.py rep kernel_regex [launch_skip]"""
import csv, io, subprocess, sys, collections
rep, kern = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      "regex:" + kern, "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file = None
tot = collections.Counter(); ins = collections.Counter(); src = {}
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path': cur_file = r[1].split('/')[-1]; continue
    if len(r) < 8 or r[0] in ('Line No', ''): continue
    try:
        ln = int(r[0]); s = int(r[6]); n = int(r[7])
    except ValueError:
        continue
    tot[(cur_file, ln)] += s; ins[(cur_file, ln)] += n; src[(cur_file, ln)] = r[1]
S = sum(tot.values())
print('total samples', S)
for (f, ln), s in tot.most_common(28):
    print(f"{f}:{ln:4d} {100*s/S:5.1f}% inst={ins[(f,ln)]:9d}  {src[(f,ln)][:100]}")

"""Aggregate an ncu source page (CSV, `--page source --csv --print-source cuda,sass` or sass only) of one kernel:
instructions executed and stall samples per CUDA source line and per opcode.  Usage: ncu_lines.py report.ncu-rep"""
import csv, io, subprocess, sys, collections, re
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {n: i for i, n in enumerate(hdr)}
ops = collections.Counter(); stall_ops = collections.Counter()
tot = 0; samp = 0
per_addr = []
for r in rows[hdr_i + 1:]:
    if len(r) < len(hdr): continue
    n = int(r[col["Instructions Executed"]] or 0); s = int(r[col["# Samples"]] or 0)
    src = r[col["Source"]].strip()
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
    op = m.group(2).split(".")[0] if m else "?"
    if op in ("LDS", "STS", "LDG", "STG", "F2F", "SHFL", "BAR", "DFMA", "DMUL", "DADD", "FFMA"):
        op = ".".join(m.group(2).split(".")[:3]) if op in ("F2F",) else op
    ops[op] += n; stall_ops[op] += s; tot += n; samp += s
    per_addr.append((n, s, src))
print(f"total warp instructions {tot:,}  samples {samp:,}")
print("opcode            inst      share   samples share")
for op, n in ops.most_common(top):
    print(f"{op:14s} {n:12,} {100*n/tot:6.1f}%  {stall_ops[op]:8,} {100*stall_ops[op]/max(samp,1):6.1f}%")

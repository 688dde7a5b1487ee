"""Dynamic warp instructions and stall samples of one kernel by SASS address region (bins of N static instructions),
with the dominant opcodes of each bin -- shows which phase of a long fused kernel the issue slots go to.
Usage: ncu_regions.py report.ncu-rep [bin]"""
import csv, io, subprocess, sys, collections, re
rep = sys.argv[1]; binsz = int(sys.argv[2]) if len(sys.argv) > 2 else 200
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]; col = {n: i for i, n in enumerate(hdr)}
ins = []
for r in rows[hdr_i + 1:]:
    if len(r) < len(hdr): continue
    src = r[col["Source"]].strip()
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
    op = m.group(2).split(".")[0] if m else "?"
    ins.append((op, int(r[col["Instructions Executed"]] or 0), int(r[col["# Samples"]] or 0)))
tot = sum(i[1] for i in ins); smp = sum(i[2] for i in ins)
print(f"static {len(ins)}  dynamic warp instr {tot:,}  samples {smp:,}")
for b in range(0, len(ins), binsz):
    chunk = ins[b:b + binsz]
    n = sum(c[1] for c in chunk); s = sum(c[2] for c in chunk)
    if n * 200 < tot and s * 200 < smp: continue
    ops = collections.Counter()
    for op, k, _ in chunk: ops[op] += k
    top = " ".join(f"{o}:{100*k/max(n,1):.0f}%" for o, k in ops.most_common(6))
    print(f"[{b:5d}-{b+len(chunk)-1:5d}] inst {100*n/tot:5.1f}%  smp {100*s/max(smp,1):5.1f}%   {top}")

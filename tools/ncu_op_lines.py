"""Which CUDA source lines execute a given SASS opcode most (ncu cuda,sass source page). Usage: ncu_op_lines.py rep OPCODE [top]"""
import csv, io, subprocess, sys, collections, re
rep, opc = sys.argv[1], sys.argv[2]; top = int(sys.argv[3]) if len(sys.argv) > 3 else 15
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
fname = "?"; cur = None; cnt = collections.Counter(); variants = collections.Counter()
for r in csv.reader(io.StringIO(out)):
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0].isdigit() and len(r) > 8: cur = (fname, int(r[0]), r[1].strip()[:80]); continue
    if r[0] == "" and len(r) > 8 and r[2].startswith("0x"):
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[3])
        if m and m.group(2).split(".")[0] == opc:
            n = int(r[7] or 0); cnt[cur] += n; variants[m.group(2)] += n
print(dict(variants.most_common(8)))
tot = sum(cnt.values())
for (f, n, src), c in cnt.most_common(top): print(f"{100*c/max(tot,1):5.1f}%  {c:12,}  {f}:{n}  {src}")

"""Warp instructions and stall samples per phase of the fused kernel (source-line buckets). Usage: ncu_phase.py rep"""
import csv, io, subprocess, sys, collections, re
rep = sys.argv[1]
src = open("wav2vec-heart-sounds_b200/csrc/fused_kernel.cuh").read().split("\n")
def line_of(marker):
    return next(i + 1 for i, l in enumerate(src) if marker in l)
marks = [("prologue", 1), ("resample", line_of("1. resample my slice")), ("despike", line_of("2. Schmidt despike")),
         ("tables", line_of("recipe tables: global")), ("pass1", line_of("3. low-pass + high-pass as one 4-state scan")),
         ("scan+carry", line_of("// 4: pass 1 done")), ("pass2", line_of("pass 2; statistics of the valid outputs")),
         ("stats", line_of("4. row statistics across the cluster")), ("store", line_of("5. normalise + write my share")), ("end", 10**9)]
kernel_start = line_of("fused_preprocess_kernel(const __grid_constant__")
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
fname = "?"; cur = None; inst = collections.Counter(); smp = collections.Counter()
# ncu lists SASS under the innermost inlined line; bucket helper-file lines by the last fused_kernel.cuh line seen in program order
rows = []
for r in csv.reader(io.StringIO(out)):
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0].isdigit() and len(r) > 8: cur = (fname, int(r[0])); continue
    if r[0] == "" and len(r) > 8 and r[2].startswith("0x"):
        m_ = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[3])
        rows.append((int(r[2], 16), cur, int(r[7] or 0), int(r[6] or 0), m_.group(2).split(".")[0] if m_ else "?"))
rows.sort()
seen = set(); rows = [r for r in rows if not (r[0] in seen or seen.add(r[0]))]
bucket = "prologue"
ops = collections.defaultdict(collections.Counter)
for addr, (f, ln), n, s, op in rows:
    if f == "fused_kernel.cuh" and ln >= kernel_start:
        bucket = [m for m, l0 in marks if l0 <= ln][-1]
    inst[bucket] += n; smp[bucket] += s; ops[bucket][op] += n
ti, ts = sum(inst.values()), sum(smp.values())
print(f"total {ti:,} warp instr, {ts:,} samples")
for m, _ in marks[:-1]:
    print(f"{m:12s} {inst[m]:12,} {100*inst[m]/ti:5.1f}% inst   {100*smp[m]/max(ts,1):5.1f}% samples")

if len(sys.argv) > 2:
    for m, _ in marks[:-1]:
        print(m, " ".join(f"{o}:{100*c/max(inst[m],1):.0f}%" for o, c in ops[m].most_common(12)))

"""Per CUDA source line: warp instructions executed and stall samples, from `ncu --page source --print-source cuda,sass`.
Usage: ncu_src.py report.ncu-rep [top]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
fname = "?"; lines = []
for r in csv.reader(io.StringIO(out)):
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] in ("Line No", "Function Name", "File Name"): continue
    if r[0].isdigit() and len(r) > 8:
        try: lines.append((fname, int(r[0]), r[1].strip(), int(r[6] or 0), int(r[7] or 0)))
        except ValueError: pass
tot_i = sum(l[4] for l in lines); tot_s = sum(l[3] for l in lines)
print(f"total inst {tot_i:,} samples {tot_s:,}")
print("---- by instructions")
for f, n, src, s, i in sorted(lines, key=lambda l: -l[4])[:top]:
    print(f"{100*i/tot_i:5.1f}% inst {100*s/max(tot_s,1):5.1f}% smp  {f}:{n}  {src[:90]}")
print("---- by samples")
for f, n, src, s, i in sorted(lines, key=lambda l: -l[3])[:top]:
    print(f"{100*s/max(tot_s,1):5.1f}% smp {100*i/tot_i:5.1f}% inst  {f}:{n}  {src[:90]}")

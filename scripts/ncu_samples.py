"""Warp-state samples of an ncu report by source line (where the warps' time goes, barrier waits included).
usage: python scripts/ncu_samples.py report.ncu-rep [top_n]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
f = lambda x: float(x.replace(",", "")) if x.replace(",", "").replace(".", "").isdigit() else 0.0
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
sec = hdr = None; data = []
for r in rows:
    if r and r[0] in ("File Name", "File Path"): sec = r[1].split("/")[-1]; hdr = None; continue
    if r and r[0] == "Line No": hdr = r; continue
    if hdr and r and r[0].isdigit():
        nm = len(hdr) - 4
        d = dict(zip(hdr[4:], r[len(r) - nm:]))
        data.append((f(d["# Samples"]), f(d["Instructions Executed"]), sec, int(r[0]), " ".join(r[1:len(r) - nm - 2]).strip()))
tot = sum(x[0] for x in data); toti = sum(x[1] for x in data)
print(f"total samples {tot:.0f}, instructions {toti:.0f}")
data.sort(key=lambda t: -t[0])
for s, ie, sec, ln, txt in data[:top]:
    print(f"{100 * s / tot:5.1f}% samples {100 * ie / toti:5.1f}% instr  {sec[:22]:22s}:{ln:<4d} {txt[:100]}")

"""Summarise an ncu report: headline counters and warp instructions / shared wavefronts per source line.
usage: python scripts/ncu_lines.py report.ncu-rep [units_per_launch] [top_n]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
units = float(sys.argv[2]) if len(sys.argv) > 2 else 65536.0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40


def f(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return 0.0


raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, unit, r = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread", "Kernel Name"]
for h, u, v in zip(hdr, unit, r):
    if h in want or ("issue_stalled" in h and "per_issue_active" in h and f(v) > 0.3):
        print(f"{h} [{u}] = {v}")
vals = dict(zip(hdr, r))
if len(sys.argv) > 5:
    # python scripts/ncu_lines.py report units top counters.json kernel_key  -> merge the headline counters into the JSON file
    import json
    import os
    path, key = sys.argv[4], sys.argv[5]
    doc = json.load(open(path)) if os.path.exists(path) else {}
    doc[key] = {
        "warp_instructions_per_situation": f(vals["smsp__inst_executed.sum"]) / units,
        "shared_wavefronts_per_situation": f(vals["l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]) / units,
        "shared_bank_conflict_wavefronts_per_situation": f(vals["l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]) / units,
        "issue_active_pct": f(vals["smsp__issue_active.avg.pct_of_peak_sustained_active"]),
        "alu_pipe_pct": f(vals["sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"]),
        "fma_pipe_pct": f(vals["sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"]),
        "lsu_pipe_pct": f(vals["sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"]),
        "shared_wavefront_pct": f(vals["l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"]),
        "gpu_time_ms_under_ncu": f(vals["gpu__time_duration.sum"]),
        "dram_bytes_read": f(vals["dram__bytes_read.sum"]) * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1}.get(
            dict(zip(hdr, unit))["dram__bytes_read.sum"], 1),
        "dram_bytes_write": f(vals["dram__bytes_write.sum"]) * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1}.get(
            dict(zip(hdr, unit))["dram__bytes_write.sum"], 1),
        "units_per_launch": units,
        "capture": os.path.basename(rep) + " (ncu --set full --clock-control none)",
    }
    json.dump(doc, open(path, "w"), indent=1)
    print("wrote", path, key)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True,
                     text=True).stdout
rows = list(csv.reader(src.splitlines()))
sec = None
hdr = None
data = {}
for r in rows:
    if r and r[0] in ("File Name", "File Path"):
        sec = r[1]
        hdr = None
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr and r and r[0].isdigit():
        nm = len(hdr) - 4
        d = dict(zip(hdr[4:], r[len(r) - nm:]))
        data.setdefault(sec, []).append((int(r[0]), " ".join(r[1:len(r) - nm - 2]), d))
allrows = []
for sec, lst in data.items():
    tot = sum(f(d["Instructions Executed"]) for _, _, d in lst)
    print(f"{sec.split('/')[-1]}: {tot / units:.0f} warp-instructions per unit")
    for ln, s, d in lst:
        allrows.append((f(d["Instructions Executed"]), sec.split("/")[-1], ln, s, d))
allrows.sort(key=lambda t: -t[0])
print("instr/unit  shared wavefronts/unit (ideal)  samples  line  source")
for ie, sec, ln, s, d in allrows[:top]:
    print(f"{ie / units:8.0f} {f(d['L1 Wavefronts Shared']) / units:7.0f} ({f(d['L1 Wavefronts Shared Ideal']) / units:5.0f}) "
          f"{d['# Samples']:>7} {sec[:18]:18s}:{ln:<4d} {s.strip()[:90]}")

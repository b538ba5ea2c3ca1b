"""Summarise an .ncu-rep (read on the CPU box) into profiles/<name>.md + the DRAM-traffic json.

usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_xxx [--traffic-json]
"""
import collections, csv, io, json, re, subprocess, sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
KEYS = [
    "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__cycles_elapsed.avg", "smsp__cycles_active.avg", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active",
    "smsp__warps_eligible.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]
lines = [f"# ncu summary: {rep.split('/')[-1]}", "",
         "Captured with `ncu --set full --clock-control none --import-source on` under gpurun (1 GPU), "
         "read on the CPU box with `ncu -i ... --page raw --csv`. One column per captured launch.", ""]
lines.append("| metric | unit | " + " | ".join(f"launch {i}" for i in range(len(data))) + " |")
lines.append("|---|---|" + "---|" * len(data))
for k in KEYS:
    if k in hdr:
        i = hdr.index(k)
        vals = [r[i][:90] for r in data]
        lines.append(f"| {k} | {units[i]} | " + " | ".join(vals) + " |")
stall = [(h, i) for i, h in enumerate(hdr) if "issue_stalled" in h and h.endswith("_per_issue_active.ratio")]
lines += ["", "Warp stall reasons (average warps stalled per issued instruction, launch 0):", ""]
for h, i in sorted(stall, key=lambda t: -float(data[0][t[1]] or 0)):
    name = h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")
    lines.append(f"- {name}: {float(data[0][i]):.3f}")
# dynamic SASS opcode mix
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(sass)))
starts = [i for i, r in enumerate(srows) if r and r[0] == "Address"]
if starts:
    sh = srows[starts[0]]
    body = srows[starts[0] + 1: (starts[1] - 1 if len(starts) > 1 else None)]
    ci, src = sh.index("Instructions Executed"), sh.index("Source")
    byop, tot = collections.Counter(), 0
    for r in body:
        if len(r) > ci and r[ci].isdigit():
            m = re.match(r"\s*(?:@!?U?P[0-9T]+\s+)?([A-Z0-9_]+)", r[src])
            byop[m.group(1) if m else "?"] += int(r[ci]); tot += int(r[ci])
    lines += ["", f"Dynamic SASS mix of launch 0 (warp-level instructions executed, total {tot}):", "",
              ", ".join(f"{op} {n}" for op, n in byop.most_common(40))]
    proof = {k: byop.get(k, 0) for k in ("UBLKCP", "SYNCS", "UTMALDG", "UTMASTG", "LDG", "STG")}
    lines += ["", f"TMA evidence in SASS (executed counts): {proof}  (UBLKCP = cp.async.bulk, SYNCS = mbarrier ops)"]
open(out + ".md", "w").write("\n".join(lines) + "\n")
if "--traffic-json" in sys.argv:
    i, j, k = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
    scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}
    rd = sum(float(r[i]) for r in data) / len(data) * scale[units[i]]
    wr = sum(float(r[j]) for r in data) / len(data) * scale[units[j]]
    json.dump({"dram_bytes_per_launch": rd + wr, "dram_read_bytes": rd, "dram_write_bytes": wr,
               "duration_us_cold": sum(float(r[k]) for r in data) / len(data), "launches_averaged": len(data),
               "kernel": data[0][hdr.index("Kernel Name")][:100], "source": out + ".md",
               "note": "L2 flushed before every launch; DRAM writes below the algorithmic 50.3 MB because the "
                       "L2 is write-back: part of the results is still in L2 when the kernel ends"},
              open("profiles/ncu_traffic.json", "w"), indent=1)
print("wrote", out + ".md")

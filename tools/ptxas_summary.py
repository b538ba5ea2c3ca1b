"""Summarise nvcc -Xptxas -v output: registers / spills per kernel."""
import re, subprocess, sys
log = open(sys.argv[1] if len(sys.argv) > 1 else 'silver2_isaacsim_b200/lib/ptxas.log').read()
pat = re.compile(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*Used (\d+) registers")
rows = []
for m in pat.finditer(log):
    name = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
    name = name.replace('h2o::', '').replace('(h2o::StepArgs)', '').replace('(StepArgs)', '')
    rows.append((int(m.group(5)), int(m.group(2)), int(m.group(3)), name))
for r in sorted(rows, key=lambda r: r[3]):
    print(f"{r[0]:4d} regs  stack {r[1]:4d}  spill {r[2]:4d}  {r[3][:120]}")

"""usage: python tools/dev/ptxas_summary.py <nvcc -Xptxas -v log> [filter]  -- registers / spills per kernel"""
import re, subprocess, sys
t = open(sys.argv[1]).read()
flt = sys.argv[2] if len(sys.argv) > 2 else "k_classify|k_emit"
for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\nptxas info\s*: Function properties for \S+\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\nptxas info\s*: Used (\d+) registers", t):
  d = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
  if re.search(flt, d):
    print(d.split("(zm::VolParams")[0][-60:], "stack", m.group(2), "spill", m.group(3), m.group(4), "regs", m.group(5))

"""Summarise an `ncu --set full` report (read here, without a GPU) into the small CSV kept under
profiles/: per kernel the launch shape, duration, DRAM bytes, unit utilisations and the top stall
reasons.   usage: python tools/dev/summarize_ncu.py <report.ncu-rep> <out.csv> "<header comment>"
Also prints {kernel: dram bytes per launch} as JSON on stdout (for profiles/traffic.json)."""
import csv, io, json, subprocess, sys

rep, out, note = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, units = rows[0], rows[1]
ik = h.index("Kernel Name")
KEEP = ["launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}
traffic = {}
with open(out, "w") as f:
  f.write(f"# {note}\n# source: ncu --set full --clock-control none --import-source on; report {rep.split('/')[-1]} (scratch, not committed)\n")
  for r in rows[2:]:
    f.write(f"\nKernel Name,{r[ik]},\n")
    for k in KEEP:
      if k in h:
        i = h.index(k)
        f.write(f"{k},{r[i]},{units[i]}\n")
    stalls = []
    for i, c in enumerate(h):
      if c.startswith("smsp__average_warps_issue_stalled_") and c.endswith("_per_issue_active.ratio") and "not_issued" not in c:
        try:
          stalls.append((float(r[i]), c))
        except ValueError:
          pass
    for v, c in sorted(stalls, reverse=True)[:8]:
      f.write(f"{c},{v:.2f},warps per issue\n")
    try:
      rd, wr = h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum")
      b = float(r[rd]) * SCALE.get(units[rd], 1.0) + float(r[wr]) * SCALE.get(units[wr], 1.0)
      name = "k_emit" if "k_emit" in r[ik] else ("k_classify_dense" if "1>" in r[ik].split("(")[0] else "k_classify")
      traffic[name] = int(b)
    except Exception:
      pass
print(json.dumps(traffic))

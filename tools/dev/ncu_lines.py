"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line."""
import csv, sys, collections
path, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30
rows = list(csv.reader(open(path)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[hi]
iaddr = hdr.index("Address"); iexec = hdr.index("Instructions Executed"); isamp = hdr.index("# Samples")
ithr = hdr.index("Thread Instructions Executed")
agg = collections.OrderedDict(); cur = None
for r in rows[hi + 1:]:
  if len(r) < len(hdr): continue
  if r[0]:  # a CUDA line header row
    cur = (r[0], r[1].strip()); agg.setdefault(cur, [0, 0, 0]); continue
  if cur is None or not r[iaddr]: continue
  try:
    agg[cur][0] += int(r[iexec]); agg[cur][1] += int(r[isamp]); agg[cur][2] += int(r[ithr])
  except ValueError: pass
tot_i = sum(v[0] for v in agg.values()); tot_s = sum(v[1] for v in agg.values())
print(f"total warp-instr {tot_i:,}  samples {tot_s:,}")
for (ln, src), v in sorted(agg.items(), key=lambda kv: -kv[1][int(sys.argv[3]) if len(sys.argv) > 3 else 1])[:top]:
  print(f"{ln:>5} inst {100*v[0]/max(tot_i,1):5.1f}%  samp {100*v[1]/max(tot_s,1):5.1f}%  thr/inst {v[2]/max(v[0],1):4.1f} | {src[:110]}")

"""Walk an `ncu --page source --csv --print-source sass` dump in address order and print the share of
executed warp instructions / stall samples between marker instructions (barriers, atomics, TMA...)."""
import csv, sys
path = sys.argv[1]; which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
marks = ("BAR.", "SYNCS", "MATCH", "UTMA", "EXIT", "ATOMG", "RED.", "ATOMS", "ATOM.", "BRA.U", "LDGSTS", "STG", "LDG")
if len(sys.argv) > 3: marks = tuple(sys.argv[3].split(","))
rows = list(csv.reader(open(path)))
secs = []; cur = None
for r in rows:
  if r and r[0] == "Kernel Name": cur = {"name": r[1], "rows": []}; secs.append(cur); continue
  if r and r[0] == "Address": cur["hdr"] = r; continue
  if cur is not None and "hdr" in cur and len(r) == len(cur["hdr"]): cur["rows"].append(r)
s = secs[which]; hdr = s["hdr"]; ie = hdr.index("Instructions Executed"); isamp = hdr.index("# Samples")
tot = sum(int(r[ie]) for r in s["rows"]); ts = sum(int(r[isamp]) for r in s["rows"])
print(s["name"][:100], "total warp-instr", tot, "samples", ts, "sass", len(s["rows"]))
cum = cs = last = lasts = 0
for i, r in enumerate(s["rows"]):
  cum += int(r[ie]); cs += int(r[isamp])
  t = r[1].strip()
  if any(k in t for k in marks):
    print(f"{i:5d} cum {100*cum/tot:5.1f}% (+{100*(cum-last)/tot:4.1f}%) samp {100*cs/ts:5.1f}% (+{100*(cs-lasts)/ts:4.1f}%) exec {int(r[ie]):>10}  {t[:80]}")
    last = cum; lasts = cs

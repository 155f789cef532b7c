"""Aggregate an `ncu --page source --csv --print-source sass` dump by SASS address order into
regions delimited by source-line ranges (inlined helpers are attributed to the surrounding region
by carrying the last non-helper region forward)."""
import csv, sys, collections
path = sys.argv[1]
# region spec: name:lo-hi,...
spec = [s.split(":") for s in sys.argv[2].split(",")]
regions = [(n, int(r.split("-")[0]), int(r.split("-")[1])) for n, r in spec]
rows = list(csv.reader(open(path)))
hi = next(i for i, r in enumerate(rows) if r and "Address" in r)
hdr = rows[hi]
iexec = hdr.index("Instructions Executed"); isamp = hdr.index("# Samples"); isrc = hdr.index("Source")
iline = None
for k in ("Line No", "File Line", "Line"):
  if k in hdr: iline = hdr.index(k)
print(hdr[:12])

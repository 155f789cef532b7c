#!/usr/bin/env python
"""Pure-Python emulation of the GPU formulation (edge-owner vertices + slot ids + perm lookups),
mirroring zmesh_b200/csrc/zm_kernels.cuh expression by expression (no tiles, no atomics).
Dev-time check of the geometric logic against the CPU oracle; not used by the product."""
import os, sys, gzip
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.oracle import OracleMesher, OracleMesh, assert_same_mesh, random_volume, voronoi_volume

import re
hdr = open(os.path.join(os.path.dirname(__file__), "..", "zmesh_b200", "csrc", "mc_tables.h")).read()
def tab(name):
  m = re.search(name + r"\[256\] = \{(.*?)\};", hdr, re.S)
  return [int(t.rstrip("ul"), 0) for t in re.findall(r"0x[0-9A-Fa-f]+(?:ull)?|\d+", m.group(1))]
TRI_NIBBLES, TRI_COUNT = tab("TRI_NIBBLES"), tab("TRI_COUNT")

corner_dx = lambda n: (0x66 >> n) & 1
corner_dy = lambda n: (0xF0 >> n) & 1
corner_dz = lambda n: (0xCC >> n) & 1
edge_a = lambda e: e if e < 8 else e - 8
edge_b = lambda e: ((e + 1) & 3) if e < 4 else (4 + ((e + 1) & 3) if e < 8 else e - 4)
def cdf(CO, n): return corner_dz(n) if CO else corner_dx(n)
def cdm(CO, n): return corner_dy(n)
def cds(CO, n): return corner_dx(n) if CO else corner_dz(n)
def edge_info(CO):
  out = []
  for e in range(12):
    a, b = edge_a(e), edge_b(e)
    mf, mm, ms = cdf(CO, a) + cdf(CO, b), cdm(CO, a) + cdm(CO, b), cds(CO, a) + cds(CO, b)
    axis = 0 if mf == 1 else (1 if mm == 1 else 2)
    out.append((mf >> 1, mm >> 1, ms >> 1, axis))
  return out

def emulate(vol, res, close):
  CO = bool(vol.flags.c_contiguous and not (vol.flags.f_contiguous and vol.ndim > 1 and False))
  if vol.flags.c_contiguous and vol.flags.f_contiguous: CO = True
  sx, sy, sz = vol.shape
  # memory-axis array A[s][m][f]
  A = vol.transpose(0, 1, 2) if CO else vol.transpose(2, 1, 0)   # C: (x,y,z)=(s,m,f); F: (z,y,x)=(s,m,f)
  A = np.ascontiguousarray(A).view(np.dtype(f"u{vol.dtype.itemsize}")).astype(np.uint64)
  pad = 1 if close else 0
  if pad: A = np.pad(A, 1)
  Es, Em, Ef = A.shape
  if min(Es, Em, Ef) < 2: return {}
  lab = lambda f, m, s: int(A[s, m, f])
  # slots
  slots = {}   # (f,m,s,slot) -> g
  verts = {}   # label -> list of keys ; perm[g] = rank
  perm = []
  for s in range(Es):
    for m in range(Em):
      for f in range(Ef):
        a = lab(f, m, s)
        for d, (df, dm, ds) in enumerate(((1, 0, 0), (0, 1, 0), (0, 0, 1))):
          if f + df >= Ef or m + dm >= Em or s + ds >= Es: continue
          b = lab(f + df, m + dm, s + ds)
          if a == b: continue
          for side, L in ((0, a), (1, b)):
            if L == 0: continue
            hf, hm, hs = 2 * f + (d == 0), 2 * m + (d == 1), 2 * s + (d == 2)
            kx, ky, kz = (hs, hm, hf) if CO else (hf, hm, hs)
            slots[(f, m, s, 2 * d + side)] = len(perm)
            perm.append(len(verts.setdefault(L, [])))
            verts[L].append((kx, ky, kz))
  EI = edge_info(CO)
  faces = {}
  for s in range(Es - 1):
    for m in range(Em - 1):
      for f in range(Ef - 1):
        cl = [lab(f + cdf(CO, n), m + cdm(CO, n), s + cds(CO, n)) for n in range(8)]
        if len(set(cl)) == 1: continue
        acc = 0
        while acc != 0xFF:
          start = ((~acc) & 0xFF & -((~acc) & 0xFF)).bit_length() - 1
          L = cl[start]
          msk = sum((1 << n) for n in range(8) if cl[n] == L)
          acc |= msk
          if L == 0: continue
          cs = ~msk & 0xFF
          nib = TRI_NIBBLES[cs]
          for t in range(TRI_COUNT[cs]):
            vi = []
            for k in range(3):
              e = (nib >> (12 * t + 4 * k)) & 0xF
              of, om, os_, d = EI[e]
              uf, um, us = f + of, m + om, s + os_
              slot = 2 * d + (0 if lab(uf, um, us) == L else 1)
              vi.append(perm[slots[(uf, um, us, slot)]])
            faces.setdefault(L, []).append((vi[1], vi[0], vi[2]))
  out = {}
  r = np.array(res, dtype=np.float32)
  for L, ks in verts.items():
    k = np.array(ks, dtype=np.float32)
    v = (k * r) / np.float32(2.0)
    out[L] = OracleMesh(v.astype(np.float32), np.array(faces[L], dtype=np.uint32))
  return out

def check(vol, res, close, name):
  got = emulate(vol, res, close)
  o = OracleMesher(res, "port"); o.mesh(vol, close=close)
  ids = sorted(o.ids())
  assert sorted(got.keys()) == ids, (name, len(got), len(ids))
  for L in ids:
    assert_same_mesh(got[L], o.get(L), what=f"{name}:{L}")
  print("ok", name, len(ids), "labels")

if __name__ == "__main__":
  for order in "CF":
    for close in (False, True):
      box = np.zeros((5, 7, 6), dtype=np.uint8, order=order); box[1:-1, 1:-1, 1:-1] = 1
      check(box, (4, 4, 40), close, f"box {order} {close}")
      check(random_volume((7, 8, 9), 6, np.uint32, seed=3, order=order), (1, 2, 3), close, f"rand {order} {close}")
      check(voronoi_volume((12, 11, 10), 5, np.uint64, order=order), (4, 4, 40), close, f"vor {order} {close}")
  vol = np.load(gzip.open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "connectomics.npy.gz")))
  check(np.asfortranarray(vol[200:216, 200:216, 200:216]), (4, 4, 40), False, "crop F")
  check(np.ascontiguousarray(vol[200:216, 200:216, 200:216]), (4, 4, 40), True, "crop C close")

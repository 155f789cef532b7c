#!/usr/bin/env python
"""Generate the committed golden fixtures under tests/golden/ (dev-time, needs /root/reference).

  connectomics.npy.gz            the reference's sample volume (data fixture, BASELINE config 1)
  unsimplified/<label>.ply.gz    a subset of the reference's own golden meshes
                                 (connectomics_npy_meshes/unsimplified, legacy orientation, res (32,32,40))
  digests_connectomics.json      per-label (nv, nf, sha256 of the canonical form) of Mesher.get()
                                 at res (4,4,40), produced by the UNMODIFIED reference C++
                                 (oracle/_ref/libzmesh_ref.so) for all 2523 labels
  ref_cases.npz                  full reference outputs (vertices, faces, normals) for a set of
                                 small seeded cases covering dtype x order x close x voxel_centered

The canonical form is oracle.oracle.canonical_digest.
"""
import gzip, json, os, shutil, sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.oracle import OracleMesher, canonical_digest, voronoi_volume, random_volume  # noqa: E402

REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")


def small_cases():
  """name -> (volume, res, close).  Deterministic; regenerated identically inside the tests."""
  cases = {}
  for dt in (np.uint8, np.uint16, np.uint32, np.uint64):
    for order in ("C", "F"):
      box = np.zeros((11, 17, 19), dtype=dt, order=order)
      box[1:-1, 1:-1, 1:-1] = 1
      for close in (False, True):
        cases[f"box_{np.dtype(dt).name}_{order}_{int(close)}"] = (box, (4, 4, 40), close)
  vol = np.load(gzip.open(os.path.join(REF, "connectomics.npy.gz")))
  crop = vol[200:232, 200:232, 200:232]
  for order in ("C", "F"):
    for close in (False, True):
      cases[f"crop32_{order}_{int(close)}"] = (np.asarray(crop, order=order), (4, 4, 40), close)
  cases["crop_odd_F_0"] = (np.asfortranarray(vol[300:337, 100:129, 50:71]), (0.1, 3.3, 7.77), False)
  cases["crop_odd_C_1"] = (np.ascontiguousarray(vol[300:337, 100:129, 50:71]), (0.1, 3.3, 7.77), True)
  cases["random14_C_0"] = (random_volume((14, 14, 14), 24, np.uint32, seed=0, order="C"), (4, 4, 40), False)
  cases["random11_F_1"] = (random_volume((11, 12, 13), 50, np.uint16, seed=1, order="F"), (1, 1, 1), True)
  cases["random12_u8_C_1"] = (random_volume((12, 12, 12), 5, np.uint8, seed=2, order="C"), (2, 3, 5), True)
  cases["voronoi28_u64_F_1"] = (voronoi_volume((28, 28, 28), 12, np.uint64, seed=0, order="F"), (4, 4, 40), True)
  cases["voronoi26_u64_C_0"] = (voronoi_volume((26, 23, 20), 10, np.uint64, seed=0, order="C"), (4, 4, 40), False)
  thin = np.zeros((2, 2, 2), dtype=np.uint32); thin[0, 0, 0] = 7
  cases["two_cubed"] = (thin, (1, 1, 1), False)
  cases["flat_1x8x8"] = (np.ones((1, 8, 8), dtype=np.uint32), (1, 1, 1), False)
  cases["flat_1x8x8_close"] = (np.ones((1, 8, 8), dtype=np.uint32), (1, 1, 1), True)
  neg = np.full((5, 6, 7), -1, dtype=np.int32); neg[2:4, 2:4, 2:5] = 3
  cases["int32_neg"] = (neg, (1, 2, 3), False)
  return cases


def main():
  os.makedirs(os.path.join(OUT, "unsimplified"), exist_ok=True)
  shutil.copyfile(os.path.join(REF, "connectomics.npy.gz"), os.path.join(OUT, "connectomics.npy.gz"))
  vol = np.load(gzip.open(os.path.join(REF, "connectomics.npy.gz")))

  # 1. the reference's own golden PLYs: 40 smallest + 8 around the median
  gd = os.path.join(REF, "connectomics_npy_meshes", "unsimplified")
  files = sorted((os.path.getsize(os.path.join(gd, f)), f) for f in os.listdir(gd) if f.endswith(".ply.gz"))
  pick = files[:40] + files[len(files) // 2: len(files) // 2 + 8]
  for _, f in pick:
    shutil.copyfile(os.path.join(gd, f), os.path.join(OUT, "unsimplified", f))
  print("copied", len(pick), "golden PLYs,", sum(s for s, _ in pick), "bytes")

  # 2. digests of get() for all labels from the reference itself
  m = OracleMesher((4, 4, 40), kind="reference")
  m.mesh(vol, close=False)
  dig = {}
  for lbl in sorted(m.ids()):
    g = m.get(lbl)
    dig[str(lbl)] = [int(len(g.vertices)), int(len(g.faces)), canonical_digest(g.vertices, g.faces)]
  with open(os.path.join(OUT, "digests_connectomics.json"), "w") as f:
    json.dump({"res": [4, 4, 40], "close": False, "voxel_centered": False, "labels": dig}, f)
  print("digests:", len(dig), "labels", sum(v[0] for v in dig.values()), "vertices", sum(v[1] for v in dig.values()), "faces")

  # 3. small seeded cases, full outputs
  blob = {}
  for name, (v, res, close) in small_cases().items():
    m = OracleMesher(res, kind="reference")
    m.mesh(v, close=close)
    ids = sorted(m.ids())
    blob[f"{name}/ids"] = np.array(ids, dtype=np.uint64)
    for lbl in ids:
      g0 = m.get(lbl, normals=True, voxel_centered=False)
      g1 = m.get(lbl, normals=False, voxel_centered=True)
      assert np.array_equal(g0.faces, g1.faces)
      blob[f"{name}/{lbl}/v0"] = g0.vertices
      blob[f"{name}/{lbl}/v1"] = g1.vertices
      blob[f"{name}/{lbl}/f"] = g0.faces
      blob[f"{name}/{lbl}/n"] = g0.normals.astype(np.float32)
      if name.startswith(("box_uint32", "crop_odd", "int32")):
        g = m.get_mesh(lbl, normals=False)
        blob[f"{name}/{lbl}/lv"] = g.vertices
        blob[f"{name}/{lbl}/lf"] = g.faces
  np.savez_compressed(os.path.join(OUT, "ref_cases.npz"), **blob)
  print("ref_cases.npz:", len(blob), "arrays", os.path.getsize(os.path.join(OUT, "ref_cases.npz")), "bytes")


if __name__ == "__main__":
  main()

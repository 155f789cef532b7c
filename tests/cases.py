"""Seeded test volumes shared by the CPU and GPU tests.  `small_cases()` must stay identical to
tools/make_golden.py:small_cases (the committed tests/golden/ref_cases.npz was generated from it
by the unmodified reference)."""
import numpy as np

from oracle.oracle import random_volume, voronoi_volume


def small_cases(connectomics):
  cases = {}
  for dt in (np.uint8, np.uint16, np.uint32, np.uint64):
    for order in ("C", "F"):
      box = np.zeros((11, 17, 19), dtype=dt, order=order)
      box[1:-1, 1:-1, 1:-1] = 1
      for close in (False, True):
        cases[f"box_{np.dtype(dt).name}_{order}_{int(close)}"] = (box, (4, 4, 40), close)
  vol = connectomics
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


def check_against_ref_cases(make_mesher, ref_cases, connectomics, assert_same_mesh, MeshT, names=None):
  """Run every small case through `make_mesher(res)` (reference-shaped API) and compare with the
  outputs the unmodified reference produced (tests/golden/ref_cases.npz)."""
  cases = small_cases(connectomics)
  n_checked = 0
  for name, (vol, res, close) in cases.items():
    if names is not None and name not in names:
      continue
    m = make_mesher(res)
    m.mesh(vol, close=close)
    want_ids = [int(x) for x in ref_cases[f"{name}/ids"]]
    assert sorted(m.ids()) == want_ids, name
    for lbl in want_ids:
      g0 = m.get(lbl, normals=True, voxel_centered=False)
      w0 = MeshT(ref_cases[f"{name}/{lbl}/v0"], ref_cases[f"{name}/{lbl}/f"], ref_cases[f"{name}/{lbl}/n"])
      assert_same_mesh(g0, w0, what=f"{name}:{lbl}:vc0")
      g1 = m.get(lbl, normals=False, voxel_centered=True)
      w1 = MeshT(ref_cases[f"{name}/{lbl}/v1"], ref_cases[f"{name}/{lbl}/f"], None)
      assert_same_mesh(g1, w1, what=f"{name}:{lbl}:vc1")
      if f"{name}/{lbl}/lv" in ref_cases:
        gl = m.get_mesh(lbl, normals=False)
        wl = MeshT(ref_cases[f"{name}/{lbl}/lv"], ref_cases[f"{name}/{lbl}/lf"], None)
        assert_same_mesh(gl, wl, what=f"{name}:{lbl}:legacy")
      n_checked += 1
  return n_checked

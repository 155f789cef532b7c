"""CPU tests: the C-ABI library builds, loads and exports every symbol include/zmesh_b200.h
declares (no compute without a GPU), fails loudly without a device, and `Mesh` codecs work."""
import ctypes
import os
import re

import numpy as np
import pytest

from tests.conftest import ROOT
from zmesh_b200.mesh import Mesh


def _declared_symbols():
  text = open(os.path.join(ROOT, "include", "zmesh_b200.h")).read()
  text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
  return sorted(set(re.findall(r"\b(zm_[a-z_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(build_all):
  from zmesh_b200 import _lib
  lib = _lib.load()
  names = _declared_symbols()
  assert len(names) >= 18
  for n in names:
    assert hasattr(lib, n), f"libzmesh_b200.so does not export {n}"
    assert n in _lib.SYMBOLS, f"{n} is declared in the header but not bound in zmesh_b200/_lib.py"
  assert b"sm_100a" in lib.zm_version()


def test_library_has_sm100a_code(build_all):
  import subprocess
  from zmesh_b200 import _lib
  out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
  assert "sm_100a" in out


def test_no_cpu_fallback(build_all):
  """Without a usable device zm_create must fail (never compute on the CPU)."""
  import torch
  if torch.cuda.is_available():
    pytest.skip("a GPU is present")
  from zmesh_b200 import Mesher
  with pytest.raises(RuntimeError, match="no CPU fallback"):
    Mesher((1, 1, 1))


def test_product_never_imports_oracle():
  pkg = os.path.join(ROOT, "zmesh_b200")
  for dirpath, _, files in os.walk(pkg):
    for f in files:
      if f.endswith((".py", ".cu", ".cuh", ".h")):
        assert "oracle" not in open(os.path.join(dirpath, f)).read().lower().replace("no cpu fallback", ""), f


def _box_mesh():
  v = np.array([[0, 0, 0], [2, 0, 0], [0, 2, 0], [0, 0, 40], [2, 2, 40]], dtype=np.float32)
  f = np.array([[0, 1, 2], [0, 2, 3], [1, 4, 2]], dtype=np.uint32)
  return Mesh(v, f, None, id=7)


def test_mesh_container():
  m = _box_mesh()
  assert len(m) == 5 and not m.empty() and m.segid == 7
  assert m.vertices.dtype == np.float32 and m.faces.dtype == np.uint32
  e = Mesh()
  assert e.empty() and e.vertices.shape == (0, 3) and e.faces.shape == (0, 3) and e.normals is None
  c = m.clone()
  assert c == m
  c.vertices[0, 0] = 1
  assert c != m
  cat = Mesh.concatenate(m, m)
  assert len(cat) == 10 and cat.faces.max() == 9
  assert m.triangles().shape == (3, 3, 3)


def test_mesh_codecs_roundtrip():
  m = _box_mesh()
  assert Mesh.from_precomputed(m.to_precomputed()) == m
  assert Mesh.from_ply(m.to_ply()) == m
  assert Mesh.from_obj(m.to_obj()) == m
  with_normals = m.clone()
  with_normals.normals = np.ones((5, 3), dtype=np.float32)
  assert Mesh.from_precomputed(with_normals.to_precomputed()) != with_normals  # normals are not stored
  with pytest.raises(ValueError):
    Mesh.from_precomputed(m.to_precomputed()[:20])


def _reference_volume(data):
  """What the reference hands its C++ mesher for a contiguous array: `reshape(data, (data.size,))`
  (zmesh/_zmesh.pyx:698-729, Fortran order tested first) and the extents data.shape[:3] with
  c_order = data.flags.c_contiguous (:970-976) -- restated here with plain indexing."""
  flat = data.ravel(order="F" if data.flags.f_contiguous else "C")
  sx, sy, sz = data.shape[:3]
  order = "C" if data.flags.c_contiguous else "F"
  return flat[: sx * sy * sz].reshape((sx, sy, sz), order=order)


def test_as_volume3d_follows_the_reference():
  from zmesh_b200.mesher import as_volume3d
  rng = np.random.default_rng(0)
  base = rng.integers(0, 5, size=(6, 5, 4), dtype=np.uint16)
  for arr in (np.ascontiguousarray(base), np.asfortranarray(base)):
    out = as_volume3d(arr)
    assert out is arr  # used in place
    for four_d in (arr[..., None], arr[..., None, None]):
      out = as_volume3d(four_d)
      assert out.shape == (6, 5, 4) and np.shares_memory(out, arr) and np.array_equal(out, base)
      assert out.flags.c_contiguous == arr.flags.c_contiguous
      assert np.array_equal(as_volume3d(four_d, close=True), base)
  strided = np.ascontiguousarray(rng.integers(0, 5, size=(12, 5, 12), dtype=np.uint8))[::2, :, ::3]
  out = as_volume3d(strided)
  assert out.flags.c_contiguous and np.array_equal(out, strided)
  # extra axes with extent > 1 (automated_test.py:195-213, transpose=True): the first sx*sy*sz elements of the buffer
  wide = rng.integers(0, 5, size=(7, 6, 5, 3), dtype=np.uint32)
  for arr in (wide, np.asfortranarray(wide), wide.T, np.ascontiguousarray(wide.T)):
    out = as_volume3d(arr)
    assert out.shape == arr.shape[:3] and np.shares_memory(out, arr)
    assert np.array_equal(out, _reference_volume(arr))
    with pytest.raises(ValueError):
      as_volume3d(arr, close=True)
  with pytest.raises(IndexError):
    as_volume3d(np.zeros((4, 4), dtype=np.uint8))


def test_fanc_fixture_shape():
  """The reference's C-vs-F regression volume (fanc_bug.npy.gz): 4-d boolean, meshed as 3-d."""
  import gzip
  from tests.conftest import GOLDEN
  from zmesh_b200.mesher import as_volume3d
  with gzip.open(os.path.join(GOLDEN, "fanc_bug.npy.gz"), "rb") as f:
    v = np.load(f)
  assert v.shape == (512, 512, 128, 1) and v.dtype == np.bool_
  assert as_volume3d(v).shape == (512, 512, 128)
  assert as_volume3d(v.T).shape == (1, 128, 512)  # no cube along x: the reference finds no ids


def test_get_slices_the_staged_block():
  """Host logic of Mesher.get (no GPU, no library): per-label views of the one bulk transfer, a fresh read from
  the device when a label is asked for twice, float64 normals, empty meshes for missing / erased / face-less labels."""
  from zmesh_b200 import mesher as M
  m = M.Mesher.__new__(M.Mesher)
  m._voxel_res = np.array((4, 4, 40), dtype=np.float32)
  m._max_label, m._erased = 2 ** 32 - 1, set()
  st = M._Stage()
  st.key = (False, None)
  st.v = np.arange(30, dtype=np.float32).reshape(10, 3)
  st.f = np.arange(12, dtype=np.uint32).reshape(4, 3)
  st.n = np.ones((10, 3), dtype=np.float32)
  st.index, st.given = {5: (0, 6, 0, 3), 9: (6, 10, 3, 4), 11: (10, 10, 4, 4)}, set()
  m._stage = st
  pristine_v, pristine_f = st.v.copy(), st.f.copy()
  fetched = []

  def fake_fetch(label, normals, voxel_centered, transpose):  # stands in for the per-label device read (zm_get)
    fetched.append((label, normals, voxel_centered, transpose))
    r = st.index[label]
    out = Mesh(pristine_v[r[0]:r[1]].copy(), pristine_f[r[2]:r[3]].copy(), None)
    out.normals = np.ones((r[1] - r[0], 3), dtype=np.float64) if normals else None
    out.id = label
    return out
  m._fetch = fake_fetch
  a = m.get(5, normals=True)
  assert isinstance(a, Mesh) and a.id == 5 and len(a) == 6 and a.faces.shape == (3, 3)
  assert a.vertices.dtype == np.float32 and a.faces.dtype == np.uint32
  assert a.normals.dtype == np.float64 and a.normals.shape == (6, 3)
  assert np.shares_memory(a.vertices, st.v) and not fetched
  # the caller owns the first result: an in-place edit must not leak into a later get() of the same label,
  # which is read back from the device (mesh.vertices += offset corrupted every later get)
  a.vertices += 100.0
  b = m.get(5, normals=True)
  assert fetched == [(5, True, False, False)]
  assert np.array_equal(b.vertices, pristine_v[0:6]) and not np.shares_memory(a.vertices, b.vertices)
  a.vertices -= 100.0
  assert a == b
  c = m.get(9)
  assert c.normals is None and c.id == 9 and np.array_equal(c.vertices, st.v[6:10]) and np.array_equal(c.faces, st.f[3:4])
  for missing in (11, 12345):
    e = m.get(missing)
    assert e.empty() and e.id == missing and e.vertices.shape == (0, 3) and e.faces.shape == (0, 3)
  m._erased.add(9)
  assert m.get(9).empty()
  with pytest.raises(NotImplementedError):
    m.get(5, reduction_factor=2)
  with pytest.raises(OverflowError):
    m.get(2 ** 32)

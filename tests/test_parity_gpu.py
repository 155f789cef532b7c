"""GPU parity tests (run on the B200 box: pytest -m gpu).  Everything goes through the C ABI via
zmesh_b200.Mesher and is compared with the CPU oracle, the committed outputs of the unmodified
reference, and the reference's own golden meshes.  Bit-exact for vertices/faces (canonical sets,
SURVEY.md section 8c); normals within 1e-5 absolute (unit vectors) with equal NaN masks."""
import gzip
import json
import os

import numpy as np
import pytest

from oracle.oracle import (OracleMesh, OracleMesher, assert_same_mesh, canonical_digest, random_volume,
                           voronoi_volume)
from tests.cases import check_against_ref_cases
from tests.conftest import GOLDEN

pytestmark = pytest.mark.gpu

NORMALS_TOL = 1e-5


@pytest.fixture(scope="module")
def Mesher(build_all):
  import zmesh_b200
  return zmesh_b200.Mesher


def compare_all_labels(gpu, cpu, normals=True, vcs=(False, True), legacy=False):
  assert gpu.ids() == sorted(cpu.ids())
  for lbl in gpu.ids():
    for vc in vcs:
      assert_same_mesh(gpu.get(lbl, normals=normals, voxel_centered=vc),
                       cpu.get(lbl, normals=normals, voxel_centered=vc), NORMALS_TOL, what=f"label {lbl} vc={vc}")
    if legacy:
      assert_same_mesh(gpu.get_mesh(lbl, normals=normals), cpu.get_mesh(lbl, normals=normals), NORMALS_TOL,
                       what=f"legacy {lbl}")
  return len(gpu.ids())


def test_reference_fixture_cases(Mesher, ref_cases, connectomics):
  """Outputs of the unmodified reference (tests/golden/ref_cases.npz): dtype x order x close x
  voxel_centered, anisotropic non-integer resolution, int32 negative labels, degenerate shapes."""
  from zmesh_b200 import Mesh
  n = check_against_ref_cases(lambda res: Mesher(res), ref_cases, connectomics,
                              lambda g, w, what: assert_same_mesh(g, w, NORMALS_TOL, what), Mesh)
  assert n > 100


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16, np.uint32, np.uint64])
@pytest.mark.parametrize("close", [False, True])
@pytest.mark.parametrize("order", ["C", "F"])
def test_executes(Mesher, dtype, close, order):
  """The reference's smoke test (automated_test.py:43-61 and :22-40), run against the drop-in."""
  labels = np.zeros((11, 17, 19), dtype=dtype, order=order)
  labels[1:-1, 1:-1, 1:-1] = 1
  mesher = Mesher((4, 4, 40))
  mesher.mesh(labels, close=close)
  for getter in (mesher.get, mesher.get_mesh):
    mesh = getter(1, normals=False)
    assert len(mesh.vertices) > 0 and len(mesh.faces) > 0 and mesh.normals is None
    mesh = getter(1, normals=True)
    assert len(mesh.vertices) > 0 and len(mesh.faces) > 0 and len(mesh.normals) > 0
    assert mesh.normals.dtype == np.float64 and mesh.vertices.dtype == np.float32 and mesh.faces.dtype == np.uint32


def test_codecs_roundtrip_on_gpu_mesh(Mesher):
  """automated_test.py:127-170 (precomputed / obj / ply)."""
  from zmesh_b200 import Mesh
  labels = np.zeros((11, 17, 19), dtype=np.uint32)
  labels[1:-1, 1:-1, 1:-1] = 1
  mesher = Mesher((4, 4, 40))
  mesher.mesh(labels)
  mesh = mesher.get(1, normals=False)
  assert Mesh.from_precomputed(mesh.to_precomputed()) == mesh
  assert Mesh.from_obj(mesh.to_obj()) == mesh
  assert Mesh.from_ply(mesh.to_ply()) == mesh
  assert Mesh.from_precomputed(mesher.get(1, normals=True).to_precomputed()) != mesher.get(1, normals=True)


def _blobs(shape, dtype, order):
  """Background label 9 (non-zero: uniform tiles of a non-zero label) with a few boxes of other labels."""
  rng = np.random.default_rng(7)
  v = np.full(shape, 9, dtype=dtype)
  for k in range(6):
    lo = [int(rng.integers(0, s - 3)) for s in shape]
    hi = [min(s, l + int(rng.integers(2, 12))) for s, l in zip(shape, lo)]
    v[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]] = (0, 3, 4, 5, 2**31 + 5, 77)[k] if np.dtype(dtype).itemsize >= 4 else (0, 3, 4, 5, 200, 77)[k]
  return np.asarray(v, order=order)


SEEDED = [
  # name, volume factory, res, close
  ("crop128_F", lambda c: np.asfortranarray(c[100:228, 100:228, 100:228]), (4, 4, 40), False),
  ("crop_C_close_odd", lambda c: np.ascontiguousarray(c[37:140, 250:331, 400:467]), (0.1, 3.3, 7.77), True),
  ("random64_u32_C", lambda c: random_volume((64, 64, 64), 1000, np.uint32, 0, "C"), (4, 4, 40), False),
  ("random40_u16_F_close", lambda c: random_volume((40, 41, 43), 300, np.uint16, 1, "F"), (1, 1, 1), True),
  ("random33_u8", lambda c: random_volume((33, 34, 35), 4, np.uint8, 2, "C"), (2, 3, 5), True),
  ("voronoi_u64_F_close", lambda c: voronoi_volume((100, 90, 80), 24, np.uint64, 0, "F"), (4, 4, 40), True),
  ("voronoi_u64_C", lambda c: voronoi_volume((70, 65, 97), 16, np.uint64, 1, "C"), (4, 4, 40), False),
  ("thin_f", lambda c: random_volume((2, 50, 60), 5, np.uint32, 3, "F"), (1, 1, 1), False),
  ("thin_s", lambda c: random_volume((70, 40, 2), 5, np.uint32, 4, "F"), (1, 1, 1), True),
  ("tile_edges", lambda c: random_volume((33, 9, 9), 3, np.uint8, 5, "F"), (1, 1, 1), False),
  ("tile_exact", lambda c: random_volume((64, 16, 16), 7, np.uint16, 6, "F"), (1, 1, 1), False),
  # mostly uniform volumes with a few blobs: the whole-region uniform test, with TMA (row pitch a multiple of
  # 16 bytes) and with the plain-load staging path (70 * 2 bytes is not), with and without the `close` border
  ("blobs_tma_close", lambda c: _blobs((96, 40, 40), np.uint32, "F"), (4, 4, 40), True),
  ("blobs_plain_u16", lambda c: _blobs((70, 33, 41), np.uint16, "F"), (1, 2, 3), False),
  ("blobs_plain_u64_C_close", lambda c: _blobs((37, 50, 67), np.uint64, "C"), (1, 1, 1), True),
]


@pytest.mark.parametrize("case", SEEDED, ids=[c[0] for c in SEEDED])
def test_seeded_volumes_match_oracle(Mesher, connectomics, case):
  name, make, res, close = case
  vol = make(connectomics)
  gpu, cpu = Mesher(res), OracleMesher(res, "port")
  gpu.mesh(vol, close=close)
  cpu.mesh(vol, close=close)
  n = compare_all_labels(gpu, cpu, normals=True, legacy=(name in ("crop128_F", "random33_u8")))
  assert n > 0


def test_many_labels_table_growth(Mesher):
  """Every voxel its own label: > 2^16 labels forces the global label table to grow, and the
  CTA-local table overflows into the direct global path."""
  vol = (np.arange(48 * 48 * 40, dtype=np.uint32) + 1).reshape((48, 48, 40))
  rng = np.random.default_rng(0)
  vol = rng.permutation(vol.ravel()).reshape(vol.shape).astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15)
  gpu, cpu = Mesher((1, 1, 1)), OracleMesher((1, 1, 1), "port")
  gpu.mesh(vol)
  cpu.mesh(vol)
  assert gpu.stats()["attempts"] > 1
  ids = gpu.ids()
  assert ids == sorted(cpu.ids()) and len(ids) == vol.size
  for lbl in ids[::997]:
    assert_same_mesh(gpu.get(lbl, normals=True), cpu.get(lbl, normals=True), NORMALS_TOL, what=str(lbl))


def test_api_semantics(Mesher):
  vol = np.zeros((20, 20, 20), dtype=np.uint16)
  vol[2:8, 2:8, 2:8] = 300
  vol[10:15, 3:9, 4:12] = 7
  m = Mesher((4, 4, 40))
  m.mesh(vol)
  assert m.ids() == [7, 300]
  missing = m.get(12345)
  assert missing.vertices.shape == (0, 3) and missing.faces.shape == (0, 3) and missing.id == 12345
  with pytest.raises(OverflowError):
    m.get(1 << 20)  # does not fit uint16 (the reference raises the same)
  with pytest.raises(NotImplementedError):
    m.get(7, reduction_factor=10)
  assert m.erase(7) is True and m.erase(7) is False
  assert m.ids() == [300] and len(m.get(7).vertices) == 0
  # captured resolution vs current voxel_res (zmesh/_zmesh.pyx:494 vs :581)
  m.voxel_res = (1, 1, 1)
  a = m.get(300, voxel_centered=True)
  ref = OracleMesher((4, 4, 40), "port")
  ref.mesh(vol)
  ref.voxel_res = (1, 1, 1)
  assert_same_mesh(a, ref.get(300, voxel_centered=True), what="captured res")
  m.clear()
  assert m.ids() == [] and len(m.get(300).vertices) == 0
  # remeshing with the same handle
  m.voxel_res = (2, 2, 2)
  m.mesh(vol, close=True)
  assert m.ids() == [7, 300]
  with pytest.raises(IndexError):
    m.mesh(np.zeros((4, 4), dtype=np.uint8))
  # compute_normals on an arbitrary mesh
  g = m.get(300)
  cn = m.compute_normals(g.clone())
  want = ref.compute_normals(g.vertices, g.faces)
  assert np.allclose(cn.normals, want, atol=NORMALS_TOL, equal_nan=True)


def test_non_contiguous_and_4d_input(Mesher, connectomics):
  base = np.ascontiguousarray(connectomics[50:114, 60:100, 70:118])
  strided = base[::2, :, ::3]
  assert not strided.flags.c_contiguous and not strided.flags.f_contiguous
  gpu, cpu = Mesher((4, 4, 40)), OracleMesher((4, 4, 40), "port")
  gpu.mesh(strided)
  cpu.mesh(strided)
  compare_all_labels(gpu, cpu, normals=False, vcs=(False,))
  four_d = np.asfortranarray(base)[..., None]
  gpu.mesh(four_d)
  cpu.mesh(np.asfortranarray(base))
  compare_all_labels(gpu, cpu, normals=False, vcs=(False,))
  b = (base % 2).astype(bool)
  gpu.mesh(b)
  cpu.mesh(b)
  compare_all_labels(gpu, cpu, normals=False, vcs=(False,))


def test_device_input(Mesher):
  import torch
  vol = voronoi_volume((64, 48, 40), 16, np.uint64, seed=2, order="C")
  t = torch.from_numpy(vol.view(np.int64)).cuda()
  gpu, cpu = Mesher((4, 4, 40)), OracleMesher((4, 4, 40), "port")
  gpu.mesh(t, close=True)
  cpu.mesh(vol, close=True)
  compare_all_labels(gpu, cpu, normals=True, vcs=(True,))


def test_c_and_f_order_give_identical_sets(Mesher, connectomics):
  """automated_test.py:172-193 strengthened to canonical equality."""
  crop = connectomics[102:230, 31:159, 17:145]
  f, c = Mesher((1, 1, 1)), Mesher((1, 1, 1))
  f.mesh(np.asfortranarray(crop))
  c.mesh(np.ascontiguousarray(crop))
  assert f.ids() == c.ids()
  for lbl in f.ids():
    assert_same_mesh(f.get(lbl), c.get(lbl), what=str(lbl))
    assert_same_mesh(f.get_mesh(lbl), c.get_mesh(lbl), what=f"legacy {lbl}")


@pytest.mark.parametrize("transpose", [True, False])
def test_fanc_bug(Mesher, fanc, transpose):
  """automated_test.py:195-213 run against the drop-in (4-d boolean input; transposed it has extent 1 along x
  and the reference finds nothing), strengthened to canonical equality and checked against the oracle."""
  lab = fanc.T if transpose else fanc
  f, c = Mesher((1, 1, 1)), Mesher((1, 1, 1))
  f.mesh(np.asfortranarray(lab))
  c.mesh(np.ascontiguousarray(lab))
  assert c.ids() == f.ids() == ([] if transpose else [1])
  for label in c.ids():
    cm, fm = c.get(label, normals=False, reduction_factor=0), f.get(label, normals=False, reduction_factor=0)
    assert np.isclose(cm.vertices.mean(), fm.vertices.mean())
    assert_same_mesh(cm, fm, what="fanc C vs F")
    cpu = OracleMesher((1, 1, 1), "port")
    cpu.mesh(np.ascontiguousarray(lab[..., 0]))
    assert_same_mesh(cm, cpu.get(label), what="fanc vs oracle")


def test_degenerate_volumes_512(Mesher):
  """BASELINE config 2: all-zero and single-label 512^3 uint32."""
  m = Mesher((4, 4, 40))
  m.mesh(np.zeros((512, 512, 512), dtype=np.uint32))
  assert m.ids() == []
  ones = np.ones((512, 512, 512), dtype=np.uint32)
  m.mesh(ones)
  assert m.ids() == []
  m.mesh(ones, close=True)
  assert m.ids() == [1]
  g = m.get(1)
  assert (len(g.vertices), len(g.faces)) == (1572864, 3145724)  # SURVEY.md section 6 probe of the reference
  # closed surface: every undirected edge is used by exactly two faces, and V - E + F = 2
  e = np.sort(np.concatenate([g.faces[:, [0, 1]], g.faces[:, [1, 2]], g.faces[:, [2, 0]]]).astype(np.uint64), axis=1)
  ek = e[:, 0] << np.uint64(32) | e[:, 1]
  uniq, cnt = np.unique(ek, return_counts=True)
  assert (cnt == 2).all()
  assert len(g.vertices) - len(uniq) + len(g.faces) == 2
  # padded coordinates: the surface lies at keys 1 and 2*512+1 (half-voxel units)
  assert g.vertices.min() == 2.0 and g.vertices[:, 2].max() == 1025.0 * 20


def test_connectomics_full_against_reference_digests(Mesher, connectomics):
  """BASELINE config 1 at full size: all 2523 labels, canonical sha256 equal to what the
  unmodified reference produced (tests/golden/digests_connectomics.json), plus the census."""
  with open(os.path.join(GOLDEN, "digests_connectomics.json")) as f:
    golden = json.load(f)
  m = Mesher(tuple(golden["res"]))
  m.mesh(connectomics, close=False)
  ids = m.ids()
  assert [str(i) for i in ids] == sorted(golden["labels"].keys(), key=int)
  st = m.stats()
  assert (st["n_vertices"], st["n_faces"]) == (44969925, 89540172)
  bulk = m.finalize()
  v, f, _ = m.fetch_all()
  order = np.argsort(bulk["labels"])
  for i in order:
    lbl = int(bulk["labels"][i])
    nv, nf, dig = golden["labels"][str(lbl)]
    vv = v[bulk["voff"][i]:bulk["voff"][i + 1]]
    ff = f[bulk["foff"][i]:bulk["foff"][i + 1]]
    assert (len(vv), len(ff)) == (nv, nf), lbl
    assert canonical_digest(vv, ff) == dig, lbl
  # the per-label accessor returns the same arrays as the bulk path
  for lbl in ids[:5] + ids[-5:]:
    g = m.get(lbl)
    assert canonical_digest(g.vertices, g.faces) == golden["labels"][str(lbl)][2]


@pytest.mark.parametrize("order", ["C", "F"])
def test_golden_ply_meshes(Mesher, connectomics, order):
  """The reference's known-answer test (automated_test.py:215-230) on the committed subset of
  connectomics_npy_meshes/unsimplified, with the strong canonical form instead of a column sort."""
  from zmesh_b200 import Mesh
  vol = np.asarray(connectomics, order=order)
  m = Mesher((32, 32, 40))
  m.mesh(vol)
  files = sorted(f for f in os.listdir(os.path.join(GOLDEN, "unsimplified")) if f.endswith(".ply.gz"))
  for fn in files:
    lbl = int(fn.split(".")[0])
    with gzip.open(os.path.join(GOLDEN, "unsimplified", fn), "rb") as f:
      gold = Mesh.from_ply(f.read())
    got = m.get_mesh(lbl, normals=False, simplification_factor=0, max_simplification_error=40)
    assert np.all(np.sort(gold.vertices[gold.faces], axis=0) == np.sort(got.vertices[got.faces], axis=0))
    assert_same_mesh(got, OracleMesh(gold.vertices, gold.faces), what=f"golden {lbl}")


def test_vertex_census_identity_random256(Mesher):
  """Size-independent property at a size the oracle would take minutes for (config 3 shape,
  256^3): V_label = number of axis-adjacent voxel pairs with exactly one endpoint == label, and
  every face index is in range, every vertex referenced."""
  vol = random_volume((256, 256, 256), 1000, np.uint32, seed=0, order="C")
  m = Mesher((4, 4, 40))
  m.mesh(vol)
  want = np.zeros(1000, dtype=np.int64)
  for ax in range(3):
    a = np.moveaxis(vol, ax, 0)[:-1].ravel()
    b = np.moveaxis(vol, ax, 0)[1:].ravel()
    d = a != b
    want += np.bincount(a[d], minlength=1000) + np.bincount(b[d], minlength=1000)
  bulk = m.finalize()
  got = dict(zip(bulk["labels"].tolist(), np.diff(bulk["voff"]).tolist()))
  assert sorted(got) == list(range(1, 1000))
  for lbl in range(1, 1000):
    assert got[lbl] == want[lbl], lbl
  v, f, _ = m.fetch_all()
  for i in range(0, len(bulk["labels"]), 97):
    ff = f[bulk["foff"][i]:bulk["foff"][i + 1]]
    nv = int(bulk["voff"][i + 1] - bulk["voff"][i])
    assert ff.max() < nv and len(np.unique(ff)) == nv


@pytest.mark.parametrize("order", ["C", "F"])
@pytest.mark.parametrize("dtype", [np.uint64, np.uint32])
def test_device_generator_matches_numpy(Mesher, order, dtype):
  """zm_synth_voronoi (bench input) is bit-identical to the numpy generator, incl. sub-blocks."""
  import torch
  from zmesh_b200.synth import voronoi_device
  full = (70, 60, 50)
  origin, shape = (10, 0, 7), (45, 60, 30)
  want = voronoi_volume(shape, 16, dtype, seed=3, order=order, origin=origin, full_shape=full)
  t = voronoi_device(shape, 16, dtype, seed=3, order=order, origin=origin, full_shape=full)
  torch.cuda.synchronize()
  got = t.cpu().numpy().view(dtype)
  assert got.shape == want.shape and np.array_equal(got, want)
  gpu, cpu = Mesher((4, 4, 40)), OracleMesher((4, 4, 40), "port")
  gpu.mesh(t)  # a CUDA tensor is consumed in place (device pointer through the C ABI)
  cpu.mesh(want)
  assert gpu.ids() == sorted(cpu.ids())
  for lbl in gpu.ids()[:20]:
    assert_same_mesh(gpu.get(lbl), cpu.get(lbl), what=str(lbl))


def test_config4_full_size_properties(Mesher):
  """BASELINE config 4 at its full size (Voronoi 1024^3 uint64, pitch 64, close=True, normals, voxel_centered) --
  far beyond what the oracle finishes in seconds -- through size-independent properties, all labels:
  the per-label vertex census (V_label = axis-adjacent voxel pairs, the virtual zero border included, with
  exactly one endpoint == label; computed independently with torch on the device), index integrity of every
  label's faces, coordinate range under close + voxel_centered, unit normals."""
  import torch
  from zmesh_b200.synth import voronoi_device
  n, pitch, res = 1024, 64, (4.0, 4.0, 40.0)
  t = voronoi_device((n, n, n), pitch, np.uint64, seed=0, order="F")
  m = Mesher(res)
  m.mesh(t, close=True)
  ids = m.ids()
  assert len(ids) == (n // pitch) ** 3
  bulk = m.finalize(normals=True, voxel_centered=True)

  # ---- census on the device (labels as signed bit patterns; every Voronoi label is non-zero) ----
  tc = t.permute(2, 1, 0)  # the contiguous view (z, y, x); the census is symmetric in the axes
  assert tc.is_contiguous()
  keys = np.sort(np.array(ids, dtype=np.uint64).view(np.int64))
  dkeys = torch.from_numpy(keys).cuda()
  want = torch.zeros(len(keys), dtype=torch.int64, device="cuda")

  def count(x):
    want.add_(torch.bincount(torch.searchsorted(dkeys, x.reshape(-1)), minlength=len(keys)))

  def pairs(a, b):
    d = a != b
    count(a[d])
    count(b[d])

  step = 32
  for z0 in range(0, n, step):
    z1 = min(n, z0 + step)
    blk = tc[z0:z1]
    for ax in (1, 2):
      pairs(blk.narrow(ax, 0, n - 1), blk.narrow(ax, 1, n - 1))
      count(blk.select(ax, 0))      # faces on the closed border: the outside voxel is 0
      count(blk.select(ax, n - 1))
    hi = min(n, z1 + 1)
    pairs(tc[z0:hi - 1], tc[z0 + 1:hi])
  count(tc[0])
  count(tc[n - 1])
  want = want.cpu().numpy()
  del tc, t
  torch.cuda.empty_cache()

  labels = bulk["labels"].view(np.int64)
  nv = np.diff(bulk["voff"]).astype(np.int64)
  nf = np.diff(bulk["foff"]).astype(np.int64)
  pos = np.searchsorted(keys, labels)
  assert np.array_equal(keys[pos], labels)
  assert np.array_equal(nv, want[pos]), "per-label vertex census"
  assert int(nv.sum()) == bulk["n_vertices"] and int(nf.sum()) == bulk["n_faces"]
  assert (nf > 0).all()

  v, f, nrm = m.fetch_all(normals=True)
  lo = np.array(res, dtype=np.float32)          # key 1 under close + voxel_centered: (res * 1 + res) / 2
  hi = lo * np.float32(n + 1)                   # key 2n + 1
  assert np.isfinite(v).all() and (v.min(axis=0) == lo).all() and (v.max(axis=0) == hi).all()
  length = np.sqrt((nrm.astype(np.float64) ** 2).sum(axis=1))
  ok = ~np.isnan(length)
  assert ok.mean() > 0.999 and np.abs(length[ok] - 1.0).max() < 1e-5
  for i in range(0, len(labels), 41):
    ff = f[bulk["foff"][i]:bulk["foff"][i + 1]]
    assert int(ff.max()) == nv[i] - 1 and len(np.unique(ff)) == nv[i], int(labels[i])


def _reference_kind():
  from oracle import oracle as O
  return "reference" if O.have_reference() else "port"


def test_config3_full_size(Mesher):
  """BASELINE config 3 at its stated size (perf.py:68-77: default_rng(0).integers(0, 1000), 512^3 uint32, C order): every
  tile overflows the per-tile staging and goes through the dense kernel.  Checked by the per-label vertex census of
  the WHOLE volume (V_label = axis-adjacent voxel pairs with exactly one endpoint == label) and, against the compiled
  reference, by bit-exact meshes of 20 labels on the volume's first 24 planes meshed as a standalone volume
  (the CPU cannot hold the full triangle soup)."""
  vol = random_volume((512, 512, 512), 1000, np.uint32, seed=0, order="C")
  m = Mesher((4, 4, 40))
  m.mesh(vol)
  st = m.stats()
  assert st["n_dense_tiles"] > 0 and st["n_tiles"] == 64 * 64 * 16
  want = np.zeros(1000, dtype=np.int64)
  for ax in range(3):
    a = np.moveaxis(vol, ax, 0)[:-1]
    b = np.moveaxis(vol, ax, 0)[1:]
    d = a != b
    want += np.bincount(a[d], minlength=1000) + np.bincount(b[d], minlength=1000)
    del a, b, d
  bulk = m.finalize()
  got = dict(zip(bulk["labels"].tolist(), np.diff(bulk["voff"]).tolist()))
  assert sorted(got) == list(range(1, 1000))
  assert all(got[lbl] == want[lbl] for lbl in range(1, 1000)), "per-label vertex census"
  assert bulk["n_vertices"] == int(want[1:].sum())
  m.clear()
  # against the reference on a slab it can hold
  slab = np.ascontiguousarray(vol[:24])
  del vol
  cpu = OracleMesher((4, 4, 40), _reference_kind())
  cpu.mesh(slab)
  m.mesh(slab)
  ids = m.ids()
  assert ids == sorted(cpu.ids()) and len(ids) == 999
  for lbl in ids[::50]:
    assert_same_mesh(m.get(lbl, normals=True), cpu.get(lbl, normals=True), NORMALS_TOL, what=f"c3 slab label {lbl}")


def test_config5_slab_against_reference(Mesher):
  """BASELINE config 5 (Voronoi 2048^3 uint64, pitch 128): the first 32 planes of the volume, 2048 x 2048 x 32, meshed
  as a standalone volume by the GPU path and by the compiled reference; every label compared by its order-independent
  fingerprint (bit-exact vertex and face sets), every 16th one also in canonical form.  (bench.py repeats the check on
  128 planes beside its cpu_baseline; the full 68.7 GB volume is compared across GPU counts by nccl_parity.)"""
  from oracle.oracle import multiset_digest, voronoi_volume_c
  slab = voronoi_volume_c((2048, 2048, 32), 128, np.uint64, 0, "F", full_shape=(2048, 2048, 2048))
  m = Mesher((4, 4, 40))
  m.mesh(slab)
  cpu = OracleMesher((4, 4, 40), _reference_kind())
  cpu.mesh(slab)
  ids = m.ids()
  assert ids == sorted(cpu.ids()) and len(ids) > 200
  for i, lbl in enumerate(ids):
    g, w = m.get(lbl), cpu.get(lbl)
    assert multiset_digest(g.vertices, g.faces) == multiset_digest(w.vertices, w.faces), f"c5 slab label {lbl}"
    if i % 16 == 0:
      assert_same_mesh(g, w, what=f"c5 slab label {lbl}")


def test_config4_slab_against_reference(Mesher):
  """BASELINE config 4 (Voronoi 1024^3 uint64, pitch 64, close=True, normals, voxel_centered) on the first 24 planes as a
  standalone volume against the compiled reference: bit-exact vertex / face sets of every label, normals within 1e-5
  with EQUAL NaN masks (the full-size test above can only bound the NaN share)."""
  from oracle.oracle import voronoi_volume_c
  slab = voronoi_volume_c((1024, 1024, 24), 64, np.uint64, 0, "F", full_shape=(1024, 1024, 1024))
  m = Mesher((4, 4, 40))
  m.mesh(slab, close=True)
  cpu = OracleMesher((4, 4, 40), _reference_kind())
  cpu.mesh(slab, close=True)
  ids = m.ids()
  assert ids == sorted(cpu.ids()) and len(ids) >= 256
  for lbl in ids[::4]:
    assert_same_mesh(m.get(lbl, normals=True, voxel_centered=True), cpu.get(lbl, normals=True, voxel_centered=True),
                     NORMALS_TOL, what=f"c4 slab label {lbl}")


def test_precomputed_objects_from_the_device(Mesher, connectomics):
  """Mesher.precomputed() -- objects laid out by a device kernel, one transfer -- returns, for every id,
  exactly the bytes the reference's encoder (zmesh/mesh.py:257-269; oracle.to_precomputed_bytes is pinned to it by
  tests/golden/codec_golden.npz) produces for get(id), and they decode to the oracle's mesh."""
  from oracle.oracle import to_precomputed_bytes
  from zmesh_b200 import Mesh
  vol = np.asfortranarray(connectomics[200:328, 200:328, 200:264])
  for vc in (False, True):
    m = Mesher((4, 4, 40))
    m.mesh(vol, close=True)
    objs = m.precomputed(voxel_centered=vc)
    ids = m.ids()
    assert sorted(objs) == ids and len(ids) > 20
    cpu = OracleMesher((4, 4, 40), "port")
    cpu.mesh(vol, close=True)
    for lbl in ids:
      g = m.get(lbl, voxel_centered=vc)
      assert bytes(objs[lbl]) == to_precomputed_bytes(g.vertices, g.faces), lbl
    for lbl in ids[::7]:
      assert_same_mesh(Mesh.from_precomputed(bytes(objs[lbl])), cpu.get(lbl, voxel_centered=vc), what=f"precomputed {lbl}")
    assert m.erase(ids[0]) and ids[0] not in m.precomputed(voxel_centered=vc)


def test_results_are_owned_by_the_caller_and_empty_before_mesh(Mesher):
  """Round-1 advice: (1) an in-place edit of a result (mesh.vertices += offset) must not leak into a later get() of the
  same label -- the first result is a view of the staged block, a repeated request is read back from the device;
  (2) a mesher that has not meshed anything answers like an empty one (the reference's __init__ builds an empty
  Mesher6464, zmesh/_zmesh.pyx:442-444)."""
  fresh = Mesher((1, 1, 1))
  assert fresh.ids() == [] and fresh.get(5).empty() and fresh.get(5).id == 5 and fresh.erase(5) is False
  vol = np.zeros((12, 12, 12), dtype=np.uint32)
  vol[2:9, 3:8, 4:10] = 5
  m = Mesher((2, 3, 4))
  m.mesh(vol)
  a = m.get(5, normals=True)
  pristine = a.vertices.copy()
  a.vertices += 100.0
  a.faces[:] = 0
  b = m.get(5, normals=True)
  assert np.array_equal(b.vertices, pristine) and b.faces.max() == len(pristine) - 1
  assert not np.shares_memory(a.vertices, b.vertices)
  c = m.get(5, normals=True)
  assert_same_mesh(b, c, NORMALS_TOL, what="third request")

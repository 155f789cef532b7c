"""CPU tests pinning the oracle (oracle/zmesh_oracle.c) to the reference:
   (a) the reference's own golden PLY meshes, (b) outputs of the unmodified reference C++
   committed as fixtures, (c) the reference shim itself when oracle/_ref is present."""
import gzip
import json
import os

import numpy as np
import pytest

from oracle.oracle import (OracleMesh, OracleMesher, assert_same_mesh, canonical_digest, have_reference,
                           random_volume, voronoi_volume)
from tests.cases import check_against_ref_cases
from tests.conftest import GOLDEN
from zmesh_b200.mesh import Mesh


def test_port_matches_reference_fixtures(ref_cases, connectomics):
  n = check_against_ref_cases(lambda res: OracleMesher(res, "port"), ref_cases, connectomics,
                              assert_same_mesh, OracleMesh)
  assert n > 100


def test_port_matches_golden_ply_and_digests(connectomics):
  """The reference's known-answer test (automated_test.py:215-230): legacy get_mesh at
  res (32,32,40) equals connectomics_npy_meshes/unsimplified/<label>.ply.gz.  The port meshes a
  crop around each label (a label's mesh only depends on its 1-voxel neighbourhood) so the CPU
  suite stays fast; coordinates are shifted back by the crop origin."""
  files = sorted(f for f in os.listdir(os.path.join(GOLDEN, "unsimplified")) if f.endswith(".ply.gz"))
  assert len(files) >= 40
  with open(os.path.join(GOLDEN, "digests_connectomics.json")) as f:
    digests = json.load(f)["labels"]
  vol = connectomics
  for fn in files:
    lbl = int(fn.split(".")[0])
    with gzip.open(os.path.join(GOLDEN, "unsimplified", fn), "rb") as f:
      gold = Mesh.from_ply(f.read())
    idx = np.argwhere(vol == lbl)
    lo = np.maximum(idx.min(axis=0) - 1, 0)
    hi = np.minimum(idx.max(axis=0) + 2, vol.shape)
    crop = np.asfortranarray(vol[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]])
    m = OracleMesher((32, 32, 40), "port")
    m.mesh(crop)
    got = m.get_mesh(lbl)
    # legacy vertices are (res0*z, res1*y, res2*x)/2 -> shift by the crop origin in that frame
    got.vertices += (np.array([lo[2], lo[1], lo[0]], dtype=np.float32) * np.array([32, 32, 40], dtype=np.float32))
    assert_same_mesh(got, OracleMesh(gold.vertices, gold.faces), what=f"golden {lbl}")
    # and the get() orientation against the digest the unmodified reference produced
    m4 = OracleMesher((4, 4, 40), "port")
    m4.mesh(crop)
    g = m4.get(lbl)
    g.vertices += lo.astype(np.float32) * np.array([4, 4, 40], dtype=np.float32)
    nv, nf, dig = digests[str(lbl)]
    assert (len(g.vertices), len(g.faces)) == (nv, nf)
    assert canonical_digest(g.vertices, g.faces) == dig, lbl


def test_port_edge_cases():
  m = OracleMesher((1, 1, 1), "port")
  m.mesh(np.zeros((8, 8, 8), dtype=np.uint32))
  assert m.ids() == []
  m.mesh(np.full((8, 8, 8), 5, dtype=np.uint32))
  assert m.ids() == []
  m.mesh(np.full((3, 3, 3), 5, dtype=np.uint32), close=True)
  assert m.ids() == [5]
  g = m.get(5)
  assert g.vertices.min() == 0.5 and g.vertices.max() == 3.5  # +1 voxel offset under close
  assert m.erase(5) is True and m.erase(5) is False
  assert len(m.get(5).vertices) == 0
  m.mesh(np.ones((1, 5, 5), dtype=np.uint8))
  assert m.ids() == []


@pytest.mark.skipif(not have_reference(), reason="oracle/_ref/libzmesh_ref.so not built here")
@pytest.mark.parametrize("order", ["C", "F"])
@pytest.mark.parametrize("close", [False, True])
def test_port_matches_reference_shim(order, close, connectomics):
  vols = [
    (np.asarray(connectomics[100:164, 100:164, 100:164], order=order), (4, 4, 40)),
    (random_volume((20, 21, 22), 30, np.uint16, seed=5, order=order), (0.1, 3.3, 7.77)),
    (voronoi_volume((33, 30, 31), 11, np.uint64, seed=3, order=order), (4, 4, 40)),
  ]
  for vol, res in vols:
    a, b = OracleMesher(res, "port"), OracleMesher(res, "reference")
    a.mesh(vol, close=close)
    b.mesh(vol, close=close)
    assert sorted(a.ids()) == sorted(b.ids())
    for lbl in a.ids():
      for vc in (False, True):
        assert_same_mesh(a.get(lbl, normals=True, voxel_centered=vc), b.get(lbl, normals=True, voxel_centered=vc),
                         what=f"{lbl}")
      assert_same_mesh(a.get_mesh(lbl, normals=True), b.get_mesh(lbl, normals=True), what=f"legacy {lbl}")


@pytest.mark.parametrize("kind", ["port", "reference"])
def test_fanc_bug_c_and_f_order(kind, fanc):
  """automated_test.py:195-213 (the C-vs-F regression volume) on the oracle, strengthened to canonical
  equality of the two orders; the port and the compiled reference must agree on the result."""
  if kind == "reference" and not have_reference():
    pytest.skip("oracle/_ref/libzmesh_ref.so not built here")
  vol = fanc[..., 0]
  f, c = OracleMesher((1, 1, 1), kind), OracleMesher((1, 1, 1), kind)
  f.mesh(np.asfortranarray(vol))
  c.mesh(np.ascontiguousarray(vol))
  assert sorted(f.ids()) == sorted(c.ids()) == [1]
  a, b = f.get(1), c.get(1)
  assert_same_mesh(a, b, what="fanc C vs F")
  assert np.isclose(a.vertices.mean(), b.vertices.mean())
  if kind == "reference":
    p = OracleMesher((1, 1, 1), "port")
    p.mesh(np.ascontiguousarray(vol))
    assert_same_mesh(p.get(1), b, what="fanc port vs reference")


def test_codecs_match_reference_encoders():
  """Mesh wire formats: the drop-in's encoders and the oracle's Precomputed restatement reproduce, byte
  for byte, what the UNMODIFIED reference encoders wrote into tests/golden/codec_golden.npz (tools/make_codec_golden.py:
  zmesh/mesh.py:257-269 to_precomputed, :348-376 to_ply, :321-346 to_obj), and decode them back."""
  import os
  from oracle.oracle import to_precomputed_bytes
  from tests.conftest import GOLDEN
  from zmesh_b200.mesh import Mesh
  g = np.load(os.path.join(GOLDEN, "codec_golden.npz"))
  names = sorted({k.split("/")[0] for k in g.files} - {"messy"})
  assert names == ["empty", "halfvoxel", "mid", "tiny"]
  for name in names:
    v, f = g[f"{name}/v"], g[f"{name}/f"]
    m = Mesh(v, f, None)
    want = {k: g[f"{name}/{k}"].tobytes() for k in ("precomputed", "ply", "obj")}
    assert m.to_precomputed() == want["precomputed"] == to_precomputed_bytes(v, f), name
    assert bytes(m.to_ply()) == want["ply"], name
    assert m.to_obj() == want["obj"], name
    back = Mesh.from_precomputed(want["precomputed"])
    assert np.array_equal(back.vertices, v) and np.array_equal(back.faces, f)
    if len(v):
      back = Mesh.from_ply(want["ply"])
      assert np.array_equal(back.vertices, v) and np.array_equal(back.faces, f)


def test_c_generator_matches_numpy_generator():
  """oracle.voronoi_volume_c (the threaded C restatement bench.py's CPU arms build their samples with) is bit-identical
  to the numpy definition of SURVEY.md 8d for both label widths, both memory orders and sub-blocks of a larger volume."""
  from oracle.oracle import voronoi_volume, voronoi_volume_c
  for dt, order in ((np.uint64, "F"), (np.uint32, "C"), (np.uint16, "F"), (np.uint8, "C")):
    a = voronoi_volume((37, 29, 23), 9, dt, 3, order, origin=(5, 11, 2), full_shape=(80, 60, 40))
    b = voronoi_volume_c((37, 29, 23), 9, dt, 3, order, origin=(5, 11, 2), full_shape=(80, 60, 40), threads=3)
    assert a.dtype == b.dtype and a.flags.f_contiguous == b.flags.f_contiguous and np.array_equal(a, b), (dt, order)


def test_multiset_digest_is_order_independent_and_winding_sensitive():
  """The O(n) fingerprint bench.py compares large samples with: invariant under vertex permutation and rotation of a
  face's corners, changed by a flipped winding, a moved vertex or a missing face."""
  from oracle.oracle import multiset_digest
  vol = voronoi_volume((24, 20, 22), 8, np.uint64, 1, "F")
  m = OracleMesher((4, 4, 40), "port")
  m.mesh(vol)
  g = m.get(m.ids()[0])
  v, f = g.vertices, g.faces.astype(np.int64)
  d = multiset_digest(v, f)
  perm = np.random.default_rng(0).permutation(len(v))
  inv = np.argsort(perm)
  assert multiset_digest(v[perm], inv[f][:, [1, 2, 0]]) == d
  assert multiset_digest(v[perm], inv[f][:, [0, 2, 1]]) != d
  v2 = v.copy(); v2[0, 0] += 0.5
  assert multiset_digest(v2, f) != d and multiset_digest(v, f[1:]) != d


def test_mesh_utilities_match_reference_results():
  """Mesh.remove_unreferenced_vertices / remove_degenerate_faces / consolidate / merge_close_vertices give exactly what
  the unmodified reference class gave on a seeded messy mesh (tests/golden/codec_golden.npz), and save / load round-trip."""
  import os
  import tempfile
  from tests.conftest import GOLDEN
  from zmesh_b200.mesh import Mesh
  g = np.load(os.path.join(GOLDEN, "codec_golden.npz"))
  v, f, n = g["messy/v"], g["messy/f"], g["messy/n"]
  for op in ("remove_unreferenced_vertices", "remove_degenerate_faces", "consolidate"):
    r = getattr(Mesh(v, f, n), op)()
    assert np.array_equal(r.vertices, g[f"messy/{op}/v"]) and np.array_equal(r.faces, g[f"messy/{op}/f"]), op
    want_n = g[f"messy/{op}/n"]
    assert (r.normals is None and want_n.size == 0) or np.array_equal(r.normals, want_n), op
  r = Mesh(g["messy/v2"], f, None).merge_close_vertices(0.8)
  assert np.array_equal(r.vertices, g["messy/merge/v"]) and np.array_equal(r.faces, g["messy/merge/f"])
  assert Mesh().consolidate().empty()
  with pytest.raises(NotImplementedError):
    Mesh(v, f, None).dust(10)
  with tempfile.TemporaryDirectory() as d:
    m = Mesh(g["mid/v"], g["mid/f"], None)
    for name in ("a.ply", "a.obj"):
      m.save(os.path.join(d, name))
      back = Mesh.load(os.path.join(d, name))
      assert np.allclose(back.vertices, m.vertices, atol=1e-4) and np.array_equal(back.faces, m.faces)
    with pytest.raises(ValueError):
      open(os.path.join(d, "a.xyz"), "wb").close()
      Mesh.load(os.path.join(d, "a.xyz"))

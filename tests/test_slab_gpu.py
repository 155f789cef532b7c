"""Slab sharding: N shards of one volume must assemble to exactly the
single-shot mesh.  Here the N shards run as N handles of ONE process on one GPU (the exchange steps
are done by hand); tests/multi_gpu_check.py runs the same thing as N processes over NCCL."""
import numpy as np
import pytest

from oracle.oracle import OracleMesher, assert_same_mesh, random_volume, voronoi_volume
from zmesh_b200.sharded import assemble, offsets_from_directories, slab_planes

pytestmark = pytest.mark.gpu


def _kat_volume():
  """Slab-boundary known-answer volume: a label that crosses every boundary plane, single
  voxels sitting on and next to boundary planes, a 2-voxel-thick sheet parallel to the slabs."""
  v = np.zeros((20, 18, 24), dtype=np.uint32, order="F")
  v[3:6, 3:6, :] = 5            # column through every slab
  for z in (5, 6, 7, 11, 12, 13, 17, 18):
    v[10, 9, z] = 9             # isolated voxels on / next to the cut planes of 2, 3, 4 shards
    v[12 + (z % 3), 4, z] = 11
  v[14:19, 10:16, 11:13] = 7    # sheet straddling the middle cut
  return v


def run_shards(Mesher, vol, res, close, nshards, normals=False):
  import torch
  axis = 0 if vol.flags.c_contiguous and not vol.flags.f_contiguous else 2
  order = "C" if axis == 0 else "F"
  full = vol.shape[axis]
  ms, dirs, keep = [], [], []
  for r in range(nshards):
    cube_lo, cube_hi, in_lo, in_hi, last = slab_planes(full, close, r, nshards)
    sl = [slice(None)] * 3
    sl[axis] = slice(in_lo, in_hi)
    sub = np.asarray(vol[tuple(sl)], order=order)
    m = Mesher(res, device=0)
    m.mesh_slab(sub, full, in_lo, cube_lo, cube_hi, last, close=close)
    ms.append(m)
    dirs.append(m.directory())
  ls, ns = [d[0] for d in dirs], [d[1] for d in dirs]
  for r, m in enumerate(ms):
    m.set_label_offsets(ls[r], offsets_from_directories(ls, ns, r))
  for r in range(nshards - 1, 0, -1):
    t = torch.empty(ms[r].plane_elems(), dtype=torch.int32, device="cuda:0")
    ms[r].export_plane(t.data_ptr())
    ms[r].sync()
    ms[r - 1].set_foreign_plane(t.data_ptr())
    keep.append(t)
  if normals:
    # contributions of every shard's top cube layer to the next shard's first-plane vertices
    nplanes = {}
    for r, m in enumerate(ms):
      if r < nshards - 1:
        nplanes[r] = torch.empty(3 * m.plane_elems(), dtype=torch.float32, device="cuda:0")
        m.set_normal_plane(nplanes[r].data_ptr())
      m.finalize(normals=True, voxel_centered=True)
      m.sync()
    for r, m in enumerate(ms):
      if r > 0:
        m.add_normal_plane(nplanes[r - 1].data_ptr())
      m.finish_normals()
    keep.append(nplanes)
  ids = sorted(set(i for m in ms for i in m.ids()))
  out = {}
  for lbl in ids:
    parts = []
    for m in ms:
      g = m.get(lbl, normals=normals, voxel_centered=True)
      part = (g.vertices, g.faces, g.normals) if normals else (g.vertices, g.faces)
      parts.append(part if len(g.vertices) or len(g.faces) else None)
    out[lbl] = assemble(parts)
  return ids, out, keep


CASES = [
  ("voronoi_u64_F", lambda: voronoi_volume((40, 36, 50), 12, np.uint64, 3, "F"), (4, 4, 40), False, (2, 3, 5)),
  ("voronoi_u64_F_close", lambda: voronoi_volume((40, 36, 50), 12, np.uint64, 3, "F"), (4, 4, 40), True, (2, 4)),
  ("random_u32_C_close", lambda: random_volume((30, 20, 33), 40, np.uint32, 7, "C"), (1, 2, 3), True, (2, 3)),
  ("random_u16_F_thin", lambda: random_volume((34, 9, 17), 6, np.uint16, 8, "F"), (1, 1, 1), False, (2, 8)),
  ("tile_aligned_u8_F", lambda: random_volume((64, 16, 33), 5, np.uint8, 9, "F"), (1, 1, 1), False, (2, 4)),
  ("boundary_kat", _kat_volume, (4, 4, 40), False, (2, 3, 4)),
  ("boundary_kat_close", _kat_volume, (4, 4, 40), True, (2, 3, 4)),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_shards_assemble_to_single_shot(build_all, case):
  from zmesh_b200 import Mesher
  name, make, res, close, shard_counts = case
  vol = make()
  cpu = OracleMesher(res, "port")
  cpu.mesh(vol, close=close)
  want_ids = sorted(cpu.ids())
  for n in shard_counts:
    ids, meshes, _keep = run_shards(Mesher, vol, res, close, n)
    assert ids == want_ids, (name, n)
    for lbl in ids:
      assert_same_mesh(meshes[lbl], cpu.get(lbl, normals=False, voxel_centered=True), what=f"{name} n={n} label {lbl}")


NORMAL_CASES = [c for c in CASES if c[0] in ("voronoi_u64_F_close", "random_u32_C_close", "boundary_kat", "tile_aligned_u8_F")]


@pytest.mark.parametrize("case", NORMAL_CASES, ids=[c[0] for c in NORMAL_CASES])
def test_shards_normals_match_single_shot(build_all, case):
  """Normals across shards: the top cube layer's contributions travel to the next shard (tolerance 1e-5
  absolute on unit vectors, the bar of the single-GPU path: float32 accumulation order differs)."""
  from zmesh_b200 import Mesher
  name, make, res, close, shard_counts = case
  vol = make()
  cpu = OracleMesher(res, "port")
  cpu.mesh(vol, close=close)
  for n in shard_counts[:2]:
    ids, meshes, _keep = run_shards(Mesher, vol, res, close, n, normals=True)
    assert ids == sorted(cpu.ids()), (name, n)
    for lbl in ids:
      assert meshes[lbl].normals is not None
      assert_same_mesh(meshes[lbl], cpu.get(lbl, normals=True, voxel_centered=True), 1e-5, what=f"{name} n={n} label {lbl}")


def test_slab_requires_exchange(build_all):
  from zmesh_b200 import Mesher
  vol = random_volume((20, 20, 20), 5, np.uint32, 1, "F")
  cube_lo, cube_hi, in_lo, in_hi, last = slab_planes(20, False, 0, 2)
  m = Mesher((1, 1, 1), device=0)
  m.mesh_slab(np.asfortranarray(vol[:, :, in_lo:in_hi]), 20, in_lo, cube_lo, cube_hi, last)
  with pytest.raises(RuntimeError, match="zm_set_foreign_plane"):
    m.get(m.ids()[0])
  with pytest.raises(ValueError):  # buffer does not cover the halo plane
    m.mesh_slab(np.asfortranarray(vol[:, :, in_lo:in_hi - 1]), 20, in_lo, cube_lo, cube_hi, last)


def test_multi_process_nccl(build_all):
  """Real one-process-per-GPU run over NCCL (needs >= 2 GPUs: `gpurun --gpus 2`)."""
  import os
  import subprocess
  import sys
  import torch
  n = torch.cuda.device_count()
  if n < 2:
    pytest.skip("needs at least 2 GPUs")
  n = 2 if n < 4 else (4 if n < 8 else 8)
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                      "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(root, "tests", "multi_gpu_check.py")],
                     capture_output=True, text=True, timeout=900)
  assert r.returncode == 0 and "MULTI_GPU_CHECK OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


MULTI_CASES = [
  ("voronoi_u64_F_close", lambda: voronoi_volume((72, 64, 96), 20, np.uint64, 3, "F"), (4, 4, 40), True),
  ("random_u32_C", lambda: random_volume((40, 20, 33), 40, np.uint32, 7, "C"), (1, 2, 3), False),
  ("boundary_kat_close", _kat_volume, (4, 4, 40), True),
  ("thin_one_device", lambda: random_volume((9, 9, 2), 4, np.uint8, 1, "F"), (1, 1, 1), False),
]


@pytest.mark.parametrize("case", MULTI_CASES, ids=[c[0] for c in MULTI_CASES])
def test_multi_device_mesher_in_one_process(build_all, case):
  """`Mesher(voxel_res, devices=[...])` (multi-GPU behind the drop-in API, one process, native NCCL
  step): every label -- vertices, faces, normals, the legacy accessor, erase -- equals the oracle's.  Needs >= 2 GPUs."""
  import torch
  from zmesh_b200 import Mesher, MultiDeviceMesher
  ndev = torch.cuda.device_count()
  if ndev < 2:
    pytest.skip("needs at least 2 GPUs")
  name, make, res, close = case
  vol = make()
  cpu = OracleMesher(res, "port")
  cpu.mesh(vol, close=close)
  for devices in ([0, 1], list(range(min(ndev, 4)))):
    m = Mesher(res, devices=devices)
    assert isinstance(m, MultiDeviceMesher)
    m.mesh(vol, close=close)
    ids = m.ids()
    assert ids == sorted(cpu.ids()), (name, devices)
    for lbl in ids:
      assert_same_mesh(m.get(lbl, voxel_centered=True), cpu.get(lbl, voxel_centered=True), what=f"{name} {devices} {lbl}")
    for lbl in ids[::3]:
      assert_same_mesh(m.get(lbl, normals=True), cpu.get(lbl, normals=True), 1e-5, what=f"{name} {devices} {lbl} normals")
      assert_same_mesh(m.get_mesh(lbl), cpu.get_mesh(lbl), what=f"{name} {devices} {lbl} legacy")
    if ids:
      assert m.erase(ids[0]) is True and m.erase(ids[0]) is False and m.ids() == ids[1:]
      assert m.get(ids[0]).empty()
    m.clear()
    assert m.ids() == []

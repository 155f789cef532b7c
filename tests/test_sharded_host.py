"""Host-side logic of the multi-GPU path on CPU: slab partition, directory all-gather (gloo,
world_size 2) and label index offsets, part assembly."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_slab_planes_cover_every_cube_once():
  from zmesh_b200.sharded import slab_planes
  for full in (2, 5, 64, 513):
    for close in (False, True):
      pad = 1 if close else 0
      ncube = full + 2 * pad - 1
      for world in (1, 2, 3, 8):
        if world > ncube:
          with pytest.raises(ValueError):
            slab_planes(full, close, 0, world)
          continue
        prev_hi = 0
        for r in range(world):
          lo, hi, in_lo, in_hi, last = slab_planes(full, close, r, world)
          assert lo == prev_hi and hi > lo
          assert in_lo == max(lo - pad, 0) and in_hi == min(hi - pad, full - 1) + 1
          assert last == (r == world - 1)
          prev_hi = hi
        assert prev_hi == ncube


def test_offsets_and_assemble():
  from zmesh_b200.sharded import assemble, offsets_from_directories
  ls = [np.array([5, 7, 9], np.uint64), np.array([], np.uint64), np.array([9, 7, 13], np.uint64)]
  ns = [np.array([10, 20, 30], np.uint64), np.array([], np.uint64), np.array([4, 4, 4], np.uint64)]
  assert offsets_from_directories(ls, ns, 0).tolist() == [0, 0, 0]
  assert offsets_from_directories(ls, ns, 1).tolist() == []
  assert offsets_from_directories(ls, ns, 2).tolist() == [30, 20, 0]
  v0 = np.zeros((2, 3), np.float32); f0 = np.array([[0, 1, 2]], np.uint32)
  v1 = np.ones((1, 3), np.float32); f1 = np.array([[2, 1, 0]], np.uint32)
  m = assemble([(v0, f0), None, (v1, f1)])
  assert m.vertices.shape == (3, 3) and m.faces.tolist() == [[0, 1, 2], [2, 1, 0]]
  assert assemble([None, None]).vertices.shape == (0, 3)


def _worker(rank, world, port, q):
  import torch.distributed as dist
  sys.path.insert(0, ROOT)
  from zmesh_b200.sharded import all_gather_directories, offsets_from_directories
  dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
  labels = [np.array([3, 2 ** 63 + 5, 8], np.uint64), np.array([8, 3], np.uint64)][rank]
  nv = [np.array([7, 1, 2], np.uint64), np.array([5, 6], np.uint64)][rank]
  ls, ns = all_gather_directories(labels, nv)
  off = offsets_from_directories(ls, ns, rank)
  q.put((rank, [l.tolist() for l in ls], off.tolist()))
  dist.barrier()
  dist.destroy_process_group()


def test_directory_exchange_gloo_world2():
  import torch.multiprocessing as mp
  ctx = mp.get_context("spawn")
  q = ctx.Queue()
  port = 29500 + os.getpid() % 2000
  procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
  for p in procs:
    p.start()
  res = sorted(q.get(timeout=120) for _ in procs)
  for p in procs:
    p.join(timeout=60)
    assert p.exitcode == 0
  assert res[0][1] == [[3, 2 ** 63 + 5, 8], [8, 3]] and res[1][1] == res[0][1]
  assert res[0][2] == [0, 0, 0]
  assert res[1][2] == [2, 7]


def test_s1_row_mask_formulation_equals_marching_scan():
  """Pure-Python restatement of k_classify's two S1 formulations (tools/emulate_s1_rowmask.py): identical bit
  planes and active-cube masks on random tiles, volume-boundary, zero-fill and slab-shard cases included."""
  import importlib.util
  import os
  from tests.conftest import ROOT
  spec = importlib.util.spec_from_file_location("emulate_s1_rowmask", os.path.join(ROOT, "tools", "emulate_s1_rowmask.py"))
  mod = importlib.util.module_from_spec(spec)
  spec.loader.exec_module(mod)
  assert mod.check(trials=24, seed=7) == 24


def test_native_slab_range_matches_python(build_all):
  """zm_slab_range (the partition the native multi-GPU step uses, pure host arithmetic -- callable without a GPU) cuts
  the volume exactly like zmesh_b200.sharded.slab_planes."""
  import ctypes as C
  from zmesh_b200 import _lib
  from zmesh_b200.sharded import slab_planes
  lib = _lib.load()
  for full in (2, 5, 64, 513, 2048):
    for close in (False, True):
      ncube = full + (2 if close else 0) - 1
      for world in (1, 2, 3, 8):
        for rank in range(world):
          slab, lo, hi = _lib.zm_slab(), C.c_uint64(0), C.c_uint64(0)
          rc = lib.zm_slab_range(full, int(close), rank, world, C.byref(slab), C.byref(lo), C.byref(hi))
          if world > ncube:
            assert rc != 0
            continue
          assert rc == 0
          cube_lo, cube_hi, in_lo, in_hi, last = slab_planes(full, close, rank, world)
          assert (slab.cube_lo, slab.cube_hi, lo.value, hi.value, bool(slab.last)) == (cube_lo, cube_hi, in_lo, in_hi, last)
          assert slab.full_extent == full and slab.buf_lo == in_lo


def test_multi_device_assembly_is_concatenation():
  """Host logic of Mesher(devices=[...]): a label's mesh is its per-device parts concatenated in device order (face
  indices are already cross-device), empty parts are skipped, the id is set."""
  from zmesh_b200 import Mesh
  from zmesh_b200.multi import MultiDeviceMesher
  m = MultiDeviceMesher.__new__(MultiDeviceMesher)
  a = Mesh(np.arange(6, dtype=np.float32).reshape(2, 3), np.array([[0, 1, 2]], np.uint32), None)
  b = Mesh(np.arange(6, 9, dtype=np.float32).reshape(1, 3), np.array([[2, 1, 0], [0, 2, 1]], np.uint32), None)
  out = m._assemble(42, [a, Mesh(), b])
  assert out.id == 42 and out.vertices.shape == (3, 3) and out.faces.tolist() == [[0, 1, 2], [2, 1, 0], [0, 2, 1]]
  assert np.array_equal(out.vertices, np.arange(9, dtype=np.float32).reshape(3, 3)) and out.normals is None
  a.normals, b.normals = np.ones((2, 3)), np.zeros((1, 3))
  assert m._assemble(1, [a, b]).normals.shape == (3, 3)
  assert m._assemble(7, [Mesh(), Mesh()]).empty() and m._assemble(7, []).id == 7
  assert m._assemble(9, [b]).faces.shape == (2, 3)


def test_mesher_devices_argument_dispatch():
  """Mesher(voxel_res, devices=[d]) is a plain Mesher on d; only lists of two or more devices take the multi-device class
  (constructing either needs a GPU: here only the dispatch of __new__ is checked)."""
  import inspect
  from zmesh_b200 import Mesher
  sig = inspect.signature(Mesher.__init__)
  assert list(sig.parameters)[1:] == ["voxel_res", "device", "devices"]
  assert sig.parameters["device"].default == -1 and sig.parameters["devices"].default is None


def test_two_label_path_equals_per_voxel_definition():
  """Pure-Python restatement of k_classify's two-label path (tools/emulate_k2_path.py): slot planes, active-cube masks,
  A's slot count, the number of records and every cube's corner masks follow from one 33-bit mask per staged row exactly
  as the per-voxel definition gives them; boundary tiles, zero fill and slab shards included."""
  import importlib.util
  import os
  from tests.conftest import ROOT
  spec = importlib.util.spec_from_file_location("emulate_k2_path", os.path.join(ROOT, "tools", "emulate_k2_path.py"))
  mod = importlib.util.module_from_spec(spec)
  spec.loader.exec_module(mod)
  assert mod.check(trials=40, seed=5) >= 24

"""`Mesher(voxel_res, devices=[...])`: the drop-in API over several GPUs of one node in ONE process (no reference
counterpart: the reference is a single-threaded CPU library).

The volume is cut along its slowest memory axis into one slab per device (zm_slab_range); one host thread per device
drives the native slab step (zm_slab_step: NCCL all-gather of the label directories, neighbour transfer of the boundary
plane, pass 2), ctypes releases the GIL around it.  Vertices are owned by voxel, so every vertex exists on exactly one
device and a label's mesh is the concatenation of its per-device parts in device order -- face indices are already
cross-device (see zmesh_b200/sharded.py) -- bit-identical, as canonical sets, to the single-GPU mesh.
"""
from __future__ import annotations

import ctypes as C
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import _lib
from .mesh import Mesh
from .mesher import Mesher, as_volume3d


class MultiDeviceMesher:
  """Same methods, keyword names and defaults as `Mesher`; returned by `Mesher(voxel_res, devices=[d0, d1, ...])`."""

  def __init__(self, voxel_res, devices):
    self._devices = [int(d) for d in devices]
    if len(set(self._devices)) != len(self._devices):
      raise ValueError("devices must be distinct")
    self._parts = [Mesher(voxel_res, device=d) for d in self._devices]
    self._pool = ThreadPoolExecutor(len(self._parts))
    self._active = 0      # devices that took part in the last mesh() (1: the volume was too thin to cut)
    self._close = False
    self._normals_key = None
    self.voxel_res = voxel_res
    world = len(self._parts)
    ids = (Mesher.nccl_unique_id(), b"".join(Mesher.nccl_unique_id() for _ in range(world - 1)))
    self._each(lambda r, p: p.comm_init(ids[0], ids[1], world, r))

  def _each(self, fn, parts=None):
    parts = self._parts if parts is None else parts
    return list(self._pool.map(lambda rp: fn(rp[0], rp[1]), list(enumerate(parts))))

  @property
  def voxel_res(self):
    return self._voxel_res

  @voxel_res.setter
  def voxel_res(self, res):
    self._voxel_res = np.array(res, dtype=np.float32)
    for p in getattr(self, "_parts", []):
      p.voxel_res = res

  # -- the hot path -------------------------------------------------------------------------------
  def mesh(self, data, close: bool = False, preserve_order: bool = True):
    if hasattr(data, "__cuda_array_interface__") and not isinstance(data, np.ndarray):
      raise NotImplementedError("devices=[...]: pass the volume as a host array (each device uploads its own slab)")
    data = np.asarray(data)
    if data.dtype.itemsize not in (1, 2, 4, 8):
      raise TypeError(f"unsupported label dtype {data.dtype}")
    data = as_volume3d(data, bool(close))
    c_order = data.flags.c_contiguous
    full = int(data.shape[0] if c_order else data.shape[2])
    world = len(self._parts)
    self._close = bool(close)
    self._normals_key = None
    ncube = full + (2 if close else 0) - 1
    if ncube < world or min(data.shape) == 0:  # too thin to cut: one device does it all
      self._active = 1
      return self._parts[0].mesh(data, close=close)
    self._active = world
    lib = _lib.load()

    def step(r, p):
      slab, lo, hi = _lib.zm_slab(), C.c_uint64(0), C.c_uint64(0)
      rc = lib.zm_slab_range(full, 1 if close else 0, r, world, C.byref(slab), C.byref(lo), C.byref(hi))
      if rc != 0:
        raise ValueError("cannot cut the volume into one slab per device")
      sub = data[lo.value:hi.value] if c_order else data[:, :, lo.value:hi.value]
      p.slab_step(sub, full, lo.value, close=close, finalize=False)
    self._each(step)

  def _live(self):
    return self._parts[:max(self._active, 1)]

  def ids(self):
    return sorted(set(i for p in self._live() for i in p.ids()))

  def _finalize(self, normals: bool, voxel_centered: bool, transpose: bool):
    """Collective pass 2 (needed when normals cross the slab boundaries)."""
    if self._active <= 1:
      return
    key = (bool(normals), bool(voxel_centered), bool(transpose), tuple(float(x) for x in self._voxel_res))
    if self._normals_key == key:
      return

    def fin(r, p):
      off = np.ascontiguousarray(p.voxel_res, dtype=np.float32)
      p._check(p._lib.zm_slab_finalize(p._h, int(normals), int(voxel_centered), int(transpose),
                                       off.ctypes.data_as(C.POINTER(C.c_float))))
      p._stage = None
    self._each(fin, self._live())
    self._normals_key = key

  def _assemble(self, label, parts) -> Mesh:
    parts = [m for m in parts if len(m.vertices) or len(m.faces)]
    if not parts:
      mesh = Mesh()
    elif len(parts) == 1:
      mesh = parts[0]
    else:
      normals = None
      if all(m.normals is not None for m in parts):
        normals = np.concatenate([m.normals for m in parts])
      mesh = Mesh(np.concatenate([m.vertices for m in parts]), np.concatenate([m.faces for m in parts]), normals)
    mesh.id = int(label)
    return mesh

  def get(self, label, normals=False, reduction_factor=0, max_error=None, voxel_centered=False) -> Mesh:
    if reduction_factor:
      raise NotImplementedError("zmesh_b200 covers reduction_factor=0 only (no mesh simplification)")
    if normals:
      self._finalize(True, voxel_centered, False)
    live = self._live()
    if any(p._stage is None for p in live):  # first request: every device runs its pass 2 and bulk transfer at once
      parts = self._each(lambda r, p: p.get(label, normals=normals, voxel_centered=voxel_centered), live)
    else:
      parts = [p.get(label, normals=normals, voxel_centered=voxel_centered) for p in live]
    return self._assemble(label, parts)

  def get_mesh(self, mesh_id, normals=False, simplification_factor=0, max_simplification_error=40,
               voxel_centered=False) -> Mesh:
    if simplification_factor:
      raise NotImplementedError("zmesh_b200 covers simplification_factor=0 only")
    if normals:
      self._finalize(True, voxel_centered, True)
    return self._assemble(mesh_id, [p.get_mesh(mesh_id, normals=normals, voxel_centered=voxel_centered) for p in self._live()])

  def erase(self, segid) -> bool:
    return any([p.erase(segid) for p in self._live()])

  def clear(self):
    for p in self._parts:
      p.clear()

  def compute_normals(self, mesh: Mesh) -> Mesh:
    return self._parts[0].compute_normals(mesh)

  def simplify(self, *args, **kwargs):
    raise NotImplementedError("zmesh_b200 covers reduction_factor=0 only (no mesh simplification)")

  def stats(self):
    return [p.stats() for p in self._live()]

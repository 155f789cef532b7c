"""`Mesher`: drop-in for the reference's Python class (zmesh/_zmesh.pyx:435-696) over the C ABI.

Same methods, keyword names, defaults and observable quirks for the path
`mesh(labels, close=) -> get(label, normals=, reduction_factor=0, voxel_centered=)`:

  * labels of any 1/2/4/8-byte dtype are compared as unsigned bit patterns (:971, :1008);
  * C- or Fortran-contiguous input is used in place, anything else is copied to C order (:499-500);
  * close=True meshes the volume as if zero-padded by one voxel and does NOT shift the
    coordinates back (:502-506);
  * the resolution used for vertices is the one captured by mesh() (:494), the voxel_centered
    offset uses the *current* `voxel_res` (:429-430, :581);
  * normals come back as float64 (:151), missing labels give an empty Mesh with `.id` set;
  * `ids()` is sorted ascending (the reference's order is unspecified).

reduction_factor > 0 (mesh simplification) is outside this package's scope and raises
NotImplementedError; there is no CPU fallback for anything.
"""
from __future__ import annotations

import ctypes as C
import os
import sys
import weakref

import numpy as np

from . import _lib
from .mesh import Mesh

_ERR = {1: ValueError, 2: RuntimeError, 3: MemoryError, 4: ValueError, 5: RuntimeError}


def _f3(a):
  a = np.ascontiguousarray(a, dtype=np.float32)
  if a.shape != (3,):
    raise ValueError("voxel_res must have three components")
  return a


def as_volume3d(data: np.ndarray, close: bool = False) -> np.ndarray:
  """The (sx, sy, sz) array the reference meshes for `data`, sharing memory with it where the reference
  does (zmesh/_zmesh.pyx:499-506 and the typed wrappers, e.g. :970-976):

    * fewer than three axes: IndexError (the reference indexes data.shape[2]);
    * neither C- nor Fortran-contiguous: copied to C order (:499-500);
    * more than three axes: the reference flattens the buffer in its memory order and hands C++ the extents
      shape[:3], i.e. it meshes the FIRST sx*sy*sz elements of the buffer laid out as (sx, sy, sz) in that
      order.  For trailing axes of extent 1 this is the array without them; in general it is what
      test_fanc_bug[transpose=True] exercises (automated_test.py:195-213: shape (1, 128, 512, 512)).
      With close=True the reference's padded copy broadcasts over the extra axes and then meshes a
      prefix of THAT buffer, which is not meaningful: only extra axes of extent 1 are accepted here.
  """
  if data.ndim < 3:
    raise IndexError("tuple index out of range")
  if not data.flags.c_contiguous and not data.flags.f_contiguous:
    data = np.ascontiguousarray(data)
  if data.ndim == 3:
    return data
  shape = tuple(int(s) for s in data.shape[:3])
  n3 = shape[0] * shape[1] * shape[2]
  if close and int(data.size) != n3:
    raise ValueError("close=True needs a 3d volume (extra axes must have extent 1)")
  order = "C" if data.flags.c_contiguous else "F"
  return data.reshape(-1, order=order)[:n3].reshape(shape, order=order)


class _PinnedBlock:
  """Page-locked host block the bulk device-to-host copy lands in.  Arrays handed to the user are views of the
  block's ctypes buffer and keep that buffer alive; the memory is returned (cudaFreeHost) by a finalizer on the
  buffer, i.e. when the mesher has dropped the block AND the last view is gone -- no reference cycle, no __del__."""

  def __init__(self, lib, nbytes: int):
    self.nbytes = max(int(nbytes), 1)
    ptr = lib.zm_host_alloc(self.nbytes)
    if not ptr:
      raise MemoryError(f"zmesh_b200: cannot allocate {self.nbytes} bytes of pinned host memory")
    self.ptr = ptr
    self._buf = (C.c_ubyte * self.nbytes).from_address(ptr)
    weakref.finalize(self._buf, lib.zm_host_free, ptr)
    self._live = []

  def view(self, dtype, shape, offset: int) -> np.ndarray:
    root = np.frombuffer(self._buf, dtype=dtype, count=int(np.prod(shape)), offset=offset)
    self._live.append(weakref.ref(root))  # every view/slice handed out has `root` as its base
    return root.reshape(shape)

  def idle(self) -> bool:
    self._live = [r for r in self._live if r() is not None]
    return not self._live


class _Stage:
  """All labels' final arrays on the host (one bulk transfer), sliced per label by get()."""
  __slots__ = ("key", "v", "f", "n", "index", "given")


class Mesher:
  """Represents a meshed volume: call mesher.mesh(labels), then mesher.get(label)."""

  def __new__(cls, voxel_res=None, device: int = -1, devices=None):
    # Mesher(voxel_res, devices=[0, 1, ...]): the same API over several GPUs of the node (zmesh_b200/multi.py)
    if cls is Mesher and devices is not None and len(devices) > 1:
      from .multi import MultiDeviceMesher
      return MultiDeviceMesher(voxel_res, devices)
    return super().__new__(cls)

  def __init__(self, voxel_res, device: int = -1, devices=None):
    if devices is not None and len(devices) == 1:
      device = devices[0]
    self._lib = _lib.load()
    self.voxel_res = voxel_res
    self._device = int(device)
    self._h = C.c_void_p()
    self._max_label = None
    self._stage = None
    self._blocks = []
    self._erased = set()
    res = self._voxel_res
    rc = self._lib.zm_create(res.ctypes.data_as(C.POINTER(C.c_float)), self._device, C.byref(self._h))
    if rc != 0:
      msg = self._lib.zm_last_error(None)
      raise _ERR.get(rc, RuntimeError)(f"zmesh_b200: {msg.decode() if msg else rc}")

  def __del__(self):
    h = getattr(self, "_h", None)
    if h:
      try:
        self._lib.zm_destroy(h)
      except Exception:
        pass
      self._h = None

  # -- plumbing -----------------------------------------------------------------------------------
  def _check(self, rc: int):
    if rc != 0:
      msg = self._lib.zm_last_error(self._h)
      raise _ERR.get(rc, RuntimeError)(f"zmesh_b200: {msg.decode() if msg else rc}")

  def _label_arg(self, label) -> int:
    label = int(label)
    if self._max_label is not None and not (0 <= label <= self._max_label):
      # the reference's typed Cython signature raises the same for out-of-range ids
      raise OverflowError("can't convert label to the meshed volume's label type")
    if not (0 <= label < 2 ** 64):
      raise OverflowError("label does not fit 64 bits")
    return label

  @property
  def voxel_res(self):
    return self._voxel_res

  @voxel_res.setter
  def voxel_res(self, res):
    self._voxel_res = np.array(res, dtype=np.float32)

  # -- the hot path -------------------------------------------------------------------------------
  def mesh(self, data, close: bool = False, preserve_order: bool = True):
    """Triggers the multi-label meshing process; afterwards call mesher.get.

    data: 3d array (numpy, or any object exposing __cuda_array_interface__ such as a CUDA
      torch tensor, which is consumed in place on the device).
    close: close meshes that touch the volume boundary (virtual one-voxel zero border).
    preserve_order: accepted for compatibility, ignored (as in the reference)."""
    res = _f3(self._voxel_res)
    self._stage = None
    self._erased = set()
    self._check(self._lib.zm_set_resolution(self._h, res.ctypes.data_as(C.POINTER(C.c_float))))

    cai = getattr(data, "__cuda_array_interface__", None)
    if cai is not None and not isinstance(data, np.ndarray):
      return self._mesh_device(data, cai, bool(close))

    data = np.asarray(data) if not isinstance(data, np.ndarray) else data
    nbytes = data.dtype.itemsize
    if nbytes not in (1, 2, 4, 8):
      raise TypeError(f"unsupported label dtype {data.dtype}")
    data = as_volume3d(data, bool(close))
    shape = tuple(int(s) for s in data.shape)
    c_order = 1 if data.flags.c_contiguous else 0
    self._max_label = (1 << (8 * nbytes)) - 1
    self._check(self._call_mesh(data.ctypes.data, nbytes, shape, c_order, close, 0))

  def _call_mesh(self, ptr, nbytes, shape, c_order, close, mem_kind):
    slab = getattr(self, "_slab", None)
    if slab is None:
      return self._lib.zm_mesh(self._h, C.c_void_p(ptr), nbytes, shape[0], shape[1], shape[2], c_order,
                               1 if close else 0, mem_kind)
    return self._lib.zm_mesh_slab(self._h, C.c_void_p(ptr), nbytes, shape[0], shape[1], shape[2], c_order,
                                  1 if close else 0, mem_kind, C.byref(slab))

  # -- multi-GPU slab pieces (driven by zmesh_b200.sharded.ShardedMesher) ----------------------------
  def mesh_slab(self, data, full_extent, buf_lo, cube_lo, cube_hi, last, close: bool = False):
    """Mesh one slab of a larger volume (see zm_mesh_slab in include/zmesh_b200.h): `data` holds the
    input planes [buf_lo, buf_lo + n) along the slowest memory axis; the shard owns the cubes whose
    origin (extended coordinates) lies in [cube_lo, cube_hi)."""
    self._slab = _lib.zm_slab(int(full_extent), int(buf_lo), int(cube_lo), int(cube_hi), 1 if last else 0)
    try:
      return self.mesh(data, close=close)
    finally:
      self._slab = None

  def directory(self):
    """(labels, n_vertices, n_faces) of the last mesh call, storage order (uint64 arrays)."""
    n = int(self._lib.zm_num_directory(self._h))
    out = [np.empty(n, dtype=np.uint64) for _ in range(3)]
    if n:
      p = [o.ctypes.data_as(C.POINTER(C.c_uint64)) for o in out]
      self._check(self._lib.zm_directory(self._h, p[0], p[1], p[2], n))
    return tuple(out)

  def set_label_offsets(self, labels, offsets):
    labels = np.ascontiguousarray(labels, dtype=np.uint64)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint32)
    self._check(self._lib.zm_set_label_offsets(self._h, labels.ctypes.data_as(C.POINTER(C.c_uint64)),
                                               offsets.ctypes.data_as(C.POINTER(C.c_uint32)), labels.size))

  # native multi-GPU step: NCCL driven from the C++ layer (zm_comm_init / zm_slab_step)
  @staticmethod
  def nccl_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    lib = _lib.load()
    rc = lib.zm_nccl_unique_id(buf)
    if rc != 0:
      msg = lib.zm_last_error(None)
      raise RuntimeError(f"zmesh_b200: {msg.decode() if msg else rc}")
    return buf.raw

  def comm_init(self, id_collectives: bytes, id_pairs: bytes, world: int, rank: int):
    """id_pairs: (world - 1) * 128 bytes, id k names the 2-rank communicator of shards k and k + 1."""
    if len(id_pairs) != 128 * max(int(world) - 1, 0):
      raise ValueError("id_pairs must hold world - 1 ids of 128 bytes")
    self._check(self._lib.zm_comm_init(self._h, id_collectives, id_pairs if id_pairs else None, int(world), int(rank)))

  def slab_step(self, data, full_extent: int, buf_lo: int, close=False, finalize=True, normals=False, voxel_centered=False):
    """One whole slab step on this rank (see zm_slab_step): `data` holds the input planes [buf_lo, buf_lo + n) along the
    slowest memory axis (numpy array, or a device array exposing __cuda_array_interface__)."""
    res = _f3(self._voxel_res)
    self._stage = None
    self._erased = set()
    self._check(self._lib.zm_set_resolution(self._h, res.ctypes.data_as(C.POINTER(C.c_float))))
    cai = getattr(data, "__cuda_array_interface__", None)
    if cai is not None and not isinstance(data, np.ndarray):
      ptr, nbytes, shape, c_order = self._device_view(data, cai)
      mem_kind = 1
    else:
      data = np.asarray(data)
      nbytes = data.dtype.itemsize
      if nbytes not in (1, 2, 4, 8):
        raise TypeError(f"unsupported label dtype {data.dtype}")
      data = as_volume3d(data, bool(close))
      shape = tuple(int(x) for x in data.shape)
      c_order = 1 if data.flags.c_contiguous else 0
      ptr, mem_kind = data.ctypes.data, 0
    self._max_label = (1 << (8 * nbytes)) - 1
    off = _f3(self._voxel_res)
    self._check(self._lib.zm_slab_step(self._h, C.c_void_p(ptr), nbytes, shape[0], shape[1], shape[2], c_order,
                                       1 if close else 0, mem_kind, int(full_extent), int(buf_lo), int(bool(finalize)),
                                       int(bool(normals)), int(bool(voxel_centered)), off.ctypes.data_as(C.POINTER(C.c_float))))

  def export_directory(self, dst_device_ptr: int, capacity: int):
    """Device-side directory exchange (no host round trip): see zm_export_directory."""
    self._check(self._lib.zm_export_directory(self._h, C.c_void_p(int(dst_device_ptr)), int(capacity)))

  def import_directories(self, all_device_ptr: int, world: int, rank: int, capacity: int):
    self._check(self._lib.zm_import_directories(self._h, C.c_void_p(int(all_device_ptr)), int(world), int(rank), int(capacity)))

  def plane_elems(self) -> int:
    return int(self._lib.zm_plane_elems(self._h))

  def export_plane(self, dst_device_ptr: int):
    self._check(self._lib.zm_export_plane(self._h, C.c_void_p(int(dst_device_ptr))))

  def set_foreign_plane(self, src_device_ptr):
    self._check(self._lib.zm_set_foreign_plane(self._h, C.c_void_p(int(src_device_ptr)) if src_device_ptr else None))

  def set_normal_plane(self, dst_device_ptr):
    """Device buffer (3 * plane_elems() float32) for the normal contributions to the next shard's vertices."""
    self._check(self._lib.zm_set_normal_plane(self._h, C.c_void_p(int(dst_device_ptr)) if dst_device_ptr else None))

  def add_normal_plane(self, src_device_ptr: int):
    self._check(self._lib.zm_add_normal_plane(self._h, C.c_void_p(int(src_device_ptr))))

  def finish_normals(self):
    self._check(self._lib.zm_finish_normals(self._h))

  def set_stream(self, cuda_stream):
    """Queue all work on a caller-owned CUDA stream (integer cudaStream_t); None restores the own one."""
    self._check(self._lib.zm_set_stream(self._h, C.c_void_p(int(cuda_stream)) if cuda_stream else None))
    self._user_stream = int(cuda_stream) if cuda_stream else None

  def stream_handle(self):
    """The caller-owned stream set by set_stream (integer cudaStream_t), or None for the handle's own."""
    return getattr(self, "_user_stream", None)

  def _device_view(self, obj, cai):
    """(pointer, label bytes, (sx, sy, sz), c_order) of a device array; the mesher's stream is ordered after the work
    that produced it."""
    shape = tuple(int(s) for s in cai["shape"])
    if len(shape) < 3:
      raise IndexError("tuple index out of range")
    if len(shape) > 3:
      if int(np.prod(shape[3:])) != 1:
        raise ValueError("only the first three axes may have extent > 1")
    typestr = cai["typestr"]
    nbytes = int(typestr[2:])
    if nbytes not in (1, 2, 4, 8):
      raise TypeError(f"unsupported label dtype {typestr}")
    strides = cai.get("strides")
    s3 = shape[:3]
    c_strides = (s3[1] * s3[2] * nbytes, s3[2] * nbytes, nbytes)
    f_strides = (nbytes, s3[0] * nbytes, s3[0] * s3[1] * nbytes)
    if strides is None or tuple(strides[:3]) == c_strides:
      c_order = 1
    elif tuple(strides[:3]) == f_strides:
      c_order = 0
    else:
      raise ValueError("device arrays must be C- or Fortran-contiguous")
    ptr = int(cai["data"][0])
    # order the mesher's stream after the work that produced the array: the stream the interface names (v3), else
    # torch's current stream for a torch tensor, else the legacy default stream
    producer = cai.get("stream")
    if producer is None:
      torch = sys.modules.get("torch")
      if torch is not None and isinstance(obj, torch.Tensor):
        producer = int(torch.cuda.current_stream(obj.device).cuda_stream)
      else:
        producer = 1
    self._check(self._lib.zm_wait_stream(self._h, C.c_void_p(int(producer)) if producer else None))
    return ptr, nbytes, s3, c_order

  def _mesh_device(self, obj, cai, close: bool):
    ptr, nbytes, s3, c_order = self._device_view(obj, cai)
    self._max_label = (1 << (8 * nbytes)) - 1
    self._check(self._call_mesh(ptr, nbytes, s3, c_order, close, 1))

  def ids(self):
    n = int(self._lib.zm_num_ids(self._h))
    out = np.empty(n, dtype=np.uint64)
    if n:
      self._check(self._lib.zm_ids(self._h, out.ctypes.data_as(C.POINTER(C.c_uint64)), n))
    return out.tolist()

  def _fetch(self, label, normals: bool, voxel_centered: bool, transpose: bool) -> Mesh:
    label = self._label_arg(label)
    nv, nf = C.c_uint64(0), C.c_uint64(0)
    self._check(self._lib.zm_get_counts(self._h, label, C.byref(nv), C.byref(nf)))
    if nv.value == 0 or nf.value == 0:
      mesh = Mesh()
      mesh.id = label
      return mesh
    verts = np.empty((nv.value, 3), dtype=np.float32)
    faces = np.empty((nf.value, 3), dtype=np.uint32)
    nrm = np.empty((nv.value, 3), dtype=np.float32) if normals else None
    off = _f3(self._voxel_res)
    self._check(self._lib.zm_get(
      self._h, label, 1 if normals else 0, 1 if voxel_centered else 0, 1 if transpose else 0,
      off.ctypes.data_as(C.POINTER(C.c_float)), C.c_void_p(verts.ctypes.data), C.c_void_p(faces.ctypes.data),
      C.c_void_p(nrm.ctypes.data) if normals else None))
    mesh = Mesh(verts, faces, None)
    if normals:
      mesh.normals = nrm.astype(np.float64)  # the reference hands back float64 (zmesh/_zmesh.pyx:151)
    mesh.id = label
    return mesh

  def get(self, label, normals=False, reduction_factor=0, max_error=None, voxel_centered=False) -> Mesh:
    """label: the integer id of the mesh
    normals: whether to calculate vertex normals
    reduction_factor: must be 0 (simplification is out of scope of this package)
    voxel_centered: centre the mesh in the voxel (0.5, 0.5, 0.5) instead of at (0, 0, 0)."""
    if reduction_factor:
      raise NotImplementedError("zmesh_b200 covers reduction_factor=0 only (no mesh simplification)")
    return self._get_staged(label, bool(normals), bool(voxel_centered))

  # All labels are finalized on the device by the first get(); their arrays cross PCIe in ONE
  # transfer into pinned host memory and get() hands out per-label views of it (SURVEY.md 8b:
  # "get() becomes a slice of pinned host memory").  Ownership: every label's first result is a view
  # the caller may mutate freely -- a label asked for AGAIN is read back from the device, so a caller's
  # in-place edit (mesh.vertices += offset) never shows up in a later get().  The block (12 bytes per
  # vertex and face of ALL labels; ~20 GB for a 2048^3 volume) stays page-locked while any handed-out
  # array is alive.  When it cannot be allocated, or is larger than ZMESH_B200_PINNED_LIMIT bytes,
  # get() falls back to one device-to-host copy per label into ordinary numpy arrays.
  def _pinned(self, nbytes: int) -> _PinnedBlock:
    for b in self._blocks:
      if b.nbytes >= nbytes and b.idle():
        return b
    self._blocks = [b for b in self._blocks if not b.idle()][-2:]
    b = _PinnedBlock(self._lib, nbytes)
    self._blocks.append(b)
    return b

  def _build_stage(self, normals: bool, voxel_centered: bool) -> "_Stage":
    key = (voxel_centered, tuple(float(x) for x in self._voxel_res) if voxel_centered else None)
    self._stage = None
    off = _f3(self._voxel_res)
    view = _lib.zm_bulk_view()
    self._check(self._lib.zm_finalize(self._h, int(normals), int(voxel_centered), 0,
                                      off.ctypes.data_as(C.POINTER(C.c_float)), C.byref(view)))
    nv, nf, nl = int(view.n_vertices), int(view.n_faces), int(view.n_labels)
    st = _Stage()
    st.key = key
    st.given = set()
    nbytes = 12 * nv * (2 if normals else 1) + 12 * nf
    limit = int(os.environ.get("ZMESH_B200_PINNED_LIMIT", "0") or 0)
    blk = None
    if not (limit and nbytes > limit):
      try:
        blk = self._pinned(nbytes)
      except MemoryError:
        blk = None
    if blk is None:  # per-label transfers (zm_get) instead of the bulk view
      st.v = st.f = None
      st.n = True if normals else None
      st.index = None
      self._stage = st
      return st
    st.v = blk.view(np.float32, (nv, 3), 0)
    st.f = blk.view(np.uint32, (nf, 3), 12 * nv)
    st.n = blk.view(np.float32, (nv, 3), 12 * nv + 12 * nf) if normals else None
    if nv or nf:
      self._check(self._lib.zm_fetch_all(self._h, C.c_void_p(st.v.ctypes.data), C.c_void_p(st.f.ctypes.data),
                                         C.c_void_p(st.n.ctypes.data) if normals else None))
    labels = np.ctypeslib.as_array(view.labels_host, shape=(nl,)).tolist() if nl else []
    voff = np.ctypeslib.as_array(view.voff_host, shape=(nl + 1,)).tolist() if nl else [0]
    foff = np.ctypeslib.as_array(view.foff_host, shape=(nl + 1,)).tolist() if nl else [0]
    st.index = {labels[i]: (voff[i], voff[i + 1], foff[i], foff[i + 1]) for i in range(nl)}
    self._stage = st
    return st

  def _get_staged(self, label, normals: bool, voxel_centered: bool) -> Mesh:
    label = self._label_arg(label)
    st = self._stage
    key = (voxel_centered, tuple(float(x) for x in self._voxel_res) if voxel_centered else None)
    if st is None or st.key != key or (normals and st.n is None):
      st = self._build_stage(normals or (st is not None and st.key == key and st.n is not None), voxel_centered)
    if st.index is None:  # no bulk block: one device-to-host copy per label
      return self._fetch(label, normals, voxel_centered, transpose=False)
    rng = st.index.get(label)
    if rng is None or label in self._erased or rng[3] == rng[2]:
      mesh = Mesh()
      mesh.id = label
      return mesh
    if label in st.given:  # the caller already owns (and may have edited) this label's staged arrays
      return self._fetch(label, normals, voxel_centered, transpose=False)
    st.given.add(label)
    # the reference hands normals back as float64 (zmesh/_zmesh.pyx:151)
    return Mesh._wrap(st.v[rng[0]:rng[1]], st.f[rng[2]:rng[3]],
                      st.n[rng[0]:rng[1]].astype(np.float64) if normals else None, label)

  def get_mesh(self, mesh_id, normals=False, simplification_factor=0, max_simplification_error=40,
               voxel_centered=False) -> Mesh:
    """Deprecated accessor kept for compatibility: like get() but with x and z swapped."""
    if simplification_factor:
      raise NotImplementedError("zmesh_b200 covers simplification_factor=0 only")
    return self._fetch(mesh_id, bool(normals), bool(voxel_centered), transpose=True)

  def compute_normals(self, mesh: Mesh) -> Mesh:
    verts = np.ascontiguousarray(mesh.vertices, dtype=np.float32)
    faces = np.ascontiguousarray(mesh.faces, dtype=np.uint32)
    out = np.zeros((verts.shape[0], 3), dtype=np.float32)
    self._check(self._lib.zm_compute_normals(self._h, C.c_void_p(verts.ctypes.data), verts.shape[0],
                                             C.c_void_p(faces.ctypes.data), faces.shape[0],
                                             C.c_void_p(out.ctypes.data)))
    mesh.normals = out.astype(np.float64)
    return mesh

  def simplify(self, *args, **kwargs):
    raise NotImplementedError("zmesh_b200 covers reduction_factor=0 only (no mesh simplification)")

  def erase(self, segid) -> bool:
    existed = C.c_int(0)
    label = self._label_arg(segid)
    self._check(self._lib.zm_erase(self._h, label, C.byref(existed)))
    self._erased.add(label)
    return bool(existed.value)

  def clear(self):
    self._stage = None
    self._check(self._lib.zm_clear(self._h))

  # -- extras (no reference counterpart) ------------------------------------------------------------
  def stats(self) -> dict:
    st = _lib.zm_stats_t()
    self._check(self._lib.zm_stats(self._h, C.byref(st)))
    return {name: getattr(st, name) for name, _ in st._fields_}

  def finalize(self, normals=False, voxel_centered=False, transpose=False):
    """Run the final gather for all labels on the device; returns the bulk directory."""
    view = _lib.zm_bulk_view()
    off = _f3(self._voxel_res)
    self._check(self._lib.zm_finalize(self._h, int(bool(normals)), int(bool(voxel_centered)), int(bool(transpose)),
                                      off.ctypes.data_as(C.POINTER(C.c_float)), C.byref(view)))
    n = int(view.n_labels)
    as_np = lambda p, k: (np.ctypeslib.as_array(p, shape=(k,)).copy() if k else np.zeros(0, dtype=np.uint64))
    return {
      "labels": as_np(view.labels_host, n), "voff": as_np(view.voff_host, n + 1), "foff": as_np(view.foff_host, n + 1),
      "n_vertices": int(view.n_vertices), "n_faces": int(view.n_faces),
      "vertices_dev": view.vertices_dev, "faces_dev": view.faces_dev, "normals_dev": view.normals_dev,
    }

  def precomputed(self, voxel_centered=False) -> dict:
    """{id: bytes-like} -- the Neuroglancer Precomputed object of every id (what Mesh.to_precomputed() returns for
    get(id), zmesh/mesh.py:257-269), laid out on the device and moved to the host in ONE transfer; the values are
    memoryviews of a pinned block (bytes(v) makes an owned copy)."""
    off = _f3(self._voxel_res)
    n, total = C.c_uint64(0), C.c_uint64(0)
    self._check(self._lib.zm_pack_precomputed(self._h, int(bool(voxel_centered)), off.ctypes.data_as(C.POINTER(C.c_float)),
                                              C.byref(n), C.byref(total)))
    if n.value == 0:
      return {}
    blk = self._pinned(int(total.value))
    buf = blk.view(np.uint8, (int(total.value),), 0)
    labels = np.empty(n.value, dtype=np.uint64)
    offs = np.empty(n.value + 1, dtype=np.uint64)
    self._check(self._lib.zm_fetch_precomputed(self._h, C.c_void_p(buf.ctypes.data), labels.ctypes.data_as(C.POINTER(C.c_uint64)),
                                               offs.ctypes.data_as(C.POINTER(C.c_uint64))))
    mv = memoryview(buf)
    o = offs.tolist()
    return {int(l): mv[o[i]:o[i + 1]] for i, l in enumerate(labels.tolist())}

  def finalize_begin(self, normals=False, voxel_centered=False, transpose=False):
    """Slab shards: start pass 2 for all tiles but the top layer (see zm_finalize_begin); follow with finalize()."""
    off = _f3(self._voxel_res)
    self._check(self._lib.zm_finalize_begin(self._h, int(bool(normals)), int(bool(voxel_centered)), int(bool(transpose)),
                                            off.ctypes.data_as(C.POINTER(C.c_float))))

  def fetch_all(self, normals=False):
    """After finalize(): all labels' arrays in one device-to-host transfer each."""
    st = self.stats()
    v = np.empty((st["n_vertices"], 3), dtype=np.float32)
    f = np.empty((st["n_faces"], 3), dtype=np.uint32)
    n = np.empty((st["n_vertices"], 3), dtype=np.float32) if normals else None
    self._check(self._lib.zm_fetch_all(self._h, C.c_void_p(v.ctypes.data), C.c_void_p(f.ctypes.data),
                                       C.c_void_p(n.ctypes.data) if normals else None))
    return v, f, n

  def sync(self):
    self._check(self._lib.zm_sync(self._h))

"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the hot path; never imported by zmesh_b200.

Two CPU checkers behind one Python class, both driven exactly as the reference's Python layer
drives its C++ (zmesh/_zmesh.pyx:435-696):

  * kind="port"      -> oracle/liboracle.so  (plain-C restatement, oracle/zmesh_oracle.c)
  * kind="reference" -> oracle/_ref/libzmesh_ref.so (the unmodified reference C++ behind a shim)

plus the canonical form used for "bit-exact per-label sets"  and the
deterministic synthetic volume generators.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) import this.
"""
from __future__ import annotations

import ctypes as C
import hashlib
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_LIB = os.path.join(HERE, "liboracle.so")
REF_LIB = os.path.join(HERE, "_ref", "libzmesh_ref.so")

_u64p = C.POINTER(C.c_uint64)
_u32p = C.POINTER(C.c_uint32)
_f32p = C.POINTER(C.c_float)


def build(force: bool = False) -> None:
  """Compile the C restatement (and, where /root/reference exists, the reference shim)."""
  need = force or not os.path.exists(PORT_LIB) or (
    os.path.getmtime(PORT_LIB) < os.path.getmtime(os.path.join(HERE, "zmesh_oracle.c")))
  if need:
    subprocess.check_call(["make", "-C", HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
  if os.path.isdir("/root/reference/zmesh") and (force or not os.path.exists(REF_LIB)):
    subprocess.check_call(["make", "-C", HERE, "_ref/libzmesh_ref.so"], stdout=subprocess.DEVNULL)


def have_reference() -> bool:
  return os.path.exists(REF_LIB)


_libs = {}


def _port():
  if "port" not in _libs:
    build()
    L = C.CDLL(PORT_LIB)
    L.zo_create.restype = C.c_void_p
    L.zo_create.argtypes = [_f32p]
    L.zo_destroy.argtypes = [C.c_void_p]
    L.zo_clear.argtypes = [C.c_void_p]
    L.zo_mesh.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int]
    L.zo_ids.restype = C.c_uint64
    L.zo_ids.argtypes = [C.c_void_p, _u64p, C.c_uint64]
    L.zo_erase.restype = C.c_int
    L.zo_erase.argtypes = [C.c_void_p, C.c_uint64]
    L.zo_get.argtypes = [C.c_void_p, C.c_uint64, C.c_int, _u64p, _u64p, _f32p, _u32p]
    L.zo_normals.argtypes = [_f32p, C.c_uint64, _u32p, C.c_uint64, _f32p]
    L.zo_voronoi.restype = None
    L.zo_voronoi.argtypes = [C.c_void_p, C.c_int, _u64p, _u64p, _u64p, C.c_uint32, C.c_uint64, C.c_int, C.c_uint64, C.c_uint64]
    _libs["port"] = L
  return _libs["port"]


def _ref():
  if "ref" not in _libs:
    if not os.path.exists(REF_LIB):
      raise RuntimeError("oracle/_ref/libzmesh_ref.so missing (build it where /root/reference exists)")
    L = C.CDLL(REF_LIB)
    L.zref_create.restype = C.c_void_p
    L.zref_create.argtypes = [C.c_int, C.c_int, _f32p]
    L.zref_destroy.argtypes = [C.c_void_p]
    L.zref_clear.argtypes = [C.c_void_p]
    L.zref_mesh.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int]
    L.zref_ids.restype = C.c_uint64
    L.zref_ids.argtypes = [C.c_void_p, _u64p, C.c_uint64]
    L.zref_erase.restype = C.c_int
    L.zref_erase.argtypes = [C.c_void_p, C.c_uint64]
    L.zref_get.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.POINTER(_f32p), _u64p, C.POINTER(_u32p), _u64p]
    L.zref_free.argtypes = [C.c_void_p]
    L.zref_normals.argtypes = [_f32p, C.c_uint64, _u32p, C.c_uint64, _f32p]
    _libs["ref"] = L
  return _libs["ref"]


class OracleMesh:
  def __init__(self, vertices, faces, normals=None, id=None):
    self.vertices = vertices
    self.faces = faces
    self.normals = normals
    self.id = id


class OracleMesher:
  """Mirror of the reference's Python `Mesher` (zmesh/_zmesh.pyx:435-696) over a CPU backend."""

  def __init__(self, voxel_res, kind: str = "port"):
    assert kind in ("port", "reference")
    self.kind = kind
    self.voxel_res = voxel_res
    self._h = None
    self._lib = _port() if kind == "port" else _ref()
    self._keep = None

  @property
  def voxel_res(self):
    return self._voxel_res

  @voxel_res.setter
  def voxel_res(self, res):
    self._voxel_res = np.array(res, dtype=np.float32)  # _zmesh.pyx:450-452

  def _destroy(self):
    if self._h is not None:
      (self._lib.zo_destroy if self.kind == "port" else self._lib.zref_destroy)(self._h)
      self._h = None

  def __del__(self):
    try:
      self._destroy()
    except Exception:
      pass

  def mesh(self, data: np.ndarray, close: bool = False):
    """_zmesh.pyx:454-508 (dense-array branch)."""
    self._destroy()
    shape = data.shape
    nbytes = np.dtype(data.dtype).itemsize
    pos_bits = 64 if (shape[0] > 1023 or shape[1] > 1023 or shape[2] > 511) else 32  # :477-492
    if not data.flags.c_contiguous and not data.flags.f_contiguous:
      data = np.ascontiguousarray(data)  # :499-500
    if close:  # :502-506
      tmp = np.zeros(np.array(data.shape) + 2, dtype=data.dtype, order="C")
      tmp[1:-1, 1:-1, 1:-1] = data
      data = tmp
    res = self._voxel_res.ctypes.data_as(_f32p)
    c_order = 1 if data.flags.c_contiguous else 0  # :975
    ptr = C.c_void_p(data.ctypes.data)
    sx, sy, sz = (int(s) for s in data.shape[:3])
    if self.kind == "port":
      self._h = self._lib.zo_create(res)
      self._lib.zo_mesh(self._h, ptr, nbytes, sx, sy, sz, c_order)
    else:
      self._h = self._lib.zref_create(pos_bits, nbytes, res)
      self._lib.zref_mesh(self._h, ptr, sx, sy, sz, c_order)

  def ids(self):
    f = self._lib.zo_ids if self.kind == "port" else self._lib.zref_ids
    n = int(f(self._h, None, 0))
    out = np.zeros(n, dtype=np.uint64)
    if n:
      f(self._h, out.ctypes.data_as(_u64p), n)
    return [int(x) for x in out]

  def _raw(self, label: int, transpose: bool):
    label = int(label)
    if self.kind == "port":
      nv, nf = C.c_uint64(0), C.c_uint64(0)
      self._lib.zo_get(self._h, label, int(transpose), C.byref(nv), C.byref(nf), None, None)
      v = np.zeros((nv.value, 3), dtype=np.float32)
      f = np.zeros((nf.value, 3), dtype=np.uint32)
      if nv.value:
        self._lib.zo_get(self._h, label, int(transpose), C.byref(nv), C.byref(nf),
                         v.ctypes.data_as(_f32p), f.ctypes.data_as(_u32p))
      return v, f
    pp, fp = _f32p(), _u32p()
    npf, nfi = C.c_uint64(0), C.c_uint64(0)
    self._lib.zref_get(self._h, label, int(transpose), C.byref(pp), C.byref(npf), C.byref(fp), C.byref(nfi))
    v = np.ctypeslib.as_array(pp, shape=(max(npf.value, 1),))[: npf.value].copy().reshape(-1, 3)
    f = np.ctypeslib.as_array(fp, shape=(max(nfi.value, 1),))[: nfi.value].copy().reshape(-1, 3)
    self._lib.zref_free(pp)
    self._lib.zref_free(fp)
    return v.astype(np.float32, copy=False), f.astype(np.uint32, copy=False)

  def compute_normals(self, vertices, faces):
    """_zmesh.pyx:138-152 -> chunk_mesh.hpp:345-384; result returned as float64 like the reference."""
    v = np.ascontiguousarray(vertices, dtype=np.float32)
    f = np.ascontiguousarray(faces, dtype=np.uint32)
    out = np.zeros((v.shape[0], 3), dtype=np.float32)
    fn = self._lib.zo_normals if self.kind == "port" else self._lib.zref_normals
    if v.shape[0]:
      fn(v.ctypes.data_as(_f32p), v.shape[0], f.ctypes.data_as(_u32p), f.size, out.ctypes.data_as(_f32p))
    return out.astype(np.float64)

  def _finish(self, label, v, f, normals, voxel_centered):
    n = self.compute_normals(v, f) if normals else None
    # _normalize_mesh (_zmesh.pyx:423-433), physical=True
    if voxel_centered:
      v += self._voxel_res
    v /= 2.0
    return OracleMesh(v, f, n, id=int(label))

  def get(self, label, normals=False, voxel_centered=False):
    """_zmesh.pyx:547-583 with reduction_factor=0."""
    v, f = self._raw(label, transpose=False)
    return self._finish(label, v, f, normals, voxel_centered)

  def get_mesh(self, label, normals=False, voxel_centered=False):
    """Legacy transposed accessor, _zmesh.pyx:514-545."""
    v, f = self._raw(label, transpose=True)
    return self._finish(label, v, f, normals, voxel_centered)

  def erase(self, label):
    f = self._lib.zo_erase if self.kind == "port" else self._lib.zref_erase
    return bool(f(self._h, int(label)))

  def clear(self):
    (self._lib.zo_clear if self.kind == "port" else self._lib.zref_clear)(self._h)


# ------------------------------------------------------------------------------------------------
# canonical form

def _rows_as_void(a: np.ndarray) -> np.ndarray:
  a = np.ascontiguousarray(a)
  return a.view(np.dtype((np.void, a.dtype.itemsize * a.shape[1]))).ravel()


def canonical_vertices(vertices: np.ndarray) -> np.ndarray:
  """float32 (V,3) -> uint32 bit patterns, rows sorted lexicographically."""
  b = np.ascontiguousarray(vertices, dtype=np.float32).view(np.uint32).reshape(-1, 3)
  order = np.lexsort((b[:, 2], b[:, 1], b[:, 0]))
  return b[order]


def canonical_faces(vertices: np.ndarray, faces: np.ndarray) -> np.ndarray:
  """(T,3) indices -> (T,9) uint32 vertex bit patterns, each triangle rotated so that its
  lexicographically smallest vertex comes first (keeps winding), rows sorted."""
  b = np.ascontiguousarray(vertices, dtype=np.float32).view(np.uint32).reshape(-1, 3)
  if len(faces) == 0:
    return np.zeros((0, 9), dtype=np.uint32)
  t = b[np.asarray(faces, dtype=np.int64)]  # (T,3,3)
  # generic lexicographic argmin over the 3 corners
  best = np.zeros(len(t), dtype=np.int64)
  for j in (1, 2):
    a = t[np.arange(len(t)), best]
    c = t[:, j]
    less = (c[:, 0] < a[:, 0]) | ((c[:, 0] == a[:, 0]) & ((c[:, 1] < a[:, 1]) | ((c[:, 1] == a[:, 1]) & (c[:, 2] < a[:, 2]))))
    best = np.where(less, j, best)
  idx = (best[:, None] + np.arange(3)[None, :]) % 3
  r = t[np.arange(len(t))[:, None], idx].reshape(len(t), 9)
  order = np.lexsort(tuple(r[:, i] for i in range(8, -1, -1)))
  return r[order]


def canonical_digest(vertices: np.ndarray, faces: np.ndarray) -> str:
  h = hashlib.sha256()
  cv = canonical_vertices(vertices)
  cf = canonical_faces(vertices, faces)
  h.update(np.uint64(len(cv)).tobytes()); h.update(cv.tobytes())
  h.update(np.uint64(len(cf)).tobytes()); h.update(cf.tobytes())
  return h.hexdigest()


def _mix64(x):
  x = np.asarray(x, dtype=np.uint64)
  with np.errstate(over="ignore"):
    x = (x ^ (x >> np.uint64(33))) * np.uint64(0xFF51AFD7ED558CCD)
    x = (x ^ (x >> np.uint64(33))) * np.uint64(0xC4CEB9FE1A85EC53)
    return x ^ (x >> np.uint64(33))


def multiset_digest(vertices: np.ndarray, faces: np.ndarray):
  """O(n) order-independent fingerprint of a mesh, for volumes whose canonical (sorted) form is too slow to build:
  (V, sum and xor of a 64-bit hash per vertex row, T, sum and xor of a hash per face), where the face hash is
  invariant under rotation of its three corners and changes under reflection (winding) -- the same equivalence
  as canonical_faces.  Equal digests of a mesh without duplicate vertices <=> equal canonical sets, up to 2^-64."""
  b = np.ascontiguousarray(vertices, dtype=np.float32).view(np.uint32).reshape(-1, 3).astype(np.uint64)
  with np.errstate(over="ignore"):
    hv = _mix64(b[:, 0] * np.uint64(0x9E3779B97F4A7C15) + _mix64(b[:, 1] * np.uint64(0xC2B2AE3D27D4EB4F) + _mix64(b[:, 2] + np.uint64(0x165667B19E3779F9))))
    f = np.asarray(faces, dtype=np.int64).reshape(-1, 3)
    ha, hb, hc = hv[f[:, 0]], hv[f[:, 1]], hv[f[:, 2]]
    pair = lambda x, y: _mix64(x * np.uint64(0x9E3779B97F4A7C15) + (y ^ np.uint64(0xD6E8FEB86659FD93)))
    hf = _mix64(pair(ha, hb) + pair(hb, hc) + pair(hc, ha))
    return (int(len(b)), int(hv.sum(dtype=np.uint64)), int(np.bitwise_xor.reduce(hv)) if len(hv) else 0,
            int(len(f)), int(hf.sum(dtype=np.uint64)), int(np.bitwise_xor.reduce(hf)) if len(hf) else 0)


def assert_same_mesh(got, want, normals_tol: float = 1e-5, what: str = ""):
  """Bit-exact canonical vertex and face sets; normals joined on vertex rows within tolerance."""
  gv, wv = canonical_vertices(got.vertices), canonical_vertices(want.vertices)
  assert gv.shape == wv.shape, f"{what}: vertex count {gv.shape} != {wv.shape}"
  assert np.array_equal(gv, wv), f"{what}: vertex sets differ"
  if len(gv) > 1:
    dup = np.all(gv[1:] == gv[:-1], axis=1)
    assert not dup.any(), f"{what}: duplicate vertices"
  gf, wf = canonical_faces(got.vertices, got.faces), canonical_faces(want.vertices, want.faces)
  assert gf.shape == wf.shape, f"{what}: face count {gf.shape} != {wf.shape}"
  assert np.array_equal(gf, wf), f"{what}: face sets differ"
  if want.normals is not None:
    assert got.normals is not None, f"{what}: normals missing"
    gb = np.ascontiguousarray(got.vertices, dtype=np.float32).view(np.uint32).reshape(-1, 3)
    wb = np.ascontiguousarray(want.vertices, dtype=np.float32).view(np.uint32).reshape(-1, 3)
    go = np.lexsort((gb[:, 2], gb[:, 1], gb[:, 0]))
    wo = np.lexsort((wb[:, 2], wb[:, 1], wb[:, 0]))
    gn, wn = np.asarray(got.normals)[go], np.asarray(want.normals)[wo]
    assert np.array_equal(np.isnan(gn), np.isnan(wn)), f"{what}: NaN masks of normals differ"
    ok = np.isnan(wn) | (np.abs(gn - wn) <= normals_tol)
    assert ok.all(), f"{what}: normals differ by up to {np.nanmax(np.abs(gn - wn))}"


def to_precomputed_bytes(vertices: np.ndarray, faces: np.ndarray) -> bytes:
  """Restatement of the reference's Mesh.to_precomputed (zmesh/mesh.py:257-269): uint32 Nv, then the float32 vertices
  in C order, then the uint32 faces in C order.  Pinned against the unmodified reference encoder by
  tests/golden/codec_golden.npz (tools/make_codec_golden.py)."""
  v = np.ascontiguousarray(vertices, dtype=np.float32)
  f = np.ascontiguousarray(faces).astype(np.uint32, copy=False)
  return b"".join((np.uint32(v.shape[0]).tobytes(), v.tobytes("C"), f.tobytes("C")))


# ------------------------------------------------------------------------------------------------
# deterministic synthetic volumes

_M64 = (1 << 64) - 1


def splitmix64(x):
  """Vectorised splitmix64 finaliser on uint64 arrays (wrap-around arithmetic)."""
  x = np.asarray(x, dtype=np.uint64)
  with np.errstate(over="ignore"):
    x = x + np.uint64(0x9E3779B97F4A7C15)
    z = x
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def voronoi_volume(shape, pitch: int, dtype=np.uint64, seed: int = 0, order: str = "F",
                   origin=(0, 0, 0), full_shape=None) -> np.ndarray:
  """Integer jittered-grid Voronoi segmentation.  Each grid cell of side
  `pitch` holds one site; a voxel takes the label of the nearest site among the 27 neighbouring
  cells, ties broken by the smaller cell id.  uint64 labels are splitmix64(c+1)|1, else c+1.
  `origin`/`full_shape` generate a sub-block of a larger volume (used by the slab tests)."""
  full_shape = tuple(full_shape or shape)
  G = [max(1, -(-s // pitch)) for s in full_shape]
  ci, cj, ck = np.meshgrid(np.arange(G[0]), np.arange(G[1]), np.arange(G[2]), indexing="ij")
  cid = (ci + G[0] * (cj + G[1] * ck)).astype(np.uint64)
  h = splitmix64(cid ^ np.uint64(seed))
  P = np.uint64(pitch)
  sx = ci * pitch + (((h & np.uint64(0xFFFF)) * P) >> np.uint64(16)).astype(np.int64)
  sy = cj * pitch + ((((h >> np.uint64(16)) & np.uint64(0xFFFF)) * P) >> np.uint64(16)).astype(np.int64)
  sz = ck * pitch + ((((h >> np.uint64(32)) & np.uint64(0xFFFF)) * P) >> np.uint64(16)).astype(np.int64)
  x = np.arange(shape[0], dtype=np.int64)[:, None, None] + origin[0]
  y = np.arange(shape[1], dtype=np.int64)[None, :, None] + origin[1]
  z = np.arange(shape[2], dtype=np.int64)[None, None, :] + origin[2]
  bi, bj, bk = x // pitch, y // pitch, z // pitch
  best_d = np.full(shape, np.iinfo(np.int64).max, dtype=np.int64)
  best_c = np.zeros(shape, dtype=np.int64)
  for dk in (-1, 0, 1):
    for dj in (-1, 0, 1):
      for di in (-1, 0, 1):
        ni, nj, nk = bi + di, bj + dj, bk + dk
        valid = (ni >= 0) & (ni < G[0]) & (nj >= 0) & (nj < G[1]) & (nk >= 0) & (nk < G[2])
        ni_c, nj_c, nk_c = np.clip(ni, 0, G[0] - 1), np.clip(nj, 0, G[1] - 1), np.clip(nk, 0, G[2] - 1)
        ni_c, nj_c, nk_c = np.broadcast_arrays(ni_c, nj_c, nk_c)
        d = (x - sx[ni_c, nj_c, nk_c]) ** 2 + (y - sy[ni_c, nj_c, nk_c]) ** 2 + (z - sz[ni_c, nj_c, nk_c]) ** 2
        c = ni_c + G[0] * (nj_c + G[1] * nk_c)
        valid = np.broadcast_to(valid, shape)
        better = valid & ((d < best_d) | ((d == best_d) & (c < best_c)))
        best_d = np.where(better, d, best_d)
        best_c = np.where(better, c, best_c)
  if np.dtype(dtype) == np.uint64:
    lab = splitmix64(best_c.astype(np.uint64) + np.uint64(1)) | np.uint64(1)
  else:
    lab = (best_c + 1).astype(dtype)
  return np.asarray(lab, dtype=dtype, order=order)


def voronoi_volume_c(shape, pitch: int, dtype=np.uint64, seed: int = 0, order: str = "F",
                     origin=(0, 0, 0), full_shape=None, threads: int = 0) -> np.ndarray:
  """voronoi_volume computed by the C restatement (oracle/zmesh_oracle.c:zo_voronoi), split over host
  threads along the slowest memory axis (ctypes releases the GIL).  Bit-identical to voronoi_volume and to
  the device generator; exists so that bench.py's CPU arms can build C4/C5-sized samples in seconds
  without loading the product library."""
  from concurrent.futures import ThreadPoolExecutor
  L = _port()
  shape = tuple(int(s) for s in shape)
  full_shape = tuple(int(s) for s in (full_shape or shape))
  out = np.empty(shape, dtype=np.dtype(dtype), order=order)
  c_order = 1 if order == "C" else 0
  a3 = lambda t: (C.c_uint64 * 3)(*[int(x) for x in t])
  sh, og, fs = a3(shape), a3(origin), a3(full_shape)
  ns = shape[0] if c_order else shape[2]
  nthreads = max(1, min(threads or (os.cpu_count() or 1), ns))
  cuts = [ns * k // nthreads for k in range(nthreads + 1)]
  nb = np.dtype(dtype).itemsize

  def work(k):
    L.zo_voronoi(C.c_void_p(out.ctypes.data), nb, sh, og, fs, int(pitch), int(seed), c_order, cuts[k], cuts[k + 1])
  with ThreadPoolExecutor(nthreads) as ex:
    list(ex.map(work, range(nthreads)))
  return out


def random_volume(shape, nlabels: int = 1000, dtype=np.uint32, seed: int = 0, order: str = "C") -> np.ndarray:
  """Config 3: default_rng(seed).integers(0, nlabels) (label 0 = background), cf. perf.py:68."""
  v = np.random.default_rng(seed).integers(0, nlabels, size=shape, dtype=dtype)
  return np.asarray(v, order=order)

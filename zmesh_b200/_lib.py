"""ctypes binding of libzmesh_b200.so (C ABI declared in include/zmesh_b200.h).

There is no fallback: if the CUDA library is missing or no sm_100a-capable device is usable the
import / first call fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ZMESH_B200_LIB") or os.path.join(HERE, "libzmesh_b200.so")  # (override: development A/B builds)
SOURCES = [os.path.join(HERE, "csrc", f) for f in ("zm_host.cu", "zm_kernels.cuh", "mc_tables.h")]
HEADER = os.path.join(os.path.dirname(HERE), "include", "zmesh_b200.h")

NVCC_FLAGS = [
  "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
  "-Xcompiler", "-fPIC", "-shared",
]


def build(force: bool = False, verbose: bool = False) -> str:
  """Compile the CUDA library in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
  srcs = SOURCES + [HEADER]
  if os.environ.get("ZMESH_B200_LIB"):  # a development build selected by hand is used as it is, never rebuilt
    return LIB_PATH
  if not force and os.path.exists(LIB_PATH):
    if all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in srcs if os.path.exists(s)):
      return LIB_PATH
  nvcc = os.environ.get("NVCC", "nvcc")
  cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH, SOURCES[0]]
  subprocess.check_call(cmd)
  return LIB_PATH


class zm_stats_t(C.Structure):
  _fields_ = [
    ("n_voxels", C.c_uint64), ("n_labels", C.c_uint64), ("n_vertices", C.c_uint64), ("n_faces", C.c_uint64),
    ("n_records", C.c_uint64), ("n_tiles", C.c_uint64), ("n_active_tiles", C.c_uint64), ("n_dense_tiles", C.c_uint64),
    ("hash_capacity", C.c_uint64), ("perm_capacity", C.c_uint64),
    ("attempts", C.c_uint32), ("launches", C.c_uint32), ("used_tma", C.c_uint32), ("launches_finalize", C.c_uint32),
    ("ms_h2d", C.c_float), ("ms_classify", C.c_float), ("ms_scan", C.c_float), ("ms_total", C.c_float),
    ("ms_faces", C.c_float), ("ms_vertices", C.c_float), ("ms_finalize", C.c_float), ("ms_exchange", C.c_float),
  ]


class zm_bulk_view(C.Structure):
  _fields_ = [
    ("n_labels", C.c_uint64), ("n_vertices", C.c_uint64), ("n_faces", C.c_uint64),
    ("labels_host", C.POINTER(C.c_uint64)), ("voff_host", C.POINTER(C.c_uint64)),
    ("foff_host", C.POINTER(C.c_uint64)),
    ("vertices_dev", C.c_void_p), ("faces_dev", C.c_void_p), ("normals_dev", C.c_void_p),
  ]


class zm_slab(C.Structure):
  _fields_ = [("full_extent", C.c_uint64), ("buf_lo", C.c_uint64), ("cube_lo", C.c_uint64), ("cube_hi", C.c_uint64),
              ("last", C.c_int)]


# every symbol include/zmesh_b200.h declares: name -> (restype, argtypes)
_f3 = C.POINTER(C.c_float)
_u64p = C.POINTER(C.c_uint64)
SYMBOLS = {
  "zm_create": (C.c_int, [_f3, C.c_int, C.POINTER(C.c_void_p)]),
  "zm_destroy": (None, [C.c_void_p]),
  "zm_set_resolution": (C.c_int, [C.c_void_p, _f3]),
  "zm_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
  "zm_wait_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
  "zm_synth_voronoi": (C.c_int, [C.c_void_p, C.c_int, _u64p, _u64p, _u64p, C.c_uint32, C.c_uint64, C.c_int, C.c_void_p]),
  "zm_mesh": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_int]),
  "zm_mesh_slab": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_int, C.POINTER(zm_slab)]),
  "zm_num_directory": (C.c_uint64, [C.c_void_p]),
  "zm_directory": (C.c_int, [C.c_void_p, _u64p, _u64p, _u64p, C.c_uint64]),
  "zm_set_label_offsets": (C.c_int, [C.c_void_p, _u64p, C.POINTER(C.c_uint32), C.c_uint64]),
  "zm_export_directory": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64]),
  "zm_import_directories": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint64]),
  "zm_plane_elems": (C.c_uint64, [C.c_void_p]),
  "zm_export_plane": (C.c_int, [C.c_void_p, C.c_void_p]),
  "zm_set_foreign_plane": (C.c_int, [C.c_void_p, C.c_void_p]),
  "zm_set_normal_plane": (C.c_int, [C.c_void_p, C.c_void_p]),
  "zm_add_normal_plane": (C.c_int, [C.c_void_p, C.c_void_p]),
  "zm_finish_normals": (C.c_int, [C.c_void_p]),
  "zm_finalize_begin": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, _f3]),
  "zm_num_ids": (C.c_uint64, [C.c_void_p]),
  "zm_ids": (C.c_int, [C.c_void_p, _u64p, C.c_uint64]),
  "zm_get_counts": (C.c_int, [C.c_void_p, C.c_uint64, _u64p, _u64p]),
  "zm_get": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_int, _f3, C.c_void_p, C.c_void_p, C.c_void_p]),
  "zm_erase": (C.c_int, [C.c_void_p, C.c_uint64, C.POINTER(C.c_int)]),
  "zm_clear": (C.c_int, [C.c_void_p]),
  "zm_finalize": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, _f3, C.POINTER(zm_bulk_view)]),
  "zm_fetch_all": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
  "zm_compute_normals": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p]),
  "zm_host_alloc": (C.c_void_p, [C.c_uint64]),
  "zm_host_free": (None, [C.c_void_p]),
  "zm_stats": (C.c_int, [C.c_void_p, C.POINTER(zm_stats_t)]),
  "zm_pack_precomputed": (C.c_int, [C.c_void_p, C.c_int, _f3, _u64p, _u64p]),
  "zm_fetch_precomputed": (C.c_int, [C.c_void_p, C.c_void_p, _u64p, _u64p]),
  "zm_nccl_unique_id": (C.c_int, [C.c_void_p]),
  "zm_comm_init": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
  "zm_comm_destroy": (C.c_int, [C.c_void_p]),
  "zm_slab_range": (C.c_int, [C.c_uint64, C.c_int, C.c_int, C.c_int, C.POINTER(zm_slab), _u64p, _u64p]),
  "zm_slab_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_int,
                             C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_int, _f3]),
  "zm_slab_finalize": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, _f3]),
  "zm_sync": (C.c_int, [C.c_void_p]),
  "zm_last_error": (C.c_char_p, [C.c_void_p]),
  "zm_version": (C.c_char_p, []),
}

_lib = None


def load():
  """Load the shared library and bind every declared symbol (raises if it is missing)."""
  global _lib
  if _lib is not None:
    return _lib
  if not os.path.exists(LIB_PATH):
    raise ImportError(
      f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
      "(zmesh_b200 has no CPU fallback)")
  lib = C.CDLL(LIB_PATH)
  for name, (res, args) in SYMBOLS.items():
    fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
    fn.restype = res
    fn.argtypes = args
  _lib = lib
  return lib

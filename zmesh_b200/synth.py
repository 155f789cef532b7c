"""Synthetic label volumes generated on the device (benchmark / test inputs)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

_TORCH_DTYPE = {1: "uint8", 2: "int16", 4: "int32", 8: "int64"}


def voronoi_device(shape, pitch: int, dtype=np.uint64, seed: int = 0, order: str = "F", origin=(0, 0, 0),
                   full_shape=None, device: int = 0):
  """Jittered-grid Voronoi segmentation written by a CUDA kernel into a new torch tensor whose
  logical shape is `shape` and whose memory order is `order` ("C" or "F").  Bit-identical to
  the numpy generator the tests use.  Returns (tensor, label_bytes); signed torch dtypes stand in for
  the unsigned ones torch lacks -- only the bit patterns matter."""
  import torch
  lib = _lib.load()
  nbytes = np.dtype(dtype).itemsize
  full_shape = tuple(full_shape or shape)
  tdt = getattr(torch, _TORCH_DTYPE[nbytes])
  dims = tuple(int(s) for s in shape)
  with torch.cuda.device(device):
    if order == "C":
      t = torch.empty(dims, dtype=tdt, device=f"cuda:{device}")
    else:
      t = torch.empty(dims[::-1], dtype=tdt, device=f"cuda:{device}").permute(2, 1, 0)
    a3 = lambda v: (C.c_uint64 * 3)(*[int(x) for x in v])
    stream = torch.cuda.current_stream().cuda_stream
    rc = lib.zm_synth_voronoi(C.c_void_p(t.data_ptr()), nbytes, a3(dims), a3(origin), a3(full_shape), int(pitch),
                              int(seed), 1 if order == "C" else 0, C.c_void_p(stream))
  if rc != 0:
    raise RuntimeError(f"zm_synth_voronoi failed ({rc})")
  return t

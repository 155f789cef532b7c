"""Multi-GPU slab sharding of the mesh+get path: one process per GPU (torch.distributed, NCCL over
NVLink), no reference counterpart (the reference is a single-threaded CPU library).

Sharding.  The volume is cut along its slowest memory axis (z for Fortran order, x for C order)
into `world` slabs of cube-origin planes; rank r reads its planes plus ONE halo plane on the high
side.  Every cube belongs to exactly one rank, so triangles are never duplicated.  Vertices are
owned by voxel (a vertex is a grid edge, owned by its lower voxel): the halo plane's slots belong
to rank r+1, so every vertex exists exactly once across ranks as well -- per-label partial meshes
CONCATENATE, there is nothing to dedup.  What has to be exchanged is only bookkeeping:

  1. all-gather of every rank's (label, n_vertices) directory -> rank r's face indices of label L
     are shifted by the number of L's vertices on ranks < r (`label_offsets`).  With a CUDA input and
     NCCL this runs entirely on the device (zm_export_directory -> all_gather_into_tensor on the
     mesher's stream -> zm_import_directories); the host-side functions below are the gloo / test path;
  2. rank r+1 sends rank r the final indices of the vertex slots in the shared plane
     (uint32 [Em][Efp][4], one NCCL send/recv between neighbours), which rank r's face kernel
     reads for the corners of its top cube layer.

  3. with normals: the faces of rank r's top cube layer touch vertices owned by rank r+1; their
     contributions are accumulated in a plane-shaped buffer (float32 [Em][Efp][4][3]), sent to rank r+1
     (one more neighbour send/recv), added there, and only then are the normals normalised.

The result stays distributed: rank r holds, for every label, the vertices it owns and the faces of
its cubes with global (cross-rank) indices; `gather_mesh(label)` concatenates the parts in rank
order, which is bit-identical (as canonical sets) to the single-GPU mesh.
"""
from __future__ import annotations

import numpy as np

from .mesh import Mesh
from .mesher import Mesher


def slab_planes(full_extent: int, close: bool, rank: int, world: int):
  """Cube-origin planes [cube_lo, cube_hi) of `rank` in EXTENDED coordinates (input plane + 1 when
  close) and the input planes [in_lo, in_hi) it has to read.  Returns (cube_lo, cube_hi, in_lo, in_hi, last)."""
  pad = 1 if close else 0
  ncube = full_extent + 2 * pad - 1  # cube-origin planes of the whole (extended) volume
  if world > ncube:
    raise ValueError("more ranks than cube planes")
  cube_lo = (ncube * rank) // world
  cube_hi = (ncube * (rank + 1)) // world
  in_lo = max(cube_lo - pad, 0)
  in_hi = min(cube_hi - pad, full_extent - 1) + 1
  return cube_lo, cube_hi, in_lo, in_hi, rank == world - 1


def offsets_from_directories(labels_by_rank, nv_by_rank, rank: int):
  """Index offset of each of rank's labels = its vertices on earlier ranks (pure numpy)."""
  mine = np.asarray(labels_by_rank[rank], dtype=np.uint64)
  if rank == 0 or mine.size == 0:
    return np.zeros(mine.size, dtype=np.uint32)
  prev_l = np.concatenate([np.asarray(labels_by_rank[q], dtype=np.uint64) for q in range(rank)])
  prev_n = np.concatenate([np.asarray(nv_by_rank[q], dtype=np.uint64) for q in range(rank)])
  if prev_l.size == 0:
    return np.zeros(mine.size, dtype=np.uint32)
  uniq, inv = np.unique(prev_l, return_inverse=True)
  tot = np.zeros(uniq.size, dtype=np.uint64)
  np.add.at(tot, inv, prev_n)
  pos = np.searchsorted(uniq, mine)
  pos_c = np.minimum(pos, uniq.size - 1)
  hit = uniq[pos_c] == mine
  out = np.where(hit, tot[pos_c], 0)
  if out.size and int(out.max()) >= 2 ** 32:
    raise ValueError("a label has more than 2^32-1 vertices")
  return out.astype(np.uint32)


_dir_capacity = 256  # rows of the fixed-size exchange buffer; grows to fit (every rank sees every size)


def all_gather_directories(labels, nv, group=None, device=None):
  """All-gather of variable-length (label, count) directories; works with gloo (CPU) and NCCL.
  One collective in the steady state: row 0 of every rank's fixed-capacity buffer carries its
  length; if any directory does not fit, all ranks see that and repeat with a larger buffer."""
  global _dir_capacity
  import torch
  import torch.distributed as dist
  world = dist.get_world_size(group)
  dev = device if device is not None else "cpu"
  n = int(labels.size)
  while True:
    cap = max(_dir_capacity, 1)
    host = np.zeros((cap + 1, 2), dtype=np.int64)
    host[0, 0] = n
    m = min(n, cap)
    if m:
      host[1:1 + m, 0] = labels[:m].astype(np.uint64).view(np.int64)
      host[1:1 + m, 1] = nv[:m].astype(np.uint64).view(np.int64)
    buf = torch.from_numpy(host).to(dev)
    outs = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(outs, buf, group=group)
    allh = torch.stack(outs).cpu().numpy()  # one device->host copy
    sizes = [int(allh[q, 0, 0]) for q in range(world)]
    if max(sizes) <= cap:
      break
    _dir_capacity = 1 << int(max(sizes) - 1).bit_length()
  ls, ns = [], []
  for q in range(world):
    a = allh[q, 1:1 + sizes[q]]
    ls.append(a[:, 0].copy().view(np.uint64))
    ns.append(a[:, 1].copy().view(np.uint64))
  return ls, ns


def assemble(parts):
  """Concatenate per-rank parts [(vertices, faces[, normals]) or None, ...] in rank order -> Mesh."""
  vs = [p[0] for p in parts if p is not None and len(p[0])]
  fs = [p[1] for p in parts if p is not None and len(p[1])]
  ns = [p[2] for p in parts if p is not None and len(p) > 2 and p[2] is not None and len(p[2])]
  v = np.concatenate(vs) if vs else np.zeros((0, 3), np.float32)
  f = np.concatenate(fs) if fs else np.zeros((0, 3), np.uint32)
  n = np.concatenate(ns) if ns and sum(len(x) for x in ns) == len(v) else None
  return Mesh(v, f, n)


class ShardedMesher:
  """One instance per rank.  mesh_slab() runs the rank's share of the path and the exchanges;
  afterwards the rank holds its parts of every label (device resident until fetched)."""

  def __init__(self, voxel_res, device: int, group=None, native: bool = True):
    """native: run the whole step in the C++ layer with its own NCCL communicators (zm_slab_step; torch.distributed
    is only used once, to hand out the communicator ids); False: the step is driven from Python over torch.distributed
    (the same kernels and exchanges, more host time per step)."""
    import torch.distributed as dist
    self.group = group
    self.rank = dist.get_rank(group)
    self.world = dist.get_world_size(group)
    self.device = int(device)
    self.mesher = Mesher(voxel_res, device=self.device)
    self.native = False
    if native and dist.get_backend(group) == "nccl":
      ids = ([Mesher.nccl_unique_id(), b"".join(Mesher.nccl_unique_id() for _ in range(self.world - 1))]
             if self.rank == 0 else [None, None])
      dist.broadcast_object_list(ids, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
      self.mesher.comm_init(ids[0], ids[1], self.world, self.rank)
      self.native = True
    self._plane_send = None
    self._plane_recv = None
    self._nplane_out = None
    self._nplane_in = None
    self._dir_cap = 1 << 15  # labels per shard the device-side directory exchange holds (doubles on overflow)
    self._dir_bufs = None

  def planes(self, full_extent: int, close: bool = False):
    return slab_planes(int(full_extent), bool(close), self.rank, self.world)

  def mesh_slab(self, data, full_extent: int, buf_lo: int, close: bool = False, finalize: bool = True,
                voxel_centered: bool = False, normals: bool = False):
    """data: this rank's planes [buf_lo, buf_lo + n) of the volume along the slab axis (numpy array or
    CUDA tensor), covering at least planes(full_extent, close)[2:4]."""
    import torch
    import torch.distributed as dist
    cube_lo, cube_hi, in_lo, in_hi, last = self.planes(full_extent, close)
    m = self.mesher
    if self.native:
      m.slab_step(data, full_extent, buf_lo, close=close, finalize=finalize, normals=normals, voxel_centered=voxel_centered)
      return m.finalize(normals=normals, voxel_centered=voxel_centered) if (finalize or normals) else None
    tm = self._timer()
    m.mesh_slab(data, full_extent, buf_lo, cube_lo, cube_hi, last, close=close)
    tm("pass1")
    dev = f"cuda:{self.device}"
    stream = torch.cuda.current_stream()
    same_stream = m.stream_handle() == int(stream.cuda_stream)
    on_device = hasattr(data, "__cuda_array_interface__") and dist.get_backend(self.group) == "nccl"
    if on_device:
      # device-side directory exchange: nothing crosses PCIe, nothing blocks the host
      cap = self._dir_cap
      if self._dir_bufs is None or self._dir_bufs[0].numel() != 2 * (1 + cap):
        self._dir_bufs = (torch.empty(2 * (1 + cap), dtype=torch.int64, device=dev),
                          torch.empty(self.world * 2 * (1 + cap), dtype=torch.int64, device=dev))
      mine, allb = self._dir_bufs
      m.export_directory(mine.data_ptr(), cap)
      if not same_stream:
        m.sync()
      dist.all_gather_into_tensor(allb, mine, group=self.group)
      if not same_stream:
        stream.synchronize()
      m.import_directories(allb.data_ptr(), self.world, self.rank, cap)
      tm("all_gather")
    else:
      labels, nv, nt = m.directory()
      ls, ns = all_gather_directories(labels, nv, self.group, dev)
      tm("all_gather")
      m.set_label_offsets(labels, offsets_from_directories(ls, ns, self.rank))
      tm("offsets")
    # boundary plane: rank r+1 -> rank r.  When the mesher queues its kernels on torch's current stream
    # (Mesher.set_stream, as bench.py does) the export kernel, the NCCL transfer and pass 2 are ordered
    # on the device and no host synchronisation is needed; pass 2 of every tile below the top tile layer
    # (finalize_begin) is queued BEFORE the stream waits for the transfer, so the transfer overlaps it.
    n = m.plane_elems()
    ops = []
    if self.rank > 0:
      if self._plane_send is None or self._plane_send.numel() != n:
        self._plane_send = torch.empty(n, dtype=torch.int32, device=dev)
      m.export_plane(self._plane_send.data_ptr())
      if not same_stream:
        m.sync()
      ops.append(dist.P2POp(dist.isend, self._plane_send, self.rank - 1, group=self.group))
    if not last:
      if self._plane_recv is None or self._plane_recv.numel() != n:
        self._plane_recv = torch.empty(n, dtype=torch.int32, device=dev)
      ops.append(dist.P2POp(dist.irecv, self._plane_recv, self.rank + 1, group=self.group))
    works = dist.batch_isend_irecv(ops) if ops else []
    want_pass2 = finalize or normals
    if want_pass2 and not last:
      if normals:
        if self._nplane_out is None or self._nplane_out.numel() != 3 * n:
          self._nplane_out = torch.empty(3 * n, dtype=torch.float32, device=dev)
        m.set_normal_plane(self._nplane_out.data_ptr())
      if same_stream:
        m.finalize_begin(normals=normals, voxel_centered=voxel_centered)
    for w in works:
      w.wait()  # makes the current stream wait for the transfer (no host block)
    if works and not same_stream:
      stream.synchronize()
    m.set_foreign_plane(self._plane_recv.data_ptr() if not last else None)
    tm("plane_exchange")
    out = None
    if want_pass2:
      try:
        out = m.finalize(normals=normals, voxel_centered=voxel_centered)
      except RuntimeError as e:
        if "exchange buffer" not in str(e):
          raise
        # (every rank checks every directory size, so every rank takes this branch) repeat the step with a larger buffer
        self._dir_cap *= 4
        return self.mesh_slab(data, full_extent, buf_lo, close=close, finalize=finalize, voxel_centered=voxel_centered,
                              normals=normals)
      tm("pass2")
      if normals:
        # normal contributions of the top cube layer to the next shard's first-plane vertices: rank r -> r+1
        ops = []
        if not last:
          ops.append(dist.P2POp(dist.isend, self._nplane_out, self.rank + 1, group=self.group))
        if self.rank > 0:
          if self._nplane_in is None or self._nplane_in.numel() != 3 * n:
            self._nplane_in = torch.empty(3 * n, dtype=torch.float32, device=dev)
          ops.append(dist.P2POp(dist.irecv, self._nplane_in, self.rank - 1, group=self.group))
        if ops:
          for w in dist.batch_isend_irecv(ops):
            w.wait()
          if not same_stream:
            stream.synchronize()
        if self.rank > 0:
          m.add_normal_plane(self._nplane_in.data_ptr())
        m.finish_normals()
        tm("normal_exchange")
    return out

  def _timer(self):
    """Per-phase wall times in self.timings (ms, accumulated) when ZM_SHARD_TIMING is set; each phase is
    closed with a device synchronisation, so this perturbs the overlap it measures."""
    import os
    if not os.environ.get("ZM_SHARD_TIMING"):
      return lambda name: None
    import time
    import torch
    self.timings = getattr(self, "timings", {})
    t = [time.perf_counter()]

    def mark(name):
      torch.cuda.synchronize()
      now = time.perf_counter()
      self.timings[name] = self.timings.get(name, 0.0) + (now - t[0]) * 1e3
      t[0] = now
    return mark

  def local_part(self, label, voxel_centered: bool = False, normals: bool = False):
    """(vertices, faces[, normals]) this rank holds for `label` (faces carry cross-rank indices) or None."""
    mesh = self.mesher.get(label, normals=normals, voxel_centered=voxel_centered)
    if len(mesh.vertices) == 0 and len(mesh.faces) == 0:
      return None
    return (mesh.vertices, mesh.faces, mesh.normals) if normals else (mesh.vertices, mesh.faces)

  def all_ids(self):
    """Sorted ids of the whole volume (every rank gets the same list)."""
    import torch.distributed as dist
    mine = self.mesher.ids()
    outs = [None] * self.world
    dist.all_gather_object(outs, mine, group=self.group)
    return sorted(set(i for o in outs for i in o))

  def gather_mesh(self, label, dst: int = 0, voxel_centered: bool = False, normals: bool = False):
    """Assemble the full mesh of `label` on rank `dst` (None elsewhere)."""
    import torch.distributed as dist
    part = self.local_part(label, voxel_centered, normals)
    outs = [None] * self.world if self.rank == dst else None
    dist.gather_object(part, outs, dst=dst, group=self.group)
    if self.rank != dst:
      return None
    mesh = assemble(outs)
    mesh.id = int(label)
    return mesh

#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path: Mesher.mesh(labels) + Mesher.get(id, rf=0) for
ALL labels, in megavoxels per second (BASELINE.json `metric`).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c5] [--impl ours|reference]

One process per GPU (torchrun for N > 1).  A "step" is one pass of the whole pipeline over the
workload volume: classify -> label scan -> emit -> final gather (-> normals).  Reported:

  value      device-resident throughput: labels already in HBM -> final per-label vertex / face
             (/ normal) arrays in HBM; CUDA events on the launching stream, max over ranks
  e2e        the same metric through the drop-in Python API with HOST buffers: mesher.mesh(numpy)
             + mesher.get(i) for every id (H2D of the volume and D2H of every mesh inside the
             timed region)
  roofline   dominant kernel: algorithmic bytes per launch / its measured duration vs HBM peak
  cpu_baseline  the reference's CPU implementation timed on this box's host cores on a bounded
             sample of the same workload (rank 0, N = 1)

Workloads (BASELINE.json `configs`): c1 connectomics.npy 512^3 u32; c2a zeros / c2b ones(close)
512^3 u32; c3 random [0,1000) 512^3 u32; c4 Voronoi 1024^3 u64 close+normals+voxel_centered;
c5 Voronoi 2048^3 u64 (default: the config the multi-GPU metric is quoted on; z-slab sharded
across ranks, strong scaling).
"""
import argparse
import gzip
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
  # name: (kind, shape, dtype, order, res, close, normals, voxel_centered, extra)
  "c1": dict(kind="connectomics", shape=(512, 512, 512), dtype="uint32", order="F", res=(4, 4, 40), close=False, normals=False, vc=False),
  "c2a": dict(kind="zeros", shape=(512, 512, 512), dtype="uint32", order="C", res=(4, 4, 40), close=False, normals=False, vc=False),
  "c2b": dict(kind="ones", shape=(512, 512, 512), dtype="uint32", order="C", res=(4, 4, 40), close=True, normals=False, vc=False),
  "c3": dict(kind="random", shape=(512, 512, 512), dtype="uint32", order="C", res=(4, 4, 40), close=False, normals=False, vc=False),
  "c4": dict(kind="voronoi", shape=(1024, 1024, 1024), dtype="uint64", order="F", res=(4, 4, 40), close=True, normals=True, vc=True, pitch=64),
  "c5": dict(kind="voronoi", shape=(2048, 2048, 2048), dtype="uint64", order="F", res=(4, 4, 40), close=False, normals=False, vc=False, pitch=128),
  # reduced-size stand-ins for development (never the default)
  "c5s": dict(kind="voronoi", shape=(512, 512, 512), dtype="uint64", order="F", res=(4, 4, 40), close=False, normals=False, vc=False, pitch=128),
  # all-zero volume of c5's size, filled on the device: every tile takes the uniform-region exit (cost floor of pass 1)
  "c5z": dict(kind="zeros_device", shape=(2048, 2048, 2048), dtype="uint64", order="F", res=(4, 4, 40), close=False, normals=False, vc=False),
}


def describe(name, wl, n_gpus=1):
  """`config` of the JSON line: names the workload and nothing run-specific, so the product arm and the
  reference arm print the same object (run-specific facts go to `details` / `cpu_baseline`)."""
  nb = np.dtype(wl["dtype"]).itemsize
  per_rank = int(np.prod(wl["shape"])) * nb / max(n_gpus, 1) / 1e9
  return {"workload": f"{name}: {wl['kind']} {'x'.join(map(str, wl['shape']))} {wl['dtype']} {wl['order']}-order"
                      + (f" pitch {wl['pitch']}" if 'pitch' in wl else ""),
          "resolution": list(wl["res"]), "close": wl["close"], "normals": wl["normals"],
          "voxel_centered": wl["vc"], "reduction_factor": 0,
          "l2": ("inputs larger than L2 (GPU arm: %.2f GB of labels per rank per step vs 126 MB of L2)" % per_rank
                 if per_rank > 0.3 else
                 "GPU arm: L2 flushed between timed steps (a 256 MB buffer is overwritten on the timing stream)"),
          "sharding": ("single GPU" if n_gpus <= 1 else
                       f"GPU arm: z-slabs with a 1-plane halo over {n_gpus} ranks, vertices owned by voxel (exactly once "
                       "across ranks), NCCL all-gather of per-label counts + neighbour send/recv of the boundary plane; "
                       "results stay distributed")}


class stdout_to_stderr:
  """NCCL prints its version banner to stdout when a communicator is created (NCCL_DEBUG=VERSION on some boxes): keep
  stdout for the one JSON line."""

  def __enter__(self):
    sys.stdout.flush()
    self.saved = os.dup(1)
    os.dup2(2, 1)

  def __exit__(self, *a):
    sys.stdout.flush()
    os.dup2(self.saved, 1)
    os.close(self.saved)


def load_peaks():
  p = os.path.join(ROOT, "MEASURED_PEAKS.json")
  if os.path.exists(p):
    try:
      return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
      pass
  return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
  """nvidia-smi clocks / throttle reasons DURING the timed region."""
  Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
       "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

  def __init__(self, index=0):
    self.index, self.rows, self.stop, self.t = index, [], threading.Event(), None

  def _run(self):
    while not self.stop.is_set():
      try:
        out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        if out:
          self.rows.append([c.strip() for c in out.splitlines()[0].split(",")])
      except Exception:
        pass
      self.stop.wait(0.2)

  def __enter__(self):
    self.t = threading.Thread(target=self._run, daemon=True)
    self.t.start()
    return self

  def __exit__(self, *a):
    self.stop.set()
    self.t.join(timeout=6)

  def summary(self):
    sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
    mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
    return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
            "reasons": reasons, "samples": len(self.rows)}


# ------------------------------------------------------------------------------------------------
# workload construction

def host_volume(name, wl, zrange=None):
  """Host numpy array of the workload (or of z-planes [z0, z1) of it)."""
  shape, dt = wl["shape"], np.dtype(wl["dtype"])
  if wl["kind"] == "connectomics":
    with gzip.open(os.path.join(ROOT, "tests", "golden", "connectomics.npy.gz"), "rb") as f:
      v = np.load(f)
  elif wl["kind"] == "zeros":
    v = np.zeros(shape, dtype=dt)
  elif wl["kind"] == "ones":
    v = np.ones(shape, dtype=dt)
  elif wl["kind"] == "random":
    v = np.random.default_rng(0).integers(0, 1000, size=shape, dtype=dt)
  else:
    raise ValueError("voronoi volumes are generated on the device")
  if zrange is not None:
    v = np.asarray(v[:, :, zrange[0]:zrange[1]], order=wl["order"])
  return v


def device_volume(name, wl, device, zrange=None):
  """torch CUDA tensor holding the workload (z-planes [z0,z1) of it when sharded)."""
  import torch
  from zmesh_b200.synth import voronoi_device
  shape = wl["shape"]
  z0, z1 = zrange if zrange is not None else (0, shape[2])
  if wl["kind"] == "voronoi":
    return voronoi_device((shape[0], shape[1], z1 - z0), wl["pitch"], np.dtype(wl["dtype"]), seed=0, order=wl["order"],
                          origin=(0, 0, z0), full_shape=shape, device=device)
  if wl["kind"] == "zeros_device":
    tdt = {1: torch.uint8, 2: torch.int16, 4: torch.int32, 8: torch.int64}[np.dtype(wl["dtype"]).itemsize]
    return torch.zeros((z1 - z0, shape[1], shape[0]), dtype=tdt, device=f"cuda:{device}").permute(2, 1, 0)
  v = host_volume(name, wl, zrange if zrange is not None else None)
  nb = v.dtype.itemsize
  sdt = {1: np.uint8, 2: np.int16, 4: np.int32, 8: np.int64}[nb]
  if v.flags.c_contiguous:
    return torch.from_numpy(v.view(sdt)).to(f"cuda:{device}")
  t = torch.from_numpy(np.ascontiguousarray(v.transpose(2, 1, 0)).view(sdt)).to(f"cuda:{device}")
  return t.permute(2, 1, 0)


# ------------------------------------------------------------------------------------------------

_REF = {}  # inherited by the forked reference workers

# planes of the workload (along its slowest memory axis) the CPU arms time: SURVEY.md 8(d) prescribes the
# 2048x2048x128 slab for C5; the others are sized for ~5-10 s of single-core reference work per pass.
CPU_SAMPLE_PLANES = {"c5": 128, "c4": 96, "c5s": 128, "c3": 24}
# rough single-core rates of the reference (MVx/s) used only to shrink the sample when many passes are asked for
CPU_RATE = {"c5": 90.0, "c4": 45.0, "c5s": 90.0, "c3": 0.9, "c1": 4.6, "c2a": 300.0, "c2b": 150.0, "c5z": 300.0}


def _ref_mesh_all(sample, wl, kind, digest=False):
  """The reference's own path on `sample`: Mesher.mesh + get for every id.  Returns (faces, {id: digest})."""
  from oracle import oracle as O
  m = O.OracleMesher(wl["res"], kind)
  m.mesh(sample, close=wl["close"])
  n, dig = 0, {}
  for i in m.ids():
    g = m.get(i, normals=wl["normals"], voxel_centered=wl["vc"])
    n += len(g.faces)
    if digest:
      dig[int(i)] = O.multiset_digest(g.vertices, g.faces)
  return n, dig


def _ref_worker(task):
  """One worker of the reference arm: the unmodified reference on planes [lo, hi) of the sample along its
  slowest memory axis."""
  lo, hi = task
  v, wl, kind = _REF["sample"], _REF["wl"], _REF["kind"]
  part = v[:, :, lo:hi] if v.flags.f_contiguous else v[lo:hi]
  return _ref_mesh_all(part, wl, kind)[0]


def cpu_sample(name, wl, seconds=8.0):
  """Bounded sample of the workload for the CPU arms (about `seconds` of single-core reference work per pass):
  the first planes of the volume along its slowest memory axis, as a standalone volume.  Built on the host
  (numpy / the oracle's C generator) -- never with the product library."""
  shape = wl["shape"]
  c_order = wl["order"] == "C"
  ns = shape[0] if c_order else shape[2]
  plane = int(np.prod(shape)) // ns
  want = int(CPU_RATE.get(name, 50.0) * 1e6 * seconds // plane) + 1
  nz = max(8, min(ns, CPU_SAMPLE_PLANES.get(name, ns), want))
  sub = (nz, shape[1], shape[2]) if c_order else (shape[0], shape[1], nz)
  what = f"planes [0,{nz}) of the slowest axis of the workload volume ({'x'.join(map(str, sub))}), meshed as a standalone volume"
  if wl["kind"] == "voronoi":
    from oracle.oracle import voronoi_volume_c
    v = voronoi_volume_c(sub, wl["pitch"], np.dtype(wl["dtype"]), 0, wl["order"], full_shape=shape)
  elif wl["kind"] == "zeros_device":
    v = np.zeros(sub, dtype=np.dtype(wl["dtype"]), order=wl["order"])
  else:
    full = host_volume(name, wl)
    v = np.asarray(full[:nz] if c_order else full[:, :, :nz], order=wl["order"])
  if nz == ns:
    what = "the full volume"
  return v, what


def run_reference(args, name, wl):
  """--impl reference: the reference's own CPU implementation (oracle/_ref when it was compiled
  from /root/reference, else the C port) timed on this box's host cores on a bounded sample.
  `value` is the reference exactly as it ships: one process, one thread (it has no threads, no SIMD
  and never releases the GIL, so one is all the threads it can use).  Beside it,
  `cpu_baseline.all_cores_value` gives what a production harness gets out of the box's cores by running
  one unmodified reference process per core on its own slab of the sample (one halo plane each, partial
  meshes NOT merged -- the reference has no such step); it is labelled as not a reference feature.
  Nothing of the product (zmesh_b200, its .so, torch) is imported on this arm."""
  import multiprocessing as mp
  from oracle import oracle as O
  O.build()
  kind = "reference" if O.have_reference() else "port"
  nsteps = args.steps + args.warmup
  sample, sample_desc = cpu_sample(name, wl, seconds=max(1.0, 170.0 / max(nsteps, 1)))
  axis_len = sample.shape[2] if sample.flags.f_contiguous else sample.shape[0]
  _REF.update(sample=sample, wl=wl, kind=kind)

  for _ in range(args.warmup):
    _ref_worker((0, axis_len))
  t0 = time.perf_counter()
  for _ in range(args.steps):
    nfaces = _ref_worker((0, axis_len))
  dt = (time.perf_counter() - t0) / args.steps
  mvx = sample.size / 1e6 / dt

  # all host cores: one reference process per core (fork: the sample is shared copy-on-write)
  ncores = max(1, min(os.cpu_count() or 1, 32))
  nproc = max(1, min(ncores, axis_len // 4))
  cuts = [axis_len * k // nproc for k in range(nproc + 1)]
  tasks = [(cuts[k], min(cuts[k + 1] + 1, axis_len)) for k in range(nproc)]  # +1: the halo plane of the slab's top cubes
  all_cores = None
  try:
    with mp.get_context("fork").Pool(nproc) as pool:
      pool.map(_ref_worker, [(0, min(4, axis_len))] * nproc, chunksize=1)  # start the workers
      t0 = time.perf_counter()
      pool.map(_ref_worker, tasks, chunksize=1)
      all_cores = sample.size / 1e6 / (time.perf_counter() - t0)
  except Exception as e:  # the single-core figure stands on its own
    print(f"all-cores figure skipped: {e}", file=sys.stderr)

  line = {
    "impl": "reference", "metric": "MVx/s mesh+get(rf=0)", "value": mvx, "unit": "MVx/s", "n_gpus": args.gpus,
    "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
    "scaling": "strong", "vs_baseline": None, "dtype": "u" + str(8 * np.dtype(wl["dtype"]).itemsize),
    "data": "synthetic" if wl["kind"] != "connectomics" else "connectomics.npy (reference sample volume)",
    "config": describe(name, wl, args.gpus),
    "cpu_baseline": {"value": mvx, "unit": "MVx/s", "cores": 1, "kind": kind, "sample": sample_desc, "faces": int(nfaces),
                     "host_cpus": os.cpu_count(), "all_cores_value": all_cores, "all_cores_processes": nproc,
                     "note": "value: the reference as it ships (single-threaded: no threads/SIMD/GIL release; 1 core is all it "
                             "can use), each step one pass over the sample, MVx/s = sample voxels / pass time.  "
                             "all_cores_value: NOT a reference feature -- one unmodified reference process per host "
                             "core, each meshing its own slab of the sample, partial meshes not merged"},
    "e2e": {"value": mvx, "unit": "MVx/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
  }
  print(json.dumps(line), flush=True)


def _canon(vertices, faces):
  """Canonical form of a mesh for the sharded-vs-single-GPU check (vertex rows sorted; faces as vertex rows,
  rotated to start at their smallest corner, sorted).  A product-vs-product comparison: no oracle involved."""
  b = np.ascontiguousarray(vertices, dtype=np.float32).view(np.uint32).reshape(-1, 3)
  cv = b[np.lexsort((b[:, 2], b[:, 1], b[:, 0]))]
  if len(faces) == 0:
    return cv.tobytes(), b""
  order = np.argsort(np.lexsort((b[:, 2], b[:, 1], b[:, 0])))  # rank of every vertex in the sorted list
  t = order[np.asarray(faces, dtype=np.int64)]
  k = np.argmin(t, axis=1)
  idx = (k[:, None] + np.arange(3)[None, :]) % 3
  r = t[np.arange(len(t))[:, None], idx]
  r = r[np.lexsort((r[:, 2], r[:, 1], r[:, 0]))]
  return cv.tobytes(), r.tobytes()


def nccl_parity_check(rank, world, local):
  """N-GPU exactness, run by every rank after the timed region: a 256x256x320 Voronoi volume meshed
  by the sharded path over NCCL must assemble, label by label, to exactly what ONE GPU produces for the whole
  volume (close=False, then close=True with normals).  Returns "ok" / "failed: ..." on rank 0."""
  import torch
  import torch.distributed as dist
  from zmesh_b200 import Mesher
  from zmesh_b200.sharded import ShardedMesher
  from zmesh_b200.synth import voronoi_device
  shape, pitch = (256, 256, 320), 40
  bad, nlab = 0, 0
  for close in (False, True):
    sm = ShardedMesher((4, 4, 40), device=local)
    cube_lo, cube_hi, in_lo, in_hi, last = sm.planes(shape[2], close)
    slab = voronoi_device((shape[0], shape[1], in_hi - in_lo), pitch, np.uint64, seed=5, order="F",
                          origin=(0, 0, in_lo), full_shape=shape, device=local)
    torch.cuda.synchronize()
    sm.mesh_slab(slab, shape[2], in_lo, close=close, normals=close)
    ids = sm.all_ids()
    ref = None
    if rank == 0:
      full = voronoi_device(shape, pitch, np.uint64, seed=5, order="F", device=local)
      torch.cuda.synchronize()
      ref = Mesher((4, 4, 40), device=local)
      ref.mesh(full, close=close)
      if ref.ids() != ids:
        bad += 1
    for lbl in ids:
      got = sm.gather_mesh(lbl, dst=0, normals=close)
      if rank == 0:
        want = ref.get(lbl, normals=close)
        nlab += 1
        if _canon(got.vertices, got.faces) != _canon(want.vertices, want.faces):
          bad += 1
        elif close:  # normals joined on the vertex rows: 1e-5 absolute, equal NaN masks
          gb = np.ascontiguousarray(got.vertices).view(np.uint32).reshape(-1, 3)
          wb = np.ascontiguousarray(want.vertices).view(np.uint32).reshape(-1, 3)
          gn = np.asarray(got.normals)[np.lexsort((gb[:, 2], gb[:, 1], gb[:, 0]))]
          wn = np.asarray(want.normals)[np.lexsort((wb[:, 2], wb[:, 1], wb[:, 0]))]
          if not (np.array_equal(np.isnan(gn), np.isnan(wn)) and bool((np.isnan(wn) | (np.abs(gn - wn) <= 1e-5)).all())):
            bad += 1
    del sm, slab
  flag = torch.tensor([bad, nlab], device=f"cuda:{local}")
  dist.broadcast(flag, 0)
  bad, nlab = int(flag[0]), int(flag[1])
  return "ok" if bad == 0 and nlab > 0 else f"failed: {bad} of {nlab} label checks differ"


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=5)
  ap.add_argument("--warmup", type=int, default=3)
  ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
  ap.add_argument("--workload", default=os.environ.get("ZM_BENCH_WORKLOAD", "c5"), choices=sorted(WORKLOADS))
  ap.add_argument("--e2e-steps", type=int, default=2)
  ap.add_argument("--no-e2e", action="store_true")
  ap.add_argument("--no-cpu", action="store_true")
  ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the sharded-vs-single-GPU exactness check after the timed region")
  args = ap.parse_args()
  args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)
  name, wl = args.workload, WORKLOADS[args.workload]

  rank = int(os.environ.get("RANK", "0"))
  world = int(os.environ.get("WORLD_SIZE", "1"))
  local_rank = int(os.environ.get("LOCAL_RANK", "0"))

  if args.impl == "reference":
    if rank == 0:
      run_reference(args, name, wl)
    return

  import torch
  import torch.distributed as dist
  import __graft_entry__ as G
  if rank == 0:
    G.build()
  if world > 1:
    torch.cuda.set_device(local_rank)
    with stdout_to_stderr():
      dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
      dist.barrier()
      torch.cuda.synchronize()
  from zmesh_b200 import Mesher

  dev = local_rank
  torch.cuda.set_device(dev)
  shape = wl["shape"]
  nvox_total = int(np.prod(shape))
  label_bytes = np.dtype(wl["dtype"]).itemsize
  sm = None
  if world > 1:
    # z-slab sharding (zmesh_b200/sharded.py): rank r reads its cube planes + one halo plane
    if wl["order"] != "F":
      raise SystemExit("multi-GPU bench: Fortran-order workloads only (c1, c4, c5, c5s)")
    from zmesh_b200.sharded import ShardedMesher
    with stdout_to_stderr():  # (the native step creates its own NCCL communicators)
      sm = ShardedMesher(wl["res"], device=dev)
      torch.cuda.synchronize()
    _, _, in_lo, in_hi, _ = sm.planes(shape[2], wl["close"])
    zr = (in_lo, in_hi)
    mesher = sm.mesher
  else:
    zr = (0, shape[2])
    mesher = Mesher(wl["res"], device=dev)
  vol = device_volume(name, wl, dev, zr if world > 1 else None)
  torch.cuda.synchronize()
  # a dedicated (non-default) stream shared by the mesher's kernels and NCCL: the multi-GPU exchanges
  # are then ordered on the device without host synchronisation
  stream = torch.cuda.Stream(device=dev)
  torch.cuda.set_stream(stream)
  mesher.set_stream(stream.cuda_stream)

  def step():
    if world > 1:
      sm.mesh_slab(vol, shape[2], zr[0], close=wl["close"], finalize=True, voxel_centered=wl["vc"], normals=wl["normals"])
    else:
      mesher.mesh(vol, close=wl["close"])
      mesher.finalize(normals=wl["normals"], voxel_centered=wl["vc"])
    return mesher.stats()

  def barrier():
    torch.cuda.synchronize()
    if world > 1:
      dist.barrier()
      torch.cuda.synchronize()

  for _ in range(args.warmup):
    st = step()
  barrier()
  if sm is not None and getattr(sm, "timings", None):
    sm.timings.clear()  # (ZM_SHARD_TIMING diagnostics: steady state only)
  acc = {k: 0.0 for k in ("ms_classify", "ms_scan", "ms_faces", "ms_vertices", "ms_total", "ms_finalize", "ms_exchange")}
  launches = 0
  ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  with ClockSampler(dev) as clocks:
    barrier()
    ev0.record(stream)
    host_ms = 0.0
    for _ in range(args.steps):
      t_h = time.perf_counter()
      st = step()
      host_ms += (time.perf_counter() - t_h) * 1e3
      for k in acc:
        acc[k] += st[k]
      launches += st["launches"] + st["launches_finalize"]
    ev1.record(stream)
    barrier()
  ms = ev0.elapsed_time(ev1) / args.steps
  if world > 1:
    t = torch.tensor([ms], device=f"cuda:{dev}", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    cnt = torch.tensor([st["n_vertices"], st["n_faces"], st["n_labels"]], device=f"cuda:{dev}", dtype=torch.int64)
    dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    # per-rank kernel times (the step is paced by the slowest rank of every phase: the exchanges couple the ranks)
    per_rank = [None] * world
    dist.all_gather_object(per_rank, {"k_classify": acc["ms_classify"] / args.steps,
                                      "k_emit": (acc["ms_faces"] + acc["ms_vertices"]) / args.steps,
                                      "exchange": acc["ms_exchange"] / args.steps, "mesh_call": acc["ms_total"] / args.steps,
                                      "host_step_ms": host_ms / args.steps,
                                      "non_empty_tiles": int(st["n_active_tiles"])})
    totV, totT, nlabels_total = int(cnt[0]), int(cnt[1]), int(cnt[2])  # (labels: summed over ranks, a label spanning k slabs counts k times)
  else:
    totV, totT, nlabels_total = st["n_vertices"], st["n_faces"], st["n_labels"]
  value = nvox_total / 1e6 / (ms / 1e3)

  # ---- roofline of the dominant kernel (rank-local averages) -----------------------------------
  peak, peak_src = load_peaks()
  nvox_local = int(np.prod(vol.shape))
  V, T = st["n_vertices"], st["n_faces"]
  # algorithmic bytes per launch: the volume for the classification pass; 12 B per
  # triangle and 12 B per vertex (+12 B with normals) for the emit pass
  kern = {
    "k_classify": (acc["ms_classify"] / args.steps, nvox_local * label_bytes),
    # pass 2 is one fused kernel: 12 B per triangle + 12 B per vertex (+12 B per vertex for normals, which
    # also adds the small normalisation kernel timed under ms_vertices)
    "k_emit": ((acc["ms_faces"] + acc["ms_vertices"]) / args.steps, 12 * T + 12 * V + (12 * V if wl["normals"] else 0)),
  }
  dom = max(kern, key=lambda k: kern[k][0])
  dom_ms, dom_bytes = kern[dom]
  achieved = dom_bytes / 1e9 / (dom_ms / 1e3) if dom_ms > 0 else 0.0
  # DRAM bytes per launch of the dominant kernel from the ncu capture of the CURRENT kernels kept in
  # profiles/traffic.json (keyed by workload; single-GPU launches only -- a rank-local slab moves 1/N of it)
  traffic = None
  tp = os.path.join(ROOT, "profiles", "traffic.json")
  if world == 1 and os.path.exists(tp):
    try:
      traffic = json.load(open(tp)).get(name, {}).get(dom)
    except Exception:
      traffic = None
  b_alg = nvox_local * label_bytes + 12 * V + 12 * T + (12 * V if wl["normals"] else 0)
  roofline = {
    "bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
    "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": dom_bytes,
    "kernel_ms": {**{k: v[0] for k, v in kern.items()}, "scan": acc["ms_scan"] / args.steps},
    "kernel_frac": {k: (v[1] / 1e9 / (v[0] / 1e3) / peak if v[0] > 0 else None) for k, v in kern.items()},
    "pipeline": {"algorithmic_bytes": b_alg, "ms": ms, "achieved": b_alg / 1e9 / (ms / 1e3),
                 "frac": b_alg / 1e9 / (ms / 1e3) / peak,
                 "note": "SURVEY 8d: N*sizeof(label) + 12 B/vertex + 12 B/triangle (+12 B/vertex normals), rank-local"},
  }

  # ---- end to end through the drop-in API with host buffers ------------------------------------
  e2e = None
  if not args.no_e2e:
    mesher.set_stream(None)
    host = vol.cpu()  # rank-local slab, host memory
    hnp = host.numpy() if not isinstance(host, np.ndarray) else host
    hnp = hnp.view(np.dtype(wl["dtype"]))
    del vol
    torch.cuda.empty_cache()
    try:
      pinned = torch.cuda.cudart().cudaHostRegister(hnp.ctypes.data, hnp.nbytes, 0)
    except Exception:
      pinned = None
    d2h = 0
    # what the platform gives a plain pinned copy while every rank copies at once (the e2e figure is PCIe-bound:
    # this is its ceiling; with 8 ranks the host side -- shared PCIe switches / memory -- sets it, not the GPUs)
    h2d_plain = None
    try:
      import ctypes
      rt = ctypes.CDLL("libcudart.so.12")  # (the runtime torch has loaded; cudaMemcpy from the REGISTERED numpy buffer)
      rt.cudaMemcpy.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
      nprobe = int(min(hnp.nbytes, 2 << 30))
      dst = torch.empty(nprobe, dtype=torch.uint8, device=f"cuda:{dev}")
      if pinned is not None and rt.cudaMemcpy(dst.data_ptr(), hnp.ctypes.data, nprobe, 1) == 0:
        barrier()
        t0 = time.perf_counter()
        for _ in range(3):
          rt.cudaMemcpy(dst.data_ptr(), hnp.ctypes.data, nprobe, 1)
        h2d_plain = 3 * nprobe / 1e9 / (time.perf_counter() - t0)
      del dst
      torch.cuda.empty_cache()
    except Exception:
      pass

    phase = {"mesh_call_ms": 0.0, "get_loop_ms": 0.0, "h2d_ms": 0.0, "classify_ms": 0.0, "emit_ms": 0.0}

    def e2e_step():
      nonlocal d2h
      t_a = time.perf_counter()
      if world > 1:
        sm.mesh_slab(hnp, shape[2], zr[0], close=wl["close"], finalize=wl["normals"], voxel_centered=wl["vc"],
                     normals=wl["normals"])
      else:
        mesher.mesh(hnp, close=wl["close"])
      t_b = time.perf_counter()
      d2h = 0
      for i in mesher.ids():
        m = mesher.get(i, normals=wl["normals"], reduction_factor=0, voxel_centered=wl["vc"])
        d2h += m.vertices.nbytes + m.faces.nbytes + (m.vertices.nbytes if wl["normals"] else 0)
      t_c = time.perf_counter()
      s2 = mesher.stats()
      phase["mesh_call_ms"] += (t_b - t_a) * 1e3
      phase["get_loop_ms"] += (t_c - t_b) * 1e3  # first get(): pass 2 on the device + ONE bulk D2H; the rest are host slices
      phase["h2d_ms"] += s2["ms_h2d"]
      phase["classify_ms"] += s2["ms_classify"] + s2["ms_scan"]
      phase["emit_ms"] += s2["ms_finalize"]

    e2e_step()  # warm-up (allocations, page faults)
    barrier()
    for k in phase:
      phase[k] = 0.0
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
      e2e_step()
    barrier()
    dt = (time.perf_counter() - t0) / args.e2e_steps
    if world > 1:
      t = torch.tensor([dt], device=f"cuda:{dev}", dtype=torch.float64)
      dist.all_reduce(t, op=dist.ReduceOp.MAX)
      dt = float(t.item())
    e2e = {"value": nvox_total / 1e6 / dt, "unit": "MVx/s", "h2d_bytes_per_step": int(hnp.nbytes),
           "d2h_bytes_per_step": int(d2h), "ms_per_step": dt * 1e3, "steps": args.e2e_steps,
           "host_buffer": "cudaHostRegister'ed numpy" if pinned is not None else "pageable numpy",
           "phases_ms": {k: v / args.e2e_steps for k, v in phase.items()},
           "pcie_gbs": {"h2d": (hnp.nbytes / 1e9) / (phase["h2d_ms"] / args.e2e_steps / 1e3) if phase["h2d_ms"] > 0 else None,
                        "h2d_plain_copy_all_ranks_at_once": h2d_plain,
                        "note": "rank 0; h2d = the volume upload inside mesh(); plain copy = cudaMemcpy of 2 GiB of the same "
                                "registered buffer while every rank does the same (the platform's ceiling for this step)"},
           "api": "zmesh_b200.Mesher.mesh(ndarray) + get(id) for every id (rank-local slab when sharded)"}

  # ---- CPU baseline beside it (rank 0, N = 1): the compiled reference on one host core, two passes over a
  #      bounded sample (the second, warm one is reported), and the SAME sample through the product path with
  #      per-label fingerprints compared (bit-exact vertex / face sets; oracle.multiset_digest) ---------------
  cpu = None
  if rank == 0 and world == 1 and not args.no_cpu:
    from oracle import oracle as O
    kind = "reference" if O.have_reference() else "port"
    sample, sdesc = cpu_sample(name, wl)
    secs = []
    for _ in range(2):
      t0 = time.perf_counter()
      nfaces, _ = _ref_mesh_all(sample, wl, kind)
      secs.append(time.perf_counter() - t0)
    _, want = _ref_mesh_all(sample, wl, kind, digest=True)
    mesher.set_stream(None)
    mesher.mesh(sample, close=wl["close"])
    got_ids = [int(i) for i in mesher.ids()]
    mism = 0 if sorted(got_ids) == sorted(want) else 1
    for i in got_ids:
      g = mesher.get(i, normals=False, voxel_centered=wl["vc"])
      if O.multiset_digest(g.vertices, g.faces) != want.get(i):
        mism += 1
    cpu = {"value": sample.size / 1e6 / secs[1], "unit": "MVx/s", "cores": 1, "kind": kind, "sample": sdesc,
           "seconds": secs[1], "seconds_cold_pass": secs[0], "passes": 2, "faces": int(nfaces), "host_cpus": os.cpu_count(),
           "parity_on_sample": mism == 0, "parity_labels": len(want), "parity_mismatches": mism,
           "parity_note": "the sample meshed by the GPU path through the drop-in API; per-label fingerprints of the "
                          "vertex and face sets (rotation-invariant, winding-sensitive) equal to the compiled reference's"}

  nccl_parity = None
  if world > 1 and not args.no_parity:
    torch.cuda.set_stream(torch.cuda.default_stream(dev))
    with stdout_to_stderr():
      nccl_parity = nccl_parity_check(rank, world, local_rank)

  if rank == 0:
    line = {
      "metric": "MVx/s mesh+get(rf=0)", "value": value, "unit": "MVx/s", "n_gpus": world, "steps": args.steps,
      "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
      "dtype": "u" + str(8 * label_bytes),
      "data": "synthetic" if wl["kind"] != "connectomics" else "connectomics.npy (reference sample volume)",
      "config": describe(name, wl, world),
      "details": {"labels": int(nlabels_total), "vertices": int(totV), "faces": int(totT), "volume_gb_per_rank": nvox_local * label_bytes / 1e9,
                  "tiles": {"all": int(st["n_tiles"]), "non_empty": int(st["n_active_tiles"]), "dense_redo": int(st["n_dense_tiles"])}},
      "parity_on_sample": (cpu or {}).get("parity_on_sample"), "nccl_parity": nccl_parity,
      "per_rank": per_rank if world > 1 else None,
      "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
      "clocks": clocks.summary(),
    }
    print(json.dumps(line), flush=True)
  if world > 1:
    if rank == 0 and getattr(sm, "timings", None):
      print("shard phase times (ms, summed over all mesh_slab calls):", {k: round(v, 2) for k, v in sm.timings.items()}, file=sys.stderr)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
  main()

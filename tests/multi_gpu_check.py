"""N-GPU exactness check (run under torchrun, one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py
Every rank meshes its z-slab with ShardedMesher (NCCL exchanges); rank 0 assembles every label and
compares the canonical digest with a single-GPU mesh of the whole volume computed on rank 0."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist


def main():
  rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
  torch.cuda.set_device(local)
  dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
  from zmesh_b200 import Mesher
  from zmesh_b200.sharded import ShardedMesher
  from zmesh_b200.synth import voronoi_device
  from oracle.oracle import assert_same_mesh, canonical_digest  # test infrastructure: the checker

  shape, pitch = (256, 256, 320), 40
  ok = True
  # (close, native): the native NCCL step of the C++ layer, and the same step driven from Python over torch.distributed
  for close, native in ((False, True), (True, True), (False, False), (True, False)):
    sm = ShardedMesher((4, 4, 40), device=local, native=native)
    cube_lo, cube_hi, in_lo, in_hi, last = sm.planes(shape[2], close)
    slab = voronoi_device((shape[0], shape[1], in_hi - in_lo), pitch, np.uint64, seed=5, order="F",
                          origin=(0, 0, in_lo), full_shape=shape, device=local)
    torch.cuda.synchronize()
    sm.mesh_slab(slab, shape[2], in_lo, close=close, normals=close)  # (normals on the close=True pass)
    ids = sm.all_ids()
    ref = None
    if rank == 0:
      full = voronoi_device(shape, pitch, np.uint64, seed=5, order="F", device=local)
      torch.cuda.synchronize()
      ref = Mesher((4, 4, 40), device=local)
      ref.mesh(full, close=close)
      assert ref.ids() == ids, (len(ref.ids()), len(ids))
    bad = 0
    for lbl in ids:
      got = sm.gather_mesh(lbl, dst=0, normals=close)
      if rank == 0:
        want = ref.get(lbl, normals=close)
        if canonical_digest(got.vertices, got.faces) != canonical_digest(want.vertices, want.faces):
          bad += 1
        elif close:
          try:
            assert_same_mesh(got, want, 1e-5, what=f"label {lbl}")
          except AssertionError as e:
            print(e, flush=True)
            bad += 1
    if rank == 0:
      print(f"close={close} native={native}: {len(ids)} labels over {world} ranks, {bad} mismatches", flush=True)
      ok = ok and bad == 0 and len(ids) > 0
  flag = torch.tensor([1 if ok else 0], device=f"cuda:{local}")
  dist.broadcast(flag, 0)
  dist.barrier()
  dist.destroy_process_group()
  if rank == 0:
    print("MULTI_GPU_CHECK", "OK" if ok else "FAILED", flush=True)
  sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
  main()

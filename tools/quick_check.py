"""Small GPU sanity run used during kernel development (also a compute-sanitizer target)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle.oracle import OracleMesher, assert_same_mesh, voronoi_volume, random_volume
from zmesh_b200 import Mesher

def run(vol, res, close, name):
  gpu = Mesher(res, device=0); gpu.mesh(vol, close=close)
  cpu = OracleMesher(res, "port"); cpu.mesh(vol, close=close)
  st = gpu.stats()
  print(name, {k: st[k] for k in ("n_labels", "n_vertices", "n_faces", "n_records", "n_active_tiles", "n_dense_tiles", "attempts", "used_tma")}, flush=True)
  assert gpu.ids() == sorted(cpu.ids()), (len(gpu.ids()), len(cpu.ids()))
  for lbl in gpu.ids():
    assert_same_mesh(gpu.get(lbl, normals=True, voxel_centered=True), cpu.get(lbl, normals=True, voxel_centered=True), 1e-5, what=f"{name}:{lbl}")
    assert_same_mesh(gpu.get_mesh(lbl, normals=True), cpu.get_mesh(lbl, normals=True), 1e-5, what=f"{name}:legacy:{lbl}")
  print(name, "ok", flush=True)

run(voronoi_volume((48, 40, 36), 12, np.uint64, seed=1, order="F"), (4, 4, 40), True, "voronoi_u64_F_close_tma")
run(voronoi_volume((64, 40, 36), 12, np.uint32, seed=1, order="C"), (4, 4, 40), False, "voronoi_u32_C_tma")
run(random_volume((33, 9, 9), 3, np.uint8, 5, "F"), (1, 1, 1), False, "tile_edges_generic")
run(random_volume((40, 41, 43), 300, np.uint16, 1, "F"), (1, 1, 1), True, "random_u16_dense")
run(random_volume((64, 32, 32), 1000, np.uint32, 0, "C"), (0.1, 3.3, 7.77), False, "random_u32_dense_tma")

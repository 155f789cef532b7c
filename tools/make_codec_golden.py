"""Generates tests/golden/codec_golden.npz with the UNMODIFIED reference encoders (zmesh/mesh.py:257-269 to_precomputed,
:348-376 to_ply, :321-346 to_obj) on seeded meshes.  Run where /root/reference exists; the fixture travels to the GPU box.
    python tools/make_codec_golden.py
"""
import importlib.util
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/zmesh/mesh.py"


def reference_mesh_class():
  spec = importlib.util.spec_from_file_location("zmesh_reference_mesh", REF)
  mod = importlib.util.module_from_spec(spec)
  spec.loader.exec_module(mod)
  return mod.Mesh


def cases():
  rng = np.random.default_rng(7)
  out = {}
  for name, nv, nf, scale in (("tiny", 4, 2, 1.0), ("mid", 257, 511, 1000.0), ("halfvoxel", 100, 64, 0.5)):
    v = (rng.integers(0, 4096, size=(nv, 3)).astype(np.float32) * np.float32(scale) / np.float32(2.0)).astype(np.float32)
    f = rng.integers(0, nv, size=(nf, 3)).astype(np.uint32)
    out[name] = (v, f)
  out["empty"] = (np.zeros((0, 3), np.float32), np.zeros((0, 3), np.uint32))
  return out


def main():
  Mesh = reference_mesh_class()
  blob = {}
  for name, (v, f) in cases().items():
    m = Mesh(v, f, None)
    blob[f"{name}/v"], blob[f"{name}/f"] = v, f
    blob[f"{name}/precomputed"] = np.frombuffer(m.to_precomputed(), dtype=np.uint8)
    blob[f"{name}/ply"] = np.frombuffer(bytes(m.to_ply()), dtype=np.uint8)
    blob[f"{name}/obj"] = np.frombuffer(m.to_obj() if isinstance(m.to_obj(), bytes) else m.to_obj().encode("utf8"), dtype=np.uint8)
  # host-side clean-up utilities (zmesh/mesh.py:117-226) on a seeded mesh with duplicate vertices, degenerate and
  # repeated faces
  rng = np.random.default_rng(11)
  v = rng.integers(0, 4, size=(40, 3)).astype(np.float32)
  f = rng.integers(0, 40, size=(90, 3)).astype(np.uint32)
  n = rng.random((40, 3)).astype(np.float32)
  blob["messy/v"], blob["messy/f"], blob["messy/n"] = v, f, n
  for op in ("remove_unreferenced_vertices", "remove_degenerate_faces", "consolidate"):
    r = getattr(Mesh(v, f, n), op)()
    blob[f"messy/{op}/v"], blob[f"messy/{op}/f"] = r.vertices, r.faces
    blob[f"messy/{op}/n"] = np.zeros((0, 3), np.float32) if r.normals is None else r.normals
  v2 = (rng.random((40, 3)) * 3).astype(np.float32)
  r = Mesh(v2, f, None).merge_close_vertices(0.8)
  blob["messy/v2"], blob["messy/merge/v"], blob["messy/merge/f"] = v2, r.vertices, r.faces
  np.savez_compressed(os.path.join(ROOT, "tests", "golden", "codec_golden.npz"), **blob)
  print("wrote", len(blob), "arrays")


if __name__ == "__main__":
  main()

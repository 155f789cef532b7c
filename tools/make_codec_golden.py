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
  np.savez_compressed(os.path.join(ROOT, "tests", "golden", "codec_golden.npz"), **blob)
  print("wrote", len(blob), "arrays")


if __name__ == "__main__":
  main()

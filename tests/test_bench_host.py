"""Host-side pieces of bench.py that do not need a GPU: both arms print the same `config`, the CPU sample is built
without the product library, the canonical form used by the N-GPU exactness check."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_config_object_is_the_same_on_both_arms():
  import bench
  for name, wl in bench.WORKLOADS.items():
    for n in (1, 8):
      a, b = bench.describe(name, wl, n), bench.describe(name, wl, n)
      assert a == b and a["workload"].startswith(name + ":") and "l2" in a and "sharding" in a
      assert not any(k in a for k in ("model", "sample", "labels", "tiles"))  # run-specific facts live elsewhere


def test_cpu_sample_is_a_slab_of_the_workload():
  import bench
  from oracle.oracle import voronoi_volume_c
  wl = bench.WORKLOADS["c5s"]
  v, what = bench.cpu_sample("c5s", wl, seconds=0.05)
  assert v.dtype == np.uint64 and v.flags.f_contiguous and v.shape[:2] == (512, 512) and 8 <= v.shape[2] <= 128
  assert "planes [0," in what
  full = voronoi_volume_c((512, 512, v.shape[2]), wl["pitch"], np.uint64, 0, "F", full_shape=wl["shape"])
  assert np.array_equal(v, full)
  z, what = bench.cpu_sample("c2a", bench.WORKLOADS["c2a"], seconds=100.0)
  assert z.shape == (512, 512, 512) and what == "the full volume" and not z.any()


def test_reference_arm_never_loads_the_product_library():
  """`bench.py --impl reference` (here on a tiny workload override) must not import zmesh_b200, torch or load
  libzmesh_b200.so: the arm is the unmodified reference (or its C port) alone."""
  code = (
    "import sys, json; sys.argv=['bench.py','--impl','reference','--workload','c5s','--steps','1','--warmup','0'];"
    "import bench; bench.CPU_SAMPLE_PLANES['c5s']=8; bench.main();"
    "bad=[m for m in sys.modules if m.split('.')[0] in ('zmesh_b200','torch')];"
    "maps=open('/proc/self/maps').read(); print('LOADED', bad, 'libzmesh_b200' in maps)")
  r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=600)
  assert r.returncode == 0, r.stderr[-2000:]
  lines = r.stdout.strip().splitlines()
  assert lines[-1] == "LOADED [] False", lines[-1]
  import json
  line = json.loads(lines[-2])
  assert line["impl"] == "reference" and line["cpu_baseline"]["cores"] == 1 and line["e2e"]["h2d_bytes_per_step"] == 0
  assert line["config"]["workload"].startswith("c5s:")


def test_canonical_form_of_the_nccl_parity_check():
  import bench
  v = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], dtype=np.float32)
  f = np.array([[0, 1, 2], [1, 3, 2]], dtype=np.uint32)
  perm = np.array([2, 0, 3, 1])
  inv = np.argsort(perm)
  f2 = inv[f][:, [1, 2, 0]].astype(np.uint32)[::-1]
  assert bench._canon(v, f) == bench._canon(v[perm], f2)
  assert bench._canon(v, f) != bench._canon(v, f[:, [0, 2, 1]])
  assert bench._canon(v, f[:0])[1] == b""

"""Equivalence check (pure Python, no GPU) of k_classify's two-label path (`k2_rows` + the phase B of `tile_body_k2`
in zm_kernels.cuh) against the per-voxel definition (`old()` of tools/emulate_s1_rowmask.py, the marching scan).

For a tile whose staged (33 x 9 x 9) region holds exactly two labels A and B (either may be the background 0), ONE
33-bit mask per staged row -- bit f = voxel f carries A -- determines everything:
  * the six slot bit planes and the active-cube mask of every row segment (must equal the per-voxel definition);
  * the number of vertex slots that belong to A (the label-table reservation issued before S3);
  * the number of non-uniform valid cubes = records per non-zero label;
  * a cube's corner mask for A = eight bits of four row masks, B's = its complement (must equal eight label compares).
Volume-boundary tiles, `close` zero fill and slab shards (Es_own = Es - 1) included.
Run: python tools/emulate_k2_path.py [trials]"""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("emulate_s1_rowmask", os.path.join(HERE, "emulate_s1_rowmask.py"))
S1 = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(S1)

TF, TM, TS, RM, RS = 32, 8, 8, 9, 9
FULL = 0xFFFFFFFF
# reference cube corners (marching_cubes.hpp:353-361) as (df, dm, ds) for Fortran order (f = x, m = y, s = z)
CORNERS = [(0, 0, 0), (1, 0, 0), (1, 0, 1), (0, 0, 1), (0, 1, 0), (1, 1, 0), (1, 1, 1), (0, 1, 1)]


def row_masks(lab, A):
  """k2_rows: 33-bit mask per staged row (bit 32 = the +f halo column)."""
  m = np.zeros(RS * RM, dtype=object)
  for r in range(RS * RM):
    ls, lm = divmod(r, RM)
    bits = 0
    for f in range(TF + 1):
      if lab[ls, lm, f] == A:
        bits |= 1 << f
    m[r] = bits
  return m


def phase_b(mask, A, B, Ef, Em, Es, Es_own, ef0, em0, es0):
  """tile_body_k2's phase B: planes, active mask, A's slots and cubes per row, from the masks alone."""
  zA, zB = A == 0, B == 0
  pl = np.zeros((64, 8), dtype=np.uint64)
  nva_total, ncubes = 0, 0
  for pw in range(TS):
    for pj in range(TM):
      r0 = pw * RM + pj
      m00, m10, m01, m11 = mask[r0], mask[r0 + 1], mask[r0 + RM], mask[r0 + RM + 1]
      lo = lambda x: x & FULL
      a00, a00f, a10, a01 = lo(m00), lo(m00 >> 1), lo(m10), lo(m01)
      ef, em, es = a00 ^ a00f, a00 ^ a10, a00 ^ a01
      nonuni = ef | em | es | lo(m10 ^ (m10 >> 1)) | lo(m01 ^ (m01 >> 1)) | lo(m11 ^ (m11 >> 1)) | lo(m01 ^ m11)
      pf, pm, ps, cube = ef, em, es, nonuni
      if not (ef0 + TF + 1 <= Ef and em0 + TM + 1 <= Em and es0 + TS + 1 <= Es):
        nfv = Ef - ef0
        VF = FULL if nfv >= 32 else (1 << nfv) - 1
        NF1 = FULL if nfv >= 33 else (1 << (nfv - 1)) - 1
        em_, es_ = em0 + pj, es0 + pw
        rowok = em_ < Em and es_ < Es_own
        nm1, ns1 = em_ + 1 < Em, es_ + 1 < Es
        pf = pf & NF1 if rowok else 0
        pm = pm & VF if rowok and nm1 else 0
        ps = ps & VF if rowok and ns1 else 0
        cube = nonuni & NF1 if rowok and nm1 and ns1 else 0
      nz = lambda isA: (~isA & FULL) if zA else (isA if zB else FULL)
      b = [pf & nz(a00), pf & nz(a00f), pm & nz(a00), pm & nz(a10), ps & nz(a00), ps & nz(a01)]
      act = b[0] | b[1] | b[2] | b[3] | b[4] | b[5] | cube
      pl[pw * TM + pj, :6] = b
      pl[pw * TM + pj, 7] = act
      popc = lambda x: bin(x).count("1")
      nva_total += (popc(b[0] & a00) + popc(b[2] & a00) + popc(b[4] & a00) + popc(b[1] & a00f) + popc(b[3] & a10) +
                    popc(b[5] & a01))
      ncubes += popc(cube)
  return pl, nva_total, ncubes


def brute_counts(lab, A, Ef, Em, Es, Es_own, ef0, em0, es0):
  """Per-voxel definition of A's slots and of the valid non-uniform cubes."""
  nva, ncubes = 0, 0
  for ls in range(TS):
    for lm in range(TM):
      for lf in range(TF):
        ef, em, es = ef0 + lf, em0 + lm, es0 + ls
        if not (ef < Ef and em < Em and es < Es_own):
          continue
        u = lab[ls, lm, lf]
        for (d, ok, v) in ((0, ef + 1 < Ef, lab[ls, lm, lf + 1]), (1, em + 1 < Em, lab[ls, lm + 1, lf]),
                           (2, es + 1 < Es, lab[ls + 1, lm, lf])):
          if ok and u != v:
            nva += (u == A and A != 0) + (v == A and A != 0)
        if ef + 1 < Ef and em + 1 < Em and es + 1 < Es:
          c = [lab[ls + ds, lm + dm, lf + df] for (df, dm, ds) in CORNERS]
          if any(x != c[0] for x in c):
            ncubes += 1
  return nva, ncubes


def check(trials=60, seed=3):
  rng = np.random.default_rng(seed)
  n = 0
  for trial in range(trials):
    A, B = [(0, 5), (7, 0), (3, 9), (2 ** 40 + 1, 2 ** 63 + 5)][trial % 4]
    # two-label region: a random half space plus speckle
    g = np.indices((RS, RM, TF + 2)).astype(np.float64)
    nrm = rng.normal(size=3)
    side = (g[0] * nrm[0] + g[1] * nrm[1] + g[2] * nrm[2]) > rng.uniform(2, 20)
    side ^= rng.random(side.shape) < (0.05 if trial % 3 else 0.0)
    lab = np.where(side, np.uint64(A), np.uint64(B)).astype(np.uint64)
    Ef, Em, Es = int(rng.integers(1, 100)), int(rng.integers(1, 30)), int(rng.integers(2, 30))
    if trial % 2 == 0:
      Ef, Em, Es = 200, 200, 200
    Es_own = Es if rng.integers(0, 2) else Es - 1
    ntf, ntm, nts = (Ef + 31) // 32, (Em + 7) // 8, (Es_own + 7) // 8
    tf, tm, ts = int(rng.integers(0, ntf)), int(rng.integers(0, ntm)), int(rng.integers(0, nts))
    ef0, em0, es0 = tf * 32, tm * 8, ts * 8
    # zero fill outside the volume keeps the region two-label only when one label is the background
    if 0 in (A, B):
      for lf in range(TF + 2):
        if ef0 + lf >= Ef: lab[:, :, lf] = 0
      for lm in range(RM):
        if em0 + lm >= Em: lab[:, lm, :] = 0
    elif not (ef0 + TF + 1 <= Ef and em0 + TM + 1 <= Em):
      continue  # (a third label, 0, enters the region: such a tile takes the general path)
    first = lab[0, 0, 0]
    other = B if first == A else A
    mask = row_masks(lab, first)
    pl, nva, ncubes = phase_b(mask, int(first), int(other), Ef, Em, Es, Es_own, ef0, em0, es0)
    want = S1.old(lab, Ef, Em, Es, Es_own, ef0, em0, es0)
    assert np.array_equal(pl, want), (trial, Ef, Em, Es, Es_own, ef0, em0, es0, np.argwhere(pl != want)[:5])
    bn, bc = brute_counts(lab, first, Ef, Em, Es, Es_own, ef0, em0, es0)
    assert (nva, ncubes) == (bn, bc), (trial, nva, bn, ncubes, bc)
    # corner masks of every cube from four row masks == eight label compares; B's mask is the complement
    for ls in range(TS):
      for lm in range(TM):
        r0 = ls * RM + lm
        rows = {(0, 0): mask[r0], (1, 0): mask[r0 + 1], (0, 1): mask[r0 + RM], (1, 1): mask[r0 + RM + 1]}
        for lf in range(0, TF, 5):
          msk = 0
          for k, (df, dm, ds) in enumerate(CORNERS):
            msk |= ((rows[(dm, ds)] >> (lf + df)) & 1) << k
          direct = sum((lab[ls + ds, lm + dm, lf + df] == first) << k for k, (df, dm, ds) in enumerate(CORNERS))
          direct_b = sum((lab[ls + ds, lm + dm, lf + df] == other) << k for k, (df, dm, ds) in enumerate(CORNERS))
          assert msk == direct and (~msk & 0xFF) == direct_b
    n += 1
  return n


if __name__ == "__main__":
  print("two-label tiles checked:", check(int(sys.argv[1]) if len(sys.argv) > 1 else 60))

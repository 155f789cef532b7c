// zmesh_b200 device code: multi-label marching cubes for sm_100a.
//
// Replaces marching_cubes::marche + CMesher::triangles2mesh of the reference
// (zi_lib/zi/mesh/marching_cubes.hpp:291-445, zmesh/cMesher.hpp:96-166) with a dedup-free
// formulation:
//
//   * a vertex of label L is a voxel-grid edge (two axis-adjacent voxels) with exactly one
//     endpoint == L.  Each grid edge is OWNED by its lower voxel, so every vertex is produced
//     exactly once -- no hashing, no sort, no unique pass.
//   * voxel u owns up to 6 vertex slots: slot 2d+0 = (edge u -> u+d, label of u),
//     slot 2d+1 = (same edge, label of u+d), d = memory axis 0 (fastest) .. 2 (slowest).
//   * vertices get a spatial id  g = rowbase[row(u)] + (#slots of earlier voxels in the row)
//     + (#lower slots of u);  perm[g] = index of the vertex inside its label's vertex list.
//     Faces find their vertex indices through perm[], which costs 4 B per vertex instead of the
//     ~6 hash probes per triangle of the reference.
//
// Two passes over the volume (tiles of 32 x 8 x 8 voxels, one CTA each):
//   pass 1 (classify): count per-label vertices/triangles, reserve per-label vertex ranks, write
//                      rowbase[] and perm[];
//   pass 2 (emit):     write the packed 64-bit half-voxel vertex keys and the uint32 faces to
//                      their final per-label ranges.
// followed by the final gather (key -> float32 with anisotropy / voxel_centered, optional normals).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "mc_tables.h"

namespace zm {

// device copies of the case tables (filled once per process by zm_upload_tables)
__device__ uint8_t TRI_COUNT_D[256];
__device__ unsigned long long TRI_NIBBLES_D[256];

constexpr int TF = 32;  // tile extent along the memory-fastest axis (= one warp per row)
constexpr int TM = 8;
constexpr int TS = 8;
constexpr int NT = 256;  // threads per CTA: warp w handles the 8 rows with ls == w
constexpr int NW = NT / 32;
constexpr int LT = 1024;      // slots of the CTA-local label table (aggregates global atomics)
constexpr int LT_PROBES = 32; // bounded probing; on failure the global table is used directly
static_assert(TS == NW, "one warp per s-plane of the tile");

enum : uint32_t {
  FLAG_HASH_FULL = 1u,   // global label table too small -> host grows it and reruns pass 1
  FLAG_PERM_FULL = 2u,   // perm[] capacity guess too small -> host reruns pass 1 with the exact size
  FLAG_INTERNAL = 4u,    // invariant violated (label missing in pass 2, ...)
  FLAG_RANK_OVERFLOW = 8u
};

struct VolParams {
  const void* data;        // device pointer, memory order (f fastest, m, s slowest)
  uint32_t nf, nm, ns;     // input extents
  uint32_t Ef, Em, Es;     // extended extents = n + 2*pad (close => virtual zero border)
  uint32_t pad;            // 1 when close
  uint32_t ntf, ntm, nts;  // tiles per axis
  uint32_t ox, oy, oz;     // shard origin in logical voxels (added to keys)
};

struct LabelTable {  // global open-addressing table, key 0 = empty (label 0 is never meshed)
  unsigned long long* keys;
  uint32_t* cntV;
  uint32_t* cntT;
  uint32_t mask;  // capacity - 1
};

struct Pass1Args {
  LabelTable ht;
  uint32_t* rowbase;  // [Es*Em*ntf]
  uint32_t* perm;     // [permcap]
  unsigned long long permcap;
  unsigned long long* cursor;  // perm segment allocator
  uint32_t* flags;
};

struct Pass2Args {
  LabelTable ht;
  const unsigned long long* offV;  // per table slot: first vertex row of the label
  const unsigned long long* offT;
  uint32_t* curT;  // per table slot: running triangle cursor
  const uint32_t* rowbase;
  const uint32_t* perm;
  unsigned long long* vkeys;  // [V_total]
  uint32_t* faces;            // [T_total][3]
  uint32_t* flags;
};

// ---------------------------------------------------------------------------------------------
// geometry of the reference's cube (marching_cubes.hpp:299-316, :353-361), in LOGICAL axes.
// corner n -> (dx,dy,dz);  edge e joins corners EA[e], EB[e].

__host__ __device__ constexpr int corner_dx(int n) { return (0x66 >> n) & 1; }  // 0,1,1,0,0,1,1,0
__host__ __device__ constexpr int corner_dy(int n) { return (0xF0 >> n) & 1; }  // 0,0,0,0,1,1,1,1
__host__ __device__ constexpr int corner_dz(int n) { return (0xCC >> n) & 1; }  // 0,0,1,1,0,0,1,1
__host__ __device__ constexpr int edge_a(int e) { return e < 8 ? e : e - 8; }
__host__ __device__ constexpr int edge_b(int e) { return e < 4 ? (e + 1) & 3 : (e < 8 ? 4 + ((e + 1) & 3) : e - 4); }

// memory-axis view: F order (x fastest): f=x, m=y, s=z;  C order (z fastest): f=z, m=y, s=x.
template <bool CO> __host__ __device__ constexpr int corner_df(int n) { return CO ? corner_dz(n) : corner_dx(n); }
template <bool CO> __host__ __device__ constexpr int corner_dm(int n) { return corner_dy(n); }
template <bool CO> __host__ __device__ constexpr int corner_ds(int n) { return CO ? corner_dx(n) : corner_dz(n); }

// per edge, 5 bits: owner-voxel offset (of, om, os) and the memory axis of the edge (2 bits).
// The midpoint M = corner_a + corner_b (half-voxel units): the axis is where M == 1, the owner
// (lower endpoint) is M >> 1.
template <bool CO>
__host__ __device__ constexpr unsigned long long edge_info_packed() {
  unsigned long long w = 0;
  for (int e = 0; e < 12; ++e) {
    int a = edge_a(e), b = edge_b(e);
    int mf = corner_df<CO>(a) + corner_df<CO>(b);
    int mm = corner_dm<CO>(a) + corner_dm<CO>(b);
    int ms = corner_ds<CO>(a) + corner_ds<CO>(b);
    int axis = mf == 1 ? 0 : (mm == 1 ? 1 : 2);
    unsigned long long v = (unsigned long long)((mf >> 1) | ((mm >> 1) << 1) | ((ms >> 1) << 2) | (axis << 3));
    w |= v << (5 * e);
  }
  return w;
}

// ---------------------------------------------------------------------------------------------

__device__ __forceinline__ uint32_t hash_label(unsigned long long x) {
  x ^= x >> 33;
  x *= 0xff51afd7ed558ccdull;
  x ^= x >> 29;
  return (uint32_t)x ^ (uint32_t)(x >> 32);
}

// Global table: find-or-insert.  Returns the slot, or -1 (and raises FLAG_HASH_FULL).
__device__ __forceinline__ int gtab_insert(const LabelTable& ht, unsigned long long label, uint32_t* flags) {
  uint32_t h = hash_label(label) & ht.mask;
  const uint32_t max_probe = ht.mask < 4095u ? ht.mask + 1u : 4096u;
  for (uint32_t p = 0; p < max_probe; ++p) {
    unsigned long long old = atomicCAS(&ht.keys[h], 0ull, label);
    if (old == 0ull || old == label) return (int)h;
    h = (h + 1u) & ht.mask;
  }
  atomicOr(flags, FLAG_HASH_FULL);
  return -1;
}

// Global table: read-only lookup (pass 2; the key must exist).
__device__ __forceinline__ int gtab_find(const LabelTable& ht, unsigned long long label, uint32_t* flags) {
  uint32_t h = hash_label(label) & ht.mask;
  const uint32_t max_probe = ht.mask < 4095u ? ht.mask + 1u : 4096u;
  for (uint32_t p = 0; p < max_probe; ++p) {
    unsigned long long k = ht.keys[h];
    if (k == label) return (int)h;
    if (k == 0ull) break;
    h = (h + 1u) & ht.mask;
  }
  atomicOr(flags, FLAG_INTERNAL);
  return -1;
}

// CTA-local table in shared memory.  A label either gets a slot (all later lookups succeed) or
// every attempt fails identically (slots are never freed and the probe sequence is a function of
// the label), in which case callers go to the global table directly.
__device__ __forceinline__ int ltab_insert(unsigned long long* keys, unsigned long long label) {
  uint32_t h = hash_label(label) & (LT - 1);
#pragma unroll 1
  for (int p = 0; p < LT_PROBES; ++p) {
    unsigned long long k = *(volatile unsigned long long*)&keys[h];
    if (k == label) return (int)h;
    if (k == 0ull) {
      unsigned long long old = atomicCAS(&keys[h], 0ull, label);
      if (old == 0ull || old == label) return (int)h;
    }
    h = (h + 1u) & (LT - 1);
  }
  return -1;
}

__device__ __forceinline__ int ltab_find(const unsigned long long* keys, unsigned long long label) {
  uint32_t h = hash_label(label) & (LT - 1);
#pragma unroll 1
  for (int p = 0; p < LT_PROBES; ++p) {
    unsigned long long k = keys[h];
    if (k == label) return (int)h;
    if (k == 0ull) return -1;
    h = (h + 1u) & (LT - 1);
  }
  return -1;
}

// ---------------------------------------------------------------------------------------------
// tile staging: region of (TF+H) x (TM+H) x (TS+H) labels, origin = tile origin, zero outside
// the input volume (this is also what makes `close` free: the virtual border reads as 0).

template <typename L, int H>
__device__ __forceinline__ void load_region(const VolParams& vp, L* lab, uint32_t ef0, uint32_t em0, uint32_t es0) {
  constexpr int RF = TF + H, RM = TM + H, RS = TS + H;
  const L* __restrict__ src = static_cast<const L*>(vp.data);
  for (int i = threadIdx.x; i < RF * RM * RS; i += NT) {
    int lf = i % RF;
    int t = i / RF;
    int lm = t % RM;
    int ls = t / RM;
    uint32_t jf = ef0 + lf - vp.pad, jm = em0 + lm - vp.pad, js = es0 + ls - vp.pad;  // wraps when < 0
    L v = 0;
    if (jf < vp.nf && jm < vp.nm && js < vp.ns) v = src[((size_t)js * vp.nm + jm) * vp.nf + jf];
    lab[i] = v;
  }
}

struct TileCoord {
  uint32_t tf, tm, ts, ef0, em0, es0;
};

__device__ __forceinline__ TileCoord tile_coord(const VolParams& vp) {
  TileCoord t;
  uint32_t b = blockIdx.x;
  t.tf = b % vp.ntf;
  b /= vp.ntf;
  t.tm = b % vp.ntm;
  t.ts = b / vp.ntm;
  t.ef0 = t.tf * TF;
  t.em0 = t.tm * TM;
  t.es0 = t.ts * TS;
  return t;
}

// 6-bit slot mask of voxel (lf,lm,ls) given its label a and its +f,+m,+s neighbours; validity of
// the voxel and of each neighbour (inside the extended volume) passed in.
template <typename L>
__device__ __forceinline__ uint32_t slot_mask(L a, L bf, L bm, L bs, bool vf, bool vm, bool vs) {
  uint32_t m = 0;
  if (vf && a != bf) m |= (a != 0 ? 1u : 0u) | (bf != 0 ? 2u : 0u);
  if (vm && a != bm) m |= (a != 0 ? 4u : 0u) | (bm != 0 ? 8u : 0u);
  if (vs && a != bs) m |= (a != 0 ? 16u : 0u) | (bs != 0 ? 32u : 0u);
  return m;
}

// The cube with origin (lf,lm,ls): 8 corner labels in the reference's corner order.
template <typename L, bool CO, int RF, int RM>
__device__ __forceinline__ void load_cube(const L* lab, int lf, int lm, int ls, unsigned long long c[8]) {
#pragma unroll
  for (int n = 0; n < 8; ++n)
    c[n] = (unsigned long long)lab[((ls + corner_ds<CO>(n)) * RM + (lm + corner_dm<CO>(n))) * RF + (lf + corner_df<CO>(n))];
}

__device__ __forceinline__ bool cube_uniform(const unsigned long long c[8]) {
  return c[0] == c[1] && c[0] == c[2] && c[0] == c[3] && c[0] == c[4] && c[0] == c[5] && c[0] == c[6] && c[0] == c[7];
}

// ---------------------------------------------------------------------------------------------
// pass 1: classify + count + reserve

template <typename L>
__host__ __device__ constexpr size_t classify_smem_bytes() { return sizeof(L) * (TF + 1) * (TM + 1) * (TS + 1); }
template <typename L>
__host__ __device__ constexpr size_t emit_smem_bytes() { return sizeof(L) * (TF + 2) * (TM + 2) * (TS + 2); }

extern __shared__ __align__(16) unsigned char zm_dyn_smem[];

template <typename L, bool CO>
__global__ void __launch_bounds__(NT) k_classify(const VolParams vp, const Pass1Args o) {
  constexpr int RF = TF + 1, RM = TM + 1;
  L* lab = reinterpret_cast<L*>(zm_dyn_smem);  // [TS+1][TM+1][TF+1]
  __shared__ unsigned long long lkeys[LT];
  __shared__ uint32_t lvcnt[LT];  // vertices per local label, then the running rank cursor
  __shared__ uint32_t ltcnt[LT];
  __shared__ uint32_t rowcnt[TM * TS];
  __shared__ uint32_t rowpre[TM * TS];
  __shared__ unsigned long long s_tilebase;
  __shared__ uint32_t s_ok;
  __shared__ uint8_t s_tricount[256];

  const TileCoord tc = tile_coord(vp);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  load_region<L, 1>(vp, lab, tc.ef0, tc.em0, tc.es0);
  __syncthreads();

  // ---- phase A: slot masks (packed in registers), row counts, activity ----
  const int ls = warp, lf = lane;
  const uint32_t ef = tc.ef0 + lf, es = tc.es0 + ls;
  unsigned long long sm6p = 0;  // byte j: 6-bit slot mask of voxel (lf, j, ls)
  unsigned long long prep = 0;  // byte j: exclusive in-row prefix of slot counts (<= 186)
  uint32_t active = 0;          // bit j: cube (lf, j, ls) exists and is non-uniform
#pragma unroll
  for (int j = 0; j < TM; ++j) {
    const int lm = j;
    const uint32_t em = tc.em0 + lm;
    const bool valid = ef < vp.Ef && em < vp.Em && es < vp.Es;
    const int idx = (ls * RM + lm) * RF + lf;
    const L a = lab[idx], bf = lab[idx + 1], bm = lab[idx + RF], bs = lab[idx + RF * RM];
    const uint32_t m = valid ? slot_mask<L>(a, bf, bm, bs, ef + 1 < vp.Ef, em + 1 < vp.Em, es + 1 < vp.Es) : 0u;
    const uint32_t c = __popc(m);
    uint32_t inc = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += t;
    }
    sm6p |= (unsigned long long)m << (8 * j);
    prep |= (unsigned long long)(inc - c) << (8 * j);
    if (lane == 31) rowcnt[ls * TM + lm] = inc;
    if (valid && ef + 1 < vp.Ef && em + 1 < vp.Em && es + 1 < vp.Es) {
      unsigned long long cl[8];
      load_cube<L, CO, RF, RM>(lab, lf, lm, ls, cl);
      if (!cube_uniform(cl)) active |= 1u << j;
    }
  }
  if (!__syncthreads_or((sm6p != 0ull || active != 0u) ? 1 : 0)) return;  // uniform tile

  for (int i = threadIdx.x; i < LT; i += NT) { lkeys[i] = 0ull; lvcnt[i] = 0u; ltcnt[i] = 0u; }
  s_tricount[threadIdx.x] = TRI_COUNT_D[threadIdx.x];
  __syncthreads();

  // ---- phase B: per-(tile,label) counts in the local table ----
#pragma unroll 1
  for (int j = 0; j < TM; ++j) {
    const int lm = j;
    const int idx = (ls * RM + lm) * RF + lf;
    const uint32_t m = (uint32_t)(sm6p >> (8 * j)) & 63u;
    if (m) {
      const L a = lab[idx];
      const uint32_t na = __popc(m & 0x15u);  // slots carrying my own label
      if (na) {
        int s = ltab_insert(lkeys, (unsigned long long)a);
        if (s >= 0) atomicAdd(&lvcnt[s], na);
      }
#pragma unroll
      for (int d = 0; d < 3; ++d)
        if (m & (2u << (2 * d))) {
          const L b = lab[idx + (d == 0 ? 1 : (d == 1 ? RF : RF * RM))];
          int s = ltab_insert(lkeys, (unsigned long long)b);
          if (s >= 0) atomicAdd(&lvcnt[s], 1u);
        }
    }
    if (active & (1u << j)) {
      unsigned long long cl[8];
      load_cube<L, CO, RF, RM>(lab, lf, lm, ls, cl);
      uint32_t acc = 0;
      while (acc != 0xFFu) {
        const int start = __ffs(~acc & 0xFFu) - 1;
        unsigned long long label = cl[0];
#pragma unroll
        for (int n = 1; n < 8; ++n) label = (n == start) ? cl[n] : label;
        uint32_t msk = 0;
#pragma unroll
        for (int n = 0; n < 8; ++n) msk |= (cl[n] == label ? 1u : 0u) << n;
        acc |= msk;
        if (label == 0ull) continue;
        const uint32_t nt = s_tricount[~msk & 0xFFu];
        if (nt == 0u) continue;
        int s = ltab_insert(lkeys, label);
        if (s >= 0) atomicAdd(&ltcnt[s], nt);
        else {
          int gs = gtab_insert(o.ht, label, o.flags);
          if (gs >= 0) atomicAdd(&o.ht.cntT[gs], nt);
        }
      }
    }
  }
  __syncthreads();

  // ---- phase C: reserve per-label rank ranges and the tile's perm segment ----
  for (int i = threadIdx.x; i < LT; i += NT) {
    const unsigned long long label = lkeys[i];
    if (label != 0ull) {
      const uint32_t nv = lvcnt[i], nt = ltcnt[i];
      const int gs = gtab_insert(o.ht, label, o.flags);
      uint32_t base = 0;
      if (gs >= 0) {
        if (nv) {
          base = atomicAdd(&o.ht.cntV[gs], nv);
          if (base + nv < base) atomicOr(o.flags, FLAG_RANK_OVERFLOW);
        }
        if (nt) atomicAdd(&o.ht.cntT[gs], nt);
      }
      lvcnt[i] = base;
    }
  }
  if (warp == 0) {
    const uint32_t a0 = rowcnt[2 * lane], a1 = rowcnt[2 * lane + 1];
    const uint32_t sum = a0 + a1;
    uint32_t inc = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += t;
    }
    rowpre[2 * lane] = inc - sum;
    rowpre[2 * lane + 1] = inc - sum + a0;
    const uint32_t total = __shfl_sync(0xffffffffu, inc, 31);
    if (lane == 0) {
      const unsigned long long base = total ? atomicAdd(o.cursor, (unsigned long long)total) : 0ull;
      const bool ok = base + total <= o.permcap;
      if (!ok) atomicOr(o.flags, FLAG_PERM_FULL);
      s_tilebase = base;
      s_ok = ok ? 1u : 0u;
    }
  }
  __syncthreads();
  if (!s_ok) return;  // capacity guess too small: the cursor still gives the exact need; host reruns
  const unsigned long long tilebase = s_tilebase;

  // rowbase for the rows of this tile (row id = (es*Em + em)*ntf + tf)
  if (threadIdx.x < TM * TS) {
    const int r = threadIdx.x;
    const uint32_t rs = tc.es0 + r / TM, rm = tc.em0 + r % TM;
    if (rs < vp.Es && rm < vp.Em)
      o.rowbase[((size_t)rs * vp.Em + rm) * vp.ntf + tc.tf] = (uint32_t)(tilebase + rowpre[r]);
  }

  // ---- phase D: vertex ranks -> perm[g] ----
#pragma unroll 1
  for (int j = 0; j < TM; ++j) {
    const uint32_t m = (uint32_t)(sm6p >> (8 * j)) & 63u;
    if (!m) continue;
    const int lm = j;
    const int idx = (ls * RM + lm) * RF + lf;
    const L a = lab[idx];
    unsigned long long g = tilebase + rowpre[ls * TM + lm] + ((uint32_t)(prep >> (8 * j)) & 255u);
#pragma unroll
    for (int s6 = 0; s6 < 6; ++s6) {
      if (!(m & (1u << s6))) continue;
      const int d = s6 >> 1;
      const unsigned long long label =
          (s6 & 1) ? (unsigned long long)lab[idx + (d == 0 ? 1 : (d == 1 ? RF : RF * RM))] : (unsigned long long)a;
      const int s = ltab_find(lkeys, label);
      uint32_t rank;
      if (s >= 0) rank = atomicAdd(&lvcnt[s], 1u);
      else {
        const int gs = gtab_insert(o.ht, label, o.flags);
        rank = gs >= 0 ? atomicAdd(&o.ht.cntV[gs], 1u) : 0u;
      }
      o.perm[g] = rank;
      ++g;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// pass 2: emit vertex keys and faces

template <typename L, bool CO>
__global__ void __launch_bounds__(NT) k_emit(const VolParams vp, const Pass2Args o) {
  constexpr int RF = TF + 2, RM = TM + 2;               // labels: halo 2
  constexpr int AF = TF + 1, AM = TM + 1, AS = TS + 1;  // voxels whose slots can be referenced
  L* lab = reinterpret_cast<L*>(zm_dyn_smem);           // [TS+2][TM+2][TF+2]
  __shared__ uint8_t own6[AF * AM * AS];
  __shared__ uint8_t pre8[AF * AM * AS];
  __shared__ uint32_t rb[AM * AS][2];  // rowbase of the row in this tile column / in the next one
  __shared__ unsigned long long lkeys[LT];
  __shared__ uint32_t ltcnt[LT];  // triangles per local label, then the running cursor
  __shared__ uint32_t lgs[LT];    // global slot of the local label
  __shared__ unsigned long long s_trinib[256];
  __shared__ uint8_t s_tricount[256];

  const TileCoord tc = tile_coord(vp);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  load_region<L, 2>(vp, lab, tc.ef0, tc.em0, tc.es0);
  __syncthreads();

  // ---- phase A: slot masks + in-row prefixes for the (TF+1)(TM+1)(TS+1) referenced voxels ----
  bool any = false;
  for (int r = warp; r < AM * AS; r += NW) {
    const int lm = r % AM, ls = r / AM;
    const uint32_t em = tc.em0 + lm, es = tc.es0 + ls;
    const bool rowvalid = em < vp.Em && es < vp.Es;
    {
      const int lf = lane;
      const uint32_t ef = tc.ef0 + lf;
      const int idx = (ls * RM + lm) * RF + lf;
      uint32_t m = 0;
      if (rowvalid && ef < vp.Ef)
        m = slot_mask<L>(lab[idx], lab[idx + 1], lab[idx + RF], lab[idx + RF * RM], ef + 1 < vp.Ef,
                         em + 1 < vp.Em, es + 1 < vp.Es);
      const uint32_t c = __popc(m);
      uint32_t inc = c;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
      }
      own6[(ls * AM + lm) * AF + lf] = (uint8_t)m;
      pre8[(ls * AM + lm) * AF + lf] = (uint8_t)(inc - c);
      any |= (m != 0u);
    }
    if (lane == 0) {  // halo column lf == TF: first voxel of the next tile's row
      const int lf = TF;
      const uint32_t ef = tc.ef0 + lf;
      const int idx = (ls * RM + lm) * RF + lf;
      uint32_t m = 0;
      if (rowvalid && ef < vp.Ef)
        m = slot_mask<L>(lab[idx], lab[idx + 1], lab[idx + RF], lab[idx + RF * RM], ef + 1 < vp.Ef,
                         em + 1 < vp.Em, es + 1 < vp.Es);
      own6[(ls * AM + lm) * AF + lf] = (uint8_t)m;
      pre8[(ls * AM + lm) * AF + lf] = 0;
      any |= (m != 0u);
      const size_t row = ((size_t)es * vp.Em + em) * vp.ntf + tc.tf;
      rb[r][0] = rowvalid ? o.rowbase[row] : 0u;
      rb[r][1] = (rowvalid && tc.tf + 1 < vp.ntf) ? o.rowbase[row + 1] : 0u;
    }
  }
  // every edge of an owned cube, and every owned edge, is owned by one of the voxels above
  if (!__syncthreads_or(any ? 1 : 0)) return;

  for (int i = threadIdx.x; i < LT; i += NT) { lkeys[i] = 0ull; ltcnt[i] = 0u; }
  s_tricount[threadIdx.x] = TRI_COUNT_D[threadIdx.x];
  s_trinib[threadIdx.x] = TRI_NIBBLES_D[threadIdx.x];
  __syncthreads();

  // ---- phase B: triangles per (tile,label) ----
  const int ls = warp, lf = lane;
  const uint32_t ef = tc.ef0 + lf, es = tc.es0 + ls;
  uint32_t active = 0;
#pragma unroll 1
  for (int j = 0; j < TM; ++j) {
    const int lm = j;
    const uint32_t em = tc.em0 + lm;
    if (!(ef + 1 < vp.Ef && em + 1 < vp.Em && es + 1 < vp.Es)) continue;
    unsigned long long cl[8];
    load_cube<L, CO, RF, RM>(lab, lf, lm, ls, cl);
    if (cube_uniform(cl)) continue;
    active |= 1u << j;
    uint32_t acc = 0;
    while (acc != 0xFFu) {
      const int start = __ffs(~acc & 0xFFu) - 1;
      unsigned long long label = cl[0];
#pragma unroll
      for (int n = 1; n < 8; ++n) label = (n == start) ? cl[n] : label;
      uint32_t msk = 0;
#pragma unroll
      for (int n = 0; n < 8; ++n) msk |= (cl[n] == label ? 1u : 0u) << n;
      acc |= msk;
      if (label == 0ull) continue;
      const uint32_t nt = s_tricount[~msk & 0xFFu];
      if (nt == 0u) continue;
      int s = ltab_insert(lkeys, label);
      if (s >= 0) atomicAdd(&ltcnt[s], nt);
    }
  }
  __syncthreads();

  // ---- phase C: reserve the tile's face ranges ----
  for (int i = threadIdx.x; i < LT; i += NT) {
    const unsigned long long label = lkeys[i];
    if (label != 0ull) {
      const int gs = gtab_find(o.ht, label, o.flags);
      const uint32_t nt = ltcnt[i];
      const uint32_t base = (gs >= 0 && nt) ? atomicAdd(&o.curT[gs], nt) : 0u;
      ltcnt[i] = base;
      lgs[i] = gs >= 0 ? (uint32_t)gs : 0xFFFFFFFFu;
    }
  }
  __syncthreads();

  // ---- phase D: owned vertices -> keys at their final position ----
#pragma unroll 1
  for (int j = 0; j < TM; ++j) {
    const int lm = j;
    const int aidx = (ls * AM + lm) * AF + lf;
    const uint32_t m = own6[aidx];
    if (!m) continue;
    const uint32_t em = tc.em0 + lm;
    const int idx = (ls * RM + lm) * RF + lf;
    const L a = lab[idx];
    uint32_t g = rb[ls * AM + lm][0] + pre8[aidx];
#pragma unroll
    for (int s6 = 0; s6 < 6; ++s6) {
      if (!(m & (1u << s6))) continue;
      const int d = s6 >> 1;
      const unsigned long long label =
          (s6 & 1) ? (unsigned long long)lab[idx + (d == 0 ? 1 : (d == 1 ? RF : RF * RM))] : (unsigned long long)a;
      const int s = ltab_find(lkeys, label);
      const int gs = s >= 0 ? (int)lgs[s] : gtab_find(o.ht, label, o.flags);
      const uint32_t rank = o.perm[g];
      ++g;
      if (gs < 0) continue;
      // half-voxel coordinates of the edge midpoint, memory axes -> logical axes
      const unsigned long long hf = 2ull * ef + (d == 0), hm = 2ull * em + (d == 1), hs = 2ull * es + (d == 2);
      const unsigned long long kx = (CO ? hs : hf) + 2ull * vp.ox;
      const unsigned long long ky = hm + 2ull * vp.oy;
      const unsigned long long kz = (CO ? hf : hs) + 2ull * vp.oz;
      o.vkeys[o.offV[gs] + rank] = (kx << 42) | (ky << 21) | kz;
    }
  }

  // ---- phase E: faces ----
  constexpr unsigned long long EINFO = edge_info_packed<CO>();
#pragma unroll 1
  for (int j = 0; j < TM; ++j) {
    if (!(active & (1u << j))) continue;
    const int lm = j;
    unsigned long long cl[8];
    load_cube<L, CO, RF, RM>(lab, lf, lm, ls, cl);
    uint32_t acc = 0;
    while (acc != 0xFFu) {
      const int start = __ffs(~acc & 0xFFu) - 1;
      unsigned long long label = cl[0];
#pragma unroll
      for (int n = 1; n < 8; ++n) label = (n == start) ? cl[n] : label;
      uint32_t msk = 0;
#pragma unroll
      for (int n = 0; n < 8; ++n) msk |= (cl[n] == label ? 1u : 0u) << n;
      acc |= msk;
      if (label == 0ull) continue;
      const uint32_t cs = ~msk & 0xFFu;
      const uint32_t nt = s_tricount[cs];
      if (nt == 0u) continue;
      const int s = ltab_find(lkeys, label);
      int gs;
      uint32_t tb;
      if (s >= 0) {
        gs = (int)lgs[s];
        tb = atomicAdd(&ltcnt[s], nt);
      } else {
        gs = gtab_find(o.ht, label, o.flags);
        tb = gs >= 0 ? atomicAdd(&o.curT[gs], nt) : 0u;
      }
      if (gs < 0) continue;
      uint32_t* fout = o.faces + 3ull * (o.offT[gs] + tb);
      const unsigned long long nib = s_trinib[cs];
      for (uint32_t t = 0; t < nt; ++t) {
        uint32_t vidx[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const int e = (int)((nib >> (12 * t + 4 * k)) & 0xFull);
          const uint32_t info = (uint32_t)(EINFO >> (5 * e)) & 31u;
          const int uf = lf + (int)(info & 1u), um = lm + (int)((info >> 1) & 1u), us = ls + (int)((info >> 2) & 1u);
          const uint32_t d = info >> 3;
          // side 0: the owner (lower) voxel carries `label`; side 1: the upper one does
          const unsigned long long lower = (unsigned long long)lab[(us * RM + um) * RF + uf];
          const uint32_t slot = 2u * d + (lower == label ? 0u : 1u);
          const int aidx = (us * AM + um) * AF + uf;
          const uint32_t g = rb[us * AM + um][uf == TF ? 1 : 0] + pre8[aidx] +
                             __popc((uint32_t)own6[aidx] & ((1u << slot) - 1u));
          vidx[k] = o.perm[g];
        }
        // reference winding of Mesher.get: (E[T[3n+1]], E[T[3n]], E[T[3n+2]])
        // (marching_cubes.hpp:338-343 then cMesher.hpp:158-162)
        fout[3 * t + 0] = vidx[1];
        fout[3 * t + 1] = vidx[0];
        fout[3 * t + 2] = vidx[2];
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// label table scan: per-slot exclusive offsets + compact list of live labels (one CTA)

struct ScanOut {
  unsigned long long* offV;  // [cap]
  unsigned long long* offT;  // [cap]
  unsigned long long* list;  // [cap][3]: label, nV, nT  (compact, table order)
  unsigned long long* totals;  // [4]: n_labels, V_total, T_total, perm cursor (copied)
};

__global__ void __launch_bounds__(1024) k_label_scan(const LabelTable ht, const ScanOut so,
                                                    const unsigned long long* cursor) {
  __shared__ unsigned long long sV[1024], sT[1024];
  __shared__ uint32_t sN[1024];
  const uint32_t cap = ht.mask + 1u;
  const uint32_t per = (cap + 1023u) / 1024u;
  const uint32_t lo = threadIdx.x * per, hi = min(cap, lo + per);
  unsigned long long v = 0, t = 0;
  uint32_t n = 0;
  for (uint32_t i = lo; i < hi; ++i) {
    if (ht.keys[i] != 0ull) {
      v += ht.cntV[i];
      t += ht.cntT[i];
      n += (ht.cntT[i] != 0u || ht.cntV[i] != 0u) ? 1u : 0u;
    }
  }
  sV[threadIdx.x] = v; sT[threadIdx.x] = t; sN[threadIdx.x] = n;
  __syncthreads();
  // Hillis-Steele inclusive scan over 1024 partials
  for (int d = 1; d < 1024; d <<= 1) {
    unsigned long long av = 0, at = 0;
    uint32_t an = 0;
    if ((int)threadIdx.x >= d) { av = sV[threadIdx.x - d]; at = sT[threadIdx.x - d]; an = sN[threadIdx.x - d]; }
    __syncthreads();
    sV[threadIdx.x] += av; sT[threadIdx.x] += at; sN[threadIdx.x] += an;
    __syncthreads();
  }
  unsigned long long bv = sV[threadIdx.x] - v, bt = sT[threadIdx.x] - t;
  uint32_t bn = sN[threadIdx.x] - n;
  for (uint32_t i = lo; i < hi; ++i) {
    so.offV[i] = bv;
    so.offT[i] = bt;
    if (ht.keys[i] != 0ull) {
      uint32_t cv = ht.cntV[i], ct = ht.cntT[i];
      if (cv != 0u || ct != 0u) {
        so.list[3ull * bn + 0] = ht.keys[i];
        so.list[3ull * bn + 1] = cv;
        so.list[3ull * bn + 2] = ct;
        ++bn;
      }
      bv += cv;
      bt += ct;
    }
  }
  if (threadIdx.x == 1023) {
    so.totals[0] = sN[1023];
    so.totals[1] = sV[1023];
    so.totals[2] = sT[1023];
    so.totals[3] = *cursor;
  }
}

// ---------------------------------------------------------------------------------------------
// final gather: key -> float32 vertex (reference: unpack_* marching_cubes.hpp:114-135 with
// offset 0, factor = captured resolution; then _normalize_mesh zmesh/_zmesh.pyx:423-433).
// Three separately rounded float32 operations, no FMA contraction.

struct FinalizeArgs {
  const unsigned long long* vkeys;
  float* verts;
  unsigned long long nV;
  float r0, r1, r2;  // captured resolution
  float c0, c1, c2;  // centering offset
  int voxel_centered;
  int transpose;
};

__device__ __forceinline__ void key_to_p(unsigned long long k, float r0, float r1, float r2, int transpose,
                                         float& p0, float& p1, float& p2) {
  float kx = __fadd_rn(0.0f, (float)(uint32_t)((k >> 42) & 0x1FFFFFull));
  float ky = __fadd_rn(0.0f, (float)(uint32_t)((k >> 21) & 0x1FFFFFull));
  float kz = __fadd_rn(0.0f, (float)(uint32_t)(k & 0x1FFFFFull));
  if (transpose) {  // cMesher.hpp:128-138
    p0 = __fmul_rn(r0, kz); p1 = __fmul_rn(r1, ky); p2 = __fmul_rn(r2, kx);
  } else {          // cMesher.hpp:139-149
    p0 = __fmul_rn(r0, kx); p1 = __fmul_rn(r1, ky); p2 = __fmul_rn(r2, kz);
  }
}

__global__ void __launch_bounds__(256) k_finalize_vertices(const FinalizeArgs a) {
  unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (; i < a.nV; i += stride) {
    float p0, p1, p2;
    key_to_p(a.vkeys[i], a.r0, a.r1, a.r2, a.transpose, p0, p1, p2);
    if (a.voxel_centered) { p0 = __fadd_rn(p0, a.c0); p1 = __fadd_rn(p1, a.c1); p2 = __fadd_rn(p2, a.c2); }
    a.verts[3 * i + 0] = __fmul_rn(p0, 0.5f);  // == p / 2.0f exactly
    a.verts[3 * i + 1] = __fmul_rn(p1, 0.5f);
    a.verts[3 * i + 2] = __fmul_rn(p2, 0.5f);
  }
}

// Normals (reference zmesh/chunk_mesh.hpp:345-384 on the pre-normalisation vertices res*k).
struct NormalsArgs {
  const unsigned long long* vkeys;
  const uint32_t* faces;
  float* normals;  // [nV][3], zeroed
  const unsigned long long* voff;  // [nSlots] per label-table slot (non-decreasing)
  const unsigned long long* foff;  // [nSlots]
  unsigned long long nT, nV;
  uint32_t nSlots;
  float r0, r1, r2;
  int transpose;
};

__device__ __forceinline__ float len3(float x, float y, float z) {
  return __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
}

// One face: n_hat = hat(cross(v1-v0, v2-v0)); N[f_k] += n_hat * |v_k - centroid|
// (chunk_mesh.hpp:355-368; float32 op for op, accumulation order is not the reference's).
__device__ __forceinline__ void face_normal_scatter(const float v0[3], const float v1[3], const float v2[3],
                                                    float* d0, float* d1, float* d2) {
  float c[3], e1[3], e2[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    c[d] = __fdiv_rn(__fadd_rn(__fadd_rn(v0[d], v1[d]), v2[d]), 3.0f);
    e1[d] = __fsub_rn(v1[d], v0[d]);
    e2[d] = __fsub_rn(v2[d], v0[d]);
  }
  float n0 = __fsub_rn(__fmul_rn(e1[1], e2[2]), __fmul_rn(e1[2], e2[1]));
  float n1 = __fsub_rn(__fmul_rn(e1[2], e2[0]), __fmul_rn(e1[0], e2[2]));
  float n2 = __fsub_rn(__fmul_rn(e1[0], e2[1]), __fmul_rn(e1[1], e2[0]));
  const float l = len3(n0, n1, n2);
  if (l != 1.0f) { n0 = __fdiv_rn(n0, l); n1 = __fdiv_rn(n1, l); n2 = __fdiv_rn(n2, l); }
  const float* vv[3] = {v0, v1, v2};
  float* dd[3] = {d0, d1, d2};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float w = len3(__fsub_rn(vv[k][0], c[0]), __fsub_rn(vv[k][1], c[1]), __fsub_rn(vv[k][2], c[2]));
    atomicAdd(dd[k] + 0, __fmul_rn(n0, w));
    atomicAdd(dd[k] + 1, __fmul_rn(n1, w));
    atomicAdd(dd[k] + 2, __fmul_rn(n2, w));
  }
}

__global__ void __launch_bounds__(256) k_normals_accumulate(const NormalsArgs a) {
  unsigned long long j = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (; j < a.nT; j += stride) {
    // table slot of face j: last i with foff[i] <= j (empty slots repeat the next live offset)
    uint32_t lo = 0, hi = a.nSlots;
    while (hi - lo > 1) {
      uint32_t mid = (lo + hi) >> 1;
      if (a.foff[mid] <= j) lo = mid; else hi = mid;
    }
    const unsigned long long vb = a.voff[lo];
    uint32_t f0 = a.faces[3 * j + 0], f1 = a.faces[3 * j + 1], f2 = a.faces[3 * j + 2];
    if (a.transpose) { uint32_t t = f0; f0 = f2; f2 = t; }  // legacy faces (t0,t2,t1) = stored row reversed
    float v0[3], v1[3], v2[3];
    key_to_p(a.vkeys[vb + f0], a.r0, a.r1, a.r2, a.transpose, v0[0], v0[1], v0[2]);
    key_to_p(a.vkeys[vb + f1], a.r0, a.r1, a.r2, a.transpose, v1[0], v1[1], v1[2]);
    key_to_p(a.vkeys[vb + f2], a.r0, a.r1, a.r2, a.transpose, v2[0], v2[1], v2[2]);
    face_normal_scatter(v0, v1, v2, a.normals + 3ull * (vb + f0), a.normals + 3ull * (vb + f1),
                        a.normals + 3ull * (vb + f2));
  }
}

// Mesher.compute_normals on an arbitrary float32 mesh (zmesh/_zmesh.pyx:138-152).
__global__ void __launch_bounds__(256) k_normals_accumulate_f32(const float* __restrict__ verts,
                                                               const uint32_t* __restrict__ faces,
                                                               unsigned long long nT, float* normals) {
  unsigned long long j = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (; j < nT; j += stride) {
    const uint32_t f0 = faces[3 * j + 0], f1 = faces[3 * j + 1], f2 = faces[3 * j + 2];
    float v0[3], v1[3], v2[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      v0[d] = verts[3ull * f0 + d];
      v1[d] = verts[3ull * f1 + d];
      v2[d] = verts[3ull * f2 + d];
    }
    face_normal_scatter(v0, v1, v2, normals + 3ull * f0, normals + 3ull * f1, normals + 3ull * f2);
  }
}

__global__ void __launch_bounds__(256) k_normals_normalize(float* normals, unsigned long long nV) {
  unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (; i < nV; i += stride) {
    float x = normals[3 * i], y = normals[3 * i + 1], z = normals[3 * i + 2];
    float l = len3(x, y, z);
    if (l != 1.0f) { x = __fdiv_rn(x, l); y = __fdiv_rn(y, l); z = __fdiv_rn(z, l); }  // 0/0 -> NaN like hat()
    normals[3 * i] = x; normals[3 * i + 1] = y; normals[3 * i + 2] = z;
  }
}

// ---------------------------------------------------------------------------------------------
// synthetic jittered-grid Voronoi volume (benchmark/test input)

__host__ __device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
  x += 0x9E3779B97F4A7C15ull;
  unsigned long long z = x;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

struct SynthArgs {
  void* dst;
  unsigned long long n;  // voxels of the block
  uint32_t sx, sy, sz;   // block shape (logical)
  uint32_t ox, oy, oz;   // block origin inside the full volume
  uint32_t gx, gy, gz;   // cells per axis of the full volume
  uint32_t pitch;
  unsigned long long seed;
  int c_order;
};

template <typename L>
__global__ void __launch_bounds__(256) k_synth_voronoi(const SynthArgs a) {
  unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (; i < a.n; i += stride) {
    uint32_t lx, ly, lz;
    if (a.c_order) { lz = (uint32_t)(i % a.sz); unsigned long long t = i / a.sz; ly = (uint32_t)(t % a.sy); lx = (uint32_t)(t / a.sy); }
    else           { lx = (uint32_t)(i % a.sx); unsigned long long t = i / a.sx; ly = (uint32_t)(t % a.sy); lz = (uint32_t)(t / a.sy); }
    const long long x = (long long)lx + a.ox, y = (long long)ly + a.oy, z = (long long)lz + a.oz;
    const int bi = (int)(x / a.pitch), bj = (int)(y / a.pitch), bk = (int)(z / a.pitch);
    long long best_d = 0x7FFFFFFFFFFFFFFFll;
    unsigned long long best_c = 0;
    for (int dk = -1; dk <= 1; ++dk)
      for (int dj = -1; dj <= 1; ++dj)
        for (int di = -1; di <= 1; ++di) {
          const int ni = bi + di, nj = bj + dj, nk = bk + dk;
          if (ni < 0 || nj < 0 || nk < 0 || ni >= (int)a.gx || nj >= (int)a.gy || nk >= (int)a.gz) continue;
          const unsigned long long c = (unsigned long long)ni + (unsigned long long)a.gx * ((unsigned long long)nj + (unsigned long long)a.gy * nk);
          const unsigned long long h = splitmix64(c ^ a.seed);
          const long long sx = (long long)ni * a.pitch + (long long)(((h & 0xFFFFull) * a.pitch) >> 16);
          const long long sy = (long long)nj * a.pitch + (long long)((((h >> 16) & 0xFFFFull) * a.pitch) >> 16);
          const long long sz = (long long)nk * a.pitch + (long long)((((h >> 32) & 0xFFFFull) * a.pitch) >> 16);
          const long long d = (x - sx) * (x - sx) + (y - sy) * (y - sy) + (z - sz) * (z - sz);
          if (d < best_d || (d == best_d && c < best_c)) { best_d = d; best_c = c; }
        }
    unsigned long long lab = sizeof(L) == 8 ? (splitmix64(best_c + 1ull) | 1ull) : (best_c + 1ull);
    static_cast<L*>(a.dst)[i] = (L)lab;
  }
}

}  // namespace zm

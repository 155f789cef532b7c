// zmesh_b200 device code: multi-label marching cubes for sm_100a.
//
// Replaces marching_cubes::marche + CMesher::triangles2mesh of the reference
// (zi_lib/zi/mesh/marching_cubes.hpp:291-445, zmesh/cMesher.hpp:96-166) with a dedup-free
// formulation:
//
//   * a vertex of label L is a voxel-grid edge (two axis-adjacent voxels) with exactly one
//     endpoint == L.  Each grid edge is OWNED by its lower voxel, so every vertex is produced
//     exactly once -- no hashing of vertices, no sort, no unique pass.
//   * voxel u owns up to 6 vertex slots: slot 2d+0 = (edge u -> u+d, label of u),
//     slot 2d+1 = (same edge, label of u+d), d = memory axis 0 (fastest) .. 2 (slowest).
//   * a ROW SEGMENT (32 voxels of one tile along the fastest axis) stores its slots as six 32-bit
//     bit planes (plane s, bit i = voxel i owns slot s) plus the spatial id of its first slot:
//     `rowinfo` = 8 words per segment (= 1 byte per voxel).  Slots are numbered plane-major inside
//     the segment:  g = rowbase + sum_{p<s} popc(plane_p) + popc(plane_s & ((1 << i) - 1)),
//     so a lookup is two shared loads and a popc;  perm[g] = index of the vertex inside its
//     label's vertex list.
//
// Pass 1 (k_classify, the only kernel that reads the label volume): one CTA per tile of 32 x 8 x 8 voxels; the
// (+1 halo) label region is staged in shared memory by TMA (cp.async.bulk.tensor.3d, zero fill outside the volume =
// the `close` border for free), the region of the tile one layer up is pulled into L2.  A tile whose whole region is
// one value is recognised with 16-byte compares and costs nothing else.  A tile with exactly two labels takes the
// bit-parallel path (tile_body_k2: one mask per staged row, everything else is arithmetic on whole rows); the general
// path compacts the active voxels, enumerates the distinct labels of each cube and keeps per-(tile,label) counts in a
// shared-memory table; tiles that overflow the per-tile staging are redone by the MODE 1 launch.  One global
// reservation per (tile,label).  Outputs, all label-free:
//   rowinfo[row segment], perm[g] (4 B/vertex), vinfo[g] (4 B/vertex: voxel-in-tile, slot,
//   tile-local label index), rec[] (8 B per (label,cube) pair: voxel-in-tile | case | tile-local
//   label index, first face row inside the (tile,label) block), tl[] (per (tile,label): label slot
//   + face base), hdr[] work list of the non-empty tiles.
// Scan (k_scan_*) turns per-label counts into per-label output offsets; k_tl_fixup folds them into
// tl[].  Pass 2 (k_emit) never touches labels again: persistent warp-specialised CTAs walk the work list, stage the
// rowinfo region of a tile with one TMA load, and write faces (uint32 triples, one triangle per
// lane) [+ face normals: one 16-byte vector atomic per corner] and float32 vertices in the final form
// fl32(fl32(fl32(res*k) [+ off]) / 2).  k_pack_precomputed lays out the Neuroglancer objects; k_export_directory /
// k_import_directories / k_export_plane / k_import_plane_normals are the device ends of the multi-GPU exchanges.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "mc_tables.h"

namespace zm {

typedef unsigned long long u64;

// device copies of the case tables (filled once per process by prepare_device)
__device__ uint8_t TRI_COUNT_D[256];
__device__ u64 TRI_NIBBLES_D[256];

#ifndef ZM_S1_ROWMASK
#define ZM_S1_ROWMASK 1  // 1: row-mask formulation of pass 1's S1 (edge_rows); 0: marching scan (scan_tile)
#endif

#ifndef ZM_S2_TWOWARPS
#define ZM_S2_TWOWARPS 1  // 1 (needs ZM_S1_ROWMASK): phase B + the row prefixes run in warps 0-1 only, one thread per row
#endif

#ifndef ZM_S3_REDRAW
#define ZM_S3_REDRAW 0  // 1: S3 lanes that draw label 0 draw their next label in the same iteration
#endif

#ifndef ZM_K2_PATH
#define ZM_K2_PATH 1  // 1: MODE 0 tiles whose staged region holds exactly two labels take the bit-parallel two-label path (tile_body<.., K2 = true>)
#endif

#ifndef ZM_S3_ATOMIC_RANK
#define ZM_S3_ATOMIC_RANK 1  // 1: S3 ranks vertices / face rows with one shared atomic per working lane (c1 k_classify 1.24 -> 1.05 ms, c5 31.5 -> 30.2); 0: match_any + REDUX + five ballots
#endif

constexpr int TF = 32;  // tile extent along the memory-fastest axis (= one warp per row)
constexpr int TM = 8;
constexpr int TS = 8;
constexpr int NT = 256;  // threads per CTA: warp w handles the s-plane w of the tile
constexpr int NW = NT / 32;
constexpr int RM = TM + 1, RS = TS + 1;  // staged label region (halo 1 on the high side)
constexpr int TILE_VOX = TF * TM * TS;
constexpr int NROWS = TM * TS;           // row segments per tile
constexpr int RI_WORDS = 8;              // rowinfo words per row segment: planes 0..5, rowbase, spare
static_assert(TS == NW, "one warp per s-plane of the tile");

// staged row length: a multiple of 16 bytes (TMA box constraint) that holds TF+1 voxels plus, for
// `close`, the 16/sizeof(L)-1 extra leading columns an aligned box start costs (see stage coords)
template <typename L> struct RowPad { static constexpr int value = ((TF + 1) * (int)sizeof(L) + 15) / 16 * 16 / (int)sizeof(L); };

enum : uint32_t {
  FLAG_HASH_FULL = 1u,  // global label table too small -> host grows it and reruns pass 1
  FLAG_CAP = 2u,        // perm / record / tile-label capacity guess too small -> host reruns pass 1
  FLAG_INTERNAL = 4u,   // invariant violated
  FLAG_DIR = 8u         // slab sharding: a shard's label directory did not fit the exchange buffer
};

// capacities of the per-tile shared-memory structures.  MODE 0 covers ordinary segmentations;
// tiles that overflow it are queued and redone by the MODE 1 launch, which processes a tile as two
// half tiles (4 s-planes each) whose capacities are the hard maxima (so it cannot overflow).
template <int MODE> struct Caps;
#ifndef ZM_CAP0
#define ZM_CAP0 2048     // MODE 0 capacity (vertex slots / records per tile) of the shared staging arrays
#endif
#ifndef ZM_CAP0_U64
#define ZM_CAP0_U64 1280     // the same for 8-byte labels: 43.3 KB per CTA, which buys the fifth resident CTA (ZM_U64_CTAS)
#endif
#ifndef ZM_U64_CTAS
#define ZM_U64_CTAS 5    // resident CTAs per SM k_classify<8-byte labels, MODE 0> is register-bounded for
#endif
#ifndef ZM_U32_CTAS
#define ZM_U32_CTAS 5    // the same for 1/2/4-byte labels
#endif
template <> struct Caps<0> { static constexpr int LT = 128, VCAP = ZM_CAP0, RCAP = ZM_CAP0, PROBES = 16, HALVES = 1; };
template <> struct Caps<1> { static constexpr int LT = 2048, VCAP = 6 * TILE_VOX / 2, RCAP = 8 * TILE_VOX / 2, PROBES = 2048, HALVES = 2; };
// staging capacities by label width as well
template <typename L, int MODE> struct TileCap {
  static constexpr int V = (MODE == 0 && sizeof(L) == 8) ? ZM_CAP0_U64 : Caps<MODE>::VCAP;
  static constexpr int R = (MODE == 0 && sizeof(L) == 8) ? ZM_CAP0_U64 : Caps<MODE>::RCAP;
};

struct VolParams {
  const void* data;        // device pointer, memory order (f fastest, m, s slowest)
  uint32_t nf, nm, ns;     // input extents
  uint32_t Ef, Em, Es;     // extended extents = n + 2*pad (close => virtual zero border)
  uint32_t Efp;            // ntf * TF: row pitch of the boundary-plane exchange buffers
  uint32_t pad;            // 1 when close
  uint32_t ntf, ntm, nts;  // tiles per axis
  uint32_t ox, oy, oz;     // shard origin in logical voxels (added to keys)
  uint32_t use_tma;
  // slab sharding along s (multi-GPU): planes [0, Es_own) own vertex slots; when Es_own == Es - 1
  // the top plane belongs to the next shard and is only read as cube corners.  s_shift maps an
  // extended local plane to a plane of the buffer (-pad for an unsharded volume).
  uint32_t Es_own;
  int32_t s_shift;
};

struct LabelTable {  // global open-addressing table, key 0 = empty (label 0 is never meshed)
  u64* keys;
  u64* cnt;  // low 32: vertices, high 32: triangles
  uint32_t mask;  // capacity - 1
};

struct __align__(16) TileHdr {  // one per non-empty (half) tile, in work-list order
  u64 recbase;
  uint32_t gbase;
  uint32_t tlbase;
  uint16_t nslots, nrec, nlab, pad;
  uint32_t tile, pad2;
};
static_assert(sizeof(TileHdr) == 32, "TileHdr is loaded as two 16-byte words");

struct __align__(16) TLEntry {  // pass 1: a = label slot | face base in label << 32;  after k_tl_fixup:
  u64 a, b;                     // a = first vertex row of the label | index offset of the label << 32
};                              // (vertices of the label on earlier shards), b = first face row of this (tile,label)

// control block (device): cursors and flags
struct Control {
  u64 cur_perm, cur_rec, cur_tl, cur_tri;
  uint32_t work_count, dense_count, flags, pad;
  u64 totals[4];  // n_labels, V_total, T_total, spare (written by the scan)
};

struct Pass1Args {
  LabelTable ht;
  Control* ctl;
  uint32_t* rowinfo;  // [Es][Em][ntf][RI_WORDS]
  uint32_t* perm;     // [capV]
  uint32_t* vinfo;    // [capV]: voxel-in-tile | slot << 11 | tile-local label index << 14
  u64* rec;           // [capR]: voxel-in-tile | case << 11 | tile-local label index << 19 | first face row in (tile,label) << 32
  TLEntry* tl;        // [capL]
  TileHdr* hdr;       // [ntiles * HALVES] work list of non-empty tiles
  uint32_t* dense_list;  // [ntiles]
  u64 capV, capR, capL;
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
// corner index at memory offset (+f), (+m), (+s)
template <bool CO> __host__ __device__ constexpr int corner_plus_f() { return CO ? 3 : 1; }
template <bool CO> __host__ __device__ constexpr int corner_plus_m() { return 4; }
template <bool CO> __host__ __device__ constexpr int corner_plus_s() { return CO ? 1 : 3; }

// per edge: bits 0-3 row delta (ds * RM + dm) of the owner voxel, bit 4 its f offset, bits 12-14
// 2*axis, bits 16-18 owner corner.  The midpoint M = corner_a + corner_b (half-voxel units): the
// axis is where M == 1, the owner (lower endpoint) is M >> 1.
template <bool CO>
__host__ __device__ constexpr uint32_t edge_info(int e) {
  int a = edge_a(e), b = edge_b(e);
  int mf = corner_df<CO>(a) + corner_df<CO>(b);
  int mm = corner_dm<CO>(a) + corner_dm<CO>(b);
  int ms = corner_ds<CO>(a) + corner_ds<CO>(b);
  int axis = mf == 1 ? 0 : (mm == 1 ? 1 : 2);
  int of = mf >> 1, om = mm >> 1, os = ms >> 1;
  int oc = 0;
  for (int n = 0; n < 8; ++n)
    if (corner_df<CO>(n) == of && corner_dm<CO>(n) == om && corner_ds<CO>(n) == os) oc = n;
  int drow = os * RM + om;
  return (uint32_t)(drow | (of << 4) | ((2 * axis) << 12) | (oc << 16));
}

// ---------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + TMA (cp.async.bulk.tensor)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(u64* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tmap, u64* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* tmap, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(tmap), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(u64* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}

// ---------------------------------------------------------------------------------------------

__device__ __forceinline__ uint32_t hash_label(u64 x) {
  x ^= x >> 33;
  x *= 0xff51afd7ed558ccdull;
  x ^= x >> 29;
  return (uint32_t)x ^ (uint32_t)(x >> 32);
}
__device__ __forceinline__ uint32_t hash_label32(uint32_t x) {
  x *= 0x9E3779B1u;
  return x ^ (x >> 15);
}
template <typename L> __device__ __forceinline__ uint32_t hash_any(L x) {
  if (sizeof(L) == 8) return hash_label((u64)x);
  return hash_label32((uint32_t)x);
}

// Global table: find-or-insert.  Returns the slot, or -1 (and raises FLAG_HASH_FULL).
__device__ __forceinline__ int gtab_insert(const LabelTable& ht, u64 label, uint32_t* flags) {
  uint32_t h = hash_label(label) & ht.mask;
  const uint32_t max_probe = ht.mask < 4095u ? ht.mask + 1u : 4096u;
  for (uint32_t p = 0; p < max_probe; ++p) {
    u64 old = ht.keys[h];
    if (old == label) return (int)h;
    if (old == 0ull) {
      old = atomicCAS(&ht.keys[h], 0ull, label);
      if (old == 0ull || old == label) return (int)h;
    }
    h = (h + 1u) & ht.mask;
  }
  atomicOr(flags, FLAG_HASH_FULL);
  return -1;
}

// CTA-local table in shared memory (keys only; 0 = empty).  Returns the slot or -1 when PROBES
// probes found neither the label nor a free slot.
template <int LT, int PROBES>
__device__ __forceinline__ int ltab_insert(u64* keys, u64 label, uint32_t h) {
  h &= (LT - 1);
#pragma unroll 1
  for (int p = 0; p < PROBES; ++p) {
    u64 k = *(volatile u64*)&keys[h];
    if (k == label) return (int)h;
    if (k == 0ull) {
      u64 old = atomicCAS(&keys[h], 0ull, label);
      if (old == 0ull || old == label) return (int)h;
    }
    h = (h + 1u) & (LT - 1);
  }
  return -1;
}

// exclusive in-warp prefix and warp total of a per-lane count c in [0, 7], via three ballots
__device__ __forceinline__ uint32_t warp_prefix3(uint32_t c, uint32_t ltm, uint32_t& total) {
  const uint32_t b0 = __ballot_sync(0xffffffffu, c & 1u);
  const uint32_t b1 = __ballot_sync(0xffffffffu, c & 2u);
  const uint32_t b2 = __ballot_sync(0xffffffffu, c & 4u);
  total = __popc(b0) + 2u * __popc(b1) + 4u * __popc(b2);
  return __popc(b0 & ltm) + 2u * __popc(b1 & ltm) + 4u * __popc(b2 & ltm);
}
// same, restricted to the lanes of `grp` (a __match_any_sync group containing this lane)
__device__ __forceinline__ uint32_t group_prefix3(uint32_t c, uint32_t grp, uint32_t ltm, uint32_t& total) {
  const uint32_t b0 = __ballot_sync(0xffffffffu, c & 1u) & grp;
  const uint32_t b1 = __ballot_sync(0xffffffffu, c & 2u) & grp;
  const uint32_t b2 = __ballot_sync(0xffffffffu, c & 4u) & grp;
  total = __popc(b0) + 2u * __popc(b1) + 4u * __popc(b2);
  return __popc(b0 & ltm) + 2u * __popc(b1 & ltm) + 4u * __popc(b2 & ltm);
}

// ---------------------------------------------------------------------------------------------
// pass 1: classify + count + reserve

template <typename L, int MODE>
struct __align__(128) P1Smem {
  static constexpr int RFP = RowPad<L>::value;
  union {
    L lab[RS * RM * RFP];            // TMA destination: must stay first (128-byte aligned); live until S3 ends
    struct {                         // live from S5 on
      u64 tla[Caps<MODE>::LT];       // tl entry (word a) of the label
      uint32_t lvb[Caps<MODE>::LT];  // first rank of the tile's vertices inside the label
      uint16_t cidx[Caps<MODE>::LT]; // tile-local (compact) label index
    };
  };
  uint32_t pl[NROWS][RI_WORDS];    // per row segment: slot bit planes 0..5; [6] slots of the row, then (S2) first slot of the row in the tile; [7] active-cube mask
  u64 lkeys[Caps<MODE>::LT];
  u64 mbar;
  u64 recbase;
  uint32_t lcnt[Caps<MODE>::LT];   // low 16: vertices of the label in this tile, high 16: triangles
  // per tile-local slot.  MODE 0: local rank (11 bits) | table slot << 11 (7 bits) | voxel-in-tile << 18 | slot << 29;
  // MODE 1: local rank << 12 | table slot, voxel-in-tile | slot << 11 in pstage
  alignas(16) uint32_t vstage[TileCap<L, MODE>::V];  // (ZM_S1_ROWMASK: holds the row masks of edge_rows until S3)
  uint32_t rstage[TileCap<L, MODE>::R];  // voxel-in-tile | case << 11 | table slot << 19
  uint32_t tc[4];                  // tile coordinates (tf, tm, ts)
  uint32_t wtot[NW];               // per s-plane: slots | active voxels << 16
  uint32_t nrec, overflow, ok1, ok2, gbase, tlbase, ci, ttot;
  uint16_t alist[TILE_VOX];        // active voxels (voxel-in-tile), compacted
  uint16_t actpre[NROWS];          // active voxels in earlier rows
  uint16_t pstage[MODE == 0 ? 2 : TileCap<L, MODE>::V];
  uint16_t rtoff[TileCap<L, MODE>::R];   // per record: first face row inside the (tile,label) block
  alignas(8) uint8_t pp8[NROWS][8];  // slots of the row in lower planes
};
static_assert(sizeof(P1Smem<u64, 1>) <= 227 * 1024 && sizeof(P1Smem<uint8_t, 1>) <= 227 * 1024, "dense mode must fit one SM");
static_assert((sizeof(P1Smem<uint32_t, 0>) + 1024) * ZM_U32_CTAS <= 228 * 1024, "MODE 0 / 4-byte labels: ZM_U32_CTAS CTAs per SM (1 KB reserved each)");
static_assert(Caps<0>::VCAP <= 2048 && ZM_CAP0_U64 <= 2048 && Caps<0>::LT <= 128, "MODE 0 vstage packing");
static_assert((sizeof(P1Smem<u64, 0>) + 1024) * ZM_U64_CTAS <= 228 * 1024, "MODE 0 / 8-byte labels: ZM_U64_CTAS CTAs per SM (1 KB reserved each)");

extern __shared__ __align__(128) unsigned char zm_dyn_smem[];

// first staged column / row / plane of a tile in input coordinates.  TMA needs a 16-byte aligned
// start along f: with the `close` border the box starts ALIGN elements (not 1) before the tile and
// the region sits ALIGN-1 columns into the staged rows.
template <typename L>
__device__ __forceinline__ void stage_origin(const VolParams& vp, uint32_t tf, uint32_t tm, uint32_t ts, int& c0, int& c1, int& c2) {
  constexpr int ALIGN = 16 / (int)sizeof(L);
  c0 = (int)(tf * TF) - (vp.pad ? ALIGN : 0);
  c1 = (int)(tm * TM) - (int)vp.pad;
  c2 = (int)(ts * TS) + vp.s_shift;
}

// Launch order of the MODE 0 tiles (a 3-d grid: x = tf + ntf * row-in-group, y = ts, z = group): f fastest, then a
// group of TM_GROUP tile rows, then ALL s layers, then the next group.  The halo plane a tile shares with its s-neighbour is then re-read ntf * TM_GROUP tiles
// later (8 MB of labels for c5) instead of ntf * ntm tiles later (268 MB > L2): ncu measured 12.6 % more
// DRAM reads than the volume with the plain order.
constexpr uint32_t TM_GROUP = 8;
// thread 0: publish the coordinates of a tile and (TMA path) start the load of its region
template <typename L, int MODE>
__device__ __forceinline__ void begin_tile(const VolParams& vp, const CUtensorMap* tmap, P1Smem<L, MODE>& S, uint32_t tf,
                                           uint32_t tm, uint32_t ts) {
  S.tc[0] = tf; S.tc[1] = tm; S.tc[2] = ts;
  if (vp.use_tma) {
    int c0, c1, c2;
    stage_origin<L>(vp, tf, tm, ts, c0, c1, c2);
    // MODE 1 re-stages a region whose memory lvb / cidx / tla were written to through the generic proxy
    if (MODE == 1) fence_proxy_async();
    mbar_expect_tx(&S.mbar, (uint32_t)(sizeof(L) * RS * RM * P1Smem<L, MODE>::RFP));
    tma_load_3d(S.lab, tmap, &S.mbar, c0, c1, c2);
  }
}

// volumes TMA cannot address (row pitch or base not 16-byte aligned): the same box with plain loads
template <typename L, int MODE>
__device__ __forceinline__ void stage_plain(const VolParams& vp, P1Smem<L, MODE>& S, uint32_t tf, uint32_t tm, uint32_t ts) {
  constexpr int RFP = P1Smem<L, MODE>::RFP;
  const L* __restrict__ src = static_cast<const L*>(vp.data);
  int c0, c1, c2;
  stage_origin<L>(vp, tf, tm, ts, c0, c1, c2);
  for (int i = threadIdx.x; i < RFP * RM * RS; i += NT) {
    const int lf = i % RFP;
    const int t = i / RFP;
    const int lm = t % RM, ls = t / RM;
    const uint32_t jf = (uint32_t)(c0 + lf), jm = (uint32_t)(c1 + lm), js = (uint32_t)(c2 + ls);  // wraps when < 0
    L v = 0;
    if (jf < vp.nf && jm < vp.nm && js < vp.ns) v = src[((size_t)js * vp.nm + jm) * vp.nf + jf];
    S.lab[i] = v;
  }
}

// true iff every staged element equals the first element of the region (then no cube of the tile is active and
// no voxel owns a slot); 16-byte compares (loads issued back to back), one barrier
template <typename L, int MODE>
__device__ __forceinline__ bool region_uniform(const P1Smem<L, MODE>& S, const L* lab, L& first) {
  constexpr int NQ = (int)sizeof(L) * RS * RM * P1Smem<L, MODE>::RFP / 16;
  constexpr int PASSES = (NQ + NT - 1) / NT;
  const uint4* q = reinterpret_cast<const uint4*>(S.lab);
  first = lab[0];
  uint4 ref;
  if (sizeof(L) == 8) {
    const u64 r = (u64)first;
    ref = make_uint4((uint32_t)r, (uint32_t)(r >> 32), (uint32_t)r, (uint32_t)(r >> 32));
  } else {
    uint32_t r = (uint32_t)first;
    if (sizeof(L) == 2) r = __byte_perm(r, r, 0x1010);
    if (sizeof(L) == 1) r = __byte_perm(r, r, 0x0000);
    ref = make_uint4(r, r, r, r);
  }
  uint4 v[PASSES];
#pragma unroll
  for (int j = 0; j < PASSES; ++j) {
    const int i = threadIdx.x + j * NT;
    v[j] = (j < NQ / NT || i < NQ) ? q[i] : ref;
  }
  uint32_t acc = 0;
#pragma unroll
  for (int j = 0; j < PASSES; ++j) acc |= (v[j].x ^ ref.x) | (v[j].y ^ ref.y) | (v[j].z ^ ref.z) | (v[j].w ^ ref.w);
  return __syncthreads_and(acc == 0u) != 0;
}

// all-zero rowinfo for the row segments of planes [h0, h0 + nh) of a tile without slots
__device__ __forceinline__ void zero_rows(const VolParams& vp, const Pass1Args& o, uint32_t tf, uint32_t tm, uint32_t ts,
                                          int h0, int nh) {
  const int tid = threadIdx.x;
  const int row = tid >> 1;
  if (tid < 2 * NROWS && row >= h0 * TM && row < (h0 + nh) * TM) {
    const uint32_t rs = ts * TS + row / TM, rm = tm * TM + row % TM;
    if (rs < vp.Es_own && rm < vp.Em)
      reinterpret_cast<uint4*>(o.rowinfo + (((size_t)rs * vp.Em + rm) * vp.ntf + tf) * RI_WORDS)[tid & 1] = make_uint4(0u, 0u, 0u, 0u);
  }
}

// S1 of a tile.  Thread (lane = f, warp = s) marches over m keeping the four labels of the
// previous row in registers: per step 4 shared loads and 4 label compares give the cube's
// uniformity and the voxel's slot mask; six ballots turn the masks of a row into its bit planes.
// No atomics and no cross-step dependencies besides the marching registers.
// INTERIOR tiles (no volume boundary within reach) skip all validity logic.
template <typename L, int MODE, bool INTERIOR>
__device__ __forceinline__ bool scan_tile(const VolParams& vp, P1Smem<L, MODE>& S, const L* lab,
                                          uint32_t ef0, uint32_t em0, uint32_t es0) {
  constexpr int RFP = P1Smem<L, MODE>::RFP;
  constexpr uint32_t FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31, ls = threadIdx.x >> 5;
  const uint32_t ef = ef0 + lane, es = es0 + ls;
  const bool okf = INTERIOR || ef < vp.Ef, oks = INTERIOR || es < vp.Es_own;
  const bool nf1 = INTERIOR || ef + 1 < vp.Ef, ns1 = INTERIOR || es + 1 < vp.Es;
  const L* p = lab + (ls * RM) * RFP + lane;
  L a = p[0], af = p[1], as_ = p[RM * RFP];
  bool nef = a != af, nes = a != as_;
  bool eq_row = !nef && !nes && a == p[RM * RFP + 1];  // the 4 corners of row j agree
  bool za = a != 0, zf = af != 0, zs = as_ != 0;
  bool any = false;
#pragma unroll
  for (int j = 0; j < TM; ++j) {
    p += RFP;
    const L am = p[0], amf = p[1], ams = p[RM * RFP], amfs = p[RM * RFP + 1];
    const bool nef2 = am != amf, nes2 = am != ams;
    const bool eq_row2 = !nef2 && !nes2 && am == amfs;
    const bool nem = a != am;
    const bool zm = am != 0;
    const bool okm = INTERIOR || em0 + j < vp.Em, nm1 = INTERIOR || em0 + j + 1 < vp.Em;
    const bool uniform = eq_row && eq_row2 && !nem;
    // which of its three grid edges carry vertices (an edge with different endpoint labels)
    bool pf = nef, pm = nem, ps = nes, act;
    if (INTERIOR) {
      act = !uniform;  // (an edge with different labels makes the cube non-uniform)
    } else {
      const bool valid = okf && okm && oks;
      pf = pf && valid && nf1; pm = pm && valid && nm1; ps = ps && valid && ns1;
      act = (pf && (za || zf)) || (pm && (za || zm)) || (ps && (za || zs)) || (valid && nf1 && nm1 && ns1 && !uniform);
    }
    const uint32_t ab = __ballot_sync(FULL, act);
    uint32_t b0 = 0, b1 = 0, b2 = 0, b3 = 0, b4 = 0, b5 = 0;
    if (ab) {  // (a voxel that owns a slot is active: rows without active voxels have empty planes)
      any = true;
      b0 = __ballot_sync(FULL, pf && za); b1 = __ballot_sync(FULL, pf && zf);
      b2 = __ballot_sync(FULL, pm && za); b3 = __ballot_sync(FULL, pm && zm);
      b4 = __ballot_sync(FULL, ps && za); b5 = __ballot_sync(FULL, ps && zs);
    }
    if (lane == 0) {
      uint32_t* row = S.pl[ls * TM + j];
      *reinterpret_cast<uint4*>(row) = make_uint4(b0, b1, b2, b3);
      *reinterpret_cast<uint4*>(row + 4) = make_uint4(b4, b5, 0u, ab);
    }
    a = am; af = amf; as_ = ams;
    nef = nef2; nes = nes2; eq_row = eq_row2;
    za = zm; zf = amf != 0; zs = ams != 0;
  }
  return any;
}

// S1, row-mask formulation (ZM_S1_ROWMASK): phase A turns every staged row (all RS * RM of them, the halo
// rows included) into five 32-bit masks with ONE pass of three label compares per voxel --
//   Ef: voxel != its +f neighbour,  Em: != +m neighbour,  Es: != +s neighbour,  Z: voxel != 0,  Zf: +f neighbour != 0
// -- and everything per cube / per slot is then bit arithmetic on whole rows (phase B, in tile_body): a cube is
// uniform iff the 7 edges of a spanning tree of its corners are (4 f-edges, the m-edges at f of the two planes, the
// s-edge at (f, m)).  The masks live in vstage, which is dead until S3.
constexpr int EM_WORDS = 8;  // words per staged row in the mask array (5 used)
template <typename L, int MODE>
__device__ __forceinline__ void edge_rows(P1Smem<L, MODE>& S, const L* lab) {
  constexpr int RFP = P1Smem<L, MODE>::RFP;
  constexpr uint32_t FULL = 0xffffffffu;
  static_assert(TileCap<L, MODE>::V >= RS * RM * EM_WORDS, "edge masks overlay vstage");
  static_assert(NW == TM, "warp w marches over s along the rows m = w");
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t* const em = S.vstage;
  {  // rows m = warp (< TM, so row m + 1 is staged), s = 0 .. RS-1: the +s neighbour is the next step's own voxel
    const L* p = lab + warp * RFP + lane;
    uint32_t* e = em + warp * EM_WORDS;
    L a = p[0];
#pragma unroll
    for (int ls = 0; ls < RS; ++ls) {
      const L af = p[1], am = p[RFP];
      const L as_ = ls < RS - 1 ? p[RM * RFP] : a;  // (no plane above the halo plane: Es unused there)
      const uint32_t ef = __ballot_sync(FULL, a != af), emm = __ballot_sync(FULL, a != am), es = __ballot_sync(FULL, a != as_);
      const uint32_t z = __ballot_sync(FULL, a != 0), zf = __ballot_sync(FULL, af != 0);
      if (lane == 0) {
        *reinterpret_cast<uint4*>(e) = make_uint4(ef, emm, es, z);
        e[4] = zf;
      }
      a = as_;
      p += RM * RFP;
      e += RM * EM_WORDS;
    }
  }
  // halo rows m = TM (only ever the "+m" row of a cube: Ef and Z are all that is read): s = warp, and s = TS for warp 0
  for (int ls = warp; ls < RS; ls += NW) {
    const L* p = lab + (ls * RM + TM) * RFP + lane;
    const L a = p[0], af = p[1];
    const uint32_t ef = __ballot_sync(FULL, a != af), z = __ballot_sync(FULL, a != 0);
    if (lane == 0) *reinterpret_cast<uint4*>(em + (ls * RM + TM) * EM_WORDS) = make_uint4(ef, 0u, 0u, z);
  }
}

template <typename L> __device__ __forceinline__ L shfl_label(L v, int src) {
  if (sizeof(L) == 8) return (L)__shfl_sync(0xffffffffu, (u64)v, src);
  return (L)__shfl_sync(0xffffffffu, (uint32_t)v, src);
}

// Two-label path, S1: if the staged region holds (at most) two labels A and B, ONE 33-bit mask per staged row
// -- bit f = (voxel f carries A) -- says everything: edges are where the mask changes, the non-zero masks follow
// from which of A / B is the background.  One compare pair + one ballot per row instead of edge_rows' three label
// compares + five ballots per voxel.  A is the region's first label; every warp discovers the other label of ITS
// rows on the way (the first voxel that is not A) and checks its rows against it; the candidates of the warps are
// compared after the barrier (k2_second_label).  Returns true when the warp met a third label (the tile then takes
// the general path).  The masks live in S.lkeys[2 ..] (64-bit, bit 32 = the +f halo column), the candidates in rstage.
constexpr int K2_MASK0 = 2;  // first lkeys entry used for the row masks
template <typename L, int MODE>
__device__ __forceinline__ bool k2_rows(P1Smem<L, MODE>& S, const L* lab, const L A) {
  constexpr int RFP = P1Smem<L, MODE>::RFP;
  constexpr uint32_t FULL = 0xffffffffu;
  static_assert(Caps<MODE>::LT >= K2_MASK0 + RS * RM, "row masks overlay the label keys");
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t* const mw = reinterpret_cast<uint32_t*>(S.lkeys + K2_MASK0);
  constexpr int NR = RS * RM, PASSES = (NR + NW - 1) / NW;
  static_assert(PASSES <= 32, "one lane per row of the warp for the halo column");
  bool bad = false, hasB = false;
  L B = A;  // (until the warp meets another label)
#pragma unroll
  for (int j = 0; j <= PASSES; ++j) {
    // passes 0 .. PASSES-1: lane f of row warp + j * NW; the last pass: the +f halo column of the warp's rows
    const int r = j < PASSES ? warp + j * NW : warp + lane * NW;
    const bool in = j < PASSES ? (j < NR / NW || r < NR) : (lane < PASSES && r < NR);
    if (j < PASSES && !in) continue;  // (warp-uniform: only pass PASSES-1 is partial)
    const L a = in ? lab[r * RFP + (j < PASSES ? lane : TF)] : A;
    const bool pa = a == A;
    const uint32_t m = __ballot_sync(FULL, pa);
    if (m != FULL && !hasB) {  // (warp-uniform) the warp's second label
      B = shfl_label<L>(a, __ffs(~m) - 1);
      hasB = true;
    }
    bad = bad || !(pa || a == B);
    if (j < PASSES) {
      if (lane == 0) mw[2 * r] = m;
    } else if (in) {
      mw[2 * r + 1] = pa ? 1u : 0u;
    }
  }
  if (lane == 0) reinterpret_cast<u64*>(S.rstage)[warp] = (u64)B;  // == A: none met
  return bad;
}

// after the barrier that follows k2_rows: 0 = no label but A in the used columns (nothing to mesh), 1 = exactly one
// other label (returned in B), 2 = the warps met different labels (general path)
template <typename L, int MODE>
__device__ __forceinline__ int k2_second_label(const P1Smem<L, MODE>& S, const L A, L& B) {
  constexpr uint32_t FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const L c = (L)reinterpret_cast<const u64*>(S.rstage)[lane & (NW - 1)];
  const uint32_t has = __ballot_sync(FULL, c != A);
  if (has == 0u) return 0;
  B = shfl_label<L>(c, __ffs(has) - 1);
  return __any_sync(FULL, c != A && c != B) ? 2 : 1;
}

enum : int { TILE_EMPTY = 0, TILE_DEFERRED = 1, TILE_DONE = 2, TILE_NOT_K2 = 3 };

// Everything after staging for planes [h0, h0 + nh) of a tile (all 8 in MODE 0).  The caller has
// staged the region and synchronised; the per-tile tables are cleared here.
template <typename L, bool CO, int MODE>
__device__ __forceinline__ int tile_body(const VolParams& vp, const Pass1Args& o, P1Smem<L, MODE>& S, const uint32_t tile,
                                         const uint32_t tf, const uint32_t tm, const uint32_t ts, const int h0, const int nh) {
  constexpr int RFP = P1Smem<L, MODE>::RFP;
  constexpr int LT = Caps<MODE>::LT, VCAP = TileCap<L, MODE>::V, RCAP = TileCap<L, MODE>::R, PROBES = Caps<MODE>::PROBES;
  constexpr uint32_t FULL = 0xffffffffu;
  constexpr int ALIGN = 16 / (int)sizeof(L);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t ltm = (1u << lane) - 1u;
  const uint32_t ef0 = tf * TF, em0 = tm * TM, es0 = ts * TS;
  const L* const lab = S.lab + (vp.pad ? ALIGN - 1 : 0);

  for (int i = tid; i < LT; i += NT) { S.lkeys[i] = 0ull; S.lcnt[i] = 0u; }
  if (tid == 0) { S.nrec = 0; S.overflow = 0; S.ci = 0; S.ttot = 0; }

#if ZM_S1_ROWMASK
  // ---- S1 (row masks): phase A: edge / non-zero masks of all staged rows; phase B: lane j < TM of warp w turns
  //      them into the bit planes and the active-cube mask of row (w, j), then S2's in-row prefixes ----
  edge_rows<L, MODE>(S, lab);
  __syncthreads();
  bool any = false;
  uint32_t packed = 0;  // slots of the row | active voxels of the row << 16   (of the thread's row)
#if ZM_S2_TWOWARPS
  const int pw = tid >> 3, pj = tid & 7;  // thread t < NROWS owns row t: plane pw, row pj
  if (warp < NROWS / 32) {
#else
  const int pw = warp, pj = lane;         // lane j < TM of warp w owns row j of plane w
  if (lane < TM) {
#endif
    const int row = pw * TM + pj;
    uint32_t b0 = 0, b1 = 0, b2 = 0, b3 = 0, b4 = 0, b5 = 0, act = 0;
    if (pw >= h0 && pw < h0 + nh) {
      const uint32_t* e0 = S.vstage + (pw * RM + pj) * EM_WORDS;
      const uint4 q = *reinterpret_cast<const uint4*>(e0);                     // Ef, Em, Es, Z of (m, s)
      const uint32_t zf = e0[4];
      const uint4 qm = *reinterpret_cast<const uint4*>(e0 + EM_WORDS);         // (m + 1, s)
      const uint4 qs = *reinterpret_cast<const uint4*>(e0 + RM * EM_WORDS);    // (m, s + 1)
      const uint32_t ef3 = e0[(RM + 1) * EM_WORDS];                            // Ef of (m + 1, s + 1)
      const uint32_t nonuni = q.x | qm.x | qs.x | ef3 | q.y | qs.y | q.z;     // cube has two different corners
      uint32_t pf = q.x, pm = q.y, ps = q.z, cube = nonuni;
      if (!(ef0 + TF + 1 <= vp.Ef && em0 + TM + 1 <= vp.Em && es0 + TS + 1 <= vp.Es)) {
        // volume boundary within reach: slots exist only on edges whose upper voxel is inside the (extended)
        // volume and whose lower voxel this shard owns; a cube needs all three upper neighbours
        const uint32_t nfv = vp.Ef - ef0;  // valid columns from ef0 (>= 1)
        const uint32_t VF = nfv >= 32u ? FULL : (1u << nfv) - 1u;          // ef < Ef
        const uint32_t NF1 = nfv >= 33u ? FULL : (1u << (nfv - 1u)) - 1u;  // ef + 1 < Ef
        const uint32_t em_ = em0 + (uint32_t)pj, es_ = es0 + (uint32_t)pw;
        const bool rowok = em_ < vp.Em && es_ < vp.Es_own;
        const bool nm1 = em_ + 1 < vp.Em, ns1 = es_ + 1 < vp.Es;
        pf = rowok ? pf & NF1 : 0u;
        pm = rowok && nm1 ? pm & VF : 0u;
        ps = rowok && ns1 ? ps & VF : 0u;
        cube = rowok && nm1 && ns1 ? nonuni & NF1 : 0u;
      }
      b0 = pf & q.w; b1 = pf & zf; b2 = pm & q.w; b3 = pm & qm.w; b4 = ps & q.w; b5 = ps & qs.w;
      act = b0 | b1 | b2 | b3 | b4 | b5 | cube;  // (interior tiles: an edge with different labels makes the cube non-uniform)
    }
    *reinterpret_cast<uint4*>(&S.pl[row][0]) = make_uint4(b0, b1, b2, b3);
    *reinterpret_cast<uint4*>(&S.pl[row][4]) = make_uint4(b4, b5, 0u, act);
    any = act != 0u;
    const uint32_t c0 = __popc(b0), c1 = c0 + __popc(b1), c2 = c1 + __popc(b2), c3 = c2 + __popc(b3);
    const uint32_t c4 = c3 + __popc(b4);
    *reinterpret_cast<uint2*>(S.pp8[row]) = make_uint2((c0 << 8) | (c1 << 16) | (c2 << 24), c3 | (c4 << 8));
    packed = (c4 + __popc(b5)) | ((uint32_t)__popc(act) << 16);
  }
#else
  // ---- S1: slot masks -> bit planes, active-cube masks ----
  bool any = false;
  if (warp >= h0 && warp < h0 + nh) {
    if (ef0 + TF + 1 <= vp.Ef && em0 + TM + 1 <= vp.Em && es0 + TS + 1 <= vp.Es)
      any = scan_tile<L, MODE, true>(vp, S, lab, ef0, em0, es0);
    else
      any = scan_tile<L, MODE, false>(vp, S, lab, ef0, em0, es0);
  } else if (lane < TM) {
    uint4* row = reinterpret_cast<uint4*>(S.pl[warp * TM + lane]);
    row[0] = make_uint4(0u, 0u, 0u, 0u);
    row[1] = make_uint4(0u, 0u, 0u, 0u);
  }
  __syncwarp();
  // ---- S2 (per warp, no serial section): in-row plane prefixes, row prefixes inside the plane ----
  uint32_t packed = 0;  // slots of the row | active voxels of the row << 16   (lane j < TM: row j of the plane)
  if (lane < TM) {
    const int row = warp * TM + lane;
    const uint4 q0 = *reinterpret_cast<const uint4*>(&S.pl[row][0]);
    const uint4 q1 = *reinterpret_cast<const uint4*>(&S.pl[row][4]);
    const uint32_t c0 = __popc(q0.x), c1 = c0 + __popc(q0.y), c2 = c1 + __popc(q0.z), c3 = c2 + __popc(q0.w);
    const uint32_t c4 = c3 + __popc(q1.x);
    *reinterpret_cast<uint2*>(S.pp8[row]) = make_uint2((c0 << 8) | (c1 << 16) | (c2 << 24), c3 | (c4 << 8));
    packed = (c4 + __popc(q1.y)) | ((uint32_t)__popc(q1.w) << 16);
  }
#endif
#if ZM_S1_ROWMASK && ZM_S2_TWOWARPS
  // rows 0..31 / 32..63 are scanned by warp 0 / 1; the bases stored before the barrier are relative to the 32-row
  // group, the readers add the first group's total (wtot[0]) for the rows of the second
  if (warp < NROWS / 32) {
    uint32_t rinc = packed;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t t = __shfl_up_sync(FULL, rinc, d);
      if (lane >= d) rinc += t;
    }
    const uint32_t ex = rinc - packed;
    S.pl[tid][6] = ex & 0xFFFFu;
    S.actpre[tid] = (uint16_t)(ex >> 16);
    if (lane == 31) S.wtot[warp] = rinc;
  }
  if (!__syncthreads_or(any ? 1 : 0)) return TILE_EMPTY;
  const uint32_t g0 = S.wtot[0];
  const uint32_t nslots = (g0 + S.wtot[1]) & 0xFFFFu, nact = (g0 + S.wtot[1]) >> 16;
  if (warp == 1) S.pl[tid][6] += g0 & 0xFFFFu;            // (read by S3 / the flush, after the next barrier)
  const uint32_t actbase = warp >= NW / 2 ? g0 >> 16 : 0u;  // rows of warps 4..7 = rows 32..63
#else
  uint32_t rinc = packed;
#pragma unroll
  for (int d = 1; d < TM; d <<= 1) {
    const uint32_t t = __shfl_up_sync(FULL, rinc, d);
    if (lane >= d) rinc += t;
  }
  if (lane == TM - 1) S.wtot[warp] = rinc;
  if (!__syncthreads_or(any ? 1 : 0)) return TILE_EMPTY;
  uint32_t nslots, nact;
  {
    const uint32_t t = S.wtot[lane & (NW - 1)];
    uint32_t winc = t;
#pragma unroll
    for (int d = 1; d < NW; d <<= 1) {
      const uint32_t x = __shfl_up_sync(FULL, winc, d, NW);
      if ((lane & (NW - 1)) >= d) winc += x;
    }
    const uint32_t wbase = __shfl_sync(FULL, winc - t, warp);
    const uint32_t total = __shfl_sync(FULL, winc, NW - 1);
    nslots = total & 0xFFFFu;
    nact = total >> 16;
    if (lane < TM) {
      const uint32_t ex = wbase + rinc - packed;
      S.pl[warp * TM + lane][6] = ex & 0xFFFFu;
      S.actpre[warp * TM + lane] = (uint16_t)(ex >> 16);
    }
  }
  const uint32_t actbase = 0u;
#endif
  if (MODE == 0 && nslots > (uint32_t)VCAP) {
    if (tid == 0) o.dense_list[atomicAdd(&o.ctl->dense_count, 1u)] = tile;
    return TILE_DEFERRED;
  }
  __syncwarp();
  // compaction of the active voxels of the warp's own plane
#pragma unroll
  for (int j = 0; j < TM; ++j) {
    const int row = warp * TM + j;
    const uint32_t mask = S.pl[row][7];
    if ((mask >> lane) & 1u) S.alist[actbase + S.actpre[row] + __popc(mask & ltm)] = (uint16_t)(row * TF + lane);
  }
  __syncthreads();

  // ---- S3: per active voxel: distinct labels of its cube -> counts, local ranks, records ----
  for (uint32_t base = warp * 32; base < nact; base += NT) {
    const uint32_t i = base + lane;
    const bool valid = i < nact;
    const uint32_t vidx = valid ? S.alist[i] : 0u;
    const int lf = vidx & 31, lm = (vidx >> 5) & 7, ls = vidx >> 8;
    // slots exist only on edges whose upper voxel is inside the (extended) volume
    const bool axf = ef0 + lf + 1 < vp.Ef, axm = em0 + lm + 1 < vp.Em, axs = es0 + ls + 1 < vp.Es;
    const uint32_t amask = (axf ? 0x03u : 0u) | (axm ? 0x0Cu : 0u) | (axs ? 0x30u : 0u);
    const bool cube = axf && axm && axs;
    const uint32_t row = vidx >> 5;
    const uint32_t ltf = (1u << lf) - 1u;
    const uint32_t rowpre = S.pl[row][6];
    L c[8];
#pragma unroll
    for (int n = 0; n < 8; ++n)
      c[n] = lab[((ls + corner_ds<CO>(n)) * RM + (lm + corner_dm<CO>(n))) * RFP + (lf + corner_df<CO>(n))];
    uint32_t acc = valid ? 0u : 0xFFu;
    while (__any_sync(FULL, acc != 0xFFu)) {
      const bool have = acc != 0xFFu;
      const int start = __ffs(~acc & 0xFFu) - 1;
      L label = c[0];
#pragma unroll
      for (int n = 1; n < 8; ++n) label = (n == start) ? c[n] : label;
      uint32_t msk = 0;
#pragma unroll
      for (int n = 0; n < 8; ++n) msk |= (c[n] == label ? 1u : 0u) << n;
      if (have) acc |= msk;
#if ZM_S3_REDRAW
      // lanes that drew the background label (never meshed) draw again right away: one more select + compare
      // instead of a whole iteration of bookkeeping for nothing.  Volumes without background never take the branch.
      {
        const bool redraw = have && label == 0 && acc != 0xFFu;
        if (__any_sync(FULL, redraw)) {
          const int start2 = __ffs(~acc & 0xFFu) - 1;
          L label2 = c[0];
#pragma unroll
          for (int n = 1; n < 8; ++n) label2 = (n == start2) ? c[n] : label2;
          uint32_t msk2 = 0;
#pragma unroll
          for (int n = 0; n < 8; ++n) msk2 |= (c[n] == label2 ? 1u : 0u) << n;
          if (redraw) { label = label2; msk = msk2; acc |= msk2; }
        }
      }
#endif
      const uint32_t cs = ~msk & 0xFFu;
      uint32_t nt = 0, mine = 0;
      if (have && label != 0) {
        if (cube) nt = __ldg(&TRI_COUNT_D[cs]);
        // neighbours (+f, +m, +s) carrying this label: the voxel owns the lower-side slot of an edge when it
        // carries the label and the neighbour does not, the upper-side slot when only the neighbour does
        const uint32_t nb = ((msk >> corner_plus_f<CO>()) & 1u) | (((msk >> corner_plus_m<CO>()) & 1u) << 2) |
                            (((msk >> corner_plus_s<CO>()) & 1u) << 4);
        mine = ((msk & 1u) ? (0x15u & ~nb) : (nb << 1)) & amask;
      }
      const uint32_t nv = __popc(mine);
      bool work = (nv | nt) != 0u;
      int hs = -1;
      if (work) {
        hs = ltab_insert<LT, PROBES>(S.lkeys, (u64)label, hash_any<L>(label));
        if (hs < 0) { S.overflow = 1u; work = false; }
      }
      // warp-aggregated add of (nv | nt << 16) to lcnt[hs]; the return value ranks the vertices
      // and gives the record its first face row inside the (tile,label) block
#if ZM_S3_ATOMIC_RANK
      // (experiment, not measured yet) let the shared-memory atomic unit serialise the lanes of a label: every
      // working lane adds its own counts and gets its own rank back -- no match_any / ballots / popcounts
      uint32_t old = 0;
      if (work) old = atomicAdd(&S.lcnt[hs], nv | (nt << 16));
      const uint32_t prev = 0u, pret = 0u;
#else
      const uint32_t grp = __match_any_sync(FULL, work ? (uint32_t)hs : 0xFFFFFFFFu);
      const uint32_t glt = grp & ltm;
      const uint32_t wv = work ? nv : 0u, wt = work ? nt : 0u;  // nv <= 3, nt <= 5
      const uint32_t tot = __reduce_add_sync(grp, wv | (wt << 16));
      const uint32_t v0 = __ballot_sync(FULL, wv & 1u), v1 = __ballot_sync(FULL, wv & 2u);
      const uint32_t t0 = __ballot_sync(FULL, wt & 1u), t1 = __ballot_sync(FULL, wt & 2u), t2 = __ballot_sync(FULL, wt & 4u);
      const uint32_t prev = __popc(v0 & glt) + 2u * __popc(v1 & glt);
      const uint32_t pret = __popc(t0 & glt) + 2u * __popc(t1 & glt) + 4u * __popc(t2 & glt);
      const int leader = __ffs(grp) - 1;
      uint32_t old = 0;
      if (work && lane == leader) old = atomicAdd(&S.lcnt[hs], tot);
      old = __shfl_sync(FULL, old, leader);
#endif
      if (work && nv) {
        uint32_t r = (old & 0xFFFFu) + prev;
        uint32_t mm = mine;
        while (mm) {
          const int s6 = __ffs(mm) - 1;
          mm &= mm - 1u;
          const uint32_t lg = rowpre + S.pp8[row][s6] + __popc(S.pl[row][s6] & ltf);
          if (MODE == 0) {
            S.vstage[lg] = r | ((uint32_t)hs << 11) | (vidx << 18) | ((uint32_t)s6 << 29);
          } else {
            S.vstage[lg] = (r << 12) | (uint32_t)hs;
            S.pstage[lg] = (uint16_t)(vidx | ((uint32_t)s6 << 11));
          }
          ++r;
        }
      }
      const bool hasrec = work && nt != 0u;
      const uint32_t rb = __ballot_sync(FULL, hasrec);
      if (rb) {
        uint32_t rbase = 0;
        if (lane == 0) rbase = atomicAdd(&S.nrec, (uint32_t)__popc(rb));
        rbase = __shfl_sync(FULL, rbase, 0);
        if (hasrec) {
          const uint32_t pos = rbase + __popc(rb & ltm);
          if (pos < (uint32_t)RCAP) {
            S.rstage[pos] = vidx | (cs << 11) | ((uint32_t)hs << 19);
            S.rtoff[pos] = (uint16_t)((old >> 16) + pret);
          } else {
            S.overflow = 1u;
          }
        }
      }
    }
  }
  __syncthreads();  // the staged labels are dead from here on (lvb / cidx reuse their memory)
  if (S.overflow) {
    if (MODE == 0) {
      if (tid == 0) o.dense_list[atomicAdd(&o.ctl->dense_count, 1u)] = tile;
    } else {
      if (tid == 0) atomicOr(&o.ctl->flags, FLAG_INTERNAL);
    }
    return TILE_DEFERRED;
  }

  // ---- S5: reservations.  Thread 0's global atomics overlap the per-label ones of the others ----
  const uint32_t nrec = S.nrec;
  if (tid == 0) {
    const u64 gbase = nslots ? atomicAdd(&o.ctl->cur_perm, (u64)nslots) : 0ull;
    const u64 recbase = nrec ? atomicAdd(&o.ctl->cur_rec, (u64)nrec) : 0ull;
    const bool ok = gbase + nslots <= o.capV && recbase + nrec <= o.capR;
    if (!ok) atomicOr(&o.ctl->flags, FLAG_CAP);
    S.gbase = (uint32_t)gbase;
    S.recbase = recbase;
    S.ok1 = ok ? 1u : 0u;
  }
  for (int i = tid; i < LT; i += NT) {
    const u64 label = S.lkeys[i];
    if (label != 0ull) {
      const uint32_t cnt = S.lcnt[i];
      const uint32_t nv = cnt & 0xFFFFu, nt = cnt >> 16;
      const int gs = gtab_insert(o.ht, label, &o.ctl->flags);
      u64 old = 0;
      if (gs >= 0) old = atomicAdd(&o.ht.cnt[gs], (u64)nv | ((u64)nt << 32));
      S.lvb[i] = (uint32_t)old;
      S.cidx[i] = (uint16_t)atomicAdd(&S.ci, 1u);
      S.tla[i] = (u64)(uint32_t)(gs >= 0 ? gs : 0) | (old & 0xFFFFFFFF00000000ull);
      if (nt) atomicAdd(&S.ttot, nt);
    }
  }
  __syncthreads();
  if (!S.ok1) {  // capacity guess too small: ALL cursors still advance, so one failed attempt gives the exact need of every buffer
    if (tid == 0 && S.ci) atomicAdd(&o.ctl->cur_tl, (u64)S.ci);
    return TILE_DONE;
  }
  const uint32_t gbase = S.gbase;
  const u64 recbase = S.recbase;
  if (tid == 0) {  // tl block + work-list entry; the latency of these atomics hides behind the flush below
    const uint32_t nlab = S.ci;
    const u64 tlbase = nlab ? atomicAdd(&o.ctl->cur_tl, (u64)nlab) : 0ull;
    const bool ok = tlbase + nlab <= o.capL;
    if (!ok) atomicOr(&o.ctl->flags, FLAG_CAP);
    else {
      TileHdr h;
      h.recbase = recbase;
      h.gbase = gbase;
      h.tlbase = (uint32_t)tlbase;
      h.nslots = (uint16_t)nslots;
      h.nrec = (uint16_t)nrec;
      h.nlab = (uint16_t)nlab;
      h.pad = 0;
      h.tile = tile;
      h.pad2 = 0;
      o.hdr[atomicAdd(&o.ctl->work_count, 1u)] = h;
      if (S.ttot) atomicAdd(&o.ctl->cur_tri, (u64)S.ttot);
    }
    S.tlbase = (uint32_t)tlbase;
    S.ok2 = ok ? 1u : 0u;
  }

  // ---- S6: flush rowinfo, perm/vinfo and the records (coalesced) ----
  {
    const int row = tid >> 2, q = tid & 3;  // 4 threads per row segment, 8 bytes each
    if (row >= h0 * TM && row < (h0 + nh) * TM) {
      const uint32_t rs = ts * TS + row / TM, rm = tm * TM + row % TM;
      if (rs < vp.Es_own && rm < vp.Em) {
        uint2 w = *reinterpret_cast<const uint2*>(&S.pl[row][2 * q]);
        if (q == 3) { w.x += gbase; w.y = 0u; }
        reinterpret_cast<uint2*>(o.rowinfo + (((size_t)rs * vp.Em + rm) * vp.ntf + tf) * RI_WORDS)[q] = w;
      }
    }
  }
  for (uint32_t i = tid; i < nslots; i += NT) {
    const uint32_t w = S.vstage[i];
    if (MODE == 0) {
      const uint32_t hs = (w >> 11) & 0x7Fu;
      o.perm[(size_t)gbase + i] = S.lvb[hs] + (w & 0x7FFu);
      o.vinfo[(size_t)gbase + i] = (w >> 18) | ((uint32_t)S.cidx[hs] << 14);
    } else {
      const uint32_t hs = w & 0xFFFu;
      o.perm[(size_t)gbase + i] = S.lvb[hs] + (w >> 12);
      o.vinfo[(size_t)gbase + i] = (uint32_t)S.pstage[i] | ((uint32_t)S.cidx[hs] << 14);
    }
  }
  for (uint32_t i = tid; i < nrec; i += NT) {
    const uint32_t w = S.rstage[i];
    o.rec[recbase + i] = (u64)((w & 0x7FFFFu) | ((uint32_t)S.cidx[w >> 19] << 19)) | ((u64)S.rtoff[i] << 32);
  }
  __syncthreads();
  if (!S.ok2) return TILE_DONE;
  const uint32_t tlbase = S.tlbase;
  for (int i = tid; i < LT; i += NT) {
    if (S.lkeys[i] != 0ull) {
      TLEntry e;
      e.a = S.tla[i];
      e.b = 0;
      o.tl[tlbase + S.cidx[i]] = e;
    }
  }
  return TILE_DONE;
}

// ---------------------------------------------------------------------------------------------
// The two-label path of MODE 0 (ZM_K2_PATH): the staged region holds exactly two labels A and B (one of them may
// be the background).  82 % of c5's non-empty tiles and a third of connectomics' are of this kind.
//   * S1: one 33-bit mask per staged row (k2_rows); edges, non-zero masks, slot planes and per-row counts of A's
//     slots are bit arithmetic on whole rows; B's corner mask of a cube is the complement of A's.
//   * every count the global reservations need is known after the row scan (slots, A's slots, the non-uniform
//     cubes = records per label), so the cursor and label-table atomics are ISSUED before the active voxels are
//     compacted and their results are picked up after S3: their L2 round trips (7 % of the kernel's warp time on
//     the critical path of the general formulation) overlap the compaction and S3.
//   * S3 splits the (chunk of 32 active voxels, label) pairs over the warps: all lanes of a warp work on the same
//     label, whose counter is the one address the shared-memory atomic ranks them on.
template <typename L, bool CO>
__device__ __forceinline__ int tile_body_k2(const VolParams& vp, const Pass1Args& o, P1Smem<L, 0>& S, const uint32_t tile,
                                            const uint32_t tf, const uint32_t tm, const uint32_t ts, const L keyA) {
  constexpr int VCAP = TileCap<L, 0>::V, RCAP = TileCap<L, 0>::R;
  constexpr uint32_t FULL = 0xffffffffu;
  constexpr int ALIGN = 16 / (int)sizeof(L);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t ltm = (1u << lane) - 1u;
  const uint32_t ef0 = tf * TF, em0 = tm * TM, es0 = ts * TS;
  const L* const lab = S.lab + (vp.pad ? ALIGN - 1 : 0);

  if (tid < 2) S.lcnt[tid] = 0u;
  if (tid == 0) S.nrec = 0;
  if (__syncthreads_or(k2_rows<L, 0>(S, lab, keyA) ? 1 : 0)) return TILE_NOT_K2;
  L keyB = keyA;
  {
    const int n2 = k2_second_label<L, 0>(S, keyA, keyB);
    if (n2 == 0) return TILE_EMPTY;
    if (n2 == 2) return TILE_NOT_K2;
  }
  const bool zA = keyA == 0, zB = keyB == 0;  // (at most one of the two is the background)

  // ---- phase B + S2: thread t < 64 owns row t (plane pw, row pj) ----
  bool any = false;
  uint32_t packed = 0;   // slots of the row | active voxels of the row << 16
  uint32_t packed2 = 0;  // slots of the row that belong to A | non-uniform (valid) cubes of the row << 16
  const int pw = tid >> 3, pj = tid & 7;
  if (warp < NROWS / 32) {
    const int row = tid;
    const u64* mk = S.lkeys + K2_MASK0 + (pw * RM + pj);
    const u64 m00 = mk[0], m10 = mk[1], m01 = mk[RM], m11 = mk[RM + 1];
    const uint32_t a00 = (uint32_t)m00, a00f = (uint32_t)(m00 >> 1), a10 = (uint32_t)m10, a01 = (uint32_t)m01;
    const uint32_t ef = a00 ^ a00f, em = a00 ^ a10, es = a00 ^ a01;
    const uint32_t nonuni = ef | em | es | (uint32_t)(m10 ^ (m10 >> 1)) | (uint32_t)(m01 ^ (m01 >> 1)) |
                            (uint32_t)(m11 ^ (m11 >> 1)) | (uint32_t)(m01 ^ m11);
    uint32_t pf = ef, pm = em, ps = es, cube = nonuni;
    if (!(ef0 + TF + 1 <= vp.Ef && em0 + TM + 1 <= vp.Em && es0 + TS + 1 <= vp.Es)) {
      // volume boundary within reach (see tile_body): slots need their upper voxel inside the extended volume and
      // a lower voxel this shard owns; a cube needs all three upper neighbours
      const uint32_t nfv = vp.Ef - ef0;
      const uint32_t VF = nfv >= 32u ? FULL : (1u << nfv) - 1u;
      const uint32_t NF1 = nfv >= 33u ? FULL : (1u << (nfv - 1u)) - 1u;
      const uint32_t em_ = em0 + (uint32_t)pj, es_ = es0 + (uint32_t)pw;
      const bool rowok = em_ < vp.Em && es_ < vp.Es_own;
      const bool nm1 = em_ + 1 < vp.Em, ns1 = es_ + 1 < vp.Es;
      pf = rowok ? pf & NF1 : 0u;
      pm = rowok && nm1 ? pm & VF : 0u;
      ps = rowok && ns1 ? ps & VF : 0u;
      cube = rowok && nm1 && ns1 ? nonuni & NF1 : 0u;
    }
    // slot 2d: the edge's lower voxel carries a label (A where its mask bit is set, else B); 2d + 1: the upper one
    const uint32_t nzlo = zA ? ~a00 : (zB ? a00 : FULL);  // lower voxel is not the background
    const uint32_t b0 = pf & nzlo, b2 = pm & nzlo, b4 = ps & nzlo;
    const uint32_t b1 = pf & (zA ? ~a00f : (zB ? a00f : FULL));
    const uint32_t b3 = pm & (zA ? ~a10 : (zB ? a10 : FULL));
    const uint32_t b5 = ps & (zA ? ~a01 : (zB ? a01 : FULL));
    const uint32_t act = b0 | b1 | b2 | b3 | b4 | b5 | cube;
    *reinterpret_cast<uint4*>(&S.pl[row][0]) = make_uint4(b0, b1, b2, b3);
    *reinterpret_cast<uint4*>(&S.pl[row][4]) = make_uint4(b4, b5, 0u, act);
    any = act != 0u;
    const uint32_t c0 = __popc(b0), c1 = c0 + __popc(b1), c2 = c1 + __popc(b2), c3 = c2 + __popc(b3);
    const uint32_t c4 = c3 + __popc(b4);
    *reinterpret_cast<uint2*>(S.pp8[row]) = make_uint2((c0 << 8) | (c1 << 16) | (c2 << 24), c3 | (c4 << 8));
    packed = (c4 + __popc(b5)) | ((uint32_t)__popc(act) << 16);
    const uint32_t nva = __popc(b0 & a00) + __popc(b2 & a00) + __popc(b4 & a00) + __popc(b1 & a00f) + __popc(b3 & a10) +
                         __popc(b5 & a01);
    packed2 = nva | ((uint32_t)__popc(cube) << 16);
    uint32_t rinc = packed;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t t = __shfl_up_sync(FULL, rinc, d);
      if (lane >= d) rinc += t;
    }
    const uint32_t ex = rinc - packed;
    S.pl[tid][6] = ex & 0xFFFFu;
    S.actpre[tid] = (uint16_t)(ex >> 16);
    const uint32_t tot2 = __reduce_add_sync(FULL, packed2);
    if (lane == 31) { S.wtot[warp] = rinc; S.wtot[2 + warp] = tot2; }
  }
  if (!__syncthreads_or(any ? 1 : 0)) return TILE_EMPTY;
  const uint32_t g0 = S.wtot[0];
  const uint32_t nslots = (g0 + S.wtot[1]) & 0xFFFFu, nact = (g0 + S.wtot[1]) >> 16;
  const uint32_t t2 = S.wtot[2] + S.wtot[3];
  const uint32_t nvA = t2 & 0xFFFFu, ncubes = t2 >> 16;
  const uint32_t nnz = (zA ? 0u : 1u) + (zB ? 0u : 1u);
  const uint32_t nrec = ncubes * nnz;  // every non-uniform cube holds both labels
  if (nslots > (uint32_t)VCAP || nrec > (uint32_t)RCAP) {
    if (tid == 0) o.dense_list[atomicAdd(&o.ctl->dense_count, 1u)] = tile;
    return TILE_DEFERRED;
  }
  if (warp == 1) S.pl[tid][6] += g0 & 0xFFFFu;
  const uint32_t actbase = warp >= NW / 2 ? g0 >> 16 : 0u;
  const bool hasA = !zA && (nvA | ncubes) != 0u, hasB = !zB && ((nslots - nvA) | ncubes) != 0u;
  const uint32_t nlab = (hasA ? 1u : 0u) + (hasB ? 1u : 0u);

  // ---- early reservations (results are read after S3) ----
  u64 r_gbase = 0, r_recbase = 0, r_oldv = 0;
  int r_gs = 0;
  if (tid == 0) {
    if (nslots) r_gbase = atomicAdd(&o.ctl->cur_perm, (u64)nslots);
    if (nrec) r_recbase = atomicAdd(&o.ctl->cur_rec, (u64)nrec);
  }
  const bool lab_lane = warp == 1 && lane < 2 && (lane == 0 ? hasA : hasB);  // lane 0: A, lane 1: B
  if (lab_lane) {
    const int gs = gtab_insert(o.ht, (u64)(lane == 0 ? keyA : keyB), &o.ctl->flags);
    r_gs = gs >= 0 ? gs : 0;
    if (gs >= 0) r_oldv = atomicAdd(&o.ht.cnt[gs], (u64)(lane == 0 ? nvA : nslots - nvA));
  }

  // ---- compaction of the active voxels of the warp's own plane ----
#pragma unroll
  for (int j = 0; j < TM; ++j) {
    const int row = warp * TM + j;
    const uint32_t mask = S.pl[row][7];
    if ((mask >> lane) & 1u) S.alist[actbase + S.actpre[row] + __popc(mask & ltm)] = (uint16_t)(row * TF + lane);
  }
  __syncthreads();

  // ---- S3: (chunk of 32 active voxels, label) tasks over the warps ----
  const uint32_t ntask = ((nact + 31u) >> 5) * nnz;
  for (uint32_t t = warp; t < ntask; t += NW) {
    const uint32_t k = nnz == 2u ? (t & 1u) : (zA ? 1u : 0u);
    const uint32_t i = (nnz == 2u ? (t >> 1) : t) * 32u + lane;
    const bool valid = i < nact;
    const uint32_t vidx = valid ? S.alist[i] : 0u;
    const int lf = vidx & 31, lm = (vidx >> 5) & 7, ls = vidx >> 8;
    const bool axf = ef0 + lf + 1 < vp.Ef, axm = em0 + lm + 1 < vp.Em, axs = es0 + ls + 1 < vp.Es;
    const uint32_t amask = (axf ? 0x03u : 0u) | (axm ? 0x0Cu : 0u) | (axs ? 0x30u : 0u);
    const bool cube = axf && axm && axs;
    const uint32_t row = vidx >> 5;
    const uint32_t ltf = (1u << lf) - 1u;
    const uint32_t rowpre = S.pl[row][6];
    const u64* mk = S.lkeys + K2_MASK0 + (ls * RM + lm);
    uint32_t b[4];  // index dm + 2 * ds: bits (f, f + 1) of the row
    b[0] = (uint32_t)(mk[0] >> lf) & 3u; b[1] = (uint32_t)(mk[1] >> lf) & 3u;
    b[2] = (uint32_t)(mk[RM] >> lf) & 3u; b[3] = (uint32_t)(mk[RM + 1] >> lf) & 3u;
    uint32_t msk = 0;
#pragma unroll
    for (int n = 0; n < 8; ++n) msk |= ((b[corner_dm<CO>(n) + 2 * corner_ds<CO>(n)] >> corner_df<CO>(n)) & 1u) << n;
    if (k) msk = ~msk & 0xFFu;
    const uint32_t cs = ~msk & 0xFFu;
    uint32_t nt = 0, mine = 0;
    if (valid) {
      if (cube) nt = __ldg(&TRI_COUNT_D[cs]);
      const uint32_t nb = ((msk >> corner_plus_f<CO>()) & 1u) | (((msk >> corner_plus_m<CO>()) & 1u) << 2) |
                          (((msk >> corner_plus_s<CO>()) & 1u) << 4);
      mine = ((msk & 1u) ? (0x15u & ~nb) : (nb << 1)) & amask;
    }
    const uint32_t nv = __popc(mine);
    uint32_t old = 0;
    if ((nv | nt) != 0u) old = atomicAdd(&S.lcnt[k], nv | (nt << 16));
    if (nv) {
      uint32_t r = old & 0xFFFFu;
      uint32_t mm = mine;
      while (mm) {
        const int s6 = __ffs(mm) - 1;
        mm &= mm - 1u;
        const uint32_t lg = rowpre + S.pp8[row][s6] + __popc(S.pl[row][s6] & ltf);
        S.vstage[lg] = r | (k << 11) | (vidx << 18) | ((uint32_t)s6 << 29);
        ++r;
      }
    }
    const bool hasrec = nt != 0u;
    const uint32_t rb = __ballot_sync(FULL, hasrec);
    if (rb) {
      uint32_t rbase = 0;
      if (lane == 0) rbase = atomicAdd(&S.nrec, (uint32_t)__popc(rb));
      rbase = __shfl_sync(FULL, rbase, 0);
      if (hasrec) {
        const uint32_t pos = rbase + __popc(rb & ltm);  // (< nrec <= RCAP)
        S.rstage[pos] = vidx | (cs << 11) | (k << 19);
        S.rtoff[pos] = (uint16_t)(old >> 16);
      }
    }
  }
  __syncthreads();  // (the staged labels have been dead since k2_rows: lvb / cidx reuse their memory)

  // ---- publish the early reservations, issue the remaining atomics ----
  const uint32_t ntA = S.lcnt[0] >> 16, ntB = S.lcnt[1] >> 16;
  u64 r_tl = 0, r_oldt = 0;
  uint32_t r_wc = 0;
  if (tid == 0) {
    const bool ok = r_gbase + nslots <= o.capV && r_recbase + nrec <= o.capR;
    if (!ok) atomicOr(&o.ctl->flags, FLAG_CAP);
    if (S.nrec != nrec) atomicOr(&o.ctl->flags, FLAG_INTERNAL);
    S.gbase = (uint32_t)r_gbase;
    S.recbase = r_recbase;
    S.ok1 = ok ? 1u : 0u;
    if (nlab) r_tl = atomicAdd(&o.ctl->cur_tl, (u64)nlab);
    if (ok) r_wc = atomicAdd(&o.ctl->work_count, 1u);
    if (ntA + ntB) atomicAdd(&o.ctl->cur_tri, (u64)(ntA + ntB));
  }
  if (lab_lane) {
    S.lvb[lane] = (uint32_t)r_oldv;
    r_oldt = atomicAdd(&o.ht.cnt[r_gs], (u64)(lane == 0 ? ntA : ntB) << 32);
  }
  const uint32_t ciB = hasA ? 1u : 0u;  // tile-local label indices: A -> 0, B -> 1 (0 when A is not meshed here)
  __syncthreads();
  if (!S.ok1) return TILE_DONE;  // capacity guess too small: the cursors give the exact need; host reruns
  const uint32_t gbase = S.gbase;
  const u64 recbase = S.recbase;

  // ---- S6: flush rowinfo, perm / vinfo and the records (coalesced) ----
  {
    const int row = tid >> 2, q = tid & 3;  // 4 threads per row segment, 8 bytes each
    const uint32_t rs = ts * TS + row / TM, rm = tm * TM + row % TM;
    if (rs < vp.Es_own && rm < vp.Em) {
      uint2 w = *reinterpret_cast<const uint2*>(&S.pl[row][2 * q]);
      if (q == 3) { w.x += gbase; w.y = 0u; }
      reinterpret_cast<uint2*>(o.rowinfo + (((size_t)rs * vp.Em + rm) * vp.ntf + tf) * RI_WORDS)[q] = w;
    }
  }
  const uint32_t lvbA = S.lvb[0], lvbB = S.lvb[1];
  for (uint32_t i = tid; i < nslots; i += NT) {
    const uint32_t w = S.vstage[i];
    const bool isB = (w >> 11) & 1u;
    o.perm[(size_t)gbase + i] = (isB ? lvbB : lvbA) + (w & 0x7FFu);
    o.vinfo[(size_t)gbase + i] = (w >> 18) | ((isB ? ciB : 0u) << 14);
  }
  for (uint32_t i = tid; i < nrec; i += NT) {
    const uint32_t w = S.rstage[i];
    o.rec[recbase + i] = (u64)((w & 0x7FFFFu) | (((w >> 19) ? ciB : 0u) << 19)) | ((u64)S.rtoff[i] << 32);
  }
  if (tid == 0) {
    const bool ok = r_tl + nlab <= o.capL;
    if (!ok) atomicOr(&o.ctl->flags, FLAG_CAP);
    else {
      TileHdr h;
      h.recbase = recbase;
      h.gbase = gbase;
      h.tlbase = (uint32_t)r_tl;
      h.nslots = (uint16_t)nslots;
      h.nrec = (uint16_t)nrec;
      h.nlab = (uint16_t)nlab;
      h.pad = 0;
      h.tile = tile;
      h.pad2 = 0;
      o.hdr[r_wc] = h;
    }
    S.tlbase = (uint32_t)r_tl;
    S.ok2 = ok ? 1u : 0u;
  }
  __syncthreads();
  if (!S.ok2) return TILE_DONE;
  if (lab_lane) {
    TLEntry e;
    e.a = (u64)(uint32_t)r_gs | (r_oldt & 0xFFFFFFFF00000000ull);
    e.b = 0;
    o.tl[S.tlbase + (lane == 0 ? 0u : ciB)] = e;
  }
  return TILE_DONE;
}

// MODE 0: one CTA per tile (the hardware CTA scheduler overlaps the TMA wait of one tile with the
// work of the others resident on the SM).  MODE 1: the queued dense tiles, two half tiles each.
#ifndef ZM_PREFETCH_LAYERS
#define ZM_PREFETCH_LAYERS 1
#endif
// s-layers ahead whose region is pulled into L2 (0: off).  c5 k_classify ms at 1 / 2 / 3 / 4 layers (512 CTAs each):
// 27.14 / 27.44 / 28.19 / 28.70; all-zero volume 12.50 / 12.69 / 12.76 / 13.85
constexpr uint32_t PREFETCH_LAYERS = ZM_PREFETCH_LAYERS;
template <typename L, bool CO, int MODE>
__global__ void __launch_bounds__(NT, MODE == 0 ? (sizeof(L) == 8 ? ZM_U64_CTAS : ZM_U32_CTAS) : 1)
k_classify(const VolParams vp, const __grid_constant__ CUtensorMap tmap, const Pass1Args o) {
  P1Smem<L, MODE>& S = *reinterpret_cast<P1Smem<L, MODE>*>(zm_dyn_smem);
  const int tid = threadIdx.x;
  if (tid == 0) {
    mbar_init(&S.mbar, 1);
    fence_mbar_init();
  }
  if constexpr (MODE == 0) {
    // 3-d grid in the launch order described at TM_GROUP: x = tf + ntf * (row inside the group), y = ts, z = group.
    // Every thread derives the coordinates itself (one division by ntf), so nothing but the mbarrier initialisation
    // precedes the first CTA barrier (the serial prologue of thread 0 used to cost 10 % of the kernel's warp time).
    const uint32_t tmg = blockIdx.x / vp.ntf;
    const uint32_t tf = blockIdx.x - tmg * vp.ntf, tm = blockIdx.z * TM_GROUP + tmg, ts = blockIdx.y;
    if (tm >= vp.ntm) return;  // (the last group may be short)
    if (tid == 0) begin_tile(vp, &tmap, S, tf, tm, ts);
    __syncthreads();  // mbarrier initialised
    if (tid == 0 && PREFETCH_LAYERS != 0 && vp.use_tma) {
      // pull the region of the tile PREFETCH_LAYERS s-layers up (launched ntf * TM_GROUP * PREFETCH_LAYERS CTAs later) into L2
      const uint32_t pts = ts + PREFETCH_LAYERS;
      if (pts < vp.nts) {
        int c0, c1, c2;
        stage_origin<L>(vp, tf, tm, pts, c0, c1, c2);
        tma_prefetch_3d(&tmap, c0, c1, c2);
      }
    }
    const uint32_t tile = tf + vp.ntf * (tm + vp.ntm * ts);  // canonical index (work list, dense list)
    if (vp.use_tma) {
      mbar_wait(&S.mbar, 0);  // every thread waits itself: the TMA writes are visible to it afterwards
    } else {
      stage_plain(vp, S, tf, tm, ts);
      __syncthreads();
    }
    int status = TILE_EMPTY;
    L first;
    if (!region_uniform(S, S.lab + (vp.pad ? 16 / (int)sizeof(L) - 1 : 0), first)) {
#if ZM_K2_PATH
      status = tile_body_k2<L, CO>(vp, o, S, tile, tf, tm, ts, first);
      if (status == TILE_NOT_K2)
#endif
        status = tile_body<L, CO, MODE>(vp, o, S, tile, tf, tm, ts, 0, TS);
    }
    if (status == TILE_EMPTY) zero_rows(vp, o, tf, tm, ts, 0, TS);
  } else {
    const uint32_t n = o.ctl->dense_count;  // written by the MODE 0 launch that precedes this one
    uint32_t parity = 0;
    for (uint32_t i = blockIdx.x; i < n; i += gridDim.x) {
      const uint32_t tile = o.dense_list[i];
      for (int half = 0; half < Caps<MODE>::HALVES; ++half) {
        const int nh = TS / Caps<MODE>::HALVES, h0 = half * nh;
        if (tid == 0) {  // (the region is staged again: lvb / cidx of the first half overwrote it)
          uint32_t b = tile;
          const uint32_t tf0 = b % vp.ntf;
          b /= vp.ntf;
          begin_tile(vp, &tmap, S, tf0, b % vp.ntm, b / vp.ntm);
        }
        __syncthreads();
        const uint32_t tf = S.tc[0], tm = S.tc[1], ts = S.tc[2];
        if (vp.use_tma) {
          mbar_wait(&S.mbar, parity);
          parity ^= 1u;
        } else {
          stage_plain(vp, S, tf, tm, ts);
          __syncthreads();
        }
        const int status = tile_body<L, CO, MODE>(vp, o, S, tile, tf, tm, ts, h0, nh);
        if (status == TILE_EMPTY) zero_rows(vp, o, tf, tm, ts, h0, nh);
        __syncthreads();
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// label table scan: per-slot exclusive offsets + compact list of live labels

struct ScanArgs {
  LabelTable ht;
  u64* offV;     // [cap]
  u64* offT;     // [cap]
  u64* list;     // [cap][3]: label, nV, nT  (compact, table order)
  u64* partial;  // [3][1024]
  Control* ctl;
  uint32_t chunk;  // table slots per CTA (multiple of 1024)
};

struct Scan3 { u64 v, t; uint32_t n; };

__device__ __forceinline__ Scan3 block_scan3(Scan3 x, Scan3& total, Scan3* sh /*[32]*/) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  Scan3 inc = x;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const u64 tv = __shfl_up_sync(0xffffffffu, inc.v, d), tt = __shfl_up_sync(0xffffffffu, inc.t, d);
    const uint32_t tn = __shfl_up_sync(0xffffffffu, inc.n, d);
    if (lane >= d) { inc.v += tv; inc.t += tt; inc.n += tn; }
  }
  if (lane == 31) sh[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    Scan3 w = sh[lane];
    Scan3 wi = w;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const u64 tv = __shfl_up_sync(0xffffffffu, wi.v, d), tt = __shfl_up_sync(0xffffffffu, wi.t, d);
      const uint32_t tn = __shfl_up_sync(0xffffffffu, wi.n, d);
      if (lane >= d) { wi.v += tv; wi.t += tt; wi.n += tn; }
    }
    Scan3 ex;
    ex.v = wi.v - w.v; ex.t = wi.t - w.t; ex.n = wi.n - w.n;
    sh[lane] = ex;
    if (lane == 31) sh[32] = wi;
  }
  __syncthreads();
  const Scan3 wb = sh[warp];
  total = sh[32];
  Scan3 r;
  r.v = wb.v + inc.v - x.v; r.t = wb.t + inc.t - x.t; r.n = wb.n + inc.n - x.n;
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(1024) k_scan_partials(const ScanArgs a) {
  __shared__ Scan3 sh[33];
  const uint32_t lo = blockIdx.x * a.chunk;
  Scan3 x;
  x.v = 0; x.t = 0; x.n = 0;
  for (uint32_t i = lo + threadIdx.x; i < lo + a.chunk; i += 1024) {
    if (a.ht.keys[i] != 0ull) {
      const u64 c = a.ht.cnt[i];
      x.v += c & 0xFFFFFFFFull;
      x.t += c >> 32;
      x.n += c != 0ull ? 1u : 0u;
    }
  }
  Scan3 total;
  (void)block_scan3(x, total, sh);
  if (threadIdx.x == 0) {
    a.partial[blockIdx.x] = total.v;
    a.partial[1024 + blockIdx.x] = total.t;
    a.partial[2048 + blockIdx.x] = total.n;
  }
}

__global__ void __launch_bounds__(1024) k_scan_apply(const ScanArgs a) {
  __shared__ Scan3 sh[33];
  __shared__ Scan3 s_base;
  Scan3 p;
  p.v = 0; p.t = 0; p.n = 0;
  if (threadIdx.x < gridDim.x) {
    p.v = a.partial[threadIdx.x];
    p.t = a.partial[1024 + threadIdx.x];
    p.n = (uint32_t)a.partial[2048 + threadIdx.x];
  }
  Scan3 total;
  const Scan3 ex = block_scan3(p, total, sh);
  if (threadIdx.x == blockIdx.x) s_base = ex;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    a.ctl->totals[0] = total.n;
    a.ctl->totals[1] = total.v;
    a.ctl->totals[2] = total.t;
    a.ctl->totals[3] = 0;
  }
  __syncthreads();
  Scan3 base = s_base;
  const uint32_t lo = blockIdx.x * a.chunk;
  for (uint32_t off = 0; off < a.chunk; off += 1024) {
    const uint32_t i = lo + off + threadIdx.x;
    const u64 key = a.ht.keys[i];
    const u64 c = key != 0ull ? a.ht.cnt[i] : 0ull;
    Scan3 x;
    x.v = c & 0xFFFFFFFFull; x.t = c >> 32; x.n = c != 0ull ? 1u : 0u;
    Scan3 tot;
    const Scan3 e = block_scan3(x, tot, sh);
    a.offV[i] = base.v + e.v;
    a.offT[i] = base.t + e.t;
    if (x.n) {
      const u64 li = (u64)base.n + e.n;
      a.list[3 * li + 0] = key;
      a.list[3 * li + 1] = x.v;
      a.list[3 * li + 2] = x.t;
    }
    base.v += tot.v; base.t += tot.t; base.n += tot.n;
  }
}

// tl[i]: (label slot | face base in label << 32)  ->  (first vertex row of the label | shard index offset << 32,
// first face row of the (tile,label) block)
__global__ void __launch_bounds__(256) k_tl_fixup(TLEntry* tl, const Control* ctl, u64 capL, const u64* offV,
                                                  const u64* offT, const uint32_t* voff) {
  u64 n = ctl->cur_tl;
  if (n > capL) n = capL;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
    const u64 a = tl[i].a;
    const uint32_t gs = (uint32_t)a;
    TLEntry e;
    e.a = offV[gs] | ((u64)(voff ? voff[gs] : 0u) << 32);
    e.b = offT[gs] + (a >> 32);
    tl[i] = e;
  }
}

// voff[slot(labels[i])] = offs[i]  (labels absent from the table are ignored)
__global__ void __launch_bounds__(256) k_set_voff(const LabelTable ht, const u64* labels, const uint32_t* offs, u64 n,
                                                  uint32_t* voff) {
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
    const u64 label = labels[i];
    if (label == 0ull) continue;
    uint32_t h = hash_label(label) & ht.mask;
    for (uint32_t p = 0; p <= ht.mask && p < 4096u; ++p) {
      const u64 k = ht.keys[h];
      if (k == label) { voff[h] = offs[i]; break; }
      if (k == 0ull) break;
      h = (h + 1u) & ht.mask;
    }
  }
}

// slab sharding, device-side directory exchange: every shard publishes its (label, vertices) list in a fixed-capacity
// buffer -- word 0 = number of labels, then (label, n_vertices) pairs -- which is all-gathered over NCCL on the
// mesher's stream; each shard then sums, per label, the vertices on EARLIER shards into voff[] (indexed by its own
// label-table slot).  No host round trip (the host-side path of zmesh_b200/sharded.py cost 0.8 ms per step at 8 GPUs).
__global__ void __launch_bounds__(256) k_export_directory(const u64* list, const Control* ctl, u64* dst, u64 capacity) {
  const u64 n = ctl->totals[0];
  if (blockIdx.x == 0 && threadIdx.x == 0) { dst[0] = n; dst[1] = 0ull; }
  const u64 m = n < capacity ? n : capacity;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (u64)gridDim.x * blockDim.x) {
    dst[2 + 2 * i] = list[3 * i];
    dst[3 + 2 * i] = list[3 * i + 1];
  }
}

// all: [world][1 + capacity][2] u64.  grid.y = shard q (every shard checks EVERY directory for overflow, so that all
// shards take the same decision; only the earlier ones, q < rank, contribute offsets).
__global__ void __launch_bounds__(256) k_import_directories(const LabelTable ht, const u64* all, uint32_t rank, u64 capacity,
                                                            uint32_t* voff, Control* ctl) {
  const uint32_t q = blockIdx.y;
  const u64* src = all + (size_t)q * 2 * (1 + capacity);
  const u64 n = src[0];
  if (n > capacity) {
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(&ctl->flags, FLAG_DIR);
    return;
  }
  if (q >= rank) return;
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
    const u64 label = src[2 + 2 * i];
    const uint32_t nv = (uint32_t)src[3 + 2 * i];
    if (label == 0ull || nv == 0u) continue;
    uint32_t h = hash_label(label) & ht.mask;
    for (uint32_t p = 0; p <= ht.mask && p < 4096u; ++p) {
      const u64 k = ht.keys[h];
      if (k == label) { atomicAdd(&voff[h], nv); break; }
      if (k == 0ull) break;  // (the label does not occur on this shard)
      h = (h + 1u) & ht.mask;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// pass 2

struct Pass2Args {
  const TileHdr* hdr;
  const uint32_t* perm;
  const uint32_t* vinfo;
  const u64* rec;
  const TLEntry* tl;
  uint32_t* faces;  // [T_total][3]
  float* verts;     // [V_total][3]
  float* normals;   // [V_total][3] or null (k_normals_normalize4 writes it; arbitrary-mesh path accumulates in it)
  float4* nacc;     // [V_total] accumulation rows of pass 2: ONE 16-byte vector atomic per triangle corner instead of three
                    // scalar ones (sm_90+: atomicAdd(float4*)); .w unused
  float r0, r1, r2;  // captured resolution
  float c0, c1, c2;  // centering offset
  uint32_t n_work;
  int voxel_centered, transpose;
  int write_faces, write_verts;
  // slab sharding: 0 = all tiles; 1 = all but the top tile layer (tiles >= top_tile_lo: the only ones whose cubes
  // reference the boundary plane of the next shard), launched while that plane is still in flight; 2 = only the top layer
  uint32_t layer_mode, top_tile_lo;
  const uint32_t* foreign;  // slab sharding: final indices of the top plane's slots, [Em][Efp][4], from the next shard
  float* fnormals;          // slab sharding + normals: contributions to the next shard's first-plane vertices, [Em][Efp][4][3]
};

__device__ __forceinline__ float len3(float x, float y, float z) {
  return __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
}

// One face: n_hat = hat(cross(v1-v0, v2-v0)); N[f_k] += n_hat * |v_k - centroid|
// (chunk_mesh.hpp:355-368; float32 op for op, accumulation order is not the reference's).  A destination is a row of
// three floats (scalar atomics) or, vec[k] = true, a 16-byte aligned float4 row (one vector atomic).
__device__ __forceinline__ void face_normal_scatter(const float v0[3], const float v1[3], const float v2[3],
                                                    float* d0, float* d1, float* d2, bool vec0 = false, bool vec1 = false,
                                                    bool vec2 = false) {
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
  const bool vec[3] = {vec0, vec1, vec2};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float w = len3(__fsub_rn(vv[k][0], c[0]), __fsub_rn(vv[k][1], c[1]), __fsub_rn(vv[k][2], c[2]));
    const float x = __fmul_rn(n0, w), y = __fmul_rn(n1, w), z = __fmul_rn(n2, w);
    if (vec[k]) {
      atomicAdd(reinterpret_cast<float4*>(dd[k]), make_float4(x, y, z, 0.0f));
    } else {
      atomicAdd(dd[k] + 0, x);
      atomicAdd(dd[k] + 1, y);
      atomicAdd(dd[k] + 2, z);
    }
  }
}

// p = res * k for the vertex on the edge of memory axis d owned by extended voxel (ef, em, es):
// half-voxel key, memory axes -> logical axes, + shard origin (reference: unpack_*,
// marching_cubes.hpp:114-135 with offset 0, factor = captured resolution; transpose = the legacy
// get_mesh orientation, cMesher.hpp:128-149).
template <bool CO>
__device__ __forceinline__ void slot_position(const VolParams& vp, const Pass2Args& a, uint32_t ef, uint32_t em,
                                              uint32_t es, uint32_t d, float& p0, float& p1, float& p2) {
  const uint32_t hf = 2u * ef + (d == 0u), hm = 2u * em + (d == 1u), hs = 2u * es + (d == 2u);
  const float kx = __fadd_rn(0.0f, (float)((CO ? hs : hf) + 2u * vp.ox));
  const float ky = __fadd_rn(0.0f, (float)(hm + 2u * vp.oy));
  const float kz = __fadd_rn(0.0f, (float)((CO ? hf : hs) + 2u * vp.oz));
  if (a.transpose) { p0 = __fmul_rn(a.r0, kz); p1 = __fmul_rn(a.r1, ky); p2 = __fmul_rn(a.r2, kx); }
  else             { p0 = __fmul_rn(a.r0, kx); p1 = __fmul_rn(a.r1, ky); p2 = __fmul_rn(a.r2, kz); }
}

// per (case, triangle): three 9-bit fields, one per output corner, in the reference winding of
// Mesher.get: (E[T[3n+1]], E[T[3n]], E[T[3n+2]])  (marching_cubes.hpp:338-343 then cMesher.hpp:158-162).
// Field: slot | row delta of the owner voxel (ds * RM + dm) << 4 | its f offset << 8, i.e. the low
// 8 bits are the word offset of (row, plane) inside the staged region.  Filled by prepare_device.
constexpr int CASE_TRIS = 5;
__device__ __align__(16) uint32_t CASE_TAB_D[2][256 * CASE_TRIS];

template <bool CO>
inline void build_case_table(uint32_t* tab) {
  uint32_t info[12];
  for (int e = 0; e < 12; ++e) info[e] = edge_info<CO>(e);
  static const int order[3] = {1, 0, 2};
  for (int cs = 0; cs < 256; ++cs)
    for (int t = 0; t < CASE_TRIS; ++t) {
      uint32_t w = 0;
      for (int k = 0; k < 3; ++k) {
        const int ed = (int)((TRI_NIBBLES[cs] >> (12 * t + 4 * order[k])) & 0xFull);
        uint32_t v = 0;
        if (t < TRI_COUNT[cs] && ed < 12) {
          const uint32_t i = info[ed];
          const uint32_t slot = ((i >> 12) & 7u) + (((uint32_t)cs >> ((i >> 16) & 7u)) & 1u);
          v = slot | ((i & 0xFu) << 4) | (((i >> 4) & 1u) << 8);
        }
        w |= v << (9 * k);
      }
      tab[cs * CASE_TRIS + t] = w;
    }
}

__device__ __forceinline__ TileHdr load_hdr(const TileHdr* p) {
  union { uint4 q[2]; TileHdr h; } u;
  u.q[0] = __ldg(reinterpret_cast<const uint4*>(p));
  u.q[1] = __ldg(reinterpret_cast<const uint4*>(p) + 1);
  return u.h;
}

// k_emit: persistent, warp-specialised CTAs over the work list of non-empty (half) tiles.
//   producer warp : walks the CTA's tiles EMIT_STAGES ahead of the consumers.  Per tile: one TMA load
//                   brings the rowinfo of the (TM+1)(TS+1) rows x 2 segments whose slots the tile's
//                   cubes can reference; the tile's header and tl entries are copied to shared memory
//                   and the region is turned into per-(row, segment, plane) slot bases.
//   consumer warps: no CTA barrier -- a warp waits for a stage (mbarrier `ready`), does its share of the
//                   tile and releases the stage (mbarrier `empty`), so a slow warp never stalls the others.
//                   Records are read 32 per warp (next batch prefetched), their triangles expanded so that
//                   every lane owns one triangle: three slot lookups (two shared loads + popc -> perm) and
//                   three 4-byte stores, consecutive lanes writing consecutive 12-byte rows.  Then one
//                   thread per vertex slot writes the float32 vertex in its final form (reference:
//                   _normalize_mesh zmesh/_zmesh.pyx:423-433: three separately rounded operations, no FMA).
constexpr int RGN_ROWS = RM * RS;            // 81
constexpr int RGN_WORDS = 2 * RI_WORDS;      // two row segments per region row
constexpr int RGN_PAD = 1312;                // words per staged region buffer (5184 B rounded up to 128 B)
// (consumer warps, stages, CTAs per SM) measured on B200, c1 / c5s k_emit ms: (7,2,5) 0.952 / 0.216, (8,2,4) 0.969 /
// 0.221, (8,3,4) 0.988 / 0.226, (12,3,3) 0.987 / 0.224, (16,3,2) 0.988 / 0.225; CTA-barrier version 1.00 / 0.228
#ifndef ZM_EMIT_CTAS
#define ZM_EMIT_CTAS 5   // resident CTAs per SM the register allocation of k_emit is bounded for
#endif
#ifndef ZM_EMIT_CONSUMERS
#define ZM_EMIT_CONSUMERS 7
#endif
#ifndef ZM_EMIT_STAGES
#define ZM_EMIT_STAGES 2
#endif
constexpr int EMIT_NC = ZM_EMIT_CONSUMERS;           // consumer warps
constexpr int EMIT_ST = ZM_EMIT_STAGES;              // tiles in flight per CTA
constexpr int EMIT_THREADS = 32 * (EMIT_NC + 1);     // + the producer warp
constexpr int TLC = 64;                      // tile-local labels whose tl entry is cached in shared memory
constexpr uint32_t EMIT_END = 0xFFFFFFFFu;   // header.tile of the end marker the producer publishes after its last tile

__device__ __forceinline__ void mbar_arrive(u64* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// SLAB: 0 unsharded; 1 slab shard, tiles that never touch the next shard's boundary plane (face indices get the label's
// cross-shard offset); 2 slab shard, tiles whose cubes may reference that plane (the top tile layer)
template <bool CO, bool NORMALS, int SLAB>
__global__ void __launch_bounds__(EMIT_THREADS, NORMALS ? 3 : ZM_EMIT_CTAS)
k_emit(const VolParams vp, const __grid_constant__ CUtensorMap rmap, const Pass2Args a) {
  constexpr uint32_t FULL = 0xffffffffu;
  __shared__ __align__(128) uint32_t R[EMIT_ST][RGN_PAD];       // TMA destinations: rowinfo of the region
  __shared__ uint32_t rbs[EMIT_ST][RGN_ROWS * RGN_WORDS];       // spatial id of the first slot of (row, segment, plane)
  __shared__ __align__(16) TLEntry tlss[EMIT_ST][TLC];
  __shared__ __align__(16) TileHdr s_hdr[EMIT_ST];
  __shared__ u64 full[EMIT_ST], ready[EMIT_ST], empty[EMIT_ST];
  __shared__ uint32_t s_tab[256 * CASE_TRIS];
  __shared__ uint8_t s_tricount[256];
  __shared__ u64 rf[EMIT_NC][32];              // per record of the warp's batch: first face row
  __shared__ u64 rv[NORMALS ? EMIT_NC : 1][32];  //                                first vertex row of the label
  __shared__ uint32_t ru[EMIT_NC][32];         //                                  region row * 16 | f << 11 | case << 16
  __shared__ uint32_t rvo[SLAB ? EMIT_NC : 1][32];  //                             index offset of the label (earlier shards)
  __shared__ uint8_t tlist[EMIT_NC][160];      // triangles of the batch: record lane << 3 | t

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t ltm = (1u << lane) - 1u;
  const uint32_t n = a.n_work, G = gridDim.x;
  const uint32_t first = blockIdx.x;
  if (first >= n) return;
  const uint32_t ntile = (n - first + G - 1) / G;  // tiles of this CTA: first, first + G, ...

  if (tid == 0) {
    for (int s = 0; s < EMIT_ST; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&ready[s], 1);
      mbar_init(&empty[s], EMIT_NC);
    }
    fence_mbar_init();
  }
  for (int j = tid; j < 256; j += EMIT_THREADS) s_tricount[j] = TRI_COUNT_D[j];
  for (int j = tid; j < 256 * CASE_TRIS; j += EMIT_THREADS) s_tab[j] = CASE_TAB_D[CO ? 1 : 0][j];
  __syncthreads();

  if (warp == EMIT_NC) {
    // ================= producer warp =================
    // The CTA's headers are read 32 at a time (lane l: tile first + (base + l) * G; the next batch is in flight while
    // the current one is published), filtered by tile layer for the split launches of slab shards, and the ones that
    // pass are published one by one.
    uint32_t pub = 0;  // tiles published so far (the consumers see exactly these, then the end marker)
    union HdrWords { uint4 q[2]; TileHdr h; };
    HdrWords nxt;
    nxt.q[0] = nxt.q[1] = make_uint4(0u, 0u, 0u, 0u);
    if ((uint32_t)lane < ntile) nxt.h = load_hdr(a.hdr + first + (size_t)lane * G);
    for (uint32_t base = 0; base < ntile; base += 32) {
      const HdrWords cur = nxt;
      const bool have = base + lane < ntile;
      if (base + 32 + lane < ntile) nxt.h = load_hdr(a.hdr + first + (size_t)(base + 32 + lane) * G);
      const bool pass = have && (SLAB == 0 || a.layer_mode == 0u || ((cur.h.tile >= a.top_tile_lo) == (a.layer_mode == 2u)));
      uint32_t todo = __ballot_sync(FULL, pass);
      while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1u;
        HdrWords hw;  // every lane holds the header of lane `src`
        hw.q[0] = make_uint4(__shfl_sync(FULL, cur.q[0].x, src), __shfl_sync(FULL, cur.q[0].y, src),
                             __shfl_sync(FULL, cur.q[0].z, src), __shfl_sync(FULL, cur.q[0].w, src));
        hw.q[1] = make_uint4(__shfl_sync(FULL, cur.q[1].x, src), __shfl_sync(FULL, cur.q[1].y, src),
                             __shfl_sync(FULL, cur.q[1].z, src), __shfl_sync(FULL, cur.q[1].w, src));
        const TileHdr& hn = hw.h;
        const int s = pub % EMIT_ST;
        const uint32_t k = pub / EMIT_ST;  // k-th use of stage s
        ++pub;
        if (k > 0) mbar_wait(&empty[s], (k - 1) & 1u);  // every consumer warp has released the stage
        if (lane == 0) {
          s_hdr[s] = hn;
          uint32_t b = hn.tile;
          const uint32_t tf = b % vp.ntf;
          b /= vp.ntf;
          // The region buffers are only ever READ through the generic proxy and those reads are ordered before
          // this refill by the `empty` mbarrier, so no proxy fence is needed.
          mbar_expect_tx(&full[s], (uint32_t)(RGN_ROWS * RGN_WORDS * 4));
          tma_load_3d(R[s], &rmap, &full[s], (int)(tf * RI_WORDS), (int)((b % vp.ntm) * TM), (int)((b / vp.ntm) * TS));
        }
        const uint32_t nlab = hn.nlab, tlbase = hn.tlbase;
        for (uint32_t j = lane; j < nlab && j < (uint32_t)TLC; j += 32) tlss[s][j] = a.tl[tlbase + j];
        mbar_wait(&full[s], k & 1u);  // every lane waits itself: the TMA writes are visible to it afterwards
        for (int e = lane; e < RGN_ROWS * 2; e += 32) {
          const uint32_t* seg = R[s] + e * RI_WORDS;
          uint32_t run = seg[6];
#pragma unroll
          for (int p = 0; p < 6; ++p) {
            rbs[s][e * RI_WORDS + p] = run;
            run += __popc(seg[p]);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&ready[s]);  // (release: header, tl cache and slot bases are visible to the waiters)
      }
    }
    {  // end marker: a header without a tile
      const int s = pub % EMIT_ST;
      const uint32_t k = pub / EMIT_ST;
      if (k > 0) mbar_wait(&empty[s], (k - 1) & 1u);
      if (lane == 0) {
        s_hdr[s].tile = EMIT_END;
        mbar_arrive(&ready[s]);
      }
    }
    return;
  }

  // ================= consumer warps =================
  const int cw = warp;
  for (uint32_t it = 0;; ++it) {
    const int s = it % EMIT_ST;
    const uint32_t k = it / EMIT_ST;
    mbar_wait(&ready[s], k & 1u);
    TileHdr h;
    {
      union { uint4 q[2]; TileHdr h; } u;
      u.q[0] = reinterpret_cast<const uint4*>(&s_hdr[s])[0];
      u.q[1] = reinterpret_cast<const uint4*>(&s_hdr[s])[1];
      h = u.h;
    }
    if (h.tile == EMIT_END) break;  // the producer has published all of the CTA's tiles
    mbar_wait(&full[s], k & 1u);  // (already complete: makes the TMA-written region visible to this thread)
    uint32_t b = h.tile;
    const uint32_t tf = b % vp.ntf;
    b /= vp.ntf;
    const uint32_t tm = b % vp.ntm, ts = b / vp.ntm;
    const uint32_t ef0 = tf * TF, em0 = tm * TM, es0 = ts * TS;
    const uint32_t* const Rc = R[s];
    const uint32_t* const rb = rbs[s];
    const TLEntry* const tls = tlss[s];
    // slab sharding: region rows [ftop9, ftop9 + RM) lie on the plane owned by the next shard
    uint32_t ftop9 = 0xFFFFu;
    if (SLAB == 2 && vp.Es_own < vp.Es && vp.Es_own >= es0 && vp.Es_own - es0 < (uint32_t)RS) ftop9 = (vp.Es_own - es0) * RM;

    // ---- faces ----
    if (a.write_faces || NORMALS) {
      const uint32_t nrec = h.nrec;
      const u64* recp = a.rec + h.recbase;
      uint32_t base = cw * 32;
      u64 wnext = base + lane < nrec ? __ldg(recp + base + lane) : 0ull;
      for (; base < nrec; base += EMIT_NC * 32) {
        const bool valid = base + lane < nrec;
        const u64 w64 = wnext;
        wnext = base + EMIT_NC * 32 + lane < nrec ? __ldg(recp + base + EMIT_NC * 32 + lane) : 0ull;  // next batch in flight
        const uint32_t w = (uint32_t)w64;
        const uint32_t vidx = w & 0x7FFu, cs = (w >> 11) & 0xFFu, ci = w >> 19;
        const uint32_t nt = valid ? s_tricount[cs] : 0u;
        uint32_t ntot;
        const uint32_t tpre = warp_prefix3(nt, ltm, ntot);
        if (valid) {
          TLEntry e;
          if (ci < (uint32_t)TLC) e = tls[ci];
          else e = a.tl[h.tlbase + ci];
          const uint32_t lf = vidx & 31u, lm = (vidx >> 5) & 7u, ls = vidx >> 8;
          ru[cw][lane] = ((ls * RM + lm) << 4) | (lf << 11) | (cs << 16);
          rf[cw][lane] = e.b + (uint32_t)(w64 >> 32);
          if (SLAB) rvo[cw][lane] = (uint32_t)(e.a >> 32);
          if (NORMALS) rv[cw][lane] = e.a & 0xFFFFFFFFull;
          for (uint32_t t = 0; t < nt; ++t) tlist[cw][tpre + t] = (uint8_t)((lane << 3) | t);
        }
        __syncwarp();
        for (uint32_t q = lane; q < ntot; q += 32) {
          const uint32_t tr = tlist[cw][q];
          const uint32_t src = tr >> 3, t = tr & 7u;
          const uint32_t uc = ru[cw][src];
          const uint32_t u0 = uc & 0x7FFu, lf0 = (uc >> 11) & 31u;
          const uint32_t tab = s_tab[(uc >> 16) * CASE_TRIS + t];
          const uint32_t voff = SLAB ? rvo[cw][src] : 0u;
          uint32_t vi[3];
          uint32_t fslot[3];  // (SLAB && NORMALS) boundary-plane slot of a corner owned by the next shard, else ~0
          float p[3][3];
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            fslot[c] = 0xFFFFFFFFu;
            const uint32_t en = (tab >> (9 * c)) & 0x1FFu;
            const uint32_t lfx = lf0 + (en >> 8), slot = en & 7u;
            const uint32_t uw = u0 + (en & 0xFFu);  // word of (row, plane) in the first segment
            const uint32_t rr = uw >> 4;
            // slot = 2*axis + side; side 0: the owner (lower) voxel carries the label, 1: the upper one
            if (SLAB == 2 && rr - ftop9 < (uint32_t)RM) {
              const uint32_t fs = 4u * ((em0 + rr - ftop9) * vp.Efp + ef0 + lfx) + slot;  // (zm_mesh_slab rejects planes of 2^32 or more slots)
              vi[c] = __ldg(a.foreign + fs);
              if (NORMALS) fslot[c] = fs;
            } else {
              const uint32_t idx = uw + ((lfx >> 5) << 3);
              const uint32_t g = rb[idx] + __popc(Rc[idx] & ((1u << (lfx & 31u)) - 1u));
              vi[c] = __ldg(a.perm + g) + voff;
            }
            if (NORMALS) {
              const uint32_t rs_ = (rr * 57u) >> 9;  // rr / 9 for rr < 90
              slot_position<CO>(vp, a, ef0 + lfx, em0 + rr - rs_ * RM, es0 + rs_, slot >> 1, p[c][0], p[c][1], p[c][2]);
            }
          }
          if (a.write_faces) {
            uint32_t* f = a.faces + 3ull * (rf[cw][src] + t);
            f[0] = vi[0]; f[1] = vi[1]; f[2] = vi[2];
          }
          if (NORMALS) {
            float4* nb = a.nacc + rv[cw][src];
            float* d0 = reinterpret_cast<float*>(nb + (vi[0] - voff));
            float* d1 = reinterpret_cast<float*>(nb + (vi[1] - voff));
            float* d2 = reinterpret_cast<float*>(nb + (vi[2] - voff));
            bool q0 = true, q1 = true, q2 = true;  // float4 rows of this shard (one vector atomic each)
            if (SLAB == 2) {  // vertices of the next shard: accumulate in the plane buffer that is sent to it (3-float rows)
              if (fslot[0] != 0xFFFFFFFFu) { d0 = a.fnormals + 3ull * fslot[0]; q0 = false; }
              if (fslot[1] != 0xFFFFFFFFu) { d1 = a.fnormals + 3ull * fslot[1]; q1 = false; }
              if (fslot[2] != 0xFFFFFFFFu) { d2 = a.fnormals + 3ull * fslot[2]; q2 = false; }
            }
            // legacy faces (t0,t2,t1) = the stored row reversed
            if (a.transpose) face_normal_scatter(p[2], p[1], p[0], d2, d1, d0, q2, q1, q0);
            else face_normal_scatter(p[0], p[1], p[2], d0, d1, d2, q0, q1, q2);
          }
        }
        __syncwarp();
      }
    }

    // ---- vertices ----
    if (a.write_verts) {
#pragma unroll 2
      for (uint32_t v_ = tid; v_ < h.nslots; v_ += EMIT_NC * 32) {
        const uint32_t rank = __ldg(a.perm + h.gbase + v_);
        const uint32_t w = __ldg(a.vinfo + h.gbase + v_);
        const uint32_t vidx = w & 0x7FFu, s6 = (w >> 11) & 7u, ci = w >> 14;
        const u64 ea = ci < (uint32_t)TLC ? tls[ci].a : a.tl[h.tlbase + ci].a;
        const u64 dst = (ea & 0xFFFFFFFFull) + rank;
        float p0, p1, p2;
        slot_position<CO>(vp, a, ef0 + (vidx & 31u), em0 + ((vidx >> 5) & 7u), es0 + (vidx >> 8), s6 >> 1, p0, p1, p2);
        if (a.voxel_centered) { p0 = __fadd_rn(p0, a.c0); p1 = __fadd_rn(p1, a.c1); p2 = __fadd_rn(p2, a.c2); }
        float* v = a.verts + 3ull * dst;
        v[0] = __fmul_rn(p0, 0.5f);  // == p / 2.0f exactly
        v[1] = __fmul_rn(p1, 0.5f);
        v[2] = __fmul_rn(p2, 0.5f);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);  // this warp is done with the stage
  }
}

// slab sharding: final (cross-shard) indices of the in-plane slots of this shard's FIRST plane, for
// the shard below whose cubes reference them: dst[(em * Efp + ef) * 4 + slot] = label offset + rank.
// One warp per work-list entry (grid-stride): entries of other planes cost one header load.
constexpr int NT_V = 128;
template <bool CO>
__global__ void __launch_bounds__(NT_V) k_export_plane(const VolParams vp, const Pass2Args a, uint32_t* dst) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t nwarps = gridDim.x * (NT_V / 32);
  const uint32_t plane_tiles = vp.ntf * vp.ntm;
  for (uint32_t w = blockIdx.x * (NT_V / 32) + (threadIdx.x >> 5); w < a.n_work; w += nwarps) {
    const TileHdr h = load_hdr(a.hdr + w);
    if (h.tile >= plane_tiles || h.nslots == 0) continue;  // not in the first tile layer
    const uint32_t tf = h.tile % vp.ntf, tm = h.tile / vp.ntf;
    const TLEntry* tl = a.tl + h.tlbase;
    for (uint32_t i = lane; i < h.nslots; i += 32) {
      const uint32_t v = __ldg(a.vinfo + h.gbase + i);
      const uint32_t vidx = v & 0x7FFu, s6 = (v >> 11) & 7u, ci = v >> 14;
      if ((vidx >> 8) != 0u || s6 >= 4u) continue;
      const uint32_t ef = tf * TF + (vidx & 31u), em = tm * TM + ((vidx >> 5) & 7u);
      dst[4ull * ((size_t)em * vp.Efp + ef) + s6] = (uint32_t)(tl[ci].a >> 32) + __ldg(a.perm + h.gbase + i);
    }
  }
}

// slab sharding + normals: add the contributions the shard below accumulated for this shard's first-plane
// vertices (src[(em * Efp + ef) * 4 + slot][3], same indexing as k_export_plane); every vertex is touched once.
template <bool CO>
__global__ void __launch_bounds__(NT_V) k_import_plane_normals(const VolParams vp, const Pass2Args a, const float* src) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t nwarps = gridDim.x * (NT_V / 32);
  const uint32_t plane_tiles = vp.ntf * vp.ntm;
  for (uint32_t w = blockIdx.x * (NT_V / 32) + (threadIdx.x >> 5); w < a.n_work; w += nwarps) {
    const TileHdr h = load_hdr(a.hdr + w);
    if (h.tile >= plane_tiles || h.nslots == 0) continue;
    const uint32_t tf = h.tile % vp.ntf, tm = h.tile / vp.ntf;
    const TLEntry* tl = a.tl + h.tlbase;
    for (uint32_t i = lane; i < h.nslots; i += 32) {
      const uint32_t v = __ldg(a.vinfo + h.gbase + i);
      const uint32_t vidx = v & 0x7FFu, s6 = (v >> 11) & 7u, ci = v >> 14;
      if ((vidx >> 8) != 0u || s6 >= 4u) continue;
      const uint32_t ef = tf * TF + (vidx & 31u), em = tm * TM + ((vidx >> 5) & 7u);
      const float* in = src + 3ull * (4ull * ((size_t)em * vp.Efp + ef) + s6);
      float4* nn = a.nacc + ((tl[ci].a & 0xFFFFFFFFull) + __ldg(a.perm + h.gbase + i));
      float4 acc = *nn;
      acc.x = __fadd_rn(acc.x, in[0]);
      acc.y = __fadd_rn(acc.y, in[1]);
      acc.z = __fadd_rn(acc.z, in[2]);
      *nn = acc;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Mesher.compute_normals on an arbitrary float32 mesh (zmesh/_zmesh.pyx:138-152).
__global__ void __launch_bounds__(256) k_normals_accumulate_f32(const float* __restrict__ verts,
                                                               const uint32_t* __restrict__ faces,
                                                               u64 nT, float* normals) {
  u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  const u64 stride = (u64)gridDim.x * blockDim.x;
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

__global__ void __launch_bounds__(256) k_normals_normalize(float* normals, u64 nV) {
  u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  const u64 stride = (u64)gridDim.x * blockDim.x;
  for (; i < nV; i += stride) {
    float x = normals[3 * i], y = normals[3 * i + 1], z = normals[3 * i + 2];
    float l = len3(x, y, z);
    if (l != 1.0f) { x = __fdiv_rn(x, l); y = __fdiv_rn(y, l); z = __fdiv_rn(z, l); }  // 0/0 -> NaN like hat()
    normals[3 * i] = x; normals[3 * i + 1] = y; normals[3 * i + 2] = z;
  }
}

// pass 2's float4 accumulation rows -> unit normals in the packed [V][3] output (hat(): 0/0 -> NaN like the reference)
__global__ void __launch_bounds__(256) k_normals_normalize4(const float4* __restrict__ acc, float* normals, u64 nV) {
  u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  const u64 stride = (u64)gridDim.x * blockDim.x;
  for (; i < nV; i += stride) {
    const float4 v = acc[i];
    float x = v.x, y = v.y, z = v.z;
    const float l = len3(x, y, z);
    if (l != 1.0f) { x = __fdiv_rn(x, l); y = __fdiv_rn(y, l); z = __fdiv_rn(z, l); }
    normals[3 * i] = x; normals[3 * i + 1] = y; normals[3 * i + 2] = z;
  }
}

// ---------------------------------------------------------------------------------------------
// Mesh wire format on the device: the Neuroglancer "Precomputed" layout of every label --
// uint32 Nv | float32 vertices [Nv][3] | uint32 faces [Nf][3] (zmesh/mesh.py:257-269) -- gathered from pass 2's final
// arrays into ONE buffer in label order, so that a single device-to-host transfer lands ready-to-write objects.
// word[i] .. word[i + 1] = the 4-byte words of label i's object.  A CTA copies chunks of PACK_CHUNK words; the label
// of a chunk's first word is found by binary search, later ones by walking the (monotone) table.
constexpr int PACK_CHUNK = 4096;
struct PackArgs {
  const u64* word;   // [nl + 1] first output word of every label
  const u64* voff;   // [nl] first vertex row,  nv[nl] vertex count
  const u64* nv;
  const u64* foff;   // [nl] first face row
  const uint32_t* verts;  // float32 bit patterns [V][3]
  const uint32_t* faces;  // [T][3]
  uint32_t* out;
  u64 total;         // word[nl]
  uint32_t nl;
};
__global__ void __launch_bounds__(256) k_pack_precomputed(const PackArgs a) {
  __shared__ uint32_t s_first;
  const u64 nchunks = (a.total + PACK_CHUNK - 1) / PACK_CHUNK;
  for (u64 c = blockIdx.x; c < nchunks; c += gridDim.x) {
    const u64 w0 = c * PACK_CHUNK;
    if (threadIdx.x == 0) {
      uint32_t lo = 0, hi = a.nl - 1;  // largest i with word[i] <= w0
      while (lo < hi) {
        const uint32_t mid = (lo + hi + 1) >> 1;
        if (a.word[mid] <= w0) lo = mid; else hi = mid - 1;
      }
      s_first = lo;
    }
    __syncthreads();
    uint32_t i = s_first;
    const u64 w1 = w0 + PACK_CHUNK < a.total ? w0 + PACK_CHUNK : a.total;
    for (u64 w = w0 + threadIdx.x; w < w1; w += blockDim.x) {
      while (w >= a.word[i + 1]) ++i;
      const u64 r = w - a.word[i];
      const u64 nv = a.nv[i];
      uint32_t val;
      if (r == 0) val = (uint32_t)nv;
      else if (r - 1 < 3 * nv) val = a.verts[3 * a.voff[i] + (r - 1)];
      else val = a.faces[3 * a.foff[i] + (r - 1 - 3 * nv)];
      a.out[w] = val;
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// synthetic jittered-grid Voronoi volume (benchmark/test input)

__host__ __device__ __forceinline__ u64 splitmix64(u64 x) {
  x += 0x9E3779B97F4A7C15ull;
  u64 z = x;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

struct SynthArgs {
  void* dst;
  u64 n;                // voxels of the block
  uint32_t sx, sy, sz;  // block shape (logical)
  uint32_t ox, oy, oz;  // block origin inside the full volume
  uint32_t gx, gy, gz;  // cells per axis of the full volume
  uint32_t pitch;
  u64 seed;
  int c_order;
};

template <typename L>
__global__ void __launch_bounds__(256) k_synth_voronoi(const SynthArgs a) {
  u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  const u64 stride = (u64)gridDim.x * blockDim.x;
  for (; i < a.n; i += stride) {
    uint32_t lx, ly, lz;
    if (a.c_order) { lz = (uint32_t)(i % a.sz); u64 t = i / a.sz; ly = (uint32_t)(t % a.sy); lx = (uint32_t)(t / a.sy); }
    else           { lx = (uint32_t)(i % a.sx); u64 t = i / a.sx; ly = (uint32_t)(t % a.sy); lz = (uint32_t)(t / a.sy); }
    const long long x = (long long)lx + a.ox, y = (long long)ly + a.oy, z = (long long)lz + a.oz;
    const int bi = (int)(x / a.pitch), bj = (int)(y / a.pitch), bk = (int)(z / a.pitch);
    long long best_d = 0x7FFFFFFFFFFFFFFFll;
    u64 best_c = 0;
    for (int dk = -1; dk <= 1; ++dk)
      for (int dj = -1; dj <= 1; ++dj)
        for (int di = -1; di <= 1; ++di) {
          const int ni = bi + di, nj = bj + dj, nk = bk + dk;
          if (ni < 0 || nj < 0 || nk < 0 || ni >= (int)a.gx || nj >= (int)a.gy || nk >= (int)a.gz) continue;
          const u64 c = (u64)ni + (u64)a.gx * ((u64)nj + (u64)a.gy * nk);
          const u64 h = splitmix64(c ^ a.seed);
          const long long sx = (long long)ni * a.pitch + (long long)(((h & 0xFFFFull) * a.pitch) >> 16);
          const long long sy = (long long)nj * a.pitch + (long long)((((h >> 16) & 0xFFFFull) * a.pitch) >> 16);
          const long long sz = (long long)nk * a.pitch + (long long)((((h >> 32) & 0xFFFFull) * a.pitch) >> 16);
          const long long d = (x - sx) * (x - sx) + (y - sy) * (y - sy) + (z - sz) * (z - sz);
          if (d < best_d || (d == best_d && c < best_c)) { best_d = d; best_c = c; }
        }
    u64 lab = sizeof(L) == 8 ? (splitmix64(best_c + 1ull) | 1ull) : (best_c + 1ull);
    static_cast<L*>(a.dst)[i] = (L)lab;
  }
}

}  // namespace zm

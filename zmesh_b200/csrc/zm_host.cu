// zmesh_b200 host layer: the C ABI of include/zmesh_b200.h over the sm_100a kernels.
//
// Host-side counterpart of the reference's CMesher facade (zmesh/cMesher.hpp:16-308): owns the
// per-label results on the device, orchestrates classify -> scan -> emit -> final gather on one
// CUDA stream, and hands label ranges back to the caller.  No CPU compute path exists here.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>  // types only: the library is bound at run time (nccl_api), the single-GPU path never needs it

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/zmesh_b200.h"
#include "zm_kernels.cuh"

namespace {

thread_local std::string g_create_error;

struct DevBuf {  // grow-only device allocation cached in the handle across calls
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap && p) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes < 256 ? 256 : bytes;
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <typename T> T* as() const { return static_cast<T*>(p); }
};

struct LabelRec {
  uint64_t label, nv, nt, voff, foff;
  bool erased;
};

struct FinalState {
  bool partial = false;  // slab shard: zm_finalize_begin has emitted all tiles but the top layer (partial_args: its arguments)
  int partial_args[3] = {0, 0, 0};
  bool faces_valid = false, verts_valid = false, normals_valid = false;
  bool normals_pending = false;  // slab shard: accumulated, awaiting zm_add_normal_plane / zm_finish_normals
  int voxel_centered = 0, transpose = 0, normals_transpose = 0;
  float off[3] = {0, 0, 0};
};

}  // namespace

struct zm_handle {
  int device = 0;
  int num_sms = 148;
  cudaStream_t stream = nullptr;      // the stream all work is queued on
  cudaStream_t own_stream = nullptr;  // created by zm_create (stream == own_stream unless zm_set_stream)
  cudaEvent_t ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  float res[3] = {1, 1, 1};
  std::string err;

  // capacity guesses carried between calls (per voxel)
  uint32_t hash_cap = 1u << 16;
  double perm_ratio = 0.12, rec_ratio = 0.12, tl_ratio = 0.004;

  // scratch + intermediates (device)
  DevBuf d_vol, d_keys, d_cnt, d_offV, d_offT, d_list, d_partial, d_ctl;
  DevBuf d_rowinfo, d_perm, d_vinfo, d_rec, d_tl, d_hdr, d_dense;
  // results (device)
  DevBuf d_faces, d_verts, d_normals;
  DevBuf d_pack, d_packtab;  // Precomputed objects of all labels (zm_pack_precomputed) and their offset table
  std::vector<uint64_t> pack_labels, pack_off;  // ids in ascending order; byte offset of every object (+ the end)
  DevBuf d_nacc;  // float4 accumulation rows of the normals (pass 2 adds one vector atomic per triangle corner)
  DevBuf d_voff, d_tmpA, d_tmpB;     // slab sharding: per-table-slot index offsets; upload scratch
  bool tl_fixed = false, have_voff = false, slab_mode = false;
  bool dir_exchange = false;  // voff came from zm_import_directories: the exchange's overflow flag is checked by finalize
  // native multi-GPU step (zm_comm_init / zm_slab_step): one communicator for the collectives on `stream`, one for the
  // neighbour transfers on `comm_stream` (so that they overlap pass 2), buffers owned by the handle
  ncclComm_t comm = nullptr;  // all shards: the directory all-gather
  // One 2-rank communicator per neighbour pair (lower: with shard rank - 1, upper: with shard rank + 1): a send/recv
  // inside an 8-rank communicator gets a fraction of the channels (measured: 64 MiB neighbour shift 0.80 ms = 83 GB/s at
  // 8 ranks vs 0.14 ms = 490 GB/s in a 2-rank communicator)
  ncclComm_t comm_lo = nullptr, comm_hi = nullptr;
  int world = 1, rank = 0;
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t ev_comm[2] = {nullptr, nullptr};
  uint64_t dir_cap = 1ull << 15;
  DevBuf d_dir_mine, d_dir_all, d_plane_send, d_plane_recv, d_nplane_out, d_nplane_in;
  const uint32_t* foreign = nullptr;  // borrowed: boundary-plane indices received from the next shard
  float* nplane_out = nullptr;        // borrowed: normal contributions to the next shard's first-plane vertices
  uint64_t capL = 0;
  zm::Control* h_ctl = nullptr;  // pinned
  uint64_t* h_list = nullptr;  // pinned, grow-only: (label, nV, nT) per label, copied asynchronously by zm_mesh
  size_t h_list_cap = 0;
  uint64_t nlabels = 0;
  bool dir_pending = false;    // the host-side directory (recs / index / sorted_ids) has not been built from h_list yet
  cudaEvent_t ev_list = nullptr;

  // what pass 2 needs to know about the meshed volume
  zm::VolParams vp{};
  bool c_order = false;
  uint32_t n_work = 0;

  // results (host)
  bool has_result = false;
  bool failed = false;  // the last zm_mesh returned an error: results are undefined until the next one succeeds
  uint64_t Vtot = 0, Ttot = 0;
  std::vector<LabelRec> recs;                   // storage (table) order
  std::unordered_map<uint64_t, uint32_t> index;  // label -> recs index
  std::vector<uint64_t> sorted_ids;
  std::vector<uint64_t> bulk_labels, bulk_voff, bulk_foff;
  FinalState fin;
  zm_stats_t stats{};
};

namespace {

using namespace zm;

#define ZM_CUDA(h, call)                                                                       \
  do {                                                                                         \
    cudaError_t _e = (call);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      char _b[512];                                                                            \
      snprintf(_b, sizeof(_b), "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
      (h)->err = _b;                                                                           \
      cudaGetLastError();                                                                      \
      return _e == cudaErrorMemoryAllocation ? ZM_ERR_OOM : ZM_ERR_CUDA;                       \
    }                                                                                          \
  } while (0)

int fail(zm_handle* h, int code, const std::string& msg) {
  h->err = msg;
  return code;
}

typedef void (*classify_fn)(const VolParams, const CUtensorMap, const Pass1Args);
typedef void (*pass2_fn)(const VolParams, const CUtensorMap, const Pass2Args);

struct KernelSet {
  classify_fn classify[2];  // MODE 0 / MODE 1
  size_t smem[2];
  int row_pad;              // staged row length (elements) = TMA box extent along f
};

template <typename L, bool CO>
KernelSet make_set() {
  KernelSet k;
  k.classify[0] = k_classify<L, CO, 0>;
  k.classify[1] = k_classify<L, CO, 1>;
  k.smem[0] = sizeof(P1Smem<L, 0>);
  k.smem[1] = sizeof(P1Smem<L, 1>);
  k.row_pad = RowPad<L>::value;
  return k;
}

KernelSet kernel_set(int label_bytes, bool c_order) {
  switch (label_bytes) {
    case 1: return c_order ? make_set<uint8_t, true>() : make_set<uint8_t, false>();
    case 2: return c_order ? make_set<uint16_t, true>() : make_set<uint16_t, false>();
    case 4: return c_order ? make_set<uint32_t, true>() : make_set<uint32_t, false>();
    default: return c_order ? make_set<unsigned long long, true>() : make_set<unsigned long long, false>();
  }
}

template <bool CO, bool NORMALS>
pass2_fn emit_kernel_slab(int slab) {
  return slab == 0 ? k_emit<CO, NORMALS, 0> : (slab == 1 ? k_emit<CO, NORMALS, 1> : k_emit<CO, NORMALS, 2>);
}
pass2_fn emit_kernel(bool c_order, bool normals, int slab) {
  if (c_order) return normals ? emit_kernel_slab<true, true>(slab) : emit_kernel_slab<true, false>(slab);
  return normals ? emit_kernel_slab<false, true>(slab) : emit_kernel_slab<false, false>(slab);
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess) {
      cudaGetLastError();
      return nullptr;
    }
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// TMA descriptor of the label volume (memory order f, m, s), box = the staged tile region.
// Returns false when the volume does not meet TMA's 16-byte base/stride rules (the kernel then
// stages tiles with ordinary loads).
bool make_tensor_map(CUtensorMap* tm, const void* data, int label_bytes, uint32_t nf, uint32_t nm, uint32_t ns,
                     int row_pad) {
  memset(tm, 0, sizeof(*tm));
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return false;
  if (getenv("ZMESH_B200_NO_TMA")) return false;
  const uint64_t eb = (uint64_t)label_bytes;
  if (((uintptr_t)data & 15u) != 0 || ((uint64_t)nf * eb) % 16 != 0) return false;
  CUtensorMapDataType dt = label_bytes == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8
                         : label_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT16
                         : label_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_UINT32
                                            : CU_TENSOR_MAP_DATA_TYPE_UINT64;
  cuuint64_t dims[3] = {nf, nm, ns};
  cuuint64_t strides[2] = {(cuuint64_t)nf * eb, (cuuint64_t)nf * nm * eb};
  cuuint32_t box[3] = {(cuuint32_t)row_pad, (cuuint32_t)RM, (cuuint32_t)RS};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(tm, dt, 3, const_cast<void*>(data), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// TMA descriptor of rowinfo viewed as uint32 [Es][Em][ntf * RI_WORDS]; box = the (TM+1)(TS+1) rows x
// two row segments a tile's cubes can reference (k_emit).  Out-of-range rows / segments read as zero
// (= no slots).
bool make_rowinfo_map(CUtensorMap* tm, const void* rowinfo, const VolParams& vp) {
  memset(tm, 0, sizeof(*tm));
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return false;
  cuuint64_t dims[3] = {(cuuint64_t)vp.ntf * RI_WORDS, vp.Em, vp.Es};
  cuuint64_t strides[2] = {(cuuint64_t)vp.ntf * RI_WORDS * 4, (cuuint64_t)vp.ntf * RI_WORDS * 4 * vp.Em};
  cuuint32_t box[3] = {(cuuint32_t)RGN_WORDS, (cuuint32_t)RM, (cuuint32_t)RS};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, const_cast<void*>(rowinfo), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

int prepare_device(zm_handle* h) {
  // case tables -> device globals; opt in to > 48 KB of shared memory per CTA
  ZM_CUDA(h, cudaMemcpyToSymbol(TRI_COUNT_D, TRI_COUNT, sizeof(TRI_COUNT)));
  static_assert(sizeof(TRI_NIBBLES) == 256 * sizeof(unsigned long long), "table size");
  ZM_CUDA(h, cudaMemcpyToSymbol(TRI_NIBBLES_D, TRI_NIBBLES, sizeof(TRI_NIBBLES)));
  {
    std::vector<uint32_t> tab(2 * 256 * CASE_TRIS);
    build_case_table<false>(tab.data());
    build_case_table<true>(tab.data() + 256 * CASE_TRIS);
    ZM_CUDA(h, cudaMemcpyToSymbol(CASE_TAB_D, tab.data(), tab.size() * sizeof(uint32_t)));
  }
  for (int lb : {1, 2, 4, 8})
    for (int co = 0; co < 2; ++co) {
      KernelSet ks = kernel_set(lb, co != 0);
      for (int mode = 0; mode < 2; ++mode)
        ZM_CUDA(h, cudaFuncSetAttribute((const void*)ks.classify[mode], cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)ks.smem[mode]));
    }
  return ZM_OK;
}

void drop_results(zm_handle* h) {
  h->has_result = false;
  h->Vtot = h->Ttot = 0;
  h->n_work = 0;
  h->recs.clear();
  h->index.clear();
  h->sorted_ids.clear();
  h->dir_pending = false;
  h->nlabels = 0;
  h->pack_labels.clear();
  h->pack_off.clear();
  h->bulk_labels.clear();
  h->bulk_voff.clear();
  h->bulk_foff.clear();
  h->fin = FinalState();
}

uint32_t grid_for(unsigned long long n, int block) {
  unsigned long long g = (n + block - 1) / block;
  const unsigned long long cap = 148ull * 16ull;
  if (g > cap) g = cap;
  if (g == 0) g = 1;
  return (uint32_t)g;
}

int run_mesh(zm_handle* h, const void* labels, int label_bytes, uint64_t sx, uint64_t sy, uint64_t sz,
             int c_order, int close, int mem_kind, const zm_slab* slab) {
  if (!h) return ZM_ERR_INVALID;
  h->err.clear();
  drop_results(h);  // Mesher.mesh deletes the previous CMesher first (zmesh/_zmesh.pyx:469)
  if (label_bytes != 1 && label_bytes != 2 && label_bytes != 4 && label_bytes != 8)
    return fail(h, ZM_ERR_INVALID, "label_bytes must be 1, 2, 4 or 8");
  if (mem_kind != ZM_MEM_HOST && mem_kind != ZM_MEM_DEVICE) return fail(h, ZM_ERR_INVALID, "bad mem_kind");
  const uint64_t pad = close ? 1 : 0;
  const uint64_t lim = (1ull << 20) - 4;
  if (sx > lim || sy > lim || sz > lim || (slab && slab->full_extent > lim))
    return fail(h, ZM_ERR_UNSUPPORTED, "extent exceeds the 21-bit half-voxel key range (2^20 voxels per axis)");
  h->tl_fixed = false;
  h->have_voff = false;
  h->dir_exchange = false;
  h->foreign = nullptr;
  h->nplane_out = nullptr;
  h->slab_mode = slab != nullptr;
  ZM_CUDA(h, cudaSetDevice(h->device));
  h->stats = zm_stats_t{};
  h->stats.n_voxels = sx * sy * sz;
  if (sx == 0 || sy == 0 || sz == 0) { h->has_result = true; return ZM_OK; }
  if (!labels) return fail(h, ZM_ERR_INVALID, "labels is NULL");

  VolParams vp{};
  if (c_order) { vp.nf = (uint32_t)sz; vp.nm = (uint32_t)sy; vp.ns = (uint32_t)sx; }
  else         { vp.nf = (uint32_t)sx; vp.nm = (uint32_t)sy; vp.ns = (uint32_t)sz; }
  vp.pad = (uint32_t)pad;
  vp.Ef = vp.nf + 2 * vp.pad; vp.Em = vp.nm + 2 * vp.pad; vp.Es = vp.ns + 2 * vp.pad;
  vp.Es_own = vp.Es;
  vp.s_shift = -(int32_t)vp.pad;
  if (slab) {
    // extended planes [cube_lo, cube_hi] of the whole volume; plane p is input plane p - pad
    const uint64_t Eg = slab->full_extent + 2 * pad;
    if (slab->cube_hi <= slab->cube_lo || slab->cube_hi > Eg - 1 || (slab->last != 0) != (slab->cube_hi == Eg - 1))
      return fail(h, ZM_ERR_INVALID, "bad slab cube range");
    const int64_t need_lo = std::max<int64_t>((int64_t)slab->cube_lo - (int64_t)pad, 0);
    const int64_t need_hi = std::min<int64_t>((int64_t)slab->cube_hi - (int64_t)pad, (int64_t)slab->full_extent - 1);
    if ((int64_t)slab->buf_lo > need_lo || (int64_t)(slab->buf_lo + vp.ns) <= need_hi)
      return fail(h, ZM_ERR_INVALID, "slab buffer does not cover the planes its cubes need");
    vp.Es = (uint32_t)(slab->cube_hi - slab->cube_lo + 1);
    vp.Es_own = slab->last ? vp.Es : vp.Es - 1;
    vp.s_shift = (int32_t)((int64_t)slab->cube_lo - (int64_t)pad - (int64_t)slab->buf_lo);
    if (c_order) vp.ox = (uint32_t)slab->cube_lo; else vp.oz = (uint32_t)slab->cube_lo;
  }
  // no cube without two voxels along every axis (marching_cubes.hpp:226-257 loops are empty)
  if (vp.Ef < 2 || vp.Em < 2 || vp.Es < 2) { h->has_result = true; return ZM_OK; }
  vp.ntf = (vp.Ef + TF - 1) / TF; vp.ntm = (vp.Em + TM - 1) / TM; vp.nts = (vp.Es_own + TS - 1) / TS;
  vp.Efp = vp.ntf * TF;
  if (slab && 4ull * vp.Em * vp.Efp >= (1ull << 32))
    return fail(h, ZM_ERR_UNSUPPORTED, "slab plane too large for 32-bit boundary-slot indices (Em * Efp >= 2^30)");
  const unsigned long long ntiles = (unsigned long long)vp.ntf * vp.ntm * vp.nts;
  if (ntiles > 0x7FFFFFFFull || vp.nts > 65535u || (vp.ntm + TM_GROUP - 1) / TM_GROUP > 65535u)
    return fail(h, ZM_ERR_UNSUPPORTED, "too many tiles for one launch; shard the volume");
  const unsigned long long nvox = (unsigned long long)vp.nf * vp.nm * vp.ns;
  const size_t vol_bytes = (size_t)nvox * label_bytes;

  cudaStream_t st = h->stream;
  ZM_CUDA(h, cudaEventRecord(h->ev[0], st));
  if (mem_kind == ZM_MEM_HOST) {
    ZM_CUDA(h, h->d_vol.ensure(vol_bytes));
    ZM_CUDA(h, cudaMemcpyAsync(h->d_vol.p, labels, vol_bytes, cudaMemcpyHostToDevice, st));
    vp.data = h->d_vol.p;
  } else {
    vp.data = labels;
  }
  ZM_CUDA(h, cudaEventRecord(h->ev[1], st));

  const KernelSet ks = kernel_set(label_bytes, c_order != 0);
  CUtensorMap tmap;
  vp.use_tma = make_tensor_map(&tmap, vp.data, label_bytes, vp.nf, vp.nm, vp.ns, ks.row_pad) ? 1u : 0u;

  const size_t nrows = (size_t)vp.Es * vp.Em * vp.ntf;
  ZM_CUDA(h, h->d_rowinfo.ensure(nrows * RI_WORDS * sizeof(uint32_t)));
  ZM_CUDA(h, h->d_hdr.ensure((size_t)ntiles * 2 * sizeof(TileHdr)));  // a dense tile is emitted as two half tiles
  ZM_CUDA(h, h->d_dense.ensure((size_t)ntiles * 4));
  ZM_CUDA(h, h->d_ctl.ensure(sizeof(Control)));
  ZM_CUDA(h, h->d_partial.ensure(3 * 1024 * 8));
  Control* d_ctl = h->d_ctl.as<Control>();

  unsigned long long capV = (unsigned long long)(h->perm_ratio * (double)nvox) + 65536ull;
  unsigned long long capR = (unsigned long long)(h->rec_ratio * (double)nvox) + 65536ull;
  unsigned long long capL = (unsigned long long)(h->tl_ratio * (double)nvox) + 65536ull;
  if (capV > 0xFFFFFFF0ull) capV = 0xFFFFFFF0ull;
  uint32_t launches = 0;
  int attempt = 0;
  Control ctl{};
  for (;; ++attempt) {
    if (attempt >= 8) return fail(h, ZM_ERR_UNSUPPORTED, "label table / capacity sizing did not converge");
    const uint32_t cap = h->hash_cap;
    ZM_CUDA(h, h->d_keys.ensure((size_t)cap * 8));
    ZM_CUDA(h, h->d_cnt.ensure((size_t)cap * 8));
    ZM_CUDA(h, h->d_offV.ensure((size_t)cap * 8));
    ZM_CUDA(h, h->d_offT.ensure((size_t)cap * 8));
    ZM_CUDA(h, h->d_list.ensure((size_t)cap * 24));
    ZM_CUDA(h, h->d_perm.ensure((size_t)capV * 4));
    ZM_CUDA(h, h->d_vinfo.ensure((size_t)capV * 4));
    ZM_CUDA(h, h->d_rec.ensure((size_t)capR * 8));
    ZM_CUDA(h, h->d_tl.ensure((size_t)capL * sizeof(TLEntry)));
    ZM_CUDA(h, cudaMemsetAsync(h->d_keys.p, 0, (size_t)cap * 8, st));
    ZM_CUDA(h, cudaMemsetAsync(h->d_cnt.p, 0, (size_t)cap * 8, st));
    ZM_CUDA(h, cudaMemsetAsync(h->d_ctl.p, 0, sizeof(Control), st));

    LabelTable ht{h->d_keys.as<u64>(), h->d_cnt.as<u64>(), cap - 1};
    Pass1Args p1{ht, d_ctl, h->d_rowinfo.as<uint32_t>(), h->d_perm.as<uint32_t>(),
                 h->d_vinfo.as<uint32_t>(), h->d_rec.as<u64>(), h->d_tl.as<TLEntry>(), h->d_hdr.as<TileHdr>(),
                 h->d_dense.as<uint32_t>(), capV, capR, capL};
    // launch order = grid order (x fastest): f, row inside a group of TM_GROUP tile rows, then all s layers, then the next group
    const dim3 grid0(vp.ntf * std::min<uint32_t>(TM_GROUP, vp.ntm), vp.nts, (vp.ntm + TM_GROUP - 1) / TM_GROUP);
    ks.classify[0]<<<grid0, NT, ks.smem[0], st>>>(vp, tmap, p1);
    ZM_CUDA(h, cudaGetLastError());
    const uint32_t dense_grid = (uint32_t)std::min<unsigned long long>(ntiles, 148ull);
    ks.classify[1]<<<dense_grid, NT, ks.smem[1], st>>>(vp, tmap, p1);
    ZM_CUDA(h, cudaGetLastError());
    ZM_CUDA(h, cudaEventRecord(h->ev[2], st));
    const uint32_t chunk = std::max<uint32_t>(1024u, cap / 1024u);
    ScanArgs sa{ht, h->d_offV.as<u64>(), h->d_offT.as<u64>(), h->d_list.as<u64>(), h->d_partial.as<u64>(), d_ctl, chunk};
    k_scan_partials<<<cap / chunk, 1024, 0, st>>>(sa);
    ZM_CUDA(h, cudaGetLastError());
    k_scan_apply<<<cap / chunk, 1024, 0, st>>>(sa);
    ZM_CUDA(h, cudaGetLastError());
    launches += 4;
    ZM_CUDA(h, cudaMemcpyAsync(h->h_ctl, d_ctl, sizeof(Control), cudaMemcpyDeviceToHost, st));
    ZM_CUDA(h, cudaEventRecord(h->ev[3], st));
    ZM_CUDA(h, cudaStreamSynchronize(st));
    ctl = *h->h_ctl;
    if (ctl.flags & FLAG_INTERNAL) return fail(h, ZM_ERR_CUDA, "internal: dense-mode tile overflowed");
    if (ctl.flags & FLAG_HASH_FULL) {
      if (h->hash_cap >= (1u << 30)) return fail(h, ZM_ERR_UNSUPPORTED, "more than 2^29 distinct labels");
      h->hash_cap <<= 3;
      continue;
    }
    // keep the table at most half full so probes stay short
    if (ctl.totals[0] * 2 > cap) {
      while ((unsigned long long)h->hash_cap < ctl.totals[0] * 4 && h->hash_cap < (1u << 30)) h->hash_cap <<= 1;
      continue;
    }
    if (ctl.cur_perm > 0xFFFFFFF0ull || ctl.cur_tl > 0xFFFFFFF0ull)
      return fail(h, ZM_ERR_UNSUPPORTED, "more than 2^32-16 vertices in one call; shard the volume");
    if (ctl.flags & FLAG_CAP) {
      capV = std::max(capV, ctl.cur_perm + 4096ull);
      capR = std::max(capR, ctl.cur_rec + 4096ull);
      capL = std::max(capL, ctl.cur_tl + 4096ull);
      continue;
    }
    break;
  }
  h->stats.attempts = (uint32_t)attempt + 1;
  const uint32_t cap = h->hash_cap;
  const unsigned long long nlabels = ctl.totals[0], Vtot = ctl.totals[1], Ttot = ctl.totals[2];
  if (Vtot != ctl.cur_perm) return fail(h, ZM_ERR_CUDA, "internal: vertex totals disagree");
  if (Ttot != ctl.cur_tri)
    return fail(h, ZM_ERR_UNSUPPORTED, "a label has more than 2^32-1 faces in one call; shard the volume");

  // The label directory crosses PCIe asynchronously; the host-side structures are built from it on first use
  // (ensure_directory) -- by pass 2 while its kernel runs, so the GPU never waits for the host's hash map and sort.
  if (nlabels) {
    if ((size_t)nlabels * 3 > h->h_list_cap) {
      if (h->h_list) cudaFreeHost(h->h_list);
      h->h_list = nullptr;
      h->h_list_cap = 0;
      const size_t want = (size_t)nlabels * 3 + 3072;
      ZM_CUDA(h, cudaHostAlloc((void**)&h->h_list, want * 8, cudaHostAllocDefault));
      h->h_list_cap = want;
    }
    ZM_CUDA(h, cudaMemcpyAsync(h->h_list, h->d_list.p, (size_t)nlabels * 24, cudaMemcpyDeviceToHost, st));
  }
  ZM_CUDA(h, cudaEventRecord(h->ev[4], st));
  ZM_CUDA(h, cudaEventRecord(h->ev_list, st));
  h->nlabels = nlabels;
  h->dir_pending = true;
  h->Vtot = Vtot;
  h->Ttot = Ttot;
  h->vp = vp;
  h->c_order = c_order != 0;
  h->n_work = ctl.work_count;
  h->capL = capL;

  h->has_result = true;
  h->stats.n_labels = nlabels;  // (labels with vertices; ensure_directory replaces it by the number of ids = labels with faces)
  h->stats.n_vertices = Vtot;
  h->stats.n_faces = Ttot;
  h->stats.n_records = ctl.cur_rec;
  h->stats.n_active_tiles = ctl.work_count;
  h->stats.n_dense_tiles = ctl.dense_count;
  h->stats.n_tiles = ntiles;
  h->stats.hash_capacity = cap;
  h->stats.perm_capacity = capV;
  h->stats.used_tma = vp.use_tma;
  h->stats.launches = launches;
  cudaEventElapsedTime(&h->stats.ms_h2d, h->ev[0], h->ev[1]);
  cudaEventElapsedTime(&h->stats.ms_classify, h->ev[1], h->ev[2]);
  cudaEventElapsedTime(&h->stats.ms_scan, h->ev[2], h->ev[3]);
  cudaEventElapsedTime(&h->stats.ms_total, h->ev[0], h->ev[3]);
  // let the capacity guesses track the data (next call of a similar volume needs one attempt)
  h->perm_ratio = std::max(0.02, std::min(6.5, 1.25 * (double)ctl.cur_perm / (double)nvox));
  h->rec_ratio = std::max(0.02, std::min(8.5, 1.25 * (double)ctl.cur_rec / (double)nvox));
  h->tl_ratio = std::max(0.002, std::min(2.0, 1.25 * (double)ctl.cur_tl / (double)nvox));
  return ZM_OK;
}

// host-side label directory (storage order = table order; offsets are running sums), built on first use
int ensure_directory(zm_handle* h) {
  if (!h->dir_pending) return ZM_OK;
  ZM_CUDA(h, cudaEventSynchronize(h->ev_list));
  h->dir_pending = false;
  const size_t nlabels = (size_t)h->nlabels;
  h->recs.resize(nlabels);
  h->index.reserve(nlabels * 2);
  uint64_t vo = 0, fo = 0;
  for (size_t i = 0; i < nlabels; ++i) {
    LabelRec& r = h->recs[i];
    r.label = h->h_list[3 * i];
    r.nv = h->h_list[3 * i + 1];
    r.nt = h->h_list[3 * i + 2];
    r.voff = vo;
    r.foff = fo;
    r.erased = false;
    vo += r.nv;
    fo += r.nt;
    h->index.emplace(r.label, (uint32_t)i);
  }
  if (vo != h->Vtot || fo != h->Ttot) {
    h->failed = true;
    return fail(h, ZM_ERR_CUDA, "internal: label directory does not add up");
  }
  h->sorted_ids.reserve(nlabels);
  for (const LabelRec& r : h->recs)
    if (r.nt) h->sorted_ids.push_back(r.label);
  std::sort(h->sorted_ids.begin(), h->sorted_ids.end());
  h->stats.n_labels = h->sorted_ids.size();
  return ZM_OK;
}

int ensure_tl_fixed(zm_handle* h) {
  if (h->tl_fixed || !h->n_work) return ZM_OK;
  k_tl_fixup<<<148 * 4, 256, 0, h->stream>>>(h->d_tl.as<TLEntry>(), h->d_ctl.as<Control>(), h->capL, h->d_offV.as<u64>(),
                                             h->d_offT.as<u64>(), h->have_voff ? h->d_voff.as<uint32_t>() : nullptr);
  ZM_CUDA(h, cudaGetLastError());
  h->tl_fixed = true;
  return ZM_OK;
}

Pass2Args pass2_args(zm_handle* h) {
  Pass2Args a{};
  a.hdr = h->d_hdr.as<TileHdr>();
  a.perm = h->d_perm.as<uint32_t>();
  a.vinfo = h->d_vinfo.as<uint32_t>();
  a.rec = h->d_rec.as<u64>();
  a.n_work = h->n_work;
  a.tl = h->d_tl.as<TLEntry>();
  a.foreign = h->foreign;
  a.fnormals = h->nplane_out;
  return a;
}

// Pass 2 (lazy): faces once per zm_mesh; vertices per (voxel_centered, transpose, offset); normals per
// transpose.  Everything is written in its final layout on the device.
// begin_only (slab shards that are not the last): launch pass 2 for every tile but the top layer -- the only ones
// whose cubes reference the next shard's boundary plane -- and return without synchronising; the finalize that
// follows (same arguments, after zm_set_foreign_plane) emits the top layer.
int do_finalize(zm_handle* h, int normals, int voxel_centered, int transpose, const float* off, bool begin_only = false) {
  if (h->failed) return fail(h, ZM_ERR_STATE, "the last zm_mesh failed; mesh again before reading results");
  float o[3] = {h->res[0], h->res[1], h->res[2]};
  if (off) { o[0] = off[0]; o[1] = off[1]; o[2] = off[2]; }
  normals = normals ? 1 : 0; voxel_centered = voxel_centered ? 1 : 0; transpose = transpose ? 1 : 0;
  FinalState& f = h->fin;
  const bool same_verts = f.verts_valid && f.voxel_centered == voxel_centered && f.transpose == transpose &&
                          (!voxel_centered || (f.off[0] == o[0] && f.off[1] == o[1] && f.off[2] == o[2]));
  const bool need_normals = normals && !(f.normals_valid && f.normals_transpose == transpose);
  const bool need_faces = !f.faces_valid;
  if (same_verts && !need_normals && !need_faces) return ZM_OK;
  const bool has_foreign = h->vp.Es_own < h->vp.Es;  // (a slab shard that is not the last)
  if (begin_only && (!has_foreign || !h->Vtot || !h->n_work || f.partial)) return ZM_OK;  // nothing to split
  if (f.partial && (f.partial_args[0] != normals || f.partial_args[1] != voxel_centered || f.partial_args[2] != transpose))
    return fail(h, ZM_ERR_STATE, "slab shard: zm_finalize must repeat the arguments of zm_finalize_begin");
  ZM_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  uint32_t launches = 0;
  if (!f.partial) {
    ZM_CUDA(h, cudaEventRecord(h->ev[5], st));
    ZM_CUDA(h, cudaEventRecord(h->ev[6], st));
  }
  if (h->Vtot && h->n_work) {
    if (has_foreign && !h->foreign && !begin_only)
      return fail(h, ZM_ERR_STATE, "slab shard: zm_set_foreign_plane must be called before the first get/finalize");
    if (normals && h->slab_mode && f.normals_pending)
      return fail(h, ZM_ERR_STATE, "slab shard: normals await zm_add_normal_plane / zm_finish_normals");
    if (need_normals && h->vp.Es_own < h->vp.Es && !h->nplane_out)
      return fail(h, ZM_ERR_STATE, "slab shard: zm_set_normal_plane must be called before a finalize with normals");
    int rc = ensure_tl_fixed(h);
    if (rc != ZM_OK) return rc;
    ZM_CUDA(h, h->d_faces.ensure((size_t)h->Ttot * 12));
    ZM_CUDA(h, h->d_verts.ensure((size_t)h->Vtot * 12));
    if (need_normals && !f.partial) {
      ZM_CUDA(h, h->d_normals.ensure((size_t)h->Vtot * 12));
      ZM_CUDA(h, h->d_nacc.ensure((size_t)h->Vtot * 16));
      ZM_CUDA(h, cudaMemsetAsync(h->d_nacc.p, 0, (size_t)h->Vtot * 16, st));
      if (h->slab_mode && h->nplane_out)
        ZM_CUDA(h, cudaMemsetAsync(h->nplane_out, 0, (size_t)zm_plane_elems(h) * 12, st));
    }
    Pass2Args a = pass2_args(h);
    a.faces = h->d_faces.as<uint32_t>();
    a.verts = h->d_verts.as<float>();
    a.normals = h->d_normals.as<float>();
    a.nacc = h->d_nacc.as<float4>();
    a.r0 = h->res[0]; a.r1 = h->res[1]; a.r2 = h->res[2];
    a.c0 = o[0]; a.c1 = o[1]; a.c2 = o[2];
    a.voxel_centered = voxel_centered;
    a.transpose = transpose;
    a.write_faces = need_faces ? 1 : 0;
    a.write_verts = same_verts ? 0 : 1;
    a.layer_mode = begin_only ? 1u : (f.partial ? 2u : 0u);
    a.top_tile_lo = h->vp.ntf * h->vp.ntm * (h->vp.nts - 1u);
    {
      CUtensorMap rmap;
      if (!make_rowinfo_map(&rmap, h->d_rowinfo.p, h->vp)) return fail(h, ZM_ERR_CUDA, "cuTensorMapEncodeTiled(rowinfo) failed");
      // slab shards: the launch that excludes the top tile layer never looks at the boundary plane
      const int slab_kind = !h->slab_mode ? 0 : ((begin_only || !has_foreign) ? 1 : 2);
      pass2_fn fn = emit_kernel(h->c_order, need_normals, slab_kind);
      int per_sm = 0;
      ZM_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*)fn, EMIT_THREADS, 0));
      const uint32_t grid = std::min<uint32_t>(h->n_work, (uint32_t)(h->num_sms * std::max(per_sm, 1)));
      fn<<<grid, EMIT_THREADS, 0, st>>>(h->vp, rmap, a);
      ZM_CUDA(h, cudaGetLastError());
      ++launches;
    }
    if (begin_only) {
      f.partial = true;
      f.partial_args[0] = normals; f.partial_args[1] = voxel_centered; f.partial_args[2] = transpose;
      h->stats.launches_finalize = launches;
      return ZM_OK;
    }
    ZM_CUDA(h, cudaEventRecord(h->ev[6], st));
    if (need_normals && !h->slab_mode) {  // (slab shards normalise in zm_finish_normals, after the plane exchange)
      k_normals_normalize4<<<grid_for(h->Vtot, 256), 256, 0, st>>>(h->d_nacc.as<float4>(), h->d_normals.as<float>(), h->Vtot);
      ZM_CUDA(h, cudaGetLastError());
      ++launches;
    }
  }
  ZM_CUDA(h, cudaEventRecord(h->ev[7], st));
  {  // (host work overlapped with pass 2)
    const int rc = ensure_directory(h);
    if (rc != ZM_OK) return rc;
  }
  if (h->dir_exchange)  // (device-side directory exchange: its overflow flag travels with this synchronisation)
    ZM_CUDA(h, cudaMemcpyAsync(&h->h_ctl->flags, &h->d_ctl.as<Control>()->flags, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  ZM_CUDA(h, cudaStreamSynchronize(st));
  if (h->dir_exchange && (h->h_ctl->flags & FLAG_DIR))
    return fail(h, ZM_ERR_STATE, "slab shard: a label directory did not fit the exchange buffer; repeat the step with a larger capacity");
  cudaEventElapsedTime(&h->stats.ms_faces, h->ev[5], h->ev[6]);
  cudaEventElapsedTime(&h->stats.ms_vertices, h->ev[6], h->ev[7]);
  cudaEventElapsedTime(&h->stats.ms_finalize, h->ev[5], h->ev[7]);
  h->stats.launches_finalize = launches + (f.partial ? h->stats.launches_finalize : 0u);
  f.partial = false;
  f.faces_valid = true;
  f.verts_valid = true;
  f.voxel_centered = voxel_centered;
  f.transpose = transpose;
  f.off[0] = o[0]; f.off[1] = o[1]; f.off[2] = o[2];
  if (need_normals) {
    f.normals_transpose = transpose;
    if (h->slab_mode) f.normals_pending = true;
    else f.normals_valid = true;
  }
  if (normals && h->slab_mode && !f.normals_valid && !f.normals_pending) {  // (nothing to accumulate on this shard)
    f.normals_pending = true;
    f.normals_transpose = transpose;
  }
  return ZM_OK;
}

// ---- NCCL, bound at run time (dlopen): inside a torch process this resolves to the library torch already loaded ----
struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};

const NcclApi& nccl_api() {
  static NcclApi api = []() {
    NcclApi a;
    void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW);
    if (!lib) return a;
    auto sym = [&](const char* n) { return dlsym(lib, n); };
    a.GetUniqueId = (decltype(a.GetUniqueId))sym("ncclGetUniqueId");
    a.CommInitRank = (decltype(a.CommInitRank))sym("ncclCommInitRank");
    a.CommDestroy = (decltype(a.CommDestroy))sym("ncclCommDestroy");
    a.AllGather = (decltype(a.AllGather))sym("ncclAllGather");
    a.Send = (decltype(a.Send))sym("ncclSend");
    a.Recv = (decltype(a.Recv))sym("ncclRecv");
    a.GroupStart = (decltype(a.GroupStart))sym("ncclGroupStart");
    a.GroupEnd = (decltype(a.GroupEnd))sym("ncclGroupEnd");
    a.GetErrorString = (decltype(a.GetErrorString))sym("ncclGetErrorString");
    a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllGather && a.Send && a.Recv && a.GroupStart &&
           a.GroupEnd && a.GetErrorString;
    return a;
  }();
  return api;
}

#define ZM_NCCL(h, call)                                                                                   \
  do {                                                                                                     \
    ncclResult_t _r = (call);                                                                              \
    if (_r != ncclSuccess) {                                                                               \
      char _b[512];                                                                                        \
      snprintf(_b, sizeof(_b), "%s failed: %s (%s:%d)", #call, nccl_api().GetErrorString(_r), __FILE__, __LINE__); \
      (h)->err = _b;                                                                                       \
      return ZM_ERR_CUDA;                                                                                  \
    }                                                                                                      \
  } while (0)

}  // namespace

extern "C" {

const char* zm_version(void) { return "zmesh_b200 0.1 (sm_100a)"; }

const char* zm_last_error(zm_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int zm_create(const float resolution[3], int device, zm_handle** out) {
  if (!out || !resolution) { g_create_error = "null argument"; return ZM_ERR_INVALID; }
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    g_create_error = std::string("no usable CUDA device: ") + cudaGetErrorString(e) +
                     " (zmesh_b200 has no CPU fallback)";
    cudaGetLastError();
    return ZM_ERR_CUDA;
  }
  if (device < 0) {
    if (cudaGetDevice(&device) != cudaSuccess) device = 0;
  }
  if (device >= ndev) { g_create_error = "device index out of range"; return ZM_ERR_INVALID; }
  zm_handle* h = new zm_handle();
  h->device = device;
  for (int i = 0; i < 3; ++i) h->res[i] = resolution[i];
  auto bail = [&](const char* what, cudaError_t err) {
    g_create_error = std::string(what) + ": " + cudaGetErrorString(err);
    cudaGetLastError();
    delete h;
    return ZM_ERR_CUDA;
  };
  if ((e = cudaSetDevice(device)) != cudaSuccess) return bail("cudaSetDevice", e);
  if ((e = cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
  if (cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || h->num_sms <= 0) h->num_sms = 148;
  h->stream = h->own_stream;
  for (auto& ev : h->ev)
    if ((e = cudaEventCreate(&ev)) != cudaSuccess) return bail("cudaEventCreate", e);
  if ((e = cudaHostAlloc((void**)&h->h_ctl, sizeof(zm::Control), cudaHostAllocDefault)) != cudaSuccess) return bail("cudaHostAlloc", e);
  if ((e = cudaEventCreateWithFlags(&h->ev_list, cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
  int rc = prepare_device(h);
  if (rc != ZM_OK) {
    g_create_error = h->err;
    zm_destroy(h);
    return rc;
  }
  *out = h;
  return ZM_OK;
}

void zm_destroy(zm_handle* h) {
  if (!h) return;
  zm_comm_destroy(h);
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  for (DevBuf* b : {&h->d_vol, &h->d_keys, &h->d_cnt, &h->d_offV, &h->d_offT, &h->d_list, &h->d_partial, &h->d_ctl,
                    &h->d_rowinfo, &h->d_perm, &h->d_vinfo, &h->d_rec, &h->d_tl, &h->d_hdr,
                    &h->d_dense, &h->d_faces, &h->d_verts, &h->d_normals, &h->d_nacc, &h->d_pack, &h->d_packtab, &h->d_voff, &h->d_tmpA, &h->d_tmpB})
    b->release();
  if (h->h_ctl) cudaFreeHost(h->h_ctl);
  if (h->h_list) cudaFreeHost(h->h_list);
  if (h->ev_list) cudaEventDestroy(h->ev_list);
  for (auto& ev : h->ev)
    if (ev) cudaEventDestroy(ev);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  delete h;
}

int zm_set_stream(zm_handle* h, void* cuda_stream) {
  if (!h) return ZM_ERR_INVALID;
  ZM_CUDA(h, cudaSetDevice(h->device));
  ZM_CUDA(h, cudaStreamSynchronize(h->stream));
  h->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : h->own_stream;
  return ZM_OK;
}

int zm_wait_stream(zm_handle* h, void* producer_stream) {
  if (!h) return ZM_ERR_INVALID;
  ZM_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t ps = producer_stream ? static_cast<cudaStream_t>(producer_stream) : cudaStreamLegacy;
  if (ps == h->stream) return ZM_OK;
  // ev[0] is re-recorded by the next zm_mesh; a stream wait keeps the state it saw when it was queued
  ZM_CUDA(h, cudaEventRecord(h->ev[0], ps));
  ZM_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev[0], 0));
  return ZM_OK;
}

int zm_synth_voronoi(void* dst, int label_bytes, const uint64_t shape[3], const uint64_t origin[3],
                     const uint64_t full_shape[3], uint32_t pitch, uint64_t seed, int c_order, void* cuda_stream) {
  if (!dst || !shape || !origin || !full_shape || pitch == 0) return ZM_ERR_INVALID;
  if (label_bytes != 1 && label_bytes != 2 && label_bytes != 4 && label_bytes != 8) return ZM_ERR_INVALID;
  SynthArgs a{};
  a.dst = dst;
  a.sx = (uint32_t)shape[0]; a.sy = (uint32_t)shape[1]; a.sz = (uint32_t)shape[2];
  a.n = (unsigned long long)shape[0] * shape[1] * shape[2];
  a.ox = (uint32_t)origin[0]; a.oy = (uint32_t)origin[1]; a.oz = (uint32_t)origin[2];
  auto cells = [&](uint64_t s) { uint64_t g = (s + pitch - 1) / pitch; return (uint32_t)(g ? g : 1); };
  a.gx = cells(full_shape[0]); a.gy = cells(full_shape[1]); a.gz = cells(full_shape[2]);
  a.pitch = pitch; a.seed = seed; a.c_order = c_order;
  if (a.n == 0) return ZM_OK;
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  const uint32_t grid = grid_for(a.n, 256) * 4;
  switch (label_bytes) {
    case 1: k_synth_voronoi<uint8_t><<<grid, 256, 0, st>>>(a); break;
    case 2: k_synth_voronoi<uint16_t><<<grid, 256, 0, st>>>(a); break;
    case 4: k_synth_voronoi<uint32_t><<<grid, 256, 0, st>>>(a); break;
    default: k_synth_voronoi<unsigned long long><<<grid, 256, 0, st>>>(a); break;
  }
  return cudaGetLastError() == cudaSuccess ? ZM_OK : ZM_ERR_CUDA;
}

int zm_set_resolution(zm_handle* h, const float resolution[3]) {
  if (!h || !resolution) return ZM_ERR_INVALID;
  for (int i = 0; i < 3; ++i) h->res[i] = resolution[i];
  h->fin = FinalState();
  return ZM_OK;
}

int zm_mesh(zm_handle* h, const void* labels, int label_bytes, uint64_t sx, uint64_t sy, uint64_t sz,
            int c_order, int close, int mem_kind) {
  const int rc = run_mesh(h, labels, label_bytes, sx, sy, sz, c_order, close, mem_kind, nullptr);
  if (h) h->failed = rc != ZM_OK;
  return rc;
}

int zm_mesh_slab(zm_handle* h, const void* labels, int label_bytes, uint64_t sx, uint64_t sy, uint64_t sz,
                 int c_order, int close, int mem_kind, const zm_slab* slab) {
  if (!slab) return ZM_ERR_INVALID;
  const int rc = run_mesh(h, labels, label_bytes, sx, sy, sz, c_order, close, mem_kind, slab);
  if (h) h->failed = rc != ZM_OK;
  return rc;
}

int zm_directory(zm_handle* h, uint64_t* labels, uint64_t* n_vertices, uint64_t* n_faces, uint64_t capacity) {
  if (!h) return ZM_ERR_INVALID;
  if (ensure_directory(h) != ZM_OK) return ZM_ERR_CUDA;
  const uint64_t n = std::min<uint64_t>(capacity, h->recs.size());
  for (uint64_t i = 0; i < n; ++i) {
    if (labels) labels[i] = h->recs[i].label;
    if (n_vertices) n_vertices[i] = h->recs[i].nv;
    if (n_faces) n_faces[i] = h->recs[i].nt;
  }
  return ZM_OK;
}

uint64_t zm_num_directory(zm_handle* h) { return h ? h->nlabels : 0; }

int zm_set_label_offsets(zm_handle* h, const uint64_t* labels, const uint32_t* offsets, uint64_t n) {
  if (!h || (n && (!labels || !offsets))) return ZM_ERR_INVALID;
  if (!h->has_result) return fail(h, ZM_ERR_STATE, "zm_mesh has not been called");
  if (h->tl_fixed) return fail(h, ZM_ERR_STATE, "label offsets must be set before the first get/finalize/export");
  if (!h->n_work) return ZM_OK;
  ZM_CUDA(h, cudaSetDevice(h->device));
  const uint32_t cap = h->hash_cap;
  ZM_CUDA(h, h->d_voff.ensure((size_t)cap * 4));
  ZM_CUDA(h, cudaMemsetAsync(h->d_voff.p, 0, (size_t)cap * 4, h->stream));
  if (n) {
    ZM_CUDA(h, h->d_tmpA.ensure(n * 8));
    ZM_CUDA(h, h->d_tmpB.ensure(n * 4));
    ZM_CUDA(h, cudaMemcpyAsync(h->d_tmpA.p, labels, n * 8, cudaMemcpyHostToDevice, h->stream));
    ZM_CUDA(h, cudaMemcpyAsync(h->d_tmpB.p, offsets, n * 4, cudaMemcpyHostToDevice, h->stream));
    LabelTable ht{h->d_keys.as<u64>(), h->d_cnt.as<u64>(), cap - 1};
    k_set_voff<<<grid_for(n, 256), 256, 0, h->stream>>>(ht, h->d_tmpA.as<u64>(), h->d_tmpB.as<uint32_t>(), n,
                                                        h->d_voff.as<uint32_t>());
    ZM_CUDA(h, cudaGetLastError());
    ZM_CUDA(h, cudaStreamSynchronize(h->stream));  // the host arrays are borrowed
  }
  h->have_voff = true;
  return ZM_OK;
}

int zm_export_directory(zm_handle* h, uint64_t* dst_device, uint64_t capacity) {
  if (!h || !dst_device) return ZM_ERR_INVALID;
  if (!h->has_result) return fail(h, ZM_ERR_STATE, "zm_mesh has not been called");
  ZM_CUDA(h, cudaSetDevice(h->device));
  if (!h->n_work && h->nlabels == 0) {  // nothing meshed on this shard (degenerate slab): an empty directory
    ZM_CUDA(h, cudaMemsetAsync(dst_device, 0, 16, h->stream));
    return ZM_OK;
  }
  k_export_directory<<<grid_for(h->nlabels + 1, 256), 256, 0, h->stream>>>(h->d_list.as<u64>(), h->d_ctl.as<Control>(),
                                                                               (u64*)dst_device, capacity);
  ZM_CUDA(h, cudaGetLastError());
  return ZM_OK;
}

int zm_import_directories(zm_handle* h, const uint64_t* all_device, uint32_t world, uint32_t rank, uint64_t capacity) {
  if (!h || !all_device || rank >= world) return ZM_ERR_INVALID;
  if (!h->has_result) return fail(h, ZM_ERR_STATE, "zm_mesh has not been called");
  if (h->tl_fixed) return fail(h, ZM_ERR_STATE, "label offsets must be set before the first get/finalize/export");
  ZM_CUDA(h, cudaSetDevice(h->device));
  // a shard without output still takes part in the overflow check, so that all shards take the same decision
  ZM_CUDA(h, h->d_ctl.ensure(sizeof(Control)));
  const bool have_table = h->n_work != 0 && h->d_keys.p && h->d_cnt.p;
  const uint32_t cap = have_table ? h->hash_cap : 1u;
  ZM_CUDA(h, h->d_voff.ensure((size_t)cap * 4));
  ZM_CUDA(h, cudaMemsetAsync(h->d_voff.p, 0, (size_t)cap * 4, h->stream));
  if (!have_table) ZM_CUDA(h, cudaMemsetAsync(h->d_ctl.p, 0, sizeof(Control), h->stream));
  LabelTable ht{h->d_keys.as<u64>(), h->d_cnt.as<u64>(), cap - 1};
  const dim3 grid(std::min<uint32_t>(grid_for(std::min<uint64_t>(capacity, 1u << 20), 256), 64u), world);
  k_import_directories<<<grid, 256, 0, h->stream>>>(ht, (const u64*)all_device, have_table ? rank : 0u, capacity,
                                                    h->d_voff.as<uint32_t>(), h->d_ctl.as<Control>());
  ZM_CUDA(h, cudaGetLastError());
  h->have_voff = have_table;
  h->dir_exchange = true;
  return ZM_OK;
}

uint64_t zm_plane_elems(zm_handle* h) { return h ? 4ull * h->vp.Em * h->vp.Efp : 0; }

int zm_export_plane(zm_handle* h, uint32_t* dst_device) {
  if (!h || !dst_device) return ZM_ERR_INVALID;
  if (!h->has_result) return fail(h, ZM_ERR_STATE, "zm_mesh has not been called");
  if (!h->n_work) return ZM_OK;
  ZM_CUDA(h, cudaSetDevice(h->device));
  int rc = ensure_tl_fixed(h);
  if (rc != ZM_OK) return rc;
  Pass2Args a = pass2_args(h);
  const uint32_t grid = std::min<uint32_t>((h->n_work + NT_V / 32 - 1) / (NT_V / 32), (uint32_t)h->num_sms * 16u);
  if (h->c_order) k_export_plane<true><<<grid, NT_V, 0, h->stream>>>(h->vp, a, dst_device);
  else k_export_plane<false><<<grid, NT_V, 0, h->stream>>>(h->vp, a, dst_device);
  ZM_CUDA(h, cudaGetLastError());
  return ZM_OK;
}

int zm_set_foreign_plane(zm_handle* h, const uint32_t* src_device) {
  if (!h) return ZM_ERR_INVALID;
  h->foreign = src_device;
  return ZM_OK;
}

int zm_set_normal_plane(zm_handle* h, float* out_device) {
  if (!h) return ZM_ERR_INVALID;
  h->nplane_out = out_device;
  return ZM_OK;
}

int zm_add_normal_plane(zm_handle* h, const float* src_device) {
  if (!h || !src_device) return ZM_ERR_INVALID;
  if (!h->has_result || !h->fin.normals_pending) return fail(h, ZM_ERR_STATE, "zm_finalize with normals has not been called on this slab shard");
  if (!h->n_work || !h->Vtot) return ZM_OK;
  ZM_CUDA(h, cudaSetDevice(h->device));
  Pass2Args a = pass2_args(h);
  a.normals = h->d_normals.as<float>();
  a.nacc = h->d_nacc.as<float4>();
  const uint32_t grid = std::min<uint32_t>((h->n_work + NT_V / 32 - 1) / (NT_V / 32), (uint32_t)h->num_sms * 16u);
  if (h->c_order) k_import_plane_normals<true><<<grid, NT_V, 0, h->stream>>>(h->vp, a, src_device);
  else k_import_plane_normals<false><<<grid, NT_V, 0, h->stream>>>(h->vp, a, src_device);
  ZM_CUDA(h, cudaGetLastError());
  return ZM_OK;
}

int zm_finish_normals(zm_handle* h) {
  if (!h) return ZM_ERR_INVALID;
  if (!h->has_result || !h->fin.normals_pending) return fail(h, ZM_ERR_STATE, "no pending normals on this handle");
  ZM_CUDA(h, cudaSetDevice(h->device));
  if (h->Vtot && h->n_work) {
    k_normals_normalize4<<<grid_for(h->Vtot, 256), 256, 0, h->stream>>>(h->d_nacc.as<float4>(), h->d_normals.as<float>(), h->Vtot);
    ZM_CUDA(h, cudaGetLastError());
    ZM_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  h->fin.normals_pending = false;
  h->fin.normals_valid = true;
  return ZM_OK;
}

uint64_t zm_num_ids(zm_handle* h) {
  if (!h || ensure_directory(h) != ZM_OK) return 0;
  return h->sorted_ids.size();
}

int zm_ids(zm_handle* h, uint64_t* out, uint64_t capacity) {
  if (!h || (!out && capacity)) return ZM_ERR_INVALID;
  if (ensure_directory(h) != ZM_OK) return ZM_ERR_CUDA;
  uint64_t n = std::min<uint64_t>(capacity, h->sorted_ids.size());
  if (n) memcpy(out, h->sorted_ids.data(), n * sizeof(uint64_t));
  return ZM_OK;
}

int zm_get_counts(zm_handle* h, uint64_t label, uint64_t* nv, uint64_t* nf) {
  if (!h || !nv || !nf) return ZM_ERR_INVALID;
  *nv = *nf = 0;
  // (before the first zm_mesh the handle answers like an empty volume: the reference's Mesher.__init__ builds an
  // empty Mesher6464, zmesh/_zmesh.pyx:442-444)
  if (h->failed) return fail(h, ZM_ERR_STATE, "the last zm_mesh failed; mesh again before reading results");
  if (ensure_directory(h) != ZM_OK) return ZM_ERR_CUDA;
  auto it = h->index.find(label);
  if (it == h->index.end() || h->recs[it->second].erased) return ZM_OK;
  *nv = h->recs[it->second].nv;
  *nf = h->recs[it->second].nt;
  return ZM_OK;
}

int zm_get(zm_handle* h, uint64_t label, int normals, int voxel_centered, int transpose,
           const float centering_offset[3], float* vertices, uint32_t* faces, float* normals_out) {
  if (!h) return ZM_ERR_INVALID;
  if (h->failed) return fail(h, ZM_ERR_STATE, "the last zm_mesh failed; mesh again before reading results");
  if (ensure_directory(h) != ZM_OK) return ZM_ERR_CUDA;
  auto it = h->index.find(label);
  if (it == h->index.end() || h->recs[it->second].erased) return ZM_OK;  // empty mesh
  const LabelRec& r = h->recs[it->second];
  if (!vertices || !faces || (normals && !normals_out)) return fail(h, ZM_ERR_INVALID, "null output buffer");
  int rc = do_finalize(h, normals, voxel_centered, transpose, centering_offset);
  if (rc != ZM_OK) return rc;
  cudaStream_t st = h->stream;
  ZM_CUDA(h, cudaMemcpyAsync(vertices, h->d_verts.as<float>() + 3 * r.voff, (size_t)r.nv * 12, cudaMemcpyDeviceToHost, st));
  ZM_CUDA(h, cudaMemcpyAsync(faces, h->d_faces.as<uint32_t>() + 3 * r.foff, (size_t)r.nt * 12, cudaMemcpyDeviceToHost, st));
  if (normals)
    ZM_CUDA(h, cudaMemcpyAsync(normals_out, h->d_normals.as<float>() + 3 * r.voff, (size_t)r.nv * 12, cudaMemcpyDeviceToHost, st));
  ZM_CUDA(h, cudaStreamSynchronize(st));
  if (transpose) {  // legacy winding (t0,t2,t1) = stored (t1,t2,t0) reversed (cMesher.hpp:152-157)
    for (uint64_t i = 0; i < r.nt; ++i) std::swap(faces[3 * i], faces[3 * i + 2]);
  }
  return ZM_OK;
}

int zm_erase(zm_handle* h, uint64_t label, int* existed) {
  if (!h) return ZM_ERR_INVALID;
  int ex = 0;
  if (ensure_directory(h) != ZM_OK) return ZM_ERR_CUDA;
  auto it = h->index.find(label);
  if (it != h->index.end() && !h->recs[it->second].erased && h->recs[it->second].nt) {
    h->recs[it->second].erased = true;
    auto pos = std::lower_bound(h->sorted_ids.begin(), h->sorted_ids.end(), label);
    if (pos != h->sorted_ids.end() && *pos == label) h->sorted_ids.erase(pos);
    ex = 1;
  }
  if (existed) *existed = ex;
  return ZM_OK;
}

int zm_clear(zm_handle* h) {
  if (!h) return ZM_ERR_INVALID;
  const bool had = h->has_result;
  drop_results(h);
  h->has_result = had;  // a cleared mesher answers like an empty one (marching_cubes.hpp:184-189)
  cudaSetDevice(h->device);
  for (DevBuf* b : {&h->d_faces, &h->d_verts, &h->d_normals, &h->d_nacc, &h->d_perm, &h->d_vinfo, &h->d_rec, &h->d_tl, &h->d_rowinfo,
                    &h->d_vol})
    b->release();
  return ZM_OK;
}

int zm_finalize(zm_handle* h, int normals, int voxel_centered, int transpose, const float centering_offset[3],
                zm_bulk_view* view) {
  if (!h) return ZM_ERR_INVALID;
  int rc = do_finalize(h, normals, voxel_centered, transpose, centering_offset);
  if (rc != ZM_OK) return rc;
  if (view) {
    rc = ensure_directory(h);
    if (rc != ZM_OK) return rc;
    if (h->bulk_labels.size() != h->recs.size() || h->bulk_voff.empty()) {
      h->bulk_labels.clear(); h->bulk_voff.clear(); h->bulk_foff.clear();
      for (const LabelRec& r : h->recs) {
        h->bulk_labels.push_back(r.label);
        h->bulk_voff.push_back(r.voff);
        h->bulk_foff.push_back(r.foff);
      }
      h->bulk_voff.push_back(h->Vtot);
      h->bulk_foff.push_back(h->Ttot);
    }
    view->n_labels = h->recs.size();
    view->n_vertices = h->Vtot;
    view->n_faces = h->Ttot;
    view->labels_host = h->bulk_labels.data();
    view->voff_host = h->bulk_voff.data();
    view->foff_host = h->bulk_foff.data();
    view->vertices_dev = h->Vtot ? h->d_verts.as<float>() : nullptr;
    view->faces_dev = h->Ttot ? h->d_faces.as<uint32_t>() : nullptr;
    view->normals_dev = (normals && h->fin.normals_valid && h->Vtot) ? h->d_normals.as<float>() : nullptr;
  }
  return ZM_OK;
}

int zm_finalize_begin(zm_handle* h, int normals, int voxel_centered, int transpose, const float centering_offset[3]) {
  if (!h) return ZM_ERR_INVALID;
  return do_finalize(h, normals, voxel_centered, transpose, centering_offset, true);
}

int zm_fetch_all(zm_handle* h, float* vertices, uint32_t* faces, float* normals_out) {
  if (!h) return ZM_ERR_INVALID;
  if (!h->has_result || !h->fin.verts_valid) return fail(h, ZM_ERR_STATE, "zm_finalize has not been called");
  if (normals_out && !(h->fin.normals_valid && h->fin.normals_transpose == h->fin.transpose))
    return fail(h, ZM_ERR_STATE, "normals were not requested in zm_finalize");
  cudaStream_t st = h->stream;
  if (h->Vtot && vertices)
    ZM_CUDA(h, cudaMemcpyAsync(vertices, h->d_verts.p, (size_t)h->Vtot * 12, cudaMemcpyDeviceToHost, st));
  if (h->Ttot && faces)
    ZM_CUDA(h, cudaMemcpyAsync(faces, h->d_faces.p, (size_t)h->Ttot * 12, cudaMemcpyDeviceToHost, st));
  if (h->Vtot && normals_out)
    ZM_CUDA(h, cudaMemcpyAsync(normals_out, h->d_normals.p, (size_t)h->Vtot * 12, cudaMemcpyDeviceToHost, st));
  ZM_CUDA(h, cudaStreamSynchronize(st));
  if (h->fin.transpose && faces)
    for (uint64_t i = 0; i < h->Ttot; ++i) std::swap(faces[3 * i], faces[3 * i + 2]);
  return ZM_OK;
}

int zm_pack_precomputed(zm_handle* h, int voxel_centered, const float centering_offset[3], uint64_t* n_objects,
                        uint64_t* total_bytes) {
  if (!h || !n_objects || !total_bytes) return ZM_ERR_INVALID;
  *n_objects = *total_bytes = 0;
  int rc = do_finalize(h, 0, voxel_centered, 0, centering_offset);
  if (rc != ZM_OK) return rc;
  rc = ensure_directory(h);
  if (rc != ZM_OK) return rc;
  h->pack_labels.clear();
  h->pack_off.clear();
  const size_t nl = h->sorted_ids.size();
  if (nl == 0) return ZM_OK;
  // table: word[nl + 1], voff[nl], nv[nl], foff[nl]  (ids in ascending order; erased labels are not in sorted_ids)
  std::vector<uint64_t> tab(4 * nl + 1);
  uint64_t w = 0;
  for (size_t i = 0; i < nl; ++i) {
    const LabelRec& r = h->recs[h->index[h->sorted_ids[i]]];
    tab[i] = w;
    tab[nl + 1 + i] = r.voff;
    tab[2 * nl + 1 + i] = r.nv;
    tab[3 * nl + 1 + i] = r.foff;
    h->pack_labels.push_back(r.label);
    h->pack_off.push_back(4 * w);
    w += 1 + 3 * r.nv + 3 * r.nt;
  }
  tab[nl] = w;
  h->pack_off.push_back(4 * w);
  ZM_CUDA(h, cudaSetDevice(h->device));
  ZM_CUDA(h, h->d_packtab.ensure(tab.size() * 8));
  ZM_CUDA(h, h->d_pack.ensure((size_t)w * 4));
  cudaStream_t st = h->stream;
  ZM_CUDA(h, cudaMemcpyAsync(h->d_packtab.p, tab.data(), tab.size() * 8, cudaMemcpyHostToDevice, st));
  PackArgs a{};
  const u64* t = h->d_packtab.as<u64>();
  a.word = t; a.voff = t + nl + 1; a.nv = t + 2 * nl + 1; a.foff = t + 3 * nl + 1;
  a.verts = h->d_verts.as<uint32_t>();
  a.faces = h->d_faces.as<uint32_t>();
  a.out = h->d_pack.as<uint32_t>();
  a.total = w;
  a.nl = (uint32_t)nl;
  const uint64_t nchunks = (w + PACK_CHUNK - 1) / PACK_CHUNK;
  k_pack_precomputed<<<(uint32_t)std::min<uint64_t>(nchunks, (uint64_t)h->num_sms * 16), 256, 0, st>>>(a);
  ZM_CUDA(h, cudaGetLastError());
  ZM_CUDA(h, cudaStreamSynchronize(st));  // (the table is a host vector)
  *n_objects = nl;
  *total_bytes = 4 * w;
  return ZM_OK;
}

int zm_fetch_precomputed(zm_handle* h, void* dst_host, uint64_t* labels_out, uint64_t* byte_offsets_out) {
  if (!h) return ZM_ERR_INVALID;
  const size_t nl = h->pack_labels.size();
  if (nl == 0) return ZM_OK;
  if (!dst_host || !labels_out || !byte_offsets_out) return fail(h, ZM_ERR_INVALID, "null output buffer");
  ZM_CUDA(h, cudaSetDevice(h->device));
  ZM_CUDA(h, cudaMemcpyAsync(dst_host, h->d_pack.p, (size_t)h->pack_off[nl], cudaMemcpyDeviceToHost, h->stream));
  memcpy(labels_out, h->pack_labels.data(), nl * 8);
  memcpy(byte_offsets_out, h->pack_off.data(), (nl + 1) * 8);
  ZM_CUDA(h, cudaStreamSynchronize(h->stream));
  return ZM_OK;
}

int zm_compute_normals(zm_handle* h, const float* vertices, uint64_t n_vertices, const uint32_t* faces,
                       uint64_t n_faces, float* normals_out) {
  if (!h) return ZM_ERR_INVALID;
  if (n_vertices == 0) return ZM_OK;
  if (!vertices || !normals_out || (n_faces && !faces)) return fail(h, ZM_ERR_INVALID, "null buffer");
  ZM_CUDA(h, cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  DevBuf dv, df, dn;
  auto cleanup = [&]() { dv.release(); df.release(); dn.release(); };
  cudaError_t e;
  if ((e = dv.ensure(n_vertices * 12)) != cudaSuccess || (e = df.ensure(n_faces * 12 + 16)) != cudaSuccess ||
      (e = dn.ensure(n_vertices * 12)) != cudaSuccess) {
    cleanup();
    return fail(h, ZM_ERR_OOM, std::string("device allocation failed: ") + cudaGetErrorString(e));
  }
  int rc = ZM_OK;
  do {
    if ((e = cudaMemcpyAsync(dv.p, vertices, n_vertices * 12, cudaMemcpyHostToDevice, st)) != cudaSuccess) break;
    if (n_faces && (e = cudaMemcpyAsync(df.p, faces, n_faces * 12, cudaMemcpyHostToDevice, st)) != cudaSuccess) break;
    if ((e = cudaMemsetAsync(dn.p, 0, n_vertices * 12, st)) != cudaSuccess) break;
    if (n_faces) {
      k_normals_accumulate_f32<<<grid_for(n_faces, 256), 256, 0, st>>>(dv.as<float>(), df.as<uint32_t>(), n_faces,
                                                                       dn.as<float>());
      if ((e = cudaGetLastError()) != cudaSuccess) break;
    }
    k_normals_normalize<<<grid_for(n_vertices, 256), 256, 0, st>>>(dn.as<float>(), n_vertices);
    if ((e = cudaGetLastError()) != cudaSuccess) break;
    if ((e = cudaMemcpyAsync(normals_out, dn.p, n_vertices * 12, cudaMemcpyDeviceToHost, st)) != cudaSuccess) break;
    e = cudaStreamSynchronize(st);
  } while (0);
  if (e != cudaSuccess) rc = fail(h, ZM_ERR_CUDA, std::string("zm_compute_normals: ") + cudaGetErrorString(e));
  cleanup();
  return rc;
}

void* zm_host_alloc(uint64_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return p;
}

void zm_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

int zm_stats(zm_handle* h, zm_stats_t* out) {
  if (!h || !out) return ZM_ERR_INVALID;
  *out = h->stats;
  return ZM_OK;
}

int zm_nccl_unique_id(void* out128) {
  if (!out128) return ZM_ERR_INVALID;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  const NcclApi& N = nccl_api();
  if (!N.ok) { g_create_error = "libnccl.so.2 could not be loaded"; return ZM_ERR_UNSUPPORTED; }
  ncclUniqueId id;
  if (N.GetUniqueId(&id) != ncclSuccess) { g_create_error = "ncclGetUniqueId failed"; return ZM_ERR_CUDA; }
  memcpy(out128, &id, 128);
  return ZM_OK;
}

int zm_comm_destroy(zm_handle* h) {
  if (!h) return ZM_ERR_INVALID;
  if (!h->comm && !h->comm_lo && !h->comm_hi && !h->comm_stream) return ZM_OK;  // (never touch NCCL -- not even load it -- without a communicator)
  const NcclApi& N = nccl_api();
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->comm_stream) cudaStreamSynchronize(h->comm_stream);
  for (ncclComm_t* c : {&h->comm, &h->comm_lo, &h->comm_hi}) {
    if (*c && N.ok) N.CommDestroy(*c);
    *c = nullptr;
  }
  for (auto& ev : h->ev_comm) { if (ev) cudaEventDestroy(ev); ev = nullptr; }
  if (h->comm_stream) cudaStreamDestroy(h->comm_stream);
  h->comm_stream = nullptr;
  for (DevBuf* b : {&h->d_dir_mine, &h->d_dir_all, &h->d_plane_send, &h->d_plane_recv, &h->d_nplane_out, &h->d_nplane_in})
    b->release();
  h->world = 1; h->rank = 0;
  return ZM_OK;
}

int zm_comm_init(zm_handle* h, const void* id_collectives, const void* id_pairs, int world, int rank) {
  if (!h || !id_collectives || (world > 1 && !id_pairs) || world < 1 || rank < 0 || rank >= world) return ZM_ERR_INVALID;
  const NcclApi& N = nccl_api();
  if (!N.ok) return fail(h, ZM_ERR_UNSUPPORTED, "libnccl.so.2 could not be loaded");
  zm_comm_destroy(h);
  ZM_CUDA(h, cudaSetDevice(h->device));
  ncclUniqueId a, lo, hi;
  memcpy(&a, id_collectives, 128);
  // id_pairs[k] (128 bytes each, k < world - 1) names the communicator of shards k and k + 1
  if (rank > 0) memcpy(&lo, static_cast<const char*>(id_pairs) + 128 * (size_t)(rank - 1), 128);
  if (rank + 1 < world) memcpy(&hi, static_cast<const char*>(id_pairs) + 128 * (size_t)rank, 128);
  ZM_NCCL(h, N.GroupStart());
  ZM_NCCL(h, N.CommInitRank(&h->comm, world, a, rank));
  if (rank > 0) ZM_NCCL(h, N.CommInitRank(&h->comm_lo, 2, lo, 1));         // (the upper shard of a pair is its rank 1)
  if (rank + 1 < world) ZM_NCCL(h, N.CommInitRank(&h->comm_hi, 2, hi, 0));
  ZM_NCCL(h, N.GroupEnd());
  ZM_CUDA(h, cudaStreamCreateWithFlags(&h->comm_stream, cudaStreamNonBlocking));
  for (auto& ev : h->ev_comm) ZM_CUDA(h, cudaEventCreate(&ev));
  h->world = world;
  h->rank = rank;
  return ZM_OK;
}

int zm_slab_range(uint64_t full_extent, int close, int rank, int world, zm_slab* slab, uint64_t* in_lo, uint64_t* in_hi) {
  if (!slab || world < 1 || rank < 0 || rank >= world) return ZM_ERR_INVALID;
  const uint64_t pad = close ? 1 : 0;
  const uint64_t ncube = full_extent + 2 * pad - 1;  // cube-origin planes of the whole (extended) volume
  if (full_extent == 0 || (uint64_t)world > ncube) return ZM_ERR_INVALID;
  slab->full_extent = full_extent;
  slab->cube_lo = ncube * (uint64_t)rank / (uint64_t)world;
  slab->cube_hi = ncube * (uint64_t)(rank + 1) / (uint64_t)world;
  slab->last = rank == world - 1;
  const uint64_t lo = slab->cube_lo > pad ? slab->cube_lo - pad : 0;
  const uint64_t hi = std::min<uint64_t>(slab->cube_hi - pad, full_extent - 1) + 1;
  slab->buf_lo = lo;
  if (in_lo) *in_lo = lo;
  if (in_hi) *in_hi = hi;
  return ZM_OK;
}

int zm_slab_finalize(zm_handle* h, int normals, int voxel_centered, int transpose, const float centering_offset[3]) {
  if (!h) return ZM_ERR_INVALID;
  if (!h->comm) return fail(h, ZM_ERR_STATE, "zm_comm_init has not been called");
  if (!h->slab_mode) return fail(h, ZM_ERR_STATE, "zm_slab_step has not been called");
  const NcclApi& N = nccl_api();
  const bool last = h->rank == h->world - 1;
  const bool want_normals = normals && !(h->fin.normals_valid && h->fin.normals_transpose == (transpose ? 1 : 0));
  const uint64_t n = zm_plane_elems(h);
  cudaStream_t st = h->stream;
  int rc;
  ZM_CUDA(h, cudaSetDevice(h->device));
  if (want_normals && !last) {
    ZM_CUDA(h, h->d_nplane_out.ensure(n * 12));
    h->nplane_out = h->d_nplane_out.as<float>();
  }
  if (!last) {  // all tiles below the top layer with the cheaper kernel variant (they never look at the boundary plane)
    rc = do_finalize(h, normals, voxel_centered, transpose, centering_offset, true);
    if (rc != ZM_OK) return rc;
  }
  rc = do_finalize(h, normals, voxel_centered, transpose, centering_offset);
  if (rc != ZM_OK) return rc;
  if (want_normals) {
    // normal contributions of the top cube layer to the next shard's first-plane vertices: rank r -> r + 1 (collective:
    // every shard of the step must make this call)
    if (h->rank > 0) ZM_CUDA(h, h->d_nplane_in.ensure(n * 12));
    ZM_NCCL(h, N.GroupStart());
    if (!last) ZM_NCCL(h, N.Send(h->d_nplane_out.p, n * 3, ncclFloat32, 1, h->comm_hi, st));
    if (h->rank > 0) ZM_NCCL(h, N.Recv(h->d_nplane_in.p, n * 3, ncclFloat32, 0, h->comm_lo, st));
    ZM_NCCL(h, N.GroupEnd());
    if (h->rank > 0 && h->Vtot && h->n_work) {
      rc = zm_add_normal_plane(h, h->d_nplane_in.as<float>());
      if (rc != ZM_OK) return rc;
    }
    rc = zm_finish_normals(h);
    if (rc != ZM_OK) return rc;
  }
  return ZM_OK;
}

int zm_slab_step(zm_handle* h, const void* labels, int label_bytes, uint64_t sx, uint64_t sy, uint64_t sz, int c_order,
                 int close, int mem_kind, uint64_t full_extent, uint64_t buf_lo, int finalize, int normals,
                 int voxel_centered, const float centering_offset[3]) {
  if (!h) return ZM_ERR_INVALID;
  if (!h->comm) return fail(h, ZM_ERR_STATE, "zm_comm_init has not been called");
  const NcclApi& N = nccl_api();
  zm_slab slab{};
  if (zm_slab_range(full_extent, close, h->rank, h->world, &slab, nullptr, nullptr) != ZM_OK)
    return fail(h, ZM_ERR_INVALID, "more shards than cube planes");
  slab.buf_lo = buf_lo;
  const bool last = slab.last != 0;
  for (int attempt = 0; attempt < 6; ++attempt) {
    int rc = run_mesh(h, labels, label_bytes, sx, sy, sz, c_order, close, mem_kind, &slab);
    h->failed = rc != ZM_OK;
    if (rc != ZM_OK) return rc;
    cudaStream_t st = h->stream;
    // 1. label directories: export -> all-gather -> per-label offsets, all on the handle's stream
    ZM_CUDA(h, cudaEventRecord(h->ev_comm[0], st));
    const uint64_t cap = h->dir_cap, words = 2 * (1 + cap);
    ZM_CUDA(h, h->d_dir_mine.ensure(words * 8));
    ZM_CUDA(h, h->d_dir_all.ensure(words * 8 * (uint64_t)h->world));
    rc = zm_export_directory(h, h->d_dir_mine.as<uint64_t>(), cap);
    if (rc != ZM_OK) return rc;
    ZM_NCCL(h, N.AllGather(h->d_dir_mine.p, h->d_dir_all.p, words, ncclUint64, h->comm, st));
    rc = zm_import_directories(h, h->d_dir_all.as<uint64_t>(), (uint32_t)h->world, (uint32_t)h->rank, cap);
    if (rc != ZM_OK) return rc;
    // 2. boundary plane: rank r + 1 -> rank r (64 MiB for c5: 0.14 ms over NVLink).  On the handle's own stream: pass 2
    //    is a persistent kernel that fills every SM, so a transfer queued beside it on a second stream does not start
    //    before pass 2 retires (measured: 0.95 ms lost per step at 8 GPUs) -- it runs first instead.
    const uint64_t n = zm_plane_elems(h);
    if (h->rank > 0) {
      ZM_CUDA(h, h->d_plane_send.ensure(n * 4));
      rc = zm_export_plane(h, h->d_plane_send.as<uint32_t>());
      if (rc != ZM_OK) return rc;
    }
    if (!last) ZM_CUDA(h, h->d_plane_recv.ensure(n * 4));
    ZM_NCCL(h, N.GroupStart());
    if (h->rank > 0) ZM_NCCL(h, N.Send(h->d_plane_send.p, n, ncclUint32, 0, h->comm_lo, st));
    if (!last) ZM_NCCL(h, N.Recv(h->d_plane_recv.p, n, ncclUint32, 1, h->comm_hi, st));
    ZM_NCCL(h, N.GroupEnd());
    h->foreign = last ? nullptr : h->d_plane_recv.as<uint32_t>();
    ZM_CUDA(h, cudaEventRecord(h->ev_comm[1], st));
    if (!(finalize || normals)) {  // (the caller finalizes later: only the overflow check is left, it needs the host)
      ZM_CUDA(h, cudaMemcpyAsync(&h->h_ctl->flags, &h->d_ctl.as<Control>()->flags, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
      ZM_CUDA(h, cudaStreamSynchronize(st));
      if (h->h_ctl->flags & FLAG_DIR) { h->dir_cap *= 4; continue; }
      return ZM_OK;
    }
    rc = zm_slab_finalize(h, normals, voxel_centered, 0, centering_offset);
    if (rc == ZM_ERR_STATE && (h->h_ctl->flags & FLAG_DIR)) {  // every shard sees every directory size: all of them repeat
      h->dir_cap *= 4;
      continue;
    }
    if (rc == ZM_OK) cudaEventElapsedTime(&h->stats.ms_exchange, h->ev_comm[0], h->ev_comm[1]);  // (waits for the slowest rank included)
    return rc;
  }
  return fail(h, ZM_ERR_UNSUPPORTED, "slab step: the label directory exchange did not converge");
}

int zm_sync(zm_handle* h) {
  if (!h) return ZM_ERR_INVALID;
  ZM_CUDA(h, cudaSetDevice(h->device));
  ZM_CUDA(h, cudaStreamSynchronize(h->stream));
  return ZM_OK;
}

}  // extern "C"

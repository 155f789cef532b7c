/* zmesh_b200 -- C ABI of the B200-native multi-label marching-cubes path.
 *
 * This is the drop-in boundary: it replaces the C++ template the reference binds from Cython,
 *
 *     cdef cppclass CMesher[P, L, S]                      (reference zmesh/_zmesh.pyx:74-108)
 *     class CMesher<PositionType, LabelType, SimplifierType>   (reference zmesh/cMesher.hpp:16-308)
 *
 * Label width is a run-time argument (the reference instantiates 8 classes,
 * zmesh/_zmesh.pyx:739-1033); vertex keys are always the 64-bit layout
 * (zi_lib/zi/mesh/marching_cubes.hpp:60-74), which is unobservable in the results.
 * Plain pointers and sizes only; every function returns a zm_status (0 = ok) unless noted, and
 * zm_last_error() gives the text.  A handle is not thread-safe (neither is the reference).
 *
 * There is no CPU fallback: every entry point that computes runs CUDA kernels built for sm_100a
 * and fails with ZM_ERR_CUDA when no such device is usable.
 */
#ifndef ZMESH_B200_H
#define ZMESH_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct zm_handle zm_handle;

typedef enum {
  ZM_OK = 0,
  ZM_ERR_INVALID = 1,     /* bad argument                                   */
  ZM_ERR_CUDA = 2,        /* CUDA runtime/driver failure (text in last_error) */
  ZM_ERR_OOM = 3,         /* device or pinned-host allocation failed        */
  ZM_ERR_UNSUPPORTED = 4, /* e.g. > 2^32-1 vertices in one call             */
  ZM_ERR_STATE = 5        /* call order (get before mesh, ...)              */
} zm_status;

enum { ZM_MEM_HOST = 0, ZM_MEM_DEVICE = 1 };

/* Replaces CMesher(const std::vector<float>& voxelresolution) (cMesher.hpp:24-26) and
 * Mesher.__init__ (zmesh/_zmesh.pyx:442-444).  `resolution` is captured here, as the reference
 * captures it when Mesher.mesh() constructs the C++ object (zmesh/_zmesh.pyx:494).
 * device < 0 selects the current CUDA device. */
int zm_create(const float resolution[3], int device, zm_handle** out);
void zm_destroy(zm_handle* h);

/* Run all work of this handle on a caller-owned CUDA stream (cudaStream_t passed as void*, e.g.
 * torch.cuda.current_stream().cuda_stream) instead of the handle's own; NULL restores the own one. */
int zm_set_stream(zm_handle* h, void* cuda_stream);

/* Order the handle's next work after everything queued so far on `producer_stream` (cudaStream_t as void*;
 * NULL or 1 = the legacy default stream, 2 = the per-thread default stream, as in __cuda_array_interface__ v3):
 * call before zm_mesh(..., ZM_MEM_DEVICE) when the label volume was written by kernels on another stream.  The
 * reference has no counterpart (its input is host memory it reads synchronously, cMesher.hpp:29-36). */
int zm_wait_stream(zm_handle* h, void* producer_stream);

/* Replaces the resolution captured by `MesherClass(self.voxel_res)` in Mesher.mesh
 * (zmesh/_zmesh.pyx:494): call before zm_mesh to re-capture. */
int zm_set_resolution(zm_handle* h, const float resolution[3]);

/* Replaces CMesher::mesh(const L* data, sx, sy, sz, c_order) (cMesher.hpp:29-36 ->
 * marching_cubes::marche, zi_lib/zi/mesh/marching_cubes.hpp:291-445, 644-656) together with the
 * `close` zero padding of Mesher.mesh (zmesh/_zmesh.pyx:502-506), which is applied virtually on
 * the device (out-of-range voxels read as 0, coordinates shifted by +1 voxel like the padded copy).
 *   labels       dense volume, logical shape (sx, sy, sz); element (x,y,z) at
 *                z + sz*(y + sy*x) if c_order else x + sx*(y + sy*z); borrowed for this call only
 *   label_bytes  1, 2, 4 or 8 (values are compared as unsigned bit patterns of that width)
 *   mem_kind     ZM_MEM_HOST (pageable or pinned) or ZM_MEM_DEVICE (pointer on the handle's device)
 * Previous results of the handle are dropped (as Mesher.mesh deletes the old CMesher, :469). */
int zm_mesh(zm_handle* h, const void* labels, int label_bytes, uint64_t sx, uint64_t sy,
            uint64_t sz, int c_order, int close, int mem_kind);

/* ---- multi-GPU slab decomposition (no reference counterpart; driven by zmesh_b200/sharded.py) ----
 * The volume is cut along its slowest memory axis (z for Fortran order, x for C order) into slabs,
 * one handle (one GPU, one process) each.  In EXTENDED plane coordinates (input plane + 1 when
 * close, else input plane) a shard meshes the cubes whose origin plane lies in [cube_lo, cube_hi)
 * and therefore reads planes [cube_lo, cube_hi]; `labels` holds input planes
 * [buf_lo, buf_lo + extent of the buffer along the slab axis).  Vertex slots are owned by voxel:
 * the shard owns planes [cube_lo, cube_hi), plus plane cube_hi when `last`; the slots of a
 * non-last shard's top plane belong to the next shard, so every vertex exists exactly once and
 * per-label partial meshes concatenate without dedup.  Face indices become cross-shard indices
 * through zm_set_label_offsets (vertices of the label on earlier shards) and the boundary-plane
 * exchange zm_export_plane -> (NCCL send/recv) -> zm_set_foreign_plane. */
typedef struct {
  uint64_t full_extent;      /* voxels of the whole volume along the slab axis               */
  uint64_t buf_lo;           /* input plane index of the buffer's first plane                */
  uint64_t cube_lo, cube_hi; /* extended planes of the cube origins this shard owns          */
  int last;                  /* 1 iff cube_hi is the last extended plane of the volume       */
} zm_slab;
int zm_mesh_slab(zm_handle* h, const void* labels, int label_bytes, uint64_t sx, uint64_t sy,
                 uint64_t sz, int c_order, int close, int mem_kind, const zm_slab* slab);

/* Per-label counts of the last zm_mesh / zm_mesh_slab in storage order (the order of the bulk view). */
uint64_t zm_num_directory(zm_handle* h);
int zm_directory(zm_handle* h, uint64_t* labels, uint64_t* n_vertices, uint64_t* n_faces, uint64_t capacity);

/* offsets[i] is added to every face index of labels[i] (= number of its vertices on earlier
 * shards).  Call after zm_mesh_slab and before the first zm_get / zm_finalize / zm_export_plane. */
int zm_set_label_offsets(zm_handle* h, const uint64_t* labels, const uint32_t* offsets, uint64_t n);

/* The same without a host round trip (all work queued on the handle's stream): zm_export_directory writes this
 * shard's directory into dst_device[2 * (1 + capacity)] -- word 0 = number of labels, then (label, n_vertices)
 * pairs; the caller all-gathers the buffers of all shards in rank order (e.g. ncclAllGather on the same stream) and
 * zm_import_directories(all, world, rank, capacity) sums, per label of this shard, the vertices on earlier shards.
 * A directory that does not fit `capacity` is reported by the next zm_finalize / zm_get as ZM_ERR_STATE: repeat the
 * step with a larger capacity. */
int zm_export_directory(zm_handle* h, uint64_t* dst_device, uint64_t capacity);
int zm_import_directories(zm_handle* h, const uint64_t* all_device, uint32_t world, uint32_t rank, uint64_t capacity);

/* Writes the cross-shard indices of the in-plane vertex slots of this shard's first plane into
 * dst_device[zm_plane_elems(h)] (uint32 [Em][Efp][4]); the shard below passes the received copy to
 * zm_set_foreign_plane (pointer borrowed until the next zm_mesh*). */
uint64_t zm_plane_elems(zm_handle* h);
int zm_export_plane(zm_handle* h, uint32_t* dst_device);
int zm_set_foreign_plane(zm_handle* h, const uint32_t* src_device);

/* Normals across slab shards.  The faces of a shard's top cube layer touch vertices owned by the next
 * shard; their contributions (compute_vertex_normals_from_faces, zmesh/chunk_mesh.hpp:345-384) are
 * accumulated in a caller-owned device buffer of 3 * zm_plane_elems(h) floats (zm_set_normal_plane, before
 * the zm_finalize with normals; zeroed by the library), sent to the next shard, added there
 * (zm_add_normal_plane, after its zm_finalize) and normalised by zm_finish_normals.  On a slab shard
 * zm_finalize(normals=1) leaves the normals un-normalised until zm_finish_normals is called. */
int zm_set_normal_plane(zm_handle* h, float* out_device);
int zm_add_normal_plane(zm_handle* h, const float* src_device);
int zm_finish_normals(zm_handle* h);

/* Replaces CMesher::ids() (cMesher.hpp:46-54).  The reference's order is unspecified
 * (unordered_map iteration); here ids are ascending.  Labels that produced no triangle are
 * absent, label 0 is never meshed (marching_cubes.hpp:438). */
uint64_t zm_num_ids(zm_handle* h); /* returns the count */
int zm_ids(zm_handle* h, uint64_t* out, uint64_t capacity);

/* Sizes for zm_get: number of vertices and faces of `label` (0, 0 if absent or erased;
 * cf. marching_cubes::count, marching_cubes.hpp:208 and cMesher.hpp:69-74). */
int zm_get_counts(zm_handle* h, uint64_t label, uint64_t* n_vertices, uint64_t* n_faces);

/* Replaces CMesher::get_mesh(label, normals=false, simplification_factor=0, ..., transpose)
 * -> triangles2mesh (cMesher.hpp:60-166) fused with the Python post-processing of Mesher.get /
 * Mesher.get_mesh: compute_normals (zmesh/_zmesh.pyx:138-152 -> zmesh/chunk_mesh.hpp:345-384) and
 * _normalize_mesh (zmesh/_zmesh.pyx:423-433):
 *     vertex_i = fl32( fl32( fl32(res_i * k_i) [+ offset_i if voxel_centered] ) / 2 )
 * with res = the captured resolution and offset = `centering_offset` (the reference adds the
 * Mesher's *current* voxel_res, which may differ from the captured one); NULL = captured.
 *   transpose  0: Mesher.get orientation; 1: legacy Mesher.get_mesh (x/z swapped, winding flipped)
 *   vertices   caller buffer, 3*n_vertices floats;  faces: 3*n_faces uint32 (label-local indices)
 *   normals_out  NULL, or 3*n_vertices floats (unit normals, NaN where the reference gives NaN)
 * Vertex/face order is unspecified (as in the reference, whose order is hash-map order). */
int zm_get(zm_handle* h, uint64_t label, int normals, int voxel_centered, int transpose,
           const float centering_offset[3], float* vertices, uint32_t* faces, float* normals_out);

/* Replaces compute_vertex_normals_from_faces (zmesh/chunk_mesh.hpp:345-384) as exposed by
 * Mesher.compute_normals (zmesh/_zmesh.pyx:138-152, :585-590) for an arbitrary host mesh:
 * vertices 3*n_vertices float32, faces 3*n_faces uint32 -> normals_out 3*n_vertices float32. */
int zm_compute_normals(zm_handle* h, const float* vertices, uint64_t n_vertices, const uint32_t* faces,
                       uint64_t n_faces, float* normals_out);

/* Replaces CMesher::erase / clear (cMesher.hpp:300-307 -> marching_cubes.hpp:184-204). */
int zm_erase(zm_handle* h, uint64_t label, int* existed);
int zm_clear(zm_handle* h);

/* ---- bulk / device-resident access (no reference counterpart; used by bench.py and by callers
 * that keep results on the GPU) ------------------------------------------------------------- */

/* Runs the final gather for ALL labels with the given options and leaves the result on the device:
 * vertices float32 [V_total][3], faces uint32 [T_total][3] (label-local indices), optional normals
 * float32 [V_total][3]; label i of the directory (STORAGE order, which is unspecified -- not
 * sorted; erased labels included) owns vertex rows [voff[i], voff[i+1]) and face rows
 * [foff[i], foff[i+1]).  With transpose != 0 the device faces keep the Mesher.get winding; the
 * host copies (zm_get, zm_fetch_all) reverse each row to the legacy winding.  Pointers stay valid
 * until the next zm_mesh / zm_clear / zm_destroy on this handle. */
typedef struct {
  uint64_t n_labels, n_vertices, n_faces;
  const uint64_t* labels_host; /* [n_labels] storage order                   */
  const uint64_t* voff_host;   /* [n_labels+1]                               */
  const uint64_t* foff_host;   /* [n_labels+1]                               */
  const float* vertices_dev;
  const uint32_t* faces_dev;
  const float* normals_dev; /* NULL unless requested */
} zm_bulk_view;
int zm_finalize(zm_handle* h, int normals, int voxel_centered, int transpose,
                const float centering_offset[3], zm_bulk_view* view);

/* Slab shards that are not the last one: run pass 2 for every tile except the top tile layer -- the only tiles whose
 * cubes reference the next shard's boundary plane -- with the kernel variant that has no boundary-plane lookups, without
 * synchronising.  The zm_finalize that follows, with the same arguments and after zm_set_foreign_plane, emits the top
 * layer.  A no-op for unsharded volumes and for the last shard.  (The plane transfer itself is best queued BEFORE this
 * call on the same stream: pass 2 is a persistent kernel that fills every SM, a transfer queued beside it on a second
 * stream does not start until it retires -- measured) */
int zm_finalize_begin(zm_handle* h, int normals, int voxel_centered, int transpose, const float centering_offset[3]);

/* Copies the finalized arrays of all labels to host buffers in one transfer each
 * (vertices 3*V_total floats, faces 3*T_total uint32, normals 3*V_total floats or NULL). */
int zm_fetch_all(zm_handle* h, float* vertices, uint32_t* faces, float* normals_out);

/* Mesh wire format on the device (replaces Mesh.to_precomputed applied label by label, zmesh/mesh.py:257-269, the step
 * that follows Mesher.get in production callers): zm_pack_precomputed finalizes (no normals) and gathers, for every id
 * in ascending order, the Neuroglancer "Precomputed" object -- uint32 Nv | float32 vertices [Nv][3] | uint32 faces
 * [Nf][3] -- into one device buffer; zm_fetch_precomputed copies it to dst_host[total_bytes] (pinned memory from
 * zm_host_alloc for full speed) in ONE transfer and returns the ids and the byte offset of every object
 * (byte_offsets_out[n_objects] = total_bytes). */
int zm_pack_precomputed(zm_handle* h, int voxel_centered, const float centering_offset[3], uint64_t* n_objects,
                        uint64_t* total_bytes);
int zm_fetch_precomputed(zm_handle* h, void* dst_host, uint64_t* labels_out, uint64_t* byte_offsets_out);

/* Page-locked host memory for zm_fetch_all destinations (full-speed D2H); NULL on failure. */
void* zm_host_alloc(uint64_t bytes);
void zm_host_free(void* p);

typedef struct {
  uint64_t n_voxels, n_labels, n_vertices, n_faces;
  uint64_t n_records;      /* (label, cube) pairs that produced triangles                          */
  uint64_t n_tiles, n_active_tiles, n_dense_tiles; /* 32x8x8 tiles: all / non-uniform / redone in dense mode */
  uint64_t hash_capacity, perm_capacity;
  uint32_t attempts;   /* classification passes run (1 unless a capacity guess was too small) */
  uint32_t launches;   /* kernels launched by the last zm_mesh                                 */
  uint32_t used_tma;   /* 1: tiles staged by TMA, 0: volume not 16-byte aligned, plain loads    */
  uint32_t launches_finalize;
  float ms_h2d, ms_classify, ms_scan, ms_total; /* CUDA-event times of the last zm_mesh          */
  float ms_faces, ms_vertices, ms_finalize;     /* last zm_finalize / first zm_get (pass 2): the fused emit
                                                   kernel, the normals normalisation (0 without normals), both */
  float ms_exchange;  /* zm_slab_step: directory all-gather + boundary-plane transfer, incl. the wait for the slowest shard */
} zm_stats_t;
int zm_stats(zm_handle* h, zm_stats_t* out);

/* Synthetic input for benchmarks/tests: integer jittered-grid Voronoi segmentation (SURVEY.md
 * section 8d) written straight into device memory.  Generates the sub-block `shape` at `origin`
 * of a volume of `full_shape`; one site per cell of side `pitch`; nearest site among the 27
 * neighbouring cells wins, ties to the smaller cell id; label = splitmix64(cell+1)|1 for 8-byte
 * labels, cell+1 otherwise.  Bit-identical to oracle/oracle.py:voronoi_volume. */
int zm_synth_voronoi(void* dst_device, int label_bytes, const uint64_t shape[3], const uint64_t origin[3],
                     const uint64_t full_shape[3], uint32_t pitch, uint64_t seed, int c_order,
                     void* cuda_stream);

/* ---- native multi-GPU step (one process per GPU; no reference counterpart) -------------------------------------
 * The whole slab step of zmesh_b200/sharded.py in ONE call, with NCCL driven from C++ (bound at run time with
 * dlopen("libnccl.so.2"); inside a torch process that is the library torch loaded): zm_mesh_slab -> directory
 * all-gather (ncclAllGather on the handle's stream) -> boundary-plane send/recv on a side stream overlapped with pass 2
 * of every tile below the top layer -> pass 2 of the top layer (-> the normals plane exchange).  The host issues a
 * dozen launches after the one synchronisation zm_mesh_slab needs; the Python/torch.distributed path costs ~2 ms of
 * exposed host time per step at 8 GPUs.
 *   zm_nccl_unique_id   128-byte ncclUniqueId (call on one rank, broadcast the bytes by any means; `world` are needed)
 *   zm_comm_init        ncclCommInitRank for this handle's device: one communicator of all shards for the directory
 *                       all-gather (id_collectives) and one 2-rank communicator per neighbour pair for the boundary
 *                       planes (id_pairs: world - 1 ids of 128 bytes, id k = shards k and k + 1; a send/recv inside a
 *                       2-rank communicator gets all NVLink channels)
 *   zm_slab_range       the cube planes [cube_lo, cube_hi) of `rank` and the input planes [in_lo, in_hi) it must supply
 *   zm_slab_step        `labels` holds input planes [buf_lo, buf_lo + extent along the slab axis) covering that range;
 *                       finalize / normals / voxel_centered as zm_finalize (results stay distributed: zm_get returns
 *                       this shard's part of a label, face indices are cross-shard) */
int zm_nccl_unique_id(void* out128);
int zm_comm_init(zm_handle* h, const void* id_collectives, const void* id_pairs, int world, int rank);
int zm_comm_destroy(zm_handle* h);
int zm_slab_range(uint64_t full_extent, int close, int rank, int world, zm_slab* slab, uint64_t* in_lo, uint64_t* in_hi);
int zm_slab_step(zm_handle* h, const void* labels, int label_bytes, uint64_t sx, uint64_t sy, uint64_t sz, int c_order,
                 int close, int mem_kind, uint64_t full_extent, uint64_t buf_lo, int finalize, int normals,
                 int voxel_centered, const float centering_offset[3]);
/* Pass 2 of a slab step that was run with finalize = 0, or a later request for normals / another vertex transform:
 * zm_finalize for a shard, including the normals plane exchange.  COLLECTIVE when normals are (re)computed: every shard
 * of the step must call it with the same arguments. */
int zm_slab_finalize(zm_handle* h, int normals, int voxel_centered, int transpose, const float centering_offset[3]);

/* Blocks until all work queued by this handle has finished. */
int zm_sync(zm_handle* h);

const char* zm_last_error(zm_handle* h); /* h may be NULL: error of the last failed zm_create */
const char* zm_version(void);

#ifdef __cplusplus
}
#endif
#endif /* ZMESH_B200_H */

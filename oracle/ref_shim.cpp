// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// C-ABI shim over the UNMODIFIED reference C++ headers (seung-lab/zmesh).  It is
// compiled *in place* against /root/reference (include path only; no reference
// source is copied into this repository) by oracle/Makefile into
// oracle/_ref/libzmesh_ref.so.  The shim instantiates the very template the
// reference's Cython layer binds (zmesh/_zmesh.pyx:74-108):
//
//     zmesh::CMesher<PositionType, LabelType, float>      (zmesh/cMesher.hpp:16-308)
//     zmesh::compute_vertex_normals_from_faces            (zmesh/chunk_mesh.hpp:345-384)
//
// for the eight (P, L) combinations the reference generates
// (zmesh/_zmesh.pyx:739-1033) and type-erases them behind plain C functions so
// tests/ and bench.py (cpu_baseline / --impl reference) can drive the real
// reference without Cython.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "zmesh/cMesher.hpp"
#include "zmesh/chunk_mesh.hpp"

namespace {

struct RefBase {
  virtual ~RefBase() {}
  virtual void mesh(const void* data, size_t sx, size_t sy, size_t sz, bool c_order) = 0;
  virtual std::vector<uint64_t> ids() = 0;
  virtual zmesh::utility::MeshObject get(uint64_t label, bool transpose) = 0;
  virtual bool erase(uint64_t label) = 0;
  virtual void clear() = 0;
};

template <typename P, typename L>
struct RefImpl final : RefBase {
  zmesh::CMesher<P, L, float> m;
  explicit RefImpl(const std::vector<float>& res) : m(res) {}
  void mesh(const void* data, size_t sx, size_t sy, size_t sz, bool c_order) override {
    m.mesh(static_cast<const L*>(data), sx, sy, sz, c_order);
  }
  std::vector<uint64_t> ids() override {
    std::vector<L> v = m.ids();
    return std::vector<uint64_t>(v.begin(), v.end());
  }
  zmesh::utility::MeshObject get(uint64_t label, bool transpose) override {
    // Mesher.get(): get_mesh(label, False, reduction_factor=0, max_error, transpose)
    // (zmesh/_zmesh.pyx:572-576); min error default 25*eps (:988).
    return m.get_mesh(static_cast<L>(label), false, 0, 40.0f, 25.0f * 1.1920929e-07f, transpose);
  }
  bool erase(uint64_t label) override { return m.erase(static_cast<L>(label)); }
  void clear() override { m.clear(); }
};

template <typename P>
RefBase* make_for_label(int label_bytes, const std::vector<float>& res) {
  switch (label_bytes) {
    case 1: return new RefImpl<P, uint8_t>(res);
    case 2: return new RefImpl<P, uint16_t>(res);
    case 4: return new RefImpl<P, uint32_t>(res);
    case 8: return new RefImpl<P, uint64_t>(res);
  }
  return nullptr;
}

}  // namespace

extern "C" {

void* zref_create(int pos_bits, int label_bytes, const float* res) {
  std::vector<float> r(res, res + 3);
  if (pos_bits == 32) return make_for_label<uint32_t>(label_bytes, r);
  if (pos_bits == 64) return make_for_label<uint64_t>(label_bytes, r);
  return nullptr;
}

void zref_destroy(void* h) { delete static_cast<RefBase*>(h); }

void zref_mesh(void* h, const void* data, uint64_t sx, uint64_t sy, uint64_t sz, int c_order) {
  static_cast<RefBase*>(h)->mesh(data, sx, sy, sz, c_order != 0);
}

// ids: two-call protocol (out == NULL returns the count).
uint64_t zref_ids(void* h, uint64_t* out, uint64_t cap) {
  std::vector<uint64_t> v = static_cast<RefBase*>(h)->ids();
  if (out) {
    uint64_t n = v.size() < cap ? v.size() : cap;
    std::memcpy(out, v.data(), n * sizeof(uint64_t));
  }
  return v.size();
}

// get: result buffers are malloc'ed here and released with zref_free.
void zref_get(void* h, uint64_t label, int transpose, float** points, uint64_t* n_point_floats,
              uint32_t** faces, uint64_t* n_face_ints) {
  zmesh::utility::MeshObject mo = static_cast<RefBase*>(h)->get(label, transpose != 0);
  *n_point_floats = mo.points.size();
  *n_face_ints = mo.faces.size();
  *points = static_cast<float*>(std::malloc(sizeof(float) * (mo.points.size() + 1)));
  *faces = static_cast<uint32_t*>(std::malloc(sizeof(uint32_t) * (mo.faces.size() + 1)));
  std::memcpy(*points, mo.points.data(), sizeof(float) * mo.points.size());
  std::memcpy(*faces, mo.faces.data(), sizeof(uint32_t) * mo.faces.size());
}

void zref_free(void* p) { std::free(p); }

int zref_erase(void* h, uint64_t label) { return static_cast<RefBase*>(h)->erase(label) ? 1 : 0; }

void zref_clear(void* h) { static_cast<RefBase*>(h)->clear(); }

// compute_normals (zmesh/_zmesh.pyx:138-152): Nf passed is faces.size (3 per face).
void zref_normals(const float* verts, uint64_t nv, const uint32_t* faces, uint64_t n_face_ints,
                  float* out) {
  std::vector<float> n = zmesh::chunk_mesh::compute_vertex_normals_from_faces(verts, nv, faces, n_face_ints);
  std::memcpy(out, n.data(), sizeof(float) * n.size());
}

}  // extern "C"

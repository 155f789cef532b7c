/* TEST INFRASTRUCTURE ONLY -- CPU oracle, never on the product path.
 *
 * Plain-C restatement of the reference algorithm of seung-lab/zmesh for the hot path
 *     Mesher.mesh(labels) -> Mesher.get(label, reduction_factor=0)
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  Parity status: PINNED -- checked against (a) the reference's own golden
 * PLY meshes for connectomics.npy (connectomics_npy_meshes/unsimplified, see
 * tests/test_oracle_golden.py) and (b) the unmodified reference C++ compiled in place
 * (oracle/_ref/libzmesh_ref.so, tests/test_oracle_vs_ref.py).
 *
 * Every function cites the reference lines it follows (paths relative to the reference root).
 * Keys are always the 64-bit layout x<<42 | y<<21 | z (marching_cubes.hpp:60-74); the 32-bit
 * layout (:76-91) unpacks to the same coordinates, so key width is unobservable.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "mc_tables_oracle.h"

/* ------------------------------------------------------------------------------------------ */
/* per-label triangle soup: std::unordered_map<Label, std::deque<triangle>>                     */
/* (zi_lib/zi/mesh/marching_cubes.hpp:172-173)                                                  */

typedef struct {
  uint64_t label;
  uint64_t* tri; /* 3 keys per triangle */
  size_t ntri, cap;
  int alive;
} zo_soup;

typedef struct {
  float res[3];   /* captured at construction (zmesh/cMesher.hpp:24-26) */
  zo_soup* soups; /* insertion order */
  size_t nsoups, soups_cap;
  uint64_t* map_key; /* open-addressing label -> soup index + 1 */
  uint32_t* map_val;
  size_t map_cap; /* power of two */
} zo_mesher;

static uint64_t mix64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull;
  x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull;
  x ^= x >> 33;
  return x;
}

static void map_insert_raw(zo_mesher* m, uint64_t label, uint32_t val) {
  size_t i = (size_t)mix64(label) & (m->map_cap - 1);
  while (m->map_val[i] != 0) i = (i + 1) & (m->map_cap - 1);
  m->map_key[i] = label;
  m->map_val[i] = val;
}

static void map_grow(zo_mesher* m) {
  size_t ncap = m->map_cap ? m->map_cap * 2 : 1024;
  free(m->map_key); free(m->map_val);
  m->map_key = (uint64_t*)calloc(ncap, sizeof(uint64_t));
  m->map_val = (uint32_t*)calloc(ncap, sizeof(uint32_t));
  m->map_cap = ncap;
  for (size_t s = 0; s < m->nsoups; ++s)
    if (m->soups[s].alive) map_insert_raw(m, m->soups[s].label, (uint32_t)s + 1);
}

/* returns soup index or -1 */
static long map_find(const zo_mesher* m, uint64_t label) {
  if (!m->map_cap) return -1;
  size_t i = (size_t)mix64(label) & (m->map_cap - 1);
  while (m->map_val[i] != 0) {
    if (m->map_key[i] == label && m->soups[m->map_val[i] - 1].alive) return (long)m->map_val[i] - 1;
    i = (i + 1) & (m->map_cap - 1);
  }
  return -1;
}

/* meshes_[label] (marching_cubes.hpp:333) */
static zo_soup* soup_for(zo_mesher* m, uint64_t label) {
  long s = map_find(m, label);
  if (s >= 0) return &m->soups[s];
  if ((m->nsoups + 1) * 2 > m->map_cap) map_grow(m);
  if (m->nsoups == m->soups_cap) {
    m->soups_cap = m->soups_cap ? m->soups_cap * 2 : 256;
    m->soups = (zo_soup*)realloc(m->soups, m->soups_cap * sizeof(zo_soup));
  }
  zo_soup* sp = &m->soups[m->nsoups];
  sp->label = label; sp->tri = NULL; sp->ntri = 0; sp->cap = 0; sp->alive = 1;
  map_insert_raw(m, label, (uint32_t)m->nsoups + 1);
  m->nsoups++;
  return sp;
}

static void soup_push(zo_soup* s, uint64_t a, uint64_t b, uint64_t c) {
  if (s->ntri == s->cap) {
    s->cap = s->cap ? s->cap * 2 : 64;
    s->tri = (uint64_t*)realloc(s->tri, s->cap * 3 * sizeof(uint64_t));
  }
  uint64_t* t = s->tri + 3 * s->ntri++;
  t[0] = a; t[1] = b; t[2] = c;
}

/* ------------------------------------------------------------------------------------------ */

/* pack_coords, 64-bit traits (marching_cubes.hpp:60-74, :108-112) */
static uint64_t pack(uint64_t x, uint64_t y, uint64_t z) { return (x << 42) | (y << 21) | z; }

zo_mesher* zo_create(const float* res) {
  zo_mesher* m = (zo_mesher*)calloc(1, sizeof(zo_mesher));
  m->res[0] = res[0]; m->res[1] = res[1]; m->res[2] = res[2];
  return m;
}

/* marching_cubes::clear (marching_cubes.hpp:184-189) */
void zo_clear(zo_mesher* m) {
  for (size_t s = 0; s < m->nsoups; ++s) free(m->soups[s].tri);
  free(m->soups); free(m->map_key); free(m->map_val);
  m->soups = NULL; m->nsoups = m->soups_cap = 0;
  m->map_key = NULL; m->map_val = NULL; m->map_cap = 0;
}

void zo_destroy(zo_mesher* m) {
  if (!m) return;
  zo_clear(m);
  free(m);
}

static uint64_t load_label(const void* data, int label_bytes, size_t i) {
  switch (label_bytes) {
    case 1: return ((const uint8_t*)data)[i];
    case 2: return ((const uint16_t*)data)[i];
    case 4: return ((const uint32_t*)data)[i];
    default: return ((const uint64_t*)data)[i];
  }
}

/* CMesher::mesh (cMesher.hpp:29-36) -> marching_cubes::marche<Tag> (marching_cubes.hpp:291-445).
 * Logical axis 0 = x, 1 = y, 2 = z regardless of memory order (get_strides :259-289).  The
 * skip_check sliding-window shortcut (:346, :366-409) only elides compares and is not restated:
 * a uniform cube yields mask 0xFF -> case 0 -> mc_edge_table[0] == 0 -> nothing appended. */
void zo_mesh(zo_mesher* m, const void* data, int label_bytes, uint64_t sx, uint64_t sy, uint64_t sz,
             int c_order) {
  if (sx < 2 || sy < 2 || sz < 2) return;
  /* element strides of the logical axes */
  size_t stx, sty, stz;
  if (c_order) { stx = (size_t)(sy * sz); sty = (size_t)sz; stz = 1; }
  else         { stx = 1; sty = (size_t)sx; stz = (size_t)(sx * sy); }

  /* cube_corners (:299-302) as logical (dx,dy,dz), and the packed edge midpoints (:304-316):
   * midpoint(p1,p2) = p1/2 + p2/2 with corners at 0 or 2 half-voxel units => corner_a + corner_b
   * in voxel units. */
  static const int cdx[8] = {0, 1, 1, 0, 0, 1, 1, 0};
  static const int cdy[8] = {0, 0, 0, 0, 1, 1, 1, 1};
  static const int cdz[8] = {0, 0, 1, 1, 0, 0, 1, 1};
  static const int ea[12] = {0, 1, 2, 3, 4, 5, 6, 7, 0, 1, 2, 3};
  static const int eb[12] = {1, 2, 3, 0, 5, 6, 7, 4, 4, 5, 6, 7};
  uint64_t mid[12];
  for (int e = 0; e < 12; ++e)
    mid[e] = pack((uint64_t)(cdx[ea[e]] + cdx[eb[e]]), (uint64_t)(cdy[ea[e]] + cdy[eb[e]]),
                  (uint64_t)(cdz[ea[e]] + cdz[eb[e]]));
  size_t coff[8];
  for (int n = 0; n < 8; ++n) coff[n] = cdx[n] * stx + cdy[n] * sty + cdz[n] * stz;

  /* mc_nested_loops (:226-257): memory-order-major traversal. */
  uint64_t n_outer = c_order ? sx - 1 : sz - 1;
  uint64_t n_inner = c_order ? sz - 1 : sx - 1;
  for (uint64_t o = 0; o < n_outer; ++o)
    for (uint64_t y = 0; y < sy - 1; ++y)
      for (uint64_t i = 0; i < n_inner; ++i) {
        uint64_t x = c_order ? o : i, z = c_order ? i : o;
        size_t ind = (size_t)(x * stx + y * sty + z * stz);
        uint64_t lab[8]; /* (:353-361) */
        for (int n = 0; n < 8; ++n) lab[n] = load_label(data, label_bytes, ind + coff[n]);

        /* distinct labels in first-corner order (:411-442) */
        unsigned acc = 0;
        while (acc != 0xFF) {
          int start = __builtin_ctz((~acc) & 0xFF);
          uint64_t label = lab[start];
          unsigned c = 0;
          for (int n = start; n < 8; ++n) c |= (unsigned)(lab[n] == label) << n;
          acc |= c;
          if (label == 0) continue; /* (:438) */
          unsigned cs = (~c) & 0xFF; /* add_face(..., ~c) (:440) */
          if (EDGE_USED[cs] == 0) continue; /* (:323-326) */
          uint64_t cur = pack(2 * x, 2 * y, 2 * z); /* (:328-331) */
          zo_soup* s = soup_for(m, label);
          int nt = TRI_COUNT[cs] < 5 ? TRI_COUNT[cs] : 5; /* (:335) */
          uint64_t nib = TRI_NIBBLES[cs];
          for (int t = 0; t < nt; ++t) { /* (:338-343): (E[T[n+2]], E[T[n+1]], E[T[n]]) + cur */
            int e0 = (int)((nib >> (12 * t)) & 0xF), e1 = (int)((nib >> (12 * t + 4)) & 0xF),
                e2 = (int)((nib >> (12 * t + 8)) & 0xF);
            soup_push(s, mid[e2] + cur, mid[e1] + cur, mid[e0] + cur);
          }
        }
      }
}

/* CMesher::ids (cMesher.hpp:46-54): key listing; order is unspecified in the reference
 * (unordered_map iteration) -- here insertion order. */
uint64_t zo_ids(const zo_mesher* m, uint64_t* out, uint64_t cap) {
  uint64_t n = 0;
  for (size_t s = 0; s < m->nsoups; ++s)
    if (m->soups[s].alive) {
      if (out && n < cap) out[n] = m->soups[s].label;
      n++;
    }
  return n;
}

/* marching_cubes::erase (marching_cubes.hpp:191-204) */
int zo_erase(zo_mesher* m, uint64_t label) {
  long s = map_find(m, label);
  if (s < 0) return 0;
  m->soups[s].alive = 0;
  free(m->soups[s].tri);
  m->soups[s].tri = NULL; m->soups[s].ntri = m->soups[s].cap = 0;
  return 1;
}

/* triangles2mesh (cMesher.hpp:96-166).  Two-call protocol: with verts == NULL only the counts
 * are returned.  Vertex index = order of first appearance (:102-116); vertices
 * res[i] * (0.0f + coord_i) via unpack_* (marching_cubes.hpp:114-135), transposed variant
 * (z,y,x) (:128-138); faces (t1,t2,t0) or, transposed, (t0,t2,t1) (:152-163). */
void zo_get(const zo_mesher* m, uint64_t label, int transpose, uint64_t* nv_out, uint64_t* nf_out,
            float* verts, uint32_t* faces) {
  long si = map_find(m, label);
  if (si < 0) { *nv_out = 0; *nf_out = 0; return; } /* count(segid) == 0 (cMesher.hpp:72-74) */
  const zo_soup* s = &m->soups[si];
  size_t n3 = 3 * s->ntri;
  size_t cap = 16;
  while (cap < 2 * n3) cap <<= 1;
  uint64_t* hk = (uint64_t*)malloc(cap * sizeof(uint64_t));
  uint32_t* hv = (uint32_t*)malloc(cap * sizeof(uint32_t));
  memset(hv, 0xFF, cap * sizeof(uint32_t));
  uint32_t idx = 0;
  for (size_t i = 0; i < n3; ++i) {
    uint64_t k = s->tri[i];
    size_t h = (size_t)mix64(k) & (cap - 1);
    while (hv[h] != 0xFFFFFFFFu && hk[h] != k) h = (h + 1) & (cap - 1);
    if (hv[h] == 0xFFFFFFFFu) {
      hk[h] = k; hv[h] = idx;
      if (verts) {
        float kx = 0.0f + (float)((k >> 42) & 0x1FFFFF);
        float ky = 0.0f + (float)((k >> 21) & 0x1FFFFF);
        float kz = 0.0f + (float)(k & 0x1FFFFF);
        if (transpose) {
          verts[3 * idx + 0] = m->res[0] * kz;
          verts[3 * idx + 1] = m->res[1] * ky;
          verts[3 * idx + 2] = m->res[2] * kx;
        } else {
          verts[3 * idx + 0] = m->res[0] * kx;
          verts[3 * idx + 1] = m->res[1] * ky;
          verts[3 * idx + 2] = m->res[2] * kz;
        }
      }
      idx++;
    }
    if (faces) {
      size_t t = i / 3, j = i % 3; /* tri.at(j) */
      /* non-transposed: faces = (at1, at2, at0): at0 -> col 2, at1 -> col 0, at2 -> col 1
       * transposed:     faces = (at0, at2, at1): at0 -> col 0, at1 -> col 2, at2 -> col 1 */
      static const int col_n[3] = {2, 0, 1};
      static const int col_t[3] = {0, 2, 1};
      faces[3 * t + (transpose ? col_t[j] : col_n[j])] = hv[h];
    }
  }
  free(hk); free(hv);
  *nv_out = idx;
  *nf_out = s->ntri;
}

/* Vec3<float> helpers (zmesh/utility.hpp:157-183) */
static float v_len(const float* v) { return sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }
static void v_hat(float* v) { /* :163-173: divide by the length unless it is exactly 1 */
  float l = v_len(v);
  if (l == 1.0f) return;
  v[0] /= l; v[1] /= l; v[2] /= l;
}

/* compute_vertex_normals_from_faces (zmesh/chunk_mesh.hpp:345-384); n_face_ints = 3 * faces
 * (the caller passes faces.size, zmesh/_zmesh.pyx:147-150). */
void zo_normals(const float* vertices, uint64_t nv, const uint32_t* faces, uint64_t n_face_ints,
                float* out) {
  memset(out, 0, sizeof(float) * 3 * nv);
  for (uint64_t i = 0; i + 2 < n_face_ints; i += 3) {
    uint32_t f[3] = {faces[i], faces[i + 1], faces[i + 2]};
    const float* v0 = vertices + 3 * (size_t)f[0];
    const float* v1 = vertices + 3 * (size_t)f[1];
    const float* v2 = vertices + 3 * (size_t)f[2];
    float center[3], a[3], b[3], nh[3];
    for (int d = 0; d < 3; ++d) {
      center[d] = ((v0[d] + v1[d]) + v2[d]) / 3.0f; /* (:363) */
      a[d] = v1[d] - v0[d];
      b[d] = v2[d] - v0[d];
    }
    nh[0] = a[1] * b[2] - a[2] * b[1]; /* cross (utility.hpp:177-183) */
    nh[1] = a[2] * b[0] - a[0] * b[2];
    nh[2] = a[0] * b[1] - a[1] * b[0];
    v_hat(nh); /* (:364) */
    for (int k = 0; k < 3; ++k) { /* (:366-368) */
      const float* vk = vertices + 3 * (size_t)f[k];
      float d[3] = {vk[0] - center[0], vk[1] - center[1], vk[2] - center[2]};
      float w = v_len(d);
      out[3 * (size_t)f[k] + 0] += nh[0] * w;
      out[3 * (size_t)f[k] + 1] += nh[1] * w;
      out[3 * (size_t)f[k] + 2] += nh[2] * w;
    }
  }
  for (uint64_t i = 0; i < nv; ++i) v_hat(out + 3 * i); /* (:372-381) */
}

/* ------------------------------------------------------------------------------------------------
 * Synthetic jittered-grid Voronoi volume (the same integer definition as
 * oracle.py:voronoi_volume and the device generator).  Test / benchmark input only: lets the CPU arm
 * of bench.py build its sample without touching the product library.  Fills planes [p0, p1) of the
 * block along its slowest memory axis, so callers can split a block over threads. */
static uint64_t zo_splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  uint64_t z = x;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

void zo_voronoi(void* dst, int label_bytes, const uint64_t shape[3], const uint64_t origin[3],
                const uint64_t full_shape[3], uint32_t pitch, uint64_t seed, int c_order, uint64_t p0, uint64_t p1) {
  const int64_t P = (int64_t)pitch;
  int64_t G[3];
  for (int a = 0; a < 3; ++a) {
    G[a] = (int64_t)((full_shape[a] + pitch - 1) / pitch);
    if (G[a] < 1) G[a] = 1;
  }
  /* memory axes: f fastest, s slowest */
  const int af = c_order ? 2 : 0, as = c_order ? 0 : 2;
  const uint64_t nf = shape[af], nm = shape[1];
  for (uint64_t ps = p0; ps < p1; ++ps)
    for (uint64_t pm = 0; pm < nm; ++pm) {
      int64_t q[3];
      q[as] = (int64_t)(ps + origin[as]);
      q[1] = (int64_t)(pm + origin[1]);
      const int64_t bs = q[as] / P, bm = q[1] / P;
      int64_t last_bf = -1;
      int64_t sx[27], sy[27], sz[27], sc[27];
      int ns = 0;
      for (uint64_t pf = 0; pf < nf; ++pf) {
        q[af] = (int64_t)(pf + origin[af]);
        const int64_t bf = q[af] / P;
        if (bf != last_bf) { /* sites of the 27 neighbouring cells, in the order (dk, dj, di) of the definition */
          last_bf = bf;
          int64_t b[3];
          b[af] = bf; b[1] = bm; b[as] = bs;
          ns = 0;
          for (int dk = -1; dk <= 1; ++dk)
            for (int dj = -1; dj <= 1; ++dj)
              for (int di = -1; di <= 1; ++di) {
                const int64_t ni = b[0] + di, nj = b[1] + dj, nk = b[2] + dk;
                if (ni < 0 || nj < 0 || nk < 0 || ni >= G[0] || nj >= G[1] || nk >= G[2]) continue;
                const uint64_t c = (uint64_t)ni + (uint64_t)G[0] * ((uint64_t)nj + (uint64_t)G[1] * (uint64_t)nk);
                const uint64_t h = zo_splitmix64(c ^ seed);
                sx[ns] = ni * P + (int64_t)(((h & 0xFFFFull) * (uint64_t)P) >> 16);
                sy[ns] = nj * P + (int64_t)((((h >> 16) & 0xFFFFull) * (uint64_t)P) >> 16);
                sz[ns] = nk * P + (int64_t)((((h >> 32) & 0xFFFFull) * (uint64_t)P) >> 16);
                sc[ns] = (int64_t)c;
                ++ns;
              }
        }
        int64_t best_d = INT64_MAX, best_c = 0;
        for (int i = 0; i < ns; ++i) {
          const int64_t dx = q[0] - sx[i], dy = q[1] - sy[i], dz = q[2] - sz[i];
          const int64_t d = dx * dx + dy * dy + dz * dz;
          if (d < best_d || (d == best_d && sc[i] < best_c)) { best_d = d; best_c = sc[i]; }
        }
        const uint64_t lab = label_bytes == 8 ? (zo_splitmix64((uint64_t)best_c + 1ull) | 1ull) : (uint64_t)best_c + 1ull;
        const size_t i = ((size_t)ps * nm + pm) * nf + pf;
        switch (label_bytes) {
          case 1: ((uint8_t*)dst)[i] = (uint8_t)lab; break;
          case 2: ((uint16_t*)dst)[i] = (uint16_t)lab; break;
          case 4: ((uint32_t*)dst)[i] = (uint32_t)lab; break;
          default: ((uint64_t*)dst)[i] = lab; break;
        }
      }
    }
}

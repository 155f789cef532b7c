"""`Mesh`: the value type returned by `Mesher.get` -- counterpart of the reference's
zmesh/mesh.py:9-97 (container semantics) and :229-376 (Precomputed / OBJ / PLY codecs)."""
from __future__ import annotations

import re
import struct
from typing import Optional

import numpy as np

_PLY_HEADER = (
  "ply\nformat binary_little_endian 1.0\nelement vertex {nv}\nproperty float x\nproperty float y\n"
  "property float z\nelement face {nf}\nproperty list int int vertex_indices\nend_header\n"
)


_U32_MAX = int(np.iinfo(np.uint32).max)


class Mesh:
  """vertices float32 (Nv,3); faces uint32 (Nf,3) (uint64 only beyond 2^32 vertices);
  normals None or (Nv,3); id = label."""

  def __init__(self, vertices=None, faces=None, normals=None, id: Optional[int] = None):
    self.vertices = (np.zeros((0, 3), dtype=np.float32) if vertices is None
                     else np.asarray(vertices, dtype=np.float32))
    index_t = np.uint64 if self.vertices.shape[0] > _U32_MAX else np.uint32
    self.faces = np.zeros((0, 3), dtype=index_t) if faces is None else np.asarray(faces, dtype=index_t)
    self.normals = None if normals is None else np.asarray(normals, dtype=np.float32)
    self.id = id

  @classmethod
  def _wrap(cls, vertices: np.ndarray, faces: np.ndarray, normals, id) -> "Mesh":
    """Mesh around arrays that already have the final dtypes and shapes (Mesher.get's per-label views): skips
    the conversions of __init__, which dominate a get() loop over thousands of small labels."""
    m = cls.__new__(cls)
    m.vertices, m.faces, m.normals, m.id = vertices, faces, normals, id
    return m

  # -- container protocol ---------------------------------------------------------------------
  @property
  def segid(self):
    return self.id

  @segid.setter
  def segid(self, value):
    self.id = value

  def __len__(self) -> int:
    return int(self.vertices.shape[0])

  def _has_normals(self) -> bool:
    return self.normals is not None and self.normals.size > 0

  def __eq__(self, other) -> bool:
    if self._has_normals() != other._has_normals():
      return False
    same = bool(np.all(self.vertices == other.vertices)) and bool(np.all(self.faces == other.faces))
    if same and self._has_normals():
      same = bool(np.all(self.normals == other.normals))
    return same

  def __ne__(self, other) -> bool:
    return not self.__eq__(other)

  __hash__ = None

  def __repr__(self) -> str:
    nn = None if self.normals is None else self.normals.shape[0]
    return f"Mesh(vertices<{self.vertices.shape[0]}>, faces<{self.faces.shape[0]}>, normals<{nn}>)"

  def empty(self) -> bool:
    return self.faces.size == 0 or self.vertices.size == 0

  @property
  def nbytes(self) -> int:
    return sum(a.nbytes for a in (self.vertices, self.faces, self.normals) if a is not None)

  def clone(self) -> "Mesh":
    n = None if self.normals is None else self.normals.copy()
    return Mesh(self.vertices.copy(), self.faces.copy(), n, id=self.id)

  def triangles(self) -> np.ndarray:
    return self.vertices[self.faces]

  @classmethod
  def concatenate(cls, *meshes, id: Optional[int] = None) -> "Mesh":
    starts = np.concatenate([[0], np.cumsum([len(m) for m in meshes])]).astype(np.uint64)
    index_t = np.uint32 if starts[-1] < np.iinfo(np.uint32).max else np.uint64
    verts = np.concatenate([m.vertices for m in meshes])
    faces = np.concatenate([m.faces.astype(index_t, copy=False) + index_t(starts[i]) for i, m in enumerate(meshes)])
    return cls(verts, faces, None, id=id)

  # -- host-side clean-up utilities (zmesh/mesh.py:117-226; plain numpy, results identical to the reference's) ---------
  def remove_unreferenced_vertices(self) -> "Mesh":
    """Drop the vertices no face refers to and renumber the faces (normals are not kept, as in the reference)."""
    if self.empty():
      return Mesh([], [], None)
    used = np.zeros(len(self.vertices), dtype=bool)
    used[self.faces] = True
    new_index = np.cumsum(used) - 1
    return Mesh(self.vertices[used], new_index[self.faces], None)

  def remove_degenerate_faces(self) -> "Mesh":
    """Drop faces that name a vertex twice, then faces that repeat another one up to the order of their corners (the
    survivors come out in the lexicographic order of their sorted corner triples, as np.unique leaves them)."""
    if self.empty():
      return Mesh([], [], None)
    f = self.faces
    f = f[(f[:, 0] != f[:, 1]) & (f[:, 1] != f[:, 2]) & (f[:, 0] != f[:, 2])]
    _, first = np.unique(np.sort(f, axis=1), axis=0, return_index=True)
    return Mesh(self.vertices, f[first], self.normals, id=self.id)

  def consolidate(self) -> "Mesh":
    """Merge bit-identical vertices, drop duplicate and degenerate faces and unreferenced vertices; a new mesh."""
    if self.empty():
      return Mesh([], [], None)
    verts, rep, inverse = np.unique(self.vertices, axis=0, return_index=True, return_inverse=True)
    faces = np.unique(inverse.reshape(-1)[self.faces], axis=0)
    normals = self.normals[rep] if self._has_normals() else None
    return Mesh(verts, faces, normals, id=self.id).remove_degenerate_faces().remove_unreferenced_vertices()

  def merge_close_vertices(self, radius: float = 1e-5) -> "Mesh":
    """Merge vertices closer than `radius` (Euclidean); needs scipy."""
    from scipy.spatial import cKDTree
    if radius is None:
      radius = np.inf
    if radius <= 0:
      raise ValueError("radius must be greater than zero: " + str(radius))
    mesh = self.consolidate()
    pairs = cKDTree(mesh.vertices).query_pairs(r=radius, p=2, eps=0, output_type="ndarray")
    remap = np.arange(len(mesh.vertices), dtype=np.uint32)
    remap[pairs[:, 1]] = pairs[:, 0]
    mesh.faces = remap[mesh.faces]
    return mesh.consolidate()

  def dust(self, *args, **kwargs):
    raise NotImplementedError("zmesh_b200 covers mesh extraction only (connected components / dust are out of scope)")

  def largest_k(self, *args, **kwargs):
    raise NotImplementedError("zmesh_b200 covers mesh extraction only (connected components / largest_k are out of scope)")

  def save(self, filename: str):
    """Write a .ply (binary little endian) or, for any other extension, a Wavefront .obj (zmesh/mesh.py:423-436)."""
    with open(filename, "wb") as f:
      f.write(self.to_ply() if filename.endswith(".ply") else self.to_obj())

  @classmethod
  def load(cls, filename: str) -> "Mesh":
    """Read a .ply / .obj file written by save(), optionally gzip-compressed (.gz)."""
    import gzip
    name = filename
    if name.endswith(".gz"):
      with gzip.open(filename, "rb") as f:
        binary = f.read()
      name = name[:-3]
    else:
      with open(filename, "rb") as f:
        binary = f.read()
    if name.endswith(".ply"):
      return cls.from_ply(binary)
    if name.endswith(".obj"):
      return cls.from_obj(binary)
    raise ValueError(f"File format not supported: {filename}")

  # -- wire formats -----------------------------------------------------------------------------
  def to_precomputed(self) -> bytes:
    """Neuroglancer layout: uint32 Nv, Nv*3 float32, then uint32 face indices (no normals)."""
    return b"".join((struct.pack("<I", self.vertices.shape[0]),
                     np.ascontiguousarray(self.vertices, dtype="<f4").tobytes(),
                     np.ascontiguousarray(self.faces).astype("<u4", copy=False).tobytes()))

  @classmethod
  def from_precomputed(cls, binary: bytes) -> "Mesh":
    (nv,) = struct.unpack_from("<I", binary, 0)
    need = 4 + 12 * nv
    if len(binary) < need:
      raise ValueError(f"Precomputed mesh buffer too small: need >= {need} bytes, got {len(binary)}")
    verts = np.frombuffer(binary, dtype="<f4", count=3 * nv, offset=4).reshape(nv, 3)
    faces = np.frombuffer(binary, dtype="<u4", offset=need)
    return cls(verts, faces.reshape(-1, 3), None)

  def to_ply(self) -> bytes:
    out = bytearray(_PLY_HEADER.format(nv=self.vertices.shape[0], nf=self.faces.shape[0]).encode("utf8"))
    out += np.ascontiguousarray(self.vertices, dtype="<f4").tobytes()
    rec = np.empty((self.faces.shape[0], 4), dtype="<u4")
    rec[:, 0] = 3
    rec[:, 1:] = self.faces
    out += rec.tobytes()
    return out

  @classmethod
  def from_ply(cls, plydata: bytes) -> "Mesh":
    """Reads the binary little-endian PLY dialect written by `to_ply` (and by the reference)."""
    head, sep, body = bytes(plydata).partition(b"end_header\n")
    if not sep or not head.startswith(b"ply"):
      raise ValueError("not a binary PLY produced by to_ply")
    counts = dict(re.findall(rb"element (vertex|face) (\d+)", head))
    nv, nf = int(counts[b"vertex"]), int(counts[b"face"])
    verts = np.frombuffer(body, dtype="<f4", count=3 * nv).reshape(nv, 3)
    faces = np.frombuffer(body, dtype="<u4", count=4 * nf, offset=12 * nv).reshape(nf, 4)[:, 1:]
    return cls(verts, faces, None)

  def to_obj(self) -> bytes:
    lines = ["v %.5f %.5f %.5f" % tuple(v) for v in self.vertices]
    lines += ["f %d %d %d" % tuple(f) for f in (self.faces.astype(np.int64) + 1)]
    return ("\n".join(lines) + "\n").encode("utf8")

  @classmethod
  def from_obj(cls, text) -> "Mesh":
    if isinstance(text, bytes):
      text = text.decode("utf8")
    verts, faces, normals = [], [], []
    for raw in text.splitlines():
      tok = raw.split()
      if not tok or tok[0].startswith("#"):
        continue
      if tok[0] == "v":
        verts.append([float(t) for t in tok[1:4]])
      elif tok[0] == "vn":
        normals.append([float(t) for t in tok[1:4]])
      elif tok[0] == "f":
        faces.append([int(t.split("/")[0]) - 1 for t in tok[1:4]])
    return cls(np.array(verts, dtype=np.float32).reshape(-1, 3),
               np.array(faces, dtype=np.uint32).reshape(-1, 3),
               np.array(normals, dtype=np.float32).reshape(-1, 3))

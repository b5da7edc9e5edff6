"""ctypes binding of oracle/_build/libbmf_oracle.so -- the plain-C restatement (bmf_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
leg, never by the product package.  `build()` compiles it with /usr/bin/gcc (seconds).
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_build", "libbmf_oracle.so")

SPHERE, TORUS_Z, CUBOID, PLANE_Y, CSG = 0, 1, 2, 3, 4
TERRAIN2D, TERRAIN2D_PERT, TERRAIN3D, TERRAIN3D_PERT = 10, 11, 12, 13
HOST_DENSITY = 100
CSG_UNION, CSG_INTERSECT, CSG_SUBTRACT = 0, 1, 2


class Sampler(C.Structure):
    _fields_ = [("kind", C.c_int32), ("world_size", C.c_float), ("g_scale", C.c_float), ("height", C.c_float),
                ("octaves", C.c_int32), ("amp", C.c_float), ("frequency", C.c_float), ("gain", C.c_float),
                ("seed", C.c_int32), ("csg_op", C.c_int32), ("csg_kind_a", C.c_int32), ("csg_kind_b", C.c_int32),
                ("csg_world_size_a", C.c_float), ("csg_world_size_b", C.c_float),
                ("csg_offset_a", C.c_float * 3), ("csg_offset_b", C.c_float * 3)]


class Mesh(C.Structure):
    _fields_ = [("n_cells", C.c_int32), ("n_verts", C.c_int32), ("n_inds", C.c_int32),
                ("dense_inds", C.POINTER(C.c_uint32)), ("cell_masks", C.POINTER(C.c_uint8)),
                ("cell_grid", C.POINTER(C.c_uint32)), ("pos", C.POINTER(C.c_float)),
                ("boundary", C.POINTER(C.c_uint8)), ("valence", C.POINTER(C.c_uint8)), ("inds", C.POINTER(C.c_uint32))]


class SeamChunk(C.Structure):
    _fields_ = [("pos", C.c_float * 3), ("size", C.c_float), ("overlap", C.c_float), ("bits", C.c_void_p), ("density", C.c_void_p)]


def build(force=False):
    srcs = [os.path.join(HERE, f) for f in ("bmf_oracle.c", "bmf_oracle.h", "fastnoise_ref.h", "mc_tables_oracle.h")]
    if not force and os.path.exists(SO) and all(os.path.getmtime(SO) >= os.path.getmtime(s) for s in srcs):
        return SO
    subprocess.run(["make", "-C", HERE, "oracle"], check=True, stdout=subprocess.DEVNULL)
    return SO


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _np(ptr, n, dtype):
    if n == 0:
        return np.zeros(0, dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True)


class Oracle:
    def __init__(self):
        build()
        self.lib = lib = C.CDLL(SO)
        lib.orc_sampler_defaults.argtypes = [C.POINTER(Sampler), C.c_int]
        lib.orc_chunk_geometry.argtypes = [C.c_void_p, C.c_float, C.c_int, C.c_float, C.c_void_p, C.c_void_p]
        lib.orc_implicit_value.restype = C.c_float
        lib.orc_implicit_value.argtypes = [C.c_int, C.c_float, C.c_void_p]
        lib.orc_implicit_gradient.argtypes = [C.c_int, C.c_float, C.c_void_p, C.c_float, C.c_void_p]
        lib.orc_sample_block.argtypes = [C.POINTER(Sampler), C.c_void_p, C.c_float, C.c_int, C.c_void_p]
        lib.orc_label_grid.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        lib.orc_cell_masks.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        lib.orc_extract.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(Mesh)]
        lib.orc_mesh_free.argtypes = [C.POINTER(Mesh)]
        lib.orc_smooth.argtypes = [C.c_void_p] * 5 + [C.c_int, C.c_void_p] + [C.c_int] * 5
        lib.orc_qef_solve.restype = C.c_float
        lib.orc_qef_solve.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        lib.orc_qef_place.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int]
        lib.orc_qef_place_gradient.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.POINTER(Sampler), C.c_void_p, C.c_float, C.c_float]
        lib.orc_sampler_gradient.argtypes = [C.POINTER(Sampler), C.c_void_p, C.c_float, C.c_void_p]
        lib.orc_color_map.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        lib.orc_collapse_bad_quads.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        lib.orc_format_unwind.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.orc_quads.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(Mesh)]
        lib.orc_seam.restype = C.c_int64
        lib.orc_seam.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.POINTER(C.c_float))]
        lib.orc_free.argtypes = [C.c_void_p]
        lib.orc_batch.restype = C.c_int64
        lib.orc_batch.argtypes = [C.POINTER(Sampler), C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]

    def sampler(self, kind, **kw):
        s = Sampler()
        self.lib.orc_sampler_defaults(C.byref(s), kind)
        for k, v in kw.items():
            if k in ("csg_offset_a", "csg_offset_b"):
                setattr(s, k, (C.c_float * 3)(*v))
            else:
                setattr(s, k, v)
        return s

    def geometry(self, pos, size, dim, overlap):
        pos = np.asarray(pos, np.float32)
        op = np.zeros(3, np.float32)
        delta = C.c_float()
        self.lib.orc_chunk_geometry(_p(pos), size, dim, overlap, _p(op), C.byref(delta))
        return op, float(delta.value)

    def implicit_value(self, kind, p, world_size=256.0):
        p = np.asarray(p, np.float32)
        return float(self.lib.orc_implicit_value(kind, world_size, _p(p)))

    def implicit_gradient(self, kind, p, h=0.01, world_size=256.0):
        p = np.asarray(p, np.float32)
        out = np.zeros(3, np.float32)
        self.lib.orc_implicit_gradient(kind, world_size, _p(p), h, _p(out))
        return out

    def sampler_gradient(self, sampler, points, h=0.01):
        """Sampler::gradient at [m,3] world-space points -> [m,3] raw differences"""
        pts = np.ascontiguousarray(points, np.float32).reshape(-1, 3)
        out = np.zeros_like(pts)
        for i in range(len(pts)):
            self.lib.orc_sampler_gradient(C.byref(sampler), _p(pts[i]), h, _p(out[i]))
        return out

    def color_map(self, pos):
        """ColorMapper::generate_colors: [n,3] positions -> [n,3] colours"""
        pos = np.ascontiguousarray(pos, np.float32).reshape(-1, 3)
        col = np.zeros_like(pos)
        self.lib.orc_color_map(_p(pos), len(pos), _p(col))
        return col

    def collapse_bad_quads(self, pos, quads):
        """MeshProcessor<4>::init + collapse_bad_quads -> dict(pos, quads (rewired, all of them), destroyed, adj_next, bad_count, flushed)"""
        pos = np.array(pos, np.float32, copy=True).reshape(-1, 3)
        q = np.array(quads, np.uint32, copy=True).reshape(-1, 4)
        destroyed = np.zeros(len(q), np.uint8)
        adj_next = np.zeros(len(pos), np.uint8)
        bad = self.lib.orc_collapse_bad_quads(_p(pos), len(pos), _p(q), len(q), _p(destroyed), _p(adj_next))
        return {"pos": pos, "quads": q, "destroyed": destroyed, "adj_next": adj_next, "bad_count": int(bad), "flushed": q[destroyed == 0].copy()}

    def sample_block(self, sampler, overlap_pos, delta, dim):
        op = np.asarray(overlap_pos, np.float32)
        d = np.empty(dim ** 3, np.float32)
        rc = self.lib.orc_sample_block(C.byref(sampler), _p(op), delta, dim, _p(d))
        if rc:
            raise ValueError("orc_sample_block rc=%d" % rc)
        return d

    def label_grid(self, density, dim):
        density = np.ascontiguousarray(density, np.float32).reshape(-1)
        bits = np.zeros(dim * dim * ((dim + 31) // 32), np.uint32)
        cm = self.lib.orc_label_grid(_p(density), dim, _p(bits))
        return bits, bool(cm)

    def cell_masks(self, bits, dim):
        bits = np.ascontiguousarray(bits, np.uint32)
        masks = np.zeros(dim ** 3, np.uint8)
        self.lib.orc_cell_masks(_p(bits), dim, _p(masks))
        return masks

    def extract(self, density, masks, dim):
        density = np.ascontiguousarray(density, np.float32).reshape(-1)
        masks = np.ascontiguousarray(masks, np.uint8).reshape(-1)
        m = Mesh()
        self.lib.orc_extract(_p(density), _p(masks), dim, C.byref(m))
        out = {"n_cells": m.n_cells, "n_verts": m.n_verts, "n_inds": m.n_inds,
               "dense_inds": _np(m.dense_inds, dim ** 3, np.uint32), "cell_masks": _np(m.cell_masks, m.n_cells, np.uint8),
               "cell_grid": _np(m.cell_grid, m.n_cells, np.uint32), "pos": _np(m.pos, 3 * m.n_verts, np.float32).reshape(-1, 3),
               "boundary": _np(m.boundary, m.n_verts, np.uint8), "valence": _np(m.valence, m.n_verts, np.uint8),
               "inds": _np(m.inds, m.n_inds, np.uint32)}
        self.lib.orc_mesh_free(C.byref(m))
        return out

    def smooth(self, pos, color, normal, boundary, valence, inds, prim_n=3, iters=2, process_boundary=False, smooth_normals=False):
        pos = np.array(pos, np.float32, copy=True).reshape(-1, 3)
        color = np.array(color, np.float32, copy=True).reshape(-1, 3)
        normal = np.array(normal, np.float32, copy=True).reshape(-1, 3) if normal is not None else np.zeros_like(pos)
        boundary = np.ascontiguousarray(boundary, np.uint8)
        valence = np.ascontiguousarray(valence, np.uint8)
        inds = np.ascontiguousarray(inds, np.uint32)
        self.lib.orc_smooth(_p(pos), _p(color), _p(normal), _p(boundary), _p(valence), len(pos), _p(inds), len(inds), prim_n, iters,
                            int(process_boundary), int(smooth_normals))
        return pos, color, normal

    def qef_solve(self, positions, normals):
        p = np.ascontiguousarray(positions, np.float32)
        n = np.ascontiguousarray(normals, np.float32)
        out = np.zeros(4, np.float32)
        err = self.lib.orc_qef_solve(_p(p), _p(n), len(p), _p(out))
        return out[:3].copy(), float(err)

    def chunk(self, sampler, pos, size, dim, overlap=0.0, iters=0, process_boundary=False, smooth_normals=False, host_density=None, qef=False):
        """Full pipeline for one chunk through the stage functions (python-level composition)."""
        op, delta = self.geometry(pos, size, dim, overlap)
        if sampler.kind == HOST_DENSITY:
            density = np.ascontiguousarray(host_density, np.float32).reshape(-1)
        else:
            density = self.sample_block(sampler, op, delta, dim)
        bits, cm = self.label_grid(density, dim)
        out = {"overlap_pos": op, "scale": delta, "density": density, "bits": bits, "contains_mesh": cm,
               "n_cells": 0, "n_verts": 0, "n_inds": 0}
        if not cm:
            return out
        masks = self.cell_masks(bits, dim)
        out["masks"] = masks
        out.update(self.extract(density, masks, dim))
        nv = out["n_verts"]
        out["color"] = np.ones((nv, 3), np.float32)
        out["normal"] = np.zeros((nv, 3), np.float32)
        if iters > 0 and nv and out["n_inds"]:
            out["pos"], out["color"], out["normal"] = self.smooth(out["pos"], out["color"], out["normal"], out["boundary"], out["valence"],
                                                                  out["inds"], 3, iters, process_boundary, smooth_normals)
            if qef:
                p = np.ascontiguousarray(out["pos"], np.float32)
                if int(qef) == 2:
                    self.lib.orc_qef_place_gradient(_p(p), _p(out["boundary"]), _p(out["valence"]), nv, _p(out["inds"]), out["n_inds"], int(process_boundary),
                                                    C.byref(sampler), _p(np.asarray(op, np.float32)), delta, 0.01)
                else:
                    self.lib.orc_qef_place(_p(p), _p(out["boundary"]), _p(out["valence"]), nv, _p(out["inds"]), out["n_inds"], int(process_boundary))
                out["pos"] = p
        return out

    def quads(self, density, bits, dim):
        """dual-marching-cubes quad mesh of one chunk (build-defined): pos (grid units), boundary, valence, inds (4 per quad)"""
        m = Mesh()
        density = np.ascontiguousarray(density, np.float32)
        bits = np.ascontiguousarray(bits, np.uint32)
        self.lib.orc_quads(_p(density), _p(bits), dim, C.byref(m))
        out = {"n_cells": m.n_cells, "n_verts": m.n_verts, "n_inds": m.n_inds,
               "pos": _np(m.pos, 3 * m.n_verts, np.float32).reshape(-1, 3), "boundary": _np(m.boundary, m.n_verts, np.uint8),
               "valence": _np(m.valence, m.n_verts, np.uint8), "inds": _np(m.inds, m.n_inds, np.uint32)}
        self.lib.orc_mesh_free(C.byref(m))
        return out

    def format_unwind(self, pos, normal, color, inds, smooth_normals=False):
        """GLChunk::format_data(vertices, indexes, true, smooth_normals): flat quads, [n_inds, 3] p / n / c"""
        pos, normal, color = (np.ascontiguousarray(a, np.float32).reshape(-1, 3) for a in (pos, normal, color))
        inds = np.ascontiguousarray(inds, np.uint32)
        p, n, c = (np.zeros((len(inds), 3), np.float32) for _ in range(3))
        self.lib.orc_format_unwind(_p(pos), _p(normal), _p(color), _p(inds), len(inds), int(smooth_normals), _p(p), _p(n), _p(c))
        return p, n, c

    def seam(self, chunks, pos_size, dim, overlaps, group=None, cross_group_only=False):
        """seam pass over chunks = [self.chunk(...) dicts] (needs their "bits" and "density") -> [n_tris, 3, 3]"""
        ps = np.ascontiguousarray(pos_size, np.float32).reshape(-1, 4)
        ov = np.broadcast_to(np.asarray(overlaps, np.float32), (len(ps),))
        arr = (SeamChunk * len(ps))()
        keep = []
        for i, ch in enumerate(chunks):
            b = np.ascontiguousarray(ch["bits"], np.uint32)
            dn = np.ascontiguousarray(ch["density"], np.float32)
            keep += [b, dn]
            arr[i].pos[:] = [float(v) for v in ps[i, :3]]
            arr[i].size = float(ps[i, 3])
            arr[i].overlap = float(ov[i])
            arr[i].bits = b.ctypes.data_as(C.c_void_p)
            arr[i].density = dn.ctypes.data_as(C.c_void_p)
        g = None if group is None else np.ascontiguousarray(group, np.int32)
        out = C.POINTER(C.c_float)()
        n = self.lib.orc_seam(arr, len(ps), dim, _p(g), int(cross_group_only), C.byref(out))
        if n < 0:
            raise ValueError("orc_seam: chunks are not aligned octree leaves")
        tris = _np(out, 9 * n, np.float32).reshape(-1, 3, 3)
        self.lib.orc_free(out)
        return tris

    def batch(self, sampler, pos_size, dim, overlaps=None, iters=0, process_boundary=False, threads=0):
        ps = np.ascontiguousarray(pos_size, np.float32).reshape(-1, 4)
        ov = None if overlaps is None else np.ascontiguousarray(overlaps, np.float32)
        counts = np.zeros((len(ps), 2), np.int32)
        total = self.lib.orc_batch(C.byref(sampler), _p(ps), len(ps), dim, _p(ov), iters, int(process_boundary), threads, _p(counts))
        return int(total), counts

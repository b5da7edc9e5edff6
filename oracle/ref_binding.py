"""ctypes binding of oracle/_ref/libbmf_ref.so -- the UNMODIFIED reference hot path compiled from
/root/reference (see oracle/Makefile, oracle/ref_export.cpp).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; never from the product package.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libbmf_ref.so")

# sampler kinds (same numbering as include/bmf_b200.h)
SPHERE, TORUS_Z, CUBOID, PLANE_Y = 0, 1, 2, 3
TERRAIN2D, TERRAIN2D_PERT, TERRAIN3D, TERRAIN3D_PERT = 10, 11, 12, 13
HOST_DENSITY = 100

# DualVertex (Vertices.hpp:5-24), 84 bytes with g++ x86-64; offsets checked against the library at load
DUALVERTEX_DTYPE = np.dtype({
    "names": ["boundary", "mask", "index", "valence", "init_valence", "adj_next", "adj_offset", "edge_mask", "s", "xyz", "p", "n", "avg", "color"],
    "formats": ["u1", "u1", "<u4", "u1", "u1", "u1", "<u4", "<u2", "<f4", ("<i4", 3), ("<f4", 3), ("<f4", 3), ("<f4", 3), ("<f4", 3)],
    "offsets": [0, 1, 4, 8, 9, 10, 12, 16, 20, 24, 36, 48, 60, 72],
    "itemsize": 84,
})


def available():
    return os.path.exists(REF_SO)


def _fp(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class RefLib:
    def __init__(self, path=REF_SO):
        self.lib = lib = C.CDLL(path)
        lib.ref_chunk_create.restype = C.c_void_p
        lib.ref_chunk_create.argtypes = [C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_int, C.c_float, C.c_void_p]
        lib.ref_chunk_process.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        lib.ref_chunk_info.argtypes = [C.c_void_p] + [C.c_void_p] * 5
        lib.ref_chunk_copy.argtypes = [C.c_void_p] * 9
        lib.ref_chunk_destroy.argtypes = [C.c_void_p]
        lib.ref_mesh_process.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        lib.ref_format_unwind.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.ref_qef_solve.restype = C.c_float
        lib.ref_qef_solve.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        lib.ref_implicit_value.restype = C.c_float
        lib.ref_implicit_value.argtypes = [C.c_int, C.c_float, C.c_void_p]
        lib.ref_implicit_gradient.argtypes = [C.c_int, C.c_float, C.c_void_p, C.c_float, C.c_void_p]
        lib.ref_world_create.restype = C.c_void_p
        lib.ref_world_create.argtypes = [C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.ref_world_split_leaves.argtypes = [C.c_void_p]
        lib.ref_watcher_run.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        lib.ref_world_add_chunks.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        lib.ref_world_count.argtypes = [C.c_void_p]
        lib.ref_world_leaf.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.ref_world_process.restype = C.c_double
        lib.ref_world_process.argtypes = [C.c_void_p, C.c_int]
        lib.ref_world_chunk_info.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.ref_world_chunk_copy.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        assert lib.ref_sizeof_dualvertex() == DUALVERTEX_DTYPE.itemsize
        offs = np.zeros(10, np.int32)
        lib.ref_dualvertex_offsets(_fp(offs))
        assert list(offs) == [0, 4, 8, 9, 10, 12, 20, 36, 48, 72], offs

    # ---- single chunk -------------------------------------------------------------------------
    def chunk(self, kind, pos, size, dim, overlap=0.0, level=0, world_size=256.0, noise_props=None, host_density=None,
              iters=0, process_boundary=False, smooth_normals=False, want=("density", "bits", "masks", "cells", "verts", "inds")):
        """Run label_grid -> label_edges -> polygonize (-> MeshProcessor<3>) and return a dict of numpy arrays."""
        pos = np.asarray(pos, np.float32)
        props = None if noise_props is None else np.asarray(noise_props, np.float32)
        hd = None if host_density is None else np.ascontiguousarray(host_density, np.float32).reshape(-1)
        h = self.lib.ref_chunk_create(kind, world_size, _fp(props), _fp(pos), size, level, dim, overlap, _fp(hd))
        if not h:
            raise ValueError("unknown sampler kind %r" % (kind,))
        try:
            if iters > 0:
                self.lib.ref_chunk_process(h, iters, int(process_boundary), int(smooth_normals))
            cm, nc, nv, ni = (C.c_int() for _ in range(4))
            geom = np.zeros(4, np.float32)
            self.lib.ref_chunk_info(h, C.byref(cm), C.byref(nc), C.byref(nv), C.byref(ni), _fp(geom))
            n = dim ** 3
            out = {"contains_mesh": bool(cm.value), "n_cells": nc.value, "n_verts": nv.value, "n_inds": ni.value,
                   "overlap_pos": geom[:3].copy(), "scale": float(geom[3])}
            density = np.empty(n, np.float32) if "density" in want else None
            bits = np.empty(n // 32, np.uint32) if "bits" in want else None
            has = bool(cm.value)
            masks = np.zeros(n, np.uint8) if ("masks" in want and has) else None
            dense = np.empty(n, np.uint32) if ("cells" in want and has) else None
            cmask = np.empty(nc.value, np.uint8) if ("cells" in want and has) else None
            cgrid = np.empty(nc.value, np.uint32) if ("cells" in want and has) else None
            verts = np.zeros(nv.value, DUALVERTEX_DTYPE) if ("verts" in want and has) else None
            inds = np.empty(ni.value, np.uint32) if ("inds" in want and has) else None
            self.lib.ref_chunk_copy(h, _fp(density), _fp(bits), _fp(masks), _fp(dense), _fp(cmask), _fp(cgrid), _fp(verts), _fp(inds))
            out.update(density=density, bits=bits, masks=masks, dense_inds=dense, cell_masks=cmask, cell_grid=cgrid, verts=verts, inds=inds)
            return out
        finally:
            self.lib.ref_chunk_destroy(h)

    def mesh_process(self, verts, inds, prim_n=3, iters=2, process_boundary=False, smooth_normals=False):
        v = np.array(verts, DUALVERTEX_DTYPE, copy=True)
        i = np.array(inds, np.uint32, copy=True)
        self.lib.ref_mesh_process(_fp(v), len(v), _fp(i), len(i), prim_n, iters, int(process_boundary), int(smooth_normals))
        return v, i

    def collapse_bad_quads(self, verts, inds):
        """MeshProcessor<4>::init + collapse_bad_quads + flush: (DualVertex records after, surviving quads [m,4])"""
        v = np.array(verts, DUALVERTEX_DTYPE, copy=True)
        i = np.array(inds, np.uint32, copy=True).reshape(-1)
        self.lib.ref_collapse_bad_quads.restype = C.c_int
        n = self.lib.ref_collapse_bad_quads(_fp(v), len(v), _fp(i), len(i))
        return v, i[:n].reshape(-1, 4).copy()

    def color_map(self, verts):
        """ColorMapper::generate_colors: DualVertex records with .color filled"""
        v = np.array(verts, DUALVERTEX_DTYPE, copy=True)
        self.lib.ref_color_map(_fp(v), len(v))
        return v

    def format_unwind(self, verts, inds, smooth_normals=False):
        """GLChunk::format_data(vertices, indexes, true, smooth_normals): per quad corner p / n / c"""
        v = np.array(verts, DUALVERTEX_DTYPE, copy=True)
        i = np.array(inds, np.uint32, copy=True)
        p, n, c = (np.zeros((len(i), 3), np.float32) for _ in range(3))
        self.lib.ref_format_unwind(_fp(v), len(v), _fp(i), len(i), int(smooth_normals), _fp(p), _fp(n), _fp(c))
        return p, n, c

    def qef_solve(self, positions, normals):
        p = np.ascontiguousarray(positions, np.float32)
        n = np.ascontiguousarray(normals, np.float32)
        out = np.zeros(4, np.float32)  # the reference zeroes solved[3] too when count is out of range (qef_simd.h:558)
        err = self.lib.ref_qef_solve(_fp(p), _fp(n), len(p), _fp(out))
        return out[:3].copy(), float(err)

    def implicit_value(self, kind, p, world_size=256.0):
        p = np.asarray(p, np.float32)
        return float(self.lib.ref_implicit_value(kind, world_size, _fp(p)))

    def implicit_gradient(self, kind, p, h=0.01, world_size=256.0):
        p = np.asarray(p, np.float32)
        out = np.zeros(3, np.float32)
        self.lib.ref_implicit_gradient(kind, world_size, _fp(p), h, _fp(out))
        return out

    # ---- world / batch ------------------------------------------------------------------------
    def world(self, kind, dim, max_level=5, min_level=1, iters=0, boundary_processing=False, split_multiplier=1.0,
              size_modifier=0.0, overlap=0.035, focus=(0, 0, 0), world_size=256.0, noise_props=None):
        wp = np.array([max_level, min_level, iters, dim, int(boundary_processing)], np.int32)
        fp = np.array([split_multiplier, size_modifier, overlap, focus[0], focus[1], focus[2]], np.float32)
        props = None if noise_props is None else np.asarray(noise_props, np.float32)
        h = self.lib.ref_world_create(kind, world_size, _fp(props), _fp(wp), _fp(fp))
        return RefWorld(self, h, dim)


class RefWorld:
    def __init__(self, ref, handle, dim):
        self.ref, self.h, self.dim = ref, handle, dim

    def watcher_run(self, focus_per_tick, cap=1 << 16):
        """WorldWatcher ticks (check_leaves -> process_batch -> process_queue -> post_process_batch) from the root, one focus
        point per tick -> (Morton codes of the renderables in link order, chunks generated per tick)"""
        f = np.ascontiguousarray(focus_per_tick, np.float32).reshape(-1, 3)
        codes = np.zeros(cap, np.uint64)
        gen = np.zeros(len(f), np.int32)
        n = self.ref.lib.ref_watcher_run(self.h, _fp(f), len(f), _fp(codes), cap, _fp(gen))
        assert n <= cap
        return codes[:n].copy(), gen

    def split_leaves(self):
        return self.ref.lib.ref_world_split_leaves(self.h)

    def add_chunks(self, pos_size, levels):
        ps = np.ascontiguousarray(pos_size, np.float32).reshape(-1, 4)
        lv = np.ascontiguousarray(levels, np.int32)
        return self.ref.lib.ref_world_add_chunks(self.h, _fp(ps), _fp(lv), len(ps))

    def count(self):
        return self.ref.lib.ref_world_count(self.h)

    def leaves(self):
        n = self.count()
        ps = np.zeros((n, 4), np.float32)
        lv = np.zeros(n, np.int32)
        mc = np.zeros(n, np.uint64)
        for i in range(n):
            l = C.c_int()
            m = C.c_uint64()
            row = np.zeros(4, np.float32)
            self.ref.lib.ref_world_leaf(self.h, i, _fp(row), C.byref(l), C.byref(m))
            ps[i], lv[i], mc[i] = row, l.value, m.value
        return ps, lv, mc

    def process(self, threads=8):
        """ChunkGenerator::process_queue over the batch; returns milliseconds."""
        return float(self.ref.lib.ref_world_process(self.h, threads))

    def chunk(self, i):
        cm, nv, ni = C.c_int(), C.c_int(), C.c_int()
        self.ref.lib.ref_world_chunk_info(self.h, i, C.byref(cm), C.byref(nv), C.byref(ni))
        verts = np.zeros(nv.value, DUALVERTEX_DTYPE)
        inds = np.zeros(ni.value, np.uint32)
        p = np.zeros((nv.value, 3), np.float32)
        c = np.zeros((nv.value, 3), np.float32)
        if nv.value:
            self.ref.lib.ref_world_chunk_copy(self.h, i, _fp(verts), _fp(inds), _fp(p), _fp(c))
        return {"contains_mesh": bool(cm.value), "verts": verts, "inds": inds, "p_data": p, "c_data": c}

    def totals(self):
        nv = ni = nm = 0
        for i in range(self.count()):
            cm, v, k = C.c_int(), C.c_int(), C.c_int()
            self.ref.lib.ref_world_chunk_info(self.h, i, C.byref(cm), C.byref(v), C.byref(k))
            nv += v.value
            ni += k.value
            nm += 1 if (cm.value and v.value) else 0
        return nm, nv, ni


def fnv1a64(data, h=1469598103934665603):
    """FNV-1a-64 over raw bytes (the hash SURVEY Appendix A quotes)."""
    b = np.frombuffer(np.ascontiguousarray(data).tobytes(), np.uint8)
    # vectorising FNV is awkward; arrays here are small (<= a few MB)
    prime = 1099511628211
    mask = (1 << 64) - 1
    for byte in b.tolist():
        h ^= byte
        h = (h * prime) & mask
    return h

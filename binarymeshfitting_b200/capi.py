"""ctypes binding of libbmf_b200.so (include/bmf_b200.h) -- the harness-side view of the C ABI used by
tests/, bench.py and __graft_entry__.py.  Host buffers are numpy arrays; nothing here computes.

The library is the product; this file fails loudly when it is missing or when no CUDA device is
usable -- there is no CPU fallback anywhere in the package.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libbmf_b200.so")

SPHERE, TORUS_Z, CUBOID, PLANE_Y, CSG = 0, 1, 2, 3, 4
TERRAIN2D, TERRAIN2D_PERT, TERRAIN3D, TERRAIN3D_PERT = 10, 11, 12, 13
HOST_DENSITY = 100
CSG_UNION, CSG_INTERSECT, CSG_SUBTRACT = 0, 1, 2
STAGES = ("sample", "count", "scan", "verts", "inds", "smooth", "total")

# symbols include/bmf_b200.h declares (tests check that the library exports every one of them)
EXPORTS = ("bmf_ctx_create", "bmf_ctx_destroy", "bmf_last_error", "bmf_version", "bmf_sampler_defaults", "bmf_sampler_set",
           "bmf_batch_submit", "bmf_batch_wait", "bmf_batch_totals", "bmf_batch_chunk_info", "bmf_batch_chunk_infos",
           "bmf_batch_download", "bmf_batch_download_async", "bmf_batch_copy_chunk", "bmf_batch_stage_ms", "bmf_ctx_launch_count", "bmf_ctx_stream", "bmf_ctx_set_kernel_timing", "bmf_ctx_set_reserved_sms", "bmf_ctx_set_batches_in_flight", "bmf_ctx_kernel_times", "bmf_batch_device_ptrs",
           "bmf_mesh_process", "bmf_mesh_process_steps", "bmf_qef_solve", "bmf_sampler_gradient", "bmf_color_map", "bmf_mesh_collapse_bad_quads",
           "bmf_seam_overlap", "bmf_batch_stitch", "bmf_seam_download", "bmf_seam_stage_ms", "bmf_quads_to_tris", "bmf_batch_download_flat_quads", "bmf_ubench_issue",
           "bmf_batch_download_enqueue", "bmf_batch_download_dma", "bmf_host_alloc", "bmf_host_free", "bmf_host_register", "bmf_host_unregister")


class SamplerDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("world_size", C.c_float), ("g_scale", C.c_float), ("height", C.c_float),
                ("octaves", C.c_int32), ("amp", C.c_float), ("frequency", C.c_float), ("gain", C.c_float),
                ("seed", C.c_int32), ("csg_op", C.c_int32), ("csg_kind_a", C.c_int32), ("csg_kind_b", C.c_int32),
                ("csg_world_size_a", C.c_float), ("csg_world_size_b", C.c_float),
                ("csg_offset_a", C.c_float * 3), ("csg_offset_b", C.c_float * 3)]


class ChunkDesc(C.Structure):
    _fields_ = [("pos", C.c_float * 3), ("size", C.c_float), ("level", C.c_int32), ("overlap", C.c_float), ("morton", C.c_uint64)]


class Params(C.Structure):
    _fields_ = [("dim", C.c_int32), ("iters", C.c_int32), ("process_boundary", C.c_int32), ("smooth_normals", C.c_int32),
                ("qef", C.c_int32), ("keep_density", C.c_int32), ("keep_masks", C.c_int32), ("density_on_device", C.c_int32), ("quads", C.c_int32)]


class ChunkInfo(C.Structure):
    _fields_ = [("contains_mesh", C.c_int32), ("n_cells", C.c_int32), ("n_verts", C.c_int32), ("n_inds", C.c_int32),
                ("vert_offset", C.c_int64), ("ind_offset", C.c_int64), ("overlap_pos", C.c_float * 3), ("scale", C.c_float),
                ("flags", C.c_int32), ("reserved", C.c_int32)]


class DownloadDesc(C.Structure):
    _fields_ = [("pos", C.c_void_p), ("normal", C.c_void_p), ("color", C.c_void_p), ("boundary", C.c_void_p), ("valence", C.c_void_p),
                ("indices32", C.c_void_p), ("indices16", C.c_void_p), ("cap_verts", C.c_int64), ("cap_inds", C.c_int64)]


CHUNK_DESC_DTYPE = np.dtype([("pos", "<f4", 3), ("size", "<f4"), ("level", "<i4"), ("overlap", "<f4"), ("morton", "<u8")])
CHUNK_INFO_DTYPE = np.dtype([("contains_mesh", "<i4"), ("n_cells", "<i4"), ("n_verts", "<i4"), ("n_inds", "<i4"),
                             ("vert_offset", "<i8"), ("ind_offset", "<i8"), ("overlap_pos", "<f4", 3), ("scale", "<f4"),
                             ("flags", "<i4"), ("reserved", "<i4")])
CHUNK_COLOR_ONE, CHUNK_NORMAL_ZERO, CHUNK_INDEX16 = 1, 2, 4
assert CHUNK_DESC_DTYPE.itemsize == C.sizeof(ChunkDesc) and CHUNK_INFO_DTYPE.itemsize == C.sizeof(ChunkInfo)

DUALVERTEX_DTYPE = np.dtype({
    "names": ["boundary", "mask", "index", "valence", "init_valence", "adj_next", "adj_offset", "edge_mask", "s", "xyz", "p", "n", "avg", "color"],
    "formats": ["u1", "u1", "<u4", "u1", "u1", "u1", "<u4", "<u2", "<f4", ("<i4", 3), ("<f4", 3), ("<f4", 3), ("<f4", 3), ("<f4", 3)],
    "offsets": [0, 1, 4, 8, 9, 10, 12, 16, 20, 24, 36, 48, 60, 72],
    "itemsize": 84,
})


class BmfError(RuntimeError):
    pass


def load_library(path=SO):
    if not os.path.exists(path):
        raise BmfError("libbmf_b200.so is missing (%s): build it with `python -m binarymeshfitting_b200.build`; "
                       "there is no CPU fallback" % path)
    lib = C.CDLL(path)
    vp = C.c_void_p
    lib.bmf_ctx_create.argtypes = [C.c_int, C.POINTER(vp)]
    lib.bmf_ctx_destroy.argtypes = [vp]
    lib.bmf_ctx_destroy.restype = None
    lib.bmf_last_error.argtypes = [vp]
    lib.bmf_last_error.restype = C.c_char_p
    lib.bmf_version.restype = C.c_char_p
    lib.bmf_sampler_defaults.argtypes = [C.POINTER(SamplerDesc), C.c_int]
    lib.bmf_sampler_defaults.restype = None
    lib.bmf_sampler_set.argtypes = [vp, C.POINTER(SamplerDesc)]
    lib.bmf_batch_submit.argtypes = [vp, vp, C.c_int, C.POINTER(Params), vp]
    lib.bmf_batch_wait.argtypes = [vp]
    lib.bmf_batch_totals.argtypes = [vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    lib.bmf_batch_chunk_info.argtypes = [vp, C.c_int, C.POINTER(ChunkInfo)]
    lib.bmf_batch_chunk_infos.argtypes = [vp, vp]
    lib.bmf_batch_download.argtypes = [vp] * 7
    lib.bmf_batch_download_async.argtypes = [vp] * 7
    lib.bmf_batch_copy_chunk.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp]
    lib.bmf_batch_stage_ms.argtypes = [vp, vp]
    lib.bmf_ctx_launch_count.argtypes = [vp]
    lib.bmf_ctx_launch_count.restype = C.c_int64
    lib.bmf_ctx_set_kernel_timing.argtypes = [vp, C.c_int]
    lib.bmf_ctx_set_reserved_sms.argtypes = [vp, C.c_int]
    lib.bmf_ctx_set_batches_in_flight.argtypes = [vp, C.c_int]
    lib.bmf_ctx_kernel_times.argtypes = [vp, C.c_int, vp, vp]
    lib.bmf_ctx_stream.argtypes = [vp]
    lib.bmf_ctx_stream.restype = vp
    lib.bmf_batch_device_ptrs.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    lib.bmf_mesh_process.argtypes = [vp, vp, vp, vp, vp, vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.bmf_mesh_process_steps.argtypes = [vp, vp, vp, vp, vp, vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.bmf_qef_solve.argtypes = [vp, vp, vp, vp, C.c_int, vp, vp]
    lib.bmf_sampler_gradient.argtypes = [vp, vp, C.c_int64, C.c_float, vp]
    lib.bmf_color_map.argtypes = [vp, vp, C.c_int64, vp]
    lib.bmf_mesh_collapse_bad_quads.argtypes = [vp, vp, C.c_int, vp, C.c_int64, vp, vp, vp, vp, vp]
    lib.bmf_quads_to_tris.argtypes = [vp, vp, C.c_int64, vp]
    lib.bmf_batch_download_flat_quads.argtypes = [vp, C.c_int, vp, vp, vp]
    lib.bmf_ubench_issue.argtypes = [vp, vp]
    lib.bmf_seam_overlap.argtypes = [C.c_int]
    lib.bmf_seam_overlap.restype = C.c_float
    lib.bmf_batch_stitch.argtypes = [vp, vp, C.c_int, C.POINTER(C.c_int64)]
    lib.bmf_seam_download.argtypes = [vp, vp]
    lib.bmf_seam_stage_ms.argtypes = [vp, vp]
    lib.bmf_batch_download_enqueue.argtypes = [vp, C.POINTER(DownloadDesc)]
    lib.bmf_batch_download_dma.argtypes = [vp, C.POINTER(DownloadDesc)]
    lib.bmf_host_alloc.argtypes = [C.c_size_t, C.POINTER(vp)]
    lib.bmf_host_free.argtypes = [vp]
    lib.bmf_host_free.restype = None
    lib.bmf_host_register.argtypes = [vp, C.c_size_t]
    lib.bmf_host_unregister.argtypes = [vp]
    return lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def make_chunk_descs(pos_size, overlaps=0.0, levels=0):
    ps = np.asarray(pos_size, np.float32).reshape(-1, 4)
    d = np.zeros(len(ps), CHUNK_DESC_DTYPE)
    d["pos"] = ps[:, :3]
    d["size"] = ps[:, 3]
    d["overlap"] = overlaps
    d["level"] = levels
    d["morton"] = np.arange(1, len(ps) + 1)
    return d


class PinnedBuffer:
    """bmf_host_alloc'ed (page-locked, device-mapped) host memory viewed as a numpy array -- or, with `existing`, a numpy view of
    memory the caller owns (e.g. a shared-memory segment) registered with bmf_host_register."""

    def __init__(self, lib, nbytes=0, existing=None):
        self.lib, self.owned = lib, existing is None
        if existing is None:
            p = C.c_void_p()
            if lib.bmf_host_alloc(max(int(nbytes), 1), C.byref(p)) != 0 or not p:
                raise BmfError("bmf_host_alloc(%d) failed" % nbytes)
            self.ptr, self.nbytes = p.value, max(int(nbytes), 1)
            self.u8 = np.ctypeslib.as_array((C.c_uint8 * self.nbytes).from_address(self.ptr))
        else:
            self.u8 = existing.view(np.uint8).reshape(-1)
            self.ptr, self.nbytes = self.u8.ctypes.data, self.u8.nbytes
            if lib.bmf_host_register(C.c_void_p(self.ptr), self.nbytes) != 0:
                raise BmfError("bmf_host_register failed")

    def view(self, dtype, offset=0, count=None):
        dt = np.dtype(dtype)
        n = (self.nbytes - offset) // dt.itemsize if count is None else count
        return self.u8[offset:offset + n * dt.itemsize].view(dt)

    def close(self):
        if getattr(self, "ptr", None):
            self.u8 = None
            if self.owned:
                self.lib.bmf_host_free(C.c_void_p(self.ptr))
            else:
                self.lib.bmf_host_unregister(C.c_void_p(self.ptr))
            self.ptr = None


class Context:
    """One bmf_ctx (one GPU).  Mirrors the C ABI one to one."""

    def __init__(self, device=0, lib=None):
        self.lib = lib or load_library()
        h = C.c_void_p()
        rc = self.lib.bmf_ctx_create(device, C.byref(h))
        if rc != 0 or not h:
            raise BmfError("bmf_ctx_create(device=%d) failed with %d: no usable CUDA device (no CPU fallback)" % (device, rc))
        self.h = h
        self.n = 0
        self.dim = 0

    def close(self):
        if getattr(self, "h", None):
            self.lib.bmf_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise BmfError("bmf error %d: %s" % (rc, self.lib.bmf_last_error(self.h).decode()))

    def sampler_desc(self, kind, **kw):
        s = SamplerDesc()
        self.lib.bmf_sampler_defaults(C.byref(s), kind)
        for k, v in kw.items():
            if k in ("csg_offset_a", "csg_offset_b"):
                setattr(s, k, (C.c_float * 3)(*v))
            else:
                setattr(s, k, v)
        return s

    def set_sampler(self, kind_or_desc, **kw):
        s = kind_or_desc if isinstance(kind_or_desc, SamplerDesc) else self.sampler_desc(kind_or_desc, **kw)
        self._check(self.lib.bmf_sampler_set(self.h, C.byref(s)))
        return s

    def submit(self, descs, dim, iters=0, process_boundary=False, smooth_normals=False, qef=False, keep_density=False, keep_masks=False,
               density=None, density_device_ptr=None, quads=False):
        descs = np.ascontiguousarray(descs, CHUNK_DESC_DTYPE)
        p = Params(dim, iters, int(process_boundary), int(smooth_normals), int(qef), int(keep_density), int(keep_masks), 1 if density_device_ptr else 0, int(quads))
        dptr = None
        if density_device_ptr:
            dptr = C.c_void_p(density_device_ptr)
        elif density is not None:
            self._density_keepalive = np.ascontiguousarray(density, np.float32).reshape(-1)
            dptr = _p(self._density_keepalive)
        self._check(self.lib.bmf_batch_submit(self.h, _p(descs), len(descs), C.byref(p), dptr))
        self.n, self.dim = len(descs), dim

    def wait(self):
        self._check(self.lib.bmf_batch_wait(self.h))

    def totals(self):
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        self._check(self.lib.bmf_batch_totals(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def chunk_infos(self):
        out = np.zeros(self.n, CHUNK_INFO_DTYPE)
        self._check(self.lib.bmf_batch_chunk_infos(self.h, _p(out)))
        return out

    def download(self, want=("pos", "normal", "color", "boundary", "valence", "inds"), out=None, wait=True):
        _, V, I = self.totals()
        out = out or {}
        def buf(name, shape, dt):
            if name not in want:
                return None
            if name not in out:
                out[name] = np.empty(shape, dt)
            return out[name]
        pos = buf("pos", (V, 3), np.float32)
        nrm = buf("normal", (V, 3), np.float32)
        col = buf("color", (V, 3), np.float32)
        bnd = buf("boundary", (V,), np.uint8)
        val = buf("valence", (V,), np.uint8)
        ind = buf("inds", (I,), np.uint32)
        fn = self.lib.bmf_batch_download if wait else self.lib.bmf_batch_download_async
        self._check(fn(self.h, _p(pos), _p(nrm), _p(col), _p(bnd), _p(val), _p(ind)))
        return out

    def download_enqueue(self, pos=None, normal=None, color=None, boundary=None, valence=None, inds32=None, inds16=None):
        """bmf_batch_download_enqueue: numpy views of PINNED host memory (PinnedBuffer.view); returns at once, wait() completes it."""
        d = DownloadDesc()
        cap_v, cap_i = None, None
        for name, a, per in (("pos", pos, 3), ("normal", normal, 3), ("color", color, 3), ("boundary", boundary, 1), ("valence", valence, 1)):
            if a is not None:
                setattr(d, name, a.ctypes.data)
                cap_v = a.size // per if cap_v is None else min(cap_v, a.size // per)
        for name, a in (("indices32", inds32), ("indices16", inds16)):
            if a is not None:
                setattr(d, name, a.ctypes.data)
                cap_i = a.size
        d.cap_verts, d.cap_inds = cap_v or 0, cap_i or 0
        self._check(self.lib.bmf_batch_download_enqueue(self.h, C.byref(d)))

    def download_dma(self, pos=None, normal=None, color=None, boundary=None, valence=None, inds32=None, inds16=None):
        """bmf_batch_download_dma: waits for the batch's kernels, enqueues copy-engine transfers into (pinned) numpy views, returns; wait() completes them"""
        d = DownloadDesc()
        cap_v, cap_i = None, None
        for name, a, per in (("pos", pos, 3), ("normal", normal, 3), ("color", color, 3), ("boundary", boundary, 1), ("valence", valence, 1)):
            if a is not None:
                setattr(d, name, a.ctypes.data)
                cap_v = a.size // per if cap_v is None else min(cap_v, a.size // per)
        for name, a in (("indices32", inds32), ("indices16", inds16)):
            if a is not None:
                setattr(d, name, a.ctypes.data)
                cap_i = a.size
        d.cap_verts, d.cap_inds = cap_v or 0, cap_i or 0
        self._check(self.lib.bmf_batch_download_dma(self.h, C.byref(d)))

    def copy_chunk(self, i, want=("verts", "inds", "bits")):
        info = ChunkInfo()
        self._check(self.lib.bmf_batch_chunk_info(self.h, i, C.byref(info)))
        n = self.dim ** 3
        verts = np.zeros(info.n_verts, DUALVERTEX_DTYPE) if "verts" in want else None
        inds = np.zeros(info.n_inds, np.uint32) if "inds" in want else None
        bits = np.zeros(n // 32, np.uint32) if "bits" in want else None
        masks = np.zeros(n, np.uint8) if "masks" in want else None
        dens = np.zeros(n, np.float32) if "density" in want else None
        self._check(self.lib.bmf_batch_copy_chunk(self.h, i, _p(verts), _p(inds), _p(bits), _p(masks), _p(dens)))
        return {"contains_mesh": bool(info.contains_mesh), "n_cells": info.n_cells, "n_verts": info.n_verts, "n_inds": info.n_inds,
                "overlap_pos": np.array(info.overlap_pos[:], np.float32), "scale": info.scale,
                "verts": verts, "inds": inds, "bits": bits, "masks": masks, "density": dens}

    def stage_ms(self):
        ms = np.zeros(len(STAGES), np.float32)
        self._check(self.lib.bmf_batch_stage_ms(self.h, _p(ms)))
        return dict(zip(STAGES, ms.tolist()))

    def set_reserved_sms(self, n):
        """leave n SMs partly free for another context's kernels (overlapped pipelines)"""
        self._check(self.lib.bmf_ctx_set_reserved_sms(self.h, int(n)))

    def set_batches_in_flight(self, n):
        """hint: the caller keeps n batches in flight on this device (one context each): prefer the kernels with the least total SM time"""
        self._check(self.lib.bmf_ctx_set_batches_in_flight(self.h, int(n)))

    def set_kernel_timing(self, on):
        self._check(self.lib.bmf_ctx_set_kernel_timing(self.h, int(on)))

    def kernel_times(self, cap=512):
        """[(kernel name, ms)] for every launch of the last batch (needs set_kernel_timing(True) before submit)."""
        names = (C.c_char_p * cap)()
        ms = np.zeros(cap, np.float32)
        n = self.lib.bmf_ctx_kernel_times(self.h, cap, names, _p(ms))
        if n < 0:
            self._check(n)
        return [(names[i].decode().strip("()"), float(ms[i])) for i in range(min(n, cap))]  # (template kernels are launched as "(k<a, b>)")

    def device_ptrs(self):
        """device pointers (ints) of the resident batch: pos, indices, bits, density (0 if not materialised)"""
        p = [C.c_void_p() for _ in range(4)]
        self._check(self.lib.bmf_batch_device_ptrs(self.h, *[C.byref(x) for x in p]))
        return {k: int(v.value or 0) for k, v in zip(("pos", "indices", "bits", "density"), p)}

    def stream_ptr(self):
        return int(self.lib.bmf_ctx_stream(self.h) or 0)

    def launch_count(self):
        return int(self.lib.bmf_ctx_launch_count(self.h))

    def mesh_process(self, pos, color, normal, boundary, inds, prim_n=3, iters=2, process_boundary=False, smooth_normals=False):
        pos = np.array(pos, np.float32, copy=True).reshape(-1, 3)
        color = np.array(color, np.float32, copy=True).reshape(-1, 3)
        normal = None if normal is None else np.array(normal, np.float32, copy=True).reshape(-1, 3)
        boundary = np.ascontiguousarray(boundary, np.uint8)
        inds = np.ascontiguousarray(inds, np.uint32)
        self._check(self.lib.bmf_mesh_process(self.h, _p(pos), _p(color), _p(normal), _p(boundary), None, len(pos), _p(inds), len(inds), prim_n, iters,
                                              int(process_boundary), int(smooth_normals)))
        return pos, color, normal

    def quads_to_tris(self, quads):
        """MeshProcessor<4>::flush_to_tris: [n,4] quad indices -> [2n,3] triangle indices"""
        q = np.ascontiguousarray(quads, np.uint32).reshape(-1, 4)
        t = np.zeros((len(q) * 2, 3), np.uint32)
        self._check(self.lib.bmf_quads_to_tris(self.h, _p(q), len(q), _p(t)))
        return t

    def download_flat_quads(self, smooth_normals=False):
        """GLChunk::format_data(..., unwind_verts=True, smooth_normals) of the resident quad batch: [n_inds, 3] p / n / c"""
        self.wait()
        ni = self.totals()[2]
        p, n, c = (np.zeros((ni, 3), np.float32) for _ in range(3))
        self._check(self.lib.bmf_batch_download_flat_quads(self.h, int(smooth_normals), _p(p), _p(n), _p(c)))
        return p, n, c

    def ubench_issue(self):
        """measured issue peaks (1e9 chain steps/s): fp32 fma, int32 imad, int32 logic step, fma + logic step interleaved"""
        g = np.zeros(4, np.float32)
        self._check(self.lib.bmf_ubench_issue(self.h, _p(g)))
        return {"fp32_fma": float(g[0]), "int32_imad": float(g[1]), "int32_logic_step": float(g[2]), "fma_plus_logic_step": float(g[3])}

    def seam_overlap(self, dim):
        """the overlap that puts a chunk's samples at its voxel-node centres (what the seam pass expects)"""
        return float(self.lib.bmf_seam_overlap(dim))

    def stitch(self, group=None, cross_group_only=False, download=True):
        """seam pass over the resident batch -> [n_tris, 3, 3] world-space triangle soup (or the count)"""
        g = None if group is None else np.ascontiguousarray(group, np.int32)
        n = C.c_int64()
        self._check(self.lib.bmf_batch_stitch(self.h, _p(g), int(cross_group_only), C.byref(n)))
        if not download:
            return int(n.value)
        tris = np.zeros((int(n.value), 3, 3), np.float32)
        self._check(self.lib.bmf_seam_download(self.h, _p(tris)))
        return tris

    def seam_ms(self):
        ms = np.zeros(2, np.float32)
        self._check(self.lib.bmf_seam_stage_ms(self.h, _p(ms)))
        return {"count": float(ms[0]), "emit": float(ms[1])}

    def sampler_gradient(self, points, h=0.01):
        """Sampler::gradient of the current sampler at [m,3] world-space points -> [m,3] raw differences"""
        pts = np.ascontiguousarray(points, np.float32).reshape(-1, 3)
        out = np.zeros_like(pts)
        self._check(self.lib.bmf_sampler_gradient(self.h, _p(pts), len(pts), h, _p(out)))
        return out

    def color_map(self, pos):
        """ColorMapper::generate_colors: [n,3] positions -> [n,3] colours"""
        pos = np.ascontiguousarray(pos, np.float32).reshape(-1, 3)
        col = np.zeros_like(pos)
        self._check(self.lib.bmf_color_map(self.h, _p(pos), len(pos), _p(col)))
        return col

    def collapse_bad_quads(self, pos, quads):
        """MeshProcessor<4>::init + collapse_bad_quads (+ flush) -> dict(pos, quads (rewired), destroyed, adj_next, bad_count, flushed)"""
        pos = np.array(pos, np.float32, copy=True).reshape(-1, 3)
        q = np.array(quads, np.uint32, copy=True).reshape(-1, 4)
        destroyed = np.zeros(len(q), np.uint8)
        adj_next = np.zeros(len(pos), np.uint8)
        flushed = np.zeros_like(q)
        nf, bad = C.c_int64(), C.c_int64()
        self._check(self.lib.bmf_mesh_collapse_bad_quads(self.h, _p(pos), len(pos), _p(q), len(q), _p(destroyed), _p(adj_next), _p(flushed), C.byref(nf), C.byref(bad)))
        return {"pos": pos, "quads": q, "destroyed": destroyed, "adj_next": adj_next, "bad_count": int(bad.value), "flushed": flushed[:nf.value].copy()}

    def qef_solve(self, positions, normals, counts):
        """positions/normals: [m,12,3]; counts: [m]."""
        p = np.ascontiguousarray(positions, np.float32).reshape(-1, 12, 3)
        n = np.ascontiguousarray(normals, np.float32).reshape(-1, 12, 3)
        c = np.ascontiguousarray(counts, np.int32)
        out = np.zeros((len(c), 3), np.float32)
        err = np.zeros(len(c), np.float32)
        self._check(self.lib.bmf_qef_solve(self.h, _p(p), _p(n), _p(c), len(c), _p(out), _p(err)))
        return out, err

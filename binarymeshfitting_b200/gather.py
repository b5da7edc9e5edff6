"""Final host gather of per-chunk meshes (north_star / SURVEY 8(e)): one process per GPU, no data-path collective.

Every rank meshes its share of ONE batch (world.partition: contiguous ranges of the Z-curve) and its GPU stores the result
straight into this rank's region of a shared host segment (POSIX shared memory, page-locked and device-mapped in every
rank: bmf_batch_download_enqueue + bmf_host_register).  "Gathering" is therefore not a copy: when a rank's batch is
complete it publishes its chunk table next to the data and raises a flag; the gathering rank (0) waits for the flags of
all ranks and assembles the batch-order chunk table {chunk -> (vertex offset, index offset, counts, flags)} over the
segment -- the same thing ChunkGenerator::process_queue leaves behind (one VerticesIndicesBlock per chunk,
ChunkGenerator.cpp:27-60, WorldOctree.cpp:354-358), with the blocks living in one arena.

Two slots alternate so that rank r's GPU can fill slot (k+1)&1 while rank 0 still reads slot k&1.

Host logic only (numpy + mmap): works without CUDA, which is how the CPU tests drive it; `register` pins the rank's own
regions through the C ABI when a library handle is given.
"""
import mmap
import os
import time

import numpy as np

from . import capi

PAGE = 4096
HEADER_BYTES = PAGE


def _up(x, a=PAGE):
    return (int(x) + a - 1) // a * a


class RegionLayout:
    """byte layout of one rank's region inside a slot: positions | [normals] | [colours] | indices (u16 or u32) | chunk table"""

    def __init__(self, cap_verts, cap_inds, n_chunks, index_bytes=2, with_color=False, with_normal=False):
        self.cap_verts, self.cap_inds, self.n_chunks = int(cap_verts), int(cap_inds), int(n_chunks)
        self.index_bytes, self.with_color, self.with_normal = int(index_bytes), bool(with_color), bool(with_normal)
        o = 0
        self.pos = o
        o = _up(o + 12 * self.cap_verts, 256)
        self.normal = o if with_normal else None
        if with_normal:
            o = _up(o + 12 * self.cap_verts, 256)
        self.color = o if with_color else None
        if with_color:
            o = _up(o + 12 * self.cap_verts, 256)
        self.inds = o
        o = _up(o + self.index_bytes * self.cap_inds, 256)
        self.table = o
        o = o + capi.CHUNK_INFO_DTYPE.itemsize * self.n_chunks
        self.nbytes = _up(o)

    def as_tuple(self):
        return (self.cap_verts, self.cap_inds, self.n_chunks, self.index_bytes, self.with_color, self.with_normal)


class HostGather:
    SLOTS = 2

    def __init__(self, name, rank, world_size, layouts, create):
        """layouts: one RegionLayout per rank (every rank passes the same list).  Rank `create`=True makes the segment."""
        self.name, self.rank, self.world_size, self.layouts = name, rank, world_size, layouts
        self.region_off = []
        o = HEADER_BYTES
        for s in range(self.SLOTS):
            row = []
            for l in layouts:
                row.append(o)
                o += l.nbytes
            self.region_off.append(row)
        self.nbytes = o
        self.path = "/dev/shm/" + name
        if create:
            fd = os.open(self.path, os.O_RDWR | os.O_CREAT | os.O_TRUNC, 0o600)
            os.ftruncate(fd, self.nbytes)
        else:
            t0 = time.time()
            while not os.path.exists(self.path) or os.path.getsize(self.path) < self.nbytes:
                if time.time() - t0 > 60:
                    raise RuntimeError("shared segment %s did not appear" % self.path)
                time.sleep(0.01)
            fd = os.open(self.path, os.O_RDWR)
        self.mm = mmap.mmap(fd, self.nbytes)
        os.close(fd)
        self.u8 = np.frombuffer(self.mm, np.uint8)
        # header: consumed[SLOTS], done[SLOTS][world_size] (int64 step tickets)
        self.hdr = self.u8[:HEADER_BYTES].view(np.int64)
        self.created = create
        self.pins = []
        if create:
            self.hdr[:] = 0
        # first touch of this rank's own regions from this process (NUMA placement follows the toucher)
        for s in range(self.SLOTS):
            a = self.region_off[s][rank]
            self.u8[a:a + layouts[rank].nbytes:PAGE] = 0

    # ---- pinning (needs the CUDA library)
    def register(self, lib):
        for s in range(self.SLOTS):
            a = self.region_off[s][self.rank]
            self.pins.append(capi.PinnedBuffer(lib, existing=self.u8[a:a + self.layouts[self.rank].nbytes]))

    # ---- views
    def _view(self, slot, rank, off, dtype, count):
        a = self.region_off[slot][rank] + off
        dt = np.dtype(dtype)
        return self.u8[a:a + dt.itemsize * count].view(dt)

    def buffers(self, slot, rank=None):
        """numpy views of a rank's region (default: this rank's) for bmf_batch_download_enqueue"""
        r = self.rank if rank is None else rank
        l = self.layouts[r]
        out = {"pos": self._view(slot, r, l.pos, np.float32, 3 * l.cap_verts),
               "inds": self._view(slot, r, l.inds, np.uint16 if l.index_bytes == 2 else np.uint32, l.cap_inds),
               "table": self._view(slot, r, l.table, capi.CHUNK_INFO_DTYPE, l.n_chunks)}
        if l.with_color:
            out["color"] = self._view(slot, r, l.color, np.float32, 3 * l.cap_verts)
        if l.with_normal:
            out["normal"] = self._view(slot, r, l.normal, np.float32, 3 * l.cap_verts)
        return out

    # ---- the handshake (step tickets count from 1)
    def _done_index(self, slot, rank):
        return self.SLOTS + slot * self.world_size + rank

    def wait_slot_free(self, step):
        """before step `step` (0-based) reuses its slot: the gathering rank must have consumed step - SLOTS"""
        need = step - self.SLOTS + 1
        if need <= 0:
            return
        slot = step % self.SLOTS
        while self.hdr[slot] < need:
            pass

    def publish(self, step, infos):
        """this rank's batch `step` is complete in its region: write the chunk table, raise the flag"""
        slot = step % self.SLOTS
        self.buffers(slot)["table"][:] = infos
        self.hdr[self._done_index(slot, self.rank)] = step + 1

    def collect(self, step, parts, n_total):
        """gathering rank: wait for every rank's `step`, return the batch-order chunk table over the segment
        (vert_offset / ind_offset in ELEMENTS relative to the start of the owning rank's pos / index array; column `rank` says which)"""
        slot = step % self.SLOTS
        for r in range(self.world_size):
            i = self._done_index(slot, r)
            while self.hdr[i] < step + 1:
                pass
        table = np.zeros(n_total, capi.CHUNK_INFO_DTYPE)
        owner = np.zeros(n_total, np.int32)
        for r, idx in enumerate(parts):
            if len(idx):
                table[idx] = self.buffers(slot, r)["table"][:len(idx)]
                owner[idx] = r
        return table, owner

    def release(self, step):
        self.hdr[step % self.SLOTS] = step + 1

    def reset(self):
        """all ranks must be quiescent (call between two barriers)"""
        if self.created:
            self.hdr[:] = 0

    def chunk_arrays(self, slot, table, owner, i):
        """(positions [n,3] f32, indices) of chunk i of the gathered batch"""
        b = self.buffers(slot, int(owner[i]))
        v0, nv, i0, ni = int(table["vert_offset"][i]), int(table["n_verts"][i]), int(table["ind_offset"][i]), int(table["n_inds"][i])
        return b["pos"][3 * v0:3 * (v0 + nv)].reshape(-1, 3), b["inds"][i0:i0 + ni]

    def close(self):
        for p in self.pins:
            p.close()
        self.pins = []
        self.hdr = None
        self.u8 = None
        try:
            self.mm.close()
        except BufferError:
            pass
        if self.created:
            try:
                os.unlink(self.path)
            except OSError:
                pass

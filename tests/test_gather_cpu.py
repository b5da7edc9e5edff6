"""CPU: the multi-rank host gather (binarymeshfitting_b200/gather.py) over world_size-2 gloo.  The per-rank extraction is
the CPU oracle standing in for each rank's GPU -- this test is about the partition, the shared segment, the slot handshake
and the batch-order chunk table, i.e. everything bench.py --gpus N does around the kernels."""
import os
import socket
import subprocess
import sys

import numpy as np

from binarymeshfitting_b200 import capi, gather

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_region_layout_is_aligned_and_disjoint():
    l = gather.RegionLayout(1000, 5000, 7, index_bytes=2, with_color=True, with_normal=True)
    offs = [l.pos, l.normal, l.color, l.inds, l.table]
    assert all(o % 256 == 0 for o in offs) and offs == sorted(offs) and l.nbytes % 4096 == 0
    assert l.normal - l.pos >= 12000 and l.table - l.inds >= 10000 and l.nbytes >= l.table + 7 * capi.CHUNK_INFO_DTYPE.itemsize


def test_single_rank_roundtrip():
    lay = [gather.RegionLayout(100, 300, 4, index_bytes=4)]
    g = gather.HostGather("bmf_test_single_%d" % os.getpid(), 0, 1, lay, create=True)
    try:
        for step in range(5):
            g.wait_slot_free(step)
            b = g.buffers(step % 2)
            b["pos"][:30] = np.arange(30) + step
            b["inds"][:6] = np.arange(6) * 2
            infos = np.zeros(4, capi.CHUNK_INFO_DTYPE)
            infos["n_verts"] = [0, 10, 0, 0]
            infos["n_inds"] = [0, 6, 0, 0]
            g.publish(step, infos)
            table, owner = g.collect(step, [np.arange(4)], 4)
            p, idx = g.chunk_arrays(step % 2, table, owner, 1)
            assert p.shape == (10, 3) and p[0, 0] == step and idx.tolist() == [0, 2, 4, 6, 8, 10]
            g.release(step)
    finally:
        g.close()
    assert not os.path.exists("/dev/shm/bmf_test_single_%d" % os.getpid())


WORKER = r"""
import os, sys, zlib
sys.path.insert(0, %(root)r)
import numpy as np
import torch.distributed as dist
from binarymeshfitting_b200 import capi, gather, world as W
from oracle import oracle_binding as ob
dist.init_process_group("gloo")
rank, ws = dist.get_rank(), dist.get_world_size()
dim = 32
props = W.WorldProperties(max_level=4, chunk_resolution=dim)
ps, lv, mc = W.split_leaves(props)
n = len(ps)
parts = [np.sort(p) for p in W.partition(mc, np.ones(n), ws)]
mine = parts[rank]
O = ob.Oracle()
s = O.sampler(ob.SPHERE)
chunks = [O.chunk(s, ps[i][:3], ps[i][3], dim, 0.035) for i in mine]
V = sum(c["n_verts"] for c in chunks); I = sum(c["n_inds"] for c in chunks)
lay = [None] * ws
dist.all_gather_object(lay, (V + 8, I + 8, len(mine), 2, False, False))
lay = [gather.RegionLayout(*t) for t in lay]
g = gather.HostGather("bmf_test_gloo_%%d" %% os.getppid(), rank, ws, lay, create=(rank == 0))
dist.barrier()
ok = True
for step in range(4):            # more steps than slots: the handshake must recycle them
    g.wait_slot_free(step)
    b = g.buffers(step %% 2)
    infos = np.zeros(len(mine), capi.CHUNK_INFO_DTYPE)
    vo = io = 0
    for k, c in enumerate(chunks):   # what the rank's GPU does: store its batch into its region
        nv, ni = c["n_verts"], c["n_inds"]
        if nv:
            b["pos"][3 * vo:3 * (vo + nv)] = (c["pos"] + step).ravel()
            b["inds"][io:io + ni] = c["inds"]
        infos[k] = (int(c["contains_mesh"]), c["n_cells"], nv, ni, vo, io, (0, 0, 0), 0, 0, 0)
        vo += nv; io += ni
    g.publish(step, infos)
    if rank == 0:
        table, owner = g.collect(step, parts, n)
        ci = cp = 0
        for i in range(n):           # batch order
            p, idx = g.chunk_arrays(step %% 2, table, owner, i)
            cp = zlib.crc32(np.ascontiguousarray(p).tobytes(), cp); ci = zlib.crc32(idx.astype(np.uint32).tobytes(), ci)
        want_i = want_p = 0
        for i in range(n):
            c = O.chunk(s, ps[i][:3], ps[i][3], dim, 0.035)
            if c["n_verts"]:
                want_p = zlib.crc32(np.ascontiguousarray(c["pos"] + step).tobytes(), want_p); want_i = zlib.crc32(c["inds"].astype(np.uint32).tobytes(), want_i)
        ok = ok and (ci, cp) == (want_i, want_p) and int(table["n_verts"].sum()) > 0
        g.release(step)
dist.barrier()
if rank == 0:
    print("GATHER", int(ok), n, [len(p) for p in parts])
g.close()
dist.destroy_process_group()
"""


def test_two_rank_gather_over_gloo(tmp_path):
    script = tmp_path / "gather_worker.py"
    script.write_text(WORKER % {"root": ROOT})
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), str(script)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("GATHER")][0].split()
    assert int(line[1]) == 1 and int(line[2]) > 8

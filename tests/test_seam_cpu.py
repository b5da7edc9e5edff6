"""Seam pass, CPU side: the oracle's restatement of the build-defined dual-cell seam mesher.

The reference's WorldStitcher is non-functional as committed, so there is nothing to pin against; what CAN be
checked are properties: chunk meshes + seam form a closed, consistently oriented surface on worlds with and without LOD
changes, and the per-group + cross-group passes (the multi-GPU scheme) partition the full seam exactly.
"""
import os

import numpy as np
import pytest

import seam_util as su
from oracle import oracle_binding as ob

DIM = 32
WORLDS = [
    ("uniform_sphere", 2, (0.0, 0.0, 0.0), ob.SPHERE, 256.0),
    ("lod_sphere", 3, (90.0, 20.0, -30.0), ob.SPHERE, 700.0),
    ("lod_torus", 3, (60.0, 0.0, 0.0), ob.TORUS_Z, 600.0),
]


def build(oracle, max_level, focus, kind, ws):
    ps, lv, mc = su.lod_world(max_level, 1, focus, DIM)
    ov = su.seam_overlap(DIM)
    s = oracle.sampler(kind, world_size=ws)
    chunks = [oracle.chunk(s, p[:3], p[3], DIM, ov) for p in ps]
    return ps, lv, mc, ov, chunks


def tri_keys(tris):
    return sorted(np.ascontiguousarray(tris, np.float32).reshape(-1, 9).view(np.uint32).tolist())


@pytest.mark.parametrize("name,max_level,focus,kind,ws", WORLDS, ids=[w[0] for w in WORLDS])
def test_seam_closes_the_world(oracle, name, max_level, focus, kind, ws):
    ps, lv, mc, ov, chunks = build(oracle, max_level, focus, kind, ws)
    seam = oracle.seam(chunks, ps, DIM, ov)
    ct = su.world_triangles(chunks, DIM)
    eps = 1e-3 * float(ps[:, 3].min()) / DIM
    _, open_before, _ = su.edge_report(ct, eps)
    kept, open_after, nonmanifold = su.edge_report(np.concatenate([ct, seam.astype(np.float64)]), eps)
    assert len(seam) > 0 and open_before > 0       # the chunk meshes alone leave gaps
    assert open_after == 0 and nonmanifold == 0     # with the seam every directed edge has exactly one opposite
    if name != "uniform_sphere":
        assert len(np.unique(lv)) > 1               # the world really has LOD changes


def test_terrain_seams_leave_open_edges_only_at_the_world_border(oracle):
    """a heightfield never closes: after the seam pass the only open edges of a noise-terrain LOD world are where the surface
    leaves the world (or runs along its outermost voxel layer), nowhere between chunks"""
    ps, lv, mc = su.lod_world(3, 1, (40.0, -10.0, 25.0), DIM)
    ov = su.seam_overlap(DIM)
    s = oracle.sampler(ob.TERRAIN2D_PERT)
    chunks = [oracle.chunk(s, p[:3], p[3], DIM, ov) for p in ps]
    seam = oracle.seam(chunks, ps, DIM, ov)
    ct = su.world_triangles(chunks, DIM)
    assert len(seam) > 0 and len(np.unique(lv)) > 1
    coarse_voxel = float(ps[:, 3].max()) / DIM
    eps = 1e-3 * float(ps[:, 3].min()) / DIM
    _, open_before, _, pts_before = su.edge_report(ct, eps, return_open=True)
    _, open_after, nonmanifold, pts = su.edge_report(np.concatenate([ct, seam.astype(np.float64)]), eps, return_open=True)
    assert open_after < open_before and nonmanifold == 0
    lo, hi = ps[:, :3].min(axis=0).astype(np.float64), (ps[:, :3] + ps[:, 3:4]).max(axis=0).astype(np.float64)
    dist_to_border = np.minimum(pts - lo, hi - pts).min(axis=1)
    assert dist_to_border.max() <= 1.01 * coarse_voxel  # every open edge hugs the outside of the world
    inner_before = (np.minimum(pts_before - lo, hi - pts_before).min(axis=1) > 1.01 * coarse_voxel).sum()
    assert inner_before > 0  # while the chunk meshes alone are open all over the interior


def test_group_passes_partition_the_seam(oracle):
    name, max_level, focus, kind, ws = WORLDS[1]
    ps, lv, mc, ov, chunks = build(oracle, max_level, focus, kind, ws)
    full = oracle.seam(chunks, ps, DIM, ov)
    order = np.argsort(mc, kind="stable")
    group = np.zeros(len(ps), np.int32)
    for g, part in enumerate(np.array_split(order, 3)):
        group[part] = g
    parts = [oracle.seam(chunks, ps, DIM, ov, group=group, cross_group_only=True)]
    for g in range(3):
        idx = np.nonzero(group == g)[0]
        parts.append(oracle.seam([chunks[i] for i in idx], ps[idx], DIM, ov))
    assert all(len(p) > 0 for p in parts)
    assert tri_keys(np.concatenate(parts)) == tri_keys(full)


SEAM_WORKER = r"""
import os, sys
sys.path.insert(0, %(root)r)
sys.path.insert(0, os.path.join(%(root)r, "tests"))
import numpy as np
import torch.distributed as dist
import seam_util as su
from binarymeshfitting_b200 import world as W
from oracle import oracle_binding as ob   # CPU stand-in for each rank's GPU: this test is about the multi-rank seam scheme
dist.init_process_group("gloo")
rank, ws = dist.get_rank(), dist.get_world_size()
dim = 32
ps, lv, mc = su.lod_world(3, 1, (90.0, 20.0, -30.0), dim)
ov = su.seam_overlap(dim)
parts = W.partition(mc, np.ones(len(mc)), ws)
group = np.zeros(len(ps), np.int32)
for g, part in enumerate(parts):
    group[part] = g
O = ob.Oracle()
s = O.sampler(ob.SPHERE, world_size=700.0)
mine = np.sort(parts[rank])
chunks = {int(i): O.chunk(s, ps[i][:3], ps[i][3], dim, ov) for i in mine}
own = O.seam([chunks[int(i)] for i in mine], ps[mine], dim, ov)            # cells inside this rank's chunks
border = W.border_chunks(ps, group)
# the gathering rank needs the border chunks of every rank (sign words + border samples); chunks are a pure function
# of their descriptor, so it simply re-samples them -- no data-path collective
gathered = [None] * ws
dist.all_gather_object(gathered, own.tobytes())
if rank == 0:
    bch = [chunks[int(i)] if int(i) in chunks else O.chunk(s, ps[i][:3], ps[i][3], dim, ov) for i in border]
    cross = O.seam(bch, ps[border], dim, ov, group=group[border], cross_group_only=True)
    parts_t = [np.frombuffer(b, np.float32).reshape(-1, 3, 3) for b in gathered] + [cross]
    allc = [O.chunk(s, p[:3], p[3], dim, ov) for p in ps]
    full = O.seam(allc, ps, dim, ov)
    key = lambda t: sorted(np.ascontiguousarray(t, np.float32).reshape(-1, 9).view(np.uint32).tolist())
    print("SEAM", len(full), sum(len(t) for t in parts_t), len(cross), len(border), int(key(np.concatenate(parts_t)) == key(full)))
dist.barrier()
dist.destroy_process_group()
"""


def test_two_rank_seam_scheme_over_gloo(tmp_path):
    import socket
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "seam_worker.py"
    script.write_text(SEAM_WORKER % {"root": root})
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), str(script)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("SEAM")][0].split()
    full, parts, cross, border, same = (int(v) for v in line[1:])
    assert full == parts and cross > 0 and border > 0 and same == 1


def test_subset_whose_minimum_corner_is_a_fine_chunk(oracle):
    """A valid sub-range of an octree's leaves (one GPU's share, the border batch of the cross-rank pass) may have its minimum
    corner set by a fine chunk: the slot lattice is anchored to the octree (through the largest chunk), not to that corner."""
    ps = np.array([[0, -256, 0, 256.0], [0, 0, 0, 256.0], [-256, -256, -128, 128.0]], np.float32)
    s = oracle.sampler(ob.SPHERE, world_size=700.0)
    ov = su.seam_overlap(DIM)
    chunks = [oracle.chunk(s, p[:3], p[3], DIM, ov) for p in ps]
    tris = oracle.seam(chunks, ps, DIM, ov)
    assert len(tris) > 0
    from binarymeshfitting_b200 import world as W
    assert W.border_chunks(ps, np.array([0, 1, 2])).tolist() == [0, 1]  # the two coarse chunks share a face; the fine one touches neither


def test_random_morton_subranges_are_accepted(oracle):
    """every contiguous Z-curve range of a LOD world's leaves is a valid seam batch (fuzz of the alignment test only: dim 32,
    plane sampler far away so the pass itself is trivial)"""
    from binarymeshfitting_b200 import world as W
    rng = np.random.default_rng(5)
    s = oracle.sampler(ob.PLANE_Y)
    for focus in ((0.0, 0.0, 0.0), (150.0, 40.0, -60.0), (-200.0, 10.0, 90.0)):
        ps, lv, mc = su.lod_world(4, 1, focus, DIM)
        ps = ps.copy()
        ps[:, 1] += 4096.0  # lift the world far above the plane: every chunk is uniformly air
        order = W.morton_order(mc)
        cache = {}
        for _ in range(40):
            a = int(rng.integers(0, len(ps) - 1))
            b = int(rng.integers(a + 1, min(len(ps), a + 24) + 1))
            idx = order[a:b]
            for i in idx:
                if int(i) not in cache:
                    cache[int(i)] = oracle.chunk(s, ps[i][:3], ps[i][3], DIM, su.seam_overlap(DIM))
            oracle.seam([cache[int(i)] for i in idx], ps[idx], DIM, su.seam_overlap(DIM))  # raises ValueError if rejected


def test_misaligned_chunks_are_rejected(oracle):
    s = oracle.sampler(ob.SPHERE)
    ps = np.array([[0, 0, 0, 32.0], [40.0, 0, 0, 64.0]], np.float32)
    chunks = [oracle.chunk(s, p[:3], p[3], DIM, 0.0) for p in ps]
    with pytest.raises(ValueError):
        oracle.seam(chunks, ps, DIM, 0.0)

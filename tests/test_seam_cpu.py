"""Seam pass, CPU side: the oracle's restatement of the build-defined dual-cell seam mesher.

The reference's WorldStitcher is non-functional as committed, so there is nothing to pin against; what CAN be
checked are properties: chunk meshes + seam form a closed, consistently oriented surface on worlds with and without LOD
changes, and the per-group + cross-group passes (the multi-GPU scheme) partition the full seam exactly.
"""
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


def test_misaligned_chunks_are_rejected(oracle):
    s = oracle.sampler(ob.SPHERE)
    ps = np.array([[0, 0, 0, 32.0], [40.0, 0, 0, 64.0]], np.float32)
    chunks = [oracle.chunk(s, p[:3], p[3], DIM, 0.0) for p in ps]
    with pytest.raises(ValueError):
        oracle.seam(chunks, ps, DIM, 0.0)

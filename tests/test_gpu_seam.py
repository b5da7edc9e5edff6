"""GPU seam pass (bmf_batch_stitch, csrc/seam.cuh) against the oracle's restatement: bit-exact triangle soup in the
same order, on LOD worlds of every sampler family, plus the group filters the multi-GPU scheme uses."""
import numpy as np
import pytest

import seam_util as su
from binarymeshfitting_b200 import capi
from oracle import oracle_binding as ob

pytestmark = pytest.mark.gpu
DIM = 32


def run(gpu, oracle, kind, ps, dim=DIM, ov=None, group=None, cross=False, density=None, **kw):
    ov = su.seam_overlap(dim) if ov is None else ov
    s = oracle.sampler(kind, **kw)
    if kind == ob.HOST_DENSITY:
        chunks = [oracle.chunk(s, p[:3], p[3], dim, ov, host_density=density[i]) for i, p in enumerate(ps)]
    else:
        chunks = [oracle.chunk(s, p[:3], p[3], dim, ov) for p in ps]
    want = oracle.seam(chunks, ps, dim, ov, group=group, cross_group_only=cross)
    gpu.set_sampler(kind, **kw)
    gpu.submit(capi.make_chunk_descs(ps, overlaps=ov), dim, iters=0, density=density)
    got = gpu.stitch(group=group, cross_group_only=cross)
    return got, want, chunks


@pytest.mark.parametrize("kind,max_level,focus,ws", [(ob.SPHERE, 2, (0.0, 0.0, 0.0), 256.0), (ob.SPHERE, 3, (90.0, 20.0, -30.0), 700.0),
                                                     (ob.TORUS_Z, 3, (60.0, 0.0, 0.0), 600.0), (ob.CUBOID, 3, (10.0, 100.0, 0.0), 900.0)])
def test_seam_implicit_lod_worlds(gpu, oracle, kind, max_level, focus, ws):
    ps, lv, mc = su.lod_world(max_level, 1, focus)
    got, want, chunks = run(gpu, oracle, kind, ps, world_size=ws)
    assert len(want) > 0
    np.testing.assert_array_equal(got.view(np.uint32), want.view(np.uint32))
    # and the GPU's own output closes the GPU's own chunk meshes
    infos, out = gpu.chunk_infos(), gpu.download(want=("pos", "inds"))
    tris = []
    for i in range(len(ps)):
        v0, i0, nv, ni = int(infos[i]["vert_offset"]), int(infos[i]["ind_offset"]), int(infos[i]["n_verts"]), int(infos[i]["n_inds"])
        if ni:
            p = infos[i]["overlap_pos"].astype(np.float64)[None, :] + out["pos"][v0:v0 + nv].astype(np.float64) * float(infos[i]["scale"])
            tris.append(p[out["inds"][i0:i0 + ni].astype(np.int64)].reshape(-1, 3, 3))
    eps = 1e-3 * float(ps[:, 3].min()) / DIM
    _, open_edges, nonmanifold = su.edge_report(np.concatenate(tris + [got.astype(np.float64)]), eps)
    assert open_edges == 0 and nonmanifold == 0


@pytest.mark.parametrize("kind", [ob.TERRAIN2D_PERT, ob.TERRAIN2D])
def test_seam_terrain2d_with_uniform_chunks(gpu, oracle, kind):
    # chunks far above / below the heightfield never get sign words written: the seam pass must use their uniform flag
    ps, lv, mc = su.lod_world(3, 1, (40.0, -10.0, 25.0))
    got, want, _ = run(gpu, oracle, kind, ps)
    assert len(want) > 0
    np.testing.assert_array_equal(got.view(np.uint32), want.view(np.uint32))


def test_seam_terrain3d_and_dim64(gpu, oracle):
    ps = np.array([[x, y, z, 16.0] for x in (0.0, 16.0) for y in (-16.0, 0.0) for z in (0.0, 16.0)] + [[32.0, -16.0, 0.0, 32.0]], np.float32)
    got, want, _ = run(gpu, oracle, ob.TERRAIN3D_PERT, ps)
    assert len(want) > 0
    np.testing.assert_array_equal(got.view(np.uint32), want.view(np.uint32))
    got, want, _ = run(gpu, oracle, ob.SPHERE, ps, dim=64, world_size=100.0)
    assert len(want) > 0
    np.testing.assert_array_equal(got.view(np.uint32), want.view(np.uint32))


def test_seam_dim128_two_levels_apart(gpu, oracle):
    # a 128^3-voxel fine chunk next to chunks two levels coarser (the stretched-row classification with s = 2), dim 128
    ps = np.array([[0, 0, 0, 16.0], [16, 0, 0, 16.0], [0, 16, 0, 16.0], [16, 16, 0, 16.0], [0, 0, 16, 16.0], [16, 0, 16, 16.0], [0, 16, 16, 16.0],
                   [16, 16, 16, 16.0], [32, 0, 0, 32.0], [-64, 0, 0, 64.0], [0, -64, 0, 64.0], [-64, -64, 0, 64.0]], np.float32)
    got, want, _ = run(gpu, oracle, ob.SPHERE, ps, dim=128, world_size=150.0)
    assert len(want) > 0
    np.testing.assert_array_equal(got.view(np.uint32), want.view(np.uint32))
    got, want, _ = run(gpu, oracle, ob.TERRAIN2D_PERT, ps, dim=32)
    np.testing.assert_array_equal(got.view(np.uint32), want.view(np.uint32))


def test_seam_after_quads_and_after_smoothing(gpu, oracle):
    # the seam pass only reads sign words and samples: it does not depend on what the batch emitted
    ps, lv, mc = su.lod_world(2, 1, (0.0, 0.0, 0.0))
    s = oracle.sampler(ob.SPHERE)
    ov = su.seam_overlap(DIM)
    chunks = [oracle.chunk(s, p[:3], p[3], DIM, ov) for p in ps]
    want = oracle.seam(chunks, ps, DIM, ov)
    gpu.set_sampler(ob.SPHERE)
    for kw in ({"quads": True}, {"iters": 3}, {"iters": 2, "smooth_normals": True}):
        gpu.submit(capi.make_chunk_descs(ps, overlaps=ov), DIM, **kw)
        got = gpu.stitch()
        np.testing.assert_array_equal(got.view(np.uint32), want.view(np.uint32))


def test_seam_host_density(gpu, oracle):
    rng = np.random.default_rng(7)
    ps = np.array([[0, 0, 0, 8.0], [8, 0, 0, 8.0], [0, 8, 0, 8.0], [8, 8, 0, 8.0], [0, 0, 8, 8.0], [8, 0, 8, 8.0], [0, 8, 8, 8.0], [8, 8, 8, 8.0],
                   [16, 0, 0, 16.0]], np.float32)
    density = rng.standard_normal((len(ps), DIM ** 3)).astype(np.float32)
    got, want, _ = run(gpu, oracle, ob.HOST_DENSITY, ps, density=density)
    assert len(want) > 1000
    np.testing.assert_array_equal(got.view(np.uint32), want.view(np.uint32))


def test_seam_group_filters(gpu, oracle):
    ps, lv, mc = su.lod_world(3, 1, (90.0, 20.0, -30.0))
    order = np.argsort(mc, kind="stable")
    group = np.zeros(len(ps), np.int32)
    for g, part in enumerate(np.array_split(order, 2)):
        group[part] = g
    full, want_full, _ = run(gpu, oracle, ob.SPHERE, ps, world_size=700.0)
    cross, want_cross, _ = run(gpu, oracle, ob.SPHERE, ps, group=group, cross=True, world_size=700.0)
    np.testing.assert_array_equal(cross.view(np.uint32), want_cross.view(np.uint32))
    parts = [cross]
    for g in range(2):
        idx = np.nonzero(group == g)[0]
        got, want, _ = run(gpu, oracle, ob.SPHERE, ps[idx], world_size=700.0)
        np.testing.assert_array_equal(got.view(np.uint32), want.view(np.uint32))
        parts.append(got)
    key = lambda t: sorted(np.ascontiguousarray(t).reshape(-1, 9).view(np.uint32).tolist())
    assert key(np.concatenate(parts)) == key(full)


def test_seam_multi_gpu_scheme_with_border_batch(gpu, oracle):
    """what N ranks do: own seams per group + one cross-group pass over the border chunks only == the full seam"""
    from binarymeshfitting_b200 import world as W
    ps, lv, mc = su.lod_world(3, 1, (90.0, 20.0, -30.0))
    parts = W.partition(mc, np.ones(len(mc)), 4)
    group = np.zeros(len(ps), np.int32)
    for g, part in enumerate(parts):
        group[part] = g
    full, _, _ = run(gpu, oracle, ob.TORUS_Z, ps, world_size=600.0)
    pieces = []
    for g in range(4):
        idx = np.sort(parts[g])
        got, want, _ = run(gpu, oracle, ob.TORUS_Z, ps[idx], world_size=600.0)
        pieces.append(got)
    border = W.border_chunks(ps, group)
    assert 0 < len(border) < len(ps)
    got, want, _ = run(gpu, oracle, ob.TORUS_Z, ps[border], group=group[border], cross=True, world_size=600.0)
    np.testing.assert_array_equal(got.view(np.uint32), want.view(np.uint32))
    pieces.append(got)
    key = lambda t: sorted(np.ascontiguousarray(t).reshape(-1, 9).view(np.uint32).tolist())
    assert key(np.concatenate(pieces)) == key(full)


def test_seam_subset_whose_minimum_corner_is_a_fine_chunk(gpu, oracle):
    """a Morton sub-range of an octree's leaves whose minimum corner is a fine chunk (z offset 1 slot, extent 2) used to
    be rejected as misaligned; the lattice is now anchored to the octree.  Also a fuzz of contiguous Z-curve ranges."""
    from binarymeshfitting_b200 import world as W
    ps = np.array([[0, -256, 0, 256.0], [0, 0, 0, 256.0], [-256, -256, -128, 128.0]], np.float32)
    got, want, _ = run(gpu, oracle, ob.SPHERE, ps, world_size=700.0)
    assert len(want) > 0
    np.testing.assert_array_equal(got.view(np.uint32), want.view(np.uint32))
    rng = np.random.default_rng(11)
    lps, lv, mc = su.lod_world(4, 1, (150.0, 40.0, -60.0))
    order = W.morton_order(mc)
    gpu.set_sampler(capi.SPHERE, world_size=700.0)
    for _ in range(60):
        a = int(rng.integers(0, len(lps) - 1))
        b = int(rng.integers(a + 1, min(len(lps), a + 40) + 1))
        gpu.submit(capi.make_chunk_descs(lps[order[a:b]], overlaps=su.seam_overlap(DIM)), DIM, iters=0)
        gpu.stitch(download=False)  # raises BmfError if the range is rejected


def test_seam_errors(gpu):
    from binarymeshfitting_b200 import Context
    ctx = Context(0)
    with pytest.raises(capi.BmfError):
        ctx.stitch()  # no batch
    ctx.set_sampler(capi.SPHERE)
    ctx.submit(capi.make_chunk_descs([[0, 0, 0, 32.0], [40.0, 0, 0, 64.0]]), 32)
    with pytest.raises(capi.BmfError):
        ctx.stitch()  # not aligned leaves of one octree
    ctx.submit(capi.make_chunk_descs([[0, 0, 0, 32.0], [0, 0, 0, 32.0]]), 32)
    with pytest.raises(capi.BmfError):
        ctx.stitch()  # overlapping chunks
    ctx.submit(capi.make_chunk_descs([[0, 0, 0, 32.0]]), 32)
    assert ctx.stitch(download=False) == 0  # a single chunk has no neighbours
    ctx.close()

"""GPU: the one-launch per-chunk extraction (csrc/fused.cuh, dim <= 64) against the multi-kernel path (BMF_FUSED=0) that the
rest of the suite has pinned to the oracle / the compiled reference: every array of the batch must be bit-identical
(DMCChunk.cpp:440-498 cell order, :514-576 index order, MeshProcessor.cpp:98-128 adjacency order through bit-exact smoothing)."""
import os

import numpy as np
import pytest

from binarymeshfitting_b200 import Context, capi

pytestmark = pytest.mark.gpu


def make_ctx(fused):
    old = os.environ.get("BMF_FUSED")
    os.environ["BMF_FUSED"] = "2" if fused else "0"  # 2 = always (1, the default, takes the per-chunk kernels only for large batches)
    try:
        return Context(0)
    finally:
        if old is None:
            del os.environ["BMF_FUSED"]
        else:
            os.environ["BMF_FUSED"] = old


def grid(n, size, origin):
    return [[origin + size * i, origin + size * j, origin + size * k, size] for i in range(n) for j in range(n) for k in range(n)]


def run(ctx, sampler, ps, dim, overlap=0.045, density=None, **kw):
    ctx.set_sampler(sampler)
    ctx.submit(capi.make_chunk_descs(ps, overlaps=overlap), dim, density=density, **kw)
    ctx.wait()
    return ctx.chunk_infos(), ctx.download()


def same(a, b):
    ia, oa = a
    ib, ob = b
    for f in ("contains_mesh", "n_cells", "n_verts", "n_inds", "vert_offset", "ind_offset", "flags"):
        assert np.array_equal(ia[f], ib[f]), f
    for k in ("pos", "normal", "color"):
        assert np.array_equal(oa[k].view(np.uint32), ob[k].view(np.uint32)), k
    for k in ("boundary", "valence", "inds"):
        assert np.array_equal(oa[k], ob[k]), k
    return int(ia["n_verts"].sum())


CASES = [
    ("sphere64", capi.SPHERE, [[-128, -128, -128, 256.0]], 64, 0.0),
    ("torus32", capi.TORUS_Z, [[-128, -128, -128, 256.0]], 32, 0.0),
    ("cuboid64_grid", capi.CUBOID, grid(2, 128.0, -128.0), 64, 0.045),
    ("terrain2d_64", capi.TERRAIN2D_PERT, grid(4, 32.0, -64.0), 64, 0.045),
    ("terrain2d_32", capi.TERRAIN2D_PERT, grid(5, 24.0, -60.0), 32, 0.045),
    ("terrain3d_32", capi.TERRAIN3D_PERT, grid(3, 32.0, -48.0), 32, 0.045),
]


@pytest.mark.parametrize("name,sampler,ps,dim,overlap", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("iters", [0, 2, 5])
def test_fused_equals_multi_kernel(name, sampler, ps, dim, overlap, iters):
    f, m = make_ctx(True), make_ctx(False)
    nv = same(run(f, sampler, ps, dim, overlap, iters=iters), run(m, sampler, ps, dim, overlap, iters=iters))
    assert nv > 0
    # far fewer launches: that is the point of the kernel
    assert f.launch_count() < m.launch_count()
    f.close(); m.close()


@pytest.mark.parametrize("dim", [32, 64])
def test_fused_random_signs_masks_and_boundary_smoothing(dim):
    """random densities: almost every cell is active (stress for the counts, 16-bit group offsets and the CSR), mask image kept,
    boundary vertices processed, smooth normals on (per-step kernels on top of the fused CSR)"""
    rng = np.random.default_rng(dim)
    n = 3
    dens = rng.standard_normal((n, dim ** 3)).astype(np.float32)
    dens[1] = np.abs(dens[1]) + 1.0  # a chunk without a mesh between two that have one
    ps = [[0.0, 0.0, 64.0 * i, 64.0] for i in range(n)]
    f, m = make_ctx(True), make_ctx(False)
    kw = dict(iters=3, process_boundary=True, smooth_normals=True, keep_masks=True, density=dens)
    a, b = run(f, capi.HOST_DENSITY, ps, dim, 0.0, **kw), run(m, capi.HOST_DENSITY, ps, dim, 0.0, **kw)
    same(a, b)
    assert a[0]["contains_mesh"].tolist() == [1, 0, 1]
    for i in (0, 2):
        ca, cb = f.copy_chunk(i, want=("masks", "bits")), m.copy_chunk(i, want=("masks", "bits"))
        assert np.array_equal(ca["masks"], cb["masks"]) and np.array_equal(ca["bits"], cb["bits"])
    f.close(); m.close()


def test_fused_arena_growth_and_reuse():
    """first batch on a fresh context (nothing allocated: the kernel reports 'does not fit', the host sizes the arenas and re-launches),
    then a larger batch (arenas grow again), then a smaller one and an empty one on the same context"""
    f, m = make_ctx(True), make_ctx(False)
    for ps, sampler in ((grid(2, 32.0, -32.0), capi.TERRAIN2D_PERT), (grid(4, 32.0, -64.0), capi.TERRAIN2D_PERT), (grid(1, 256.0, -128.0), capi.SPHERE),
                        ([[0.0, 1000.0, 0.0, 16.0], [16.0, 1000.0, 0.0, 16.0]], capi.TERRAIN2D_PERT)):
        same(run(f, sampler, ps, 64, iters=2), run(m, sampler, ps, 64, iters=2))
    assert f.totals() == (0, 0, 0)
    f.close(); m.close()


def test_fused_many_chunks_lookback_order():
    """a batch with more mesh chunks than resident CTAs (look-back chains across waves) in a scrambled order"""
    ps = grid(16, 16.0, -128.0)
    rng = np.random.default_rng(7)
    ps = [ps[i] for i in rng.permutation(len(ps))]
    f, m = make_ctx(True), make_ctx(False)
    a, b = run(f, capi.TERRAIN2D_PERT, ps, 32, iters=2), run(m, capi.TERRAIN2D_PERT, ps, 32, iters=2)
    same(a, b)
    assert int(a[0]["contains_mesh"].sum()) > 2 * 148
    f.close(); m.close()


def test_tma_staging_equals_load_loop():
    """the per-chunk kernels stage the sign words / word counts with cp.async.bulk + mbarrier by default; BMF_TMA=0 is the 128-bit load loop"""
    ps = grid(6, 32.0, -96.0)
    outs = []
    for tma in ("1", "0"):
        old = os.environ.get("BMF_TMA")
        os.environ["BMF_TMA"] = tma
        try:
            c = make_ctx(True)
        finally:
            if old is None:
                del os.environ["BMF_TMA"]
            else:
                os.environ["BMF_TMA"] = old
        outs.append(run(c, capi.TERRAIN2D_PERT, ps, 64, iters=2))
        outs.append(run(c, capi.TERRAIN2D_PERT, ps, 32, iters=2))  # same context again: the barrier's phase keeps alternating
        c.close()
    assert same(outs[0], outs[2]) > 0 and same(outs[1], outs[3]) > 0

"""GPU quad emission (bmf_params.quads, csrc/quads.cuh) against the oracle twin: bit-exact vertices, indices, flags; the
MeshProcessor<4> pass on the emitted quads; flush_to_tris."""
import numpy as np
import pytest

from binarymeshfitting_b200 import capi
from oracle import oracle_binding as ob

pytestmark = pytest.mark.gpu


def check_batch(gpu, oracle, kind, ps, dim, iters=0, pb=False, overlap=0.0, density=None, **kw):
    gpu.set_sampler(kind, **kw)
    gpu.submit(capi.make_chunk_descs(ps, overlaps=overlap), dim, iters=iters, process_boundary=pb, quads=True, density=density)
    gpu.wait()
    infos, out = gpu.chunk_infos(), gpu.download()
    s = oracle.sampler(kind, **kw)
    total = 0
    for i, p in enumerate(ps):
        ch = oracle.chunk(s, p[:3], p[3], dim, overlap, host_density=None if density is None else density[i])
        if not ch["contains_mesh"]:
            assert infos[i]["n_verts"] == 0 and infos[i]["n_inds"] == 0
            continue
        q = oracle.quads(ch["density"], ch["bits"], dim)
        assert (int(infos[i]["n_cells"]), int(infos[i]["n_verts"]), int(infos[i]["n_inds"])) == (q["n_cells"], q["n_verts"], q["n_inds"])
        v0, i0, nv, ni = int(infos[i]["vert_offset"]), int(infos[i]["ind_offset"]), q["n_verts"], q["n_inds"]
        np.testing.assert_array_equal(out["inds"][i0:i0 + ni], q["inds"])
        np.testing.assert_array_equal(out["boundary"][v0:v0 + nv], q["boundary"])
        np.testing.assert_array_equal(out["valence"][v0:v0 + nv], q["valence"])
        want = q["pos"]
        if iters and nv and ni:
            want, _, _ = oracle.smooth(q["pos"], np.ones((nv, 3), np.float32), np.zeros((nv, 3), np.float32), q["boundary"], q["valence"], q["inds"], 4, iters, pb, False)
        np.testing.assert_array_equal(out["pos"][v0:v0 + nv].view(np.uint32), np.ascontiguousarray(want, np.float32).view(np.uint32))
        assert np.all(out["color"][v0:v0 + nv] == 1.0)
        total += nv
    assert total > 0
    return out


@pytest.mark.parametrize("kind", [ob.SPHERE, ob.TORUS_Z, ob.CUBOID])
@pytest.mark.parametrize("dim", [32, 64, 128])
def test_quads_config1_single_chunk(gpu, oracle, kind, dim):
    # BASELINE config 1: one chunk of an implicit primitive, quads, no processing
    check_batch(gpu, oracle, kind, np.array([[-128, -128, -128, 256.0]], np.float32), dim)


def test_quads_dim_256(gpu, oracle):
    check_batch(gpu, oracle, ob.TORUS_Z, np.array([[-128, -128, -128, 256.0]], np.float32), 256)


def test_quads_batch_terrain_and_smoothing(gpu, oracle):
    ps = np.array([[x, y, z, 32.0] for x in (-32.0, 0.0) for y in (-32.0, 0.0) for z in (-32.0, 0.0)], np.float32)
    check_batch(gpu, oracle, ob.TERRAIN2D_PERT, ps, 32, overlap=0.045)
    check_batch(gpu, oracle, ob.TERRAIN2D_PERT, ps, 32, iters=2, overlap=0.045)
    check_batch(gpu, oracle, ob.TERRAIN3D_PERT, ps[:2], 32, iters=3, pb=True, overlap=0.045)
    check_batch(gpu, oracle, ob.SPHERE, np.array([[-128, -128, -128, 256.0]], np.float32), 64, iters=4)


def test_quads_random_density_all_patch_cases(gpu, oracle):
    # random signs exercise every corner configuration, including the multi-patch ones
    rng = np.random.default_rng(11)
    ps = np.array([[0, 0, 0, 1.0], [1, 0, 0, 1.0]], np.float32)
    density = rng.standard_normal((2, 32 ** 3)).astype(np.float32)
    out = check_batch(gpu, oracle, ob.HOST_DENSITY, ps, 32, density=density)
    assert out["valence"].max() >= 4


@pytest.mark.parametrize("smooth", [False, True])
def test_flat_quads_format(gpu, oracle, smooth):
    """GLChunk::format_data(vertices, indexes, true, smooth_normals) on the resident quad batch (oracle pinned to the
    compiled reference in tests/test_oracle_vs_ref.py)"""
    ps = np.array([[x, -16.0, z, 32.0] for x in (-32.0, 0.0) for z in (-32.0, 0.0)], np.float32)
    gpu.set_sampler(ob.TERRAIN2D_PERT)
    gpu.submit(capi.make_chunk_descs(ps, overlaps=0.045), 32, iters=4 if smooth else 2, smooth_normals=smooth, quads=True)
    gpu.wait()
    infos, out = gpu.chunk_infos(), gpu.download()
    p, n, c = gpu.download_flat_quads(smooth)
    assert len(p) == gpu.totals()[2] > 0
    for i in range(len(ps)):
        v0, i0, nv, ni = int(infos[i]["vert_offset"]), int(infos[i]["ind_offset"]), int(infos[i]["n_verts"]), int(infos[i]["n_inds"])
        if not ni:
            continue
        wp, wn, wc = oracle.format_unwind(out["pos"][v0:v0 + nv], out["normal"][v0:v0 + nv], out["color"][v0:v0 + nv], out["inds"][i0:i0 + ni], smooth)
        np.testing.assert_array_equal(p[i0:i0 + ni].view(np.uint32), wp.view(np.uint32))
        np.testing.assert_array_equal(c[i0:i0 + ni].view(np.uint32), wc.view(np.uint32))
        np.testing.assert_array_equal(n[i0:i0 + ni].view(np.uint32), wn.view(np.uint32))


def test_quads_arena_growth_relaunch(oracle):
    """a fresh context meets a small quad batch, then a much larger one: the emitters are re-launched after the arenas grow"""
    from binarymeshfitting_b200 import Context
    ctx = Context(0)
    small = np.array([[x, y, z, 32.0] for x in (-32.0, 0.0) for y in (-32.0, 0.0) for z in (-32.0, 0.0)], np.float32)
    big = np.array([[x, y, z, 32.0] for x in (-64.0, -32.0, 0.0, 32.0) for y in (-64.0, -32.0, 0.0, 32.0) for z in (-64.0, -32.0, 0.0, 32.0)], np.float32)
    for ps in (small, big, small):
        check_batch(ctx, oracle, ob.TERRAIN2D_PERT, ps, 32, iters=2, overlap=0.045)
    ctx.close()


def test_flush_to_tris(gpu):
    q = np.arange(40, dtype=np.uint32).reshape(-1, 4)
    t = gpu.quads_to_tris(q)
    want = np.concatenate([q[:, [0, 1, 2]], q[:, [2, 3, 0]]], axis=1).reshape(-1, 3)
    np.testing.assert_array_equal(t, want)


def test_quads_reject_incompatible_options(gpu):
    gpu.set_sampler(capi.SPHERE)
    with pytest.raises(capi.BmfError):
        gpu.submit(capi.make_chunk_descs([[-128, -128, -128, 256.0]]), 32, quads=True, qef=True, iters=1)

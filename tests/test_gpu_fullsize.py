"""GPU, BASELINE.json's full sizes: the benchmark workload (4096 x 64^3, configs[2]) and the LOD world (configs[3])
against golden totals/checksums recorded from the compiled reference, plus size-independent properties."""
import json
import os
import zlib

import numpy as np
import pytest

from binarymeshfitting_b200 import capi, world as W

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "golden.json")))
ARR = np.load(os.path.join(HERE, "golden", "golden_arrays.npz"))


def mesh_properties(infos, out):
    """holds for any input: every index < its chunk's vertex count, index count multiple of 3, the valences of a
    chunk sum to its index count, no unreferenced vertex on the interior, boundary flags are 0/1"""
    vo, io = infos["vert_offset"], infos["ind_offset"]
    for i in np.flatnonzero(infos["n_inds"] > 0)[:: max(1, len(infos) // 200)]:
        nv, ni = int(infos["n_verts"][i]), int(infos["n_inds"][i])
        idx = out["inds"][io[i]:io[i] + ni]
        assert ni % 3 == 0 and int(idx.max()) < nv
        val = out["valence"][vo[i]:vo[i] + nv]
        assert int(val.astype(np.int64).sum()) == ni
        np.testing.assert_array_equal(np.bincount(idx, minlength=nv).astype(np.uint8), val)
    assert set(np.unique(out["boundary"]).tolist()) <= {0, 1}
    assert int(infos["n_verts"].sum()) == len(out["pos"]) and int(infos["n_inds"].sum()) == len(out["inds"])


def test_bench_workload_matches_reference_totals_and_checksums(gpu):
    g = GOLD["bench_workload"]
    ps = W.grid_chunks(16, 16.0)
    descs = capi.make_chunk_descs(ps, overlaps=np.float32(np.float32(0.035) + np.float32(0.005) * np.float32(2)))
    gpu.set_sampler(capi.TERRAIN2D_PERT)
    gpu.submit(descs, 64, iters=2)
    gpu.wait()
    infos = gpu.chunk_infos()
    out = gpu.download()
    assert (int(infos["n_verts"].sum()), int(infos["n_inds"].sum())) == (g["verts"], g["inds"])
    np.testing.assert_array_equal(np.stack([infos["n_verts"], infos["n_inds"]], axis=1), ARR["bench_counts"])
    # chunks are laid out in batch order, so the concatenated buffers hash like the reference's per-chunk sequence
    assert zlib.crc32(out["inds"].tobytes()) & 0xFFFFFFFF == g["inds_crc"]
    assert zlib.crc32(out["pos"].tobytes()) & 0xFFFFFFFF == g["pos_crc"]  # positions after 2 smoothing iterations, bit-exact
    mesh_properties(infos, out)


@pytest.mark.parametrize("w", GOLD["worlds"], ids=lambda w: w["key"])
def test_lod_world_matches_reference(gpu, w):
    props = W.WorldProperties(max_level=w["max_level"], chunk_resolution=w["dim"], process_iters=w["iters"])
    ps, lv, mc = W.split_leaves(props, 256, tuple(w["focus"]))
    gpu.set_sampler(w["kind"])
    gpu.submit(W.make_descs(props, ps, lv, mc), w["dim"], iters=w["iters"])
    gpu.wait()
    infos = gpu.chunk_infos()
    out = gpu.download()
    np.testing.assert_array_equal(np.stack([infos["n_verts"], infos["n_inds"]], axis=1), ARR[w["key"] + "_counts"])
    assert (int(infos["n_verts"].sum()), int(infos["n_inds"].sum())) == (w["verts"], w["inds"])
    assert zlib.crc32(out["inds"].tobytes()) & 0xFFFFFFFF == w["inds_crc"]
    mesh_properties(infos, out)


def test_idempotent_resubmit_and_arena_reuse(gpu):
    """submitting a small batch after a large one (and back) reuses the arenas and reproduces the same bytes"""
    gpu.set_sampler(capi.TERRAIN3D_PERT)
    small = capi.make_chunk_descs(W.grid_chunks(2, 32.0, origin=(-32, -32, -32)), overlaps=0.045)
    big = capi.make_chunk_descs(W.grid_chunks(6, 32.0, origin=(-96, -96, -96)), overlaps=0.045)
    crcs = []
    for d in (small, big, small):
        gpu.submit(d, 32, iters=2)
        gpu.wait()
        o = gpu.download()
        crcs.append((zlib.crc32(o["inds"].tobytes()), zlib.crc32(o["pos"].tobytes()), gpu.totals()))
    assert crcs[0] == crcs[2] and crcs[1][2][1] > crcs[0][2][1]


def test_growing_batches_all_samplers(gpu):
    """arenas (device and pinned) grow across submits of increasing size for every sampler family; results of the
    small batch are reproduced afterwards"""
    for kind in (capi.TERRAIN2D_PERT, capi.SPHERE, capi.TERRAIN3D):
        gpu.set_sampler(kind)
        first = None
        for n in (2, 5, 9, 2):
            d = capi.make_chunk_descs(W.grid_chunks(n, 256.0 / n), overlaps=0.045)
            gpu.submit(d, 32, iters=2)
            gpu.wait()
            o = gpu.download()
            sig = (gpu.totals(), zlib.crc32(o["inds"].tobytes()), zlib.crc32(o["pos"].tobytes()))
            if n == 2 and first is None:
                first = sig
        assert sig == first


def test_per_chunk_smoothing_equals_per_step_kernels(gpu):
    """k_smooth_chunks (all half-steps of a chunk in one CTA) against the per-step kernels k_dual / k_primal (BMF_SMOOTH_FUSED=0), NaN-normal quirk included"""
    import os
    from binarymeshfitting_b200 import Context, capi, world
    ps = world.grid_chunks(8, 16.0, origin=(-64.0, -48.0, -64.0))  # 512 chunks: enough for the fused path
    descs = capi.make_chunk_descs(ps, overlaps=0.045)
    gpu.set_sampler(capi.TERRAIN2D_PERT)
    gpu.submit(descs, 64, iters=3)
    gpu.wait()
    want = gpu.download(want=("pos", "inds", "normal"))
    # the per-step kernels (the reference's own structure) on the same batch: same positions AND the same normals, including
    # the NaN the reference leaves in every processed vertex when iters is 2, 3 or >= 5 and smooth normals are off
    os.environ["BMF_SMOOTH_FUSED"] = "0"
    try:
        ctx0 = Context(0)
    finally:
        del os.environ["BMF_SMOOTH_FUSED"]
    ctx0.set_sampler(capi.TERRAIN2D_PERT)
    for iters in (1, 2, 3, 4, 5):
        ctx0.submit(descs, 64, iters=iters)
        ctx0.wait()
        a = ctx0.download(want=("pos", "normal"))
        gpu.submit(descs, 64, iters=iters)
        gpu.wait()
        b = gpu.download(want=("pos", "normal"))
        np.testing.assert_array_equal(a["pos"].view(np.uint32), b["pos"].view(np.uint32))
        np.testing.assert_array_equal(a["normal"], b["normal"])
        assert bool(np.isnan(b["normal"]).any()) == (iters in (2, 3, 5))
    ctx0.close()
    assert len(want["pos"]) > 100000


@pytest.mark.parametrize("kind", ["TERRAIN2D_PERT", "TERRAIN3D_PERT", "TORUS_Z"])
def test_batches_in_flight_hint_changes_kernels_not_results(gpu, kind):
    """bmf_ctx_set_batches_in_flight(n > 1) picks the per-chunk kernels for a batch of any size (for 2-D terrains: sign words made inside
    k_chunk_count); everything a caller can read back must be identical to the latency-first kernels' output, sign words included."""
    from binarymeshfitting_b200 import world
    props = world.WorldProperties(max_level=4, chunk_resolution=64, process_iters=2)
    lps, lv, mc = world.split_leaves(props)
    descs = capi.make_chunk_descs(lps, overlaps=0.045, levels=lv)
    assert len(descs) < 8 * 100  # small enough for the per-segment kernels by default
    gpu.set_sampler(getattr(capi, kind))
    got = []
    for hint in (1, 4, 1):
        gpu.set_batches_in_flight(hint)
        gpu.set_kernel_timing(True)
        gpu.submit(descs, 64, iters=2)
        gpu.wait()
        names = [n for n, _ in gpu.kernel_times()]
        gpu.set_kernel_timing(False)
        assert any("k_chunk_emit" in n for n in names) == (hint > 1), names
        infos = gpu.chunk_infos()
        out = gpu.download()
        mesh = np.flatnonzero(infos["n_verts"] > 0)
        bits = [gpu.copy_chunk(int(i), want=("bits",))["bits"] for i in list(mesh[:3]) + [0, len(descs) - 1]]
        got.append((infos, out, bits))
    gpu.set_batches_in_flight(1)
    assert int(got[0][0]["n_verts"].sum()) > 10000
    for infos, out, bits in got[1:]:
        for f in ("contains_mesh", "n_cells", "n_verts", "n_inds", "vert_offset", "ind_offset", "flags"):
            np.testing.assert_array_equal(infos[f], got[0][0][f])
        for k in ("pos", "normal", "color"):
            np.testing.assert_array_equal(out[k].view(np.uint32), got[0][1][k].view(np.uint32))
        for k in ("inds", "boundary", "valence"):
            np.testing.assert_array_equal(out[k], got[0][1][k])
        for a, b in zip(bits, got[0][2]):
            np.testing.assert_array_equal(a, b)


def test_mode_switching_on_one_context_keeps_every_result_right(gpu):
    """one context, many kinds of batches back to back (triangles / quads / seams / smoothing variants / host density): no state
    of an earlier batch (colour fill, published chunk table, quad flags, arena contents) may leak into a later one"""
    from oracle import oracle_binding as ob
    orc = ob.Oracle()
    ps = np.array([[x, y, z, 32.0] for x in (-32.0, 0.0) for y in (-32.0, 0.0) for z in (-32.0, 0.0)], np.float32)
    rng = np.random.default_rng(3)
    dens = rng.standard_normal((len(ps), 32 ** 3)).astype(np.float32)

    def tri_check(kind, iters, sn=False, density=None, **kw):
        gpu.set_sampler(kind, **kw)
        gpu.submit(capi.make_chunk_descs(ps, overlaps=0.045), 32, iters=iters, smooth_normals=sn, density=density)
        gpu.wait()
        infos, out = gpu.chunk_infos(), gpu.download()
        s = orc.sampler(kind, **kw)
        for i, p in enumerate(ps):
            o = orc.chunk(s, p[:3], p[3], 32, 0.045, iters=iters, smooth_normals=sn, host_density=None if density is None else density[i])
            assert (int(infos[i]["n_verts"]), int(infos[i]["n_inds"])) == (o["n_verts"], o["n_inds"])
            if o["n_verts"]:
                v0, i0 = int(infos[i]["vert_offset"]), int(infos[i]["ind_offset"])
                np.testing.assert_array_equal(out["inds"][i0:i0 + o["n_inds"]], o["inds"])
                np.testing.assert_array_equal(out["pos"][v0:v0 + o["n_verts"]].view(np.uint32), o["pos"].view(np.uint32))
                np.testing.assert_array_equal(out["color"][v0:v0 + o["n_verts"]].view(np.uint32), o["color"].view(np.uint32))
                # NaN-aware: with smooth normals off the reference leaves normalize(0) = NaN in every processed vertex for some iteration counts
                np.testing.assert_array_equal(out["normal"][v0:v0 + o["n_verts"]], o["normal"])

    def quad_check(kind, iters):
        gpu.set_sampler(kind)
        gpu.submit(capi.make_chunk_descs(ps, overlaps=0.045), 32, iters=iters, quads=True)
        gpu.wait()
        infos, out = gpu.chunk_infos(), gpu.download()
        s = orc.sampler(kind)
        for i, p in enumerate(ps):
            ch = orc.chunk(s, p[:3], p[3], 32, 0.045)
            if not ch["contains_mesh"]:
                continue
            q = orc.quads(ch["density"], ch["bits"], 32)
            nv, ni = q["n_verts"], q["n_inds"]
            want = q["pos"]
            if iters and nv and ni:
                want, _, _ = orc.smooth(q["pos"], np.ones((nv, 3), np.float32), np.zeros((nv, 3), np.float32), q["boundary"], q["valence"], q["inds"], 4, iters, False, False)
            v0, i0 = int(infos[i]["vert_offset"]), int(infos[i]["ind_offset"])
            np.testing.assert_array_equal(out["inds"][i0:i0 + ni], q["inds"])
            np.testing.assert_array_equal(out["pos"][v0:v0 + nv].view(np.uint32), np.ascontiguousarray(want, np.float32).view(np.uint32))

    tri_check(capi.TERRAIN2D_PERT, 2)
    quad_check(capi.TERRAIN2D_PERT, 2)
    tri_check(capi.TERRAIN2D_PERT, 4, sn=True)
    tri_check(capi.HOST_DENSITY, 1, density=dens)
    quad_check(capi.SPHERE, 0)
    tri_check(capi.TERRAIN3D_PERT, 2)
    assert gpu.stitch(download=False) >= 0
    tri_check(capi.TERRAIN2D_PERT, 2)
    tri_check(capi.TORUS_Z, 0)

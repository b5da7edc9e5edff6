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


def test_cluster_smoothing_variant_is_bit_identical(gpu):
    """BMF_SMOOTH_CLUSTER=1: the two-CTA-cluster (DSMEM) form of the fused smoothing kernel gives the same bits"""
    import os
    from binarymeshfitting_b200 import Context, capi, world
    ps = world.grid_chunks(8, 16.0, origin=(-64.0, -48.0, -64.0))  # 512 chunks: enough for the fused path
    descs = capi.make_chunk_descs(ps, overlaps=0.045)
    gpu.set_sampler(capi.TERRAIN2D_PERT)
    gpu.submit(descs, 64, iters=3)
    gpu.wait()
    want = gpu.download(want=("pos", "inds"))
    os.environ["BMF_SMOOTH_CLUSTER"] = "1"
    try:
        ctx = Context(0)
    finally:
        del os.environ["BMF_SMOOTH_CLUSTER"]
    ctx.set_sampler(capi.TERRAIN2D_PERT)
    ctx.set_kernel_timing(True)
    ctx.submit(descs, 64, iters=3)
    ctx.wait()
    assert any(name == "k_smooth_chunks2" for name, _ in ctx.kernel_times())
    got = ctx.download(want=("pos", "inds"))
    ctx.close()
    assert len(want["pos"]) > 100000
    np.testing.assert_array_equal(got["inds"], want["inds"])
    np.testing.assert_array_equal(got["pos"].view(np.uint32), want["pos"].view(np.uint32))

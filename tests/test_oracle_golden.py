"""CPU: the plain-C oracle (oracle/bmf_oracle.c) against the committed golden vectors, which were produced by the
COMPILED REFERENCE (tests/golden/make_golden.py).  This is what pins the oracle when /root/reference is absent."""
import json
import os
import zlib

import numpy as np
import pytest

from oracle import oracle_binding as ob
from oracle.ref_binding import fnv1a64

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "golden.json")))
ARR = np.load(os.path.join(HERE, "golden", "golden_arrays.npz"))


def crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xFFFFFFFF


@pytest.mark.parametrize("row", GOLD["appendix_a"], ids=lambda r: "%s%d" % (r["fn"], r["dim"]))
def test_appendix_a_vectors(oracle, row):
    o = oracle.chunk(oracle.sampler(row["kind"]), (-128, -128, -128), 256.0, row["dim"])
    assert (o["n_cells"], o["n_verts"], o["n_inds"]) == (row["cells"], row["verts"], row["inds"])
    assert "%016x" % fnv1a64(o["bits"]) == row["bits_fnv"]
    assert "%016x" % fnv1a64(o["inds"]) == row["inds_fnv"]
    assert "%016x" % fnv1a64(o["pos"]) == row["pos_fnv"]
    assert crc(o["masks"]) == row["masks_crc"]
    assert crc(o["valence"]) == row["valence_crc"] and crc(o["boundary"]) == row["boundary_crc"]
    np.testing.assert_allclose(o["pos"][0], row["v0"], rtol=0, atol=0)
    np.testing.assert_allclose(o["pos"][-1], row["vlast"], rtol=0, atol=0)


@pytest.mark.parametrize("c", GOLD["chunks"], ids=lambda c: "k%d_d%d_it%d" % (c["kind"], c["dim"], c["iters"]))
def test_chunk_cases(oracle, c):
    o = oracle.chunk(oracle.sampler(c["kind"]), c["pos"], c["size"], c["dim"], c["overlap"], iters=c["iters"], process_boundary=c["pb"], smooth_normals=c["sn"])
    assert o["contains_mesh"] == c["contains_mesh"]
    assert crc(o["density"]) == c["density_crc"] and crc(o["bits"]) == c["bits_crc"]
    assert (o["n_cells"], o["n_verts"], o["n_inds"]) == (c["cells"], c["verts"], c["inds"])
    if c["contains_mesh"]:
        assert crc(o["masks"]) == c["masks_crc"] and crc(o["inds"]) == c["inds_crc"]
        assert crc(o["pos"]) == c["pos_crc"] and crc(o["color"]) == c["color_crc"]
        assert crc(o["valence"]) == c["valence_crc"] and crc(o["boundary"]) == c["boundary_crc"]


@pytest.mark.parametrize("kind", [ob.TERRAIN2D, ob.TERRAIN2D_PERT, ob.TERRAIN3D, ob.TERRAIN3D_PERT])
def test_noise_density_samples(oracle, kind):
    op, delta = oracle.geometry((-64, -64, -64), 128.0, 32, 0.045)
    d = oracle.sample_block(oracle.sampler(kind), op, delta, 32)
    np.testing.assert_array_equal(d[::1021].view(np.uint32), ARR["density_kind%d" % kind].view(np.uint32))


def test_qef_known_answers(oracle):
    for k in GOLD["qef"]:
        x, err = oracle.qef_solve(k["p"], k["n"])
        # the reference uses _mm_rsqrt_ps (12-bit approximation): agreement to ~1e-6 on well-conditioned systems
        np.testing.assert_allclose(x, k["x"], atol=2e-6)
        assert abs(err - k["err"]) <= 2e-6 * max(1.0, k["err"])


def test_qef_random_systems_vs_reference_outputs(oracle):
    """64 random 2..12-plane systems solved by the compiled reference (x86 rsqrt approximation).  Bar: 1e-4 of the
    unit cell for all but threshold-flip outliers, which are counted, not hidden."""
    p, n, cnt, out = ARR["qef_p"], ARR["qef_n"], ARR["qef_counts"], ARR["qef_out"]
    bad = 0
    for j in range(len(cnt)):
        x, _ = oracle.qef_solve(p[j, :cnt[j]], n[j, :cnt[j]])
        if np.abs(x - out[j]).max() > 1e-4:
            bad += 1
    assert bad <= 1, "%d of %d systems differ by more than 1e-4" % (bad, len(cnt))


def test_empty_and_degenerate_inputs(oracle):
    for fill in (1.0, -1.0):
        o = oracle.chunk(oracle.sampler(ob.HOST_DENSITY), (0, 0, 0), 1.0, 32, host_density=np.full(32 ** 3, fill, np.float32))
        assert not o["contains_mesh"] and o["n_verts"] == 0
    x, err = oracle.qef_solve(np.zeros((1, 3), np.float32), np.zeros((1, 3), np.float32))
    assert not x.any() and err == 0.0
    pos, col, nrm = oracle.smooth(np.zeros((0, 3)), np.zeros((0, 3)), None, np.zeros(0), np.zeros(0), np.zeros(0), iters=3)
    assert pos.shape == (0, 3)


def test_mc_table_properties():
    """packed triangle table: counts are multiples of 3 (<= 15) and every edge id is < 12."""
    import re
    txt = open(os.path.join(os.path.dirname(HERE), "oracle", "mc_tables_oracle.h")).read()
    allv = [int(v, 16) for v in re.findall(r"0x([0-9a-f]{16})ull", txt)]
    assert len(allv) == 512  # tri_pack, then patch_pack
    vals, patches = allv[:256], allv[256:]
    for m, v in enumerate(vals):
        n = v >> 60
        assert n % 3 == 0 and n <= 15
        assert all(((v >> (4 * i)) & 15) < 12 for i in range(n))
    assert vals[0] == 0 and vals[255] == 0
    # patch table (quad emission): disjoint patches that cover exactly the edges the triangles use
    for m, (v, pv) in enumerate(zip(vals, patches)):
        used = 0
        for i in range(v >> 60):
            used |= 1 << ((v >> (4 * i)) & 15)
        npatch, union = pv >> 60, 0
        assert npatch <= 4
        for k in range(npatch):
            pm = (pv >> (12 * k)) & 0xFFF
            assert pm and not (pm & union)
            union |= pm
        assert union == used and (pv >> (12 * npatch)) & ((1 << (60 - 12 * npatch)) - 1) == 0


# ---- round 2: gradient, collapse_bad_quads, ColorMapper (tests/golden/make_post_golden.py, outputs of the compiled reference)
POST = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "post_golden.npz"))


@pytest.mark.parametrize("kind", [0, 1, 2, 3])
def test_gradient_golden(oracle, kind):
    g = oracle.sampler_gradient(oracle.sampler(kind), POST["grad_points"])
    np.testing.assert_array_equal(g.view(np.uint32), POST["grad_%d" % kind].view(np.uint32))


def test_color_map_golden(oracle):
    np.testing.assert_array_equal(oracle.color_map(POST["color_points"]).view(np.uint32), POST["color_rgb"].view(np.uint32))


@pytest.mark.parametrize("name", ["sphere32", "torus64"])
def test_collapse_bad_quads_golden(oracle, name):
    o = oracle.collapse_bad_quads(POST["cq_%s_pos_in" % name], POST["cq_%s_quads_in" % name])
    np.testing.assert_array_equal(o["flushed"], POST["cq_%s_flushed" % name])
    np.testing.assert_array_equal(o["pos"].view(np.uint32), POST["cq_%s_pos_out" % name].view(np.uint32))
    np.testing.assert_array_equal(o["adj_next"], POST["cq_%s_adj_next" % name])
    assert o["bad_count"] == len(o["quads"]) - len(o["flushed"]) > 0

"""GPU: MeshProcessor<4>::collapse_bad_quads (MeshProcessor.cpp:308-396) and ColorMapper::generate_colors (ColorMapper.cpp:15-60) on the
device against the oracle twins (pinned to the compiled reference, tests/test_oracle_vs_ref.py) and against the committed golden
vectors the compiled reference produced (tests/golden/post_golden.npz)."""
import os
import struct
import subprocess
import zlib

import numpy as np
import pytest

from binarymeshfitting_b200 import capi
from oracle import oracle_binding as ob

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
POST = np.load(os.path.join(ROOT, "tests", "golden", "post_golden.npz"))


def crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xFFFFFFFF


def gpu_quads(gpu, kind, dim, density=None, **kw):
    gpu.set_sampler(kind, **kw)
    gpu.submit(capi.make_chunk_descs([[-128, -128, -128, 256.0]]), dim, quads=True, density=density)
    gpu.wait()
    out = gpu.download()
    return out["pos"], out["inds"].reshape(-1, 4)


@pytest.mark.parametrize("name", ["sphere32", "torus64"])
def test_collapse_bad_quads_golden(gpu, name):
    g = gpu.collapse_bad_quads(POST["cq_%s_pos_in" % name], POST["cq_%s_quads_in" % name])
    np.testing.assert_array_equal(g["flushed"], POST["cq_%s_flushed" % name])
    np.testing.assert_array_equal(g["pos"].view(np.uint32), POST["cq_%s_pos_out" % name].view(np.uint32))
    np.testing.assert_array_equal(g["adj_next"], POST["cq_%s_adj_next" % name])
    assert g["bad_count"] == int(g["destroyed"].sum()) == len(g["quads"]) - len(g["flushed"]) > 0


@pytest.mark.parametrize("case", ["sphere64", "cuboid64", "terrain3d_64", "random32", "random64", "csg128"])
def test_collapse_bad_quads_on_device_quad_meshes(gpu, oracle, case):
    """quad meshes straight from the device's DMC quad emitter (bmf_params.quads), collapsed on the device and by the oracle's serial loop"""
    if case.startswith("random"):
        dim = int(case[6:])
        rng = np.random.default_rng(dim)
        f = rng.standard_normal((dim, dim, dim)).astype(np.float32)
        for ax in range(3):
            f = (f + np.roll(f, 1, ax) + np.roll(f, -1, ax)).astype(np.float32)
        pos, quads = gpu_quads(gpu, ob.HOST_DENSITY, dim, density=f.reshape(1, -1))
    elif case == "csg128":
        pos, quads = gpu_quads(gpu, ob.CSG, 128, csg_op=ob.CSG_UNION, csg_kind_a=ob.SPHERE, csg_kind_b=ob.TORUS_Z, csg_world_size_a=256.0, csg_world_size_b=300.0,
                               csg_offset_a=(0.0, 0.0, 0.0), csg_offset_b=(20.0, -10.0, 5.0))
    else:
        kind = {"sphere64": ob.SPHERE, "cuboid64": ob.CUBOID, "terrain3d_64": ob.TERRAIN3D_PERT}[case]
        pos, quads = gpu_quads(gpu, kind, 64)
    assert len(quads) > 1000
    g, o = gpu.collapse_bad_quads(pos, quads), oracle.collapse_bad_quads(pos, quads)
    assert g["bad_count"] == o["bad_count"]
    np.testing.assert_array_equal(g["destroyed"], o["destroyed"])
    np.testing.assert_array_equal(g["quads"], o["quads"])
    np.testing.assert_array_equal(g["flushed"], o["flushed"])
    np.testing.assert_array_equal(g["adj_next"], o["adj_next"])
    np.testing.assert_array_equal(g["pos"].view(np.uint32), o["pos"].view(np.uint32))
    if case != "cuboid64":
        assert o["bad_count"] > 0
    # flush_to_tris of the survivors (MeshProcessor.cpp:73-91)
    t = gpu.quads_to_tris(g["flushed"])
    assert t.shape == (2 * len(g["flushed"]), 3) and np.array_equal(t[0], g["flushed"][0][[0, 1, 2]]) and np.array_equal(t[1], g["flushed"][0][[2, 3, 0]])


def test_collapse_bad_quads_edge_cases(gpu, oracle):
    # nothing to do / empty
    g = gpu.collapse_bad_quads(np.zeros((0, 3), np.float32), np.zeros((0, 4), np.uint32))
    assert g["bad_count"] == 0 and len(g["flushed"]) == 0
    # a single quad, a strip, and a cube (8 valence-3 vertices, every quad has next == 4: the `continue` branch)
    one = (np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0]], np.float32), np.array([[0, 1, 2, 3]], np.uint32))
    cube_p = np.array([[x, y, z] for x in (0, 1) for y in (0, 1) for z in (0, 1)], np.float32)
    cube_q = np.array([[0, 1, 3, 2], [4, 6, 7, 5], [0, 4, 5, 1], [2, 3, 7, 6], [0, 2, 6, 4], [1, 5, 7, 3]], np.uint32)
    for p, q in (one, (cube_p, cube_q)):
        g, o = gpu.collapse_bad_quads(p, q), oracle.collapse_bad_quads(p, q)
        assert g["bad_count"] == o["bad_count"] and np.array_equal(g["quads"], o["quads"]) and np.array_equal(g["pos"], o["pos"]) and np.array_equal(g["adj_next"], o["adj_next"])
    with pytest.raises(capi.BmfError, match="out of range"):
        gpu.collapse_bad_quads(one[0], np.array([[0, 1, 2, 9]], np.uint32))


def test_color_map(gpu, oracle):
    np.testing.assert_array_equal(gpu.color_map(POST["color_points"]).view(np.uint32), POST["color_rgb"].view(np.uint32))  # compiled reference's output
    rng = np.random.default_rng(4)
    pos = np.concatenate([(rng.random((50000, 3), dtype=np.float32) * 63).astype(np.float32), np.zeros((1, 3), np.float32), np.full((1, 3), -1e4, np.float32)])
    g, o = gpu.color_map(pos), oracle.color_map(pos)
    np.testing.assert_array_equal(g.view(np.uint32), o.view(np.uint32))
    assert g.min() >= 0.28 - 1e-6 and g.max() <= 1.0
    assert gpu.color_map(np.zeros((0, 3), np.float32)).shape == (0, 3)


def test_cpp_mirror_collapse_flush_and_colors(oracle, tmp_path):
    """the sequence the reference left commented out (DebugScene.cpp:253-262): MeshProcessor<4>::init, collapse_bad_quads, flush / flush_to_tris,
    then ColorMapper::generate_colors, through the C++ mirror"""
    pos, quads = POST["cq_torus64_pos_in"], POST["cq_torus64_quads_in"]
    f = tmp_path / "mesh.bin"
    f.write_bytes(struct.pack("<II", len(pos), len(quads)) + np.ascontiguousarray(pos, np.float32).tobytes() + np.ascontiguousarray(quads, np.uint32).tobytes())
    exe = os.path.join(ROOT, "binarymeshfitting_b200", "host", "host_test")
    o = oracle.collapse_bad_quads(pos, quads)
    col = oracle.color_map(o["pos"])
    for tris in (0, 1):
        r = subprocess.run([exe, "quadpost", str(f), str(tris)], capture_output=True, text=True, timeout=120)
        assert r.returncode == 0, r.stderr
        rec = dict(kv.split("=") for kv in r.stdout.split()[1:])
        fl = o["flushed"]
        want = fl if not tris else np.stack([fl[:, [0, 1, 2]], fl[:, [2, 3, 0]]], axis=1).reshape(-1, 3)
        assert int(rec["bad"]) == o["bad_count"] and int(rec["verts"]) == len(pos) and int(rec["inds"]) == want.size
        assert int(rec["inds_crc"]) == crc(want.astype(np.uint32)) and int(rec["pos_crc"]) == crc(o["pos"])
        assert int(rec["color_crc"]) == crc(col) and int(rec["adj_next_crc"]) == crc(o["adj_next"])

"""GPU: the C++ mirror of the reference classes (binarymeshfitting_b200/host/bmf_host.hpp) driven by a C++ program
written like the reference's own callers (DMCChunk staged calls, MeshProcessor<3> sequence of ChunkGenerator.cpp:112-123,
ChunkGenerator::process_queue over a leaf list), checked against the oracle."""
import json
import os
import re
import subprocess
import zlib

import numpy as np
import pytest

from binarymeshfitting_b200 import world as W
from oracle import oracle_binding as ob

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "binarymeshfitting_b200", "host", "host_test")
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json")))
ARR = np.load(os.path.join(ROOT, "tests", "golden", "golden_arrays.npz"))


def crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xFFFFFFFF


def run(*args):
    if not os.path.exists(EXE):
        subprocess.run(["bash", os.path.join(ROOT, "binarymeshfitting_b200", "host", "build_host_test.sh")], check=True)
    r = subprocess.run([EXE] + [str(a) for a in args], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    out = {}
    for line in r.stdout.splitlines():
        parts = line.split()
        tag = parts[0] if "=" not in parts[0] else "misc"
        out.setdefault(tag, {}).update({k: v for k, v in (p.split("=") for p in parts if "=" in p)})
    return out


def check(rec, o, processed=False):
    assert int(rec["contains_mesh"]) == int(o["contains_mesh"])
    assert (int(rec["verts"]), int(rec["inds"])) == (o["n_verts"], o["n_inds"])
    assert int(rec["bits_crc"]) == crc(o["bits"])
    if o["n_verts"]:
        assert int(rec["inds_crc"]) == crc(o["inds"]) and int(rec["pos_crc"]) == crc(o["pos"])
        assert int(rec["boundary_crc"]) == crc(o["boundary"]) and int(rec["valence_crc"]) == crc(o["valence"])
    if not processed:
        assert int(rec["cells"]) == o["n_cells"]


@pytest.mark.parametrize("kind,dim,overlap,iters", [(ob.SPHERE, 64, 0.0, 0), (ob.TORUS_Z, 64, 0.045, 2), (ob.TERRAIN2D_PERT, 32, 0.045, 2), (ob.TERRAIN3D_PERT, 32, 0.055, 4)])
def test_dmcchunk_staged_calls_and_meshprocessor(oracle, kind, dim, overlap, iters):
    out = run("chunk", kind, dim, overlap, iters)
    pos, size = (-128, -128, -128), 256.0
    o0 = oracle.chunk(oracle.sampler(kind), pos, size, dim, np.float32(overlap))
    check(out["extract"], o0)
    assert int(out["extract"]["density_crc"]) == crc(o0["density"])
    assert np.float32(out["extract"]["scale"]) == np.float32(o0["scale"])
    assert int(out["misc"]["valence_sum_before_polygonize"]) == 0  # label_edges publishes vertices, polygonize the valences
    if iters:
        o1 = oracle.chunk(oracle.sampler(kind), pos, size, dim, np.float32(overlap), iters=iters)
        check(out["processed"], o1, processed=True)


@pytest.mark.parametrize("mode", [0, 1])
def test_dmcchunk_interleaved_and_threaded_chunks_keep_their_own_results(oracle, mode):
    """ChunkGenerator.cpp:93-108 runs the three DMCChunk stages per chunk inside an OpenMP loop: two chunks interleaved on one
    thread (grid A, grid B, edges A, ...) or driven from two threads must each publish their OWN vertices and indices."""
    out = run("interleave", 64, mode)
    for tag, kind in (("A", ob.SPHERE), ("B", ob.TORUS_Z)):
        check(out[tag], oracle.chunk(oracle.sampler(kind), (-128, -128, -128), 256.0, 64, np.float32(0.0)))


def test_arbitrary_host_callback_sampler(oracle, tmp_path):
    f = tmp_path / "density.bin"
    out = run("hostfn", 32, f)
    dens = np.fromfile(f, np.float32)
    o = oracle.chunk(oracle.sampler(ob.HOST_DENSITY), (-64, -64, -64), 128.0, 32, 0.0, host_density=dens)
    assert o["n_verts"] > 100
    check(out["hostfn"], o)


@pytest.mark.parametrize("w", [g for g in GOLD["worlds"] if g["dim"] == 32], ids=lambda w: w["key"])
def test_chunkgenerator_process_queue(tmp_path, w):
    props = W.WorldProperties(max_level=w["max_level"], chunk_resolution=w["dim"], process_iters=w["iters"])
    ps, lv, mc = W.split_leaves(props, 256, tuple(w["focus"]))
    f = tmp_path / "leaves.txt"
    with open(f, "w") as fh:
        for p, l, m in zip(ps, lv, mc):
            fh.write("%r %r %r %r %d %d\n" % (float(p[0]), float(p[1]), float(p[2]), float(p[3]), int(l), int(m)))
    out = run("world", w["kind"], w["dim"], w["max_level"], w["iters"], f)["world"]
    assert (int(out["chunks"]), int(out["with_mesh"]), int(out["verts"]), int(out["inds"])) == (w["leaves"], w["chunks_with_mesh"], w["verts"], w["inds"])
    assert int(out["inds_crc"]) == w["inds_crc"]  # golden: the compiled reference's ChunkGenerator::process_queue
    assert int(out["needs_upload"]) == w["chunks_with_mesh"]


@pytest.mark.parametrize("devices", [None, "0,0", "0,0,0"])
@pytest.mark.parametrize("w", [g for g in GOLD["worlds"] if g["dim"] == 32 and g["focus"] == [0, 0, 0]], ids=lambda w: w["key"])
def test_cpp_split_leaves_and_multi_context_partition(w, devices):
    """all in C++: WorldOctree::init -> split_leaves -> process_queue, optionally dealt to several contexts (the
    multi-GPU partition; here several contexts on GPU 0) -- same leaves and the same per-chunk meshes as the reference"""
    args = ["lod", w["kind"], w["dim"], w["max_level"], w["iters"]] + ([devices] if devices else [])
    out = run(*args)["lod"]
    assert (int(out["chunks"]), int(out["with_mesh"]), int(out["verts"]), int(out["inds"])) == (w["leaves"], w["chunks_with_mesh"], w["verts"], w["inds"])
    assert int(out["inds_crc"]) == w["inds_crc"]
    leaves = ARR[w["key"] + "_leaves"][:, :4].astype(np.float32)
    assert int(out["leaves_crc"]) == crc(leaves)


def test_worldstitcher_mirror_closes_a_lod_world(oracle):
    """C++: properties.enable_stitching -> process_queue samples at voxel-node centres, ChunkGenerator::stitcher
    (WorldStitcher mirror) emits the seam soup; same triangles as the oracle's seam pass over the same leaves"""
    focus, dim, max_level = (90.0, 20.0, -30.0), 32, 3
    out = run("stitch", ob.TORUS_Z, dim, max_level, *focus)["stitch"]
    props = W.WorldProperties(max_level=max_level, chunk_resolution=dim)
    ps, lv, mc = W.split_leaves(props, 256, focus)
    ov = np.float32(-0.5) / np.float32(dim)
    s = oracle.sampler(ob.TORUS_Z)
    chunks = [oracle.chunk(s, p[:3], p[3], dim, ov) for p in ps]
    seam = oracle.seam(chunks, ps, dim, ov)
    assert int(out["chunks"]) == len(ps) and len(seam) > 0
    assert (int(out["verts"]), int(out["inds"])) == (sum(c["n_verts"] for c in chunks), sum(c["n_inds"] for c in chunks))
    assert int(out["seam_verts"]) == 3 * len(seam) and int(out["seam_crc"]) == crc(seam)
    assert abs(float(out["color_g"]) - 1.0) < 1e-6
    tri_sum = sum(zlib.crc32(t.tobytes()) & 0xFFFFFFFF for t in np.ascontiguousarray(seam, np.float32))
    assert int(out["tri_sum"]) == tri_sum
    # the multi-GPU scheme (here: several contexts on GPU 0): per-device seams + one cross-device pass over the border chunks
    for devices in ("0,0", "0,0,0"):
        multi = run("stitch", ob.TORUS_Z, dim, max_level, *focus, devices)["stitch"]
        assert int(multi["seam_verts"]) == 3 * len(seam) and int(multi["tri_sum"]) == tri_sum


def test_worldwatcher_mirror_follows_a_moving_focus(oracle):
    """C++ WorldWatcher mirror (synchronous tick) against the same policy in Python (world.LodWatcher): same final leaf
    list in the same link order, and the meshes of exactly those leaves (generated incrementally) match the oracle"""
    target, dim, max_level, steps = (150.0, 40.0, -60.0), 32, 5, 12
    out = run("fly", ob.SPHERE, dim, max_level, *target, steps)["fly"]
    props = W.WorldProperties(max_level=max_level, chunk_resolution=dim)
    w = W.LodWatcher(props, 256, (0.0, 0.0, 0.0))
    generated, ticks = 0, 0
    for k in range(1, steps + 65):
        t = np.float32(1.0) if k >= steps else np.float32(k) / np.float32(steps)
        focus = tuple(np.float32(c) * t for c in target)
        gen = w.tick(focus)
        generated += len(gen)
        ticks += 1
        if k >= steps and not gen:
            break
    ps, lv, mc = w.leaves()
    assert (int(out["leaves"]), int(out["generated"]), int(out["ticks"])) == (len(ps), generated, ticks)
    assert int(out["codes_crc"]) == crc(mc.astype("<u8"))
    total, counts = oracle.batch(oracle.sampler(ob.SPHERE), ps, dim, overlaps=[W.chunk_overlap(props, int(l)) for l in lv])
    assert (int(out["verts"]), int(out["inds"])) == (int(counts[:, 0].sum()), int(counts[:, 1].sum()))


WATCHER_GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "watcher_golden.json")))


@pytest.mark.parametrize("g", WATCHER_GOLD[:3], ids=lambda g: g["name"])
def test_worldwatcher_mirror_matches_the_reference_watcher(tmp_path, g):
    """C++ WorldWatcher mirror started from the root, one tick per focus point: the batches and the final renderables list of
    the compiled reference's watcher (golden vectors, tests/golden/make_watcher_golden.py)"""
    f = tmp_path / "path.txt"
    f.write_text("".join("%r %r %r\n" % tuple(p) for p in g["path"]))
    out = run("watch", ob.SPHERE, 32, g["max_level"], f)["watch"]
    assert [int(v) for v in out["gens"].strip(",").split(",")] == g["generated_per_tick"]
    assert int(out["leaves"]) == g["renderables"] and int(out["codes_crc"]) == g["codes_crc"]

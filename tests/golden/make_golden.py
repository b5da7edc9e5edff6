#!/usr/bin/env python
"""Generate tests/golden/golden.json + golden_arrays.npz from the COMPILED REFERENCE (oracle/_ref/libbmf_ref.so,
the unmodified translation units of /root/reference).  Run in the authoring container:

    make -C oracle ref && python tests/golden/make_golden.py

The reference ships no tests or fixtures of its own; these vectors are its outputs on
seeded/analytic inputs.  Hashes: FNV-1a-64 over raw little-endian bytes (SURVEY Appendix A convention) for
small arrays, zlib.crc32 for large ones.  Noise cases go through the restated FastNoiseSIMD
(oracle/fastnoise_ref.h) -- the reference's own NoiseSampler.cpp call sites on top of an UNPINNED noise.
"""
import json
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_binding as rb  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xFFFFFFFF


def main():
    R = rb.RefLib()
    gold = {"appendix_a": [], "chunks": [], "qef": [], "worlds": [], "smooth": []}
    arrays = {}

    # SURVEY Appendix A (single chunk, overlap 0, no processing)
    names = {"sphere": rb.SPHERE, "torus_z": rb.TORUS_Z, "cuboid": rb.CUBOID, "plane_y": rb.PLANE_Y}
    for name, dim in [("sphere", 32), ("sphere", 64), ("torus_z", 64), ("cuboid", 64), ("plane_y", 64), ("sphere", 128), ("torus_z", 128), ("cuboid", 128)]:
        o = R.chunk(names[name], (-128, -128, -128), 256.0, dim)
        gold["appendix_a"].append({
            "fn": name, "kind": names[name], "dim": dim, "cells": o["n_cells"], "verts": o["n_verts"], "inds": o["n_inds"],
            "bits_fnv": "%016x" % rb.fnv1a64(o["bits"]), "inds_fnv": "%016x" % rb.fnv1a64(o["inds"]),
            "pos_fnv": "%016x" % rb.fnv1a64(np.ascontiguousarray(o["verts"]["p"])),
            "masks_crc": crc(o["masks"]), "valence_crc": crc(o["verts"]["init_valence"]), "boundary_crc": crc(o["verts"]["boundary"]),
            "v0": [float(x) for x in o["verts"]["p"][0]], "vlast": [float(x) for x in o["verts"]["p"][-1]]})

    # assorted chunks: overlap, offsets, noise terrains, smoothing
    cases = [
        (rb.TORUS_Z, (-128, -128, -128), 256.0, 64, 0.045, 2, False, False),
        (rb.SPHERE, (-100.0, -90.0, -80.0), 200.0, 64, 0.02, 0, False, False),
        (rb.SPHERE, (-128, -128, -128), 256.0, 64, 0.06, 5, True, False),
        (rb.CUBOID, (-128, -128, -128), 256.0, 32, 0.055, 4, False, True),
        (rb.TERRAIN2D, (-64, -64, -64), 128.0, 64, 0.045, 2, False, False),
        (rb.TERRAIN2D_PERT, (-64, -64, -64), 128.0, 64, 0.045, 2, False, False),
        (rb.TERRAIN2D_PERT, (-16, 0, -16), 16.0, 64, 0.045, 0, False, False),
        (rb.TERRAIN3D, (-64, -64, -64), 128.0, 32, 0.045, 2, False, False),
        (rb.TERRAIN3D_PERT, (-64, -64, -64), 128.0, 32, 0.045, 2, False, False),
        (rb.TERRAIN3D_PERT, (0, -32, 0), 32.0, 64, 0.035, 0, False, False),
    ]
    for kind, pos, size, dim, ov, iters, pb, sn in cases:
        o = R.chunk(kind, pos, size, dim, overlap=ov, iters=iters, process_boundary=pb, smooth_normals=sn)
        e = {"kind": kind, "pos": list(map(float, pos)), "size": size, "dim": dim, "overlap": ov, "iters": iters, "pb": pb, "sn": sn,
             "contains_mesh": o["contains_mesh"], "cells": o["n_cells"], "verts": o["n_verts"], "inds": o["n_inds"],
             "density_crc": crc(o["density"]), "bits_crc": crc(o["bits"])}
        if o["contains_mesh"]:
            e.update(masks_crc=crc(o["masks"]), inds_crc=crc(o["inds"]), pos_crc=crc(np.ascontiguousarray(o["verts"]["p"])),
                     color_crc=crc(np.ascontiguousarray(o["verts"]["color"])), valence_crc=crc(o["verts"]["init_valence"]),
                     boundary_crc=crc(o["verts"]["boundary"]))
        gold["chunks"].append(e)

    # a few raw density samples of each noise terrain (the restated FastNoiseSIMD behind the reference's call sites)
    for kind in (rb.TERRAIN2D, rb.TERRAIN2D_PERT, rb.TERRAIN3D, rb.TERRAIN3D_PERT):
        o = R.chunk(kind, (-64, -64, -64), 128.0, 32, overlap=0.045, want=("density",))
        arrays["density_kind%d" % kind] = o["density"][:: 1021].copy()

    # QEF known answers (qef_solve_from_points_3d on this CPU, incl. its _mm_rsqrt_ps)
    rng = np.random.default_rng(11)
    qp = rng.random((64, 12, 3), dtype=np.float32)
    qn = rng.normal(size=(64, 12, 3)).astype(np.float32)
    qn /= np.linalg.norm(qn, axis=2, keepdims=True)
    qc = rng.integers(2, 13, 64).astype(np.int32)
    qo = np.zeros((64, 3), np.float32)
    qe = np.zeros(64, np.float32)
    for j in range(64):
        qo[j], qe[j] = R.qef_solve(qp[j, :qc[j]], qn[j, :qc[j]])
    arrays.update(qef_p=qp, qef_n=qn, qef_counts=qc, qef_out=qo, qef_err=qe)
    gold["qef"] = [
        {"p": [[1, 0, 0], [0, 2, 0], [0, 0, 3], [1, 5, 7]], "n": [[1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 0, 0]], "x": [1.0, 2.0, 3.0], "err": 10.3125},
        {"p": [[.5, .25, 0], [0, .5, .75], [.3, 0, .6]], "n": [[.6, .8, 0], [0, .6, .8], [.8, 0, .6]], "x": [0.181318581, 0.48901099, 0.758241773], "err": 1.20324755},
    ]

    # LOD worlds: leaf lists + whole-world mesh totals (WorldOctree::split_leaves -> ChunkGenerator::process_queue)
    for kind, dim, ml, iters, focus in [(rb.SPHERE, 32, 5, 0, (0, 0, 0)), (rb.SPHERE, 64, 5, 2, (0, 0, 0)), (rb.TERRAIN2D_PERT, 32, 5, 2, (0, 0, 0)),
                                        (rb.TERRAIN2D_PERT, 32, 6, 0, (37.5, -80.25, 100.0))]:
        w = R.world(kind, dim, max_level=ml, iters=iters, focus=focus)
        n = w.split_leaves()
        ps, lv, mc = w.leaves()
        w.process(8)
        nm, nv, ni = w.totals()
        h = 0
        per = np.zeros((n, 2), np.int32)
        for i in range(n):
            c = w.chunk(i)
            per[i] = (len(c["verts"]), len(c["inds"]))
            h = zlib.crc32(c["inds"].tobytes(), h)
        key = "world_%d_%d_%d_%d" % (kind, dim, ml, iters)
        arrays[key + "_leaves"] = np.concatenate([ps, lv[:, None].astype(np.float32)], axis=1)
        arrays[key + "_morton"] = mc
        arrays[key + "_counts"] = per
        gold["worlds"].append({"key": key, "kind": kind, "dim": dim, "max_level": ml, "iters": iters, "focus": list(focus), "leaves": n,
                               "chunks_with_mesh": nm, "verts": nv, "inds": ni, "inds_crc": h & 0xFFFFFFFF})

    # the benchmark workload (configs[2]: 16^3 grid of 64^3 chunks, terrain2d_pert, 2 iterations): per-chunk counts + checksums
    sys.path.insert(0, ROOT)
    from binarymeshfitting_b200 import world as W
    ps = W.grid_chunks(16, 16.0)
    w = R.world(rb.TERRAIN2D_PERT, 64, max_level=99, iters=2)
    w.add_chunks(ps, np.zeros(len(ps), np.int32))
    w.process(8)
    per = np.zeros((len(ps), 2), np.int32)
    hi = hp = 0
    for i in range(len(ps)):
        c = w.chunk(i)
        per[i] = (len(c["verts"]), len(c["inds"]))
        hi = zlib.crc32(c["inds"].tobytes(), hi)
        hp = zlib.crc32(np.ascontiguousarray(c["verts"]["p"]).tobytes(), hp)
    arrays["bench_counts"] = per
    gold["bench_workload"] = {"chunks": len(ps), "dim": 64, "sampler": "terrain2d_pert", "iters": 2, "overlap": 0.045, "verts": int(per[:, 0].sum()),
                              "inds": int(per[:, 1].sum()), "inds_crc": hi & 0xFFFFFFFF, "pos_crc": hp & 0xFFFFFFFF}

    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(gold, f, indent=1)
    np.savez_compressed(os.path.join(HERE, "golden_arrays.npz"), **arrays)
    print("wrote golden.json (%d appendix rows, %d chunk cases, %d worlds) and golden_arrays.npz" % (len(gold["appendix_a"]), len(gold["chunks"]), len(gold["worlds"])))


if __name__ == "__main__":
    main()

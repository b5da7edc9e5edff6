#!/usr/bin/env python
"""Generate tests/golden/post_golden.npz from the COMPILED REFERENCE (oracle/_ref/libbmf_ref.so): known answers for the three pieces
around the extraction path that round 2 added -- Sampler::gradient (ImplicitSampler.hpp:38-49), MeshProcessor<4>::collapse_bad_quads
(MeshProcessor.cpp:308-396) and ColorMapper::generate_colors (ColorMapper.cpp:15-60).  Run in the authoring container:

    make -C oracle ref && python tests/golden/make_post_golden.py

The quad meshes fed to collapse_bad_quads come from the oracle's dual-marching-cubes emitter (the reference has no quad producer);
they are stored, so the fixture stands on its own.  ColorMapper goes through the restated FastNoiseSIMD (UNPINNED noise)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle_binding as ob  # noqa: E402
from oracle import ref_binding as rb  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    R, O = rb.RefLib(), ob.Oracle()
    out = {}
    rng = np.random.default_rng(2024)
    pts = (rng.random((256, 3), dtype=np.float32) * 300 - 150).astype(np.float32)
    out["grad_points"] = pts
    for kind in (rb.SPHERE, rb.TORUS_Z, rb.CUBOID, rb.PLANE_Y):
        out["grad_%d" % kind] = np.stack([R.implicit_gradient(kind, p) for p in pts])
    cpos = (rng.random((2048, 3), dtype=np.float32) * 512 - 256).astype(np.float32)
    v = np.zeros(len(cpos), rb.DUALVERTEX_DTYPE)
    v["p"] = cpos
    out["color_points"], out["color_rgb"] = cpos, R.color_map(v)["color"]
    for name, kind, dim in (("sphere32", ob.SPHERE, 32), ("torus64", ob.TORUS_Z, 64)):
        s = O.sampler(kind)
        op, delta = O.geometry((-128.0, -128.0, -128.0), 256.0, dim, 0.0)
        dens = O.sample_block(s, op, delta, dim)
        bits, _ = O.label_grid(dens, dim)
        q = O.quads(dens, bits, dim)
        vv = np.zeros(q["n_verts"], rb.DUALVERTEX_DTYPE)
        vv["p"], vv["color"], vv["boundary"], vv["init_valence"] = q["pos"], 1.0, q["boundary"], q["valence"]
        rv, rq = R.collapse_bad_quads(vv, q["inds"])
        out["cq_%s_pos_in" % name], out["cq_%s_quads_in" % name] = q["pos"], q["inds"].reshape(-1, 4)
        out["cq_%s_pos_out" % name], out["cq_%s_adj_next" % name], out["cq_%s_flushed" % name] = rv["p"], rv["adj_next"], rq
    np.savez_compressed(os.path.join(HERE, "post_golden.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()

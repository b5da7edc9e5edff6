"""Golden vectors for the incremental LOD policy, generated FROM THE COMPILED REFERENCE (oracle/_ref/libbmf_ref.so):
WorldWatcher::check_leaves -> process_batch -> ChunkGenerator::process_queue -> post_process_batch, driven headlessly
from the root by oracle/ref_export.cpp::ref_watcher_run.  Run in the authoring container (needs /root/reference):

    python tests/golden/make_watcher_golden.py
"""
import json
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle_binding as ob, ref_binding as rb  # noqa: E402


def paths():
    p1 = [(0.0, 0.0, 0.0)] * 8
    p2 = [(0.0, 0.0, 0.0)] * 6 + [(12.0 * k, 3.0 * k, -5.0 * k) for k in range(1, 25)] + [(288.0, 72.0, -120.0)] * 6
    p3 = [(0.0, 0.0, 0.0)] * 8 + [(-9.0 * k, float(np.float32(20.0 * np.sin(k / 3.0))), 7.0 * k) for k in range(1, 40)]
    p4 = [(200.0, -100.0, 50.0)] * 10 + [(200.0 - 20.0 * k, -100.0 + 10.0 * k, 50.0) for k in range(1, 21)] + [(-200.0, 100.0, 50.0)] * 10
    return [("static_l5", 5, p1), ("fly_l5", 5, p2), ("weave_l6", 6, p3), ("cross_l4", 4, p4)]


def main():
    R = rb.RefLib()
    out = []
    for name, max_level, path in paths():
        w = R.world(ob.SPHERE, 32, max_level=max_level)
        codes, gen = w.watcher_run(np.array(path, np.float32))
        out.append({"name": name, "max_level": max_level, "path": [[float(np.float32(c)) for c in p] for p in path], "generated_per_tick": gen.tolist(),
                    "renderables": int(len(codes)), "codes_crc": zlib.crc32(codes.astype("<u8").tobytes()) & 0xFFFFFFFF})
        print(name, len(codes), gen.tolist())
    with open(os.path.join(HERE, "watcher_golden.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()

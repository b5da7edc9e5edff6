"""per-kernel times of the seam pass on the benchmark grid and on the LOD world"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from binarymeshfitting_b200 import Context, capi, world

ctx = Context(0)
def probe(name, ps, lv, kind, dim, iters):
    ctx.set_sampler(kind)
    d = capi.make_chunk_descs(ps, overlaps=ctx.seam_overlap(dim), levels=lv)
    for r in range(3):
        ctx.set_kernel_timing(r == 2)
        ctx.submit(d, dim, iters=iters)
        ctx.wait()
        nt = ctx.stitch(download=False)
    kt = [(n, round(ms, 4)) for n, ms in ctx.kernel_times() if "seam" in n]
    ctx.set_kernel_timing(False)
    print(name, "tris", nt, ctx.seam_ms(), kt)

ps = world.grid_chunks(16, 16.0)
probe("grid4096 2d", ps, 0, capi.TERRAIN2D_PERT, 64, 2)
props = world.WorldProperties(max_level=5, chunk_resolution=64, process_iters=2)
lps, lv, mc = world.split_leaves(props)
probe("lod232 2d", lps, lv, capi.TERRAIN2D_PERT, 64, 2)
probe("lod232 3d", lps, lv, capi.TERRAIN3D_PERT, 64, 2)
probe("lod232 sphere", lps, lv, capi.SPHERE, 64, 0)

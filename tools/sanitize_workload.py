import sys; sys.path.insert(0, '.')
import numpy as np
from binarymeshfitting_b200 import capi, world, Context
ctx = Context(0)
# small batches through every sampler / option combination
ps = world.grid_chunks(3, 40.0, origin=(-60, -60, -60))
for kind in (capi.SPHERE, capi.CSG, capi.TERRAIN2D, capi.TERRAIN2D_PERT, capi.TERRAIN3D, capi.TERRAIN3D_PERT):
    ctx.set_sampler(kind)
    for dim in (32, 64):
        for iters, sn, qef, kd, km in ((0, False, False, False, False), (2, False, False, True, True), (4, True, True, False, False)):
            ctx.submit(capi.make_chunk_descs(ps, overlaps=0.045), dim, iters=iters, smooth_normals=sn, qef=qef, keep_density=kd, keep_masks=km)
            ctx.wait(); out = ctx.download(); print(kind, dim, iters, ctx.totals())
ctx.set_sampler(capi.TORUS_Z)
ctx.submit(capi.make_chunk_descs([[-128, -128, -128, 256.0]]), 128, iters=2); ctx.wait(); print(ctx.totals())
ctx.submit(capi.make_chunk_descs([[-128, -128, -128, 256.0]]), 256, iters=1); ctx.wait(); print(ctx.totals())
rng = np.random.default_rng(0)
f = rng.normal(size=2 * 32**3).astype(np.float32)
ctx.set_sampler(capi.HOST_DENSITY)
ctx.submit(capi.make_chunk_descs([[0,0,0,1.0],[1,0,0,1.0]]), 32, iters=2, density=f); ctx.wait(); print(ctx.totals())
p, c, n = ctx.mesh_process(rng.random((100,3),dtype=np.float32), np.ones((100,3),np.float32), np.zeros((100,3),np.float32), np.zeros(100,np.uint8), rng.integers(0,100,300).astype(np.uint32), 3, 3, True, True)
P = rng.random((50,12,3),dtype=np.float32); N = rng.normal(size=(50,12,3)).astype(np.float32)
print(ctx.qef_solve(P, N, rng.integers(2,13,50).astype(np.int32))[0][:2])
# seam pass over a small LOD world (2-D terrain with uniform chunks, 3-D noise, implicit) incl. the group filters
props = world.WorldProperties(max_level=3, chunk_resolution=32)
lps, lv, mc = world.split_leaves(props, 256, (40.0, -10.0, 25.0))
grp = (np.arange(len(lps)) % 2).astype(np.int32)
for kind in (capi.TERRAIN2D_PERT, capi.SPHERE):
    ctx.set_sampler(kind)
    ctx.submit(capi.make_chunk_descs(lps, overlaps=ctx.seam_overlap(32), levels=lv), 32, iters=2)
    print("seam", kind, len(ctx.stitch()), len(ctx.stitch(group=grp, cross_group_only=True)))
ctx.set_sampler(capi.TERRAIN3D_PERT)
ctx.submit(capi.make_chunk_descs(world.grid_chunks(2, 16.0, origin=(0, -16, 0)), overlaps=ctx.seam_overlap(64)), 64, iters=2)
print("seam 3d", len(ctx.stitch()))
print("sanitize workload done")

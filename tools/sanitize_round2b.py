"""compute-sanitizer workload for the kernels changed at the end of round 2: sign words made inside k_chunk_count (GEN), the warp-shared perturb corners
(shared memory + __syncwarp in the sampling kernels), k_smooth_chunks at 40 registers / two in flight, the capacity verdict inside k_scan_chunks and the
merged descriptor upload.  Run as:  compute-sanitizer --tool memcheck|racecheck python tools/sanitize_round2b.py"""
import sys; sys.path.insert(0, '.')
import numpy as np
from binarymeshfitting_b200 import capi, world, Context
ctx = Context(0)
ctx.set_batches_in_flight(4)  # per-chunk kernels at any batch size
ps = world.grid_chunks(3, 40.0, origin=(-60, -60, -60))
for kind in (capi.TERRAIN2D, capi.TERRAIN2D_PERT):
    ctx.set_sampler(kind)
    for dim, km in ((32, True), (64, False)):
        ctx.submit(capi.make_chunk_descs(ps, overlaps=0.045), dim, iters=2, keep_masks=km)
        ctx.wait(); ctx.download(); print(kind, dim, ctx.totals())
ctx.set_sampler(capi.TERRAIN3D_PERT)
ctx.submit(capi.make_chunk_descs(world.grid_chunks(2, 16.0, origin=(0, -16, 0)), overlaps=0.045), 32, iters=2)
ctx.wait(); print("3d", ctx.totals())
# enough chunks for k_smooth_chunks (>= 2 x SMs): 7^3 chunks of 32^3 straddling the surface
ctx.set_sampler(capi.TERRAIN2D_PERT)
big = capi.make_chunk_descs(world.grid_chunks(7, 8.0, origin=(-28.0, -36.0, -28.0)), overlaps=0.045)
for rep in range(2):
    ctx.submit(big, 32, iters=2)
    ctx.wait()
print("big", ctx.totals(), int((ctx.chunk_infos()["n_verts"] > 0).sum()), "chunks with mesh")
out = ctx.download(want=("pos", "inds"))
ctx.set_batches_in_flight(1)
ctx.submit(big, 32, iters=2); ctx.wait()
ref = ctx.download(want=("pos", "inds"))
assert np.array_equal(out["inds"], ref["inds"]) and np.array_equal(out["pos"].view(np.uint32), ref["pos"].view(np.uint32))
print("colors", ctx.color_map(np.random.default_rng(0).random((64, 3), dtype=np.float32) * 50)[:1])
print("round 2b sanitize workload done")

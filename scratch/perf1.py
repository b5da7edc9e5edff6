import sys, time; sys.path.insert(0, '.')
import numpy as np
from binarymeshfitting_b200 import capi, world, Context
ctx = Context(0)
def run(name, kind, descs, dim, iters, reps=4, **kw):
    ctx.set_sampler(kind)
    for r in range(reps):
        t = time.perf_counter()
        ctx.submit(descs, dim, iters=iters, **kw); ctx.wait()
        dt = time.perf_counter() - t
    st = ctx.stage_ms(); tot = ctx.totals()
    nvox = len(descs) * dim**3
    print("%-34s n=%5d dim=%3d it=%d wall %.3f ms  %.1f Gvox/s | " % (name, len(descs), dim, iters, dt*1e3, nvox/dt/1e9) + " ".join("%s %.3f" % (k, v) for k, v in st.items()) + " | C,V,I=%s" % (tot,), flush=True)
ps = world.grid_chunks(16, 16.0)
d = capi.make_chunk_descs(ps, overlaps=0.045)
for it in (0, 2):
    run("grid4096 terrain2d_pert", capi.TERRAIN2D_PERT, d, 64, it)
for it in (0, 2):
    run("grid4096 terrain3d_pert", capi.TERRAIN3D_PERT, d, 64, it)
run("grid4096 sphere", capi.SPHERE, d, 64, 0)
run("grid4096 sphere keepdens", capi.SPHERE, d, 64, 0, keep_density=True)
props = world.WorldProperties(max_level=5, chunk_resolution=64, process_iters=2)
ps, lv, mc = world.split_leaves(props)
dl = world.make_descs(props, ps, lv, mc)
run("LOD232 terrain2d_pert", capi.TERRAIN2D_PERT, dl, 64, 2, reps=6)
run("LOD232 terrain3d_pert", capi.TERRAIN3D_PERT, dl, 64, 2, reps=6)
run("LOD232 sphere", capi.SPHERE, dl, 64, 2, reps=6)
one = capi.make_chunk_descs([[-64,-64,-64,128.0]], overlaps=0.045)
run("single128 terrain3d_pert", capi.TERRAIN3D_PERT, one, 128, 2, reps=6)
run("single128 terrain2d_pert", capi.TERRAIN2D_PERT, one, 128, 2, reps=6)
# host density path (K2 pack) with device-resident density
import torch
dens = torch.randn(1024 * 64**3, device='cuda')
dd = capi.make_chunk_descs(world.grid_chunks(16,16.0)[:1024])
ctx.set_sampler(capi.HOST_DENSITY)
for r in range(3):
    t=time.perf_counter(); ctx.submit(dd, 64, density_device_ptr=dens.data_ptr()); ctx.wait(); dt=time.perf_counter()-t
print("hostdens(random) 1024x64^3", dt*1e3, ctx.stage_ms(), ctx.totals())

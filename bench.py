#!/usr/bin/env python
"""bench.py -- headline benchmark of the chunk-extraction hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--sampler NAME] [--iters I]

Workload (config.workload): BASELINE.json configs[2] -- a 16x16x16 grid of 4096 chunks of 64^3 voxels
(1.07 Gvoxel) of noise terrain, full pipeline sample -> sign bits -> cell masks -> vertex/index
emission -> 2 MeshProcessor<3> smoothing iterations.  One "step" = one ChunkGenerator::process_queue
of that batch.  N > 1: one process per GPU (torchrun), every rank meshes its own 4096-chunk region of
the same world (weak scaling, no data-path collective); value = all ranks' voxels / max-over-ranks time.

Prints ONE JSON line (rank 0).  `value` is device-timed with the inputs (chunk descriptors -> geometry)
resident; `e2e` goes through the C ABI with host descriptors in and the renderer-facing SoA meshes
(positions, colours, indices) copied back to pinned host memory inside the timed region.
--impl reference times the reference's own CPU implementation (oracle/_ref: the unmodified
reference translation units, compiled from /root/reference in the authoring container) on the host cores.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SAMPLERS = {"sphere": 0, "torus_z": 1, "cuboid": 2, "plane_y": 3, "terrain2d": 10, "terrain2d_pert": 11, "terrain3d": 12, "terrain3d_pert": 13}
BASE_OVERLAP = 0.035  # WorldProperties::overlap (WorldOctree.cpp:31)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--sampler", default="terrain2d_pert", choices=sorted(SAMPLERS))
    ap.add_argument("--iters", type=int, default=2)
    ap.add_argument("--dim", type=int, default=64)
    ap.add_argument("--chunks-per-axis", type=int, default=16)
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary measurements (3-D noise, LOD rebuild, single 128^3)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload(args, rank):
    """configs[2]: n^3 grid of size-16 chunks covering 256^3 world units; rank r takes the region shifted by r*256 in x."""
    from binarymeshfitting_b200 import world
    n = args.chunks_per_axis
    size = 256.0 / n
    ps = world.grid_chunks(n, size, origin=(-128.0 + 256.0 * rank, -128.0, -128.0))
    overlap = np.float32(np.float32(BASE_OVERLAP) + np.float32(0.005) * np.float32(args.iters)) if args.iters > 0 else np.float32(BASE_OVERLAP)
    return ps, float(overlap)


def workload_name(args):
    n = args.chunks_per_axis
    return "%d x %d^3 chunks (%dx%dx%d grid, %s, %d smoothing iters)" % (n ** 3, args.dim, n, n, n, args.sampler, args.iters)


# ---------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for k, nme in enumerate(names):
                if f[5 + k].lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------- ours
# algorithmic work of each kernel per voxel / vertex / index (DESIGN.md "Kernels"), used for the roofline line
def kernel_roofline(name, ms, nvox, V, I, n_chunks, dim, peaks, iters=2):
    hbm = peaks["hbm_gbs"]
    words = nvox / 32.0
    algo = {
        # sign words out (+ noise sheet in for the 2-D terrains)
        "k_terrain2d_bits": words * 4 + nvox * 4.0 / dim,
        "k_terrain2d_density": words * 4 + nvox * 4.0 / dim,
        "k_sample_implicit": words * 4,
        "k_pack_density": nvox * 4 + words * 4,
        # bits in, packed counts out
        "k_count<4>": words * 8, "k_count<8>": words * 8,
        # counts in, two bases out
        "k_bases<4>": words * 12, "k_bases<8>": words * 12,
        # 13 B per vertex out (position + boundary flag), two crossing-edge samples in
        "k_verts3": 13.0 * V + 8.0 * V,
        # 4 B per index out + 4 B per-class use counter per vertex; one 8-byte cell record in per ~3 indices
        "k_inds3": 4.0 * I + 4.0 * V + 8.0 * I / 9.0,
        # smoothing, per launch (gathers counted at their algorithmic size, wherever the cache serves them from):
        # dual: 3 indices + vertex base + 3 positions in, 1 centroid out per triangle
        "k_dual<N>": (12 + 4 + 36 + 12) * (I / 3.0),
        # primal: offset + valence + boundary + 12 B out per vertex; 4 B adjacency + 12 B centroid per (vertex, triangle) incidence
        "k_primal": 18.0 * V + 16.0 * I,
        # CSR fill: index + class counters + offset in, adjacency entry out per incidence; cell record + vertex base per cell / triangle
        "k_adj_fill": 16.0 * I + 4.0 * I / 3.0 + 8.0 * I / 9.0,
        # all half-steps of the batch in one launch (one CTA per chunk, iterations out of shared memory): SURVEY 8(d)'s K5
        # figure is per iteration, so the algorithmic bytes are iters x (dual + primal) of the two lines above
        "k_smooth_chunks": iters * ((12 + 4 + 36 + 12) * (I / 3.0) + 18.0 * V + 16.0 * I),
    }
    if name in algo:
        ach = algo[name] / (ms * 1e-3) / 1e9
        out = {"kernel": name, "bound": "hbm", "achieved": round(ach, 1), "peak": hbm, "unit": "GB/s", "frac": round(ach / hbm, 4),
               "traffic": None, "peak_source": peaks["source"], "algorithmic_bytes_per_launch": int(algo[name])}
        if name == "k_smooth_chunks":
            # what has to cross HBM when the iterations stay on chip: positions in and out once, the index / adjacency streams once per half-step
            out["bytes_that_must_cross_hbm"] = int(24.0 * V + iters * (4.0 * I + 6.0 * V + 4.0 * I))
            out["note"] = ("algorithmic bytes = SURVEY 8(d) K5 figure (per-iteration gathers counted at their algorithmic size); the kernel keeps "
                           "positions and dual points in shared memory, so most of those bytes never reach HBM -- see bytes_that_must_cross_hbm and traffic")
        return out
    return None


def load_kernel_profile(workload):
    p = os.path.join(ROOT, "profiles", "r1_kernel_profile.json")
    if not os.path.exists(p):
        return None
    with open(p) as f:
        d = json.load(f)
    return d["kernels"] if d.get("workload") == workload else None


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": float(d["hbm_gbs"]), "source": "of measured (MEASURED_PEAKS.json, burst copy)"}
    return {"hbm_gbs": 6650.0, "source": "of fallback (B200_PROFILING.md 6.65 TB/s)"}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from binarymeshfitting_b200 import Context, capi, world

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    if world_size != args.gpus and world_size > 1:
        args.gpus = world_size
    torch.cuda.set_device(local_rank)
    if world_size > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world_size == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world_size == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    ctx = Context(local_rank)  # raises if the CUDA library or the device is missing: no fallback
    stream = torch.cuda.ExternalStream(ctx.stream_ptr(), device=torch.device("cuda", local_rank))
    kind = SAMPLERS[args.sampler]
    ctx.set_sampler(kind)
    ps, overlap = workload(args, rank)
    descs = capi.make_chunk_descs(ps, overlaps=overlap)
    dim, K, W = args.dim, args.steps, max(args.warmup, 3)
    n_chunks = len(descs)
    nvox = n_chunks * dim ** 3

    def step():
        ctx.submit(descs, dim, iters=args.iters)

    # ---- device-timed throughput (`value`)
    for _ in range(W):
        step()
    ctx.wait()
    launches0 = ctx.launch_count()
    clocks = ClockSampler(local_rank)
    barrier()
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(stream)
    for _ in range(K):
        step()
    e1.record(stream)
    ctx.wait()
    barrier()
    wall = time.perf_counter() - t0
    dev_s = e0.elapsed_time(e1) * 1e-3
    launches = ctx.launch_count() - launches0
    t_max = max_over_ranks(dev_s)
    total_vox = sum_over_ranks(float(nvox)) * K
    value = total_vox / t_max
    stage = ctx.stage_ms()
    _, V, I = ctx.totals()

    # ---- end to end through the C ABI: host descriptors in, renderer-facing SoA meshes out (pinned host memory).
    # Two contexts on the GPU ping-pong (the documented double-buffered use of the ABI): the D2H copies of batch i
    # run on one stream while the kernels of batch i+1 run on the other.  Every step still pays its own H2D
    # (descriptors) and D2H (positions + colours + indices + per-chunk counts) inside the timed region.
    ctxs = [ctx, Context(local_rank)]
    ctxs[1].set_sampler(kind)
    bufs = []
    for _ in range(2):
        pos_h = torch.empty((max(V, 1), 3), dtype=torch.float32).pin_memory()
        col_h = torch.empty((max(V, 1), 3), dtype=torch.float32).pin_memory()
        ind_h = torch.empty((max(I, 1),), dtype=torch.int32).pin_memory()
        bufs.append({"_keep": (pos_h, col_h, ind_h), "pos": pos_h.numpy()[:V], "color": col_h.numpy()[:V], "inds": ind_h.numpy().view(np.uint32)[:I]})

    def e2e_run(steps):
        prev = None
        for i in range(steps):
            c, b = ctxs[i & 1], bufs[i & 1]
            c.submit(descs, dim, iters=args.iters)
            c.download(want=("pos", "color", "inds"), out=b, wait=False)
            if prev is not None:
                prev.wait()
                prev.chunk_infos()
            prev = c
        prev.wait()
        prev.chunk_infos()

    e2e_run(4)
    barrier()
    t0 = time.perf_counter()
    e2e_run(K)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    clk = clocks.stop()  # sampled across both timed regions (device-timed steps and the end-to-end steps)
    e2e_val = total_vox / e2e_s
    out = bufs[(K - 1) & 1]
    h2d = int(descs.nbytes + 16 * n_chunks)          # descriptors (host ABI) + the geometry records the library uploads
    d2h = int(24 * V + 4 * I + 40 * n_chunks + 32)   # positions + colours + indices + per-chunk counts + totals
    checksum = int(out["inds"][: min(I, 1 << 20)].astype(np.uint64).sum()) if I else 0
    # the same without overlap (one context, submit -> download -> wait per step), for reference
    t0 = time.perf_counter()
    for _ in range(max(2, K // 4)):
        ctx.submit(descs, dim, iters=args.iters)
        ctx.download(want=("pos", "color", "inds"), out=bufs[0])
        ctx.chunk_infos()
    e2e_serial_ms = (time.perf_counter() - t0) / max(2, K // 4) * 1e3
    ctxs[1].close()

    # ---- per-kernel times (CUDA events on the launching stream) -> dominant kernel + roofline
    ctx.set_kernel_timing(True)
    agg = {}
    reps = 5
    for _ in range(reps):
        step()
        for name, ms in ctx.kernel_times():
            agg.setdefault(name, []).append(ms)
    ctx.set_kernel_timing(False)
    per_kernel = {k: sum(v) / reps for k, v in agg.items()}
    ktot = sum(per_kernel.values())
    dominant = max(per_kernel, key=per_kernel.get)
    peaks = load_peaks()
    n_launch_dom = len(agg[dominant]) / reps
    launch_ms = per_kernel[dominant] / n_launch_dom
    roof = kernel_roofline(dominant, launch_ms, nvox, V, I, n_chunks, dim, peaks, args.iters)
    if roof is None:
        roof = {"kernel": dominant, "bound": "hbm", "achieved": None, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": None, "traffic": None,
                "peak_source": peaks["source"], "note": "no HBM model for this kernel (FP32/INT issue bound): see the `issue` view"}
    roof["share_of_step"] = round(per_kernel[dominant] / ktot, 4)
    roof["launch_ms"] = round(launch_ms, 4)
    # The kernels of this path are integer / FP32 ISSUE bound, not HBM bound (north_star: "FP32/INT pipe utilisation for
    # the noise ... stages").  Issue view of the dominant kernel: warp instructions per launch (counted by ncu on this very
    # workload, profiles/r1_kernel_profile.json) / live CUDA-event time, against 4 schedulers x SMs x the SM clock sampled
    # during the timed region.
    prof = load_kernel_profile(workload_name(args))
    base = dominant.split("<")[0]
    if prof and base in prof and prof[base].get("warp_inst") and clk.get("sm_mhz"):
        kp = prof[base]
        sms = torch.cuda.get_device_properties(local_rank).multi_processor_count
        peak_inst = sms * 4 * clk["sm_mhz"] * 1e6
        ach_inst = kp["warp_inst"] / (launch_ms * 1e-3)
        roof["issue"] = {"bound": "issue", "achieved": round(ach_inst / 1e9, 1), "peak": round(peak_inst / 1e9, 1), "unit": "Gwarp-inst/s",
                         "frac": round(ach_inst / peak_inst, 4), "warp_inst_per_launch": int(kp["warp_inst"]),
                         "ncu_issue_active_pct": kp.get("issue_active_pct"), "ncu_alu_pipe_pct": kp.get("alu_pipe_pct"), "ncu_fma_pipe_pct": kp.get("fma_pipe_pct"),
                         "source": "profiles/r1_kernel_profile.json (ncu) + live event time + live SM clock"}
        if kp.get("dram_read_bytes") is not None:
            roof["traffic"] = int(kp["dram_read_bytes"] + kp["dram_write_bytes"])  # dram__bytes_read.sum + dram__bytes_write.sum per launch (ncu)
    # the HBM view of every kernel of the step that has a byte model, each against the measured copy peak
    hbm_lines = {}
    for name, ms in per_kernel.items():
        r = kernel_roofline(name, ms / (len(agg[name]) / reps), nvox, V, I, n_chunks, dim, peaks, args.iters)
        if r:
            hbm_lines[name] = {"ms": round(ms, 4), "GB/s": r["achieved"], "frac": r["frac"]}
    # whole step against SURVEY 8(d)'s fused sample->mesh figure: N/8 (sign words) + N (cell masks) + 14 V + 4 I bytes
    step_bytes = nvox / 8 + nvox + 14.0 * V + 4.0 * I
    step_roof = {"algorithmic_bytes": int(step_bytes), "achieved": round(step_bytes / (t_max / K) / 1e9, 1), "peak": peaks["hbm_gbs"], "unit": "GB/s",
                 "frac": round(step_bytes / (t_max / K) / 1e9 / peaks["hbm_gbs"], 4),
                 "note": "SURVEY 8(d) fused figure (counts N bytes of cell masks the device path never writes)"}
    # the same step against SURVEY 8(d)'s density-in -> mesh-out figure (4N + N/8 + N + 14V + 4I; the figure behind the
    # survey's "1.13 Tvoxel/s HBM roofline, 60 % target"), although this fused path never reads a density block
    din = 5.125 * nvox + 14.0 * V + 4.0 * I
    step_roof["density_in_figure"] = {"algorithmic_bytes": int(din), "hbm_roofline_voxels_per_s": nvox / (din / (peaks["hbm_gbs"] * 1e9)),
                                      "frac": round((nvox / (t_max / K)) / (nvox / (din / (peaks["hbm_gbs"] * 1e9))), 4)}

    result = {
        "metric": "voxels/sec sampled+meshed", "value": value, "unit": "voxels/s", "n_gpus": args.gpus, "steps": K, "warmup": W,
        "ms_per_step": t_max / K * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+u32",
        "data": "synthetic (procedural noise terrain, seed 1337; no dataset)",
        "config": {"workload": workload_name(args), "chunks_per_gpu": n_chunks, "dim": dim, "sampler": args.sampler, "iters": args.iters,
                   "overlap": overlap, "l2": "working set per step (bits+counts+bases %.0f MB + meshes) exceeds the 126 MB L2; no explicit flush" % (nvox / 8 * 4 / 1e6),
                   "partition": "one 4096-chunk region per GPU, no collective"},
        "chunks_per_s": value / dim ** 3,
        "wall_ms_per_step": wall / K * 1e3,
        "e2e": {"value": e2e_val, "unit": "voxels/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_s / K * 1e3,
                "checksum": checksum, "mode": "2 contexts ping-pong (D2H of batch i overlaps kernels of batch i+1)", "serial_ms_per_step": e2e_serial_ms},
        "gpu_launches": int(launches),
        "clocks": clk,
        "roofline": roof,
        "step_roofline": step_roof,
        "kernels_ms": {k: round(v, 4) for k, v in sorted(per_kernel.items(), key=lambda kv: -kv[1])},
        "hbm_kernels": hbm_lines,
        "stage_ms": {k: round(v, 4) for k, v in stage.items()},
        "mesh": {"verts": int(V), "indices": int(I)},
    }

    if rank == 0 and args.gpus == 1 and not args.no_extras:
        result["extras"] = extras(ctx, args, capi, world)
    if rank == 0 and args.gpus == 1 and not args.no_cpu_baseline:
        result["cpu_baseline"] = cpu_baseline(args, ps, overlap)
    ctx.close()
    if world_size > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(result))


def timed_density(ctx, descs, dim, iters, dptr, reps=5):
    import time as _t
    for _ in range(2):
        ctx.submit(descs, dim, iters=iters, density_device_ptr=dptr)
    ctx.wait()
    t0 = _t.perf_counter()
    for _ in range(reps):
        ctx.submit(descs, dim, iters=iters, density_device_ptr=dptr)
    ctx.wait()
    return (_t.perf_counter() - t0) / reps


def extras(ctx, args, capi, world):
    """Secondary numbers the metric names: 3-D fractal noise, ms per LOD rebuild, the single 128^3 chunk."""
    import time as _t
    ex = {}

    def timed(descs, dim, iters, reps=5):
        for _ in range(2):
            ctx.submit(descs, dim, iters=iters)
        ctx.wait()
        t0 = _t.perf_counter()
        for _ in range(reps):
            ctx.submit(descs, dim, iters=iters)
        ctx.wait()
        return (_t.perf_counter() - t0) / reps

    ps, overlap = workload(args, 0)
    d = capi.make_chunk_descs(ps, overlaps=overlap)
    ctx.set_sampler(capi.TERRAIN3D_PERT)
    s = timed(d, args.dim, args.iters, reps=3)
    ex["terrain3d_pert_4096x64"] = {"ms_per_step": s * 1e3, "voxels_per_s": len(d) * args.dim ** 3 / s, "stage_ms": ctx.stage_ms()}
    # the noise stage is FP32 / INT32 issue bound and its denominators are to be MEASURED: dependent-free chains
    # of FFMA / IMAD / (2 LOP3 + LEA) per thread, 2048 threads per SM (csrc/smooth.cuh k_ubench_issue); 1e9 thread-level steps/s
    try:
        pk = ctx.ubench_issue()
        sass_per_step = {"fp32_fma": 1, "int32_imad": 1, "int32_logic_step": 3, "fma_plus_logic_step": 2}  # SASS instructions per chain step (cuobjdump)
        peaks_inst = {k: v * sass_per_step[k] for k, v in pk.items()}  # 1e9 thread-level instructions/s
        inst_per_voxel = 2168.0  # ncu smsp__thread_inst_executed of k_terrain3d<NT_SIMPLEX> per voxel (profiles/r1_ncu_full_summary_v1.txt)
        ach = ex["terrain3d_pert_4096x64"]["voxels_per_s"] * inst_per_voxel / 1e9
        ex["issue_peaks_measured"] = {"chain_steps_G_per_s": pk, "thread_inst_G_per_s": peaks_inst, "unit": "1e9 thread-level operations per second, whole GPU",
                                      "nominal_fp32_fma_G_per_s": 148 * 128 * 1.965}
        ex["terrain3d_pert_4096x64"]["issue"] = {"thread_inst_G_per_s": ach, "inst_per_voxel": inst_per_voxel,
                                                 "frac_of_measured_fma_peak": ach / peaks_inst["fp32_fma"],
                                                 "note": "all instructions of the kernel against the measured full-rate (FFMA) issue peak; INT32 logic / IMAD issue at half "
                                                         "that rate on B200 (measured above), and ncu puts this kernel's ALU (INT) pipe at 80 % and its FMA pipe at 41 % busy: "
                                                         "it is bound by INT32 issue"}
    except Exception as e:  # noqa: BLE001
        ex["issue_peaks_measured"] = {"error": str(e)}
    # config 4: LOD world, 2048^3 effective voxels at the finest level (dim 64, max_level 5) = 232 leaves
    props = world.WorldProperties(max_level=5, chunk_resolution=64, process_iters=2)
    lps, lv, mc = world.split_leaves(props)
    ld = world.make_descs(props, lps, lv, mc)
    for name, kind in (("terrain2d_pert", capi.TERRAIN2D_PERT), ("terrain3d_pert", capi.TERRAIN3D_PERT)):
        ctx.set_sampler(kind)
        s = timed(ld, 64, 2, reps=10)
        ex["lod_rebuild_232x64_%s" % name] = {"ms": s * 1e3, "voxels_per_s": len(ld) * 64 ** 3 / s}
    # config 4 as named ("multi-level chunks + WorldStitcher seams"): the same world sampled at voxel-node centres, chunk
    # meshes + 2 smoothing iterations, then the seam pass (bmf_batch_stitch); wall time incl. its one host round trip
    try:
        sd = capi.make_chunk_descs(lps, overlaps=ctx.seam_overlap(64), levels=lv)
        for name, kind in (("terrain2d_pert", capi.TERRAIN2D_PERT), ("terrain3d_pert", capi.TERRAIN3D_PERT)):
            ctx.set_sampler(kind)
            walls, nt = [], 0
            for r in range(6):
                t0 = _t.perf_counter()
                ctx.submit(sd, 64, iters=2)
                nt = ctx.stitch(download=False)
                walls.append(_t.perf_counter() - t0)
            ex["lod_rebuild_with_seams_232x64_%s" % name] = {"ms": min(walls[1:]) * 1e3, "seam_tris": nt, "seam_device_ms": ctx.seam_ms(),
                                                            "mesh": dict(zip(("cells", "verts", "indices"), ctx.totals()))}
        # seams of the uniform benchmark grid (every chunk border of the 16^3 grid)
        ctx.set_sampler(SAMPLERS[args.sampler])
        gd = capi.make_chunk_descs(ps, overlaps=ctx.seam_overlap(args.dim))
        ctx.submit(gd, args.dim, iters=args.iters)
        nt = ctx.stitch(download=False)
        nt = ctx.stitch(download=False)
        sm = ctx.seam_ms()
        pts = len(gd) * (6 * args.dim ** 2 + 2)
        ex["seam_pass_4096x64_%s" % args.sampler] = {"seam_tris": nt, "device_ms": sm, "lattice_points": pts,
                                                     "lattice_points_per_s": pts / ((sm["count"] + sm["emit"]) * 1e-3)}
    except Exception as e:  # noqa: BLE001
        ex["lod_rebuild_with_seams"] = {"error": str(e)}
    # "ms per LOD rebuild", incremental reading: the WorldWatcher tick (world.LodWatcher) while the focus flies 4 units per
    # tick along x; every tick's batch (8 children per split, the parent per group) is meshed with 2 smoothing iterations
    try:
        ctx.set_sampler(SAMPLERS[args.sampler])
        wprops = world.WorldProperties(max_level=5, chunk_resolution=64, process_iters=2)
        lw = world.LodWatcher(wprops, 256, (0.0, 0.0, 0.0))
        tick_ms, sizes = [], []
        for k in range(1, 61):
            gen = lw.tick((4.0 * k, 0.0, 0.0))
            if not gen:
                continue
            gps, glv, gmc = lw.arrays(gen)
            gd = world.make_descs(wprops, gps, glv, gmc)
            t0 = _t.perf_counter()
            ctx.submit(gd, 64, iters=2)
            ctx.wait()
            tick_ms.append((_t.perf_counter() - t0) * 1e3)
            sizes.append(len(gd))
        if tick_ms:
            ex["lod_incremental_fly_%s" % args.sampler] = {"ticks_with_work": len(tick_ms), "chunks_per_tick_mean": sum(sizes) / len(sizes), "chunks_per_tick_max": max(sizes),
                                                           "ms_per_tick_median": sorted(tick_ms)[len(tick_ms) // 2], "ms_per_tick_max": max(tick_ms),
                                                           "note": "submit + wait per tick (host wall clock); the tick policy itself is host code and not timed"}
    except Exception as e:  # noqa: BLE001
        ex["lod_incremental_fly"] = {"error": str(e)}
    # config 1: single 64^3 chunk of the implicit sphere, no processing -- triangles (the reference's emitter) and quads (bmf_params.quads)
    try:
        ctx.set_sampler(capi.SPHERE)
        c1 = capi.make_chunk_descs([[-128, -128, -128, 256.0]])
        for name, qd in (("tris", False), ("quads", True)):
            for _ in range(3):
                ctx.submit(c1, 64, iters=0, quads=qd)
            ctx.wait()
            t0 = _t.perf_counter()
            for _ in range(20):
                ctx.submit(c1, 64, iters=0, quads=qd)
            ctx.wait()
            s = (_t.perf_counter() - t0) / 20
            ex["single_64_sphere_%s" % name] = {"ms": s * 1e3, "mesh": dict(zip(("cells", "verts", "indices"), ctx.totals()))}
        # quads on the benchmark batch (emission only)
        ctx.set_sampler(SAMPLERS[args.sampler])
        for _ in range(2):
            ctx.submit(d, args.dim, iters=0, quads=True)
        ctx.wait()
        t0 = _t.perf_counter()
        for _ in range(3):
            ctx.submit(d, args.dim, iters=0, quads=True)
        ctx.wait()
        s = (_t.perf_counter() - t0) / 3
        ex["quads_4096x64_%s" % args.sampler] = {"ms": s * 1e3, "voxels_per_s": len(d) * args.dim ** 3 / s, "stage_ms": ctx.stage_ms(),
                                                 "mesh": dict(zip(("cells", "verts", "indices"), ctx.totals()))}
    except Exception as e:  # noqa: BLE001
        ex["single_64_sphere_quads"] = {"error": str(e)}
    # config 2: single 128^3 chunk, 2 smoothing iterations
    one = capi.make_chunk_descs([[-64, -64, -64, 128.0]], overlaps=0.045)
    for name, kind in (("terrain2d_pert", capi.TERRAIN2D_PERT), ("terrain3d_pert", capi.TERRAIN3D_PERT)):
        ctx.set_sampler(kind)
        s = timed(one, 128, 2, reps=20)
        ex["single_128_%s" % name] = {"ms": s * 1e3, "voxels_per_s": 128 ** 3 / s}
    # config 4, dense reading: the whole 2048^3-voxel world at the finest LOD = 32x32x32 chunks of 64^3 (8.6 Gvoxel)
    try:
        ctx.set_sampler(SAMPLERS[args.sampler])
        dd = capi.make_chunk_descs(world.grid_chunks(32, 16.0, origin=(-256.0, -256.0, -256.0)), overlaps=overlap)
        s = timed(dd, 64, args.iters, reps=3)
        ex["dense_2048_cubed_%s" % args.sampler] = {"chunks": len(dd), "ms": s * 1e3, "voxels_per_s": len(dd) * 64 ** 3 / s, "mesh": dict(zip(("cells", "verts", "indices"), ctx.totals()))}
    except Exception as e:  # noqa: BLE001
        ex["dense_2048_cubed_%s" % args.sampler] = {"error": str(e)}
    # K2 on its own: the density block of the benchmark workload (4.3 GB, resident in HBM) -> sign words.  This is the
    # path's one pure streaming kernel (HOST_DENSITY / staged label_grid); its roofline is the measured copy bandwidth.
    try:
        ctx.set_sampler(capi.TERRAIN2D_PERT)
        ctx.submit(d, args.dim, iters=0, keep_density=True)
        ctx.wait()
        dptr = ctx.device_ptrs()["density"]
        ctx.set_sampler(capi.HOST_DENSITY)
        ctx.set_kernel_timing(True)
        best = None
        for _ in range(5):
            ctx.submit(d, args.dim, iters=0, density_device_ptr=dptr)
            t = [ms for name, ms in ctx.kernel_times() if name == "k_pack_density"]
            best = t[0] if best is None else min(best, t[0])
        ctx.set_kernel_timing(False)
        nb = len(d) * args.dim ** 3 * (4 + 1.0 / 8)
        peaks = load_peaks()
        ex["k2_pack_density_4096x64"] = {"ms": best, "algorithmic_bytes": int(nb), "GB/s": nb / best / 1e6, "peak": peaks["hbm_gbs"],
                                         "frac": nb / best / 1e6 / peaks["hbm_gbs"], "peak_source": peaks["source"], "bound": "hbm"}
        # the whole density-in -> mesh-out pipeline (K2-K5) on that resident block, against SURVEY 8(d)'s byte figure
        # 4N + N/8 + N + 14V + 4I (its "HBM roofline 1.13 Tvoxel/s" for the sphere; recomputed here for this mesh)
        s = timed_density(ctx, d, args.dim, args.iters, dptr)
        _, V2, I2 = ctx.totals()
        nvox2 = len(d) * args.dim ** 3
        by = 5.125 * nvox2 + 14.0 * V2 + 4.0 * I2
        ex["density_in_mesh_out_4096x64"] = {"ms": s * 1e3, "voxels_per_s": nvox2 / s, "algorithmic_bytes": int(by), "GB/s": by / s / 1e9,
                                             "frac": by / s / 1e9 / peaks["hbm_gbs"], "hbm_roofline_voxels_per_s": nvox2 / (by / (peaks["hbm_gbs"] * 1e9))}
    except Exception as e:  # noqa: BLE001 -- an extra must never take the headline down
        ex["k2_pack_density_4096x64"] = {"error": str(e)}
    ctx.set_sampler(SAMPLERS[args.sampler])
    return ex


# ---------------------------------------------------------------------------------------------- CPU reference
def _ref_worker(kind, dim, iters, ps, threads, steps, warmup, conn):
    """One worker process: the reference's ChunkGenerator::process_queue over its share with <= 8 OMP threads."""
    from oracle import ref_binding as rb
    R = rb.RefLib()
    w = R.world(kind, dim, max_level=99, iters=iters, boundary_processing=False)  # max_level 99: no chunk is "at max level" -> overlap rule = base + 0.005*iters
    w.add_chunks(ps, np.zeros(len(ps), np.int32))
    conn.send("ready")
    while True:
        msg = conn.recv()
        if msg == "stop":
            break
        ms = w.process(threads)
        conn.send(ms)
    nm, nv, ni = w.totals()
    conn.send((nm, nv, ni))


class RefPool:
    """All host cores: floor(cores/8) worker processes x 8 OpenMP threads (8 is the reference's hard limit:
    Sampler::noise_samplers[8] indexed by omp_get_thread_num(), Sampler.hpp:30 / DMCChunk.cpp:109)."""

    def __init__(self, kind, dim, iters, ps, max_procs=None):
        import multiprocessing as mp
        cores = os.cpu_count() or 8
        self.threads = min(8, cores)
        self.nproc = max(1, cores // 8)
        if max_procs:
            self.nproc = min(self.nproc, max_procs)
        self.nproc = min(self.nproc, max(1, len(ps) // 8))
        ctxm = mp.get_context("spawn")
        self.conns, self.procs = [], []
        shares = np.array_split(np.arange(len(ps)), self.nproc)
        for sh in shares:
            a, b = ctxm.Pipe()
            p = ctxm.Process(target=_ref_worker, args=(kind, dim, iters, ps[sh], self.threads, 0, 0, b), daemon=True)
            p.start()
            self.conns.append(a)
            self.procs.append(p)
        for c in self.conns:
            c.recv()

    @property
    def cores(self):
        return self.threads * self.nproc

    def step(self):
        t0 = time.perf_counter()
        for c in self.conns:
            c.send("go")
        for c in self.conns:
            c.recv()
        return time.perf_counter() - t0

    def close(self):
        tot = [0, 0, 0]
        for c in self.conns:
            c.send("stop")
        for c in self.conns:
            r = c.recv()
            tot = [a + b for a, b in zip(tot, r)]
        for p in self.procs:
            p.join(timeout=5)
        return tot


def bounded_sample(ps, budget_chunks):
    """strided subset of the workload (keeps the mesh / no-mesh mix) of at most budget_chunks chunks"""
    stride = max(1, int(np.ceil(len(ps) / budget_chunks)))
    return ps[::stride], stride


def cpu_baseline(args, ps, overlap):
    from oracle import ref_binding as rb
    if not rb.available():
        return {"value": None, "unit": "voxels/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref/libbmf_ref.so missing"}
    kind = SAMPLERS[args.sampler]
    three_d = args.sampler.startswith("terrain3d")
    sample, stride = bounded_sample(ps, 512 if three_d else 4096)
    pool = RefPool(kind, args.dim, args.iters, sample)
    pool.step()  # warm-up (pools allocate)
    best = min(pool.step() for _ in range(3))
    cores = pool.cores
    tot = pool.close()
    nv = len(sample) * args.dim ** 3
    # the reference as shipped is limited to 8 OpenMP threads (noise_samplers[8]); report that number too
    eight = None
    if cores > 8:
        p8 = RefPool(kind, args.dim, args.iters, sample, max_procs=1)
        p8.step()
        eight = nv / min(p8.step() for _ in range(2))
        p8.close()
    cpu_model = ""
    try:
        with open("/proc/cpuinfo") as f:
            cpu_model = next((l.split(":", 1)[1].strip() for l in f if l.startswith("model name")), "")
    except OSError:
        pass
    return {"value": nv / best, "unit": "voxels/s", "cores": cores, "kind": "reference", "cpu_model": cpu_model, "value_8_threads": eight,
            "sample": "every %d-th chunk of the workload (%d chunks), 1 warm-up + best of 3, %d process(es) x %d OMP threads; noise = scalar restatement of FastNoiseSIMD" %
                      (stride, len(sample), pool.nproc, pool.threads),
            "ms": best * 1e3, "chunks_per_s": len(sample) / best, "mesh": {"chunks_with_mesh": tot[0], "verts": tot[1], "indices": tot[2]}}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import ref_binding as rb
    if not rb.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libbmf_ref.so missing (built from /root/reference in the authoring container)"}))
        return
    ps, overlap = workload(args, 0)
    kind = SAMPLERS[args.sampler]
    three_d = args.sampler.startswith("terrain3d")
    sample, stride = bounded_sample(ps, 512 if three_d else 4096)
    pool = RefPool(kind, args.dim, args.iters, sample)
    K, W = args.steps, max(args.warmup, 1)
    for _ in range(W):
        pool.step()
    t = [pool.step() for _ in range(K)]
    total = sum(t)
    cores = pool.cores
    tot = pool.close()
    nv = len(sample) * args.dim ** 3
    value = nv * K / total
    print(json.dumps({
        "impl": "reference", "metric": "voxels/sec sampled+meshed", "value": value, "unit": "voxels/s", "n_gpus": args.gpus, "steps": K, "warmup": W,
        "ms_per_step": total / K * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+u32",
        "data": "synthetic (procedural noise terrain, seed 1337; no dataset)",
        "config": {"workload": workload_name(args), "chunks_per_gpu": len(ps), "dim": args.dim, "sampler": args.sampler, "iters": args.iters, "overlap": overlap},
        "chunks_per_s": value / args.dim ** 3,
        "cpu_baseline": {"value": value, "unit": "voxels/s", "cores": cores, "kind": "reference",
                         "sample": "every %d-th chunk of the workload (%d chunks) per step, %d process(es) x %d OMP threads (8 = the reference's thread limit); "
                                   "noise = scalar restatement of FastNoiseSIMD" % (stride, len(sample), pool.nproc, pool.threads)},
        "e2e": {"value": value, "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "mesh": {"chunks_with_mesh": tot[0], "verts": tot[1], "indices": tot[2]},
    }))


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)

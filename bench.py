#!/usr/bin/env python
"""bench.py -- headline benchmark of the chunk-extraction hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--sampler NAME] [--iters I]

Workload (config.workload): BASELINE.json configs[2] -- ONE world: a 16x16x16 grid of 4096 chunks of 64^3 voxels
(1.07 Gvoxel) of noise terrain, full pipeline sample -> sign bits -> cell masks -> vertex/index emission -> 2
MeshProcessor<3> smoothing iterations.  One "step" = one ChunkGenerator::process_queue of that batch.

N > 1 (one process per GPU, torchrun): STRONG scaling of that one world.  world.partition() deals contiguous ranges of
the chunks' Z-curve (Morton) order to the ranks; every rank meshes its share; no data-path collective.  `value` =
the world's voxels / max-over-ranks device time.  `e2e` = host descriptors in, every rank's GPU storing its meshes
straight into its region of a shared, pinned host segment (binarymeshfitting_b200/gather.py), rank 0 assembling the
batch-order chunk table over it -- the "final host gather of per-chunk meshes" -- inside the timed region; rank 0 then
checks (outside the timed region) that the gathered batch hashes like the single-GPU golden of the compiled reference.
extras.weak keeps round 1's replica run (every rank its own 4096-chunk region); extras also carries the LOD world with
cross-rank seams and the dense 2048^3 world.

Prints ONE JSON line (rank 0).  --impl reference times the reference's own CPU implementation (oracle/_ref: the unmodified
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
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SAMPLERS = {"sphere": 0, "torus_z": 1, "cuboid": 2, "plane_y": 3, "terrain2d": 10, "terrain2d_pert": 11, "terrain3d": 12, "terrain3d_pert": 13}
BASE_OVERLAP = 0.035  # WorldProperties::overlap (WorldOctree.cpp:31)
PROFILE_JSON = os.path.join(ROOT, "profiles", "r2_kernel_profile.json")


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
    ap.add_argument("--in-flight", type=int, default=0, help="contexts (streams + arenas) the steps are dealt to round robin: batches in flight at once (>= 3); "
                    "default 4, 6 from four ranks up (a rank's share of the world is then a chain of short kernels: 0.117 / 0.094 ms per 400-chunk batch with 4 / 6 in flight)")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary measurements (3-D noise, LOD rebuild, single 128^3, ...)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--reserve-sms", type=int, default=0, help="e2e only: bmf_ctx_set_reserved_sms for both contexts (measured: no effect with the copy-engine download)")
    return ap.parse_args()


def workload(args, region=0):
    """configs[2]: n^3 grid of size-16 chunks covering 256^3 world units (region r: the same grid shifted by r*256 in x)."""
    from binarymeshfitting_b200 import world
    n = args.chunks_per_axis
    size = 256.0 / n
    ps = world.grid_chunks(n, size, origin=(-128.0 + 256.0 * region, -128.0, -128.0))
    overlap = np.float32(np.float32(BASE_OVERLAP) + np.float32(0.005) * np.float32(args.iters)) if args.iters > 0 else np.float32(BASE_OVERLAP)
    return ps, float(overlap)


def workload_name(args):
    n = args.chunks_per_axis
    return "%d x %d^3 chunks (%dx%dx%d grid, %s, %d smoothing iters)" % (n ** 3, args.dim, n, n, n, args.sampler, args.iters)


def config_of(args, overlap):
    """identical in both arms (the driver compares them)"""
    return {"workload": workload_name(args), "chunks": args.chunks_per_axis ** 3, "dim": args.dim, "sampler": args.sampler, "iters": args.iters, "overlap": overlap}


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


# ---------------------------------------------------------------------------------------------- roofline accounting
def kernel_bytes(name, st):
    """ALGORITHMIC bytes of one launch of a kernel, counted over the chunks the launch actually PROCESSES:
    st = dict(dim, n_chunks, n_mesh (chunks that contain a mesh), n_sampled (chunks whose sign words are written), V, I, iters).
    Returns None for the issue-bound kernels (noise sheets, 3-D noise), which have no HBM model."""
    d = st["dim"]
    vox_mesh = st["n_mesh"] * d ** 3
    vox_sampled = st["n_sampled"] * d ** 3
    V, I, it = float(st["V"]), float(st["I"]), st["iters"]
    w = 4.0 / 32.0  # bytes of sign word per voxel
    m = {
        # sign words out + the noise sheet row in (4 B per 32 voxels), non-uniform chunks only
        "k_terrain2d_bits": vox_sampled * 2 * w,
        "k_terrain2d_density": vox_sampled * (w + 4.0),
        "k_sample_implicit": vox_sampled * w,
        "k_pack_density": vox_sampled * (4.0 + w),
        # mesh chunks only: sign words in, packed counts out
        "k_count": vox_mesh * 2 * w,
        # sign words + counts in, index bases out (+ 16 B vertex record per active word, + 8 B record per surface cell: ~ V/1.5 + I/9 cells)
        "k_bases": vox_mesh * 3 * w + 8.0 * (V / 1.5 + I / 9.0),
        # per-chunk front end, dim <= 64 (csrc/fused.cuh).  count: sign words in, packed word counts out.  emit: sign words + word counts in; out: positions 12 B,
        # zeroed normals 12 B, boundary 1 B, valence 1 B, adjacency offsets 4 B per vertex (+ two crossing-edge samples in, 8 B), indices 4 B + CSR 4 B per index,
        # prim_vbase 4 B per triangle, and an 8-byte cell record written and read back per surface cell (~ V/1.5 vertex cells + I/9 index cells)
        "k_chunk_count": vox_mesh * 2 * w,
        "k_chunk_emit": vox_mesh * 2 * w + (12.0 + 12.0 + 1.0 + 1.0 + 4.0 + 8.0) * V + 8.0 * I + 4.0 * I / 3.0 + 16.0 * (V / 1.5 + I / 9.0),
        # 13 B per vertex out (position + boundary flag), two crossing-edge samples in
        "k_verts3": (13.0 + 8.0) * V,
        # 4 B per index out + 4 B use counter per vertex + one 8-byte cell record per ~9 indices
        "k_inds3": 4.0 * I + 4.0 * V + 8.0 * I / 9.0,
        "k_valence_offsets": 9.0 * V,
        "k_adj_fill": 16.0 * I + 4.0 * I / 3.0 + 8.0 * I / 9.0,
        "k_dual": (12 + 4 + 36 + 12) * (I / 3.0),
        "k_primal": 18.0 * V + 16.0 * I,
        # all half-steps in one launch out of shared memory: what MUST cross HBM is positions in and out once (24 V) and the index / adjacency /
        # valence streams once per half-step
        "k_smooth_chunks": 24.0 * V + it * (8.0 * I + 6.0 * V),
    }
    return m.get(name.split("<")[0])


def load_kernel_profile(workload):
    if not os.path.exists(PROFILE_JSON):
        return None
    with open(PROFILE_JSON) as f:
        d = json.load(f)
    return d["kernels"] if d.get("workload") == workload else None


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": float(d["hbm_gbs"]), "source": "of measured (MEASURED_PEAKS.json, burst copy)"}
    return {"hbm_gbs": 6650.0, "source": "of fallback (B200_PROFILING.md 6.65 TB/s)"}


def kernel_report(ctx, submit, st, peaks, prof, clk, sms, reps=5):
    """per-kernel live CUDA-event times of one step + HBM view (algorithmic bytes over PROCESSED chunks, ncu DRAM traffic beside it)
    + issue view (ncu warp instructions / live time against 4 schedulers x SMs x live SM clock)"""
    ctx.set_kernel_timing(True)
    agg = {}
    for _ in range(reps):
        submit()
        for name, ms in ctx.kernel_times():
            agg.setdefault(name, []).append(ms)
    ctx.set_kernel_timing(False)
    per_kernel = {k: sum(v) / reps for k, v in agg.items()}
    ktot = sum(per_kernel.values())
    lines = {}
    for name, ms in sorted(per_kernel.items(), key=lambda kv: -kv[1]):
        n_launch = len(agg[name]) / reps
        launch_ms = ms / n_launch
        e = {"ms_per_step": round(ms, 4), "launches_per_step": n_launch, "share_of_step": round(ms / ktot, 4)}
        b = kernel_bytes(name, st)
        if b is not None:
            e["algorithmic_bytes"] = int(b)
            e["GB/s"] = round(b / (launch_ms * 1e-3) / 1e9, 1)
            e["frac_of_hbm_peak"] = round(b / (launch_ms * 1e-3) / 1e9 / peaks["hbm_gbs"], 4)
        kp = prof.get(name.split("<")[0]) if prof else None
        if kp:
            if kp.get("dram_read_bytes") is not None:
                e["ncu_dram_traffic_bytes"] = int(kp["dram_read_bytes"] + kp["dram_write_bytes"])
                if b:
                    e["traffic_over_algorithmic"] = round(e["ncu_dram_traffic_bytes"] / b, 2)
            if kp.get("warp_inst") and clk.get("sm_mhz"):
                peak_inst = sms * 4 * clk["sm_mhz"] * 1e6
                e["issue_frac"] = round(kp["warp_inst"] / (launch_ms * 1e-3) / peak_inst, 4)
                e["ncu_issue_active_pct"] = kp.get("issue_active_pct")
                e["ncu_alu_pipe_pct"], e["ncu_fma_pipe_pct"] = kp.get("alu_pipe_pct"), kp.get("fma_pipe_pct")
        lines[name] = e
    return per_kernel, lines


def roofline_of(dominant, lines, peaks):
    e = lines[dominant]
    r = {"kernel": dominant, "bound": "hbm", "achieved": e.get("GB/s"), "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": e.get("frac_of_hbm_peak"),
         "traffic": e.get("ncu_dram_traffic_bytes"), "peak_source": peaks["source"], "algorithmic_bytes_per_launch": e.get("algorithmic_bytes"),
         "share_of_step": e["share_of_step"], "launch_ms": round(e["ms_per_step"] / e["launches_per_step"], 4),
         "accounting": "algorithmic bytes counted over the chunks the launch processes (mesh-containing chunks only), see kernels[] for every kernel of the step "
                       "with ncu DRAM traffic beside the model"}
    if e.get("issue_frac") is not None:
        r["issue"] = {"bound": "issue", "frac": e["issue_frac"], "ncu_issue_active_pct": e.get("ncu_issue_active_pct"), "ncu_alu_pipe_pct": e.get("ncu_alu_pipe_pct"),
                      "ncu_fma_pipe_pct": e.get("ncu_fma_pipe_pct"), "source": os.path.relpath(PROFILE_JSON, ROOT) + " (ncu) + live event time + live SM clock"}
    if r["frac"] is None:
        r["note"] = "no HBM model for this kernel (FP32/INT issue bound): see the `issue` view"
    return r


# ---------------------------------------------------------------------------------------------- ours
class Dist:
    def __init__(self):
        import torch
        self.torch = torch
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.size = int(os.environ.get("WORLD_SIZE", "1"))
        torch.cuda.set_device(self.local_rank)
        if self.size > 1:
            import torch.distributed as dist
            self.dist = dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))

    def barrier(self):
        if self.size > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def _reduce(self, x, op):
        if self.size == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def max(self, x):
        return self._reduce(x, self.dist.ReduceOp.MAX) if self.size > 1 else x

    def sum(self, x):
        return self._reduce(x, self.dist.ReduceOp.SUM) if self.size > 1 else x

    def allgather(self, obj):
        if self.size == 1:
            return [obj]
        out = [None] * self.size
        self.dist.all_gather_object(out, obj)
        return out

    def close(self):
        if self.size > 1:
            self.dist.barrier()
            self.dist.destroy_process_group()


def pin_to_gpu_numa_node(local_rank):
    """best effort: run this rank's host threads (and first-touch its host buffers) on the NUMA node its GPU hangs off"""
    try:
        import torch
        p = torch.cuda.get_device_properties(local_rank)
        bus = "%04x:%02x:%02x.0" % (getattr(p, "pci_domain_id", 0), p.pci_bus_id, p.pci_device_id)
        with open("/sys/bus/pci/devices/%s/numa_node" % bus) as f:
            node = int(f.read().strip())
        if node < 0:
            return "numa_node unknown (-1)"
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpus = set()
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        os.sched_setaffinity(0, cpus)
        return "node %d (%d cpus)" % (node, len(cpus))
    except Exception as e:  # noqa: BLE001
        return "not pinned (%s)" % type(e).__name__


class SharedWorld:
    """ONE batch of chunks partitioned over the ranks + the shared host segment its meshes are gathered in."""

    def __init__(self, D, ctxs, descs_all, mortons, dim, iters, tag, mode="compact", columns=False):
        from binarymeshfitting_b200 import gather, world
        self.D, self.ctxs, self.dim, self.iters, self.mode = D, ctxs, dim, iters, mode
        self.descs_all = descs_all
        self.n_total = len(descs_all)
        # columns: heightfield sampler -> ranges of the column-major curve (whole chunk columns per rank: no noise sheet is sampled twice)
        self.columns = columns
        parts = world.partition(mortons, np.ones(self.n_total), D.size, columns) if D.size > 1 else [np.arange(self.n_total)]
        c = ctxs[0]
        self.cost_model = None
        if D.size > 1:
            # cost-balanced ranges ("cost ~ measured surface count from the previous rebuild"): every rank meshes its equal-count
            # share once, the per-chunk vertex counts are exchanged, and the Z-curve is cut again by cost = EMPTY + vertices, where EMPTY = what an
            # empty chunk costs in vertex units (fit of the 4096- and 32768-chunk single-GPU steps: 7.8 ns per chunk, 0.21 ns per vertex)
            mine0 = np.sort(parts[D.rank])
            nv0 = np.zeros(0, np.int64)
            if len(mine0):
                c.submit(np.ascontiguousarray(descs_all[mine0]), dim, iters=iters)
                nv0 = c.chunk_infos()["n_verts"].astype(np.int64)
            # ... for shares large enough for the per-chunk kernels on their own; a smaller share (from four ranks up) is a chain of short kernels whose time
            # follows the number of chunks -- i.e. of noise sheets -- much more: 39 ns per chunk against 0.2 ns per vertex measured over the 8 ranks' own times
            empty = 37.0 if self.n_total / D.size >= 8 * 148 else 160.0
            cost = np.full(self.n_total, empty)
            for idx, nv in D.allgather((mine0, nv0)):
                cost[idx] += nv
            parts = world.partition(mortons, cost, D.size, columns)
            self.cost_model = "%d + n_verts of the previous rebuild" % empty
        self.parts = [np.sort(p) for p in parts]  # batch order inside a part
        self.mine = self.parts[D.rank]
        self.descs = np.ascontiguousarray(descs_all[self.mine])
        # size the regions from one real run of this rank's share (+25 %)
        if len(self.descs):
            c.submit(self.descs, dim, iters=iters)
            _, V, I = c.totals()
        else:
            V = I = 0
        self.V, self.I = V, I
        compact = mode == "compact"
        mine_layout = (int(V * 1.25) + 1024, int(I * 1.25) + 4096, len(self.descs), 2 if compact else 4, not compact, False)
        lay = [gather.RegionLayout(*t) for t in D.allgather(mine_layout)]
        port = os.environ.get("MASTER_PORT", "0")
        self.g = gather.HostGather("bmf_bench_%s_%s_%d" % (port, tag, os.getppid() if D.size > 1 else os.getpid()), D.rank, D.size, lay, create=(D.rank == 0))
        D.barrier()
        self.g.register(c.lib)
        self.compact = compact
        self.vox = self.n_total * dim ** 3

    def enqueue(self, c, slot):
        """copy-engine download of c's resident batch into this rank's region of `slot` (the host waits for the batch's kernels first)"""
        b = self.g.buffers(slot)
        if self.compact:
            c.download_dma(pos=b["pos"], inds16=b["inds"])
        else:
            c.download_dma(pos=b["pos"], color=b["color"], inds32=b["inds"])

    def run(self, steps, collect=True):
        """the e2e loop of every rank: the NC contexts take the batches round robin, rank 0 gathers.  Returns rank 0's last (table, owner).
        Iteration k: the kernels of batch k are queued on context k % NC; THEN the host waits for the kernels of batch k - LAG (another context) and
        queues its DMA -- which runs beside the kernels of the batches after it; THEN the DMA of batch k - LAG - 1 is awaited, its chunk table
        published and (rank 0) gathered.  LAG = NC - 2: a context is free again before its turn comes round, and kernels are queued LAG steps ahead
        of their download (command fetch is slow while a download saturates the PCIe link: reads do not pass posted writes)."""
        g, last = self.g, None
        have = len(self.descs) > 0
        nc = len(self.ctxs)
        lag = max(1, nc - 2)
        for k in range(steps + lag + 1):
            if nc < 3 and k >= lag + 1:
                last = self._finish(k - lag - 1, self.ctxs[(k - lag - 1) % nc], collect)
            if k < steps and have:
                self.ctxs[k % nc].submit(self.descs, self.dim, iters=self.iters)
            if lag <= k < steps + lag and have:
                g.wait_slot_free(k - lag)
                self.enqueue(self.ctxs[(k - lag) % nc], (k - lag) % g.SLOTS)
            if nc >= 3 and k >= lag + 1:
                last = self._finish(k - lag - 1, self.ctxs[(k - lag - 1) % nc], collect)
        return last

    def _finish(self, step, c, collect):
        g = self.g
        if len(self.descs):
            c.wait()
            g.publish(step, c.chunk_infos())
        else:
            g.publish(step, np.zeros(0, g.buffers(0)["table"].dtype))
        if self.D.rank == 0 and collect:
            out = g.collect(step, self.parts, self.n_total)
            g.release(step)
            return out
        return None

    def dma_only(self, K):
        """the ceiling of the e2e loop on this box: the resident batches of two contexts downloaded again and again (no kernels besides the uint16
        pack), all ranks at once.  Returns seconds per download (max over ranks)."""
        D, cs = self.D, self.ctxs[:2]
        have = len(self.descs) > 0
        if have:
            for c in cs:
                c.submit(self.descs, self.dim, iters=self.iters)
                c.wait()
        D.barrier()
        t0 = time.perf_counter()
        if have:
            for k in range(K):
                self.enqueue(cs[k & 1], k % self.g.SLOTS)
                cs[(k + 1) & 1].wait()
            cs[(K - 1) & 1].wait()
        D.barrier()
        return D.max(time.perf_counter() - t0) / K

    def bytes_per_step(self):
        """D2H bytes of one step summed over ranks: what the GPUs really store (needs the per-rank totals)"""
        per = 12 * self.V + (2 if self.compact else 4) * self.I + (0 if self.compact else 12 * self.V) + 56 * len(self.descs) + 128
        return int(self.D.sum(float(per)))

    def gathered_crcs(self, steps_done, table, owner):
        """rank 0: CRC-32 of the gathered batch in batch order (indices widened to uint32), like tests/test_gpu_fullsize.py"""
        slot = (steps_done - 1) % self.g.SLOTS
        ci = cp = 0
        for i in np.flatnonzero(table["n_verts"] > 0):
            p, idx = self.g.chunk_arrays(slot, table, owner, int(i))
            cp = zlib.crc32(np.ascontiguousarray(p).tobytes(), cp)
            ci = zlib.crc32(idx.astype(np.uint32).tobytes(), ci)
        return int(table["n_verts"].sum()), int(table["n_inds"].sum()), ci & 0xFFFFFFFF, cp & 0xFFFFFFFF

    def close(self):
        self.g.close()


def time_e2e(D, sw, K):
    for c in sw.ctxs:
        c.set_batches_in_flight(len(sw.ctxs))  # the contexts overlap their batches: least SM time first (bmf_ctx_set_batches_in_flight)
    try:
        return _time_e2e(D, sw, K)
    finally:
        for c in sw.ctxs:
            c.set_batches_in_flight(1)


def _time_e2e(D, sw, K):
    sw.run(2 * len(sw.ctxs))  # untimed: every context meshes this world twice (the first submit of a new world grows its arenas and re-launches the emitters)
    D.barrier()
    sw.g.reset()
    D.barrier()
    t0 = time.perf_counter()
    last = sw.run(K)
    D.barrier()
    s = D.max(time.perf_counter() - t0)
    D.barrier()
    sw.g.reset()
    D.barrier()
    return s, last


def time_device(D, ctx, stream, submit, K, W):
    torch = D.torch
    for _ in range(W):
        submit()
    ctx.wait()
    D.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(stream)
    for _ in range(K):
        submit()
    e1.record(stream)
    ctx.wait()
    D.barrier()
    wall = time.perf_counter() - t0
    return D.max(e0.elapsed_time(e1) * 1e-3), wall


def time_device_pipelined(D, ctxs, streams, submit_on, K, W):
    """K steps with len(ctxs) batches in flight: step k goes to context k % NC (its own stream and arenas), so the tail of one batch's launches
    overlaps the head of the next one's.  Device time = from an event recorded before the first submit to the LAST of the events recorded
    on every stream after the final submit; max over ranks."""
    torch = D.torch
    nc = len(ctxs)
    for k in range(max(W, 2 * nc)):  # every context twice: the first submit of a kind grows its arenas and re-launches
        submit_on(ctxs[k % nc])
    for c in ctxs:
        c.wait()
    torch.cuda.synchronize()
    D.barrier()
    e0 = torch.cuda.Event(enable_timing=True)
    ends = [torch.cuda.Event(enable_timing=True) for _ in streams]
    l0 = sum(c.launch_count() for c in ctxs)
    t0 = time.perf_counter()
    e0.record(streams[0])
    for k in range(K):
        submit_on(ctxs[k % nc])
    for e, s in zip(ends, streams):
        e.record(s)
    for c in ctxs:
        c.wait()
    torch.cuda.synchronize()
    D.barrier()
    wall = time.perf_counter() - t0
    mine = max(e0.elapsed_time(e) for e in ends) * 1e-3
    time_device_pipelined.per_rank_ms = [round(t / K * 1e3, 4) for t in D.allgather(mine)]  # every rank's own device time per step (the value uses the slowest)
    return D.max(mine), wall, sum(c.launch_count() for c in ctxs) - l0


def run_ours(args):
    from binarymeshfitting_b200 import Context, capi, world
    D = Dist()
    torch = D.torch
    if D.size > 1:
        args.gpus = D.size
    numa = pin_to_gpu_numa_node(D.local_rank)
    n_ctx = max(3, args.in_flight) if args.in_flight else (6 if D.size >= 4 else 4)
    ctxs_all = [Context(D.local_rank) for _ in range(n_ctx)]  # raises if the CUDA library or the device is missing: no fallback
    ctxs = ctxs_all[:4]  # the end-to-end pipeline is download-bound: four contexts keep the copy engine busy, more only lengthen its ramp
    ctx = ctxs[0]
    stream = torch.cuda.ExternalStream(ctx.stream_ptr(), device=torch.device("cuda", D.local_rank))
    kind = SAMPLERS[args.sampler]
    for c in ctxs_all:
        c.set_sampler(kind)
    ps, overlap = workload(args)
    descs_all = capi.make_chunk_descs(ps, overlaps=overlap)
    mortons = world.grid_mortons(args.chunks_per_axis)
    descs_all["morton"] = mortons
    dim, K, W = args.dim, args.steps, max(args.warmup, 3)
    columns = args.sampler.startswith("terrain2d")
    sw = SharedWorld(D, ctxs, descs_all, mortons, dim, args.iters, "main", mode="compact", columns=columns)
    nvox = sw.vox

    def step():
        if len(sw.descs):
            ctx.submit(sw.descs, dim, iters=args.iters)

    # ---- device-timed throughput (`value`): this rank's share of the world, inputs (descriptors) resident
    launches0 = ctx.launch_count()
    clocks = ClockSampler(D.local_rank)
    clocks.start()
    t_one, wall_one = time_device(D, ctx, stream, step, K, W)
    stage = ctx.stage_ms()
    # the same K steps with several batches in flight (one per context / stream): what a job that streams batches through the GPU gets
    streams = [torch.cuda.ExternalStream(c.stream_ptr(), device=torch.device("cuda", D.local_rank)) for c in ctxs_all]

    def submit_on(c):
        if len(sw.descs):
            c.submit(sw.descs, dim, iters=args.iters)

    for c in ctxs_all:
        c.set_batches_in_flight(len(ctxs_all))
    t_max, wall, launches = time_device_pipelined(D, ctxs_all, streams, submit_on, K, W)  # launches: this rank's kernels inside the timed region
    for c in ctxs_all:
        c.set_batches_in_flight(1)
    value = nvox * K / t_max
    infos_mine = ctx.chunk_infos() if len(sw.descs) else np.zeros(0, capi.CHUNK_INFO_DTYPE)
    n_mesh_mine = int((infos_mine["contains_mesh"] != 0).sum())
    V, I = sw.V, sw.I

    # ---- end to end (headline): compact download, device-driven, gathered on rank 0
    for c in ctxs:
        c.set_reserved_sms(args.reserve_sms)
    e2e_s, last = time_e2e(D, sw, K)
    d2h = sw.bytes_per_step()
    h2d = int(descs_all.nbytes + 16 * len(descs_all))  # descriptors (host ABI) + the geometry records the library uploads
    verify = None
    if D.rank == 0:
        tv, ti, ci, cp = sw.gathered_crcs(K, *last)
        verify = {"verts": tv, "indices": ti, "inds_crc": ci, "pos_crc": cp}
        gold_p = os.path.join(ROOT, "tests", "golden", "golden.json")
        if os.path.exists(gold_p):
            g = json.load(open(gold_p))["bench_workload"]
            if (g["chunks"], g["dim"], g["sampler"], g["iters"]) == (len(descs_all), dim, args.sampler, args.iters):
                verify["matches_compiled_reference_golden"] = bool((tv, ti, ci, cp) == (g["verts"], g["inds"], g["inds_crc"], g["pos_crc"]))
    clk = clocks.stop()  # sampled across both timed regions
    e2e = {"value": nvox * K / e2e_s, "unit": "voxels/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_s / K * 1e3,
           "d2h_GB_per_s": d2h / (e2e_s / K) / 1e9,
           "mode": "opt-in compact download (positions + uint16 chunk-local indices packed on the device; colour == 1 and the chunk table say what was skipped), "
                   "copy engine straight into a shared pinned host segment (bmf_batch_download_dma), %d contexts round robin so that the DMA of batch i runs beside "
                   "the kernels of the batches after it, rank 0 assembles the batch-order chunk table" % len(ctxs),
           "gathered": verify}
    dma_s = sw.dma_only(K)
    e2e["d2h_ceiling"] = {"ms_per_step": dma_s * 1e3, "GB_per_s": d2h / dma_s / 1e9, "voxels_per_s": nvox / dma_s,
                          "what": "the same downloads without the kernels, all ranks at once: what the link(s) and the host memory of this box allow end to end"}
    sw.close()
    # the reference-layout download (positions + colours + uint32 indices, GLChunk::format_data's arrays) through the same path
    sw2 = SharedWorld(D, ctxs, descs_all, mortons, dim, args.iters, "ref", mode="reference_layout", columns=columns)
    s2, _ = time_e2e(D, sw2, K)
    e2e["reference_layout"] = {"value": nvox * K / s2, "ms_per_step": s2 / K * 1e3, "d2h_bytes_per_step": sw2.bytes_per_step(),
                               "d2h_GB_per_s": sw2.bytes_per_step() / (s2 / K) / 1e9, "streams": "positions + colours + uint32 indices"}
    sw2.close()
    for c in ctxs:
        c.set_reserved_sms(0)

    result = {
        "metric": "voxels/sec sampled+meshed", "value": value, "unit": "voxels/s", "n_gpus": args.gpus, "steps": K, "warmup": W,
        "ms_per_step": t_max / K * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32+u32",
        "data": "synthetic (procedural noise terrain, seed 1337; no dataset)",
        "config": config_of(args, overlap),
        "chunks_per_s": value / dim ** 3,
        "wall_ms_per_step": wall / K * 1e3,
        "in_flight": {"batches": len(ctxs_all), "ms_per_step_per_rank": getattr(time_device_pipelined, "per_rank_ms", None), "note": "step k is submitted to context k % NC (own stream, own arenas); every step is a complete batch, consecutive steps overlap on the GPU",
                      "single_stream": {"value": nvox * K / t_one, "ms_per_step": t_one / K * 1e3, "wall_ms_per_step": wall_one / K * 1e3,
                                        "note": "the same K steps back to back on ONE stream: the latency of a step; kernels[], roofline and stage_ms are measured in this mode"}},
        "e2e": e2e,
        "gpu_launches": int(D.sum(float(launches))),
        "clocks": clk,
        "partition": {"scheme": "world.partition: contiguous ranges of the depth-normalised Morton order%s, balanced by cost (%s); no data-path collective" % (
                          " regrouped column-major ((z, x) bits before the y bits: a rank owns whole chunk columns and no noise sheet is sampled twice)" if sw.columns else "", sw.cost_model or "one rank"),
                      "chunks_per_rank": [int(len(p)) for p in sw.parts], "mesh_chunks_per_rank": [int(x) for x in D.allgather(n_mesh_mine)],
                      "numa": numa},
        "l2": "working set per step (sign words of the mesh chunks + per-word records + meshes, > 200 MB at N=1) exceeds the 126 MB L2; no explicit flush",
    }

    # ---- per-kernel times + roofline (rank 0's share)
    peaks = load_peaks()
    prof = load_kernel_profile(workload_name(args)) if D.size == 1 else None
    sms = torch.cuda.get_device_properties(D.local_rank).multi_processor_count
    st = {"dim": dim, "n_chunks": len(sw.descs), "n_mesh": n_mesh_mine, "n_sampled": n_mesh_mine, "V": V, "I": I, "iters": args.iters}
    if len(sw.descs):
        per_kernel, lines = kernel_report(ctx, step, st, peaks, prof, clk, sms)
        dominant = max(per_kernel, key=per_kernel.get)
        result["roofline"] = roofline_of(dominant, lines, peaks)
        result["kernels"] = lines
    # whole step against SURVEY 8(d)'s fused sample->mesh figure, over the chunks that are actually meshed
    n_mesh_all = int(D.sum(float(n_mesh_mine)))
    Vt, It = D.sum(float(V)), D.sum(float(I))
    step_bytes = n_mesh_all * dim ** 3 * (1.0 / 8 + 1.0) + 14.0 * Vt + 4.0 * It
    result["step_roofline"] = {"algorithmic_bytes": int(step_bytes), "achieved": round(step_bytes / (t_max / K) / 1e9, 1), "peak": peaks["hbm_gbs"] * args.gpus, "unit": "GB/s",
                               "frac": round(step_bytes / (t_max / K) / 1e9 / (peaks["hbm_gbs"] * args.gpus), 4),
                               "note": "SURVEY 8(d) fused figure N/8 + N + 14V + 4I over the %d mesh-containing chunks only (the other %d are culled by a one-thread-per-chunk "
                                       "classifier); no kernel of this step is DRAM-bound -- the integer kernels are issue-bound, see kernels[]" % (n_mesh_all, len(descs_all) - n_mesh_all)}
    result["surface_chunks"] = {"chunks_with_mesh": n_mesh_all, "voxels_per_s_over_mesh_chunks": n_mesh_all * dim ** 3 * K / t_max,
                                "us_per_meshed_chunk": t_max / K * 1e6 / max(n_mesh_all, 1) * args.gpus,
                                "note": "the headline counts all %d chunks like the CPU arm (which samples every voxel); %d %% of them are trivially empty" %
                                        (len(descs_all), round(100 - 100.0 * n_mesh_all / len(descs_all)))}
    result["stage_ms"] = {k: round(v, 4) for k, v in stage.items()}
    result["mesh"] = {"verts": int(Vt), "indices": int(It)}

    ex = {}
    if not args.no_extras:
        if D.size == 1:
            ex = extras_single(ctx, ctxs, args, capi, world, peaks, prof, clk, sms, D)
        else:
            ex = extras_multi(D, ctxs, args, capi, world, stream)
    if D.rank == 0 and ex:
        result["extras"] = ex
    if D.rank == 0 and D.size == 1 and not args.no_cpu_baseline:
        result["cpu_baseline"] = cpu_baseline(args, ps, overlap)
    for c in ctxs_all:
        c.close()
    D.close()
    if D.rank == 0:
        print(json.dumps(result))


def timed_wall(ctx, descs, dim, iters, reps=5, **kw):
    for _ in range(2):
        ctx.submit(descs, dim, iters=iters, **kw)
    ctx.wait()
    t0 = time.perf_counter()
    for _ in range(reps):
        ctx.submit(descs, dim, iters=iters, **kw)
    ctx.wait()
    return (time.perf_counter() - t0) / reps


def surface_dense_batch(ctx, capi, world, args, overlap):
    """a batch in which EVERY chunk crosses the surface: the mesh-containing chunks of the headline world, tiled to the headline's chunk count"""
    ps, _ = workload(args)
    d = capi.make_chunk_descs(ps, overlaps=overlap)
    ctx.submit(d, args.dim, iters=0)
    inf = ctx.chunk_infos()
    mesh = np.flatnonzero(inf["n_verts"] > 0)
    reps = int(np.ceil(len(d) / max(len(mesh), 1)))
    return np.ascontiguousarray(np.tile(d[mesh], reps)[:len(d)])


def extras_single(ctx, ctxs, args, capi, world, peaks, prof, clk, sms, D):
    """Secondary numbers the metric names, one GPU."""
    ex = {}
    ps, overlap = workload(args)
    d = capi.make_chunk_descs(ps, overlaps=overlap)
    dim = args.dim
    stream = D.torch.cuda.ExternalStream(ctx.stream_ptr(), device=D.torch.device("cuda", D.local_rank))

    # ---- surface-dense companion of the headline: every chunk crosses the surface
    try:
        ctx.set_sampler(SAMPLERS[args.sampler])
        sd = surface_dense_batch(ctx, capi, world, args, overlap)
        t, _ = time_device(D, ctx, stream, lambda: ctx.submit(sd, dim, iters=args.iters), 10, 3)
        _, V2, I2 = ctx.totals()
        st = {"dim": dim, "n_chunks": len(sd), "n_mesh": len(sd), "n_sampled": len(sd), "V": V2, "I": I2, "iters": args.iters}
        per_kernel, lines = kernel_report(ctx, lambda: ctx.submit(sd, dim, iters=args.iters), st, peaks, None, clk, sms, reps=3)
        sb = len(sd) * dim ** 3 * (1.0 / 8 + 1.0) + 14.0 * V2 + 4.0 * I2
        ex["surface_dense_%dx%d_%s" % (len(sd), dim, args.sampler)] = {
            "what": "the %d mesh-containing chunks of the headline world tiled to %d chunks: every chunk crosses the surface" % (len(np.unique(sd["pos"], axis=0)), len(sd)),
            "ms_per_step": t / 10 * 1e3, "voxels_per_s": len(sd) * dim ** 3 * 10 / t, "mesh": {"verts": int(V2), "indices": int(I2)},
            "step_roofline": {"algorithmic_bytes": int(sb), "GB/s": sb / (t / 10) / 1e9, "frac": sb / (t / 10) / 1e9 / peaks["hbm_gbs"], "figure": "SURVEY 8(d) fused N/8 + N + 14V + 4I"},
            "kernels": lines}
    except Exception as e:  # noqa: BLE001 -- an extra must never take the headline down
        ex["surface_dense"] = {"error": str(e)}

    # ---- 3-D fractal noise: the bounding stage of "FastNoiseSIMD fractal terrain" (issue-bound; its own roofline)
    try:
        ctx.set_sampler(capi.TERRAIN3D_PERT)
        t, _ = time_device(D, ctx, stream, lambda: ctx.submit(d, dim, iters=args.iters), 3, 3)
        e3 = {"ms_per_step": t / 3 * 1e3, "voxels_per_s": len(d) * dim ** 3 * 3 / t, "stage_ms": ctx.stage_ms()}
        ctx.set_kernel_timing(True)
        ctx.submit(d, dim, iters=args.iters)
        kt = dict((n, ms) for n, ms in ctx.kernel_times())
        ctx.set_kernel_timing(False)
        noise_ms = max(kt.values())
        noise_name = max(kt, key=kt.get)
        pk = ctx.ubench_issue()
        sass_per_step = {"fp32_fma": 1, "int32_imad": 1, "int32_logic_step": 3, "fma_plus_logic_step": 2}  # SASS instructions per chain step (cuobjdump)
        peaks_inst = {k: v * sass_per_step[k] for k, v in pk.items()}  # 1e9 thread-level instructions/s
        prof3 = None
        p3 = os.path.join(ROOT, "profiles", "r2_noise3d_profile.json")
        if os.path.exists(p3):
            prof3 = json.load(open(p3))
        inst_per_voxel = (prof3 or {}).get("thread_inst_per_voxel", 2168.0)
        ach = len(d) * dim ** 3 / (noise_ms * 1e-3) * inst_per_voxel / 1e9
        e3["roofline"] = {"kernel": noise_name, "bound": "issue (FP32 + INT32 pipes)", "launch_ms": noise_ms, "share_of_step": noise_ms / sum(kt.values()),
                          "thread_inst_per_voxel": inst_per_voxel, "achieved": ach, "unit": "1e9 thread-level instructions/s",
                          "peak": peaks_inst["fp32_fma"], "frac": ach / peaks_inst["fp32_fma"],
                          "peak_source": "of measured (bmf_ubench_issue: dependency-free FFMA chains, 2048 threads/SM, this run)",
                          "measured_issue_peaks": peaks_inst, "nominal_fp32_fma": 148 * 128 * 1.965,
                          "ncu": prof3, "note": "every instruction of the kernel against the measured full-rate (FFMA) issue peak; INT32 logic/IMAD issue at about half that rate on "
                                                "B200, so the integer share (hashing, lattice selects) is what bounds it"}
        ex["terrain3d_pert_4096x64"] = e3
    except Exception as e:  # noqa: BLE001
        ex["terrain3d_pert_4096x64"] = {"error": str(e)}

    # ---- config 4: LOD world, 2048^3 effective voxels at the finest level (dim 64, max_level 5) = 232 leaves
    props = world.WorldProperties(max_level=5, chunk_resolution=64, process_iters=2)
    lps, lv, mc = world.split_leaves(props)
    ld = world.make_descs(props, lps, lv, mc)
    for name, kind in (("terrain2d_pert", capi.TERRAIN2D_PERT), ("terrain3d_pert", capi.TERRAIN3D_PERT)):
        ctx.set_sampler(kind)
        s = timed_wall(ctx, ld, 64, 2, reps=10)
        t, _ = time_device(D, ctx, stream, lambda: ctx.submit(ld, 64, iters=2), 10, 3)
        inf = ctx.chunk_infos()
        _, Vl, Il = ctx.totals()
        nm = int((inf["contains_mesh"] != 0).sum())
        sb = nm * 64 ** 3 * (1.0 / 8 + 1.0) + 14.0 * Vl + 4.0 * Il
        ex["lod_rebuild_232x64_%s" % name] = {"ms": s * 1e3, "device_ms": t / 10 * 1e3, "voxels_per_s": len(ld) * 64 ** 3 / s, "chunks_with_mesh": nm,
                                              "launches_per_rebuild": None, "frac_of_hbm_roofline": sb / (t / 10) / 1e9 / peaks["hbm_gbs"]}
        l0 = ctx.launch_count()
        ctx.submit(ld, 64, iters=2)
        ctx.wait()
        ex["lod_rebuild_232x64_%s" % name]["launches_per_rebuild"] = ctx.launch_count() - l0
        if kind == capi.TERRAIN2D_PERT and len(ctxs) >= 4:
            # the same rebuild when the host has several of them in flight (several worlds / frames ahead): four contexts round robin, per-chunk kernels
            for c in ctxs[:4]:
                c.set_sampler(kind)
                c.set_batches_in_flight(4)
            for k in range(8):
                ctxs[k % 4].submit(ld, 64, iters=2)
            best = 1e9
            for rep in range(3):
                for c in ctxs[:4]:
                    c.wait()
                t0 = time.perf_counter()
                for k in range(40):
                    ctxs[k % 4].submit(ld, 64, iters=2)
                for c in ctxs[:4]:
                    c.wait()
                best = min(best, (time.perf_counter() - t0) / 40)
            for c in ctxs[:4]:
                c.set_batches_in_flight(1)
            ex["lod_rebuild_232x64_%s" % name]["ms_per_rebuild_four_in_flight"] = best * 1e3
    # config 4 as named ("multi-level chunks + WorldStitcher seams")
    try:
        sdsc = capi.make_chunk_descs(lps, overlaps=ctx.seam_overlap(64), levels=lv)
        for name, kind in (("terrain2d_pert", capi.TERRAIN2D_PERT), ("terrain3d_pert", capi.TERRAIN3D_PERT)):
            ctx.set_sampler(kind)
            walls, nt = [], 0
            for r in range(6):
                t0 = time.perf_counter()
                ctx.submit(sdsc, 64, iters=2)
                nt = ctx.stitch(download=False)
                walls.append(time.perf_counter() - t0)
            ex["lod_rebuild_with_seams_232x64_%s" % name] = {"ms": min(walls[1:]) * 1e3, "seam_tris": nt, "seam_device_ms": ctx.seam_ms(),
                                                            "mesh": dict(zip(("cells", "verts", "indices"), ctx.totals()))}
        ctx.set_sampler(SAMPLERS[args.sampler])
        gd = capi.make_chunk_descs(ps, overlaps=ctx.seam_overlap(dim))
        ctx.submit(gd, dim, iters=args.iters)
        nt = ctx.stitch(download=False)
        nt = ctx.stitch(download=False)
        sm = ctx.seam_ms()
        pts = len(gd) * (6 * dim ** 2 + 2)
        ex["seam_pass_4096x64_%s" % args.sampler] = {"seam_tris": nt, "device_ms": sm, "lattice_points": pts,
                                                     "lattice_points_per_s": pts / ((sm["count"] + sm["emit"]) * 1e-3)}
    except Exception as e:  # noqa: BLE001
        ex["lod_rebuild_with_seams"] = {"error": str(e)}
    # "ms per LOD rebuild", incremental reading: the WorldWatcher tick while the focus flies 4 units per tick along x
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
            t0 = time.perf_counter()
            ctx.submit(gd, 64, iters=2)
            ctx.wait()
            tick_ms.append((time.perf_counter() - t0) * 1e3)
            sizes.append(len(gd))
        if tick_ms:
            ex["lod_incremental_fly_%s" % args.sampler] = {"ticks_with_work": len(tick_ms), "chunks_per_tick_mean": sum(sizes) / len(sizes), "chunks_per_tick_max": max(sizes),
                                                           "ms_per_tick_median": sorted(tick_ms)[len(tick_ms) // 2], "ms_per_tick_max": max(tick_ms),
                                                           "note": "submit + wait per tick (host wall clock); the tick policy itself is host code and not timed"}
    except Exception as e:  # noqa: BLE001
        ex["lod_incremental_fly"] = {"error": str(e)}
    # config 1: single 64^3 chunk of the implicit sphere, no processing -- triangles and quads
    try:
        ctx.set_sampler(capi.SPHERE)
        c1 = capi.make_chunk_descs([[-128, -128, -128, 256.0]])
        for name, qd in (("tris", False), ("quads", True)):
            s = timed_wall(ctx, c1, 64, 0, reps=20, quads=qd)
            ex["single_64_sphere_%s" % name] = {"ms": s * 1e3, "mesh": dict(zip(("cells", "verts", "indices"), ctx.totals()))}
        ctx.set_sampler(SAMPLERS[args.sampler])
        s = timed_wall(ctx, d, dim, 0, reps=3, quads=True)
        ex["quads_4096x64_%s" % args.sampler] = {"ms": s * 1e3, "voxels_per_s": len(d) * dim ** 3 / s, "stage_ms": ctx.stage_ms(),
                                                 "mesh": dict(zip(("cells", "verts", "indices"), ctx.totals()))}
    except Exception as e:  # noqa: BLE001
        ex["single_64_sphere_quads"] = {"error": str(e)}
    # config 2: single 128^3 chunk, 2 smoothing iterations
    one = capi.make_chunk_descs([[-64, -64, -64, 128.0]], overlaps=0.045)
    for name, kind in (("terrain2d_pert", capi.TERRAIN2D_PERT), ("terrain3d_pert", capi.TERRAIN3D_PERT)):
        ctx.set_sampler(kind)
        s = timed_wall(ctx, one, 128, 2, reps=20)
        ex["single_128_%s" % name] = {"ms": s * 1e3, "voxels_per_s": 128 ** 3 / s}
    # config 5: 128^3 CSG with gradients: QEF placement from implicit_gradient normals + smoothed normals, triangles
    try:
        ctx.set_sampler(capi.CSG, csg_op=capi.CSG_SUBTRACT, csg_kind_a=capi.SPHERE, csg_kind_b=capi.CUBOID, csg_world_size_a=256.0, csg_world_size_b=300.0,
                        csg_offset_a=(0.0, 0.0, 0.0), csg_offset_b=(20.0, -10.0, 5.0))
        c5 = capi.make_chunk_descs([[-128, -128, -128, 256.0]], overlaps=0.055)
        for q in (1, 2):
            try:
                s = timed_wall(ctx, c5, 128, 4, reps=10, smooth_normals=True, qef=q)
                ex["csg_128_qef%d" % q] = {"ms": s * 1e3, "voxels_per_s": 128 ** 3 / s, "mesh": dict(zip(("cells", "verts", "indices"), ctx.totals()))}
            except Exception as e:  # noqa: BLE001
                ex["csg_128_qef%d" % q] = {"error": str(e)}
    except Exception as e:  # noqa: BLE001
        ex["csg_128"] = {"error": str(e)}
    # config 4, dense reading: the whole 2048^3-voxel world at the finest LOD = 32x32x32 chunks of 64^3 (8.6 Gvoxel)
    try:
        ctx.set_sampler(SAMPLERS[args.sampler])
        dd = capi.make_chunk_descs(world.grid_chunks(32, 16.0, origin=(-256.0, -256.0, -256.0)), overlaps=overlap)
        t, _ = time_device(D, ctx, stream, lambda: ctx.submit(dd, 64, iters=args.iters), 3, 3)
        inf = ctx.chunk_infos()
        nm = int((inf["contains_mesh"] != 0).sum())
        _, Vd, Id = ctx.totals()
        sb = nm * 64 ** 3 * (1.0 / 8 + 1.0) + 14.0 * Vd + 4.0 * Id
        ex["dense_2048_cubed_%s" % args.sampler] = {"chunks": len(dd), "chunks_with_mesh": nm, "ms": t / 3 * 1e3, "voxels_per_s": len(dd) * 64 ** 3 * 3 / t,
                                                    "voxels_per_s_over_mesh_chunks": nm * 64 ** 3 * 3 / t, "frac_of_hbm_roofline": sb / (t / 3) / 1e9 / peaks["hbm_gbs"],
                                                    "mesh": {"verts": int(Vd), "indices": int(Id)}}
    except Exception as e:  # noqa: BLE001
        ex["dense_2048_cubed_%s" % args.sampler] = {"error": str(e)}
    # K2 on its own: the density block of the benchmark workload (4.3 GB, resident in HBM) -> sign words: the path's one pure streaming kernel
    try:
        ctx.set_sampler(capi.TERRAIN2D_PERT)
        ctx.submit(d, dim, iters=0, keep_density=True)
        ctx.wait()
        dptr = ctx.device_ptrs()["density"]
        ctx.set_sampler(capi.HOST_DENSITY)
        ctx.set_kernel_timing(True)
        best = None
        for _ in range(5):
            ctx.submit(d, dim, iters=0, density_device_ptr=dptr)
            t = [ms for name, ms in ctx.kernel_times() if name == "k_pack_density"]
            best = t[0] if best is None else min(best, t[0])
        ctx.set_kernel_timing(False)
        nb = len(d) * dim ** 3 * (4 + 1.0 / 8)
        ex["k2_pack_density_4096x64"] = {"ms": best, "algorithmic_bytes": int(nb), "GB/s": nb / best / 1e6, "peak": peaks["hbm_gbs"],
                                         "frac": nb / best / 1e6 / peaks["hbm_gbs"], "peak_source": peaks["source"], "bound": "hbm"}
        s = timed_wall(ctx, d, dim, args.iters, density_device_ptr=dptr)
        _, V2, I2 = ctx.totals()
        nvox2 = len(d) * dim ** 3
        by = 5.125 * nvox2 + 14.0 * V2 + 4.0 * I2
        ex["density_in_mesh_out_4096x64"] = {"ms": s * 1e3, "voxels_per_s": nvox2 / s, "algorithmic_bytes": int(by), "GB/s": by / s / 1e9,
                                             "frac": by / s / 1e9 / peaks["hbm_gbs"], "hbm_roofline_voxels_per_s": nvox2 / (by / (peaks["hbm_gbs"] * 1e9)),
                                             "note": "the one configuration of this path that really streams HBM: every chunk's 1 MiB density block is read (4N bytes), whatever it contains"}
    except Exception as e:  # noqa: BLE001
        ex["k2_pack_density_4096x64"] = {"error": str(e)}
    # end-to-end variants at N=1: round 1's path (synchronising copy-engine download of positions + colours + indices, one context)
    try:
        ctx.set_sampler(SAMPLERS[args.sampler])
        ctx.submit(d, dim, iters=args.iters)
        _, V, I = ctx.totals()
        pb = {k: capi.PinnedBuffer(ctx.lib, nb) for k, nb in (("pos", 12 * V), ("color", 12 * V), ("inds", 4 * I))}
        out = {"pos": pb["pos"].view(np.float32).reshape(-1, 3), "color": pb["color"].view(np.float32).reshape(-1, 3), "inds": pb["inds"].view(np.uint32)}
        reps = 5
        for _ in range(2):
            ctx.submit(d, dim, iters=args.iters)
            ctx.download(want=("pos", "color", "inds"), out=out)
        t0 = time.perf_counter()
        for _ in range(reps):
            ctx.submit(d, dim, iters=args.iters)
            ctx.download(want=("pos", "color", "inds"), out=out)
            ctx.chunk_infos()
        s = (time.perf_counter() - t0) / reps
        ex["e2e_copy_engine_serial_reference_layout"] = {"ms_per_step": s * 1e3, "voxels_per_s": len(d) * dim ** 3 / s, "d2h_GB_per_s": (24 * V + 4 * I) / s / 1e9,
                                                         "note": "bmf_batch_download (waits for the batch, then cudaMemcpyAsync x3, then waits): round 1's serial path"}
        for b in pb.values():
            b.close()
    except Exception as e:  # noqa: BLE001
        ex["e2e_copy_engine_serial_reference_layout"] = {"error": str(e)}
    ctx.set_sampler(SAMPLERS[args.sampler])
    return ex


def extras_multi(D, ctxs, args, capi, world, stream):
    """N > 1: the other worlds the north star names, partitioned over the ranks, + round 1's replica (weak) run"""
    ex = {}
    ctx = ctxs[0]
    dim = args.dim
    kind = SAMPLERS[args.sampler]
    _, overlap = workload(args)
    # ---- weak: every rank its own 4096-chunk region (what round 1's SCALE line measured)
    try:
        ps_r, _ = workload(args, region=D.rank)
        dr = capi.make_chunk_descs(ps_r, overlaps=overlap)
        t, _ = time_device(D, ctx, stream, lambda: ctx.submit(dr, dim, iters=args.iters), 10, 3)
        ex["weak"] = {"what": "every rank meshes its own %d-chunk region (replicas, no partition)" % len(dr), "ms_per_step": t / 10 * 1e3,
                      "voxels_per_s": D.size * len(dr) * dim ** 3 * 10 / t}
    except Exception as e:  # noqa: BLE001
        ex["weak"] = {"error": str(e)}
    # ---- config 4, dense reading: ONE world of 32768 chunks (8.6 Gvoxel), partitioned
    try:
        n = 32
        dps = world.grid_chunks(n, 16.0, origin=(-256.0, -256.0, -256.0))
        dd = capi.make_chunk_descs(dps, overlaps=overlap)
        mc = world.grid_mortons(n)
        dd["morton"] = mc
        sw = SharedWorld(D, ctxs, dd, mc, dim, args.iters, "dense", columns=args.sampler.startswith("terrain2d"))
        t, _ = time_device(D, ctx, stream, lambda: ctx.submit(sw.descs, dim, iters=args.iters), 5, 3)
        s, last = time_e2e(D, sw, 5)
        rec = {"chunks": len(dd), "chunks_per_rank": [int(len(p)) for p in sw.parts], "device_ms_per_step": t / 5 * 1e3, "voxels_per_s": sw.vox * 5 / t,
               "e2e_ms_per_step": s / 5 * 1e3, "e2e_voxels_per_s": sw.vox * 5 / s, "d2h_bytes_per_step": sw.bytes_per_step()}
        if D.rank == 0:
            tv, ti, ci, cp = sw.gathered_crcs(5, *last)
            # the same world on this one GPU (untimed): the gathered batch must hash identically
            ctx.submit(dd, dim, iters=args.iters)
            ctx.wait()
            o = ctx.download(want=("pos", "inds"))
            rec["gathered"] = {"verts": tv, "indices": ti, "inds_crc": ci, "pos_crc": cp,
                               "equals_single_gpu_run": bool((ci, cp) == (zlib.crc32(o["inds"].tobytes()) & 0xFFFFFFFF, zlib.crc32(o["pos"].tobytes()) & 0xFFFFFFFF))}
            del o
        D.barrier()
        sw.close()
        ex["dense_2048_cubed_%s" % args.sampler] = rec
    except Exception as e:  # noqa: BLE001
        ex["dense_2048_cubed_%s" % args.sampler] = {"error": repr(e)}
    # ---- config 4 as named: the 232-leaf LOD world, multi-level chunks + seams, cross-rank seam pass on rank 0
    try:
        props = world.WorldProperties(max_level=5, chunk_resolution=64, process_iters=2)
        lps, lv, mc = world.split_leaves(props)
        ov = ctx.seam_overlap(64)
        ld = capi.make_chunk_descs(lps, overlaps=ov, levels=lv)
        ld["morton"] = mc
        parts = [np.sort(p) for p in world.partition(mc, np.ones(len(mc)), D.size)]
        group = np.zeros(len(lps), np.int32)
        for g, p in enumerate(parts):
            group[p] = g
        mine = parts[D.rank]
        border = world.border_chunks(lps, group)
        for c in ctxs:
            c.set_sampler(kind)
        walls, own_tris, cross_tris = [], None, None
        for rep in range(6):
            D.barrier()
            t0 = time.perf_counter()
            if len(mine):
                ctx.submit(np.ascontiguousarray(ld[mine]), 64, iters=2)
                own_tris = ctx.stitch()                       # dual cells inside this rank's chunks
                ctx.download(want=("pos", "inds"))            # this rank's chunk meshes to the host
            else:
                own_tris = np.zeros((0, 3, 3), np.float32)
            if D.rank == 0 and len(border):
                # chunks are pure functions of their descriptors: the gathering rank re-samples the border chunks instead of receiving them
                ctxs[1].submit(np.ascontiguousarray(ld[border]), 64, iters=0)
                cross_tris = ctxs[1].stitch(group=group[border], cross_group_only=True)
            D.barrier()
            walls.append(D.max(time.perf_counter() - t0))
        counts = D.allgather(int(len(own_tris)))
        rec = {"leaves": len(lps), "leaves_per_rank": [int(len(p)) for p in parts], "border_chunks": int(len(border)), "ms_per_rebuild": min(walls[1:]) * 1e3,
               "seam_tris_per_rank": counts, "what": "partition by Z-curve range, per-rank chunk meshes + own seams, host gather, cross-rank seam pass on rank 0 over the border chunks"}
        keys = D.allgather(np.ascontiguousarray(own_tris, np.float32).reshape(-1, 9).view(np.uint32).tobytes())
        if D.rank == 0:
            rec["cross_rank_seam_tris"] = 0 if cross_tris is None else int(len(cross_tris))
            ctx.submit(ld, 64, iters=2)
            full = ctx.stitch()
            got = [np.frombuffer(b, np.uint32).reshape(-1, 9) for b in keys]
            if cross_tris is not None:
                got.append(np.ascontiguousarray(cross_tris, np.float32).reshape(-1, 9).view(np.uint32))
            a = np.concatenate(got) if got else np.zeros((0, 9), np.uint32)
            b = np.ascontiguousarray(full, np.float32).reshape(-1, 9).view(np.uint32)
            same = len(a) == len(b) and np.array_equal(a[np.lexsort(a.T[::-1])], b[np.lexsort(b.T[::-1])])
            rec["seam_equals_single_gpu_seam"] = bool(same)
            rec["seam_tris_single_gpu"] = int(len(b))
        ex["lod_world_232x64_with_cross_rank_seams_%s" % args.sampler] = rec
    except Exception as e:  # noqa: BLE001
        ex["lod_world_with_cross_rank_seams"] = {"error": repr(e)}
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
            "sample": "every %d-th chunk of the workload (%d chunks), 1 warm-up + best of 3, %d process(es) x %d OMP threads; noise = scalar restatement of FastNoiseSIMD "
                      "(a real SIMD FastNoiseSIMD build would be faster on the noise share)" % (stride, len(sample), pool.nproc, pool.threads),
            "ms": best * 1e3, "chunks_per_s": len(sample) / best, "mesh": {"chunks_with_mesh": tot[0], "verts": tot[1], "indices": tot[2]}}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import ref_binding as rb
    if not rb.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libbmf_ref.so missing (built from /root/reference in the authoring container)"}))
        return
    ps, overlap = workload(args)
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
        "ms_per_step": total / K * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32+u32",
        "data": "synthetic (procedural noise terrain, seed 1337; no dataset)",
        "config": config_of(args, overlap),
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

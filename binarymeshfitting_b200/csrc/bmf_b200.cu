// bmf_b200.cu -- the C ABI of include/bmf_b200.h: context, device arenas, stage orchestration.
// Everything that computes is a hand-written sm_100a kernel in noise.cuh / extract.cuh / smooth.cuh;
// this file only owns memory, streams, events and the launch sequence.  There is no CPU fallback:
// a missing device or a failed launch is an error (BMF_ERR_CUDA).
#include "../../include/bmf_b200.h"
#include "extract.cuh"
#include "smooth.cuh"
#include "seam.cuh"
#include "quads.cuh"
#include "download.cuh"
#include "fused.cuh"
#include "post.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <array>
#include <cstring>
#include <map>
#include <string>
#include <vector>

using namespace bmf;

namespace
{

template <typename T>
struct DevBuf
{
	T* p = nullptr;
	size_t cap = 0;
	cudaError_t reserve(size_t n)
	{
		if (n <= cap) return cudaSuccess;
		if (p)
		{
			cudaError_t fe = cudaFree(p);
			p = nullptr;
			cap = 0;
			if (fe != cudaSuccess) return fe;
		}
		size_t want = n + n / 8 + 64;
		cudaError_t e = cudaMalloc((void**)&p, want * sizeof(T));
		if (e == cudaSuccess) cap = want;
		return e;
	}
	void release()
	{
		if (p) cudaFree(p);
		p = nullptr;
		cap = 0;
	}
};

struct HostTotals
{
	unsigned long long v[16]; // the device's totals block (download.cuh TOT_*): cells, verts, indices, >2^32 flag, list counters x2, -, arena-too-small flag, largest chunk, download error
};

} // namespace

struct bmf_ctx
{
	int device = 0;
	cudaStream_t stream = nullptr;
	std::string err;
	int64_t launches = 0;

	bool sampler_set = false;
	bmf_sampler_desc sampler_desc;
	SamplerDev sampler;

	// resident batch
	int n = 0;
	bool have_batch = false;
	bool finished = false;
	bmf_params params;
	Layout L;
	std::vector<bmf_chunk_desc> descs;
	std::vector<ChunkGeom> geom_host;
	std::vector<ChunkCounts> counts_host;
	unsigned long long totals[3] = { 0, 0, 0 };
	const float* ext_density = nullptr; // caller-owned device density (density_on_device)
	const float* density_cur = nullptr; // density block the emitters read crossing-edge samples from (or null)
	int relaunches = 0;                 // batches whose emitters had to be re-launched after growing an arena
	bool density_valid = false, masks_valid = false;
	bool counts_published = false; // the chunk table of the resident batch has been queued for the host (once per batch, after the emitters)
	size_t color_ones = 0; // the first color_ones floats of the colour arena are known to be exactly 1.0f

	// per-batch host -> device upload: chunk geometry, sheet geometry and the chunk -> sheet map are ONE device block filled by ONE copy (three
	// copies of pageable memory were 20 us of a submit's 90 us on the host); geom / sheet_geom / sheet_of point into it
	DevBuf<unsigned char> upload;
	std::vector<unsigned char> upload_host;
	struct { ChunkGeom* p = nullptr; } geom, sheet_geom;
	struct { int* p = nullptr; } sheet_of;
	DevBuf<uint32_t> sheet_mm;
	DevBuf<uint8_t> uni, gflags;
	uint8_t* uni_pinned = nullptr;
	size_t uni_pinned_cap = 0;
	bool uni_valid = false;
	std::vector<ChunkGeom> sheet_geom_host;
	std::vector<int> sheet_of_host, sheet_tab;
	DevBuf<uint32_t> flags, bits, wcnt, wib, seg_tot, chunk_tot;
	DevBuf<uint4> wv4; // per-word vertex record {first vertex id, ex, ey, ez}
	DevBuf<uint2> vcells, icells; // compact surface-cell lists (sized after the scan: <= cells each)
	// fused per-chunk extraction (fused.cuh): one CTA per mesh chunk does label_edges + polygonize + MeshProcessor::init
	int in_flight_hint = 1;   // bmf_ctx_set_batches_in_flight: > 1 = the caller overlaps batches on other contexts, total SM time matters more than latency
	int fused_extract = 1;    // per-chunk kernels for dim <= 64 triangle batches: 1 = when the batch has chunks for every SM several times over (a chunk is one
	                          // CTA's serial job there: ~0.1 ms, so a few hundred chunks finish sooner spread over the whole GPU by the per-segment kernels),
	                          // BMF_FUSED=0 never, BMF_FUSED=2 always
	bool batch_fused = false; // the resident batch went through k_chunk_mesh
	size_t fused_smem = 0;
	DevBuf<unsigned long long> fz_prof, sm_prof;
	int sm_prof_n = 0;
	DevBuf<int> emit_list, mixed_list;
	DevBuf<uint16_t> pack16; // uint16 copy of the index buffer (bmf_batch_download_dma)
	bool fused_prof = false; // BMF_FUSED_PROF=1: per-phase SM clocks of k_chunk_mesh, mean printed to stderr when the batch completes
	int sm_count = 148;
	int use_tma = 1;     // the per-chunk kernels stage a chunk's sign words / word counts with cp.async.bulk + mbarrier (TMA); BMF_TMA=0: 128-bit load loops
	                     // (A/B on the 4096-chunk batch: k_chunk_count 0.058 -> 0.055 ms, k_chunk_emit 0.256 -> 0.249 ms, identical output)
	int reserve_sms = 0; // bmf_ctx_set_reserved_sms: the persistent kernels leave this many SMs (partly) free for another context's kernels
	int work_sms() const { return std::max(1, sm_count - reserve_sms); } // SMs the persistent kernels size their grids for
	int smooth_fused = 1;      // batch path: all smoothing half-steps of a chunk in one CTA out of shared memory (BMF_SMOOTH_FUSED=0: per-step kernels)
	size_t smooth_smem = 0;    // dynamic shared memory of k_smooth_chunks
	int smooth_ctas_per_sm = 8; // resident CTAs per SM of the grid-stride smoothing kernels (tuning knob: BMF_SMOOTH_CTAS_PER_SM)
	DevBuf<float> density, hmap;
	DevBuf<uint8_t> masks;
	DevBuf<ChunkCounts> counts;
	DevBuf<unsigned long long> totals_dev;
	DevBuf<float> pos, color, normal;
	DevBuf<uint8_t> boundary, valence;
	DevBuf<uint32_t> inds;
	// smoothing temporaries
	DevBuf<uint32_t> adj_off, cursor, adj, prim_vbase, block_sums, cls;
	DevBuf<float> dp, dc, dn;
	// qef scratch
	DevBuf<float> qp, qn, qo, qe;
	DevBuf<int32_t> qc;
	// quad emission (quads.cuh)
	DevBuf<uint32_t> wq, wqq;
	DevBuf<uint4> wqv;
	bool quads_processed = false; // MeshProcessor<4> has run on the resident quad batch
	// seam pass (seam.cuh)
	DevBuf<SeamChunk> seam_chunks;
	DevBuf<int32_t> seam_map, seam_group;
	DevBuf<uint8_t> seam_clean;
	DevBuf<uint16_t> seam_layers;
	DevBuf<uint32_t> seam_active, seam_counters, seam_act;
	DevBuf<uint32_t> seam_blk, seam_cnt;
	DevBuf<unsigned long long> seam_base; // [n] chunk bases + [1] total
	DevBuf<float> seam_tris;
	unsigned long long* seam_total_pinned = nullptr;
	int64_t seam_n_tris = -1; // -1: no seam pass run on the resident batch
	cudaEvent_t seam_ev[3] = {};
	float seam_ms[2] = { 0, 0 };

	// device-driven download (bmf_batch_download_enqueue) of the resident batch, kept so that a re-launch after an arena grew can repeat it
	bool dl_pending = false;
	DownloadArgs dl_args;
	int download_ctas = 32; // grid of k_download (BMF_DOWNLOAD_CTAS): the PCIe link is the limit, not the SMs

	HostTotals* totals_pinned = nullptr;
	ChunkCounts* counts_pinned = nullptr;
	size_t counts_pinned_cap = 0;

	cudaEvent_t ev[BMF_NUM_STAGES + 1] = {};
	float stage_ms[BMF_NUM_STAGES] = {};

	// optional per-launch timing (bmf_ctx_set_kernel_timing): one event pair per launch of the last batch
	bool ktiming = false;
	std::vector<cudaEvent_t> kev;
	std::vector<const char*> kname;
	size_t kused = 0;
};

namespace
{

int fail(bmf_ctx* c, int code, const char* what, cudaError_t e = cudaSuccess)
{
	if (c)
	{
		c->err = what;
		if (e != cudaSuccess)
		{
			c->err += ": ";
			c->err += cudaGetErrorString(e);
		}
	}
	return code;
}

#define BMF_CUDA(call)                                                   \
	do                                                                   \
	{                                                                    \
		cudaError_t e__ = (call);                                        \
		if (e__ != cudaSuccess) return fail(ctx, BMF_ERR_CUDA, #call, e__); \
	} while (0)

void ktime_mark(bmf_ctx* ctx, const char* name)
{
	if (ctx->kev.size() < 2 * (ctx->kused + 1))
	{
		cudaEvent_t a, b;
		cudaEventCreate(&a);
		cudaEventCreate(&b);
		ctx->kev.push_back(a);
		ctx->kev.push_back(b);
		ctx->kname.push_back(name);
	}
	ctx->kname[ctx->kused] = name;
	cudaEventRecord(ctx->kev[2 * ctx->kused], ctx->stream);
}

#define BMF_LAUNCH(kernel, grid, block, smem, ...)                              \
	do                                                                          \
	{                                                                           \
		if (ctx->ktiming) ktime_mark(ctx, #kernel);                             \
		kernel<<<(grid), (block), (smem), ctx->stream>>>(__VA_ARGS__);          \
		if (ctx->ktiming) cudaEventRecord(ctx->kev[2 * ctx->kused++ + 1], ctx->stream); \
		ctx->launches++;                                                        \
		cudaError_t e__ = cudaGetLastError();                                   \
		if (e__ != cudaSuccess) return fail(ctx, BMF_ERR_CUDA, #kernel, e__);   \
	} while (0)

float bounding(float gain, int octaves)
{
	float amp = gain, amp_fractal = 1.0f;
	for (int i = 1; i < octaves; i++)
	{
		amp_fractal += amp;
		amp *= gain;
	}
	return 1.0f / amp_fractal;
}

// FastNoiseSIMD object state after the setter sequence of each terrain block function
// (NoiseSampler.cpp:120-125, 160-167, 203-206, 236-242) on a fresh per-thread sampler (library defaults).
void build_sampler(const bmf_sampler_desc& d, SamplerDev* s)
{
	memset(s, 0, sizeof(*s));
	s->kind = d.kind;
	s->world_size = d.world_size;
	NoiseState& ns = s->ns;
	ns.seed = d.seed;
	ns.frequency = 0.01f;
	ns.base = NT_SIMPLEX;
	ns.fractal = 1;
	ns.octaves = 3;
	ns.lacunarity = 2.0f;
	ns.gain = 0.5f;
	ns.fractal_type = FT_FBM;
	ns.perturb = 0;
	ns.perturb_amp = 1.0f / 511.5f;
	ns.perturb_frequency = 0.5f;
	ns.perturb_octaves = 3;
	ns.perturb_lacunarity = 2.0f;
	ns.perturb_gain = 0.5f;
	s->n_mul = 1;
	switch (d.kind)
	{
	case BMF_SAMPLER_TERRAIN2D:
		s->g = 1.0f; s->nm = 64.0f;
		ns.base = NT_VALUE; ns.octaves = 12; ns.gain = 0.5f; ns.lacunarity = 2.0f; ns.fractal_type = FT_FBM;
		break;
	case BMF_SAMPLER_TERRAIN2D_PERT:
		s->g = d.g_scale; s->nm = d.height;
		ns.base = NT_VALUE; ns.perturb = 2; ns.perturb_octaves = d.octaves; ns.perturb_amp = d.amp / 511.5f;
		ns.perturb_frequency = d.frequency; ns.perturb_gain = d.gain; ns.fractal_type = FT_FBM;
		break;
	case BMF_SAMPLER_TERRAIN3D:
		s->g = 0.15f; s->nm = 1.0f; s->dy_half = 1; s->n_mul = 0;
		ns.base = NT_VALUE; ns.octaves = 4; ns.fractal_type = FT_RIGIDMULTI;
		break;
	case BMF_SAMPLER_TERRAIN3D_PERT:
		s->g = 0.15f; s->nm = 48.0f;
		ns.base = NT_SIMPLEX; ns.perturb = 2; ns.octaves = 8; ns.perturb_amp = 1.0f / 511.5f; ns.perturb_frequency = 0.05f;
		ns.fractal_type = FT_RIGIDMULTI;
		break;
	default: break;
	}
	ns.fractal_bounding = bounding(ns.gain, ns.octaves);
	ns.perturb_bounding = bounding(ns.perturb_gain, ns.perturb_octaves);
	s->csg_op = d.csg_op; s->csg_kind_a = d.csg_kind_a; s->csg_kind_b = d.csg_kind_b;
	s->csg_ws_a = d.csg_world_size_a; s->csg_ws_b = d.csg_world_size_b;
	for (int i = 0; i < 3; i++)
	{
		s->csg_off_a[i] = d.csg_offset_a[i];
		s->csg_off_b[i] = d.csg_offset_b[i];
	}
}

bool valid_dim(int d) { return d == 32 || d == 64 || d == 128 || d == 256; }
bool is_terrain2d(int k) { return k == BMF_SAMPLER_TERRAIN2D || k == BMF_SAMPLER_TERRAIN2D_PERT; }
bool is_terrain3d(int k) { return k == BMF_SAMPLER_TERRAIN3D || k == BMF_SAMPLER_TERRAIN3D_PERT; }
bool is_implicit(int k) { return k >= BMF_SAMPLER_SPHERE && k <= BMF_SAMPLER_CSG; }

inline unsigned grid_for(size_t n, int block) { return (unsigned)((n + block - 1) / block); }
int nan_half_step(int iters);
unsigned long long* smooth_prof(bmf_ctx* ctx, int n_chunks);

// MeshProcessor<N> on device arrays (all batch-wide); chunks_dev maps index positions to vertex bases.
template <int N>
int run_smooth(bmf_ctx* ctx, size_t n_verts, size_t n_inds, float* pos, float* color, float* normal, const uint8_t* boundary,
               const uint8_t* valence, const uint32_t* inds, const ChunkCounts* chunks_dev, int n_chunks, int iters, int pb, int smooth, int qef, int final_primal = 1,
               bool grid_path = false, const unsigned long long* tot = nullptr)
{
	// tot != null (batch path): n_verts / n_inds are arena capacities used to size the launches; the kernels read the
	// real counts of the batch from tot[] on the device, so nothing here waits for the host
	// grid_path: the mesh comes from the resident batch (cell lists + per-word bases are valid and every colour is exactly 1)
	if (n_verts == 0 || n_inds < (size_t)N || iters <= 0) return BMF_OK;
	const size_t n_prims = n_inds / N;
	BMF_CUDA(ctx->adj_off.reserve(n_verts));
	if (!grid_path) BMF_CUDA(ctx->cursor.reserve(n_verts));
	BMF_CUDA(ctx->adj.reserve(n_inds));
	BMF_CUDA(ctx->prim_vbase.reserve(n_prims));
	BMF_CUDA(ctx->dp.reserve(3 * n_prims));
	if (!grid_path) BMF_CUDA(ctx->dc.reserve(3 * n_prims));
	float* dcp = grid_path ? nullptr : ctx->dc.p;
	const bool need_dn = smooth || qef;
	if (need_dn) BMF_CUDA(ctx->dn.reserve(3 * n_prims));
	const size_t per_block = (size_t)CTA * SCAN_ITEMS;
	const unsigned nblk = grid_for(n_verts, (int)per_block);
	// the smoothing kernels are grid-stride: at most smooth_ctas_per_sm CTAs per SM stay resident for the whole launch
	const unsigned gmax = (unsigned)(ctx->sm_count * ctx->smooth_ctas_per_sm);
	const unsigned g_prims = std::min(grid_for(n_prims, CTA), gmax), g_verts = std::min(grid_for(n_verts, CTA), gmax);
	const unsigned g_qef = std::min(grid_for(n_verts, 128), gmax * 2);
	BMF_CUDA(ctx->block_sums.reserve(nblk));

	// init: adj_offset = exclusive prefix of init_valence (MeshProcessor.cpp:33-39); on the batch path k_valence_offsets did it already
	if (!grid_path)
	{
		BMF_LAUNCH(k_scan8_partial, nblk, CTA, 0, valence, n_verts, ctx->block_sums.p);
		BMF_LAUNCH(k_scan_block_sums, 1, SCAN_CTA, 0, ctx->block_sums.p, (int)nblk);
		BMF_LAUNCH(k_scan8_final, nblk, CTA, 0, valence, n_verts, ctx->block_sums.p, ctx->adj_off.p);
	}
	if (grid_path && ctx->batch_fused)
	{
		// k_chunk_mesh built the CSR (adj, prim_vbase) together with the mesh
	}
	else if (grid_path)
	{
		BMF_LAUNCH(k_adj_fill, ctx->sm_count * 8, CTA, 0, ctx->L, ctx->wib.p, chunks_dev, ctx->icells.p, ctx->totals_dev.p + 4, inds, ctx->cls.p, ctx->adj_off.p, ctx->adj.p,
		           ctx->prim_vbase.p, tot);
	}
	else
	{
		BMF_CUDA(cudaMemsetAsync(ctx->cursor.p, 0, n_verts * sizeof(uint32_t), ctx->stream));
		BMF_LAUNCH(k_csr_fill<N>, grid_for(n_prims, CTA), CTA, 0, inds, n_prims, chunks_dev, n_chunks, ctx->adj_off.p, ctx->cursor.p, ctx->adj.p, ctx->prim_vbase.p);
		BMF_LAUNCH(k_csr_sort, grid_for(n_verts, CTA), CTA, 0, ctx->adj_off.p, valence, n_verts, ctx->adj.p);
	}

	// one CTA per chunk only pays when there are chunks for every SM; a few (or one large) chunks keep the per-step kernels,
	// which spread every half-step over the whole GPU
	if (grid_path && tot && N == 3 && !smooth && !qef && final_primal && ctx->smooth_fused && n_chunks >= 2 * ctx->sm_count)
	{
		// the whole optimize_dual_grid(iters) + optimize_primal_grid sequence, one CTA per chunk out of shared memory:
		// iters dual steps interleaved with iters primal steps (the last primal is the driver's extra call)
		// the in-loop primal step with set_colors (m == 3, or m == 0 when iters <= 3) turns the zero normals of the processed vertices into
		// NaN in the reference (and in k_primal); nan_step = its half-step index, -1 if there is none
		const int nan_step = nan_half_step(iters);
		BMF_LAUNCH(k_smooth_chunks, (unsigned)std::min(n_chunks, ctx->work_sms()), SMOOTH_CTA, ctx->smooth_smem, chunks_dev, n_chunks, inds, ctx->adj_off.p, ctx->adj.p,
		           valence, boundary, pos, ctx->dp.p, 2 * iters, pb, const_cast<unsigned long long*>(tot), (unsigned)(ctx->smooth_smem / sizeof(float)), normal, nan_step,
		           ctx->batch_fused ? ctx->emit_list.p : nullptr, smooth_prof(ctx, n_chunks));
		return BMF_OK;
	}

	// optimize_dual_grid (MeshProcessor.cpp:130-236)
	const int hard_norm_max = 10;
	const int max_norms = (iters / 2 - 3 < hard_norm_max ? iters / 2 - 3 : hard_norm_max);
	for (int m = 0; m < iters; m++)
	{
		const int face = (m == 0 || m < max_norms || m < 3) ? 1 : 0;
		BMF_LAUNCH(k_dual<N>, g_prims, CTA, 0, inds, ctx->prim_vbase.p, n_prims, pos, color, normal, ctx->dp.p, dcp,
		           need_dn ? ctx->dn.p : nullptr, smooth, face, tot);
		if (m < iters - 1)
		{
			const int set_colors = (m == 3) || (m == 0 && iters <= 3);
			BMF_LAUNCH(k_primal, g_verts, CTA, 0, ctx->adj_off.p, ctx->adj.p, valence, boundary, n_verts, ctx->dp.p, dcp,
			           ctx->dn.p, pos, color, normal, smooth, set_colors, pb, tot);
		}
	}
	// the driver's extra primal call (ChunkGenerator.cpp:120)
	if (final_primal)
		BMF_LAUNCH(k_primal, g_verts, CTA, 0, ctx->adj_off.p, ctx->adj.p, valence, boundary, n_verts, ctx->dp.p, dcp, ctx->dn.p, pos,
		           color, normal, smooth, 0, pb, tot);
	if (qef)
	{
		// build-defined placement: planes = (dual_p, face normal) of the final positions' primitives
		if (qef == 2)
		{
			// planes = (dual_p, normalised sampler gradient at dual_p): MeshProcessor.cpp:224 (commented out in the reference), h = 0.01 (ImplicitSampler.hpp:38)
			BMF_LAUNCH(k_dual<N>, g_prims, CTA, 0, inds, ctx->prim_vbase.p, n_prims, pos, color, normal, ctx->dp.p, dcp, ctx->dn.p, 0, 0, tot);
			BMF_LAUNCH(k_dual_gradient, g_prims, CTA, 0, ctx->sampler, chunks_dev, ctx->geom.p, n_chunks, n_prims, ctx->dp.p, ctx->dn.p, 0.01f, tot);
		}
		else
			BMF_LAUNCH(k_dual<N>, g_prims, CTA, 0, inds, ctx->prim_vbase.p, n_prims, pos, color, normal, ctx->dp.p, dcp, ctx->dn.p, 1, 1, tot);
		BMF_LAUNCH(k_qef_place, g_qef, 128, 0, ctx->adj_off.p, ctx->adj.p, valence, boundary, n_verts, ctx->dp.p, ctx->dn.p, pos, pb, tot);
	}
	return BMF_OK;
}

__global__ void __launch_bounds__(CTA) k_valence_from_inds(const uint32_t* __restrict__ inds, size_t n, uint8_t* __restrict__ valence)
{
	const size_t i = (size_t)blockIdx.x * CTA + threadIdx.x;
	if (i >= n) return;
	const size_t gv = inds[i];
	atomicAdd(reinterpret_cast<unsigned int*>(valence + (gv & ~(size_t)3)), 1u << (8 * (gv & 3)));
}

int elapsed(bmf_ctx* ctx, int a, int b, float* out);

struct MeshCaps
{
	size_t cells, verts, inds;
};

// what the output / smoothing arenas can hold right now (0 = something is not allocated yet)
MeshCaps mesh_caps(const bmf_ctx* ctx)
{
	const bmf_params& p = ctx->params;
	const bool smoothing = p.iters > 0, need_dn = smoothing && (p.smooth_normals || p.qef);
	MeshCaps c;
	c.cells = std::min(ctx->vcells.cap, ctx->icells.cap);
	size_t v = std::min({ ctx->pos.cap / 3, ctx->color.cap / 3, ctx->normal.cap / 3, ctx->boundary.cap, ctx->valence.cap, ctx->cls.cap, ctx->adj_off.cap });
	size_t i = ctx->inds.cap;
	if (smoothing) i = std::min({ i, ctx->adj.cap, ctx->prim_vbase.cap * 3, ctx->dp.cap, need_dn ? ctx->dn.cap : (size_t)-1 });
	c.verts = v > 32 ? v - 32 : 0; // the emitters touch up to 16 bytes past the last vertex (byte-packed words)
	c.inds = i > 32 ? i - 32 : 0; // k_smooth_chunks reads the index / adjacency streams in aligned 16-byte words, up to 16 entries past a chunk's last one
	return c;
}

// grow the arenas to hold (cells, verts, inds) with headroom, so that the following batches of similar size fit
int reserve_mesh(bmf_ctx* ctx, size_t cells, size_t verts, size_t inds)
{
	const bmf_params& p = ctx->params;
	const bool smoothing = p.iters > 0, need_dn = smoothing && (p.smooth_normals || p.qef);
	const size_t C = cells + cells / 4 + 64, V = verts + verts / 4 + 64, I = inds + inds / 4 + 64;
	BMF_CUDA(ctx->vcells.reserve(C));
	BMF_CUDA(ctx->icells.reserve(C));
	BMF_CUDA(ctx->pos.reserve(3 * V));
	if (3 * V > ctx->color.cap) ctx->color_ones = 0;
	BMF_CUDA(ctx->color.reserve(3 * V));
	BMF_CUDA(ctx->normal.reserve(3 * V));
	BMF_CUDA(ctx->boundary.reserve(V));
	BMF_CUDA(ctx->valence.reserve(V));
	BMF_CUDA(ctx->cls.reserve(V));
	BMF_CUDA(ctx->adj_off.reserve(V));
	BMF_CUDA(ctx->inds.reserve(I));
	if (smoothing)
	{
		BMF_CUDA(ctx->adj.reserve(I));
		BMF_CUDA(ctx->prim_vbase.reserve(I / 3 + 1));
		BMF_CUDA(ctx->dp.reserve(I));
		if (need_dn) BMF_CUDA(ctx->dn.reserve(I));
	}
	return BMF_OK;
}

// BMF_FUSED_PROF=1: per-chunk timeline of k_smooth_chunks (start / end ns, SM, path), summarised on stderr when the batch completes
unsigned long long* smooth_prof(bmf_ctx* ctx, int n_chunks)
{
	if (!ctx->fused_prof) return nullptr;
	if (ctx->sm_prof.reserve(4 * (size_t)n_chunks) != cudaSuccess) return nullptr;
	cudaMemsetAsync(ctx->sm_prof.p, 0, 4 * (size_t)n_chunks * sizeof(unsigned long long), ctx->stream);
	ctx->sm_prof_n = n_chunks;
	return ctx->sm_prof.p;
}

// last launch of a batch's sequence: totals and (once per batch) the chunk table to the host through mapped pinned memory (k_publish)
int publish(bmf_ctx* ctx)
{
	const size_t words = ctx->counts_published ? 0 : (size_t)ctx->n * (sizeof(ChunkCounts) / sizeof(uint32_t));
	BMF_LAUNCH(k_publish, words ? std::min(grid_for(words, CTA), (unsigned)(ctx->sm_count * 2)) : 1u, CTA, 0, ctx->totals_dev.p, ctx->totals_pinned->v,
	           reinterpret_cast<const uint32_t*>(ctx->counts.p), reinterpret_cast<uint32_t*>(ctx->counts_pinned), words);
	ctx->counts_published = true;
	return BMF_OK;
}

// K4 + K5 of the resident batch, sized by arena capacity and guarded on the device (k_check_caps): no host round trip
int launch_mesh(bmf_ctx* ctx, bool caps_checked /* k_scan_chunks, launched just before with the same capacities, has already formed the verdict */)
{
	const bmf_params* params = &ctx->params;
	const Layout L = ctx->L;
	const int n = ctx->n, nseg = n * L.S, kind = ctx->sampler.kind;
	cudaStream_t st = ctx->stream;
	const MeshCaps caps = mesh_caps(ctx);
	unsigned long long* tot = ctx->totals_dev.p;
	unsigned long long* list_count = tot + 4;
	if (ctx->batch_fused)
	{
		// ---- dim <= 64, triangles: k_chunk_emit (fused.cuh) does label_edges' emission, polygonize and MeshProcessor::init of a chunk in one CTA
		if (!caps_checked) BMF_LAUNCH(k_check_caps, 1, 32, 0, tot, (unsigned long long)caps.cells, (unsigned long long)caps.verts, (unsigned long long)caps.inds);
		BMF_CUDA(cudaEventRecord(ctx->ev[3], st));
		if (caps.cells == 0 || caps.verts == 0 || caps.inds == 0)
		{
			// nothing allocated yet: k_check_caps has raised the flag unless the batch is empty; bmf_batch_wait sizes the arenas and calls again
			BMF_CUDA(cudaEventRecord(ctx->ev[4], st));
			BMF_CUDA(cudaEventRecord(ctx->ev[5], st));
			BMF_CUDA(cudaEventRecord(ctx->ev[6], st));
			return publish(ctx);
		}
		if (ctx->color_ones < 3 * caps.verts)
		{
			BMF_LAUNCH(k_fill_f32, grid_for(ctx->color.cap, CTA), CTA, 0, ctx->color.p, ctx->color.cap, 1.0f);
			ctx->color_ones = ctx->color.cap;
		}
		FusedArgs A;
		memset(&A, 0, sizeof(A));
		A.L = L; A.n = n;
		A.bits = ctx->bits.p; A.wcnt = ctx->wcnt.p; A.emit_list = ctx->emit_list.p;
		A.chunks = ctx->counts.p;
		A.tot = tot;
		A.vcells = ctx->vcells.p; A.icells = ctx->icells.p;
		A.s = ctx->sampler;
		A.src.density = ctx->density_cur;
		A.src.hmap = (!ctx->density_cur && is_terrain2d(kind)) ? ctx->hmap.p : nullptr;
		A.src.sheet_of = ctx->sheet_of.p;
		A.geom = ctx->geom.p;
		A.pos = ctx->pos.p; A.boundary = ctx->boundary.p; A.normal = ctx->normal.p; A.cls = ctx->cls.p; A.inds = ctx->inds.p;
		A.valence = ctx->valence.p; A.adj_off = ctx->adj_off.p;
		A.adj = params->iters > 0 ? ctx->adj.p : nullptr;
		A.prim_vbase = ctx->prim_vbase.p;
		if (ctx->fused_prof)
		{
			BMF_CUDA(ctx->fz_prof.reserve(16 * (size_t)n));
			BMF_CUDA(cudaMemsetAsync(ctx->fz_prof.p, 0, 16 * (size_t)n * sizeof(unsigned long long), st));
			A.prof = ctx->fz_prof.p;
		}
		if (ctx->use_tma)
			BMF_LAUNCH((k_chunk_emit<FUSED_NT, true>), (unsigned)std::min(n, 2 * ctx->work_sms()), FUSED_NT, ctx->fused_smem, A);
		else
			BMF_LAUNCH((k_chunk_emit<FUSED_NT, false>), (unsigned)std::min(n, 2 * ctx->work_sms()), FUSED_NT, ctx->fused_smem, A);
		BMF_CUDA(cudaEventRecord(ctx->ev[4], st));
		BMF_CUDA(cudaEventRecord(ctx->ev[5], st));
		if (params->iters > 0)
		{
			int rc = run_smooth<3>(ctx, caps.verts, caps.inds, ctx->pos.p, ctx->color.p, ctx->normal.p, ctx->boundary.p, ctx->valence.p, ctx->inds.p, ctx->counts.p, n,
			                       params->iters, params->process_boundary, params->smooth_normals, params->qef, 1, true, tot);
			if (rc) return rc;
		}
		BMF_CUDA(cudaEventRecord(ctx->ev[6], st));
		return publish(ctx);
	}
	if (!caps_checked) BMF_LAUNCH(k_check_caps, 1, 32, 0, tot, (unsigned long long)caps.cells, (unsigned long long)caps.verts, (unsigned long long)caps.inds);
	BMF_CUDA(cudaEventRecord(ctx->ev[3], st));
	if (caps.cells == 0 || caps.verts == 0 || caps.inds == 0)
	{
		// nothing allocated yet (first batch of this kind): k_check_caps has raised the flag unless the batch is empty;
		// bmf_batch_wait sizes the arenas from the totals and calls this function again
		BMF_CUDA(cudaEventRecord(ctx->ev[4], st));
		BMF_CUDA(cudaEventRecord(ctx->ev[5], st));
		BMF_CUDA(cudaEventRecord(ctx->ev[6], st));
		return publish(ctx);
	}
	const size_t V = caps.verts, I = caps.inds;
	const size_t smem_count = (size_t)(L.P + 1) * L.wp * sizeof(uint32_t);
	const size_t smem_bases = smem_count + (size_t)L.ws * (sizeof(uint4) + sizeof(uint32_t)); // + the CTA's work list of active words

	// ---- K4
	DensitySource src;
	src.density = ctx->density_cur;
	src.hmap = (!ctx->density_cur && is_terrain2d(kind)) ? ctx->hmap.p : nullptr;
	src.sheet_of = ctx->sheet_of.p;
	if (params->quads)
	{
		// dual-marching-cubes quads (quads.cuh); MeshProcessor<4> runs when the batch is completed (it needs the real counts)
		BMF_LAUNCH(k_q_bases, (unsigned)n, CTA, 0, ctx->bits.p, L, ctx->wq.p, ctx->counts.p, ctx->wqv.p, ctx->wqq.p, ctx->vcells.p, list_count, tot);
		BMF_CUDA(cudaEventRecord(ctx->ev[4], st));
		BMF_LAUNCH(k_zero_u32, ctx->sm_count * 4, CTA, 0, reinterpret_cast<uint32_t*>(ctx->normal.p), 3 * V, tot, 1, 3);
		if (ctx->color_ones < 3 * V)
		{
			BMF_LAUNCH(k_fill_f32, grid_for(ctx->color.cap, CTA), CTA, 0, ctx->color.p, ctx->color.cap, 1.0f);
			ctx->color_ones = ctx->color.cap;
		}
		BMF_LAUNCH(k_q_emit, (unsigned)(ctx->sm_count * 8), CTA, 0, ctx->bits.p, L, ctx->wq.p, ctx->wqv.p, ctx->wqq.p, ctx->vcells.p, list_count, ctx->counts.p,
		           ctx->sampler, src, ctx->geom.p, ctx->pos.p, ctx->boundary.p, ctx->valence.p, ctx->inds.p, tot);
		BMF_CUDA(cudaEventRecord(ctx->ev[5], st));
		BMF_CUDA(cudaEventRecord(ctx->ev[6], st));
		return publish(ctx);
	}
	if (L.wpt == 4)
		BMF_LAUNCH(k_bases<4>, nseg, CTA, smem_bases, ctx->bits.p, L, ctx->wcnt.p, ctx->seg_tot.p, ctx->counts.p, ctx->wv4.p, ctx->wib.p, ctx->vcells.p,
		           ctx->icells.p, list_count, tot);
	else
		BMF_LAUNCH(k_bases<8>, nseg, CTA, smem_bases, ctx->bits.p, L, ctx->wcnt.p, ctx->seg_tot.p, ctx->counts.p, ctx->wv4.p, ctx->wib.p, ctx->vcells.p,
		           ctx->icells.p, list_count, tot);
	BMF_LAUNCH(k_verts3, ctx->sm_count * 8, CTA, 0, L, ctx->wv4.p, ctx->counts.p, ctx->sampler, src, ctx->geom.p, ctx->vcells.p, list_count, ctx->pos.p, ctx->boundary.p, ctx->cls.p, ctx->normal.p, tot);
	BMF_CUDA(cudaEventRecord(ctx->ev[4], st));
	if (ctx->color_ones < 3 * V)
	{
		// calculate_dual_vertex: color = (1,1,1) (DMCChunk.cpp:681).  The batch path never writes another colour
		// (see run_smooth), so the arena is filled once to its capacity and reused.
		BMF_LAUNCH(k_fill_f32, grid_for(ctx->color.cap, CTA), CTA, 0, ctx->color.p, ctx->color.cap, 1.0f);
		ctx->color_ones = ctx->color.cap;
	}
	BMF_LAUNCH(k_inds3, ctx->sm_count * 8, CTA, 0, L, ctx->wv4.p, ctx->wib.p, ctx->counts.p, ctx->icells.p, list_count, ctx->inds.p, ctx->cls.p, tot);
	BMF_LAUNCH(k_valence_offsets, n, CTA, 0, ctx->cls.p, ctx->counts.p, ctx->valence.p, ctx->adj_off.p, tot, nullptr);
	BMF_CUDA(cudaEventRecord(ctx->ev[5], st));

	// ---- K5 (+K6): MeshProcessor<3>(true, SMOOTH_NORMALS) as ChunkGenerator.cpp:110-124 drives it
	if (params->iters > 0)
	{
		int rc = run_smooth<3>(ctx, V, I, ctx->pos.p, ctx->color.p, ctx->normal.p, ctx->boundary.p, ctx->valence.p, ctx->inds.p, ctx->counts.p, n,
		                       params->iters, params->process_boundary, params->smooth_normals, params->qef, 1, true, tot);
		if (rc) return rc;
	}
	BMF_CUDA(cudaEventRecord(ctx->ev[6], st));
	return publish(ctx);
}

// the in-loop primal step with set_colors (m == 3, or m == 0 when iters <= 3; MeshProcessor.cpp:229-232) turns the zero normals of the
// processed vertices into NaN when smooth normals are off; returns that half-step index, -1 if there is none
int nan_half_step(int iters)
{
	for (int m = 0; m < iters - 1; m++)
		if (m == 3 || (m == 0 && iters <= 3)) return 2 * m + 1;
	return -1;
}

int launch_download(bmf_ctx* ctx)
{
	DownloadArgs A = ctx->dl_args;
	A.d_pos = ctx->pos.p; A.d_normal = ctx->normal.p; A.d_color = ctx->color.p;
	A.d_boundary = ctx->boundary.p; A.d_valence = ctx->valence.p; A.d_inds = ctx->inds.p;
	BMF_LAUNCH(k_download, (unsigned)ctx->download_ctas, CTA, 0, A, ctx->totals_dev.p, ctx->totals_pinned->v);
	return BMF_OK;
}

// completes the resident batch: waits for the stream, publishes totals / per-chunk counts to the host and, if an output
// arena was too small for this batch, grows it and runs the emitters again (the front half of the pipeline is kept)
int finish(bmf_ctx* ctx)
{
	if (ctx->finished) return BMF_OK;
	BMF_CUDA(cudaSetDevice(ctx->device));
	BMF_CUDA(cudaStreamSynchronize(ctx->stream));
	const HostTotals t = *ctx->totals_pinned;
	if (t.v[3]) return fail(ctx, BMF_ERR_INVALID, "bmf_batch_submit: batch exceeds 2^32 cells/vertices/indices; split it");
	ctx->totals[0] = t.v[0];
	ctx->totals[1] = t.v[1];
	ctx->totals[2] = t.v[2];
	ctx->counts_host.assign(ctx->counts_pinned, ctx->counts_pinned + ctx->n);
	if (t.v[7])
	{
		int rc = reserve_mesh(ctx, (size_t)t.v[0], (size_t)t.v[1], (size_t)t.v[2]);
		if (rc) return rc;
		BMF_CUDA(cudaMemsetAsync(ctx->totals_dev.p + 4, 0, 2 * sizeof(unsigned long long), ctx->stream)); // list counters
		rc = launch_mesh(ctx, false);
		if (rc) return rc;
		if (ctx->dl_pending && !(ctx->params.quads && ctx->params.iters > 0))
		{
			rc = launch_download(ctx); // the first attempt returned at once (TOT_SMALL was set)
			if (rc) return rc;
		}
		BMF_CUDA(cudaStreamSynchronize(ctx->stream));
		if (ctx->totals_pinned->v[7]) return fail(ctx, BMF_ERR_NOMEM, "bmf_batch_wait: output arenas still too small after growing");
		ctx->relaunches++;
	}
	if (ctx->params.quads && ctx->params.iters > 0 && !ctx->quads_processed && ctx->totals[1] && ctx->totals[2] >= 4)
	{
		// MeshProcessor<4>(true, smooth_normals): init + optimize_dual_grid(iters, pb) + optimize_primal_grid(false, false, pb)
		// (the sequence ChunkGenerator.cpp:271-280 / DebugScene.cpp:256-262 run on quads), generic CSR path, real counts
		BMF_CUDA(cudaEventRecord(ctx->ev[5], ctx->stream));
		int rc = run_smooth<4>(ctx, (size_t)ctx->totals[1], (size_t)ctx->totals[2], ctx->pos.p, ctx->color.p, ctx->normal.p, ctx->boundary.p, ctx->valence.p,
		                       ctx->inds.p, ctx->counts.p, ctx->n, ctx->params.iters, ctx->params.process_boundary, ctx->params.smooth_normals, 0);
		if (rc) return rc;
		ctx->color_ones = 0; // the generic path rewrites colours (with the same value); do not rely on the fill any more
		BMF_CUDA(cudaEventRecord(ctx->ev[6], ctx->stream));
		BMF_CUDA(cudaStreamSynchronize(ctx->stream));
		ctx->quads_processed = true;
	}
	for (int s = 0; s < 6; s++) elapsed(ctx, s, s + 1, &ctx->stage_ms[s]);
	elapsed(ctx, 0, 6, &ctx->stage_ms[BMF_STAGE_TOTAL]);
	ctx->finished = true;
	if (ctx->fused_prof && ctx->batch_fused && ctx->fz_prof.p)
	{
		std::vector<unsigned long long> h(16 * (size_t)ctx->n);
		cudaMemcpy(h.data(), ctx->fz_prof.p, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
		double sum[16] = {};
		int m = 0;
		for (int j = 0; j < ctx->n; j++)
		{
			const unsigned long long* r = &h[16 * (size_t)j];
			if (!r[0] || !r[7]) continue;
			m++;
			for (int k = 1; k <= 7; k++) sum[k] += (double)(r[k] - r[k - 1]);
		}
		fprintf(stderr, "k_chunk_emit phases (mean SM cycles over %d chunks): stage %.0f scan %.0f lists %.0f verts %.0f inds %.0f valence %.0f adj %.0f\n",
		        m, m ? sum[1] / m : 0, m ? sum[2] / m : 0, m ? sum[3] / m : 0, m ? sum[4] / m : 0, m ? sum[5] / m : 0, m ? sum[6] / m : 0, m ? sum[7] / m : 0);
	}
	if (ctx->fused_prof && ctx->sm_prof.p && ctx->sm_prof_n)
	{
		std::vector<unsigned long long> h(4 * (size_t)ctx->sm_prof_n);
		cudaMemcpy(h.data(), ctx->sm_prof.p, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
		unsigned long long t0 = ~0ull, t1 = 0;
		for (int j = 0; j < ctx->sm_prof_n; j++)
			if (h[4 * (size_t)j + 1]) { t0 = std::min(t0, h[4 * (size_t)j]); t1 = std::max(t1, h[4 * (size_t)j + 1]); }
		double busy[3] = {}, verts[3] = {};
		int cnt[3] = {};
		std::vector<double> sm_busy(1024, 0.0), sm_last(1024, 0.0);
		for (int j = 0; j < ctx->sm_prof_n; j++)
		{
			const unsigned long long* r = &h[4 * (size_t)j];
			if (!r[1]) continue;
			const int path = (int)(r[2] >> 8) & 3, sm = (int)(r[2] & 255);
			busy[path] += (double)(r[1] - r[0]); verts[path] += (double)r[3]; cnt[path]++;
			sm_busy[sm] += (double)(r[1] - r[0]); sm_last[sm] = std::max(sm_last[sm], (double)(r[1] - t0));
		}
		double bsum = 0, bmax = 0, lmin = 1e30; int nsm = 0;
		for (int k = 0; k < 1024; k++) if (sm_busy[k] > 0) { bsum += sm_busy[k]; bmax = std::max(bmax, sm_busy[k]); lmin = std::min(lmin, sm_last[k]); nsm++; }
		fprintf(stderr, "k_smooth_chunks timeline: span %.1f us over %d SMs; busy per SM mean %.1f max %.1f us; earliest SM finished at %.1f us\n", (t1 - t0) * 1e-3, nsm, nsm ? bsum / nsm * 1e-3 : 0, bmax * 1e-3, lmin * 1e-3);
		for (int k = 0; k < 3; k++)
			fprintf(stderr, "   path %d (%s): %d chunks, %.0f verts, %.1f us CTA time, %.2f ns per vertex\n", k, k == 0 ? "P+D in smem" : k == 1 ? "D in smem" : "global", cnt[k], verts[k], busy[k] * 1e-3, verts[k] ? busy[k] / verts[k] : 0);
		ctx->sm_prof_n = 0;
	}
	if (ctx->dl_pending)
	{
		ctx->dl_pending = false;
		const unsigned long long e = ctx->totals_pinned->v[TOT_DLERR];
		if (e == 1) return fail(ctx, BMF_ERR_NOMEM, "bmf_batch_download_enqueue: host buffers too small for this batch (nothing was written)");
		if (e == 2) return fail(ctx, BMF_ERR_INVALID, "bmf_batch_download_enqueue: uint16 indices requested but a chunk has >= 65536 vertices (nothing was written)");
	}
	return BMF_OK;
}

int elapsed(bmf_ctx* ctx, int a, int b, float* out)
{
	cudaError_t e = cudaEventElapsedTime(out, ctx->ev[a], ctx->ev[b]);
	if (e != cudaSuccess) *out = 0.0f;
	return 0;
}

} // namespace

extern "C" {

int bmf_batch_download_async(bmf_ctx* ctx, float* pos, float* normal, float* color, uint8_t* boundary, uint8_t* valence, uint32_t* indices);

const char* bmf_version(void) { return "bmf_b200 0.1 (sm_100a)"; }

const char* bmf_last_error(const bmf_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int bmf_ctx_create(int device, bmf_ctx** out)
{
	if (!out) return BMF_ERR_INVALID;
	*out = nullptr;
	int count = 0;
	cudaError_t e = cudaGetDeviceCount(&count);
	if (e != cudaSuccess || count == 0 || device < 0 || device >= count)
	{
		fprintf(stderr, "bmf_b200: no usable CUDA device %d (%s); there is no CPU fallback\n", device,
		        e != cudaSuccess ? cudaGetErrorString(e) : "device index out of range");
		return BMF_ERR_CUDA;
	}
	bmf_ctx* ctx = new bmf_ctx();
	ctx->device = device;
	if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess)
	{
		delete ctx;
		return BMF_ERR_CUDA;
	}
	for (int i = 0; i <= BMF_NUM_STAGES; i++) cudaEventCreate(&ctx->ev[i]);
	cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
	if (ctx->sm_count <= 0) ctx->sm_count = 148;
	if (const char* e = getenv("BMF_SMOOTH_CTAS_PER_SM")) { const int v = atoi(e); if (v > 0 && v <= 4096) ctx->smooth_ctas_per_sm = v; }
	if (const char* e = getenv("BMF_SMOOTH_FUSED")) ctx->smooth_fused = atoi(e) != 0;
	if (const char* e = getenv("BMF_DOWNLOAD_CTAS")) { const int v = atoi(e); if (v > 0 && v <= 4096) ctx->download_ctas = v; }
	{
		int optin = 0;
		cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
		ctx->smooth_smem = optin > 4096 ? (size_t)optin - 1024 : 0; // leave room for the kernel's static shared memory
		{
			// ... and 8 KB for CTAs of other kernels (another stream's sampling kernel) beside a smoothing CTA; a chunk whose positions no longer fit next to its
			// dual points keeps them in global memory, which costs the same per vertex.  BMF_SMOOTH_SMEM_LEAVE=<bytes> overrides.
			size_t leave = 8192;
			if (const char* e = getenv("BMF_SMOOTH_SMEM_LEAVE")) leave = (size_t)std::max(0, atoi(e));
			if (leave + 65536 < ctx->smooth_smem) ctx->smooth_smem -= leave;
		}
		if (!ctx->smooth_smem || cudaFuncSetAttribute(k_smooth_chunks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->smooth_smem) != cudaSuccess)
		{
			cudaGetLastError();
			ctx->smooth_fused = 0;
		}
	}
	if (const char* e = getenv("BMF_FUSED")) ctx->fused_extract = atoi(e);
	if (const char* e = getenv("BMF_FUSED_PROF")) ctx->fused_prof = atoi(e) != 0;
	if (const char* e = getenv("BMF_TMA")) ctx->use_tma = atoi(e) != 0;
	ctx->fused_smem = fused_smem_bytes(make_layout(64), FUSED_NT);
	if (cudaFuncSetAttribute(k_chunk_emit<FUSED_NT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->fused_smem) != cudaSuccess ||
	    cudaFuncSetAttribute(k_chunk_emit<FUSED_NT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ctx->fused_smem) != cudaSuccess)
	{
		cudaGetLastError();
		ctx->fused_extract = 0;
	}
	cudaFuncSetAttribute(k_scan_chunks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)SCAN_CTA * SCAN_PER_THREAD * sizeof(ChunkCounts)));
	// k_bases keeps its segment's sign planes plus a work list of active words in dynamic shared memory (57 KB at dim 256)
	cudaFuncSetAttribute(k_bases<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
	cudaFuncSetAttribute(k_bases<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
	cudaGetLastError();
	cudaMallocHost((void**)&ctx->totals_pinned, sizeof(HostTotals));
	*out = ctx;
	return BMF_OK;
}

void bmf_ctx_destroy(bmf_ctx* ctx)
{
	if (!ctx) return;
	cudaSetDevice(ctx->device);
	cudaStreamSynchronize(ctx->stream);
	ctx->upload.release(); ctx->geom.p = ctx->sheet_geom.p = nullptr; ctx->sheet_of.p = nullptr; ctx->flags.release(); ctx->bits.release(); ctx->wcnt.release(); ctx->wv4.release(); ctx->wib.release();
	ctx->seg_tot.release(); ctx->chunk_tot.release(); ctx->vcells.release(); ctx->icells.release(); ctx->density.release(); ctx->hmap.release(); ctx->masks.release();
	ctx->counts.release(); ctx->totals_dev.release(); ctx->pos.release(); ctx->color.release(); ctx->normal.release();
	ctx->boundary.release(); ctx->valence.release(); ctx->inds.release(); ctx->adj_off.release(); ctx->cls.release(); ctx->cursor.release();
	ctx->adj.release(); ctx->prim_vbase.release(); ctx->block_sums.release(); ctx->dp.release(); ctx->dc.release(); ctx->dn.release();
	ctx->qp.release(); ctx->qn.release(); ctx->qo.release(); ctx->qe.release(); ctx->qc.release();
	if (ctx->totals_pinned) cudaFreeHost(ctx->totals_pinned);
	if (ctx->counts_pinned) cudaFreeHost(ctx->counts_pinned);
	if (ctx->uni_pinned) cudaFreeHost(ctx->uni_pinned);
	if (ctx->seam_total_pinned) cudaFreeHost(ctx->seam_total_pinned);
	ctx->wq.release(); ctx->wqq.release(); ctx->wqv.release();
	ctx->fz_prof.release(); ctx->sm_prof.release(); ctx->emit_list.release(); ctx->mixed_list.release(); ctx->pack16.release();
	ctx->seam_chunks.release(); ctx->seam_map.release(); ctx->seam_group.release(); ctx->seam_clean.release(); ctx->seam_layers.release(); ctx->seam_active.release(); ctx->seam_counters.release(); ctx->seam_act.release(); ctx->seam_blk.release(); ctx->seam_cnt.release();
	ctx->seam_base.release(); ctx->seam_tris.release();
	for (cudaEvent_t e : ctx->seam_ev)
		if (e) cudaEventDestroy(e);
	ctx->sheet_mm.release(); ctx->uni.release(); ctx->gflags.release();
	for (int i = 0; i <= BMF_NUM_STAGES; i++)
		if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
	for (cudaEvent_t e : ctx->kev) cudaEventDestroy(e);
	cudaStreamDestroy(ctx->stream);
	delete ctx;
}

void bmf_sampler_defaults(bmf_sampler_desc* d, int kind)
{
	if (!d) return;
	memset(d, 0, sizeof(*d));
	d->kind = kind;
	d->world_size = 256.0f;
	// WorldOctree.cpp:47-54
	d->g_scale = 0.25f;
	d->height = 75.0f;
	d->octaves = 13;
	d->amp = 0.87f;
	d->frequency = 0.585f;
	d->gain = 0.488f;
	d->seed = 1337;
	d->csg_op = BMF_CSG_UNION;
	d->csg_kind_a = BMF_SAMPLER_SPHERE;
	d->csg_kind_b = BMF_SAMPLER_TORUS_Z;
	d->csg_world_size_a = d->csg_world_size_b = 256.0f;
}

int bmf_sampler_set(bmf_ctx* ctx, const bmf_sampler_desc* desc)
{
	if (!ctx || !desc) return BMF_ERR_INVALID;
	const int k = desc->kind;
	if (!(is_implicit(k) || is_terrain2d(k) || is_terrain3d(k) || k == BMF_SAMPLER_HOST_DENSITY))
		return fail(ctx, BMF_ERR_INVALID, "bmf_sampler_set: unknown sampler kind");
	if (k == BMF_SAMPLER_CSG && !(desc->csg_kind_a >= 0 && desc->csg_kind_a <= 3 && desc->csg_kind_b >= 0 && desc->csg_kind_b <= 3 && desc->csg_op >= 0 && desc->csg_op <= 2))
		return fail(ctx, BMF_ERR_INVALID, "bmf_sampler_set: CSG operands must be primitive kinds 0..3 and op 0..2");
	if (k == BMF_SAMPLER_TERRAIN2D_PERT && desc->octaves < 1)
		return fail(ctx, BMF_ERR_INVALID, "bmf_sampler_set: octaves must be >= 1");
	ctx->sampler_desc = *desc;
	build_sampler(*desc, &ctx->sampler);
	ctx->sampler_set = true;
	return BMF_OK;
}

int bmf_batch_submit(bmf_ctx* ctx, const bmf_chunk_desc* chunks, int n, const bmf_params* params, const float* density_in)
{
	if (!ctx) return BMF_ERR_INVALID;
	if (!chunks || !params || n <= 0) return fail(ctx, BMF_ERR_INVALID, "bmf_batch_submit: null argument or empty batch");
	if (!ctx->sampler_set) return fail(ctx, BMF_ERR_STATE, "bmf_batch_submit: no sampler set");
	if (!valid_dim(params->dim)) return fail(ctx, BMF_ERR_INVALID, "bmf_batch_submit: dim must be 32, 64, 128 or 256");
	if (params->iters < 0) return fail(ctx, BMF_ERR_INVALID, "bmf_batch_submit: iters < 0");
	if (params->quads && (params->qef || params->keep_masks)) return fail(ctx, BMF_ERR_INVALID, "bmf_batch_submit: quads cannot be combined with qef or keep_masks");
	if (params->qef < 0 || params->qef > 2) return fail(ctx, BMF_ERR_INVALID, "bmf_batch_submit: qef must be 0, 1 (face normals) or 2 (sampler gradients)");
	if (params->qef == 2 && !is_implicit(ctx->sampler.kind)) return fail(ctx, BMF_ERR_INVALID, "bmf_batch_submit: qef = 2 needs an analytic sampler (the value callback of the noise samplers is the constant 0, NoiseSampler.cpp:99-102)");
	const int kind = ctx->sampler.kind;
	if (kind == BMF_SAMPLER_HOST_DENSITY && !density_in) return fail(ctx, BMF_ERR_INVALID, "bmf_batch_submit: HOST_DENSITY needs a density block");
	BMF_CUDA(cudaSetDevice(ctx->device));

	ctx->have_batch = false;
	ctx->finished = false;
	ctx->uni_valid = false;
	ctx->seam_n_tris = -1;
	ctx->quads_processed = false;
	ctx->dl_pending = false;
	ctx->kused = 0;
	ctx->n = n;
	ctx->params = *params;
	ctx->descs.assign(chunks, chunks + n);
	const Layout L = ctx->L = make_layout(params->dim);
	const int d = L.d;
	const size_t nvox = (size_t)d * d * d;
	const size_t n_words = (size_t)n * L.wc;
	const int nseg = n * L.S;

	// DMCChunk::label_grid geometry (DMCChunk.cpp:94-98), IEEE single ops in this order
	ctx->geom_host.resize(n);
	for (int i = 0; i < n; i++)
	{
		const bmf_chunk_desc& c = chunks[i];
		ChunkGeom g;
		g.delta = c.size * (1.0f + c.overlap * 2.0f) / (float)(d - 1);
		const float so = c.size * c.overlap;
		g.ox = c.pos[0] - so;
		g.oy = c.pos[1] - so;
		g.oz = c.pos[2] - so;
		ctx->geom_host[i] = g;
	}
	const bool fused = ctx->batch_fused = ctx->fused_extract && !params->quads && d <= 64 && (ctx->fused_extract >= 2 || ctx->in_flight_hint > 1 || n >= 8 * ctx->sm_count);
	bool gen2d = false; // 2-D terrain on the per-chunk path: k_chunk_count makes the sign words itself, k_terrain2d_bits is not launched
	BMF_CUDA(ctx->flags.reserve(n));
	BMF_CUDA(ctx->bits.reserve(n_words));
	BMF_CUDA(ctx->wcnt.reserve(n_words));
	BMF_CUDA(ctx->chunk_tot.reserve(3 * (size_t)n));
	if (fused) BMF_CUDA(ctx->emit_list.reserve(n));
	if (!fused)
	{
		BMF_CUDA(ctx->wv4.reserve(n_words));
		BMF_CUDA(ctx->wib.reserve(n_words));
		BMF_CUDA(ctx->seg_tot.reserve(3 * (size_t)nseg));
	}
	BMF_CUDA(ctx->counts.reserve(n));
	BMF_CUDA(ctx->totals_dev.reserve(TOT_SLOTS));
	if ((size_t)n > ctx->counts_pinned_cap)
	{
		BMF_CUDA(cudaStreamSynchronize(ctx->stream)); // an earlier batch may still be copying into the old buffer
		if (ctx->counts_pinned) cudaFreeHost(ctx->counts_pinned);
		ctx->counts_pinned = nullptr;
		BMF_CUDA(cudaMallocHost((void**)&ctx->counts_pinned, sizeof(ChunkCounts) * (size_t)n));
		ctx->counts_pinned_cap = n;
	}

	const bool host_density = kind == BMF_SAMPLER_HOST_DENSITY;
	const bool need_density = host_density || is_terrain3d(kind) || params->keep_density;
	const float* density_dev = nullptr;
	ctx->ext_density = nullptr;
	if (host_density && params->density_on_device)
	{
		density_dev = density_in;
		ctx->ext_density = density_in;
	}
	else if (need_density)
	{
		BMF_CUDA(ctx->density.reserve(n * nvox));
		density_dev = ctx->density.p;
	}
	ctx->density_valid = need_density;
	ctx->masks_valid = params->keep_masks != 0;
	if (params->keep_masks) BMF_CUDA(ctx->masks.reserve(n * nvox));
	int n_sheets = 0;
	if (is_terrain2d(kind))
	{
		// the noise sheet is a function of (overlap_pos.x, overlap_pos.z, delta) only: one sheet per unique triple
		// (open-addressing table over the three 32-bit patterns; sheets are numbered in order of first appearance.  A std::map here was a fifth of a
		// small batch's submit time on the host)
		size_t cap = 64;
		while (cap < 2 * (size_t)n) cap <<= 1;
		ctx->sheet_tab.assign(cap, -1);
		ctx->sheet_of_host.resize(n);
		ctx->sheet_geom_host.clear();
		for (int i = 0; i < n; i++)
		{
			const ChunkGeom& g = ctx->geom_host[i];
			uint32_t k[3];
			memcpy(&k[0], &g.ox, 4); memcpy(&k[1], &g.oz, 4); memcpy(&k[2], &g.delta, 4);
			uint64_t h = (uint64_t)k[0] * 0x9E3779B97F4A7C15ull ^ (uint64_t)k[1] * 0xC2B2AE3D27D4EB4Full ^ (uint64_t)k[2] * 0x165667B19E3779F9ull;
			size_t slot = (size_t)((h ^ (h >> 29)) & (cap - 1));
			for (;;)
			{
				const int s = ctx->sheet_tab[slot];
				if (s < 0)
				{
					ctx->sheet_tab[slot] = (int)ctx->sheet_geom_host.size();
					ctx->sheet_geom_host.push_back(g);
					break;
				}
				const ChunkGeom& q = ctx->sheet_geom_host[s];
				if (memcmp(&q.ox, &g.ox, 4) == 0 && memcmp(&q.oz, &g.oz, 4) == 0 && memcmp(&q.delta, &g.delta, 4) == 0) break;
				slot = (slot + 1) & (cap - 1);
			}
			ctx->sheet_of_host[i] = ctx->sheet_tab[slot];
		}
		n_sheets = (int)ctx->sheet_geom_host.size();
		BMF_CUDA(ctx->hmap.reserve((size_t)n_sheets * d * d));
		BMF_CUDA(ctx->sheet_mm.reserve(3 * (size_t)n_sheets));
		BMF_CUDA(ctx->uni.reserve(n));
		if ((size_t)n > ctx->uni_pinned_cap)
		{
			BMF_CUDA(cudaStreamSynchronize(ctx->stream));
			if (ctx->uni_pinned) cudaFreeHost(ctx->uni_pinned);
			ctx->uni_pinned = nullptr;
			BMF_CUDA(cudaMallocHost((void**)&ctx->uni_pinned, (size_t)n));
			ctx->uni_pinned_cap = n;
		}
	}

	cudaStream_t st = ctx->stream;
	{
		// [geom n][sheet_geom n_sheets][sheet_of n] (16-byte records first: every part stays aligned)
		const size_t off_sg = sizeof(ChunkGeom) * (size_t)n, off_so = off_sg + sizeof(ChunkGeom) * (size_t)n_sheets, up_bytes = off_so + sizeof(int) * (size_t)(n_sheets ? n : 0);
		ctx->upload_host.resize(up_bytes);
		memcpy(ctx->upload_host.data(), ctx->geom_host.data(), off_sg);
		if (n_sheets)
		{
			memcpy(ctx->upload_host.data() + off_sg, ctx->sheet_geom_host.data(), sizeof(ChunkGeom) * (size_t)n_sheets);
			memcpy(ctx->upload_host.data() + off_so, ctx->sheet_of_host.data(), sizeof(int) * (size_t)n);
		}
		BMF_CUDA(ctx->upload.reserve(up_bytes));
		ctx->geom.p = reinterpret_cast<ChunkGeom*>(ctx->upload.p);
		ctx->sheet_geom.p = reinterpret_cast<ChunkGeom*>(ctx->upload.p + off_sg);
		ctx->sheet_of.p = reinterpret_cast<int*>(ctx->upload.p + off_so);
		BMF_CUDA(cudaMemcpyAsync(ctx->upload.p, ctx->upload_host.data(), up_bytes, cudaMemcpyHostToDevice, st));
	}
	if (host_density && !params->density_on_device)
		BMF_CUDA(cudaMemcpyAsync(ctx->density.p, density_in, sizeof(float) * n * nvox, cudaMemcpyHostToDevice, st));
	BMF_CUDA(cudaMemsetAsync(ctx->flags.p, 0, sizeof(uint32_t) * n, st));
	BMF_CUDA(cudaMemsetAsync(ctx->totals_dev.p + TOT_MESH, 0, 4 * sizeof(unsigned long long), st)); // list lengths and tickets: TOT_MESH, TOT_TICKET, TOT_CAND, TOT_CTICKET
	if (!fused) BMF_CUDA(cudaMemsetAsync(ctx->chunk_tot.p, 0, sizeof(uint32_t) * 3 * n, st)); // k_count adds to it; k_chunk_count stores


	// ---- K1 / K2
	BMF_CUDA(cudaEventRecord(ctx->ev[0], st));
	float* dens_w = need_density ? ctx->density.p : nullptr;
	if (is_implicit(kind))
	{
		BMF_LAUNCH(k_sample_implicit, (unsigned)(n_words / SAMPLE_WORDS_PER_CTA), CTA, 0, ctx->sampler, ctx->geom.p, L, ctx->bits.p, dens_w, ctx->flags.p);
	}
	else if (is_terrain2d(kind))
	{
		BMF_CUDA(cudaMemsetAsync(ctx->sheet_mm.p, 0xFF, sizeof(uint32_t) * n_sheets, st));
		BMF_CUDA(cudaMemsetAsync(ctx->sheet_mm.p + n_sheets, 0, sizeof(uint32_t) * 2 * n_sheets, st));
		BMF_LAUNCH(k_terrain2d_sheet<NT_VALUE>, grid_for((size_t)n_sheets * d * d, CTA), CTA, 0, ctx->sampler, ctx->sheet_geom.p, d, L.ld, ctx->hmap.p, n_sheets, ctx->sheet_mm.p);
		if (dens_w)
			BMF_LAUNCH(k_terrain2d_density, (unsigned)(n_words / SAMPLE_WORDS_PER_CTA), CTA, 0, ctx->sampler, ctx->geom.p, L, ctx->hmap.p, ctx->sheet_of.p, ctx->bits.p,
			           dens_w, ctx->flags.p);
		else
		{
			// no density block wanted: chunks entirely above / below the surface are classified from the sheet range and
			// skipped; the rest get compare-only sign words (binary search + warp bit transpose, no per-voxel arithmetic)
			// (totals_dev[12] = length of the list of chunks the surface crosses; k_scan_chunks does not touch that slot)
			BMF_CUDA(ctx->mixed_list.reserve(n));
			BMF_LAUNCH(k_terrain2d_classify, grid_for(n, CTA), CTA, 0, ctx->sampler, ctx->geom.p, d, ctx->sheet_of.p, ctx->sheet_mm.p, n_sheets, n, ctx->flags.p, ctx->uni.p,
			           ctx->mixed_list.p, ctx->totals_dev.p + TOT_CAND);
			if (fused)
				gen2d = true; // k_chunk_count makes the sign words of the listed chunks itself (fused.cuh, GEN)
			else
			{
				const size_t groups = (size_t)n * L.d * L.zc * L.zc / (CTA / 32);
				BMF_LAUNCH(k_terrain2d_bits, (unsigned)std::min(groups, (size_t)ctx->sm_count * 16), CTA, 0, ctx->sampler, ctx->geom.p, L, ctx->hmap.p, ctx->sheet_of.p,
				           ctx->mixed_list.p, ctx->totals_dev.p + TOT_CAND, ctx->bits.p, ctx->flags.p);
			}
			ctx->uni_valid = true;
		}
	}
	else if (kind == BMF_SAMPLER_TERRAIN3D)
	{
		BMF_LAUNCH(k_terrain3d<NT_VALUE>, (unsigned)(n_words / (CTA / 32)), CTA, 0, ctx->sampler, ctx->geom.p, L, ctx->bits.p, dens_w, ctx->flags.p);
	}
	else if (kind == BMF_SAMPLER_TERRAIN3D_PERT)
	{
		BMF_LAUNCH(k_terrain3d<NT_SIMPLEX>, (unsigned)(n_words / (CTA / 32)), CTA, 0, ctx->sampler, ctx->geom.p, L, ctx->bits.p, dens_w, ctx->flags.p);
	}
	else
	{
		BMF_CUDA(ctx->gflags.reserve(n_words / PACK_UNROLL));
		BMF_LAUNCH(k_pack_density, (unsigned)(n_words / (PACK_UNROLL * (CTA / 32))), CTA, 0, density_dev, ctx->bits.p, ctx->gflags.p, n_words);
		BMF_LAUNCH(k_reduce_flags, n, CTA, 0, ctx->gflags.p, L.wc / PACK_UNROLL, ctx->flags.p);
	}
	BMF_CUDA(cudaEventRecord(ctx->ev[1], st));

	// ---- K3
	const size_t smem_count = (size_t)(L.P + 1) * L.wp * sizeof(uint32_t);
	uint8_t* masks_w = params->keep_masks ? ctx->masks.p : nullptr;
	if (fused)
	{
		// chunks without a mesh keep chunk_tot = 0 (k_scan_chunks only reads the totals of mesh chunks); TOT_MESH / TOT_CTICKET restart
		const bool have_cand = ctx->uni_valid; // the 2-D terrain classifier has listed the chunks it could not cull
		Gen2D G;
		G.s = ctx->sampler; G.geom = ctx->geom.p; G.hmap = ctx->hmap.p; G.sheet_of = ctx->sheet_of.p;
		if (gen2d)
			BMF_LAUNCH((k_chunk_count<COUNT_NT, false, true>), (unsigned)std::min(n, 6 * ctx->work_sms()), COUNT_NT, (size_t)(L.d + 1) * L.wp * sizeof(uint32_t), ctx->bits.p, ctx->flags.p, L, n,
			           ctx->mixed_list.p, ctx->totals_dev.p + TOT_CAND, ctx->wcnt.p, ctx->chunk_tot.p, masks_w, ctx->totals_dev.p, G);
		else if (ctx->use_tma)
			BMF_LAUNCH((k_chunk_count<COUNT_NT, true, false>), (unsigned)std::min(n, 6 * ctx->work_sms()), COUNT_NT, (size_t)(L.d + 1) * L.wp * sizeof(uint32_t), ctx->bits.p, ctx->flags.p, L, n,
			           have_cand ? ctx->mixed_list.p : nullptr, ctx->totals_dev.p + TOT_CAND, ctx->wcnt.p, ctx->chunk_tot.p, masks_w, ctx->totals_dev.p, G);
		else
			BMF_LAUNCH((k_chunk_count<COUNT_NT, false, false>), (unsigned)std::min(n, 6 * ctx->work_sms()), COUNT_NT, (size_t)(L.d + 1) * L.wp * sizeof(uint32_t), ctx->bits.p, ctx->flags.p, L, n,
			           have_cand ? ctx->mixed_list.p : nullptr, ctx->totals_dev.p + TOT_CAND, ctx->wcnt.p, ctx->chunk_tot.p, masks_w, ctx->totals_dev.p, G);
	}
	else if (params->quads)
	{
		BMF_CUDA(ctx->wq.reserve(n_words));
		BMF_CUDA(ctx->wqq.reserve(n_words));
		BMF_CUDA(ctx->wqv.reserve(n_words));
		BMF_LAUNCH(k_q_count, (unsigned)(n_words / CTA), CTA, 0, ctx->bits.p, ctx->flags.p, L, ctx->wq.p, ctx->chunk_tot.p);
	}
	else if (L.wpt == 4)
		BMF_LAUNCH(k_count<4>, nseg, CTA, smem_count, ctx->bits.p, ctx->flags.p, L, ctx->wcnt.p, ctx->seg_tot.p, ctx->chunk_tot.p, masks_w);
	else
		BMF_LAUNCH(k_count<8>, nseg, CTA, smem_count, ctx->bits.p, ctx->flags.p, L, ctx->wcnt.p, ctx->seg_tot.p, ctx->chunk_tot.p, masks_w);
	BMF_CUDA(cudaEventRecord(ctx->ev[2], st));

	// ---- scan, then the emitters straight away: their launches are sized by the arenas' capacity and guarded on the
	// device, so the host does not wait here (bmf_batch_wait / any query completes the batch)
	const MeshCaps caps0 = mesh_caps(ctx); // the capacities launch_mesh is about to size the emitters with: the scan forms k_check_caps' verdict itself
	BMF_LAUNCH(k_scan_chunks, 1, SCAN_CTA, (size_t)SCAN_CTA * SCAN_PER_THREAD * sizeof(ChunkCounts), ctx->chunk_tot.p, ctx->flags.p, n, ctx->counts.p, ctx->totals_dev.p,
	           fused ? ctx->emit_list.p : nullptr, (unsigned long long)caps0.cells, (unsigned long long)caps0.verts, (unsigned long long)caps0.inds);
	ctx->counts_published = false;
	ctx->density_cur = density_dev;
	ctx->have_batch = true;
	int rc = launch_mesh(ctx, true);
	if (rc)
	{
		ctx->have_batch = false;
		return rc;
	}
	return BMF_OK;
}

int bmf_batch_wait(bmf_ctx* ctx)
{
	if (!ctx) return BMF_ERR_INVALID;
	if (!ctx->have_batch) return fail(ctx, BMF_ERR_STATE, "bmf_batch_wait: no batch submitted");
	int rc = finish(ctx);
	if (rc) return rc;
	BMF_CUDA(cudaStreamSynchronize(ctx->stream)); // copies enqueued after the batch (bmf_batch_download_async)
	return BMF_OK;
}

int bmf_batch_totals(bmf_ctx* ctx, int64_t* n_cells, int64_t* n_verts, int64_t* n_inds)
{
	if (!ctx) return BMF_ERR_INVALID;
	if (!ctx->have_batch) return fail(ctx, BMF_ERR_STATE, "bmf_batch_totals: no batch submitted");
	{
		int frc = finish(ctx);
		if (frc) return frc;
	}
	if (n_cells) *n_cells = (int64_t)ctx->totals[0];
	if (n_verts) *n_verts = (int64_t)ctx->totals[1];
	if (n_inds) *n_inds = (int64_t)ctx->totals[2];
	return BMF_OK;
}

int bmf_batch_chunk_info(bmf_ctx* ctx, int i, bmf_chunk_info* out)
{
	if (!ctx || !out) return BMF_ERR_INVALID;
	if (!ctx->have_batch) return fail(ctx, BMF_ERR_STATE, "bmf_batch_chunk_info: no batch submitted");
	{
		int frc = finish(ctx);
		if (frc) return frc;
	}
	if (i < 0 || i >= ctx->n) return fail(ctx, BMF_ERR_INVALID, "bmf_batch_chunk_info: chunk index out of range");
	const ChunkCounts& c = ctx->counts_host[i];
	const ChunkGeom& g = ctx->geom_host[i];
	out->contains_mesh = (int32_t)c.contains_mesh;
	out->n_cells = (int32_t)c.n_cells;
	out->n_verts = (int32_t)c.n_verts;
	out->n_inds = (int32_t)c.n_inds;
	out->vert_offset = (int64_t)c.vert_base;
	out->ind_offset = (int64_t)c.ind_base;
	out->overlap_pos[0] = g.ox; out->overlap_pos[1] = g.oy; out->overlap_pos[2] = g.oz;
	out->scale = g.delta;
	out->flags = 0;
	out->reserved = 0;
	const bmf_params& pr = ctx->params;
	if (ctx->color_ones >= 3 * (size_t)(c.vert_base + c.n_verts)) out->flags |= BMF_CHUNK_COLOR_ONE;
	if (!pr.smooth_normals && !pr.qef && nan_half_step(pr.iters) < 0) out->flags |= BMF_CHUNK_NORMAL_ZERO;
	if (c.n_verts < 65536u) out->flags |= BMF_CHUNK_INDEX16;
	return BMF_OK;
}

int bmf_batch_chunk_infos(bmf_ctx* ctx, bmf_chunk_info* out)
{
	if (!ctx || !out) return BMF_ERR_INVALID;
	if (!ctx->have_batch) return fail(ctx, BMF_ERR_STATE, "bmf_batch_chunk_infos: no batch submitted");
	for (int i = 0; i < ctx->n; i++)
	{
		int rc = bmf_batch_chunk_info(ctx, i, out + i);
		if (rc) return rc;
	}
	return BMF_OK;
}

int bmf_batch_download(bmf_ctx* ctx, float* pos, float* normal, float* color, uint8_t* boundary, uint8_t* valence, uint32_t* indices)
{
	int rc = bmf_batch_download_async(ctx, pos, normal, color, boundary, valence, indices);
	return rc ? rc : bmf_batch_wait(ctx);
}

int bmf_batch_download_async(bmf_ctx* ctx, float* pos, float* normal, float* color, uint8_t* boundary, uint8_t* valence, uint32_t* indices)
{
	if (!ctx) return BMF_ERR_INVALID;
	if (!ctx->have_batch) return fail(ctx, BMF_ERR_STATE, "bmf_batch_download: no batch submitted");
	{
		int frc = finish(ctx);
		if (frc) return frc;
	}
	BMF_CUDA(cudaSetDevice(ctx->device));
	const size_t V = ctx->totals[1], I = ctx->totals[2];
	cudaStream_t st = ctx->stream;
	if (V)
	{
		if (pos) BMF_CUDA(cudaMemcpyAsync(pos, ctx->pos.p, sizeof(float) * 3 * V, cudaMemcpyDeviceToHost, st));
		if (normal) BMF_CUDA(cudaMemcpyAsync(normal, ctx->normal.p, sizeof(float) * 3 * V, cudaMemcpyDeviceToHost, st));
		if (color) BMF_CUDA(cudaMemcpyAsync(color, ctx->color.p, sizeof(float) * 3 * V, cudaMemcpyDeviceToHost, st));
		if (boundary) BMF_CUDA(cudaMemcpyAsync(boundary, ctx->boundary.p, V, cudaMemcpyDeviceToHost, st));
		if (valence) BMF_CUDA(cudaMemcpyAsync(valence, ctx->valence.p, V, cudaMemcpyDeviceToHost, st));
	}
	if (I && indices) BMF_CUDA(cudaMemcpyAsync(indices, ctx->inds.p, sizeof(uint32_t) * I, cudaMemcpyDeviceToHost, st));
	return BMF_OK;
}

int bmf_batch_download_enqueue(bmf_ctx* ctx, const bmf_download_desc* desc)
{
	if (!ctx || !desc) return BMF_ERR_INVALID;
	if (!ctx->have_batch) return fail(ctx, BMF_ERR_STATE, "bmf_batch_download_enqueue: no batch submitted");
	if (desc->indices32 && desc->indices16) return fail(ctx, BMF_ERR_INVALID, "bmf_batch_download_enqueue: pass indices32 or indices16, not both");
	if (desc->cap_verts < 0 || desc->cap_inds < 0) return fail(ctx, BMF_ERR_INVALID, "bmf_batch_download_enqueue: negative capacity");
	BMF_CUDA(cudaSetDevice(ctx->device));
	if (ctx->params.quads && ctx->params.iters > 0)
	{
		// MeshProcessor<4> on a quad batch runs when the batch is completed (it needs the real counts on the host)
		int frc = finish(ctx);
		if (frc) return frc;
		ctx->finished = false;
	}
	DownloadArgs A;
	memset(&A, 0, sizeof(A));
	void* host[7] = { desc->pos, desc->normal, desc->color, desc->boundary, desc->valence, desc->indices32, desc->indices16 };
	void* dev[7] = { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr };
	const uintptr_t align[7] = { 4, 4, 4, 1, 1, 4, 2 };
	for (int k = 0; k < 7; k++)
	{
		if (!host[k]) continue;
		if (reinterpret_cast<uintptr_t>(host[k]) & (align[k] - 1)) return fail(ctx, BMF_ERR_INVALID, "bmf_batch_download_enqueue: a host buffer is not aligned to its element type");
		if (cudaHostGetDevicePointer(&dev[k], host[k], 0) != cudaSuccess)
		{
			cudaGetLastError();
			return fail(ctx, BMF_ERR_INVALID, "bmf_batch_download_enqueue: host buffers must be page-locked and device-mapped (bmf_host_alloc / bmf_host_register)");
		}
	}
	A.pos = (float*)dev[0]; A.normal = (float*)dev[1]; A.color = (float*)dev[2]; A.boundary = (uint8_t*)dev[3]; A.valence = (uint8_t*)dev[4];
	A.inds32 = (uint32_t*)dev[5]; A.inds16 = (uint16_t*)dev[6];
	A.cap_verts = (unsigned long long)desc->cap_verts;
	A.cap_inds = (unsigned long long)desc->cap_inds;
	ctx->dl_args = A;
	ctx->dl_pending = true;
	ctx->finished = false; // bmf_batch_wait must also complete (and check) this download
	return launch_download(ctx);
}

int bmf_batch_download_dma(bmf_ctx* ctx, const bmf_download_desc* desc)
{
	if (!ctx || !desc) return BMF_ERR_INVALID;
	if (!ctx->have_batch) return fail(ctx, BMF_ERR_STATE, "bmf_batch_download_dma: no batch submitted");
	if (desc->indices32 && desc->indices16) return fail(ctx, BMF_ERR_INVALID, "bmf_batch_download_dma: pass indices32 or indices16, not both");
	if (desc->cap_verts < 0 || desc->cap_inds < 0) return fail(ctx, BMF_ERR_INVALID, "bmf_batch_download_dma: negative capacity");
	{
		int frc = finish(ctx); // the host learns the sizes of the transfers here
		if (frc) return frc;
	}
	BMF_CUDA(cudaSetDevice(ctx->device));
	const size_t V = ctx->totals[1], I = ctx->totals[2];
	const bool want_v = desc->pos || desc->normal || desc->color || desc->boundary || desc->valence, want_i = desc->indices32 || desc->indices16;
	if ((want_v && V > (size_t)desc->cap_verts) || (want_i && I > (size_t)desc->cap_inds))
		return fail(ctx, BMF_ERR_NOMEM, "bmf_batch_download_dma: host buffers too small for this batch (nothing was written)");
	if (desc->indices16 && ctx->totals_pinned->v[TOT_MAXV] > 65535ull)
		return fail(ctx, BMF_ERR_INVALID, "bmf_batch_download_dma: uint16 indices requested but a chunk has >= 65536 vertices (nothing was written)");
	cudaStream_t st = ctx->stream;
	if (V)
	{
		if (desc->pos) BMF_CUDA(cudaMemcpyAsync(desc->pos, ctx->pos.p, sizeof(float) * 3 * V, cudaMemcpyDeviceToHost, st));
		if (desc->normal) BMF_CUDA(cudaMemcpyAsync(desc->normal, ctx->normal.p, sizeof(float) * 3 * V, cudaMemcpyDeviceToHost, st));
		if (desc->color) BMF_CUDA(cudaMemcpyAsync(desc->color, ctx->color.p, sizeof(float) * 3 * V, cudaMemcpyDeviceToHost, st));
		if (desc->boundary) BMF_CUDA(cudaMemcpyAsync(desc->boundary, ctx->boundary.p, V, cudaMemcpyDeviceToHost, st));
		if (desc->valence) BMF_CUDA(cudaMemcpyAsync(desc->valence, ctx->valence.p, V, cudaMemcpyDeviceToHost, st));
	}
	if (I && desc->indices32) BMF_CUDA(cudaMemcpyAsync(desc->indices32, ctx->inds.p, sizeof(uint32_t) * I, cudaMemcpyDeviceToHost, st));
	if (I && desc->indices16)
	{
		BMF_CUDA(ctx->pack16.reserve(I + 8));
		BMF_LAUNCH(k_pack_indices16, std::min(grid_for((I + 7) / 8, CTA), (unsigned)(ctx->sm_count * 4)), CTA, 0, ctx->inds.p, I, ctx->pack16.p);
		BMF_CUDA(cudaMemcpyAsync(desc->indices16, ctx->pack16.p, sizeof(uint16_t) * I, cudaMemcpyDeviceToHost, st));
	}
	return BMF_OK;
}

int bmf_host_alloc(size_t bytes, void** out)
{
	if (!out) return BMF_ERR_INVALID;
	*out = nullptr;
	return cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable | cudaHostAllocMapped) == cudaSuccess ? BMF_OK : BMF_ERR_NOMEM;
}

void bmf_host_free(void* p)
{
	if (p) cudaFreeHost(p);
}

int bmf_host_register(void* p, size_t bytes)
{
	if (!p || !bytes) return BMF_ERR_INVALID;
	return cudaHostRegister(p, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped) == cudaSuccess ? BMF_OK : BMF_ERR_CUDA;
}

int bmf_host_unregister(void* p)
{
	if (!p) return BMF_ERR_INVALID;
	return cudaHostUnregister(p) == cudaSuccess ? BMF_OK : BMF_ERR_CUDA;
}

int bmf_batch_copy_chunk(bmf_ctx* ctx, int i, void* dual_vertices, uint32_t* indices, uint32_t* bits, uint8_t* masks, float* density)
{
	if (!ctx) return BMF_ERR_INVALID;
	if (!ctx->have_batch) return fail(ctx, BMF_ERR_STATE, "bmf_batch_copy_chunk: no batch submitted");
	{
		int frc = finish(ctx);
		if (frc) return frc;
	}
	if (i < 0 || i >= ctx->n) return fail(ctx, BMF_ERR_INVALID, "bmf_batch_copy_chunk: chunk index out of range");
	BMF_CUDA(cudaSetDevice(ctx->device));
	const ChunkCounts& c = ctx->counts_host[i];
	const Layout& L = ctx->L;
	const size_t nvox = (size_t)L.d * L.d * L.d;
	cudaStream_t st = ctx->stream;
	uint8_t uniform = 0; // chunk skipped by the 2-D terrain classifier: all ones (1) / all zeros (2); fetched on demand (this is the only host use)
	if (ctx->uni_valid)
	{
		BMF_CUDA(cudaMemcpyAsync(ctx->uni_pinned, ctx->uni.p + i, 1, cudaMemcpyDeviceToHost, ctx->stream));
		BMF_CUDA(cudaStreamSynchronize(ctx->stream));
		uniform = ctx->uni_pinned[0];
	}
	if (bits && uniform) memset(bits, uniform == 1 ? 0xFF : 0x00, sizeof(uint32_t) * L.wc);
	else if (bits) BMF_CUDA(cudaMemcpyAsync(bits, ctx->bits.p + (size_t)i * L.wc, sizeof(uint32_t) * L.wc, cudaMemcpyDeviceToHost, st));
	if (masks)
	{
		if (!ctx->masks_valid) return fail(ctx, BMF_ERR_STATE, "bmf_batch_copy_chunk: masks were not kept (params.keep_masks)");
		BMF_CUDA(cudaMemcpyAsync(masks, ctx->masks.p + (size_t)i * nvox, nvox, cudaMemcpyDeviceToHost, st));
	}
	if (density)
	{
		if (!ctx->density_valid) return fail(ctx, BMF_ERR_STATE, "bmf_batch_copy_chunk: density was not kept (params.keep_density)");
		const float* src = ctx->ext_density ? ctx->ext_density : ctx->density.p;
		BMF_CUDA(cudaMemcpyAsync(density, src + (size_t)i * nvox, sizeof(float) * nvox, cudaMemcpyDeviceToHost, st));
	}
	if (indices && c.n_inds) BMF_CUDA(cudaMemcpyAsync(indices, ctx->inds.p + c.ind_base, sizeof(uint32_t) * c.n_inds, cudaMemcpyDeviceToHost, st));
	std::vector<float> p, nn, col;
	std::vector<uint8_t> bd, val;
	if (dual_vertices && c.n_verts)
	{
		const size_t nv = c.n_verts;
		p.resize(3 * nv); nn.resize(3 * nv); col.resize(3 * nv); bd.resize(nv); val.resize(nv);
		BMF_CUDA(cudaMemcpyAsync(p.data(), ctx->pos.p + 3 * c.vert_base, sizeof(float) * 3 * nv, cudaMemcpyDeviceToHost, st));
		BMF_CUDA(cudaMemcpyAsync(nn.data(), ctx->normal.p + 3 * c.vert_base, sizeof(float) * 3 * nv, cudaMemcpyDeviceToHost, st));
		BMF_CUDA(cudaMemcpyAsync(col.data(), ctx->color.p + 3 * c.vert_base, sizeof(float) * 3 * nv, cudaMemcpyDeviceToHost, st));
		BMF_CUDA(cudaMemcpyAsync(bd.data(), ctx->boundary.p + c.vert_base, nv, cudaMemcpyDeviceToHost, st));
		BMF_CUDA(cudaMemcpyAsync(val.data(), ctx->valence.p + c.vert_base, nv, cudaMemcpyDeviceToHost, st));
	}
	int rc = bmf_batch_wait(ctx);
	if (rc) return rc;
	if (dual_vertices && c.n_verts)
	{
		// DualVertex, 84 bytes (Vertices.hpp:5-24): boundary@0 index@4 valence@8 init_valence@9 adj_next@10
		// adj_offset@12 s@20 p@36 n@48 color@72; fields the reference leaves uninitialised are zero here
		uint8_t* out = (uint8_t*)dual_vertices;
		const bool processed = ctx->params.iters > 0 && c.n_inds > 0;
		uint32_t off = 0;
		for (size_t v = 0; v < c.n_verts; v++)
		{
			uint8_t* r = out + 84 * v;
			memset(r, 0, 84);
			r[0] = bd[v];
			uint32_t idx = (uint32_t)v;
			memcpy(r + 4, &idx, 4);
			r[8] = processed ? val[v] : 0;
			r[9] = val[v];
			r[10] = processed ? val[v] : 0;
			uint32_t ao = processed ? off : 0;
			memcpy(r + 12, &ao, 4);
			memcpy(r + 36, &p[3 * v], 12);
			memcpy(r + 48, &nn[3 * v], 12);
			memcpy(r + 72, &col[3 * v], 12);
			off += val[v];
		}
	}
	return BMF_OK;
}

int bmf_batch_stage_ms(bmf_ctx* ctx, float* ms)
{
	if (!ctx || !ms) return BMF_ERR_INVALID;
	int rc = bmf_batch_wait(ctx);
	if (rc) return rc;
	memcpy(ms, ctx->stage_ms, sizeof(ctx->stage_ms));
	return BMF_OK;
}

int64_t bmf_ctx_launch_count(const bmf_ctx* ctx) { return ctx ? ctx->launches : 0; }

int bmf_ctx_set_kernel_timing(bmf_ctx* ctx, int on)
{
	if (!ctx) return BMF_ERR_INVALID;
	ctx->ktiming = on != 0;
	return BMF_OK;
}

int bmf_ctx_kernel_times(bmf_ctx* ctx, int cap, const char** names, float* ms)
{
	if (!ctx) return BMF_ERR_INVALID;
	if (!ctx->have_batch) return fail(ctx, BMF_ERR_STATE, "bmf_ctx_kernel_times: no batch submitted");
	int rc = bmf_batch_wait(ctx);
	if (rc) return rc;
	int n = (int)ctx->kused;
	for (int i = 0; i < n && i < cap; i++)
	{
		float t = 0.0f;
		cudaEventElapsedTime(&t, ctx->kev[2 * i], ctx->kev[2 * i + 1]);
		if (names) names[i] = ctx->kname[i];
		if (ms) ms[i] = t;
	}
	return n;
}

int bmf_ctx_set_reserved_sms(bmf_ctx* ctx, int n)
{
	if (!ctx) return BMF_ERR_INVALID;
	if (n < 0 || n >= ctx->sm_count) return fail(ctx, BMF_ERR_INVALID, "bmf_ctx_set_reserved_sms: 0 <= n < number of SMs");
	ctx->reserve_sms = n;
	return BMF_OK;
}

int bmf_ctx_set_batches_in_flight(bmf_ctx* ctx, int n)
{
	if (!ctx) return BMF_ERR_INVALID;
	ctx->in_flight_hint = n > 1 ? n : 1;
	return BMF_OK;
}

void* bmf_ctx_stream(const bmf_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int bmf_batch_device_ptrs(bmf_ctx* ctx, void** pos, void** indices, void** bits, void** density)
{
	if (!ctx) return BMF_ERR_INVALID;
	if (!ctx->have_batch) return fail(ctx, BMF_ERR_STATE, "bmf_batch_device_ptrs: no batch submitted");
	{
		int frc = finish(ctx);
		if (frc) return frc;
	}
	if (pos) *pos = ctx->pos.p;
	if (indices) *indices = ctx->inds.p;
	if (bits) *bits = ctx->bits.p;
	if (density) *density = ctx->density_valid ? (void*)(ctx->ext_density ? ctx->ext_density : ctx->density.p) : nullptr;
	return BMF_OK;
}

int bmf_mesh_process_steps(bmf_ctx* ctx, float* pos, float* color, float* normal, const uint8_t* boundary, const uint8_t* valence_in, int n_verts,
                           const uint32_t* indices, int n_inds, int prim_n, int iters, int process_boundary, int smooth_normals, int final_primal);

int bmf_mesh_process(bmf_ctx* ctx, float* pos, float* color, float* normal, const uint8_t* boundary, const uint8_t* valence_in, int n_verts,
                     const uint32_t* indices, int n_inds, int prim_n, int iters, int process_boundary, int smooth_normals)
{
	return bmf_mesh_process_steps(ctx, pos, color, normal, boundary, valence_in, n_verts, indices, n_inds, prim_n, iters, process_boundary, smooth_normals, 1);
}

int bmf_mesh_process_steps(bmf_ctx* ctx, float* pos, float* color, float* normal, const uint8_t* boundary, const uint8_t* valence_in, int n_verts,
                           const uint32_t* indices, int n_inds, int prim_n, int iters, int process_boundary, int smooth_normals, int final_primal)
{
	(void)valence_in; // init_valence is recomputed from the index buffer (identical whenever the caller's was consistent)
	if (!ctx) return BMF_ERR_INVALID;
	if (!pos || !color || !boundary || !indices || n_verts < 0 || n_inds < 0) return fail(ctx, BMF_ERR_INVALID, "bmf_mesh_process: null argument");
	if (prim_n != 3 && prim_n != 4) return fail(ctx, BMF_ERR_INVALID, "bmf_mesh_process: prim_n must be 3 or 4");
	if (smooth_normals && !normal) return fail(ctx, BMF_ERR_INVALID, "bmf_mesh_process: smooth_normals needs a normal array");
	if (n_verts == 0 || n_inds < prim_n || iters <= 0) return BMF_OK;
	{
		// init_valence is a uint8 in the reference (Vertices.hpp:12) and the device keeps four byte counters per 32-bit word: a vertex
		// used more than 255 times would carry into its neighbour's counter, so such a mesh is refused instead of being smoothed wrongly
		std::vector<uint16_t> uses((size_t)n_verts, 0);
		for (int i = 0; i < n_inds; i++)
		{
			if (indices[i] >= (uint32_t)n_verts) return fail(ctx, BMF_ERR_INVALID, "bmf_mesh_process: index out of range");
			if (i < (n_inds / prim_n) * prim_n && ++uses[indices[i]] > 255) return fail(ctx, BMF_ERR_INVALID, "bmf_mesh_process: a vertex is referenced more than 255 times (init_valence is 8 bits)");
		}
	}
	BMF_CUDA(cudaSetDevice(ctx->device));
	ctx->have_batch = false; // the arenas are reused
	ctx->color_ones = 0;
	const size_t V = n_verts, I = (size_t)(n_inds / prim_n) * prim_n;
	cudaStream_t st = ctx->stream;
	BMF_CUDA(ctx->pos.reserve(3 * V + 4));
	BMF_CUDA(ctx->color.reserve(3 * V + 4));
	BMF_CUDA(ctx->normal.reserve(3 * V + 4));
	BMF_CUDA(ctx->boundary.reserve(V + 16));
	BMF_CUDA(ctx->valence.reserve(V + 16));
	BMF_CUDA(ctx->inds.reserve(I + 4));
	BMF_CUDA(ctx->counts.reserve(1));
	BMF_CUDA(cudaMemcpyAsync(ctx->pos.p, pos, sizeof(float) * 3 * V, cudaMemcpyHostToDevice, st));
	BMF_CUDA(cudaMemcpyAsync(ctx->color.p, color, sizeof(float) * 3 * V, cudaMemcpyHostToDevice, st));
	if (normal) BMF_CUDA(cudaMemcpyAsync(ctx->normal.p, normal, sizeof(float) * 3 * V, cudaMemcpyHostToDevice, st));
	else BMF_CUDA(cudaMemsetAsync(ctx->normal.p, 0, sizeof(float) * 3 * V, st));
	BMF_CUDA(cudaMemcpyAsync(ctx->boundary.p, boundary, V, cudaMemcpyHostToDevice, st));
	BMF_CUDA(cudaMemcpyAsync(ctx->inds.p, indices, sizeof(uint32_t) * I, cudaMemcpyHostToDevice, st));
	ChunkCounts one;
	memset(&one, 0, sizeof(one));
	one.contains_mesh = 1; one.n_verts = (uint32_t)V; one.n_inds = (uint32_t)I;
	BMF_CUDA(cudaMemcpyAsync(ctx->counts.p, &one, sizeof(one), cudaMemcpyHostToDevice, st));
	BMF_CUDA(cudaMemsetAsync(ctx->valence.p, 0, V + 16, st));
	BMF_LAUNCH(k_valence_from_inds, grid_for(I, CTA), CTA, 0, ctx->inds.p, I, ctx->valence.p);
	int rc = (prim_n == 3)
		? run_smooth<3>(ctx, V, I, ctx->pos.p, ctx->color.p, ctx->normal.p, ctx->boundary.p, ctx->valence.p, ctx->inds.p, ctx->counts.p, 1, iters, process_boundary, smooth_normals, 0, final_primal)
		: run_smooth<4>(ctx, V, I, ctx->pos.p, ctx->color.p, ctx->normal.p, ctx->boundary.p, ctx->valence.p, ctx->inds.p, ctx->counts.p, 1, iters, process_boundary, smooth_normals, 0, final_primal);
	if (rc) return rc;
	BMF_CUDA(cudaMemcpyAsync(pos, ctx->pos.p, sizeof(float) * 3 * V, cudaMemcpyDeviceToHost, st));
	BMF_CUDA(cudaMemcpyAsync(color, ctx->color.p, sizeof(float) * 3 * V, cudaMemcpyDeviceToHost, st));
	if (normal) BMF_CUDA(cudaMemcpyAsync(normal, ctx->normal.p, sizeof(float) * 3 * V, cudaMemcpyDeviceToHost, st));
	BMF_CUDA(cudaStreamSynchronize(st));
	return BMF_OK;
}

int bmf_qef_solve(bmf_ctx* ctx, const float* positions, const float* normals, const int32_t* counts, int m, float* out_pos, float* out_err)
{
	if (!ctx) return BMF_ERR_INVALID;
	if (!positions || !normals || !counts || !out_pos || !out_err || m < 0) return fail(ctx, BMF_ERR_INVALID, "bmf_qef_solve: null argument");
	if (m == 0) return BMF_OK;
	BMF_CUDA(cudaSetDevice(ctx->device));
	cudaStream_t st = ctx->stream;
	const size_t M = m;
	BMF_CUDA(ctx->qp.reserve(36 * M));
	BMF_CUDA(ctx->qn.reserve(36 * M));
	BMF_CUDA(ctx->qc.reserve(M));
	BMF_CUDA(ctx->qo.reserve(3 * M));
	BMF_CUDA(ctx->qe.reserve(M));
	BMF_CUDA(cudaMemcpyAsync(ctx->qp.p, positions, sizeof(float) * 36 * M, cudaMemcpyHostToDevice, st));
	BMF_CUDA(cudaMemcpyAsync(ctx->qn.p, normals, sizeof(float) * 36 * M, cudaMemcpyHostToDevice, st));
	BMF_CUDA(cudaMemcpyAsync(ctx->qc.p, counts, sizeof(int32_t) * M, cudaMemcpyHostToDevice, st));
	BMF_LAUNCH(k_qef_batch, grid_for(M, 128), 128, 0, ctx->qp.p, ctx->qn.p, ctx->qc.p, m, ctx->qo.p, ctx->qe.p);
	BMF_CUDA(cudaMemcpyAsync(out_pos, ctx->qo.p, sizeof(float) * 3 * M, cudaMemcpyDeviceToHost, st));
	BMF_CUDA(cudaMemcpyAsync(out_err, ctx->qe.p, sizeof(float) * M, cudaMemcpyDeviceToHost, st));
	BMF_CUDA(cudaStreamSynchronize(st));
	return BMF_OK;
}

int bmf_sampler_gradient(bmf_ctx* ctx, const float* points, int64_t m, float h, float* out)
{
	if (!ctx) return BMF_ERR_INVALID;
	if (m < 0 || (m && (!points || !out))) return fail(ctx, BMF_ERR_INVALID, "bmf_sampler_gradient: bad arguments");
	if (!ctx->sampler_set) return fail(ctx, BMF_ERR_STATE, "bmf_sampler_gradient: no sampler set");
	if (ctx->sampler.kind == BMF_SAMPLER_HOST_DENSITY) return fail(ctx, BMF_ERR_STATE, "bmf_sampler_gradient: a HOST_DENSITY sampler has no device value function (call the host callback)");
	if (m == 0) return BMF_OK;
	BMF_CUDA(cudaSetDevice(ctx->device));
	// scratch: the QEF arenas are free outside bmf_qef_solve
	const size_t M = (size_t)m;
	BMF_CUDA(ctx->qp.reserve(3 * M));
	BMF_CUDA(ctx->qo.reserve(3 * M));
	cudaStream_t st = ctx->stream;
	BMF_CUDA(cudaMemcpyAsync(ctx->qp.p, points, sizeof(float) * 3 * M, cudaMemcpyHostToDevice, st));
	BMF_LAUNCH(k_sampler_gradient, std::min(grid_for(M, CTA), (unsigned)(ctx->sm_count * 8)), CTA, 0, ctx->sampler, ctx->qp.p, M, h, ctx->qo.p);
	BMF_CUDA(cudaMemcpyAsync(out, ctx->qo.p, sizeof(float) * 3 * M, cudaMemcpyDeviceToHost, st));
	BMF_CUDA(cudaStreamSynchronize(st));
	return BMF_OK;
}

int bmf_color_map(bmf_ctx* ctx, const float* pos, int64_t n, float* color)
{
	if (!ctx) return BMF_ERR_INVALID;
	if (n < 0 || (n && (!pos || !color))) return fail(ctx, BMF_ERR_INVALID, "bmf_color_map: bad arguments");
	if (n == 0) return BMF_OK;
	BMF_CUDA(cudaSetDevice(ctx->device));
	// ColorMapper::ColorMapper + get_noise (ColorMapper.cpp:6-9, 45-47): a fresh FastNoiseSIMD object (library defaults, seed 1337),
	// SetNoiseType(SimplexFractal), SetFractalOctaves(4), SetFractalType(FBM)
	bmf_sampler_desc d;
	bmf_sampler_defaults(&d, BMF_SAMPLER_SPHERE);
	SamplerDev sd;
	build_sampler(d, &sd);
	NoiseState ns = sd.ns;
	ns.seed = 1337;
	ns.base = NT_SIMPLEX; ns.fractal = 1; ns.octaves = 4; ns.fractal_type = FT_FBM; ns.perturb = 0;
	ns.fractal_bounding = bounding(ns.gain, ns.octaves);
	const size_t N = (size_t)n;
	BMF_CUDA(ctx->qp.reserve(3 * N));
	BMF_CUDA(ctx->qo.reserve(3 * N));
	cudaStream_t st = ctx->stream;
	BMF_CUDA(cudaMemcpyAsync(ctx->qp.p, pos, sizeof(float) * 3 * N, cudaMemcpyHostToDevice, st));
	BMF_LAUNCH(k_color_map, std::min(grid_for(N, CTA), (unsigned)(ctx->sm_count * 8)), CTA, 0, ns, ctx->qp.p, N, ctx->qo.p);
	BMF_CUDA(cudaMemcpyAsync(color, ctx->qo.p, sizeof(float) * 3 * N, cudaMemcpyDeviceToHost, st));
	BMF_CUDA(cudaStreamSynchronize(st));
	return BMF_OK;
}

int bmf_mesh_collapse_bad_quads(bmf_ctx* ctx, float* pos, int n_verts, uint32_t* quads, int64_t n_quads, uint8_t* destroyed, uint8_t* adj_next, uint32_t* flushed,
                                int64_t* n_flushed, int64_t* bad_count)
{
	if (!ctx) return BMF_ERR_INVALID;
	if (n_flushed) *n_flushed = 0;
	if (bad_count) *bad_count = 0;
	if (n_verts < 0 || n_quads < 0 || (n_quads && (!pos || !quads))) return fail(ctx, BMF_ERR_INVALID, "bmf_mesh_collapse_bad_quads: bad arguments");
	if (n_quads > 0x3FFFFFFFll) return fail(ctx, BMF_ERR_INVALID, "bmf_mesh_collapse_bad_quads: more than 2^30 quads");
	if (n_verts == 0 || n_quads == 0) return BMF_OK; // MeshProcessor::init returns at once (MeshProcessor.cpp:28-29)
	const size_t V = (size_t)n_verts, Q = (size_t)n_quads, I = 4 * Q;
	{
		std::vector<uint16_t> uses(V, 0);
		for (size_t i = 0; i < I; i++)
		{
			if (quads[i] >= (uint32_t)n_verts) return fail(ctx, BMF_ERR_INVALID, "bmf_mesh_collapse_bad_quads: index out of range");
			if (++uses[quads[i]] > 255) return fail(ctx, BMF_ERR_INVALID, "bmf_mesh_collapse_bad_quads: a vertex is referenced more than 255 times (adj_next is 8 bits)");
		}
	}
	BMF_CUDA(cudaSetDevice(ctx->device));
	ctx->have_batch = false; // the arenas are reused
	ctx->color_ones = 0;
	cudaStream_t st = ctx->stream;
	BMF_CUDA(ctx->pos.reserve(3 * V + 4));
	BMF_CUDA(ctx->valence.reserve(V + 16));
	BMF_CUDA(ctx->boundary.reserve(Q + 16));   // Primitive::destroyed
	BMF_CUDA(ctx->inds.reserve(I + 4));
	BMF_CUDA(ctx->cls.reserve(I + 4));         // the flushed index buffer
	BMF_CUDA(ctx->adj_off.reserve(V));
	BMF_CUDA(ctx->cursor.reserve(V));
	BMF_CUDA(ctx->adj.reserve(2 * I + 4));     // init's lists + one new 4-entry list per collapse (adj_block.push_back, :384-387)
	BMF_CUDA(ctx->prim_vbase.reserve(Q));
	BMF_CUDA(ctx->counts.reserve(1));
	BMF_CUDA(ctx->totals_dev.reserve(TOT_SLOTS));
	const size_t per_block = (size_t)CTA * SCAN_ITEMS;
	const unsigned nblk = grid_for(V, (int)per_block);
	BMF_CUDA(ctx->block_sums.reserve(nblk));
	BMF_CUDA(cudaMemcpyAsync(ctx->pos.p, pos, sizeof(float) * 3 * V, cudaMemcpyHostToDevice, st));
	BMF_CUDA(cudaMemcpyAsync(ctx->inds.p, quads, sizeof(uint32_t) * I, cudaMemcpyHostToDevice, st));
	ChunkCounts one;
	memset(&one, 0, sizeof(one));
	one.contains_mesh = 1; one.n_verts = (uint32_t)V; one.n_inds = (uint32_t)I;
	BMF_CUDA(cudaMemcpyAsync(ctx->counts.p, &one, sizeof(one), cudaMemcpyHostToDevice, st));
	// MeshProcessor<4>::init (MeshProcessor.cpp:25-55, 98-128): adj_next = uses, adj_offset = their exclusive prefix, lists in (prim, corner) order
	BMF_CUDA(cudaMemsetAsync(ctx->valence.p, 0, V + 16, st));
	BMF_LAUNCH(k_valence_from_inds, grid_for(I, CTA), CTA, 0, ctx->inds.p, I, ctx->valence.p);
	BMF_LAUNCH(k_scan8_partial, nblk, CTA, 0, ctx->valence.p, V, ctx->block_sums.p);
	BMF_LAUNCH(k_scan_block_sums, 1, SCAN_CTA, 0, ctx->block_sums.p, (int)nblk);
	BMF_LAUNCH(k_scan8_final, nblk, CTA, 0, ctx->valence.p, V, ctx->block_sums.p, ctx->adj_off.p);
	BMF_CUDA(cudaMemsetAsync(ctx->cursor.p, 0, V * sizeof(uint32_t), st));
	BMF_LAUNCH(k_csr_fill<4>, grid_for(Q, CTA), CTA, 0, ctx->inds.p, Q, ctx->counts.p, 1, ctx->adj_off.p, ctx->cursor.p, ctx->adj.p, ctx->prim_vbase.p);
	BMF_LAUNCH(k_csr_sort, grid_for(V, CTA), CTA, 0, ctx->adj_off.p, ctx->valence.p, V, ctx->adj.p);
	BMF_LAUNCH(k_collapse_bad_quads, 1, COLLAPSE_CTA, 0, ctx->inds.p, (uint32_t)Q, ctx->pos.p, ctx->valence.p, ctx->adj_off.p, ctx->adj.p, (uint32_t)I, ctx->boundary.p,
	           ctx->cls.p, ctx->totals_dev.p);
	unsigned long long res[2] = { 0, 0 };
	BMF_CUDA(cudaMemcpyAsync(res, ctx->totals_dev.p, sizeof(res), cudaMemcpyDeviceToHost, st));
	BMF_CUDA(cudaMemcpyAsync(pos, ctx->pos.p, sizeof(float) * 3 * V, cudaMemcpyDeviceToHost, st));
	BMF_CUDA(cudaMemcpyAsync(quads, ctx->inds.p, sizeof(uint32_t) * I, cudaMemcpyDeviceToHost, st));
	if (destroyed) BMF_CUDA(cudaMemcpyAsync(destroyed, ctx->boundary.p, Q, cudaMemcpyDeviceToHost, st));
	if (adj_next) BMF_CUDA(cudaMemcpyAsync(adj_next, ctx->valence.p, V, cudaMemcpyDeviceToHost, st));
	BMF_CUDA(cudaStreamSynchronize(st));
	if (flushed && res[1]) 
	{
		BMF_CUDA(cudaMemcpyAsync(flushed, ctx->cls.p, sizeof(uint32_t) * 4 * (size_t)res[1], cudaMemcpyDeviceToHost, st));
		BMF_CUDA(cudaStreamSynchronize(st));
	}
	if (bad_count) *bad_count = (int64_t)res[0];
	if (n_flushed) *n_flushed = (int64_t)res[1];
	return BMF_OK;
}

int bmf_quads_to_tris(bmf_ctx* ctx, const uint32_t* quads, int64_t n_quads, uint32_t* tris)
{
	if (!ctx || n_quads < 0 || (n_quads && (!quads || !tris))) return fail(ctx, BMF_ERR_INVALID, "bmf_quads_to_tris: bad arguments");
	if (n_quads == 0) return BMF_OK;
	BMF_CUDA(cudaSetDevice(ctx->device));
	// scratch: the adjacency / dual-point arenas of the smoothing stage are free outside a submit
	BMF_CUDA(ctx->adj.reserve(4 * (size_t)n_quads));
	BMF_CUDA(ctx->cursor.reserve(6 * (size_t)n_quads));
	cudaStream_t st = ctx->stream;
	BMF_CUDA(cudaMemcpyAsync(ctx->adj.p, quads, sizeof(uint32_t) * 4 * (size_t)n_quads, cudaMemcpyHostToDevice, st));
	BMF_LAUNCH(k_quads_to_tris, std::min(grid_for((size_t)n_quads, CTA), (unsigned)(ctx->sm_count * 8)), CTA, 0, ctx->adj.p, (size_t)n_quads, ctx->cursor.p);
	BMF_CUDA(cudaMemcpyAsync(tris, ctx->cursor.p, sizeof(uint32_t) * 6 * (size_t)n_quads, cudaMemcpyDeviceToHost, st));
	BMF_CUDA(cudaStreamSynchronize(st));
	return BMF_OK;
}

int bmf_batch_download_flat_quads(bmf_ctx* ctx, int smooth_normals, float* p_data, float* n_data, float* c_data)
{
	if (!ctx) return BMF_ERR_INVALID;
	if (!ctx->have_batch) return fail(ctx, BMF_ERR_STATE, "bmf_batch_download_flat_quads: no batch submitted");
	if (!ctx->params.quads) return fail(ctx, BMF_ERR_STATE, "bmf_batch_download_flat_quads: the resident batch is not a quad batch");
	int rc = finish(ctx);
	if (rc) return rc;
	const size_t I = (size_t)ctx->totals[2], Q = I / 4;
	if (Q == 0) return BMF_OK;
	BMF_CUDA(cudaSetDevice(ctx->device));
	BMF_CUDA(ctx->dp.reserve(3 * I));
	BMF_CUDA(ctx->dn.reserve(3 * I));
	BMF_CUDA(ctx->dc.reserve(3 * I));
	cudaStream_t st = ctx->stream;
	BMF_LAUNCH(k_format_unwind, std::min(grid_for(Q, CTA), (unsigned)(ctx->sm_count * 8)), CTA, 0, ctx->pos.p, ctx->normal.p, ctx->color.p, ctx->inds.p, Q, ctx->counts.p,
	           ctx->n, smooth_normals ? 1 : 0, ctx->dp.p, ctx->dn.p, ctx->dc.p);
	if (p_data) BMF_CUDA(cudaMemcpyAsync(p_data, ctx->dp.p, sizeof(float) * 3 * I, cudaMemcpyDeviceToHost, st));
	if (n_data) BMF_CUDA(cudaMemcpyAsync(n_data, ctx->dn.p, sizeof(float) * 3 * I, cudaMemcpyDeviceToHost, st));
	if (c_data) BMF_CUDA(cudaMemcpyAsync(c_data, ctx->dc.p, sizeof(float) * 3 * I, cudaMemcpyDeviceToHost, st));
	BMF_CUDA(cudaStreamSynchronize(st));
	return BMF_OK;
}

int bmf_ubench_issue(bmf_ctx* ctx, float* gops /* [4] */)
{
	if (!ctx || !gops) return BMF_ERR_INVALID;
	BMF_CUDA(cudaSetDevice(ctx->device));
	BMF_CUDA(ctx->qe.reserve(4));
	cudaEvent_t a, b;
	BMF_CUDA(cudaEventCreate(&a));
	BMF_CUDA(cudaEventCreate(&b));
	const int iters = 4096;
	const unsigned grid = (unsigned)(ctx->sm_count * 8 * 4);
	for (int op = 0; op < 4; op++)
	{
		float best = 0.0f;
		for (int rep = 0; rep < 4; rep++)
		{
			cudaEventRecord(a, ctx->stream);
			if (op == 0) BMF_LAUNCH(k_ubench_issue<0>, grid, CTA, 0, iters, 1.0001f, 12345, ctx->qe.p);
			if (op == 1) BMF_LAUNCH(k_ubench_issue<1>, grid, CTA, 0, iters, 1.0001f, 12345, ctx->qe.p);
			if (op == 2) BMF_LAUNCH(k_ubench_issue<2>, grid, CTA, 0, iters, 1.0001f, 12345, ctx->qe.p);
			if (op == 3) BMF_LAUNCH(k_ubench_issue<3>, grid, CTA, 0, iters, 1.0001f, 12345, ctx->qe.p);
			cudaEventRecord(b, ctx->stream);
			BMF_CUDA(cudaEventSynchronize(b));
			float ms = 0.0f;
			cudaEventElapsedTime(&ms, a, b);
			// thread-level operations per second: OP 2 is two instructions (LOP3 + IADD3/SHF...) per chain step as written; reported per
			// chain step, the harness converts with the SASS instruction count per step
			const double ops = (double)grid * CTA * (double)iters * 8.0 * (op == 3 ? 2.0 : 1.0);
			const float g = (float)(ops / (ms * 1e-3) / 1e9);
			if (rep > 0 && g > best) best = g;
		}
		gops[op] = best;
	}
	cudaEventDestroy(a);
	cudaEventDestroy(b);
	return BMF_OK;
}

float bmf_seam_overlap(int dim) { return dim > 0 ? -0.5f / (float)dim : 0.0f; }

int bmf_batch_stitch(bmf_ctx* ctx, const int32_t* group, int cross_group_only, int64_t* n_tris)
{
	if (!ctx) return BMF_ERR_INVALID;
	if (!ctx->have_batch) return fail(ctx, BMF_ERR_STATE, "bmf_batch_stitch: no batch submitted");
	if (cross_group_only && !group) return fail(ctx, BMF_ERR_INVALID, "bmf_batch_stitch: cross_group_only needs group ids");
	int rc = finish(ctx);
	if (rc) return rc;
	const int n = ctx->n;
	const Layout L = ctx->L;
	const int d = L.d;

	// the chunk lattice: one slot = the extent of the finest chunk; every chunk must be a power-of-two number of slots
	// wide and sit at a multiple of its own extent (leaves of one octree do: WorldOctree.cpp:175-210)
	// The lattice is anchored to the OCTREE, not to the batch: a sub-range of an octree's leaves (one GPU's share, the border
	// chunks of the cross-rank pass) may have its minimum corner set by a fine chunk that does not sit on a multiple of the
	// coarser chunks' extent.  Every leaf sits at root + k * size, so the largest chunk of the batch is congruent to the root
	// modulo its own size, and stepping down from it in whole multiples of that size to (or below) the minimum corner gives an
	// origin every chunk of a valid leaf set is aligned to.
	float smin = ctx->descs[0].size, smax = ctx->descs[0].size;
	double lo[3] = { ctx->descs[0].pos[0], ctx->descs[0].pos[1], ctx->descs[0].pos[2] }, o[3];
	int largest = 0;
	for (int i = 0; i < n; i++)
	{
		const bmf_chunk_desc& c = ctx->descs[i];
		if (!(c.size > 0.0f)) return fail(ctx, BMF_ERR_INVALID, "bmf_batch_stitch: chunk size must be positive");
		smin = std::min(smin, c.size);
		if (c.size > smax) { smax = c.size; largest = i; }
		for (int a = 0; a < 3; a++) lo[a] = std::min(lo[a], (double)c.pos[a]);
	}
	for (int a = 0; a < 3; a++)
	{
		const double pl = (double)ctx->descs[largest].pos[a];
		o[a] = pl - std::ceil((pl - lo[a]) / (double)smax - 1e-6) * (double)smax;
	}
	std::vector<SeamChunk> sc(n);
	int g[3] = { 0, 0, 0 };
	for (int i = 0; i < n; i++)
	{
		const bmf_chunk_desc& c = ctx->descs[i];
		const double e = (double)c.size / (double)smin;
		const long long ei = llround(e);
		if (ei < 1 || ei > (1 << 20) || std::fabs(e - (double)ei) > 1e-3 * e || (ei & (ei - 1)))
			return fail(ctx, BMF_ERR_INVALID, "bmf_batch_stitch: chunk sizes must be power-of-two multiples of the smallest one");
		int org[3];
		for (int a = 0; a < 3; a++)
		{
			const double q = ((double)c.pos[a] - o[a]) / (double)smin;
			const long long qi = llround(q);
			if (std::fabs(q - (double)qi) > 1e-3 || qi < 0 || qi > (1 << 20) || (qi % ei))
				return fail(ctx, BMF_ERR_INVALID, "bmf_batch_stitch: chunks must be aligned leaves of one octree");
			org[a] = (int)qi;
			g[a] = std::max(g[a], (int)(qi + ei));
		}
		sc[i].ox = org[0]; sc[i].oy = org[1]; sc[i].oz = org[2];
		sc[i].lg = ilog2((int)ei);
	}
	const size_t slots = (size_t)g[0] * g[1] * g[2];
	if (slots > ((size_t)1 << 24)) return fail(ctx, BMF_ERR_INVALID, "bmf_batch_stitch: more than 2^24 finest-chunk slots");
	std::vector<int32_t> map(slots, -1);
	for (int i = 0; i < n; i++)
	{
		const int e = 1 << sc[i].lg;
		for (int x = sc[i].ox; x < sc[i].ox + e; x++)
			for (int y = sc[i].oy; y < sc[i].oy + e; y++)
				for (int z = sc[i].oz; z < sc[i].oz + e; z++)
				{
					int32_t& m = map[((size_t)x * g[1] + y) * g[2] + z];
					if (m >= 0) return fail(ctx, BMF_ERR_INVALID, "bmf_batch_stitch: chunks overlap");
					m = i;
				}
	}

	BMF_CUDA(cudaSetDevice(ctx->device));
	cudaStream_t st = ctx->stream;
	SeamArgs A;
	A.G.gx = g[0]; A.G.gy = g[1]; A.G.gz = g[2];
	A.G.n = n;
	A.G.npts = 6 * d * d + 2;
	A.G.bpc = (A.G.npts + CTA - 1) / CTA;
	if ((size_t)n * A.G.bpc > 0x7FFFFFFFull) return fail(ctx, BMF_ERR_INVALID, "bmf_batch_stitch: batch too large");
	const unsigned nblk = (unsigned)((size_t)n * A.G.bpc);
	BMF_CUDA(ctx->seam_chunks.reserve(n));
	BMF_CUDA(ctx->seam_map.reserve(slots));
	BMF_CUDA(ctx->seam_blk.reserve((size_t)n * SEAM_PARTS)); // triangles per (chunk, part)
	BMF_CUDA(ctx->seam_cnt.reserve(n));
	BMF_CUDA(ctx->seam_clean.reserve(n));
	BMF_CUDA(ctx->seam_layers.reserve(n));
	BMF_CUDA(ctx->seam_active.reserve(n));
	BMF_CUDA(ctx->seam_counters.reserve(2));
	const int act_words = (A.G.npts + 31) / 32 + 1;
	BMF_CUDA(ctx->seam_act.reserve((size_t)n * act_words));
	BMF_CUDA(ctx->seam_base.reserve((size_t)n + 1));
	if (group) BMF_CUDA(ctx->seam_group.reserve(n));
	if (!ctx->seam_total_pinned) BMF_CUDA(cudaMallocHost((void**)&ctx->seam_total_pinned, sizeof(unsigned long long)));
	for (cudaEvent_t& e : ctx->seam_ev)
		if (!e) BMF_CUDA(cudaEventCreate(&e));
	// pageable sources: these copies are staged by the runtime before the call returns
	BMF_CUDA(cudaMemcpyAsync(ctx->seam_chunks.p, sc.data(), sizeof(SeamChunk) * n, cudaMemcpyHostToDevice, st));
	BMF_CUDA(cudaMemcpyAsync(ctx->seam_map.p, map.data(), sizeof(int32_t) * slots, cudaMemcpyHostToDevice, st));
	if (group) BMF_CUDA(cudaMemcpyAsync(ctx->seam_group.p, group, sizeof(int32_t) * n, cudaMemcpyHostToDevice, st));
	BMF_CUDA(cudaMemsetAsync(ctx->seam_cnt.p, 0, sizeof(uint32_t) * n, st));
	BMF_CUDA(cudaMemsetAsync(ctx->seam_blk.p, 0, sizeof(uint32_t) * (size_t)n * SEAM_PARTS, st));
	BMF_CUDA(cudaMemsetAsync(ctx->seam_counters.p, 0, sizeof(uint32_t) * 2, st));
	BMF_CUDA(cudaMemsetAsync(ctx->seam_act.p, 0, sizeof(uint32_t) * (size_t)n * act_words, st));
	A.L = L;
	A.chunks = ctx->seam_chunks.p;
	A.slot_map = ctx->seam_map.p;
	A.bits = ctx->bits.p;
	A.uni = ctx->uni_valid ? ctx->uni.p : nullptr;
	A.clean = ctx->seam_clean.p;
	A.act = ctx->seam_act.p;
	A.act_words = act_words;
	A.group = group ? ctx->seam_group.p : nullptr;
	A.cross_group_only = cross_group_only ? 1 : 0;
	A.geom = ctx->geom.p;
	A.s = ctx->sampler;
	A.src.density = ctx->density_cur;
	A.src.hmap = (!ctx->density_cur && is_terrain2d(ctx->sampler.kind)) ? ctx->hmap.p : nullptr;
	A.src.sheet_of = ctx->sheet_of.p;

	BMF_CUDA(cudaEventRecord(ctx->seam_ev[0], st));
	const unsigned pgrid = std::min(nblk, (unsigned)(ctx->sm_count * 8));
	BMF_LAUNCH(k_seam_layers, (unsigned)n, CTA, 0, L, ctx->bits.p, ctx->flags.p, n, ctx->seam_layers.p);
	BMF_LAUNCH(k_seam_cull, grid_for((size_t)n * 32, CTA), CTA, 0, A.G, ctx->seam_chunks.p, ctx->seam_map.p, ctx->flags.p, ctx->seam_layers.p, ctx->seam_clean.p,
	           ctx->seam_active.p, ctx->seam_counters.p);
	BMF_LAUNCH(k_seam_classify, std::min((unsigned)n * 6u, (unsigned)(ctx->sm_count * 8)), CTA, 0, A, ctx->seam_active.p, ctx->seam_counters.p, ctx->seam_act.p);
	BMF_LAUNCH(k_seam_pass<false>, pgrid, CTA, 0, A, ctx->seam_active.p, ctx->seam_counters.p, ctx->seam_blk.p, ctx->seam_cnt.p, ctx->seam_base.p, nullptr);
	BMF_LAUNCH(k_seam_scan, 1, SEAM_SCAN_CTA, 0, ctx->seam_cnt.p, n, ctx->seam_base.p, ctx->seam_base.p + n);
	BMF_CUDA(cudaEventRecord(ctx->seam_ev[1], st));
	BMF_CUDA(cudaMemcpyAsync(ctx->seam_total_pinned, ctx->seam_base.p + n, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
	BMF_CUDA(cudaStreamSynchronize(st));
	const unsigned long long T = *ctx->seam_total_pinned;
	if (T > 0x7FFFFFFFull) return fail(ctx, BMF_ERR_INVALID, "bmf_batch_stitch: more than 2^31 seam triangles; split the batch");
	if (T)
	{
		BMF_CUDA(ctx->seam_tris.reserve(9 * (size_t)T));
		BMF_LAUNCH(k_seam_pass<true>, pgrid, CTA, 0, A, ctx->seam_active.p, ctx->seam_counters.p, ctx->seam_blk.p, ctx->seam_cnt.p, ctx->seam_base.p, ctx->seam_tris.p);
	}
	BMF_CUDA(cudaEventRecord(ctx->seam_ev[2], st));
	BMF_CUDA(cudaStreamSynchronize(st));
	cudaEventElapsedTime(&ctx->seam_ms[0], ctx->seam_ev[0], ctx->seam_ev[1]);
	cudaEventElapsedTime(&ctx->seam_ms[1], ctx->seam_ev[1], ctx->seam_ev[2]);
	ctx->seam_n_tris = (int64_t)T;
	if (n_tris) *n_tris = (int64_t)T;
	return BMF_OK;
}

int bmf_seam_download(bmf_ctx* ctx, float* positions)
{
	if (!ctx) return BMF_ERR_INVALID;
	if (ctx->seam_n_tris < 0) return fail(ctx, BMF_ERR_STATE, "bmf_seam_download: no seam pass run on the resident batch");
	if (!positions || ctx->seam_n_tris == 0) return BMF_OK;
	BMF_CUDA(cudaSetDevice(ctx->device));
	BMF_CUDA(cudaMemcpyAsync(positions, ctx->seam_tris.p, sizeof(float) * 9 * (size_t)ctx->seam_n_tris, cudaMemcpyDeviceToHost, ctx->stream));
	BMF_CUDA(cudaStreamSynchronize(ctx->stream));
	return BMF_OK;
}

int bmf_seam_stage_ms(bmf_ctx* ctx, float* ms)
{
	if (!ctx || !ms) return BMF_ERR_INVALID;
	if (ctx->seam_n_tris < 0) return fail(ctx, BMF_ERR_STATE, "bmf_seam_stage_ms: no seam pass run on the resident batch");
	ms[0] = ctx->seam_ms[0];
	ms[1] = ctx->seam_ms[1];
	return BMF_OK;
}

} // extern "C"

// extract.cuh -- K1 sampling front-ends, K2 sign pack, K3 cell-mask/count, segment scan, K4 vertex and
// index emission.  What they compute is the closed form of DMCChunk::label_grid / label_edges /
// polygonize (DMCChunk.cpp:79-166, 168-508, 514-576); how they compute it is
// B200-first:
//   * a 32-bit sign word is one __ballot_sync over 32 consecutive z (the reference's inner loop
//     DMCChunk.cpp:132-143 collapses to one instruction);
//   * cell masks are never materialised per cell on the hot path: 32 cells are classified at once with
//     word-wide logic on the four row words and their z+1 funnel shifts ("pseudo-SIMD" at warp-word width),
//     popc gives the cell / vertex counts, only ACTIVE cells touch the triangle table;
//   * the reference's serial x->y->z scan (DMCChunk.cpp:449-498) becomes count -> exclusive scan -> emit
//     over fixed-size segments of whole x-planes staged in shared memory, so vertex ids, cell order and
//     the index buffer are bit-identical to the serial order;
//   * the dense IndexesBlock (4 B/voxel) and the 124-byte DMC_Cell records are not produced at all: a
//     4-byte vertex base per 32-cell word replaces them.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "noise.cuh"
#include "mc_tables.h"

namespace bmf
{

static constexpr int CTA = 256;

struct ChunkGeom
{
	float ox, oy, oz; // DMCChunk::overlap_pos
	float delta;      // DMCChunk::scale
};

// chunk flags accumulated by the sampling / pack kernels (DMCChunk.cpp:141-162)
enum { CF_MIXED = 1, CF_ZERO = 2, CF_ONES = 4 };
__host__ __device__ __forceinline__ bool flags_contain_mesh(uint32_t f) { return (f & CF_MIXED) || ((f & CF_ZERO) && (f & CF_ONES)); }

struct SamplerDev
{
	int32_t kind; // bmf_sampler_kind
	float world_size;
	float g, nm;       // terrain: g_scale, height multiplier
	int32_t dy_half;   // terrain3d: dy * 0.5
	int32_t n_mul;     // 1: density = -dy - n*nm ; 0: -dy - n
	NoiseState ns;
	int32_t csg_op, csg_kind_a, csg_kind_b;
	float csg_ws_a, csg_ws_b;
	float csg_off_a[3], csg_off_b[3];
};

struct Layout
{
	int d;       // dim
	int zc;      // words per row = d/32
	int wp;      // words per plane = d*zc
	int wc;      // words per chunk
	int P;       // planes per segment
	int ws;      // words per segment
	int S;       // segments per chunk
	int wpt;     // words per thread in segment kernels = ws / CTA
	int lzc, ld, lwp, lwc, lS; // log2 of zc, d, wp, wc, S (all powers of two: index math is shifts and masks)
};

__host__ __device__ __forceinline__ int ilog2(int v) { int l = 0; while ((1 << l) < v) l++; return l; }

__host__ inline Layout make_layout(int d)
{
	Layout L;
	L.d = d; L.zc = d / 32; L.wp = d * L.zc; L.wc = d * L.wp;
	L.P = (d == 32) ? 32 : (d == 64) ? 8 : (d == 128) ? 2 : 1;
	L.ws = L.P * L.wp; L.S = d / L.P; L.wpt = L.ws / CTA;
	L.lzc = ilog2(L.zc); L.ld = ilog2(d); L.lwp = ilog2(L.wp); L.lwc = ilog2(L.wc); L.lS = ilog2(L.S);
	return L;
}

__constant__ uint64_t c_tri_pack[256] = BMF_TRI_PACK_INIT;
// the same table in global memory: a constant-bank read with a different index in every lane is serialised (32 replays per warp),
// so the kernels that stage the table in shared memory fill it from here with one coalesced load
__device__ const uint64_t g_tri_pack[256] = BMF_TRI_PACK_INIT;

// ---- density of one grid point for the analytic / heightmap samplers ----------------------------------
// implicit_block (ImplicitSampler.hpp:14-36): coordinate = p + (float)i * scale
// Sampler::value at a world-space point for the analytic kinds (primitive or the build-defined CSG of two primitives)
__device__ __forceinline__ float analytic_value(const SamplerDev& s, float px, float py, float pz)
{
	if (s.kind == 4)
	{
		float a = implicit_value(s.csg_kind_a, s.csg_ws_a, px - s.csg_off_a[0], py - s.csg_off_a[1], pz - s.csg_off_a[2]);
		float b = implicit_value(s.csg_kind_b, s.csg_ws_b, px - s.csg_off_b[0], py - s.csg_off_b[1], pz - s.csg_off_b[2]);
		return s.csg_op == 0 ? fmaxf(a, b) : s.csg_op == 1 ? fminf(a, b) : fminf(a, -b);
	}
	return implicit_value(s.kind, s.world_size, px, py, pz);
}

__device__ __forceinline__ float implicit_point(const SamplerDev& s, const ChunkGeom& g, int x, int y, int z)
{
	float px = g.ox + (float)x * g.delta;
	float py = g.oy + (float)y * g.delta;
	float pz = g.oz + (float)z * g.delta;
	return analytic_value(s, px, py, pz);
}

// Sampler::gradient = implicit_gradient bound to the sampler's VALUE callback (ImplicitSampler.hpp:38-49, 57; NoiseSampler.hpp:35-47,
// 76-136): six evaluations, raw differences -- not normalised, not divided by 2h.  The value callback of every noise sampler is
// NoiseSamplers::noise3d, the constant 0 (NoiseSampler.cpp:99-102), so their gradient is (0-0, 0-0, 0-0).
__device__ __forceinline__ void sampler_gradient_at(const SamplerDev& s, float px, float py, float pz, float h, float out[3])
{
	const bool analytic = s.kind >= 0 && s.kind <= 4;
	const float dxp = analytic ? analytic_value(s, px + h, py, pz) : 0.0f, dxm = analytic ? analytic_value(s, px - h, py, pz) : 0.0f;
	const float dyp = analytic ? analytic_value(s, px, py + h, pz) : 0.0f, dym = analytic ? analytic_value(s, px, py - h, pz) : 0.0f;
	const float dzp = analytic ? analytic_value(s, px, py, pz + h) : 0.0f, dzm = analytic ? analytic_value(s, px, py, pz - h) : 0.0f;
	out[0] = dxp - dxm; out[1] = dyp - dym; out[2] = dzp - dzm;
}

// terrain*_block tail (NoiseSampler.cpp:136-143, 180-187, 216-223, 250-257): density from a noise value
__device__ __forceinline__ float terrain_density(const SamplerDev& s, const ChunkGeom& g, int y, float n)
{
	float dy = ((float)y * g.delta + g.oy) * s.g;
	if (s.dy_half) dy = dy * 0.5f;
	return s.n_mul ? (-dy - n * s.nm) : (-dy - n);
}

// where the emitters read a density sample back from
struct DensitySource
{
	const float* density; // [n][d^3] or null
	const float* hmap;    // [n_sheets][d*d] noise sheets of the 2-D terrains, or null
	const int* sheet_of;  // [n] sheet index of each chunk (chunks stacked in y share one sheet)
};

__device__ __forceinline__ float density_at(const SamplerDev& s, const DensitySource& src, const ChunkGeom& g, int d, int chunk, int x, int y, int z)
{
	if (src.density) return src.density[(size_t)chunk * d * d * d + ((size_t)x * d + y) * d + z];
	if (src.hmap) return terrain_density(s, g, y, src.hmap[(size_t)src.sheet_of[chunk] * d * d + (size_t)x * d + z]);
	return implicit_point(s, g, x, y, z);
}

// block-level flag merge: one atomicOr per CTA (a per-warp "read first, add only new bits" variant was 7x slower:
// millions of same-address volatile reads serialise in L2)
__device__ __forceinline__ void merge_flags(uint32_t f, uint32_t* chunk_flags)
{
	__shared__ uint32_t s_f;
	if (threadIdx.x == 0) s_f = 0;
	__syncthreads();
	f |= __shfl_xor_sync(0xffffffffu, f, 16);
	f |= __shfl_xor_sync(0xffffffffu, f, 8);
	f |= __shfl_xor_sync(0xffffffffu, f, 4);
	f |= __shfl_xor_sync(0xffffffffu, f, 2);
	f |= __shfl_xor_sync(0xffffffffu, f, 1);
	if ((threadIdx.x & 31) == 0 && f) atomicOr(&s_f, f);
	__syncthreads();
	if (threadIdx.x == 0 && s_f) atomicOr(chunk_flags, s_f);
}

__device__ __forceinline__ uint32_t word_flags(uint32_t w) { return w == 0 ? CF_ZERO : (w == 0xFFFFFFFFu ? CF_ONES : CF_MIXED); }

// ---- K1a: implicit primitives -> (density) + sign words.  One warp per 32-z word, WORDS_PER_CTA words per CTA.
static constexpr int SAMPLE_WORDS_PER_CTA = 64;

__global__ void __launch_bounds__(CTA) k_sample_implicit(SamplerDev s, const ChunkGeom* __restrict__ geom, Layout L,
                                                          uint32_t* __restrict__ bits, float* __restrict__ density, uint32_t* __restrict__ flags)
{
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int lcpc = L.lwc - 6; // log2(CTAs per chunk), SAMPLE_WORDS_PER_CTA == 64
	const int chunk = blockIdx.x >> lcpc;
	const int w0 = (blockIdx.x & ((1 << lcpc) - 1)) * SAMPLE_WORDS_PER_CTA;
	const ChunkGeom g = geom[chunk];
	uint32_t f = 0;
	for (int k = warp; k < SAMPLE_WORDS_PER_CTA; k += CTA / 32)
	{
		int w = w0 + k;
		int zb = w & (L.zc - 1), y = (w >> L.lzc) & (L.d - 1), x = w >> L.lwp;
		int z = zb * 32 + lane;
		float v = implicit_point(s, g, x, y, z);
		if (density) density[(size_t)chunk * L.wc * 32 + (size_t)w * 32 + lane] = v;
		uint32_t word = __ballot_sync(0xffffffffu, v < 0.0f);
		if (lane == 0)
		{
			bits[(size_t)chunk * L.wc + w] = word;
			f |= word_flags(word);
		}
	}
	merge_flags(f, flags + chunk);
}

// ---- K1b: 2-D terrains.  Noise sheet: one thread per (ix, iz) column (NOISE_BLOCK with size_y = 1,
// NoiseSampler.cpp:117,152), then density = -dy - n*height per voxel and the sign word.
// The sheet depends only on (overlap_pos.x, overlap_pos.z, delta): chunks stacked in y get the same
// floats, so the host deduplicates and `geom` here holds one entry per UNIQUE sheet.
// order-preserving map float -> uint32 (for atomicMin / atomicMax on floats); NaNs are handled separately
__device__ __forceinline__ uint32_t float_order(float f)
{
	const uint32_t u = __float_as_uint(f);
	return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float float_unorder(uint32_t o) { return __uint_as_float((o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o); }

// sheet_mm[0*ns + s] = min, [1*ns + s] = max of t = n*height over sheet s (order-mapped), [2*ns + s] = 1 if any NaN
template <int BASE>
__global__ void __launch_bounds__(CTA) k_terrain2d_sheet(SamplerDev s, const ChunkGeom* __restrict__ geom, int d, int ld, float* __restrict__ hmap, int n_chunks,
                                                          uint32_t* __restrict__ sheet_mm)
{
	__shared__ uint32_t s_mm[3];
	size_t i = (size_t)blockIdx.x * CTA + threadIdx.x; // d*d is a multiple of CTA: a CTA never straddles two sheets, no thread is out of range
	int chunk = (int)(i >> (2 * ld));
	int r = (int)(i & (((size_t)1 << (2 * ld)) - 1));
	int ix = r >> ld, iz = r & (d - 1);
	const ChunkGeom g = geom[chunk];
	float sg = g.delta * s.g;
	float vx = (float)ix * sg + g.ox * s.g;
	float vy = (float)0 * sg + 0.0f;
	float vz = (float)iz * sg + g.oz * s.g;
	__shared__ float4 s_corner[CTA / 32][8]; // gradient_perturb_warp: the lattice corners a warp shares
	const float n = noise_eval_warp<BASE>(s.ns, vx, vy, vz, s_corner[threadIdx.x >> 5], threadIdx.x & 31);
	hmap[i] = n;
	// per-sheet range of t = n*height (the same product k_terrain2d_bits compares against)
	const float t = n * s.nm;
	const bool nan = t != t;
	uint32_t lo = nan ? 0xFFFFFFFFu : float_order(t), hi = nan ? 0u : float_order(t), bad = nan ? 1u : 0u;
#pragma unroll
	for (int o = 16; o >= 1; o >>= 1)
	{
		lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
		hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
		bad |= __shfl_xor_sync(0xffffffffu, bad, o);
	}
	if (threadIdx.x == 0) { s_mm[0] = 0xFFFFFFFFu; s_mm[1] = 0u; s_mm[2] = 0u; }
	__syncthreads();
	if ((threadIdx.x & 31) == 0)
	{
		atomicMin(&s_mm[0], lo);
		atomicMax(&s_mm[1], hi);
		if (bad) atomicOr(&s_mm[2], 1u);
	}
	__syncthreads();
	if (threadIdx.x == 0)
	{
		atomicMin(sheet_mm + chunk, s_mm[0]);
		atomicMax(sheet_mm + n_chunks + chunk, s_mm[1]);
		if (s_mm[2]) atomicOr(sheet_mm + 2 * n_chunks + chunk, 1u);
	}
}

// Chunks that lie entirely above or below the heightfield: with -dy monotone in y, every bit of the chunk is 1 if
// -dy(0) < min t, and 0 if !(-dy(d-1) < max t) (no NaN in the sheet).  Such a chunk contains no mesh
// (DMCChunk.cpp:159-162); its flags are set here and k_terrain2d_bits skips it, so its sign words are never written
// (bmf_batch_copy_chunk synthesises them on request).  uni: 0 = mixed, 1 = all air (ones), 2 = all solid (zeros).
__global__ void __launch_bounds__(CTA) k_terrain2d_classify(SamplerDev s, const ChunkGeom* __restrict__ geom, int d, const int* __restrict__ sheet_of,
                                                             const uint32_t* __restrict__ sheet_mm, int n_sheets, int n_chunks, uint32_t* __restrict__ flags,
                                                             uint8_t* __restrict__ uni, int* __restrict__ mixed_list, unsigned long long* __restrict__ mixed_count)
{
	const int c = blockIdx.x * CTA + threadIdx.x;
	if (c >= n_chunks) return;
	const ChunkGeom g = geom[c];
	const int sh = sheet_of[c];
	float top = ((float)0 * g.delta + g.oy) * s.g, bot = ((float)(d - 1) * g.delta + g.oy) * s.g;
	if (s.dy_half) { top = top * 0.5f; bot = bot * 0.5f; }
	top = -top; bot = -bot; // largest / smallest -dy of the chunk when -dy is non-increasing in y
	uint8_t u = 0;
	const bool monotone = g.delta >= 0.0f && s.g >= 0.0f && top >= bot; // rounding is monotone, so the chain y -> -dy(y) is too; false on NaN
	if (monotone && !sheet_mm[2 * n_sheets + sh])
	{
		const float tmin = float_unorder(sheet_mm[sh]), tmax = float_unorder(sheet_mm[n_sheets + sh]);
		if (top < tmin) u = 1;
		else if (!(bot < tmax)) u = 2;
	}
	uni[c] = u;
	if (u) flags[c] = (u == 1) ? CF_ONES : CF_ZERO;
	else if (mixed_list)
	{
		// the chunks k_terrain2d_bits has to visit (order irrelevant); one atomic per warp
		const uint32_t bal = __activemask();
		const uint32_t mine = __ballot_sync(bal, true);
		const int leader = __ffs(mine) - 1, lane = threadIdx.x & 31;
		unsigned long long o = 0;
		if (lane == leader) o = atomicAdd(mixed_count, (unsigned long long)__popc(mine));
		o = __shfl_sync(mine, o, leader);
		mixed_list[o + __popc(mine & ((1u << lane) - 1u))] = c;
	}
}

__global__ void __launch_bounds__(CTA) k_terrain2d_density(SamplerDev s, const ChunkGeom* __restrict__ geom, Layout L, const float* __restrict__ hmap,
                                                            const int* __restrict__ sheet_of, uint32_t* __restrict__ bits, float* __restrict__ density,
                                                            uint32_t* __restrict__ flags)
{
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int lcpc = L.lwc - 6; // log2(CTAs per chunk), SAMPLE_WORDS_PER_CTA == 64
	const int chunk = blockIdx.x >> lcpc;
	const int w0 = (blockIdx.x & ((1 << lcpc) - 1)) * SAMPLE_WORDS_PER_CTA;
	const ChunkGeom g = geom[chunk];
	const float* hm = hmap + ((size_t)sheet_of[chunk] << (2 * L.ld));
	uint32_t f = 0;
	for (int k = warp; k < SAMPLE_WORDS_PER_CTA; k += CTA / 32)
	{
		int w = w0 + k;
		int zb = w & (L.zc - 1), y = (w >> L.lzc) & (L.d - 1), x = w >> L.lwp;
		int z = zb * 32 + lane;
		float v = terrain_density(s, g, y, hm[x * L.d + z]);
		if (density) density[(size_t)chunk * L.wc * 32 + (size_t)w * 32 + lane] = v;
		uint32_t word = __ballot_sync(0xffffffffu, v < 0.0f);
		if (lane == 0)
		{
			bits[(size_t)chunk * L.wc + w] = word;
			f |= word_flags(word);
		}
	}
	merge_flags(f, flags + chunk);
}

// ---- K1b': sign words of a 2-D terrain WITHOUT evaluating the density per voxel.
// density(x,y,z) = (-dy(y)) - t(x,z) with t = n(x,z)*height: in IEEE arithmetic with gradual underflow
// a - b < 0  <=>  a < b (the difference of two floats is zero only if they are equal; NaN compares false
// either way), so bit(x,y,z) = (-dy(y) < t(x,z)).  A warp owns a 32(y) x 32(z) tile of one x-plane; lane j
// holds -dy(y0+j) and t(z0+j).  -dy is monotone in y (checked per tile, NaN-safe), so every lane finds the
// first y of ITS column where the bit turns on with a 5-step shuffle binary search, forms the column's 32-bit
// y-mask with one shift, and a 5-stage warp bit-matrix transpose turns the 32 column masks into the 32 row
// words -> one coalesced store.  ~70 instructions per 1024 voxels; a non-monotone tile falls back to 32 ballots.
// the word of row y = yb * 32 + lane of the (x, yb, zb) tile; `sheet` = the chunk's d x d noise sheet
__device__ __forceinline__ uint32_t terrain2d_tile_word(const SamplerDev& s, const ChunkGeom& g, const Layout& L, const float* __restrict__ sheet, int x, int yb, int zb, int lane)
{
	const float n = sheet[((size_t)x << L.ld) + zb * 32 + lane];
	const float t = n * s.nm;
	const int y = yb * 32 + lane;
	float dy = ((float)y * g.delta + g.oy) * s.g;
	if (s.dy_half) dy = dy * 0.5f;
	const float ndy = -dy;
	uint32_t mine;
	const float nxt = __shfl_down_sync(0xffffffffu, ndy, 1);
	const bool monotone = __all_sync(0xffffffffu, lane == 31 || ndy >= nxt); // false if any NaN
	const float top = __shfl_sync(0xffffffffu, ndy, 0), bot = __shfl_sync(0xffffffffu, ndy, 31); // largest / smallest -dy of the tile
	if (monotone && __all_sync(0xffffffffu, top < t))
		mine = 0xFFFFFFFFu; // the whole tile is above the surface: every row word is all air
	else if (monotone && __all_sync(0xffffffffu, !(bot < t)))
		mine = 0u; // the whole tile is below the surface (NaN columns are never air)
	else if (monotone)
	{
		// number of rows y (from the bottom of the tile) whose bit is still 0 in this lane's column
		int cnt = 0;
#pragma unroll
		for (int st = 16; st >= 1; st >>= 1)
		{
			const float a = __shfl_sync(0xffffffffu, ndy, cnt + st - 1);
			if (!(a < t)) cnt += st;
		}
		const float a31 = __shfl_sync(0xffffffffu, ndy, 31);
		if (cnt == 31 && !(a31 < t)) cnt = 32;
		uint32_t v = cnt >= 32 ? 0u : (0xFFFFFFFFu << cnt); // lane = z, bit = y
		// 32x32 bit-matrix transpose across the warp: afterwards lane = y, bit = z
#pragma unroll
		for (int j = 16; j >= 1; j >>= 1)
		{
			const uint32_t m0 = j == 16 ? 0x0000FFFFu : j == 8 ? 0x00FF00FFu : j == 4 ? 0x0F0F0F0Fu : j == 2 ? 0x33333333u : 0x55555555u;
			const uint32_t o = __shfl_xor_sync(0xffffffffu, v, j);
			v = (lane & j) ? ((v & ~m0) | ((o >> j) & m0)) : ((v & m0) | ((o << j) & ~m0));
		}
		mine = v;
	}
	else
	{
		mine = 0;
#pragma unroll
		for (int j = 0; j < 32; j++)
		{
			const float a = __shfl_sync(0xffffffffu, ndy, j);
			const uint32_t word = __ballot_sync(0xffffffffu, a < t);
			if (lane == j) mine = word;
		}
	}
	return mine;
}

__device__ __forceinline__ void terrain2d_bits_task(const SamplerDev& s, const ChunkGeom* __restrict__ geom, const Layout& L, const float* __restrict__ hmap,
                                                    const int* __restrict__ sheet_of, int chunk, int task /* (x, yb, zb) inside the chunk */, int lane,
                                                    uint32_t* __restrict__ bits, uint32_t* __restrict__ flags)
{
	const int zb = task & (L.zc - 1);
	const int yb = (task >> L.lzc) & (L.zc - 1);
	const int x = (task >> (2 * L.lzc)) & (L.d - 1);
	const ChunkGeom g = geom[chunk];
	const uint32_t mine = terrain2d_tile_word(s, g, L, hmap + ((size_t)sheet_of[chunk] << (2 * L.ld)), x, yb, zb, lane);
	const int y = yb * 32 + lane;
	bits[(size_t)chunk * L.wc + ((((size_t)x << L.ld) + y) << L.lzc) + zb] = mine;
	merge_flags(word_flags(mine), flags + chunk);
}

__global__ void __launch_bounds__(CTA) k_terrain2d_bits(SamplerDev s, const ChunkGeom* __restrict__ geom, Layout L, const float* __restrict__ hmap,
                                                         const int* __restrict__ sheet_of, const int* __restrict__ mixed_list,
                                                         const unsigned long long* __restrict__ mixed_count, uint32_t* __restrict__ bits, uint32_t* __restrict__ flags)
{
	// grid-stride over (listed chunk, CTA-sized group of tasks): the chunks that lie entirely above / below the surface are not on the
	// list, so no CTA is launched for them (an empty CTA per 8 tasks of every culled chunk used to cost more than the kernel's work)
	const int lane = threadIdx.x & 31;
	const int ltc = 2 * L.lzc + L.ld - 3;  // log2(CTA-groups per chunk): tasks per chunk = d * zc^2, CTA / 32 == 8 tasks per group
	const size_t n_groups = (size_t)*mixed_count << ltc;
	for (size_t grp = blockIdx.x; grp < n_groups; grp += gridDim.x)
		terrain2d_bits_task(s, geom, L, hmap, sheet_of, mixed_list[grp >> ltc], (int)(((grp & (((size_t)1 << ltc) - 1)) << 3) + (threadIdx.x >> 5)), lane, bits, flags);
}

// ---- K1c: 3-D terrains: one noise evaluation per voxel, density always materialised (4 B/voxel is
// noise next to ~1.5 k ALU ops/voxel) so the emitters can read crossing-edge samples back.
template <int BASE>
__global__ void __launch_bounds__(CTA) k_terrain3d(SamplerDev s, const ChunkGeom* __restrict__ geom, Layout L,
                                                    uint32_t* __restrict__ bits, float* __restrict__ density, uint32_t* __restrict__ flags)
{
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int lcpc = L.lwc - 3; // log2(CTAs per chunk), CTA/32 == 8 words per CTA
	const int chunk = blockIdx.x >> lcpc;
	const int w = (blockIdx.x & ((1 << lcpc) - 1)) * (CTA / 32) + warp;
	const ChunkGeom g = geom[chunk];
	int zb = w & (L.zc - 1), y = (w >> L.lzc) & (L.d - 1), x = w >> L.lwp;
	int z = zb * 32 + lane;
	float sg = g.delta * s.g;
	float vx = (float)x * sg + g.ox * s.g;
	float vy = (float)y * sg + g.oy * s.g;
	float vz = (float)z * sg + g.oz * s.g;
	__shared__ float4 s_corner[CTA / 32][8]; // gradient_perturb_warp: the lattice corners a warp shares
	float n = noise_eval_warp<BASE>(s.ns, vx, vy, vz, s_corner[warp], lane);
	float v = terrain_density(s, g, y, n);
	density[(size_t)chunk * L.wc * 32 + (size_t)w * 32 + lane] = v;
	uint32_t word = __ballot_sync(0xffffffffu, v < 0.0f);
	uint32_t f = 0;
	if (lane == 0)
	{
		bits[(size_t)chunk * L.wc + w] = word;
		f = word_flags(word);
	}
	merge_flags(f, flags + chunk);
}

// ---- K2: density block -> sign words (label_grid's pack, DMCChunk.cpp:118-157).  Pure streaming:
// the density of a batch is one flat array of 32-float groups, each group is one word.  Every lane issues
// PACK_UNROLL independent coalesced 4-byte loads before the first ballot so enough bytes are in flight; consecutive
// warps read consecutive kilobytes (DRAM-page friendly).  No barrier, no atomic: the chunk-flag contribution of a
// warp's PACK_UNROLL words goes to one byte of `gflags`, which k_reduce_flags ORs per chunk afterwards.
static constexpr int PACK_UNROLL = 8;

__global__ void __launch_bounds__(CTA) k_pack_density(const float* __restrict__ density, uint32_t* __restrict__ bits, uint8_t* __restrict__ gflags,
                                                       size_t n_words)
{
	const int lane = threadIdx.x & 31;
	const size_t warp_global = ((size_t)blockIdx.x * CTA + threadIdx.x) >> 5;
	const size_t w0 = warp_global * PACK_UNROLL;
	if (w0 >= n_words) return; // n_words is a multiple of PACK_UNROLL * (CTA/32)
	float v[PACK_UNROLL];
#pragma unroll
	for (int k = 0; k < PACK_UNROLL; k++) v[k] = __ldcs(density + (w0 + k) * 32 + lane);
	uint32_t f = 0, mine = 0;
#pragma unroll
	for (int k = 0; k < PACK_UNROLL; k++)
	{
		const uint32_t word = __ballot_sync(0xffffffffu, v[k] < 0.0f);
		if (lane == k) mine = word;
		f |= word_flags(word);
	}
	if (lane < PACK_UNROLL) bits[w0 + lane] = mine;
	if (lane == 0) gflags[warp_global] = (uint8_t)f;
}

// chunk flags = OR of the chunk's group flags (words_per_chunk / PACK_UNROLL bytes, a multiple of 128); one CTA per chunk
__global__ void __launch_bounds__(CTA) k_reduce_flags(const uint8_t* __restrict__ gflags, int groups_per_chunk, uint32_t* __restrict__ flags)
{
	__shared__ uint32_t s_f;
	if (threadIdx.x == 0) s_f = 0;
	__syncthreads();
	const uint32_t* g = reinterpret_cast<const uint32_t*>(gflags + (size_t)blockIdx.x * groups_per_chunk);
	uint32_t f = 0;
	for (int i = threadIdx.x; i < groups_per_chunk / 4; i += CTA) f |= g[i];
	f |= f >> 16;
	f |= f >> 8;
	f &= 0xFF;
	f |= __shfl_xor_sync(0xffffffffu, f, 16);
	f |= __shfl_xor_sync(0xffffffffu, f, 8);
	f |= __shfl_xor_sync(0xffffffffu, f, 4);
	f |= __shfl_xor_sync(0xffffffffu, f, 2);
	f |= __shfl_xor_sync(0xffffffffu, f, 1);
	if ((threadIdx.x & 31) == 0 && f) atomicOr(&s_f, f);
	__syncthreads();
	if (threadIdx.x == 0) flags[blockIdx.x] = s_f;
}

// ---- shared-memory staging of sign planes ----------------------------------------------------------------
// Copies `planes` x-planes starting at x0 of one chunk into smem, zero-filling planes at x >= d (the
// closed form's B == 0 outside the grid).
__device__ __forceinline__ void stage_planes(uint32_t* sb, const uint32_t* __restrict__ chunk_bits, const Layout& L, int x0, int planes)
{
	const int total = planes * L.wp;
	const int valid = min(planes, L.d - x0) * L.wp;
	const uint4* src = reinterpret_cast<const uint4*>(chunk_bits + (size_t)x0 * L.wp);
	uint4* dst = reinterpret_cast<uint4*>(sb);
	for (int i = threadIdx.x; i < total / 4; i += CTA)
		dst[i] = (i < valid / 4) ? src[i] : make_uint4(0, 0, 0, 0);
}

struct WordBits
{
	uint32_t A, A1, B, B1, C, C1, D, D1; // rows (x,y) (x,y+1) (x+1,y) (x+1,y+1); *1 = shifted so bit z holds sample z+1
};

// lx = plane index inside the staged window (plane lx+1 must be staged too)
__device__ __forceinline__ WordBits load_word_bits(const uint32_t* sb, const Layout& L, int lx, int y, int zb)
{
	WordBits r;
	const int base = (((lx << L.ld) + y) << L.lzc) + zb;
	const bool zn = zb + 1 < L.zc, yn = y + 1 < L.d;
	r.A = sb[base];
	r.A1 = __funnelshift_r(r.A, zn ? sb[base + 1] : 0u, 1);
	r.B = yn ? sb[base + L.zc] : 0u;
	r.B1 = __funnelshift_r(r.B, (yn && zn) ? sb[base + L.zc + 1] : 0u, 1);
	r.C = sb[base + L.wp];
	r.C1 = __funnelshift_r(r.C, zn ? sb[base + L.wp + 1] : 0u, 1);
	r.D = yn ? sb[base + L.wp + L.zc] : 0u;
	r.D1 = __funnelshift_r(r.D, (yn && zn) ? sb[base + L.wp + L.zc + 1] : 0u, 1);
	return r;
}

struct WordClass
{
	uint32_t active;   // cells with mask8 not in {0,255}
	uint32_t ex, ey, ez; // cells owning an X / Y / Z edge vertex
	uint32_t interior; // cells that polygonize (x,y,z < d-1)
};

__device__ __forceinline__ WordClass classify(const WordBits& b, const Layout& L, int x, int y, int zb)
{
	WordClass c;
	const uint32_t any = b.A | b.A1 | b.B | b.B1 | b.C | b.C1 | b.D | b.D1;
	const uint32_t all = b.A & b.A1 & b.B & b.B1 & b.C & b.C1 & b.D & b.D1;
	const uint32_t zvalid = (zb == L.zc - 1) ? 0x7FFFFFFFu : 0xFFFFFFFFu;
	const bool xn = x + 1 < L.d, yn = y + 1 < L.d;
	c.active = any & ~all;
	c.ex = xn ? (b.A ^ b.C) : 0u;
	c.ey = yn ? (b.A ^ b.B) : 0u;
	c.ez = (b.A ^ b.A1) & zvalid;
	c.interior = (xn && yn) ? zvalid : 0u;
	return c;
}

// corner mask of the cell at `bit`: bit (4 dx + 2 dy + dz).  The (z, z + 1) samples of a row are two consecutive bits of the row
// extended by its successor's first bit (= bit 31 of the shifted copy), so one funnel shift per row fetches both.
__device__ __forceinline__ uint32_t mask8_of(const WordBits& b, int bit)
{
	const uint32_t a = __funnelshift_r(b.A, b.A1 >> 31, bit) & 3u, bb = __funnelshift_r(b.B, b.B1 >> 31, bit) & 3u;
	const uint32_t c = __funnelshift_r(b.C, b.C1 >> 31, bit) & 3u, dd = __funnelshift_r(b.D, b.D1 >> 31, bit) & 3u;
	return a | (bb << 2) | (c << 4) | (dd << 6);
}

// packed per-word counts: cells [0,8) verts [8,16) indices [16,32)
__device__ __forceinline__ uint32_t pack_counts(uint32_t nc, uint32_t nv, uint32_t ni) { return nc | (nv << 8) | (ni << 16); }

// exclusive scan of K counters over the CTA (thread order); totals returned through tot[K] (valid in all threads)
template <int K>
__device__ __forceinline__ void block_scan(uint32_t (&v)[K], uint32_t (&tot)[K])
{
	__shared__ uint32_t s_w[K][CTA / 32];
	__shared__ uint32_t s_t[K];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	uint32_t in[K];
#pragma unroll
	for (int q = 0; q < K; q++) in[q] = v[q];
#pragma unroll
	for (int o = 1; o < 32; o <<= 1)
	{
#pragma unroll
		for (int q = 0; q < K; q++)
		{
			const uint32_t t = __shfl_up_sync(0xffffffffu, in[q], o);
			if (lane >= o) in[q] += t;
		}
	}
	if (lane == 31)
	{
#pragma unroll
		for (int q = 0; q < K; q++) s_w[q][warp] = in[q];
	}
	__syncthreads();
	if (warp == 0)
	{
#pragma unroll
		for (int q = 0; q < K; q++)
		{
			const uint32_t w = lane < CTA / 32 ? s_w[q][lane] : 0;
			uint32_t j = w;
#pragma unroll
			for (int o = 1; o < CTA / 32; o <<= 1)
			{
				const uint32_t t = __shfl_up_sync(0xffffffffu, j, o);
				if (lane >= o) j += t;
			}
			if (lane < CTA / 32) s_w[q][lane] = j - w;
			if (lane == CTA / 32 - 1) s_t[q] = j;
		}
	}
	__syncthreads();
#pragma unroll
	for (int q = 0; q < K; q++)
	{
		v[q] = in[q] - v[q] + s_w[q][warp];
		tot[q] = s_t[q];
	}
	__syncthreads();
}

__device__ __forceinline__ void block_scan3(uint32_t& a, uint32_t& b, uint32_t& c, uint32_t tot[3])
{
	uint32_t v[3] = { a, b, c }, t[3];
	block_scan<3>(v, t);
	a = v[0]; b = v[1]; c = v[2];
	tot[0] = t[0]; tot[1] = t[1]; tot[2] = t[2];
}

// ---- K3: cell-mask build + counts.  One CTA per segment (P whole x-planes); each thread owns wpt
// consecutive words.  Writes one packed count per word and one (cells, verts, indices) total per segment;
// optionally the 8-bit mask image (MasksBlock, DMCChunk.cpp:184-438) when the caller wants it back.
template <int WPT>
__global__ void __launch_bounds__(CTA) k_count(const uint32_t* __restrict__ bits, const uint32_t* __restrict__ flags, Layout L, uint32_t* __restrict__ wcnt,
                                                uint32_t* __restrict__ seg_tot, uint32_t* __restrict__ chunk_tot, uint8_t* __restrict__ masks)
{
	extern __shared__ uint32_t sb[];
	const int seg = blockIdx.x;
	const int chunk = seg >> L.lS, x0 = (seg & (L.S - 1)) * L.P;
	if (!flags_contain_mesh(flags[chunk]))
	{
		// label_edges returns at once for a chunk without a mesh (DMCChunk.cpp:170-171): nothing downstream reads its words
		if (threadIdx.x < 3) seg_tot[3 * (size_t)seg + threadIdx.x] = 0;
		return;
	}
	stage_planes(sb, bits + (size_t)chunk * L.wc, L, x0, L.P + 1);
	__syncthreads();

	uint32_t tc = 0, tv = 0, ti = 0;
	uint32_t packed[WPT];
#pragma unroll
	for (int k = 0; k < WPT; k++)
	{
		const int lw = threadIdx.x * WPT + k; // word inside the segment
		const int zb = lw & (L.zc - 1), y = (lw >> L.lzc) & (L.d - 1), lx = lw >> L.lwp;
		const WordBits b = load_word_bits(sb, L, lx, y, zb);
		const WordClass c = classify(b, L, x0 + lx, y, zb);
		uint32_t nc = __popc(c.active);
		uint32_t nv = __popc(c.ex) + __popc(c.ey) + __popc(c.ez);
		uint32_t ni = 0;
		uint32_t m = c.active & c.interior;
		while (m)
		{
			int bit = __ffs(m) - 1;
			m &= m - 1;
			ni += (uint32_t)(c_tri_pack[mask8_of(b, bit)] >> 60);
		}
		packed[k] = pack_counts(nc, nv, ni);
		tc += nc; tv += nv; ti += ni;
		if (masks)
		{
			// MasksBlock byte image: 32 cells of this word, little-endian, 8 cells per uint64
			uint8_t* mrow = masks + (size_t)chunk * L.wc * 32 + ((size_t)(x0 + lx) * L.d + y) * L.d + zb * 32;
#pragma unroll
			for (int q = 0; q < 8; q++)
			{
				uint32_t v4 = mask8_of(b, 4 * q) | (mask8_of(b, 4 * q + 1) << 8) | (mask8_of(b, 4 * q + 2) << 16) | (mask8_of(b, 4 * q + 3) << 24);
				reinterpret_cast<uint32_t*>(mrow)[q] = v4;
			}
		}
	}
	uint32_t* out = wcnt + (size_t)seg * L.ws + threadIdx.x * WPT;
	if (WPT == 4) *reinterpret_cast<uint4*>(out) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
	else
	{
		*reinterpret_cast<uint4*>(out) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
		*reinterpret_cast<uint4*>(out + 4) = make_uint4(packed[4 % WPT], packed[5 % WPT], packed[6 % WPT], packed[7 % WPT]);
	}
	// CTA totals
	uint32_t tot[3];
	block_scan3(tc, tv, ti, tot);
	if (threadIdx.x < 3)
	{
		seg_tot[3 * (size_t)seg + threadIdx.x] = tot[threadIdx.x];
		if (tot[threadIdx.x]) atomicAdd(chunk_tot + 3 * (size_t)chunk + threadIdx.x, tot[threadIdx.x]); // integer adds commute: deterministic
	}
}

// ---- chunk scan: one CTA walks the per-chunk totals (accumulated by k_count) in coalesced tiles of SCAN_CTA
// with a running carry and writes the chunk table.  A chunk that does not contain a mesh (DMCChunk.cpp:159-162,
// label_edges :170-171) has zero totals because k_count skipped it.  Segment bases inside a chunk are a <= S-term
// sum that k_bases forms itself, so the scan length is the number of chunks, not of segments.
struct ChunkCounts
{
	uint32_t contains_mesh;
	uint32_t n_cells, n_verts, n_inds;
	uint64_t cell_base, vert_base, ind_base;
};

static constexpr int SCAN_CTA = 1024;

static constexpr int SCAN_PER_THREAD = 4; // consecutive chunks per thread: 4096 chunks per tile, so the usual batch is one tile (one block scan, not four)

static constexpr int SIZE_CLASSES = 32; // of the work list below: class = 31 - min(31, n_verts / 256), so class 0 holds the largest chunks

__global__ void __launch_bounds__(SCAN_CTA) k_scan_chunks(const uint32_t* __restrict__ chunk_tot, const uint32_t* __restrict__ flags, int n_chunks,
                                                           ChunkCounts* __restrict__ chunks,
                                                           unsigned long long* __restrict__ totals /* cells, verts, inds, overflow, list counters */,
                                                           int* __restrict__ work_list /* or null: the chunks that have vertices, LARGEST FIRST (by size
                                                           class) -- the order the per-chunk kernels take them in, so that the tail of their launches is short */,
                                                           unsigned long long cap_cells, unsigned long long cap_verts, unsigned long long cap_inds /* k_check_caps' verdict
                                                           for the emitters that follow, formed here so that a submission has one launch less */)
{
	__shared__ uint32_t s_cls_cnt[SIZE_CLASSES], s_cls_pos[SIZE_CLASSES];
	if (threadIdx.x < SIZE_CLASSES) s_cls_cnt[threadIdx.x] = 0;
	// 64-bit throughout: a tile of dim-256 chunks can hold more than 2^32 indices (15 * 256^3 per chunk), and a sum that
	// wrapped inside the tile would slip past the > 2^32 guard below
	typedef unsigned long long u64;
	constexpr int REC = (int)(sizeof(ChunkCounts) / sizeof(uint32_t));
	extern __shared__ __align__(16) uint32_t s_tab[]; // [SCAN_CTA * SCAN_PER_THREAD][REC]: the tile's table records, written out coalesced (one SM storing
	                                                  // 40-byte records straight from its threads touches 32 sectors per instruction and took 20 us per 4096 chunks)
	__shared__ u64 s_w[3][SCAN_CTA / 32];
	__shared__ u64 s_tot[3];
	__shared__ uint32_t s_maxv;
	const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
	if (t == 0) s_maxv = 0;
	__syncthreads(); // the class counters are zero before the first chunk is counted
	uint32_t maxv = 0;
	u64 carry0 = 0, carry1 = 0, carry2 = 0;
	for (int base = 0; base < n_chunks; base += SCAN_CTA * SCAN_PER_THREAD)
	{
		uint32_t a[SCAN_PER_THREAD], b[SCAN_PER_THREAD], c[SCAN_PER_THREAD], f[SCAN_PER_THREAD];
		u64 ia = 0, ib = 0, ic = 0;
#pragma unroll
		for (int k = 0; k < SCAN_PER_THREAD; k++)
		{
			const int i = base + t * SCAN_PER_THREAD + k;
			a[k] = b[k] = c[k] = f[k] = 0;
			if (i < n_chunks)
			{
				f[k] = flags_contain_mesh(flags[i]) ? 1u : 0u;
				if (f[k]) { a[k] = chunk_tot[3 * (size_t)i]; b[k] = chunk_tot[3 * (size_t)i + 1]; c[k] = chunk_tot[3 * (size_t)i + 2]; }
			}
			maxv = max(maxv, b[k]);
			ia += a[k]; ib += b[k]; ic += c[k];
			if (work_list && b[k]) atomicAdd(&s_cls_cnt[SIZE_CLASSES - 1 - min(SIZE_CLASSES - 1, (int)(b[k] >> 8))], 1u);
		}
		const u64 ta0 = ia, tb0 = ib, tc0 = ic; // this thread's totals
#pragma unroll
		for (int o = 1; o < 32; o <<= 1)
		{
			u64 ta = __shfl_up_sync(0xffffffffu, ia, o), tb = __shfl_up_sync(0xffffffffu, ib, o), tc = __shfl_up_sync(0xffffffffu, ic, o);
			if (lane >= o) { ia += ta; ib += tb; ic += tc; }
		}
		if (lane == 31) { s_w[0][warp] = ia; s_w[1][warp] = ib; s_w[2][warp] = ic; }
		__syncthreads();
		if (warp == 0)
		{
			u64 va = s_w[0][lane], vb = s_w[1][lane], vc = s_w[2][lane];
			u64 ja = va, jb = vb, jc = vc;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1)
			{
				u64 ta = __shfl_up_sync(0xffffffffu, ja, o), tb = __shfl_up_sync(0xffffffffu, jb, o), tc = __shfl_up_sync(0xffffffffu, jc, o);
				if (lane >= o) { ja += ta; jb += tb; jc += tc; }
			}
			s_w[0][lane] = ja - va; s_w[1][lane] = jb - vb; s_w[2][lane] = jc - vc;
			if (lane == 31) { s_tot[0] = ja; s_tot[1] = jb; s_tot[2] = jc; }
		}
		__syncthreads();
		u64 r0 = carry0 + (ia - ta0 + s_w[0][warp]), r1 = carry1 + (ib - tb0 + s_w[1][warp]), r2 = carry2 + (ic - tc0 + s_w[2][warp]);
#pragma unroll
		for (int k = 0; k < SCAN_PER_THREAD; k++)
		{
			const int i = base + t * SCAN_PER_THREAD + k;
			if (i < n_chunks)
			{
				ChunkCounts cc;
				cc.contains_mesh = f[k];
				cc.n_cells = a[k]; cc.n_verts = b[k]; cc.n_inds = c[k];
				cc.cell_base = r0; cc.vert_base = r1; cc.ind_base = r2;
				*reinterpret_cast<ChunkCounts*>(s_tab + (size_t)(t * SCAN_PER_THREAD + k) * REC) = cc;
			}
			r0 += a[k]; r1 += b[k]; r2 += c[k];
		}
		carry0 += s_tot[0]; carry1 += s_tot[1]; carry2 += s_tot[2];
		__syncthreads();
		{
			const int n_tile = min(n_chunks - base, SCAN_CTA * SCAN_PER_THREAD);
			uint32_t* out = reinterpret_cast<uint32_t*>(chunks + base);
			for (int w = t; w < n_tile * REC; w += SCAN_CTA) out[w] = s_tab[w];
		}
		__syncthreads();
	}
#pragma unroll
	for (int o = 16; o >= 1; o >>= 1) maxv = max(maxv, __shfl_xor_sync(0xffffffffu, maxv, o));
	if (lane == 0 && maxv) atomicMax(&s_maxv, maxv);
	__syncthreads();
	if (work_list)
	{
		// counting sort by size class: class bases, then every chunk with vertices takes the next slot of its class (order inside a class is irrelevant)
		if (t == 0)
		{
			uint32_t run = 0;
			for (int k = 0; k < SIZE_CLASSES; k++) { s_cls_pos[k] = run; run += s_cls_cnt[k]; }
			totals[10] = run; // TOT_MESH: length of the list
		}
		__syncthreads();
		for (int i = t; i < n_chunks; i += SCAN_CTA)
		{
			const uint32_t nv = flags_contain_mesh(flags[i]) ? chunk_tot[3 * (size_t)i + 1] : 0u;
			if (nv) work_list[atomicAdd(&s_cls_pos[SIZE_CLASSES - 1 - min(SIZE_CLASSES - 1, (int)(nv >> 8))], 1u)] = i;
		}
	}
	if (t == 0)
	{
		const bool overflow = carry0 >= 0xFFFFFFFFull || carry1 >= 0xFFFFFFFFull || carry2 >= 0xFFFFFFFFull;
		totals[0] = carry0; totals[1] = carry1; totals[2] = carry2; totals[3] = overflow ? 1ull : 0ull;
		totals[4] = 0; totals[5] = 0; // surface-cell list counters (k_bases)
		totals[8] = s_maxv;           // largest chunk (vertices): decides whether chunk-local indices fit 16 bits (download.cuh)
		totals[9] = 0;                // download error flag
		totals[11] = 0;               // chunk ticket of k_chunk_emit (fused.cuh)
		totals[7] = (carry0 > cap_cells || carry1 > cap_verts || carry2 > cap_inds || overflow) ? 1ull : 0ull; // = k_check_caps
		totals[6] = 0;                // chunk work counter of k_smooth_chunks
	}
}

// Output arenas are sized from the previous batches so that a submission never has to wait for the host.  This
// one-thread kernel compares the totals of THIS batch with the capacities the emitters were launched with; if
// anything does not fit it raises tot[7] and every emitter below returns at once -- the host then grows the
// arenas and re-launches them (bmf_batch_wait).  tot = {cells, verts, indices, >2^32, list counters x2, -, too small}
// Device memory only: a store to host memory in the MIDDLE of the launch sequence makes the kernel wait for the PCIe write
// queue -- and when another context's mesh download is in flight (two contexts ping-pong), for that whole download.
__global__ void k_check_caps(unsigned long long* __restrict__ tot, unsigned long long cap_cells, unsigned long long cap_verts, unsigned long long cap_inds)
{
	if (blockIdx.x == 0 && threadIdx.x == 0)
	{
		tot[7] = (tot[0] > cap_cells || tot[1] > cap_verts || tot[2] > cap_inds || tot[3]) ? 1ull : 0ull;
		tot[6] = 0; // chunk work counter of k_smooth_chunks
	}
}

// LAST kernel of a batch's launch sequence: totals and (once per batch) the chunk table go to the host through mapped pinned memory written
// by the device itself -- chunks_words -> host_words as coalesced 4-byte words over all CTAs (n_words = 0: already there).  No copy-engine
// transfer, no host round trip; the stream's completion orders the stores for the host.
__global__ void __launch_bounds__(CTA) k_publish(const unsigned long long* __restrict__ tot, unsigned long long* __restrict__ tot_host /* mapped pinned host copy */,
                                                  const uint32_t* __restrict__ chunks_words, uint32_t* __restrict__ host_words, size_t n_words)
{
	if (blockIdx.x == 0 && threadIdx.x < 10) tot_host[threadIdx.x] = threadIdx.x == 6 || threadIdx.x == 9 ? 0ull : tot[threadIdx.x];
	for (size_t i = (size_t)blockIdx.x * CTA + threadIdx.x; i < n_words; i += (size_t)gridDim.x * CTA) host_words[i] = chunks_words[i];
}

// ---- K4a: per-word output bases + compaction of the CELLS that emit anything.  One CTA per segment with
// active cells (everything else leaves at once); the segment's sign planes are staged like in k_count.
// wv4[word] = {chunk-local id of the word's first vertex, X/Y/Z edge-owner masks}, wib[word] = batch-wide position of
// its first index (wv4 is only written -- and only ever read -- for words with active cells).
// Every cell that owns iso-vertices and every cell that polygonizes is appended to a compact list (a range per
// CTA reserved with one atomic; list order does not matter, each record carries its own output position):
//   vertex cell : {word, bit | edge flags << 5 | rank of its first vertex inside the word << 8}
//   index cell  : {word, bit | offset of its first index inside the word << 5 | mask8 << 16}
// so the emitters run one THREAD per surface cell, whatever the orientation of the surface inside the words.
template <int WPT>
__global__ void __launch_bounds__(CTA) k_bases(const uint32_t* __restrict__ bits, Layout L, const uint32_t* __restrict__ wcnt, const uint32_t* __restrict__ seg_tot,
                                                const ChunkCounts* __restrict__ chunks,
                                                uint4* __restrict__ wv4, uint32_t* __restrict__ wib, uint2* __restrict__ vcells,
                                                uint2* __restrict__ icells, unsigned long long* __restrict__ list_count /* [2] */,
                                                const unsigned long long* __restrict__ tot)
{
	extern __shared__ uint32_t sb[];
	__shared__ uint32_t s_base[2], s_pre[2];
	if (tot[7]) return; // an output arena is too small for this batch (k_check_caps): the host grows it and re-launches
	const int seg = blockIdx.x;
	const int chunk = seg >> L.lS, x0 = (seg & (L.S - 1)) * L.P;
	const ChunkCounts cc = chunks[chunk];
	if (!cc.contains_mesh || seg_tot[3 * (size_t)seg] == 0) return;
	// vertices / indices of the chunk's earlier segments (S <= 256 = CTA terms)
	if (threadIdx.x < 2) s_pre[threadIdx.x] = 0;
	__syncthreads();
	if ((int)threadIdx.x < (seg & (L.S - 1)))
	{
		const uint32_t* st = seg_tot + 3 * (((size_t)chunk << L.lS) + threadIdx.x);
		if (st[1]) atomicAdd(&s_pre[0], st[1]);
		if (st[2]) atomicAdd(&s_pre[1], st[2]);
	}
	stage_planes(sb, bits + (size_t)chunk * L.wc, L, x0, L.P + 1);
	uint32_t cnt[WPT];
	{
		const uint32_t* in = wcnt + (size_t)seg * L.ws + threadIdx.x * WPT;
#pragma unroll
		for (int k = 0; k < WPT; k += 4)
		{
			uint4 v = *reinterpret_cast<const uint4*>(in + k);
			cnt[k] = v.x; cnt[k + 1] = v.y; cnt[k + 2] = v.z; cnt[k + 3] = v.w;
		}
	}
	__syncthreads();
	uint32_t sc[5] = { 0, 0, 0, 0, 0 }; // verts, indices, vertex cells, index cells, words with active cells
	uint32_t nvc[WPT], nic[WPT];
#pragma unroll
	for (int k = 0; k < WPT; k++)
	{
		nvc[k] = nic[k] = 0;
		if (cnt[k] == 0) continue;
		const int lw = threadIdx.x * WPT + k;
		const int zb = lw & (L.zc - 1), y = (lw >> L.lzc) & (L.d - 1), lx = lw >> L.lwp;
		const WordBits b = load_word_bits(sb, L, lx, y, zb);
		const WordClass c = classify(b, L, x0 + lx, y, zb);
		nvc[k] = __popc(c.ex | c.ey | c.ez);
		nic[k] = __popc(c.active & c.interior);
		sc[0] += (cnt[k] >> 8) & 0xFF;
		sc[1] += cnt[k] >> 16;
		sc[2] += nvc[k];
		sc[3] += nic[k];
		sc[4] += 1;
	}
	uint32_t cta_tot[5];
	block_scan<5>(sc, cta_tot);
	if (threadIdx.x == 0)
	{
		s_base[0] = (uint32_t)atomicAdd(&list_count[0], (unsigned long long)cta_tot[2]);
		s_base[1] = (uint32_t)atomicAdd(&list_count[1], (unsigned long long)cta_tot[3]);
	}
	__syncthreads();
	// Phase 2a: every thread walks its own words once more, only to hand each word WITH active cells -- with the four
	// running positions it starts at -- to a shared work list; the per-cell work below then runs on dense warps instead of
	// on the few threads whose words happen to be crossed by the surface.
	uint4* s_job = reinterpret_cast<uint4*>(sb + (size_t)(L.P + 1) * L.wp); // [ws] {local word | vertex id << 12 .., ...}: see below
	uint32_t* s_job_w = reinterpret_cast<uint32_t*>(s_job + L.ws);
	{
		uint32_t rv = s_pre[0] + sc[0];                          // chunk-local vertex id
		uint32_t ri = (uint32_t)cc.ind_base + s_pre[1] + sc[1];  // batch-wide index position
		uint32_t ov = s_base[0] + sc[2], oi = s_base[1] + sc[3];
		uint32_t slot = sc[4];
		uint32_t oix[WPT];
#pragma unroll
		for (int k = 0; k < WPT; k++)
		{
			oix[k] = ri;
			if (cnt[k] == 0) continue; // nothing references a word without active cells
			s_job[slot] = make_uint4(rv, ri, ov, oi);
			s_job_w[slot] = threadIdx.x * WPT + k;
			slot++;
			rv += (cnt[k] >> 8) & 0xFF;
			ri += cnt[k] >> 16;
			ov += nvc[k];
			oi += nic[k];
		}
		uint32_t* o2 = wib + (size_t)seg * L.ws + threadIdx.x * WPT;
#pragma unroll
		for (int k = 0; k < WPT; k += 4)
			*reinterpret_cast<uint4*>(o2 + k) = make_uint4(oix[k], oix[k + 1], oix[k + 2], oix[k + 3]);
	}
	__syncthreads();
	// Phase 2b: one thread per listed word
	for (uint32_t j = threadIdx.x; j < cta_tot[4]; j += CTA)
	{
		const uint4 job = s_job[j];
		const int lw = (int)s_job_w[j];
		const uint32_t gw = (uint32_t)((size_t)seg * L.ws + lw);
		const int zb = lw & (L.zc - 1), y = (lw >> L.lzc) & (L.d - 1), lx = lw >> L.lwp;
		const WordBits b = load_word_bits(sb, L, lx, y, zb);
		const WordClass c = classify(b, L, x0 + lx, y, zb);
		// vertex record of the word: first vertex id + the three edge-owner masks (everything a vertex id needs)
		wv4[gw] = make_uint4(job.x, c.ex, c.ey, c.ez);
		uint32_t ov = job.z, oi = job.w;
		uint32_t m = c.ex | c.ey | c.ez, rank = 0;
		while (m)
		{
			const int bit = __ffs(m) - 1;
			m &= m - 1;
			const uint32_t fl = ((c.ex >> bit) & 1u) | (((c.ey >> bit) & 1u) << 1) | (((c.ez >> bit) & 1u) << 2);
			vcells[ov++] = make_uint2(gw, (uint32_t)bit | (fl << 5) | (rank << 8));
			rank += __popc(fl);
		}
		m = c.active & c.interior;
		uint32_t ofs = 0;
		while (m)
		{
			const int bit = __ffs(m) - 1;
			m &= m - 1;
			const uint32_t m8 = mask8_of(b, bit);
			icells[oi++] = make_uint2(gw, (uint32_t)bit | (ofs << 5) | (m8 << 16));
			ofs += (uint32_t)(c_tri_pack[m8] >> 60);
		}
	}
}

// ---- K4b: vertex emission, one THREAD per cell that owns iso-vertices.
// calculate_cell / calculate_isovertex / _get_intersection (DMCChunk.cpp:593-674): X, Y, Z edge of a cell in that order.
__global__ void __launch_bounds__(CTA) k_verts3(Layout L, const uint4* __restrict__ wv4, const ChunkCounts* __restrict__ chunks, SamplerDev s, DensitySource src,
                                                 const ChunkGeom* __restrict__ geom, const uint2* __restrict__ vcells,
                                                 const unsigned long long* __restrict__ list_count, float* __restrict__ pos, uint8_t* __restrict__ boundary,
                                                 uint32_t* __restrict__ cls, float* __restrict__ normal, const unsigned long long* __restrict__ tot)
{
	if (tot[7]) return;
	const uint32_t n_cells = (uint32_t)list_count[0];
	const uint32_t stride = gridDim.x * CTA;
	const int d = L.d;
	for (uint32_t i = blockIdx.x * CTA + threadIdx.x; i < n_cells; i += stride)
	{
		const uint2 rec = vcells[i];
		const uint32_t gw = rec.x;
		const int bit = rec.y & 31, fl = (rec.y >> 5) & 7, rank = rec.y >> 8;
		const int chunk = (int)(gw >> L.lwc);
		const int w = (int)(gw & (uint32_t)(L.wc - 1));
		const int zb = w & (L.zc - 1), y = (w >> L.lzc) & (d - 1), x = w >> L.lwp;
		const int z = zb * 32 + bit;
		const ChunkGeom g = geom[chunk];
		size_t v = (size_t)chunks[chunk].vert_base + wv4[gw].x + rank;
		const float s0 = density_at(s, src, g, d, chunk, x, y, z);
		const bool b0 = x == 0 || y == 0 || z == 0 || x == d - 1 || y == d - 1 || z == d - 1;
#pragma unroll
		for (int axis = 0; axis < 3; axis++)
		{
			if (!((fl >> axis) & 1)) continue;
			const int x1 = x + (axis == 0), y1 = y + (axis == 1), z1 = z + (axis == 2);
			const float s1 = density_at(s, src, g, d, chunk, x1, y1, z1);
			const float mu = (0.0f - s0) / (s1 - s0);
			// (p1 - p0) * mu + p0 per component, p in grid units
			pos[3 * v + 0] = ((float)x1 - (float)x) * mu + (float)x;
			pos[3 * v + 1] = ((float)y1 - (float)y) * mu + (float)y;
			pos[3 * v + 2] = ((float)z1 - (float)z) * mu + (float)z;
			boundary[v] = (b0 || x1 == d - 1 || y1 == d - 1 || z1 == d - 1) ? 1 : 0;
			// every vertex is written exactly once, here: its use counters (k_inds3 adds to them) and its normal start at zero
			cls[v] = 0u;
			normal[3 * v + 0] = 0.0f; normal[3 * v + 1] = 0.0f; normal[3 * v + 2] = 0.0f;
			v++;
		}
	}
}

// chunk-local id of the vertex on `axis` of cell (x,y,z): one 16-byte record of its word + popcounts below the cell
__device__ __forceinline__ uint32_t vertex_id_rec(const uint4* __restrict__ rec4, const Layout& L, int x, int y, int z, int axis)
{
	const int bit = z & 31;
	const uint4 r = rec4[(((x << L.ld) + y) << L.lzc) + (z >> 5)];
	const uint32_t lt = (1u << bit) - 1u;
	uint32_t id = r.x + __popc(r.y & lt) + __popc(r.z & lt) + __popc(r.w & lt);
	if (axis >= 1) id += (r.y >> bit) & 1u;
	if (axis >= 2) id += (r.z >> bit) & 1u;
	return id;
}

// ---- K4c: index emission (polygonize / polygonize_cell, DMCChunk.cpp:514-576) + per-class use counts.  The cell's
// corner mask travels in its record; the ids of the edge vertices the triangle table names (EDGE_V, DMCChunk.cpp:32,
// 543-565) come from the 16-byte vertex records of the neighbouring words.
// A warp takes 32 polygonizing cells.  A cell uses 3..12 edges and emits 3..15 indices, so one thread per cell would
// leave most lanes idle most of the time; instead both steps are flattened over the warp: lane j takes the j-th
// (cell, used edge) pair of the 32 cells, then the j-th (cell, index) pair -- prefix sums over the warp,
// the owning cell of every pair written into a small shared table by the cells themselves, ids handed over through
// shared memory.
__global__ void __launch_bounds__(CTA) k_inds3(Layout L, const uint4* __restrict__ wv4, const uint32_t* __restrict__ wib,
                                                const ChunkCounts* __restrict__ chunks, const uint2* __restrict__ icells,
                                                const unsigned long long* __restrict__ list_count, uint32_t* __restrict__ inds, uint32_t* __restrict__ cls,
                                                const unsigned long long* __restrict__ tot)
{
	__shared__ uint64_t s_tri[256];
	__shared__ uint32_t s_id[CTA][12];
	__shared__ uint64_t s_tp[CTA];
	__shared__ uint32_t s_xyz[CTA], s_chunk[CTA], s_used[CTA], s_out[CTA], s_vb[CTA];
	__shared__ uint32_t s_pre[CTA / 32][33];
	__shared__ uint8_t s_own[CTA / 32][480]; // owning cell (0..31) of every (cell, edge) / (cell, index) pair of the warp
	if (tot[7]) return;
	s_tri[threadIdx.x] = g_tri_pack[threadIdx.x];
	__syncthreads();
	const uint32_t n_cells = (uint32_t)list_count[1];
	const uint32_t stride = gridDim.x * CTA;
	const int d = L.d, lane = threadIdx.x & 31, wbase = threadIdx.x & ~31;
	uint32_t* pre = s_pre[threadIdx.x >> 5];
	for (uint32_t i0 = blockIdx.x * CTA + wbase; i0 < n_cells; i0 += stride) // warp-uniform
	{
		const uint32_t i = i0 + lane;
		uint32_t used = 0, n = 0;
		if (i < n_cells)
		{
			const uint2 rec = icells[i];
			const uint32_t gw = rec.x, ofs = (rec.y >> 5) & 0x7FF, m8 = (rec.y >> 16) & 0xFF;
			const uint64_t tp = s_tri[m8];
			n = (uint32_t)(tp >> 60);
			// the edges the table uses = the cell's sign-changing edges: X-edges e0-3 <-> corner pairs (a, a+4), Y-edges e4-7 <-> pairs
			// (0,2),(1,3),(4,6),(5,7), Z-edges e8-11 <-> pairs (0,1),(2,3),(4,5),(6,7)
			const uint32_t tx = (m8 ^ (m8 >> 4)) & 0xFu, ty = m8 ^ (m8 >> 2), tz = m8 ^ (m8 >> 1);
			used = tx | ((ty & 3u) << 4) | (((ty >> 4) & 3u) << 6) | ((tz & 1u) << 8) | (((tz >> 2) & 1u) << 9) | (((tz >> 4) & 1u) << 10) | (((tz >> 6) & 1u) << 11);
			const int w = (int)(gw & (uint32_t)(L.wc - 1));
			s_tp[threadIdx.x] = tp;
			s_xyz[threadIdx.x] = (uint32_t)(w >> L.lwp) | ((uint32_t)((w >> L.lzc) & (d - 1)) << 10) | ((uint32_t)((w & (L.zc - 1)) * 32 + (rec.y & 31)) << 20);
			s_chunk[threadIdx.x] = gw >> L.lwc;
			s_used[threadIdx.x] = used;
			s_out[threadIdx.x] = wib[gw] + ofs;
			s_vb[threadIdx.x] = (uint32_t)chunks[gw >> L.lwc].vert_base;
		}
		// ---- step 1: vertex ids of the used edges, one (cell, edge) pair per lane and round
		uint32_t inc = __popc(used);
#pragma unroll
		for (int o = 1; o < 32; o <<= 1)
		{
			const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
			if (lane >= o) inc += u;
		}
		uint8_t* own = s_own[threadIdx.x >> 5];
		pre[lane + 1] = inc;
		if (lane == 0) pre[0] = 0;
		{
			const uint32_t u = __popc(used);
			for (uint32_t q = inc - u; q < inc; q++) own[q] = (uint8_t)lane;
		}
		__syncwarp();
		const uint32_t n_pairs = pre[32];
		for (uint32_t j = lane; j < n_pairs; j += 32)
		{
			const int c = own[j];
			const int e = (int)__fns(s_used[wbase + c], 0, (int)(j - pre[c]) + 1);
			const uint32_t xyz = s_xyz[wbase + c];
			const int x = (int)(xyz & 1023u), y = (int)((xyz >> 10) & 1023u), z = (int)(xyz >> 20);
			const int axis = e >> 2, hi = (e >> 1) & 1, lo = e & 1;
			const int dx = axis == 0 ? 0 : hi;
			const int dy = axis == 0 ? hi : (axis == 1 ? 0 : lo);
			const int dz = axis == 2 ? 0 : lo;
			s_id[wbase + c][e] = vertex_id_rec(wv4 + ((size_t)s_chunk[wbase + c] << L.lwc), L, x + dx, y + dy, z + dz, axis);
		}
		__syncwarp();
		// ---- step 2: the indices, one (cell, index) pair per lane and round
		inc = n;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1)
		{
			const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
			if (lane >= o) inc += u;
		}
		pre[lane + 1] = inc;
		for (uint32_t q = inc - n; q < inc; q++) own[q] = (uint8_t)lane;
		__syncwarp();
		const uint32_t n_inds = pre[32];
		for (uint32_t j = lane; j < n_inds; j += 32)
		{
			const int c = own[j];
			const uint32_t t = j - pre[c];
			const uint32_t e = (uint32_t)(s_tp[wbase + c] >> (4 * t)) & 15u;
			const uint32_t vid = s_id[wbase + c][e];
			inds[(size_t)s_out[wbase + c] + t] = vid;
			// init_valence++ (DMCChunk.cpp:573), kept per "cell class": a vertex is shared by at most the four cells
			// around its grid edge and class = 3 - (e & 3) is this cell's place among them in scan order; byte c of
			// cls[v] counts the uses by the class-c cell (the sort-free CSR build in smooth.cuh needs the split)
			atomicAdd(cls + s_vb[wbase + c] + vid, 1u << (8 * (3 - (e & 3))));
		}
		__syncwarp(); // the next round overwrites the warp's shared slots
	}
}

// init_valence of every vertex (= sum of its four per-class use counts) and, in the same pass, adj_offset = exclusive
// prefix of init_valence (MeshProcessor.cpp:33-39).  The valences of a chunk add up to its index count, so the
// batch-wide prefix at the chunk's first vertex is simply its ind_base: one CTA per chunk scans its own vertices
// in tiles with a running carry -- no device-wide scan, one launch.
static constexpr int VAL_ITEMS = 8; // (32 items per thread was measured slower: 64 us against 38 us)

__global__ void __launch_bounds__(CTA) k_valence_offsets(const uint32_t* __restrict__ cls, const ChunkCounts* __restrict__ chunks, uint8_t* __restrict__ valence,
                                                          uint32_t* __restrict__ adj_off, const unsigned long long* __restrict__ tot,
                                                          uint32_t* __restrict__ host_table /* mapped pinned copy of the chunk table, or null */)
{
	// one CTA per chunk: the natural place to hand the chunk's table record to the host (posted PCIe writes, nobody waits for them)
	constexpr int REC = (int)(sizeof(ChunkCounts) / sizeof(uint32_t));
	if (host_table && threadIdx.x < REC)
		host_table[(size_t)blockIdx.x * REC + threadIdx.x] = reinterpret_cast<const uint32_t*>(chunks)[(size_t)blockIdx.x * REC + threadIdx.x];
	if (tot[7]) return;
	const ChunkCounts cc = chunks[blockIdx.x];
	if (cc.n_verts == 0) return;
	const size_t v0 = (size_t)cc.vert_base;
	uint32_t carry = (uint32_t)cc.ind_base;
	// a tile is VAL_ITEMS rows of CTA consecutive vertices; thread t holds vertex t of every row, so that every load and store of
	// a warp is one contiguous run (a thread-contiguous layout made every byte store of a warp touch 8 different sectors)
	for (uint32_t base = 0; base < cc.n_verts; base += CTA * VAL_ITEMS)
	{
		uint32_t val[VAL_ITEMS], sc[VAL_ITEMS], rt[VAL_ITEMS];
#pragma unroll
		for (int k = 0; k < VAL_ITEMS; k++)
		{
			const uint32_t i = base + k * CTA + threadIdx.x;
			const uint32_t w = i < cc.n_verts ? cls[v0 + i] : 0u;
			val[k] = (w & 0xFF) + ((w >> 8) & 0xFF) + ((w >> 16) & 0xFF) + (w >> 24);
			sc[k] = val[k];
		}
		block_scan<VAL_ITEMS>(sc, rt); // per row: exclusive prefix over the threads, row total
		uint32_t run = carry;
#pragma unroll
		for (int k = 0; k < VAL_ITEMS; k++)
		{
			const uint32_t i = base + k * CTA + threadIdx.x;
			if (i < cc.n_verts)
			{
				valence[v0 + i] = (uint8_t)val[k];
				adj_off[v0 + i] = run + sc[k];
			}
			run += rt[k];
		}
		carry = run;
	}
}

} // namespace bmf

// fused.cuh -- label_edges + polygonize + MeshProcessor::init of a whole batch in TWO per-chunk kernels (dim <= 64, triangles):
//   k_chunk_count  one CTA per chunk: the chunk's sign field goes to shared memory once, every 32-cell word is classified once,
//                  its packed (vertices, indices) count is written out and the chunk's totals are reduced;
//   k_scan_chunks / k_check_caps (extract.cuh): chunk bases in batch order, capacity verdict, chunk table for the host;
//   k_chunk_emit   one CTA per mesh chunk (persistent, atomic ticket): sign field + word counts back into shared memory, prefix
//                  inside the chunk, then vertices, indices, init_valence / adj_offset and the CSR adjacency -- everything a
//                  vertex id needs comes out of shared memory.
// They replace k_count -> k_bases -> k_verts3 -> k_inds3 -> k_valence_offsets -> k_adj_fill (per-segment CTAs, three classifications of
// every word, per-word 16-byte records and index bases round-tripping through HBM) for the chunk sizes the reference's worlds use
// (chunk_resolution 32 / 64, WorldOctree.cpp:20-33).
//
// What stays exactly as in the multi-kernel path -- and therefore bit-identical to the reference's serial scan
// (DMCChunk.cpp:440-498 cell order, :514-576 index order, MeshProcessor.cpp:98-128 adjacency order):
//   * vertex ids / index positions are the exclusive prefix, in x -> y -> z word order, of the per-word vertex / index counts;
//   * chunk bases are the exclusive prefix of the chunk totals in BATCH order.
// (A first version did the chunk prefix with a decoupled look-back inside one kernel; the CTAs spent a sixth of their time waiting
// for the slowest predecessor's classification, so the prefix went back to its own tiny launch.)
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "extract.cuh"
#include "download.cuh"

namespace bmf
{

enum
{
	TOT_MESH = 10,    // length of the work list (k_scan_chunks)
	TOT_TICKET = 11,  // next emit-list position (k_chunk_emit)
	TOT_CAND = 12,    // length of the candidate list (k_terrain2d_classify)
	TOT_CTICKET = 13  // next candidate (k_chunk_count)
};

typedef unsigned long long u64;
static constexpr int FUSED_NT = 512; // threads per chunk CTA: two CTAs per SM at 64^3 (107 KB of shared memory each)
static constexpr int FUSED_CAPV = 8192; // chunks with at most this many vertices keep their use counters and adjacency offsets in shared memory

struct FusedArgs
{
	Layout L;
	int n;
	const uint32_t* bits;
	const uint32_t* wcnt;  // [n][wc] packed (vertices | indices << 16) of every word, written by k_chunk_count
	const int* emit_list;  // the chunks that have vertices, largest first (k_scan_chunks)
	const ChunkCounts* chunks;
	u64* tot;
	uint2* vcells;
	uint2* icells;
	SamplerDev s;
	DensitySource src;
	const ChunkGeom* geom;
	float* pos;
	uint8_t* boundary;
	float* normal;
	uint32_t* cls;
	uint32_t* inds;
	uint8_t* valence;
	uint32_t* adj_off;
	uint32_t* adj;        // null: no smoothing follows, the CSR is not built
	uint32_t* prim_vbase;
	u64* prof;            // debugging aid (BMF_FUSED_PROF=1): [chunk][16] SM clock at the phase boundaries, thread 0; null otherwise
};

#define BMF_FUSED_MARK(k) do { if (A.prof && tid == 0) A.prof[(size_t)chunk * 16 + (k)] = (u64)clock64(); } while (0)

// dynamic shared memory of k_chunk_emit<NT> for a layout (bytes)
__host__ __device__ inline size_t fused_smem_bytes(const Layout& L, int nt)
{
	const size_t scratch = (size_t)44 * nt + 256, jobs = (size_t)2 * L.wc;
	return ((size_t)(L.d + 1) * L.wp + L.wc) * 4 + (size_t)(L.wc / 32) * 8 + (scratch > jobs ? scratch : jobs) + (size_t)FUSED_CAPV * 2;
}

template <int NT, int K>
__device__ __forceinline__ void block_scan_nt(uint32_t (&v)[K], uint32_t (&tot)[K], uint32_t (*s_w)[NT / 32], uint32_t* s_t)
{
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
			const uint32_t w = lane < NT / 32 ? s_w[q][lane] : 0;
			uint32_t j = w;
#pragma unroll
			for (int o = 1; o < NT / 32; o <<= 1)
			{
				const uint32_t t = __shfl_up_sync(0xffffffffu, j, o);
				if (lane >= o) j += t;
			}
			if (lane < NT / 32) s_w[q][lane] = j - w;
			if (lane == NT / 32 - 1) s_t[q] = j;
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

struct FusedView
{
	const uint32_t* sb;
	const uint32_t* off;
	const uint2* gb;
};

__device__ __forceinline__ void word_bases(const FusedView& S, int w, uint32_t& vbase, uint32_t& ibase)
{
	const uint2 g = S.gb[w >> 5];
	const uint32_t o = S.off[w];
	vbase = g.x + (o & 0xFFFFu);
	ibase = g.y + (o >> 16);
}

// chunk-local id of the vertex on `axis` of the cell at bit `bit` of word `w`, everything out of shared memory (vertex_id_rec's arithmetic;
// the plane at x = d is staged as zeros, so the loads below never leave the staged window)
__device__ __forceinline__ uint32_t vertex_id_smem(const FusedView& S, const Layout& L, int w, int bit, int axis)
{
	const uint32_t A = S.sb[w];
	const uint32_t ex = (w < L.wc - L.wp) ? (A ^ S.sb[w + L.wp]) : 0u;                       // x + 1 < d
	const uint32_t ey = (((w >> L.lzc) & (L.d - 1)) != L.d - 1) ? (A ^ S.sb[w + L.zc]) : 0u; // y + 1 < d
	const bool zn = (w & (L.zc - 1)) != L.zc - 1;                                            // another word follows in this row
	const uint32_t A1 = __funnelshift_r(A, zn ? S.sb[w + 1] : 0u, 1);
	const uint32_t ez = (A ^ A1) & (zn ? 0xFFFFFFFFu : 0x7FFFFFFFu);
	const uint32_t lt = (1u << bit) - 1u;
	uint32_t id = S.gb[w >> 5].x + (S.off[w] & 0xFFFFu) + __popc(ex & lt) + __popc(ey & lt) + __popc(ez & lt);
	if (axis >= 1) id += (ex >> bit) & 1u;
	if (axis >= 2) id += (ey >> bit) & 1u;
	return id;
}

// ---- K3 for whole chunks: classification of every word, ONCE.  wcnt[word] = vertices | indices << 16, chunk_tot = (cells, vertices,
// indices) of the chunk; optionally the MasksBlock byte image (DMCChunk.cpp:184-438).  Persistent CTAs take chunks from `cand` (the
// chunks the 2-D terrain classifier could not cull) or, without such a list, from 0..n-1; a chunk without a mesh (label_edges returns at
// once, DMCChunk.cpp:170-171) costs one ticket.  COUNT_NT threads per chunk so that six CTAs share an SM and no wave is left half empty.
static constexpr int COUNT_NT = 256;

// ---- TMA (bulk async copy) staging of a chunk's sign words: one elected thread arms an mbarrier with the byte count and issues ONE
// cp.async.bulk global -> shared for the whole 4 / 32 KB image; everybody waits on the barrier's phase.  A/B against the 128-bit load loop:
// BMF_TMA=1 (DESIGN.md section 4 has the numbers).
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// the elected thread: arrive on the barrier announcing `bytes` of asynchronous transfers (all the bulk_copy calls of this phase together)
__device__ __forceinline__ void mbar_expect(uint64_t* bar, uint32_t bytes)
{
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // earlier generic-proxy accesses to the destinations are ordered before the async writes
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy(void* dst_smem, const void* src_global, uint32_t bytes, uint64_t* bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)), "l"(src_global), "r"(bytes),
	             "r"(smem_u32(bar))
	             : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst_smem, const void* src_global, uint32_t bytes, uint64_t* bar)
{
	mbar_expect(bar, bytes);
	bulk_copy(dst_smem, src_global, bytes, bar);
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
	asm volatile("{\n"
	             ".reg .pred p;\n"
	             "BMF_WAIT_%=:\n"
	             "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
	             "@p bra BMF_DONE_%=;\n"
	             "bra BMF_WAIT_%=;\n"
	             "BMF_DONE_%=:\n"
	             "}" ::"r"(smem_u32(bar)),
	             "r"(parity)
	             : "memory");
}

// GEN (2-D noise terrains): the chunk's sign words are not loaded but MADE here, from its noise sheet and its y coordinates (terrain2d_tile_word,
// a warp per 32 x 32 (y, z) tile), into shared memory and -- for the emitter and bmf_batch_copy_chunk -- into `bits`; the chunk's flags are formed
// in the CTA.  That is the whole of k_terrain2d_bits, without its launch, its pass over the words and this kernel's load of them.
struct Gen2D
{
	SamplerDev s;
	const ChunkGeom* geom;
	const float* hmap;
	const int* sheet_of;
};

template <int NT, bool TMA, bool GEN>
__global__ void __launch_bounds__(NT) k_chunk_count(uint32_t* __restrict__ bits, uint32_t* __restrict__ flags, Layout L, int n, const int* __restrict__ cand,
                                                     const u64* __restrict__ cand_count, uint32_t* __restrict__ wcnt, uint32_t* __restrict__ chunk_tot,
                                                     uint8_t* __restrict__ masks, u64* __restrict__ tot /* [TOT_CTICKET] */, Gen2D G)
{
	__shared__ uint32_t s_flags;
	extern __shared__ __align__(16) uint32_t dyn[];
	__shared__ u64 s_tri[256];
	__shared__ uint32_t s_red[3][NT / 32];
	__shared__ int s_chunk;
	__shared__ __align__(8) uint64_t s_bar;
	const int d = L.d, wc = L.wc, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
	uint32_t* sb = dyn;
	for (int i = tid; i < 256; i += NT) s_tri[i] = g_tri_pack[i];
	if (TMA && tid == 0) mbar_init(&s_bar, 1);
	uint32_t phase = 0;
	const u64 n_cand = cand ? *cand_count : (u64)n;
	for (;;)
	{
		__syncthreads();
		if (tid == 0)
		{
			int c = -1;
			for (;;)
			{
				const u64 j = atomicAdd(tot + TOT_CTICKET, 1ull);
				if (j >= n_cand) break;
				const int k = cand ? cand[j] : (int)j;
				if (GEN || flags_contain_mesh(flags[k])) { c = k; break; } // GEN: the flags are not known before the words are made
			}
			s_chunk = c;
			s_flags = 0;
		}
		__syncthreads();
		const int chunk = s_chunk;
		if (chunk < 0) return;
		if (GEN)
		{
			const ChunkGeom g = G.geom[chunk];
			const float* sheet = G.hmap + ((size_t)G.sheet_of[chunk] << (2 * L.ld));
			uint32_t* out = bits + (size_t)chunk * wc;
			uint32_t f = 0;
			const int n_tasks = d << (2 * L.lzc); // (x, yb, zb)
			for (int task = wid; task < n_tasks; task += NT / 32)
			{
				const int zb = task & (L.zc - 1), yb = (task >> L.lzc) & (L.zc - 1), x = task >> (2 * L.lzc);
				const uint32_t mine = terrain2d_tile_word(G.s, g, L, sheet, x, yb, zb, lane);
				const int w = ((((x << L.ld) + yb * 32 + lane)) << L.lzc) + zb;
				sb[w] = mine;
				out[w] = mine;
				f |= word_flags(mine);
			}
			for (int i = tid; i < L.wp; i += NT) sb[wc + i] = 0u; // plane x = d: B == 0 outside the grid
			f |= __shfl_xor_sync(0xffffffffu, f, 16);
			f |= __shfl_xor_sync(0xffffffffu, f, 8);
			f |= __shfl_xor_sync(0xffffffffu, f, 4);
			f |= __shfl_xor_sync(0xffffffffu, f, 2);
			f |= __shfl_xor_sync(0xffffffffu, f, 1);
			if (lane == 0 && f) atomicOr(&s_flags, f);
			__syncthreads();
			const uint32_t fl = s_flags;
			if (tid == 0) flags[chunk] = fl; // this CTA is the only writer: the classifier leaves the chunks it lists untouched
			if (!flags_contain_mesh(fl)) continue; // (chunk_tot of such a chunk is never read)
		}
		else if (TMA)
		{
			if (tid == 0) bulk_load(sb, bits + (size_t)chunk * wc, (uint32_t)wc * 4u, &s_bar);
			for (int i = tid; i < L.wp; i += NT) sb[wc + i] = 0u; // plane x = d: B == 0 outside the grid
			mbar_wait(&s_bar, phase);
			phase ^= 1u;
		}
		else
		{
			const uint4* src = reinterpret_cast<const uint4*>(bits + (size_t)chunk * wc);
			uint4* dst = reinterpret_cast<uint4*>(sb);
			for (int i = tid; i < wc / 4; i += NT) dst[i] = src[i];
			for (int i = tid; i < L.wp; i += NT) sb[wc + i] = 0u; // plane x = d: B == 0 outside the grid
		}
		__syncthreads();
		uint32_t cells = 0, tv = 0, ti = 0;
		for (int j = 0; j < wc / NT; j++)
		{
			// a warp takes 32 consecutive words (coalesced count stores) of round j -- a DIFFERENT slot of the round every time: with a fixed
			// slot a warp would see the same band of y rows at every x, and the one warp whose band holds the surface would do all the
			// per-cell work while the others wait at the barrier
			const int w = j * NT + (((wid + j) & (NT / 32 - 1)) << 5) + lane;
			const int zb = w & (L.zc - 1), y = (w >> L.lzc) & (d - 1), x = w >> L.lwp;
			const WordBits b = load_word_bits(sb, L, x, y, zb);
			const WordClass c = classify(b, L, x, y, zb);
			uint32_t nv = 0, ni = 0;
			if (c.active)
			{
				cells += __popc(c.active);
				nv = __popc(c.ex) + __popc(c.ey) + __popc(c.ez);
				uint32_t m = c.active & c.interior;
				while (m)
				{
					const int bit = __ffs(m) - 1;
					m &= m - 1;
					ni += (uint32_t)(s_tri[mask8_of(b, bit)] >> 60);
				}
			}
			wcnt[(size_t)chunk * wc + w] = nv | (ni << 16);
			tv += nv; ti += ni;
			if (masks)
			{
				uint8_t* mrow = masks + (size_t)chunk * wc * 32 + ((size_t)x * d + y) * d + zb * 32;
#pragma unroll
				for (int q = 0; q < 8; q++)
					reinterpret_cast<uint32_t*>(mrow)[q] = mask8_of(b, 4 * q) | (mask8_of(b, 4 * q + 1) << 8) | (mask8_of(b, 4 * q + 2) << 16) | (mask8_of(b, 4 * q + 3) << 24);
			}
		}
#pragma unroll
		for (int o = 16; o >= 1; o >>= 1)
		{
			cells += __shfl_xor_sync(0xffffffffu, cells, o);
			tv += __shfl_xor_sync(0xffffffffu, tv, o);
			ti += __shfl_xor_sync(0xffffffffu, ti, o);
		}
		if (lane == 0) { s_red[0][wid] = cells; s_red[1][wid] = tv; s_red[2][wid] = ti; }
		__syncthreads();
		if (tid < 3)
		{
			uint32_t t = 0;
			for (int k = 0; k < NT / 32; k++) t += s_red[tid][k];
			chunk_tot[3 * (size_t)chunk + tid] = t; // (k_scan_chunks turns these into bases and into the largest-first work list of k_chunk_emit)
		}
	}
}

template <int NT, bool TMA>
__global__ void __launch_bounds__(NT, (NT <= 512 ? 2 : 1)) k_chunk_emit(FusedArgs A)
{
	extern __shared__ __align__(16) uint32_t dyn[];
	constexpr int NW = NT / 32;
	const Layout L = A.L;
	const int d = L.d, wc = L.wc, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
	uint32_t* sb = dyn;
	uint32_t* s_off = sb + (d + 1) * L.wp;
	uint2* s_gb = reinterpret_cast<uint2*>(s_off + wc);
	uint8_t* s_r = reinterpret_cast<uint8_t*>(s_gb + wc / 32);
	// region R, phases 1 / 5b: the list of words that emit anything
	uint16_t* s_job = reinterpret_cast<uint16_t*>(s_r);
	// region R, phases 5d / 5f: per-thread cell slots and per-warp flattening tables
	u64* s_tp = reinterpret_cast<u64*>(s_r);                                    // [NT] the cell's triangle-table row
	u64* s_lc = s_tp + NT;                                                      // [NT] 3 bits per index: earlier uses of the same edge inside the row
	uint32_t* s_xyz = reinterpret_cast<uint32_t*>(s_lc + NT);                   // [NT] word | bit << 13 of the cell
	uint32_t* s_out = s_xyz + NT;                                               // [NT] 5d: first index position inside the chunk; 5f: first primitive
	uint32_t* s_pre_all = s_out + NT;                                           // [NW][40]
	uint8_t* s_own_all = reinterpret_cast<uint8_t*>(s_pre_all + NW * 40);       // [NW][480]
	// per-vertex use counters, 4 bits per "cell class" (a cell uses one vertex at most 5 times), 16 bits per vertex, two vertices per word
	const size_t r_bytes = ((size_t)44 * NT + 256 > (size_t)2 * wc) ? (size_t)44 * NT + 256 : (size_t)2 * wc;
	uint32_t* s_cls = reinterpret_cast<uint32_t*>(s_r + r_bytes);                // [FUSED_CAPV / 2]
	uint32_t* s_aoff = sb;                                                       // [FUSED_CAPV] adjacency offsets, over the sign words once 5d is done
	__shared__ u64 s_tri[256], s_loc[256];
	__shared__ uint32_t s_edge[16];
	__shared__ int s_chunk;
	__shared__ uint32_t s_njobs, s_nvc, s_nic;
	__shared__ __align__(8) uint64_t s_bar;
	__shared__ uint2 s_wtot[NW], s_woff[NW];
	__shared__ uint32_t s_scan[VAL_ITEMS][NW], s_scant[VAL_ITEMS];
	FusedView S;
	S.sb = sb; S.off = s_off; S.gb = s_gb;

	if (A.tot[TOT_SMALL]) return; // an output arena is too small for this batch (k_check_caps): the host grows it and re-launches
	for (int i = tid; i < 256; i += NT)
	{
		// the triangle-table row of corner mask i and, 3 bits per index t, how many of the indices before t name the same edge (<= 4)
		const u64 tp = g_tri_pack[i];
		const int n = (int)(tp >> 60);
		u64 lc = 0;
		for (int t = 1; t < n; t++)
		{
			const uint32_t e = (uint32_t)(tp >> (4 * t)) & 15u;
			uint32_t c = 0;
			for (int q = 0; q < t; q++) c += (((uint32_t)(tp >> (4 * q)) & 15u) == e) ? 1u : 0u;
			lc |= (u64)c << (3 * t);
		}
		s_tri[i] = tp;
		s_loc[i] = lc;
	}
	if (TMA && tid == 0) mbar_init(&s_bar, 1);
	uint32_t phase = 0;
	if (tid < 12)
	{
		// edge e of a cell (EDGE_V, DMCChunk.cpp:32, 543-565): X-edges 0-3 at (y+hi, z+lo), Y-edges 4-7 at (x+hi, z+lo), Z-edges 8-11 at (x+hi, y+lo)
		// -> word offset of the cell that owns the edge's vertex | dz << 16 | axis << 20
		const int e = tid, axis = e >> 2, hi = (e >> 1) & 1, lo = e & 1;
		const int dx = axis == 0 ? 0 : hi, dy = axis == 0 ? hi : (axis == 1 ? 0 : lo), dz = axis == 2 ? 0 : lo;
		s_edge[e] = (uint32_t)(dx * L.wp + dy * L.zc) | ((uint32_t)dz << 16) | ((uint32_t)axis << 20);
	}
	for (;;)
	{
		__syncthreads(); // the previous chunk's last phase still reads shared memory
		// ---- next chunk of the work list (k_scan_chunks has ordered it largest first: longest-processing-time order keeps the launch's tail short)
		if (tid == 0)
		{
			const u64 j = atomicAdd(A.tot + TOT_TICKET, 1ull);
			s_chunk = j < A.tot[TOT_MESH] ? A.emit_list[j] : -1;
			s_njobs = 0; s_nvc = 0; s_nic = 0;
		}
		__syncthreads();
		const int chunk = s_chunk;
		if (chunk < 0) return;
		const ChunkCounts cc = A.chunks[chunk];
		const uint32_t V = cc.n_verts;
		const u64 cb = cc.cell_base, vb = cc.vert_base, ib = cc.ind_base;
		// use counters of this chunk's vertices, 16 bits each, in the chunk's own stretch of the global array (V / 2 of its V words)
		const uint32_t cap_aoff = (uint32_t)((d + 1) * L.wp + wc + (wc / 32) * 2); // words of the sign-word / offset region s_aoff takes over
		const bool fits = V <= (uint32_t)FUSED_CAPV && V <= cap_aoff;
		uint32_t* const cls32 = A.cls + vb; // accumulated with fire-and-forget global REDs (shared-memory atomics on the same words were slower: two
		                                    // vertices per word and neighbouring triangles in one warp collide); phase 5e copies them to shared memory
		BMF_FUSED_MARK(0);

		// ---- phase 1: sign words (+ a zero plane at x = d: B == 0 outside the grid) and word counts -> shared memory; words that emit
		// anything -> job list (one shared-memory atomic per warp and step)
		{
			const uint32_t* cnt = A.wcnt + (size_t)chunk * wc;
			if (TMA)
			{
				// both 4 * wc byte images with two bulk copies; the stores below are issued while they are in flight
				if (tid == 0)
				{
					mbar_expect(&s_bar, 8u * (uint32_t)wc);
					bulk_copy(sb, A.bits + (size_t)chunk * wc, 4u * (uint32_t)wc, &s_bar);
					bulk_copy(s_off, cnt, 4u * (uint32_t)wc, &s_bar);
				}
			}
			else
			{
				const uint4* src = reinterpret_cast<const uint4*>(A.bits + (size_t)chunk * wc);
				uint4* dst = reinterpret_cast<uint4*>(sb);
				for (int i = tid; i < wc / 4; i += NT) dst[i] = src[i];
			}
			for (int i = tid; i < L.wp; i += NT) sb[wc + i] = 0u;
			if (TMA)
			{
				for (uint32_t i = tid; i < (V + 1) / 2; i += NT) cls32[i] = 0u;
				for (uint32_t i = tid; i < 3 * V; i += NT) A.normal[3 * (size_t)vb + i] = 0.0f;
				mbar_wait(&s_bar, phase);
				phase ^= 1u;
				for (int w = tid; w < wc; w += NT)
				{
					const uint32_t c = s_off[w];
					const uint32_t bal = __ballot_sync(0xffffffffu, c != 0u);
					uint32_t o = 0;
					if (lane == 0 && bal) o = atomicAdd(&s_njobs, (uint32_t)__popc(bal));
					o = __shfl_sync(0xffffffffu, o, 0);
					if (c) s_job[o + __popc(bal & ((1u << lane) - 1u))] = (uint16_t)w;
				}
			}
			else
			{
			constexpr int U1 = 8; // word counts in flight per thread
			for (int w0 = tid; w0 < wc; w0 += U1 * NT)
			{
				uint32_t c[U1];
#pragma unroll
				for (int u = 0; u < U1; u++) c[u] = (w0 + u * NT < wc) ? __ldg(cnt + w0 + u * NT) : 0u;
#pragma unroll
				for (int u = 0; u < U1; u++)
				{
					const int w = w0 + u * NT;
					if (w >= wc) break; // uniform over the CTA: wc is a multiple of NT
					s_off[w] = c[u];
					const uint32_t bal = __ballot_sync(0xffffffffu, c[u] != 0u);
					uint32_t o = 0;
					if (lane == 0 && bal) o = atomicAdd(&s_njobs, (uint32_t)__popc(bal));
					o = __shfl_sync(0xffffffffu, o, 0);
					if (c[u]) s_job[o + __popc(bal & ((1u << lane) - 1u))] = (uint16_t)w;
				}
			}
			// every vertex's use counters (phase 5d adds to them) and normal start at zero: coalesced here instead of per vertex in 5c
			for (uint32_t i = tid; i < (V + 1) / 2; i += NT) cls32[i] = 0u;
			for (uint32_t i = tid; i < 3 * V; i += NT) A.normal[3 * (size_t)vb + i] = 0.0f;
			}
		}
		__syncthreads();
		BMF_FUSED_MARK(1);

		// ---- phase 3: exclusive prefix of the packed counts in word order.  Warp w owns a contiguous range of words and scans it 32 words
		// (one group) at a time: a word keeps its 16 + 16 bit offset inside its group (a group holds <= 3072 vertices and <= 15360 indices),
		// the group keeps the 32-bit bases.  The per-group scans are independent; group, warp and CTA prefixes follow.
		{
			const int Sw = wc / NW, G = Sw >> 5; // words / groups per warp
			uint32_t gv = 0, gi = 0;           // lane g: totals of the warp's group g
			for (int g = 0; g < G; g++)
			{
				const int w = wid * Sw + (g << 5) + lane;
				const uint32_t own = s_off[w];
				uint32_t inc = own;
#pragma unroll
				for (int o = 1; o < 32; o <<= 1)
				{
					const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
					if (lane >= o) inc += u;
				}
				s_off[w] = inc - own;
				const uint32_t t = __shfl_sync(0xffffffffu, inc, 31);
				if (lane == g) { gv = t & 0xFFFFu; gi = t >> 16; }
			}
			uint32_t iv = gv, ii = gi;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1)
			{
				const uint32_t a = __shfl_up_sync(0xffffffffu, iv, o), b = __shfl_up_sync(0xffffffffu, ii, o);
				if (lane >= o) { iv += a; ii += b; }
			}
			if (lane == 31) s_wtot[wid] = make_uint2(iv, ii);
			__syncthreads();
			if (wid == 0)
			{
				const uint2 v = lane < NW ? s_wtot[lane] : make_uint2(0u, 0u);
				uint32_t jv = v.x, ji = v.y;
#pragma unroll
				for (int o = 1; o < 32; o <<= 1)
				{
					const uint32_t a = __shfl_up_sync(0xffffffffu, jv, o), b = __shfl_up_sync(0xffffffffu, ji, o);
					if (lane >= o) { jv += a; ji += b; }
				}
				if (lane < NW) s_woff[lane] = make_uint2(jv - v.x, ji - v.y);
			}
			__syncthreads();
			if (lane < G)
			{
				const uint2 wo = s_woff[wid];
				s_gb[wid * G + lane] = make_uint2(wo.x + iv - gv, wo.y + ii - gi);
			}
		}
		__syncthreads();
		BMF_FUSED_MARK(2);
		const uint32_t njobs = s_njobs;

		// ---- phase 5b: compact lists of the cells that own vertices / that polygonize (order irrelevant: every record carries
		// its own output position), appended with one shared-memory atomic per word
		uint2* vl = A.vcells + cb;
		uint2* il = A.icells + cb;
		for (uint32_t q = tid; q < njobs; q += NT)
		{
			const int w = s_job[q];
			const int zb = w & (L.zc - 1), y = (w >> L.lzc) & (d - 1), x = w >> L.lwp;
			const WordBits b = load_word_bits(sb, L, x, y, zb);
			const WordClass c = classify(b, L, x, y, zb);
			uint32_t vbase, ibase;
			word_bases(S, w, vbase, ibase);
			uint32_t m = c.ex | c.ey | c.ez;
			if (m)
			{
				uint32_t o = atomicAdd(&s_nvc, (uint32_t)__popc(m)), rank = 0;
				while (m)
				{
					const int bit = __ffs(m) - 1;
					m &= m - 1;
					const uint32_t fl = ((c.ex >> bit) & 1u) | (((c.ey >> bit) & 1u) << 1) | (((c.ez >> bit) & 1u) << 2);
					vl[o++] = make_uint2((uint32_t)w | ((uint32_t)bit << 13) | (fl << 18), vbase + rank);
					rank += __popc(fl);
				}
			}
			m = c.active & c.interior;
			if (m)
			{
				uint32_t o = atomicAdd(&s_nic, (uint32_t)__popc(m)), ofs = 0;
				while (m)
				{
					const int bit = __ffs(m) - 1;
					m &= m - 1;
					const uint32_t m8 = mask8_of(b, bit);
					il[o++] = make_uint2((uint32_t)w | ((uint32_t)bit << 13) | (m8 << 18), ibase + ofs);
					ofs += (uint32_t)(s_tri[m8] >> 60);
				}
			}
		}
		__syncthreads();
		BMF_FUSED_MARK(3);
		const uint32_t nvc = s_nvc, nic = s_nic;

		// ---- phase 5c: iso-vertices, one thread per vertex-owning cell (calculate_isovertex, DMCChunk.cpp:657-674: X, Y, Z edge).  The cell's
		// own sample and its three possible neighbour samples are fetched together (one memory latency instead of up to four), the record of
		// the next round is in flight meanwhile.  Same expressions as density_at(), so the same bits.
		{
			const ChunkGeom g = A.geom[chunk];
			const float* dens = A.src.density ? A.src.density + (size_t)chunk * d * d * d : nullptr;
			const float* hm = (!dens && A.src.hmap) ? A.src.hmap + (size_t)A.src.sheet_of[chunk] * d * d : nullptr;
			uint2 vrec_nx = make_uint2(0u, 0u);
			if ((uint32_t)tid < nvc) vrec_nx = __ldcg(vl + tid);
			for (uint32_t i = tid; i < nvc; i += NT)
			{
				const uint2 rec = vrec_nx;
				if (i + NT < nvc) vrec_nx = __ldcg(vl + i + NT);
				const int w = rec.x & 0x1FFF, bit = (rec.x >> 13) & 31, fl = (rec.x >> 18) & 7;
				const int zb = w & (L.zc - 1), y = (w >> L.lzc) & (d - 1), x = w >> L.lwp;
				const int z = zb * 32 + bit;
				size_t v = (size_t)vb + rec.y;
				float s0, s1[3];
				if (dens)
				{
					const int xn = min(x + 1, d - 1), yn = min(y + 1, d - 1), zn = min(z + 1, d - 1); // a clamped neighbour is never used (no edge there)
					s0 = dens[((size_t)x * d + y) * d + z];
					s1[0] = dens[((size_t)xn * d + y) * d + z];
					s1[1] = dens[((size_t)x * d + yn) * d + z];
					s1[2] = dens[((size_t)x * d + y) * d + zn];
				}
				else if (hm)
				{
					const int xn = min(x + 1, d - 1), zn = min(z + 1, d - 1);
					const float h0 = hm[(size_t)x * d + z], hx = hm[(size_t)xn * d + z], hz = hm[(size_t)x * d + zn];
					s0 = terrain_density(A.s, g, y, h0);
					s1[0] = terrain_density(A.s, g, y, hx);
					s1[1] = terrain_density(A.s, g, y + 1, h0);
					s1[2] = terrain_density(A.s, g, y, hz);
				}
				else
				{
					s0 = implicit_point(A.s, g, x, y, z);
					s1[0] = (fl & 1) ? implicit_point(A.s, g, x + 1, y, z) : 0.0f;
					s1[1] = (fl & 2) ? implicit_point(A.s, g, x, y + 1, z) : 0.0f;
					s1[2] = (fl & 4) ? implicit_point(A.s, g, x, y, z + 1) : 0.0f;
				}
				const bool b0 = x == 0 || y == 0 || z == 0 || x == d - 1 || y == d - 1 || z == d - 1;
#pragma unroll
				for (int axis = 0; axis < 3; axis++)
				{
					if (!((fl >> axis) & 1)) continue;
					const int x1 = x + (axis == 0), y1 = y + (axis == 1), z1 = z + (axis == 2);
					const float mu = (0.0f - s0) / (s1[axis] - s0);
					A.pos[3 * v + 0] = ((float)x1 - (float)x) * mu + (float)x;
					A.pos[3 * v + 1] = ((float)y1 - (float)y) * mu + (float)y;
					A.pos[3 * v + 2] = ((float)z1 - (float)z) * mu + (float)z;
					A.boundary[v] = (b0 || x1 == d - 1 || y1 == d - 1 || z1 == d - 1) ? 1 : 0;
					v++;
				}
			}
		}
		// region R changes hands: job list -> flattening tables
		__syncthreads();
		BMF_FUSED_MARK(4);

		uint32_t* pre = s_pre_all + wid * 40;
		uint8_t* own = s_own_all + wid * 480;
		const int wbase = tid & ~31;
		// ---- phase 5d: indices (polygonize_cell, DMCChunk.cpp:537-576) + per-class use counts.  A warp takes 32 polygonizing cells and
		// spreads their (cell, index) pairs over its lanes; the vertex id of an index is looked up in shared memory.
		uint32_t* const inds_c = A.inds + ib;
		uint2 rec_nx = make_uint2(0u, 0u); // the record of the NEXT round is fetched while this round's pairs are processed
		if ((uint32_t)tid < nic) rec_nx = __ldcg(il + tid);
		for (uint32_t i0 = (uint32_t)wbase; i0 < nic; i0 += NT)
		{
			const uint32_t i = i0 + lane;
			const uint2 rec = rec_nx;
			if (i + NT < nic) rec_nx = __ldcg(il + i + NT);
			uint32_t n = 0;
			if (i < nic)
			{
				const u64 tp = s_tri[(rec.x >> 18) & 0xFF];
				n = (uint32_t)(tp >> 60);
				s_tp[tid] = tp;
				s_xyz[tid] = rec.x & 0x3FFFFu; // word | bit << 13
				s_out[tid] = rec.y;
			}
			// the unit of work is a TRIANGLE of a cell (three consecutive indices): the per-unit bookkeeping is paid once per three indices
			n = (n * 11u) >> 5; // triangles of the cell (n / 3 for n <= 15)
			uint32_t inc = n;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1)
			{
				const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
				if (lane >= o) inc += u;
			}
			pre[lane + 1] = inc;
			if (lane == 0) pre[0] = 0;
			for (uint32_t q = inc - n; q < inc; q++) own[q] = (uint8_t)lane;
			__syncwarp();
			const uint32_t n_tris = pre[32];
			for (uint32_t p = lane; p < n_tris; p += 32)
			{
				const int c = own[p];
				const uint32_t tt = p - pre[c];
				const uint32_t e3 = (uint32_t)(s_tp[wbase + c] >> (12 * tt)) & 0xFFFu; // the triangle's three edge ids
				const uint32_t wb = s_xyz[wbase + c];
				uint32_t* const out = inds_c + s_out[wbase + c] + 3u * tt;
#pragma unroll
				for (int q = 0; q < 3; q++)
				{
					const uint32_t e = (e3 >> (4 * q)) & 15u, ed = s_edge[e];
					int w = (int)(wb & 0x1FFFu) + (int)(ed & 0xFFFFu), bit = (int)(wb >> 13) + (int)((ed >> 16) & 1u);
					if (bit == 32) { w++; bit = 0; } // the z + 1 neighbour of the word's last cell
					const uint32_t vid = vertex_id_smem(S, L, w, bit, (int)(ed >> 20));
					out[q] = vid;
					// init_valence++ (DMCChunk.cpp:573) per "cell class" 3 - (e & 3): see k_inds3 / k_adj_fill
					atomicAdd(cls32 + (vid >> 1), 1u << (4 * (3 - (e & 3)) + 16 * (vid & 1)));
				}
			}
			__syncwarp();
		}
		__syncthreads();
		BMF_FUSED_MARK(5);

		// ---- phase 5e: init_valence and adj_offset = exclusive prefix of init_valence (MeshProcessor.cpp:33-39); the valences of a
		// chunk add up to its index count, so the batch-wide prefix at the chunk's first vertex is its ind_base
		{
			uint32_t carry = (uint32_t)ib;
			for (uint32_t base = 0; base < V; base += NT * VAL_ITEMS)
			{
				uint32_t val[VAL_ITEMS], sc[VAL_ITEMS], rt[VAL_ITEMS];
#pragma unroll
				for (int k = 0; k < VAL_ITEMS; k++)
				{
					const uint32_t i = base + k * NT + tid;
					uint32_t w = 0;
					if (i < V) w = (__ldcg(cls32 + (i >> 1)) >> (16 * (i & 1))) & 0xFFFFu;
					if (fits && i < V) reinterpret_cast<uint16_t*>(s_cls)[i] = (uint16_t)w;
					val[k] = (w & 15u) + ((w >> 4) & 15u) + ((w >> 8) & 15u) + (w >> 12);
					sc[k] = val[k];
				}
				block_scan_nt<NT, VAL_ITEMS>(sc, rt, s_scan, s_scant);
				uint32_t run = carry;
#pragma unroll
				for (int k = 0; k < VAL_ITEMS; k++)
				{
					const uint32_t i = base + k * NT + tid;
					if (i < V)
					{
						A.valence[(size_t)vb + i] = (uint8_t)val[k];
						A.adj_off[(size_t)vb + i] = run + sc[k];
						if (fits) s_aoff[i] = run + sc[k];
					}
					run += rt[k];
				}
				carry = run;
			}
		}
		if (!A.adj) continue;
		__syncthreads();
		BMF_FUSED_MARK(6);

		// ---- phase 5f: CSR adjacency in ascending primitive order without sort or cursor atomics (see k_adj_fill).  Latency-bound (index ->
		// use counters / adjacency offset -> store, all through L2): UF pairs per lane are in flight at once.
		const uint32_t* const aoff_c = A.adj_off + vb;
		const uint32_t prim_c = (uint32_t)(ib / 3); // every cell emits whole triangles, so a chunk's and a cell's first index are multiples of 3
		for (uint32_t t = tid; t < cc.n_inds / 3u; t += NT) A.prim_vbase[prim_c + t] = (uint32_t)vb; // Primitive -> its chunk's first vertex
		rec_nx = make_uint2(0u, 0u);
		if ((uint32_t)tid < nic) rec_nx = __ldcg(il + tid);
		for (uint32_t i0 = (uint32_t)wbase; i0 < nic; i0 += NT)
		{
			const uint32_t i = i0 + lane;
			const uint2 rec = rec_nx;
			if (i + NT < nic) rec_nx = __ldcg(il + i + NT);
			uint32_t n = 0;
			if (i < nic)
			{
				const uint32_t m8 = (rec.x >> 18) & 0xFF;
				const u64 tp = s_tri[m8];
				n = (uint32_t)(tp >> 60);
				s_tp[tid] = tp;
				s_lc[tid] = s_loc[m8];
				s_out[tid] = rec.y / 3u; // the cell's first primitive inside the chunk
			}
			n = (n * 11u) >> 5; // triangles of the cell
			uint32_t inc = n;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1)
			{
				const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
				if (lane >= o) inc += u;
			}
			pre[lane + 1] = inc;
			if (lane == 0) pre[0] = 0;
			for (uint32_t q = inc - n; q < inc; q++) own[q] = (uint8_t)lane;
			__syncwarp();
			const uint32_t n_tris = pre[32];
			constexpr int UF = 2; // triangles per lane in flight (six index -> counter / offset -> store chains)
			for (uint32_t p0 = lane; p0 < n_tris; p0 += 32 * UF)
			{
				uint32_t prim[UF], e3[UF], lc3[UF], idx[UF][3], cw[UF][3], ao[UF][3];
#pragma unroll
				for (int u = 0; u < UF; u++)
				{
					const uint32_t p = p0 + 32 * u;
					prim[u] = 0; e3[u] = 0; lc3[u] = 0;
					idx[u][0] = idx[u][1] = idx[u][2] = 0;
					if (p < n_tris)
					{
						const int c = own[p];
						const uint32_t tt = p - pre[c];
						e3[u] = (uint32_t)(s_tp[wbase + c] >> (12 * tt)) & 0xFFFu;  // the triangle's three edge ids
						lc3[u] = (uint32_t)(s_lc[wbase + c] >> (9 * tt)) & 0x1FFu; // their ranks among the same edge's uses inside the cell
						const uint32_t lp = s_out[wbase + c] + tt;                // primitive inside the chunk
						prim[u] = prim_c + lp;
						const uint32_t* src = inds_c + 3u * lp;
						idx[u][0] = __ldcg(src); idx[u][1] = __ldcg(src + 1); idx[u][2] = __ldcg(src + 2);
					}
				}
				if (fits)
				{
#pragma unroll
					for (int u = 0; u < UF; u++)
#pragma unroll
						for (int q = 0; q < 3; q++)
						{
							cw[u][q] = reinterpret_cast<const uint16_t*>(s_cls)[idx[u][q]];
							ao[u][q] = s_aoff[idx[u][q]];
						}
				}
				else
				{
#pragma unroll
					for (int u = 0; u < UF; u++)
#pragma unroll
						for (int q = 0; q < 3; q++)
						{
							cw[u][q] = __ldcg(cls32 + (idx[u][q] >> 1)) >> (16 * (idx[u][q] & 1));
							ao[u][q] = __ldcg(aoff_c + idx[u][q]);
						}
				}
#pragma unroll
				for (int u = 0; u < UF; u++)
				{
					if (p0 + 32 * u >= n_tris) continue;
#pragma unroll
					for (int q = 0; q < 3; q++)
					{
						const uint32_t cl = 3u - ((e3[u] >> (4 * q)) & 3u); // class of the cell among the (at most four) cells around the vertex's edge
						const uint32_t lower = cw[u][q] & ((1u << (4 * cl)) - 1u);
						const uint32_t before = (lower & 15u) + ((lower >> 4) & 15u) + ((lower >> 8) & 15u);
						A.adj[ao[u][q] + before + ((lc3[u] >> (3 * q)) & 7u)] = prim[u];
					}
				}
			}
			__syncwarp();
		}
		__syncthreads();
		BMF_FUSED_MARK(7);
	}
}

} // namespace bmf

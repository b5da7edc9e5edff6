// fused.cuh -- ONE launch for label_edges + polygonize + MeshProcessor::init of a whole batch (dim <= 64, triangles):
// one CTA per mesh-containing chunk, the chunk's sign field resident in shared memory from the first classification to the last
// adjacency entry.  It replaces k_count -> k_scan_chunks -> k_check_caps -> k_bases -> k_verts3 -> k_inds3 -> k_valence_offsets ->
// k_adj_fill (eight launches, three classifications of every word and the per-word count / base / record arrays round-tripping
// through HBM) for the chunk sizes the reference's worlds use (chunk_resolution 32 / 64, WorldOctree.cpp:20-33).
//
// What stays exactly as in the multi-kernel path -- and therefore bit-identical to the reference's serial scan
// (DMCChunk.cpp:440-498 cell order, :514-576 index order, MeshProcessor.cpp:98-128 adjacency order):
//   * vertex ids / index positions are the exclusive prefix, in x -> y -> z word order, of the per-word vertex / index counts;
//   * chunk bases are the exclusive prefix of the chunk totals in BATCH order.
// How it is computed here:
//   * the prefix inside a chunk is a warp-sequential scan over the 8192 (64^3) or 1024 (32^3) per-word counts in shared memory
//     (a 16-word group base + a 16-bit offset per word: 36 KB instead of 64 KB, so two CTAs fit one SM);
//   * the prefix ACROSS chunks is a decoupled look-back over the compacted list of mesh chunks (k_mesh_list): CTAs take list
//     positions from an atomic ticket, so every predecessor of a waiting CTA is running or done -- no deadlock, no second pass;
//   * a vertex id anywhere in the chunk is (group base + word offset + three popcounts) out of shared memory -- five LDS instead
//     of a 16-byte record fetched from L2 -- so index emission needs neither the record array nor the per-edge staging table;
//   * the cell lists, use counters and adjacency offsets of a chunk are written and read back by the same CTA: they stay in L2.
// Arena capacity is checked per chunk against the bases the look-back delivers; the CTA of the last mesh chunk publishes the
// totals and the "too small" flag (host + device), exactly what k_check_caps did.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "extract.cuh"
#include "download.cuh"

namespace bmf
{

enum { TOT_MESH = 10 /* number of mesh-containing chunks (k_mesh_list) */, TOT_TICKET = 11 /* next list position (k_chunk_mesh) */ };

typedef unsigned long long u64;
static constexpr int FUSED_NT = 512; // threads per chunk CTA: two CTAs per SM at 64^3 (89 KB of shared memory each)

// ---- ordered compaction of the chunks that contain a mesh; resets the per-batch device state of the fused path
__global__ void __launch_bounds__(SCAN_CTA) k_mesh_list(const uint32_t* __restrict__ flags, int n, int* __restrict__ mesh_list, unsigned int* __restrict__ status,
                                                         ChunkCounts* __restrict__ chunks, uint32_t* __restrict__ host_table, u64* __restrict__ tot, u64* __restrict__ tot_host)
{
	__shared__ uint32_t s_w[SCAN_CTA / 32];
	__shared__ uint32_t s_tot;
	const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
	uint32_t carry = 0;
	for (int base = 0; base < n; base += SCAN_CTA)
	{
		const int i = base + t;
		const bool f = i < n && flags_contain_mesh(flags[i]);
		const uint32_t bal = __ballot_sync(0xffffffffu, f);
		if (lane == 0) s_w[warp] = __popc(bal);
		__syncthreads();
		if (warp == 0)
		{
			const uint32_t v = s_w[lane];
			uint32_t inc = v;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1)
			{
				const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
				if (lane >= o) inc += u;
			}
			s_w[lane] = inc - v;
			if (lane == 31) s_tot = inc;
		}
		__syncthreads();
		if (f)
		{
			const uint32_t p = carry + s_w[warp] + __popc(bal & ((1u << lane) - 1u));
			mesh_list[p] = i;
			status[p] = 0;
		}
		carry += s_tot;
		__syncthreads();
	}
	if (t == 0)
	{
		for (int k = 0; k < TOT_SLOTS; k++) tot[k] = 0;
		tot[TOT_MESH] = carry;
	}
	if (carry == 0)
	{
		// no chunk contains a mesh: the chunk table and the totals are all zero, and k_chunk_mesh has nothing to do
		constexpr int REC = (int)(sizeof(ChunkCounts) / sizeof(uint32_t));
		for (size_t k = t; k < (size_t)n * REC; k += SCAN_CTA)
		{
			reinterpret_cast<uint32_t*>(chunks)[k] = 0;
			if (host_table) host_table[k] = 0;
		}
		if (t == 0)
		{
			for (int k = 0; k < TOT_SLOTS; k++) tot_host[k] = 0;
			__threadfence_system();
		}
	}
}

struct FusedArgs
{
	Layout L;
	int n;
	const uint32_t* bits;
	const uint32_t* flags;
	const int* mesh_list;
	ChunkCounts* chunks;
	uint32_t* host_table; // mapped pinned copy of the chunk table, or null (already published for this batch)
	unsigned int* status; // look-back state per list position: 0 nothing, 1 aggregate, 2 inclusive prefix
	u64* agg;             // [n][3] (cells, verts, indices) of the chunk at a list position
	u64* inc;             // [n][3] inclusive prefix up to and including it
	u64* tot;
	u64* tot_host;
	u64 cap_cells, cap_verts, cap_inds;
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
	uint8_t* masks;       // MasksBlock byte image (keep_masks) or null
	u64* prof;            // debugging aid (BMF_FUSED_PROF=1): [list position][16] SM clock at the phase boundaries, thread 0; null otherwise
};

#define BMF_FUSED_MARK(k) do { if (A.prof && tid == 0) A.prof[(size_t)j * 16 + (k)] = (u64)clock64(); } while (0)

// dynamic shared memory of k_chunk_mesh<NT> for a layout (bytes)
__host__ __device__ inline size_t fused_smem_bytes(const Layout& L, int nt)
{
	const size_t scratch = (size_t)36 * nt + 256, jobs = (size_t)2 * L.wc;
	return ((size_t)(L.d + 1) * L.wp + L.wc) * 4 + (size_t)(L.wc / 16) * 8 + (scratch > jobs ? scratch : jobs);
}

__device__ __forceinline__ u64 ld_volatile_u64(const u64* p) { return *reinterpret_cast<const volatile u64*>(p); }

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
	const u64* gb;
};

__device__ __forceinline__ void word_bases(const FusedView& S, int w, uint32_t& vbase, uint32_t& ibase)
{
	const u64 g = S.gb[w >> 4];
	const uint32_t o = S.off[w];
	vbase = (uint32_t)g + (o & 0xFFFFu);
	ibase = (uint32_t)(g >> 32) + (o >> 16);
}

// chunk-local id of the vertex on `axis` of cell (x,y,z), everything out of shared memory (vertex_id_rec's arithmetic)
__device__ __forceinline__ uint32_t vertex_id_smem(const FusedView& S, const Layout& L, int x, int y, int z, int axis)
{
	const int zb = z >> 5, bit = z & 31;
	const int w = (((x << L.ld) + y) << L.lzc) + zb;
	const uint32_t A = S.sb[w];
	const uint32_t ex = (x + 1 < L.d) ? (A ^ S.sb[w + L.wp]) : 0u;
	const uint32_t ey = (y + 1 < L.d) ? (A ^ S.sb[w + L.zc]) : 0u;
	const bool zn = zb + 1 < L.zc;
	const uint32_t A1 = __funnelshift_r(A, zn ? S.sb[w + 1] : 0u, 1);
	const uint32_t ez = (A ^ A1) & (zn ? 0xFFFFFFFFu : 0x7FFFFFFFu);
	const uint32_t lt = (1u << bit) - 1u;
	const u64 g = S.gb[w >> 4];
	uint32_t id = (uint32_t)g + (S.off[w] & 0xFFFFu) + __popc(ex & lt) + __popc(ey & lt) + __popc(ez & lt);
	if (axis >= 1) id += (ex >> bit) & 1u;
	if (axis >= 2) id += (ey >> bit) & 1u;
	return id;
}

template <int NT>
__global__ void __launch_bounds__(NT, (NT <= 512 ? 2 : 1)) k_chunk_mesh(FusedArgs A)
{
	extern __shared__ __align__(16) uint32_t dyn[];
	constexpr int NW = NT / 32;
	const Layout L = A.L;
	const int d = L.d, wc = L.wc, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
	uint32_t* sb = dyn;
	uint32_t* s_off = sb + (d + 1) * L.wp;
	u64* s_gb = reinterpret_cast<u64*>(s_off + wc);
	uint8_t* s_r = reinterpret_cast<uint8_t*>(s_gb + wc / 16);
	// region R, phase 2 / 5b: the list of words with active cells
	uint16_t* s_job = reinterpret_cast<uint16_t*>(s_r);
	// region R, phases 5d / 5f: per-thread cell slots and per-warp flattening tables
	u64* s_tp = reinterpret_cast<u64*>(s_r);                                    // [NT]
	uint32_t* s_xyz = reinterpret_cast<uint32_t*>(s_tp + NT);                   // [NT]
	uint32_t* s_out = s_xyz + NT;                                               // [NT]
	uint32_t* s_pre_all = s_out + NT;                                           // [NW][40]
	uint8_t* s_own_all = reinterpret_cast<uint8_t*>(s_pre_all + NW * 40);       // [NW][480]
	__shared__ u64 s_tri[256];
	__shared__ int s_ticket;
	__shared__ uint32_t s_cells, s_njobs, s_nvc, s_nic, s_fits;
	__shared__ u64 s_wtot[NW], s_woff[NW];
	__shared__ u64 s_base[3];
	__shared__ uint32_t s_scan[VAL_ITEMS][NW], s_scant[VAL_ITEMS];
	FusedView S;
	S.sb = sb; S.off = s_off; S.gb = s_gb;

	for (int i = tid; i < 256; i += NT) s_tri[i] = c_tri_pack[i];
	const int M = (int)A.tot[TOT_MESH];
	constexpr int REC = (int)(sizeof(ChunkCounts) / sizeof(uint32_t));
	for (;;)
	{
		__syncthreads(); // the previous chunk's last phase still reads shared memory
		if (tid == 0)
		{
			s_ticket = (int)atomicAdd(A.tot + TOT_TICKET, 1ull);
			s_cells = 0; s_njobs = 0; s_nvc = 0; s_nic = 0;
		}
		__syncthreads();
		const int j = s_ticket;
		if (j >= M) return;
		const int chunk = A.mesh_list[j];
		BMF_FUSED_MARK(0);

		// ---- phase 1: the chunk's sign words -> shared memory (+ a zero plane at x = d: B == 0 outside the grid)
		{
			const uint4* src = reinterpret_cast<const uint4*>(A.bits + (size_t)chunk * wc);
			uint4* dst = reinterpret_cast<uint4*>(sb);
			for (int i = tid; i < wc / 4; i += NT) dst[i] = src[i];
			for (int i = tid; i < L.wp; i += NT) sb[wc + i] = 0u;
		}
		__syncthreads();
		BMF_FUSED_MARK(1);

		// ---- phase 2: classify every word once: (vertices, indices) of the word -> s_off, words with active cells -> job list
		{
			uint32_t cells = 0;
			for (int w = tid; w < wc; w += NT)
			{
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
					s_job[atomicAdd(&s_njobs, 1u)] = (uint16_t)w;
				}
				s_off[w] = nv | (ni << 16);
				if (A.masks)
				{
					uint8_t* mrow = A.masks + (size_t)chunk * wc * 32 + ((size_t)x * d + y) * d + zb * 32;
#pragma unroll
					for (int q = 0; q < 8; q++)
						reinterpret_cast<uint32_t*>(mrow)[q] = mask8_of(b, 4 * q) | (mask8_of(b, 4 * q + 1) << 8) | (mask8_of(b, 4 * q + 2) << 16) | (mask8_of(b, 4 * q + 3) << 24);
				}
			}
#pragma unroll
			for (int o = 16; o >= 1; o >>= 1) cells += __shfl_xor_sync(0xffffffffu, cells, o);
			if (lane == 0 && cells) atomicAdd(&s_cells, cells);
		}
		__syncthreads();
		BMF_FUSED_MARK(2);

		// ---- phase 3: exclusive prefix of the packed counts in word order.  Warp w scans its contiguous range of words 32 at a
		// time with a running carry; a word keeps a 16-bit offset inside its 16-word group, the group keeps the 64-bit base.
		{
			const int Sw = wc / NW; // words per warp (a multiple of 32)
			u64 carry = 0;
			for (int st = 0; st < Sw; st += 32)
			{
				const int w = wid * Sw + st + lane;
				const uint32_t cnt = s_off[w];
				const u64 own = (u64)(cnt & 0xFFFFu) | ((u64)(cnt >> 16) << 32);
				u64 inc = own;
#pragma unroll
				for (int o = 1; o < 32; o <<= 1)
				{
					const u64 u = __shfl_up_sync(0xffffffffu, inc, o);
					if (lane >= o) inc += u;
				}
				const u64 ex = inc - own;
				const u64 gex = __shfl_sync(0xffffffffu, ex, lane & 16);
				const u64 off = ex - gex;
				s_off[w] = (uint32_t)off | ((uint32_t)(off >> 32) << 16);
				if ((lane & 15) == 0) s_gb[w >> 4] = carry + gex;
				carry += __shfl_sync(0xffffffffu, inc, 31);
			}
			if (lane == 0) s_wtot[wid] = carry;
		}
		__syncthreads();
		BMF_FUSED_MARK(3);
		if (wid == 0)
		{
			const u64 v = lane < NW ? s_wtot[lane] : 0ull;
			u64 inc = v;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1)
			{
				const u64 u = __shfl_up_sync(0xffffffffu, inc, o);
				if (lane >= o) inc += u;
			}
			if (lane < NW) s_woff[lane] = inc - v;
			const u64 total = __shfl_sync(0xffffffffu, inc, 31);
			const u64 V = (uint32_t)total, I = total >> 32, C = s_cells;

			// ---- phase 4: this chunk's bases = exclusive prefix over the mesh chunks before it (decoupled look-back, warp 0)
			u64 e0 = 0, e1 = 0, e2 = 0;
			if (j > 0)
			{
				if (lane == 0)
				{
					A.agg[3 * (size_t)j] = C; A.agg[3 * (size_t)j + 1] = V; A.agg[3 * (size_t)j + 2] = I;
					__threadfence();
					*reinterpret_cast<volatile unsigned int*>(A.status + j) = 1u;
				}
				int p = j - 1;
				for (;;)
				{
					const int idx = p - lane;
					unsigned int st = 2u; // positions before the list start: an inclusive prefix of zero
					if (idx >= 0)
						do { st = *reinterpret_cast<const volatile unsigned int*>(A.status + idx); } while (st == 0u);
					__threadfence();
					const uint32_t pm = __ballot_sync(0xffffffffu, st == 2u);
					const int first = pm ? __ffs(pm) - 1 : 32;
					u64 a0 = 0, a1 = 0, a2 = 0;
					if (idx >= 0 && lane <= first)
					{
						const u64* src = (lane == first ? A.inc : A.agg) + 3 * (size_t)idx;
						a0 = ld_volatile_u64(src); a1 = ld_volatile_u64(src + 1); a2 = ld_volatile_u64(src + 2);
					}
#pragma unroll
					for (int o = 16; o >= 1; o >>= 1)
					{
						a0 += __shfl_xor_sync(0xffffffffu, a0, o);
						a1 += __shfl_xor_sync(0xffffffffu, a1, o);
						a2 += __shfl_xor_sync(0xffffffffu, a2, o);
					}
					e0 += a0; e1 += a1; e2 += a2;
					if (pm) break;
					p -= 32;
				}
			}
			if (lane == 0)
			{
				A.inc[3 * (size_t)j] = e0 + C; A.inc[3 * (size_t)j + 1] = e1 + V; A.inc[3 * (size_t)j + 2] = e2 + I;
				__threadfence();
				*reinterpret_cast<volatile unsigned int*>(A.status + j) = 2u;
				s_base[0] = e0; s_base[1] = e1; s_base[2] = e2;
				const u64 t0 = e0 + C, t1 = e1 + V, t2 = e2 + I;
				const bool overflow = t0 >= 0xFFFFFFFFull || t1 >= 0xFFFFFFFFull || t2 >= 0xFFFFFFFFull;
				s_fits = (!overflow && t0 <= A.cap_cells && t1 <= A.cap_verts && t2 <= A.cap_inds) ? 1u : 0u;
				if (V) atomicMax(A.tot + TOT_MAXV, V);
				if (j == M - 1)
				{
					// the last mesh chunk: the batch totals and the verdict on the arenas (what k_scan_chunks + k_check_caps published)
					const u64 small = s_fits ? 0ull : 1ull;
					A.tot[TOT_CELLS] = t0; A.tot[TOT_VERTS] = t1; A.tot[TOT_INDS] = t2;
					A.tot[TOT_OVERFLOW] = overflow ? 1ull : 0ull;
					A.tot[TOT_SMALL] = small;
					A.tot_host[TOT_CELLS] = t0; A.tot_host[TOT_VERTS] = t1; A.tot_host[TOT_INDS] = t2;
					A.tot_host[TOT_OVERFLOW] = overflow ? 1ull : 0ull;
					A.tot_host[TOT_LIST0] = 0; A.tot_host[TOT_LIST1] = 0; A.tot_host[TOT_WORK] = 0;
					A.tot_host[TOT_SMALL] = small;
					A.tot_host[TOT_DLERR] = 0;
					__threadfence_system();
				}
			}
			BMF_FUSED_MARK(4);
			// the chunk table: this chunk's record and those of the meshless chunks up to the next mesh chunk (and, for the first
			// list position, of the meshless chunks before it), whose bases are simply the running prefix
			{
				const u64 C1 = e0 + C, V1 = e1 + V, I1 = e2 + I;
				const int next = (j + 1 < M) ? A.mesh_list[j + 1] : A.n;
				const int first_rec = (j == 0) ? 0 : chunk;
				for (int c = first_rec + lane; c < next; c += 32)
				{
					ChunkCounts cc;
					const bool me = c == chunk, before = c < chunk;
					cc.contains_mesh = me ? 1u : 0u;
					cc.n_cells = me ? (uint32_t)C : 0u; cc.n_verts = me ? (uint32_t)V : 0u; cc.n_inds = me ? (uint32_t)I : 0u;
					cc.cell_base = (me || before) ? e0 : C1; cc.vert_base = (me || before) ? e1 : V1; cc.ind_base = (me || before) ? e2 : I1;
					A.chunks[c] = cc;
					if (A.host_table)
					{
						const uint32_t* r = reinterpret_cast<const uint32_t*>(&cc);
#pragma unroll
						for (int q = 0; q < REC; q++) A.host_table[(size_t)c * REC + q] = r[q];
					}
				}
			}
		}
		__syncthreads();
		BMF_FUSED_MARK(5);
		for (int g = tid; g < wc / 16; g += NT) s_gb[g] += s_woff[(g * 16) / (wc / NW)];
		const u64 total = s_woff[NW - 1] + s_wtot[NW - 1];
		const uint32_t V = (uint32_t)total, I = (uint32_t)(total >> 32);
		const uint32_t njobs = s_njobs;
		const u64 cb = s_base[0], vb = s_base[1], ib = s_base[2];
		const bool fits = s_fits != 0;
		__syncthreads();
		if (!fits || V == 0) continue; // an arena is too small (the host grows it and re-launches) or nothing to emit

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
		BMF_FUSED_MARK(6);
		const uint32_t nvc = s_nvc, nic = s_nic;

		// ---- phase 5c: iso-vertices, one thread per vertex-owning cell (calculate_isovertex, DMCChunk.cpp:657-674: X, Y, Z edge)
		{
			const ChunkGeom g = A.geom[chunk];
			for (uint32_t i = tid; i < nvc; i += NT)
			{
				const uint2 rec = __ldcg(vl + i);
				const int w = rec.x & 0x1FFF, bit = (rec.x >> 13) & 31, fl = (rec.x >> 18) & 7;
				const int zb = w & (L.zc - 1), y = (w >> L.lzc) & (d - 1), x = w >> L.lwp;
				const int z = zb * 32 + bit;
				size_t v = (size_t)vb + rec.y;
				const float s0 = density_at(A.s, A.src, g, d, chunk, x, y, z);
				const bool b0 = x == 0 || y == 0 || z == 0 || x == d - 1 || y == d - 1 || z == d - 1;
#pragma unroll
				for (int axis = 0; axis < 3; axis++)
				{
					if (!((fl >> axis) & 1)) continue;
					const int x1 = x + (axis == 0), y1 = y + (axis == 1), z1 = z + (axis == 2);
					const float s1 = density_at(A.s, A.src, g, d, chunk, x1, y1, z1);
					const float mu = (0.0f - s0) / (s1 - s0);
					A.pos[3 * v + 0] = ((float)x1 - (float)x) * mu + (float)x;
					A.pos[3 * v + 1] = ((float)y1 - (float)y) * mu + (float)y;
					A.pos[3 * v + 2] = ((float)z1 - (float)z) * mu + (float)z;
					A.boundary[v] = (b0 || x1 == d - 1 || y1 == d - 1 || z1 == d - 1) ? 1 : 0;
					A.cls[v] = 0u;
					A.normal[3 * v + 0] = 0.0f; A.normal[3 * v + 1] = 0.0f; A.normal[3 * v + 2] = 0.0f;
					v++;
				}
			}
		}
		// region R changes hands: job list -> flattening tables
		__syncthreads();
		BMF_FUSED_MARK(7);

		uint32_t* pre = s_pre_all + wid * 40;
		uint8_t* own = s_own_all + wid * 480;
		const int wbase = tid & ~31;
		// ---- phase 5d: indices (polygonize_cell, DMCChunk.cpp:537-576) + per-class use counts.  A warp takes 32 polygonizing cells and
		// spreads their (cell, index) pairs over its lanes; the vertex id of an index is looked up in shared memory.
		for (uint32_t i0 = (uint32_t)wbase; i0 < nic; i0 += NT)
		{
			const uint32_t i = i0 + lane;
			uint32_t n = 0;
			if (i < nic)
			{
				const uint2 rec = __ldcg(il + i);
				const int w = rec.x & 0x1FFF, bit = (rec.x >> 13) & 31;
				const u64 tp = s_tri[(rec.x >> 18) & 0xFF];
				n = (uint32_t)(tp >> 60);
				s_tp[tid] = tp;
				s_xyz[tid] = (uint32_t)(w >> L.lwp) | ((uint32_t)((w >> L.lzc) & (d - 1)) << 10) | ((uint32_t)((w & (L.zc - 1)) * 32 + bit) << 20);
				s_out[tid] = rec.y;
			}
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
			const uint32_t n_pairs = pre[32];
			for (uint32_t p = lane; p < n_pairs; p += 32)
			{
				const int c = own[p];
				const uint32_t t = p - pre[c];
				const uint32_t e = (uint32_t)(s_tp[wbase + c] >> (4 * t)) & 15u;
				const uint32_t xyz = s_xyz[wbase + c];
				const int x = (int)(xyz & 1023u), y = (int)((xyz >> 10) & 1023u), z = (int)(xyz >> 20);
				// edge e of the cell (EDGE_V, DMCChunk.cpp:32, 543-565): X-edges 0-3 at (y+hi, z+lo), Y-edges 4-7 at (x+hi, z+lo), Z-edges 8-11 at (x+hi, y+lo)
				const int axis = e >> 2, hi = (e >> 1) & 1, lo = e & 1;
				const int dx = axis == 0 ? 0 : hi;
				const int dy = axis == 0 ? hi : (axis == 1 ? 0 : lo);
				const int dz = axis == 2 ? 0 : lo;
				const uint32_t vid = vertex_id_smem(S, L, x + dx, y + dy, z + dz, axis);
				A.inds[(size_t)ib + s_out[wbase + c] + t] = vid;
				// init_valence++ (DMCChunk.cpp:573) per "cell class" 3 - (e & 3): see k_inds3 / k_adj_fill
				atomicAdd(A.cls + (size_t)vb + vid, 1u << (8 * (3 - (e & 3))));
			}
			__syncwarp();
		}
		__syncthreads();
		BMF_FUSED_MARK(8);

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
					const uint32_t w = i < V ? __ldcg(A.cls + (size_t)vb + i) : 0u;
					val[k] = (w & 0xFF) + ((w >> 8) & 0xFF) + ((w >> 16) & 0xFF) + (w >> 24);
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
					}
					run += rt[k];
				}
				carry = run;
			}
		}
		if (!A.adj) continue;
		__syncthreads();
		BMF_FUSED_MARK(9);

		// ---- phase 5f: CSR adjacency in ascending primitive order without sort or cursor atomics (see k_adj_fill)
		for (uint32_t i0 = (uint32_t)wbase; i0 < nic; i0 += NT)
		{
			const uint32_t i = i0 + lane;
			uint32_t n = 0;
			if (i < nic)
			{
				const uint2 rec = __ldcg(il + i);
				const u64 tp = s_tri[(rec.x >> 18) & 0xFF];
				n = (uint32_t)(tp >> 60);
				s_tp[tid] = tp;
				s_out[tid] = rec.y;
			}
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
			const uint32_t n_pairs = pre[32];
			for (uint32_t p = lane; p < n_pairs; p += 32)
			{
				const int c = own[p];
				const uint32_t t = p - pre[c];
				const u64 tp = s_tp[wbase + c];
				const uint32_t e = (uint32_t)(tp >> (4 * t)) & 15u;
				const size_t out = (size_t)ib + s_out[wbase + c] + t;
				const uint32_t prim = (uint32_t)(out / 3);
				if ((t % 3) == 0) A.prim_vbase[prim] = (uint32_t)vb;
				u64 x = tp ^ (0x1111111111111111ull * e);
				x = (x | (x >> 1) | (x >> 2) | (x >> 3)) & 0x1111111111111111ull;
				const u64 below = t ? (~0ull >> (64 - 4 * t)) : 0ull;
				const uint32_t local = t - (uint32_t)__popcll(x & below);
				const size_t v = (size_t)vb + __ldcg(A.inds + out);
				const uint32_t cl = 3u - (e & 3u);
				const uint32_t lower = __ldcg(A.cls + v) & ((1u << (8 * cl)) - 1u);
				const uint32_t before = (lower & 0xFF) + ((lower >> 8) & 0xFF) + ((lower >> 16) & 0xFF);
				A.adj[__ldcg(A.adj_off + v) + before + local] = prim;
			}
			__syncwarp();
		}
		__syncthreads();
		BMF_FUSED_MARK(10);
	}
}

} // namespace bmf

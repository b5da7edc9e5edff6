// seam.cuh -- the seam pass between chunks of different (or equal) LOD: WorldStitcher::stitch_all / stitch_cell /
// stitch_indexes (WorldStitcher.cpp:26-49, 184-239, 491-572).
//
// What the reference intends (it is non-functional as committed): every chunk is a block of d^3 voxel
// nodes (DMCChunk::generate_octree, DMCChunk.cpp:699-781: node pos = chunk pos + xyz * size/dim, one sample per
// node); a DC-style cell/face/edge recursion over the world octree collects, around every common corner, the 8 leaf
// nodes that meet there; those 8 nodes form a DUAL CELL that is polygonised with the marching-cubes table from the
// nodes' positions and samples (stitch_indexes) into a non-indexed triangle soup.
//
// What this build defines (UNPINNED -- the reference produces nothing to compare with): the same dual cells, found
// without recursion.  With chunks sampled at voxel-node centres (overlap = -1/(2 dim): sample i of a chunk sits at
// pos + (i + 1/2) size/dim), a chunk's own mesh covers the dual cells whose 8 nodes are all its own; the seam pass
// covers every other dual cell -- one per lattice point P on the boundary shell of a chunk's (d+1)^3 voxel-corner
// lattice.  Its 8 nodes are the voxels (of whatever chunk and LOD) that contain P -+ eps in each octant; a coarse voxel
// may serve several octants (the degenerate hexahedra of Schaefer & Warren's dual grids).  P is emitted exactly once, by
// the lowest-index chunk among the FINEST chunks touching it; cells touching the outside of the world are skipped.
// Corner numbering, edge numbering and triangle table are the chunk mesher's (corner = 4 dx + 2 dy + dz), so the seam
// continues the chunks' triangulation across equal-LOD borders and closes the cracks across LOD changes.
//
// Signs come from the resident sign words (or the uniform-chunk flags of the 2-D terrains); densities are read or
// re-evaluated only for the corners of sign-changing cells.  Output order is deterministic: (chunk, shell point,
// table order) via count -> scan -> emit.
#pragma once
#include "extract.cuh"

namespace bmf
{

struct SeamChunk
{
	int32_t ox, oy, oz; // origin in slots (one slot = the extent of the finest chunk of the batch)
	int32_t lg;         // log2(extent in slots) = log2(voxel size in finest-voxel units)
};

struct SeamGrid
{
	int gx, gy, gz; // slots per axis
	int n;          // chunks
	int bpc;        // CTAs per chunk
	int npts;       // shell lattice points per chunk = 6 d^2 + 2
};

// t-th point of the boundary shell of the (d+1)^3 lattice: the two x faces, then the y faces without their x borders,
// then the z faces without their x and y borders
__host__ __device__ __forceinline__ void seam_shell_point(int t, int d, int& i, int& j, int& k)
{
	const int a = (d + 1) * (d + 1);
	if (t < 2 * a)
	{
		i = t < a ? 0 : d;
		if (t >= a) t -= a;
		j = t / (d + 1);
		k = t - j * (d + 1);
		return;
	}
	t -= 2 * a;
	const int b = (d - 1) * (d + 1);
	if (t < 2 * b)
	{
		j = t < b ? 0 : d;
		if (t >= b) t -= b;
		const int q = t / (d + 1);
		i = 1 + q;
		k = t - q * (d + 1);
		return;
	}
	t -= 2 * b;
	const int c = (d - 1) * (d - 1);
	k = t < c ? 0 : d;
	if (t >= c) t -= c;
	const int q = t / (d - 1);
	i = 1 + q;
	j = 1 + (t - q * (d - 1));
}

struct SeamArgs
{
	SeamGrid G;
	Layout L;
	const SeamChunk* chunks;
	const int32_t* slot_map; // [gx][gy][gz] chunk index or -1
	const uint32_t* bits;
	const uint8_t* uni;      // per-chunk uniform flag (2-D terrains without a density block) or null
	const int32_t* group;    // per-chunk group id or null
	int cross_group_only;    // 1: only cells whose nodes span more than one group
	const ChunkGeom* geom;
	SamplerDev s;
	DensitySource src;
};

// the dual cell at shell point t of chunk c: returns the number of triangles kept and (if tri) their 9 floats each
__device__ __forceinline__ int seam_cell(const SeamArgs& A, const uint64_t* __restrict__ s_tri, int c, int t, float* tri)
{
	const int d = A.L.d, ld = A.L.ld;
	int i, j, k;
	seam_shell_point(t, d, i, j, k);
	const SeamChunk me = A.chunks[c];
	const int Px = (me.ox << ld) + (i << me.lg), Py = (me.oy << ld) + (j << me.lg), Pz = (me.oz << ld) + (k << me.lg);
	int cm[8], vx[8], vy[8], vz[8];
	uint32_t mask = 0;
	int owner = c;
	bool same_group = true;
	const int g0 = A.group ? A.group[c] : 0;
#pragma unroll
	for (int o = 0; o < 8; o++)
	{
		const int qx = Px - 1 + (o >> 2), qy = Py - 1 + ((o >> 1) & 1), qz = Pz - 1 + (o & 1);
		if (qx < 0 || qy < 0 || qz < 0) return 0;
		const int sx = qx >> ld, sy = qy >> ld, sz = qz >> ld;
		if (sx >= A.G.gx || sy >= A.G.gy || sz >= A.G.gz) return 0;
		const int m = A.slot_map[((size_t)sx * A.G.gy + sy) * A.G.gz + sz];
		if (m < 0) return 0; // outside the world (or a leaf that is not in the batch)
		const SeamChunk cm_ = A.chunks[m];
		if (cm_.lg < me.lg) return 0; // a finer chunk touches P: it owns the cell
		if (cm_.lg == me.lg && m < owner) owner = m;
		if (A.group && A.group[m] != g0) same_group = false;
		cm[o] = m;
		vx[o] = (qx - (cm_.ox << ld)) >> cm_.lg;
		vy[o] = (qy - (cm_.oy << ld)) >> cm_.lg;
		vz[o] = (qz - (cm_.oz << ld)) >> cm_.lg;
		uint32_t bit;
		const uint8_t u = A.uni ? A.uni[m] : (uint8_t)0;
		if (u) bit = (u == 1) ? 1u : 0u;
		else bit = (A.bits[(size_t)m * A.L.wc + ((((size_t)vx[o] << ld) + vy[o]) << A.L.lzc) + (vz[o] >> 5)] >> (vz[o] & 31)) & 1u;
		mask |= bit << o;
	}
	if (owner != c || mask == 0 || mask == 255) return 0;
	if (A.cross_group_only && same_group) return 0;

	const uint64_t tp = s_tri[mask];
	const int n = (int)(tp >> 60);
	// node positions (the chunk's own sample coordinates, ImplicitSampler.hpp:24-30) and samples of the 8 corners
	float px[8], py[8], pz[8], sv[8];
#pragma unroll
	for (int o = 0; o < 8; o++)
	{
		const ChunkGeom g = A.geom[cm[o]];
		px[o] = g.ox + (float)vx[o] * g.delta;
		py[o] = g.oy + (float)vy[o] * g.delta;
		pz[o] = g.oz + (float)vz[o] * g.delta;
		sv[o] = density_at(A.s, A.src, g, d, cm[o], vx[o], vy[o], vz[o]);
	}
	int kept = 0;
	for (int q = 0; q < n; q += 3)
	{
		float v[9];
#pragma unroll
		for (int r = 0; r < 3; r++)
		{
			const int e = (int)(tp >> (4 * (q + r))) & 15;
			// edge e joins corners a < b (tools/gen_mc_tables.py edge_corners)
			const int axis = e >> 2, hi = (e >> 1) & 1, lo = e & 1;
			const int a = axis == 0 ? ((hi << 1) | lo) : axis == 1 ? ((hi << 2) | lo) : ((hi << 2) | (lo << 1));
			const int b = a | (axis == 0 ? 4 : axis == 1 ? 2 : 1);
			float ax = 0, ay = 0, az = 0, as = 0, bx = 0, by = 0, bz = 0, bs = 0;
#pragma unroll
			for (int o = 0; o < 8; o++)
			{
				if (o == a) { ax = px[o]; ay = py[o]; az = pz[o]; as = sv[o]; }
				if (o == b) { bx = px[o]; by = py[o]; bz = pz[o]; bs = sv[o]; }
			}
			// _get_intersection (WorldStitcher.cpp:476-481)
			const float mu = (0.0f - as) / (bs - as);
			v[3 * r + 0] = (bx - ax) * mu + ax;
			v[3 * r + 1] = (by - ay) * mu + ay;
			v[3 * r + 2] = (bz - az) * mu + az;
		}
		// "TODO: don't push degenerate triangles" (WorldStitcher.cpp:566): two corners at the same place
		const bool e01 = v[0] == v[3] && v[1] == v[4] && v[2] == v[5];
		const bool e12 = v[3] == v[6] && v[4] == v[7] && v[5] == v[8];
		const bool e02 = v[0] == v[6] && v[1] == v[7] && v[2] == v[8];
		if (e01 || e12 || e02) continue;
		if (tri)
		{
#pragma unroll
			for (int r = 0; r < 9; r++) tri[9 * kept + r] = v[r];
		}
		kept++;
	}
	return kept;
}

// pass 1: triangles per CTA and per chunk
__global__ void __launch_bounds__(CTA) k_seam_count(SeamArgs A, uint32_t* __restrict__ blk_cnt, uint32_t* __restrict__ chunk_cnt)
{
	__shared__ uint64_t s_tri[256];
	__shared__ uint32_t s_sum;
	s_tri[threadIdx.x] = c_tri_pack[threadIdx.x];
	if (threadIdx.x == 0) s_sum = 0;
	__syncthreads();
	const int c = blockIdx.x / A.G.bpc, b = blockIdx.x - c * A.G.bpc;
	const int t = b * CTA + threadIdx.x;
	uint32_t cnt = 0;
	if (t < A.G.npts) cnt = (uint32_t)seam_cell(A, s_tri, c, t, nullptr);
#pragma unroll
	for (int o = 16; o >= 1; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
	if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(&s_sum, cnt);
	__syncthreads();
	if (threadIdx.x == 0)
	{
		blk_cnt[blockIdx.x] = s_sum;
		if (s_sum) atomicAdd(chunk_cnt + c, s_sum);
	}
}

// exclusive scan of the per-chunk triangle counts (one CTA, tiles of SEAM_SCAN_CTA with a running carry)
static constexpr int SEAM_SCAN_CTA = 1024;
__global__ void __launch_bounds__(SEAM_SCAN_CTA) k_seam_scan(const uint32_t* __restrict__ chunk_cnt, int n, unsigned long long* __restrict__ chunk_base,
                                                              unsigned long long* __restrict__ total)
{
	__shared__ unsigned long long s_warp[SEAM_SCAN_CTA / 32];
	__shared__ unsigned long long s_carry;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	if (threadIdx.x == 0) s_carry = 0;
	__syncthreads();
	for (int base = 0; base < n; base += SEAM_SCAN_CTA)
	{
		const int i = base + threadIdx.x;
		const unsigned long long v = i < n ? chunk_cnt[i] : 0ull;
		unsigned long long inc = v;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1)
		{
			const unsigned long long u = __shfl_up_sync(0xffffffffu, inc, o);
			if (lane >= o) inc += u;
		}
		if (lane == 31) s_warp[warp] = inc;
		__syncthreads();
		if (warp == 0)
		{
			unsigned long long w = s_warp[lane];
#pragma unroll
			for (int o = 1; o < 32; o <<= 1)
			{
				const unsigned long long u = __shfl_up_sync(0xffffffffu, w, o);
				if (lane >= o) w += u;
			}
			s_warp[lane] = w; // inclusive over warps
		}
		__syncthreads();
		const unsigned long long before = s_carry + (warp ? s_warp[warp - 1] : 0ull) + inc - v;
		if (i < n) chunk_base[i] = before;
		__syncthreads();
		if (threadIdx.x == SEAM_SCAN_CTA - 1) s_carry = before + v;
		__syncthreads();
	}
	if (threadIdx.x == 0) *total = s_carry;
}

// pass 2: the same cells again, written at chunk base + CTAs before this one + threads before this one
__global__ void __launch_bounds__(CTA) k_seam_emit(SeamArgs A, const uint32_t* __restrict__ blk_cnt, const unsigned long long* __restrict__ chunk_base,
                                                    float* __restrict__ out)
{
	__shared__ uint64_t s_tri[256];
	__shared__ uint32_t s_w[CTA / 32];
	__shared__ unsigned long long s_base;
	s_tri[threadIdx.x] = c_tri_pack[threadIdx.x];
	const int c = blockIdx.x / A.G.bpc, b = blockIdx.x - c * A.G.bpc;
	if (blk_cnt[blockIdx.x] == 0) return; // whole CTA
	if (threadIdx.x == 0) s_base = 0;
	__syncthreads();
	// triangles of the CTAs of this chunk before this one
	{
		unsigned long long part = 0;
		for (int q = threadIdx.x; q < b; q += CTA) part += blk_cnt[(size_t)c * A.G.bpc + q];
#pragma unroll
		for (int o = 16; o >= 1; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
		if ((threadIdx.x & 31) == 0 && part) atomicAdd(&s_base, part);
	}
	const int t = b * CTA + threadIdx.x;
	float tri[45];
	uint32_t cnt = 0;
	if (t < A.G.npts) cnt = (uint32_t)seam_cell(A, s_tri, c, t, tri);
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	uint32_t inc = cnt;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1)
	{
		const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
		if (lane >= o) inc += u;
	}
	if (lane == 31) s_w[warp] = inc;
	__syncthreads();
	uint32_t before = inc - cnt;
	for (int w = 0; w < warp; w++) before += s_w[w];
	if (cnt)
	{
		float* dst = out + 9 * (size_t)(chunk_base[c] + s_base + before);
		for (uint32_t q = 0; q < 9 * cnt; q++) dst[q] = tri[q];
	}
}

} // namespace bmf

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

// the triangle table in global memory (g_tri_pack, extract.cuh): only the rare sign-changing cells read it

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
	const uint8_t* clean;    // per-chunk, bit f: no dual cell in the interior of face f (x0,x1,y0,y1,z0,z1) changes sign
	const uint32_t* act;     // per-chunk bitmap over the shell points (k_seam_classify): 0 = certainly nothing to emit
	int act_words;           // words per chunk in act
	const int32_t* group;    // per-chunk group id or null
	int cross_group_only;    // 1: only cells whose nodes span more than one group
	const ChunkGeom* geom;
	SamplerDev s;
	DensitySource src;
};

// One entry per direction (dx,dy,dz) in {-1,0,1}^3 around a chunk: the chunk on that side if it is a single
// same-or-coarser chunk (then it covers the whole face / edge / corner region), else invalid -- a finer or missing
// neighbour means this chunk owns no dual cell touching that region.
struct SeamNbr
{
	int32_t m;          // chunk index, < 0 = invalid
	int32_t bx, by, bz; // origin in finest-voxel units
	int32_t lg;
	int32_t u;          // 0: read sign words, 1: uniformly air, 2: uniformly solid
	int32_t grp;
};

// the chunk on side (dx,dy,dz) of chunk c (see SeamNbr)
__device__ __forceinline__ SeamNbr seam_nbr(const SeamArgs& A, int c, int dx, int dy, int dz)
{
	const SeamChunk me = A.chunks[c];
	const int e = 1 << me.lg;
	const int sx = me.ox + (dx < 0 ? -1 : dx > 0 ? e : 0), sy = me.oy + (dy < 0 ? -1 : dy > 0 ? e : 0), sz = me.oz + (dz < 0 ? -1 : dz > 0 ? e : 0);
	SeamNbr n;
	n.m = -1;
	n.bx = n.by = n.bz = n.lg = n.u = n.grp = 0;
	if (sx >= 0 && sy >= 0 && sz >= 0 && sx < A.G.gx && sy < A.G.gy && sz < A.G.gz)
	{
		const int m = A.slot_map[((size_t)sx * A.G.gy + sy) * A.G.gz + sz];
		if (m >= 0)
		{
			const SeamChunk o = A.chunks[m];
			if (o.lg >= me.lg)
			{
				n.m = m;
				n.bx = o.ox << A.L.ld; n.by = o.oy << A.L.ld; n.bz = o.oz << A.L.ld;
				n.lg = o.lg;
				n.u = A.uni ? (int32_t)A.uni[m] : 0;
				n.grp = A.group ? A.group[m] : 0;
			}
		}
	}
	return n;
}

// filled by the first 27 threads of the CTA for chunk c (direction index = 9 (dx+1) + 3 (dy+1) + (dz+1))
__device__ __forceinline__ void seam_fill_nbrs(const SeamArgs& A, int c, SeamNbr* tab)
{
	if (threadIdx.x < 27)
	{
		const int q = threadIdx.x;
		tab[q] = seam_nbr(A, c, q / 9 - 1, (q / 3) % 3 - 1, q % 3 - 1);
	}
}

// the dual cell at shell point t of chunk c: returns the number of triangles kept and (if tri) their 9 floats each
__device__ __forceinline__ int seam_cell(const SeamArgs& A, const SeamNbr* __restrict__ tab, int c, int t, float* tri)
{
	const int d = A.L.d, ld = A.L.ld;
	int i, j, k;
	seam_shell_point(t, d, i, j, k);
	const SeamNbr me = tab[13];
	const int Px = me.bx + (i << me.lg), Py = me.by + (j << me.lg), Pz = me.bz + (k << me.lg);
	const int hx = me.bx + (d << me.lg), hy = me.by + (d << me.lg), hz = me.bz + (d << me.lg);
	int cm[8], vx[8], vy[8], vz[8], un[8];
	int owner = c;
	bool same_group = true, ok = true;
#pragma unroll
	for (int o = 0; o < 8; o++)
	{
		const int qx = Px - 1 + (o >> 2), qy = Py - 1 + ((o >> 1) & 1), qz = Pz - 1 + (o & 1);
		const int dirx = qx < me.bx ? 0 : qx >= hx ? 2 : 1, diry = qy < me.by ? 0 : qy >= hy ? 2 : 1, dirz = qz < me.bz ? 0 : qz >= hz ? 2 : 1;
		const SeamNbr nb = tab[9 * dirx + 3 * diry + dirz];
		ok = ok && nb.m >= 0; // outside the world, a leaf that is not in the batch, or a finer chunk (which owns the cell)
		if (nb.lg == me.lg && nb.m >= 0 && nb.m < owner) owner = nb.m;
		if (nb.grp != me.grp) same_group = false;
		cm[o] = nb.m;
		un[o] = nb.u;
		vx[o] = (qx - nb.bx) >> nb.lg;
		vy[o] = (qy - nb.by) >> nb.lg;
		vz[o] = (qz - nb.bz) >> nb.lg;
	}
	if (!ok || owner != c) return 0;
	if (A.cross_group_only && same_group) return 0;
	uint32_t mask = 0;
#pragma unroll
	for (int o = 0; o < 8; o++)
	{
		uint32_t bit;
		if (un[o]) bit = (un[o] == 1) ? 1u : 0u;
		else bit = (A.bits[(size_t)cm[o] * A.L.wc + ((((size_t)vx[o] << ld) + vy[o]) << A.L.lzc) + (vz[o] >> 5)] >> (vz[o] & 31)) & 1u;
		mask |= bit << o;
	}
	if (mask == 0 || mask == 255) return 0;

	const uint64_t tp = __ldg(g_tri_pack + mask);
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

// pass 0a: which signs occur in each of the six boundary voxel layers of a chunk (bit 2f = some solid, bit 2f+1 = some
// air; f = x0,x1,y0,y1,z0,z1).  One CTA per chunk; uniform chunks answer from their flags without reading sign words.
__global__ void __launch_bounds__(CTA) k_seam_layers(Layout L, const uint32_t* __restrict__ bits, const uint32_t* __restrict__ flags, int n,
                                                      uint16_t* __restrict__ layers)
{
	__shared__ uint32_t s_f;
	const int c = blockIdx.x;
	const uint32_t f = flags[c];
	if (f == CF_ONES || f == CF_ZERO)
	{
		if (threadIdx.x == 0) layers[c] = f == CF_ONES ? 0xAAA : 0x555;
		return;
	}
	if (threadIdx.x == 0) s_f = 0;
	__syncthreads();
	const uint32_t* b = bits + (size_t)c * L.wc;
	const int d = L.d, zc = L.zc;
	uint32_t acc = 0;
	auto see = [&](uint32_t w, int face) { acc |= ((w != 0xFFFFFFFFu) ? 1u : 0u) << (2 * face) | ((w != 0u) ? 2u : 0u) << (2 * face); };
	for (int q = threadIdx.x; q < L.wp; q += CTA)
	{
		see(b[q], 0);
		see(b[(size_t)(d - 1) * L.wp + q], 1);
	}
	for (int q = threadIdx.x; q < d * zc; q += CTA)
	{
		const int x = q >> L.lzc, zb = q & (zc - 1);
		see(b[((size_t)x * d + 0) * zc + zb], 2);
		see(b[((size_t)x * d + d - 1) * zc + zb], 3);
	}
	for (int q = threadIdx.x; q < d * d; q += CTA)
	{
		const uint32_t lo = b[(size_t)q * zc] & 1u, hi = b[(size_t)q * zc + zc - 1] >> 31;
		acc |= (lo ? 2u : 1u) << 8;
		acc |= (hi ? 2u : 1u) << 10;
	}
#pragma unroll
	for (int o = 16; o >= 1; o >>= 1) acc |= __shfl_xor_sync(0xffffffffu, acc, o);
	if ((threadIdx.x & 31) == 0) atomicOr(&s_f, acc);
	__syncthreads();
	if (threadIdx.x == 0) layers[c] = (uint16_t)s_f;
}

// pass 0b, one warp per chunk:
//  * a chunk whose whole neighbourhood (itself and every chunk sharing a face, edge or corner with it, at any LOD) is
//    uniformly air or uniformly solid owns no sign-changing dual cell: it is left out of the active list;
//  * per face: the cells in the face's interior only see this chunk's boundary layer and the layer of the one
//    same-or-coarser neighbour across it (finer or missing neighbours: this chunk owns nothing there); if both
//    layers are uniformly of the same sign the face is "clean".
__global__ void __launch_bounds__(CTA) k_seam_cull(SeamGrid G, const SeamChunk* __restrict__ chunks, const int32_t* __restrict__ slot_map,
                                                    const uint32_t* __restrict__ flags, const uint16_t* __restrict__ layers, uint8_t* __restrict__ clean,
                                                    uint32_t* __restrict__ active, uint32_t* __restrict__ counters /* [0] active chunks */)
{
	const int c = (int)(((size_t)blockIdx.x * CTA + threadIdx.x) >> 5), lane = threadIdx.x & 31;
	if (c >= G.n) return;
	const uint32_t f = flags[c];
	const SeamChunk me = chunks[c];
	const int e = 1 << me.lg;
	bool same = f == CF_ONES || f == CF_ZERO;
	if (same)
	{
		const int w = e + 2;
		for (int q = lane; q < w * w * w && same; q += 32)
		{
			const int x = q / (w * w), r = q - x * w * w, y = r / w, z = r - y * w;
			if (x > 0 && x < w - 1 && y > 0 && y < w - 1 && z > 0 && z < w - 1) continue; // inside the chunk itself
			const int sx = me.ox + x - 1, sy = me.oy + y - 1, sz = me.oz + z - 1;
			if (sx < 0 || sy < 0 || sz < 0 || sx >= G.gx || sy >= G.gy || sz >= G.gz) continue;
			const int m = slot_map[((size_t)sx * G.gy + sy) * G.gz + sz];
			if (m >= 0 && flags[m] != f) same = false;
		}
	}
	same = __all_sync(0xffffffffu, same);
	if (same) return;
	uint32_t cl = 0;
	if (lane < 6)
	{
		const int axis = lane >> 1, hi = lane & 1;
		int s[3] = { me.ox, me.oy, me.oz };
		s[axis] += hi ? e : -1;
		int m = -1;
		if (s[0] >= 0 && s[1] >= 0 && s[2] >= 0 && s[0] < G.gx && s[1] < G.gy && s[2] < G.gz) m = slot_map[((size_t)s[0] * G.gy + s[1]) * G.gz + s[2]];
		if (m < 0 || chunks[m].lg < me.lg) cl = 1;
		else
		{
			const uint32_t mine = (layers[c] >> (2 * lane)) & 3u, theirs = (layers[m] >> (2 * (lane ^ 1))) & 3u;
			cl = (mine == theirs && (mine == 1u || mine == 2u)) ? 1u : 0u;
		}
	}
	cl = __ballot_sync(0xffffffffu, cl != 0) & 0x3Fu;
	if (lane == 0)
	{
		clean[c] = (uint8_t)cl;
		active[atomicAdd(counters, 1u)] = (uint32_t)c;
	}
}

// sign word w of row (x, y) of chunk m (constant for the uniform chunks of the 2-D terrains, whose words are never written)
__device__ __forceinline__ uint32_t seam_row_word(const SeamArgs& A, int m, int u, int x, int y, int w)
{
	if (u) return u == 1 ? 0xFFFFFFFFu : 0u;
	return A.bits[(size_t)m * A.L.wc + ((((size_t)x << A.L.ld) + y) << A.L.lzc) + w];
}

__device__ __forceinline__ void seam_or_bits(uint32_t* act, uint32_t t0, uint32_t mask)
{
	if (!mask) return;
	const uint32_t sh = t0 & 31u;
	atomicOr(act + (t0 >> 5), mask << sh);
	if (sh && (mask >> (32u - sh))) atomicOr(act + (t0 >> 5) + 1, mask >> (32u - sh));
}

// pass 0c (persistent over (active chunk, face)): which shell points can emit anything at all.  For a face whose
// neighbour is a chunk of the SAME level the test is exact and word-parallel for the x and y faces -- 32 dual cells per
// thread from the four sign rows that meet at the face (two of this chunk, two of the neighbour): a cell is active iff
// its 2x2x2 voxels are not all of one sign.  A coarser neighbour's rows are stretched to this chunk's resolution first.
// Edge / corner points are marked "maybe" and resolved by the generic per-point path; faces that are clean, not owned
// (lower-index neighbour of the same level, finer or missing neighbour) or filtered out by the group rule stay 0.
__global__ void __launch_bounds__(CTA) k_seam_classify(SeamArgs A, const uint32_t* __restrict__ active, const uint32_t* __restrict__ counters, uint32_t* __restrict__ actmap)
{
	const uint32_t n_tasks = counters[0] * 6u;
	const int d = A.L.d, zc = A.L.zc;
	const uint32_t fa = (uint32_t)(d + 1) * (d + 1), fb = (uint32_t)(d - 1) * (d + 1), fc = (uint32_t)(d - 1) * (d - 1);
	for (uint32_t task = blockIdx.x; task < n_tasks; task += gridDim.x)
	{
		const int c = (int)active[task / 6u], F = (int)(task % 6u), axis = F >> 1, hi = F & 1;
		uint32_t* act = actmap + (size_t)c * A.act_words;
		// edge and corner points of the shell (on two or three faces at once): generic path
		if (axis == 0)
		{
			const uint32_t base = (uint32_t)hi * fa;
			for (int q = threadIdx.x; q < 4 * (d + 1); q += CTA)
			{
				const int side = q / (d + 1), r = q - side * (d + 1);
				const int j = side == 0 ? 0 : side == 1 ? d : r, k = side < 2 ? r : (side == 2 ? 0 : d);
				seam_or_bits(act, base + (uint32_t)j * (d + 1) + k, 1u);
			}
		}
		else if (axis == 1)
		{
			const uint32_t base = 2 * fa + (uint32_t)hi * fb;
			for (int q = threadIdx.x; q < 2 * (d - 1); q += CTA)
			{
				const int i = 1 + (q >> 1), k = (q & 1) ? d : 0;
				seam_or_bits(act, base + (uint32_t)(i - 1) * (d + 1) + k, 1u);
			}
		}
		const SeamNbr nb = seam_nbr(A, c, axis == 0 ? (hi ? 1 : -1) : 0, axis == 1 ? (hi ? 1 : -1) : 0, axis == 2 ? (hi ? 1 : -1) : 0);
		const SeamChunk me = A.chunks[c];
		if (nb.m < 0) continue;                                       // finer or missing neighbour: nothing owned on this face
		if ((A.clean[c] >> F) & 1) continue;                          // both layers uniformly of one sign
		if (nb.lg == me.lg && nb.m < c) continue;                     // the neighbour owns the face
		if (A.cross_group_only && nb.grp == (A.group ? A.group[c] : 0)) continue;
		const int mu = A.uni ? (int)A.uni[c] : 0;
		// The neighbour's voxels across the face: same in-face coordinates for a chunk of the same level; for a chunk s
		// levels coarser, voxel (p, q) of this chunk faces voxel (P0 + (p >> s), Q0 + (q >> s)) of the neighbour (this chunk
		// is aligned to the neighbour's voxel lattice), i.e. every neighbour bit is seen 2^s times in a row.
		const int s = nb.lg - me.lg;
		const int obx = (int)me.ox << A.L.ld, oby = (int)me.oy << A.L.ld, obz = (int)me.oz << A.L.ld;
		const int X0 = (obx - nb.bx) >> nb.lg, Y0 = (oby - nb.by) >> nb.lg, Z0 = (obz - nb.bz) >> nb.lg;
		const int own = hi ? d - 1 : 0, opp = hi ? 0 : d - 1;
		if (s > 5)
		{
			// more than 5 levels apart (a stretched run would be shorter than one bit per word): generic path for the whole face
			for (int q = threadIdx.x; q < (d - 1) * (d - 1); q += CTA)
			{
				const int u = 1 + q / (d - 1), v = 1 + q % (d - 1);
				const uint32_t t = axis == 0 ? (uint32_t)hi * fa + (uint32_t)u * (d + 1) + v
				                 : axis == 1 ? 2 * fa + (uint32_t)hi * fb + (uint32_t)(u - 1) * (d + 1) + v : 2 * fa + 2 * fb + (uint32_t)hi * fc + q;
				seam_or_bits(act, t, 1u);
			}
		}
		else if (axis < 2)
		{
			for (int q = threadIdx.x; q < (d - 1) * zc; q += CTA)
			{
				const int r = 1 + q / zc, w = q - (r - 1) * zc; // lattice row r (j for x faces, i for y faces), word w of the z run
				uint32_t any = 0, all = 0xFFFFFFFFu, anyp = 0, allp = 1u;
#pragma unroll
				for (int rr = 0; rr < 2; rr++)
				{
					const int line = r - 1 + rr;
					// this chunk's row
					{
						const int x = axis == 0 ? own : line, y = axis == 0 ? line : own;
						const uint32_t v = seam_row_word(A, c, mu, x, y, w);
						any |= v; all &= v;
						if (w > 0)
						{
							const uint32_t vp = seam_row_word(A, c, mu, x, y, w - 1) >> 31;
							anyp |= vp; allp &= vp;
						}
					}
					// the neighbour's row, stretched by 2^s
					{
						const int x = axis == 0 ? opp : X0 + (line >> s), y = axis == 0 ? Y0 + (line >> s) : opp;
						uint32_t v, vp = 0;
						if (s == 0 || nb.u)
						{
							v = seam_row_word(A, nb.m, nb.u, x, y, w);
							if (w > 0) vp = seam_row_word(A, nb.m, nb.u, x, y, w - 1) >> 31;
						}
						else
						{
							const int z0 = Z0 + ((32 * w) >> s); // first neighbour voxel of this run; the run never straddles a word
							const uint32_t cw = seam_row_word(A, nb.m, 0, x, y, z0 >> 5) >> (z0 & 31);
							v = 0;
#pragma unroll
							for (int bit = 0; bit < 32; bit++) v |= ((cw >> (bit >> s)) & 1u) << bit;
							if (w > 0)
							{
								const int zp = Z0 + ((32 * w - 1) >> s);
								vp = (seam_row_word(A, nb.m, 0, x, y, zp >> 5) >> (zp & 31)) & 1u;
							}
						}
						any |= v; all &= v;
						if (w > 0) { anyp |= vp; allp &= vp; }
					}
				}
				const uint32_t a2 = any | ((any << 1) | (w > 0 ? anyp : 0u));
				const uint32_t l2 = all & ((all << 1) | (w > 0 ? allp : 0u));
				uint32_t m32 = a2 & ~l2;
				if (w == 0) m32 &= ~1u; // k = 0 is an edge point
				const uint32_t t0 = (axis == 0 ? (uint32_t)hi * fa + (uint32_t)r * (d + 1) : 2 * fa + (uint32_t)hi * fb + (uint32_t)(r - 1) * (d + 1)) + 32u * w;
				seam_or_bits(act, t0, m32);
			}
		}
		else
		{
			const int wo = hi ? zc - 1 : 0, bo = hi ? 31 : 0, zn = hi ? 0 : d - 1;
			for (int q = threadIdx.x; q < (d - 1) * (d - 1); q += CTA)
			{
				const int i = 1 + q / (d - 1), j = 1 + q % (d - 1);
				uint32_t any = 0, all = 1;
#pragma unroll
				for (int dx = 0; dx < 2; dx++)
#pragma unroll
					for (int dy = 0; dy < 2; dy++)
					{
						const int x = i - 1 + dx, y = j - 1 + dy;
						const uint32_t a = (seam_row_word(A, c, mu, x, y, wo) >> bo) & 1u;
						const uint32_t b = (seam_row_word(A, nb.m, nb.u, X0 + (x >> s), Y0 + (y >> s), zn >> 5) >> (zn & 31)) & 1u;
						any |= a | b; all &= a & b;
					}
				if (any && !all) seam_or_bits(act, 2 * fa + 2 * fb + (uint32_t)hi * fc + q, 1u);
			}
		}
	}
}

// ---- passes 1 and 2 (persistent over (active chunk, part)) ------------------------------------------------------------
// A task owns a contiguous range of a chunk's activity words.  256 words at a time, the set bits are compacted IN ORDER
// into a shared-memory list (popc + block scan), so the expensive per-cell work runs on dense warps whatever the
// pattern of the marked points; list order = lattice order, so positions stay the defined ones.
static constexpr int SEAM_PARTS = 4;

// exclusive scan of v over the CTA in thread order; returns the CTA total through `total` (s_w: CTA/32 + 1 words)
__device__ __forceinline__ uint32_t seam_block_scan(uint32_t v, uint32_t* s_w, uint32_t& total)
{
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	uint32_t inc = v;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1)
	{
		const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
		if (lane >= o) inc += u;
	}
	__syncthreads(); // s_w may still be read from the previous scan
	if (lane == 31) s_w[warp] = inc;
	__syncthreads();
	uint32_t before = inc - v, tot = 0;
#pragma unroll
	for (int w = 0; w < CTA / 32; w++)
	{
		const uint32_t x = s_w[w];
		if (w < warp) before += x;
		tot += x;
	}
	total = tot;
	return before;
}

template <bool EMIT>
__global__ void __launch_bounds__(CTA, EMIT ? 2 : 4) k_seam_pass(SeamArgs A, const uint32_t* __restrict__ active, const uint32_t* __restrict__ counters,
                                                                  uint32_t* __restrict__ part_cnt, uint32_t* __restrict__ chunk_cnt,
                                                                  const unsigned long long* __restrict__ chunk_base, float* __restrict__ out)
{
	__shared__ SeamNbr s_tab[27];
	__shared__ uint16_t s_list[CTA * 32];
	__shared__ uint32_t s_w[CTA / 32];
	const uint32_t n_tasks = counters[0] * (uint32_t)SEAM_PARTS;
	const int wpp = (A.act_words + SEAM_PARTS - 1) / SEAM_PARTS; // activity words per part
	for (uint32_t task = blockIdx.x; task < n_tasks; task += gridDim.x)
	{
		const int c = (int)active[task / (uint32_t)SEAM_PARTS], part = (int)(task % (uint32_t)SEAM_PARTS);
		if (EMIT && chunk_cnt[c] == 0) continue; // whole CTA
		__syncthreads(); // the previous task is done with s_tab / s_list
		seam_fill_nbrs(A, c, s_tab);
		const uint32_t* act = A.act + (size_t)c * A.act_words;
		unsigned long long base = 0;
		if (EMIT)
		{
			base = chunk_base[c];
			for (int q = 0; q < part; q++) base += part_cnt[(size_t)c * SEAM_PARTS + q];
		}
		uint32_t mine = 0; // triangles counted by this thread (pass 1)
		const int w_end = min(A.act_words, (part + 1) * wpp);
		for (int w0 = part * wpp; w0 < w_end; w0 += CTA)
		{
			const int w = w0 + threadIdx.x;
			uint32_t word = w < w_end ? act[w] : 0u;
			uint32_t n_act;
			uint32_t off = seam_block_scan((uint32_t)__popc(word), s_w, n_act); // first barrier inside also orders s_tab and the previous round's s_list reads
			if (n_act == 0) continue;
			while (word)
			{
				const int bit = __ffs(word) - 1;
				word &= word - 1;
				s_list[off++] = (uint16_t)((threadIdx.x << 5) | bit);
			}
			__syncthreads();
			for (uint32_t e0 = 0; e0 < n_act; e0 += CTA)
			{
				const uint32_t e = e0 + threadIdx.x;
				float tri[EMIT ? 45 : 1];
				uint32_t cnt = 0;
				if (e < n_act) cnt = (uint32_t)seam_cell(A, s_tab, c, (w0 << 5) + (int)s_list[e], EMIT ? tri : nullptr);
				if (EMIT)
				{
					uint32_t round_total;
					const uint32_t before = seam_block_scan(cnt, s_w, round_total);
					if (cnt)
					{
						float* dst = out + 9 * (size_t)(base + before);
						for (uint32_t q = 0; q < 9 * cnt; q++) dst[q] = tri[q];
					}
					base += round_total;
				}
				else
					mine += cnt;
			}
		}
		if (!EMIT)
		{
			uint32_t total;
			seam_block_scan(mine, s_w, total);
			if (threadIdx.x == 0)
			{
				part_cnt[(size_t)c * SEAM_PARTS + part] = total;
				if (total) atomicAdd(chunk_cnt + c, total);
			}
		}
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

} // namespace bmf

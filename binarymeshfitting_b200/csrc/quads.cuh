// quads.cuh -- quad emission (bmf_params.quads): the producer MeshProcessor<4> never had in the reference (README.md:72
// recommends quads, `Processing::MeshProcessor<4>` and Tables::NumVertices / EdgeTable exist, nothing emits a quad;
// SURVEY 8(f)#4).  Build-defined, UNPINNED; the oracle twin is orc_quads (oracle/bmf_oracle.c), same definition:
//
// Nielson's dual marching cubes on the chunk's (d-1)^3 cells.  One dual vertex per surface patch of a cell (c_patch_pack:
// the connected components of the cell's tri_table triangles), placed at the mean of the patch's edge crossing points
// (edges in ascending id, crossing = the triangle emitter's iso-vertex formula); one quad per sign-changing grid edge
// that has all four cells around it, joining the patch vertices that contain the edge.  Vertex ids follow the serial
// x->y->z cell scan (patches in table order); quads are emitted in the same scan by the cell at the edge's lower end, X
// then Y then Z edge, wound like the triangle emitter's triangles.
//
// Same count -> scan -> emit structure as the triangle path, one thread per 32-cell word; a cell's vertex id anywhere in
// the chunk is `record.base + prefix(bit planes of the patch counts)` -- one 16-byte load, like the triangle path's
// word records.
#pragma once
#include "extract.cuh"
#include "smooth.cuh"

namespace bmf
{

__constant__ uint64_t c_patch_pack[256] = BMF_PATCH_PACK_INIT;

// sign rows of the 32 cells of word (x, y, zb): r[2 dx + dy] = row (x+dx, y+dy), s[..] = the same row one voxel up in z
struct QWord
{
	uint32_t r[4], s[4];
};

__device__ __forceinline__ QWord q_load(const uint32_t* __restrict__ cb, const Layout& L, int x, int y, int zb)
{
	QWord q;
#pragma unroll
	for (int k = 0; k < 4; k++)
	{
		const size_t row = ((((size_t)(x + (k >> 1)) << L.ld) + (y + (k & 1))) << L.lzc);
		const uint32_t w = cb[row + zb];
		const uint32_t nx = zb + 1 < L.zc ? cb[row + zb + 1] : 0u;
		q.r[k] = w;
		q.s[k] = (w >> 1) | (nx << 31);
	}
	return q;
}

// corner mask of the cell at bit z of the word (corner = 4 dx + 2 dy + dz)
__device__ __forceinline__ uint32_t q_mask8(const QWord& q, int z)
{
	return ((q.r[0] >> z) & 1u) | (((q.s[0] >> z) & 1u) << 1) | (((q.r[1] >> z) & 1u) << 2) | (((q.s[1] >> z) & 1u) << 3) | (((q.r[2] >> z) & 1u) << 4) |
	       (((q.s[2] >> z) & 1u) << 5) | (((q.r[3] >> z) & 1u) << 6) | (((q.s[3] >> z) & 1u) << 7);
}

// bits of the word that are cells of the (d-1)^3 grid
__device__ __forceinline__ uint32_t q_valid_bits(const Layout& L, int x, int y, int zb)
{
	if (x >= L.d - 1 || y >= L.d - 1) return 0u;
	return zb == L.zc - 1 ? 0x7FFFFFFFu : 0xFFFFFFFFu;
}

// ---- count: per word (dual vertices | quads << 16) and per chunk (cells, vertices, indices)
__global__ void __launch_bounds__(CTA) k_q_count(const uint32_t* __restrict__ bits, const uint32_t* __restrict__ flags, Layout L, uint32_t* __restrict__ wq,
                                                  uint32_t* __restrict__ chunk_tot)
{
	const size_t gw = (size_t)blockIdx.x * CTA + threadIdx.x; // the CTA's 256 words belong to one chunk
	const int chunk = (int)(gw >> L.lwc);
	if (!flags_contain_mesh(flags[chunk])) return; // whole CTA; its words are never read
	const int w = (int)(gw & (size_t)(L.wc - 1));
	const int zb = w & (L.zc - 1), y = (w >> L.lzc) & (L.d - 1), x = w >> L.lwp;
	uint32_t nc = 0, nv = 0, nq = 0;
	uint32_t valid = q_valid_bits(L, x, y, zb);
	if (valid)
	{
		const QWord q = q_load(bits + (size_t)chunk * L.wc, L, x, y, zb);
		// cells whose eight corners are not all equal
		const uint32_t any = q.r[0] | q.r[1] | q.r[2] | q.r[3] | q.s[0] | q.s[1] | q.s[2] | q.s[3];
		const uint32_t all = q.r[0] & q.r[1] & q.r[2] & q.r[3] & q.s[0] & q.s[1] & q.s[2] & q.s[3];
		uint32_t act = any & ~all & valid;
		nc = __popc(act);
		// quads owned by the cells of this word: sign-changing X / Y / Z edge at the cell's lower corner with all four cells around it
		const uint32_t zin = zb == 0 ? 0xFFFFFFFEu : 0xFFFFFFFFu; // z >= 1
		if (y >= 1) nq += __popc((q.r[0] ^ q.r[2]) & valid & zin);
		if (x >= 1) nq += __popc((q.r[0] ^ q.r[1]) & valid & zin);
		if (x >= 1 && y >= 1) nq += __popc((q.r[0] ^ q.s[0]) & valid);
		while (act)
		{
			const int z = __ffs(act) - 1;
			act &= act - 1;
			nv += (uint32_t)(c_patch_pack[q_mask8(q, z)] >> 60);
		}
	}
	wq[gw] = nv | (nq << 16);
	uint32_t ni = 4 * nq, tot[3];
	block_scan3(nc, nv, ni, tot);
	if (threadIdx.x < 3 && tot[threadIdx.x]) atomicAdd(chunk_tot + 3 * (size_t)chunk + threadIdx.x, tot[threadIdx.x]);
}

// ---- bases: ordered scan of the words of a chunk (one CTA per chunk, tiles with a running carry).
// wqv[word] = {chunk-local id of the word's first dual vertex, bit planes 0..2 of its cells' patch counts};
// wqq[word] = chunk-local number of the word's first quad.
static constexpr int Q_ITEMS = 4;

__global__ void __launch_bounds__(CTA) k_q_bases(const uint32_t* __restrict__ bits, Layout L, const uint32_t* __restrict__ wq, const ChunkCounts* __restrict__ chunks,
                                                  uint4* __restrict__ wqv, uint32_t* __restrict__ wqq, uint2* __restrict__ wlist, unsigned long long* __restrict__ list_count,
                                                  const unsigned long long* __restrict__ tot)
{
	__shared__ uint32_t s_lbase;
	if (tot[7]) return;
	const int chunk = blockIdx.x;
	const ChunkCounts cc = chunks[chunk];
	if (!cc.contains_mesh || cc.n_verts == 0) return;
	const uint32_t* cb = bits + (size_t)chunk * L.wc;
	uint32_t carry_v = 0, carry_q = 0;
	for (int base = 0; base < L.wc; base += CTA * Q_ITEMS)
	{
		const int w0 = base + threadIdx.x * Q_ITEMS;
		uint32_t cnt[Q_ITEMS], act[Q_ITEMS], p0[Q_ITEMS], p1[Q_ITEMS], p2[Q_ITEMS], sum[3] = { 0, 0, 0 }, tt[3];
#pragma unroll
		for (int k = 0; k < Q_ITEMS; k++)
		{
			const int w = w0 + k;
			cnt[k] = wq[(size_t)chunk * L.wc + w];
			act[k] = p0[k] = p1[k] = p2[k] = 0;
			if (cnt[k] & 0xFFFFu)
			{
				const int zb = w & (L.zc - 1), y = (w >> L.lzc) & (L.d - 1), x = w >> L.lwp;
				const QWord q = q_load(cb, L, x, y, zb);
				const uint32_t any = q.r[0] | q.r[1] | q.r[2] | q.r[3] | q.s[0] | q.s[1] | q.s[2] | q.s[3];
				const uint32_t all = q.r[0] & q.r[1] & q.r[2] & q.r[3] & q.s[0] & q.s[1] & q.s[2] & q.s[3];
				act[k] = any & ~all & q_valid_bits(L, x, y, zb);
				uint32_t a = act[k];
				while (a)
				{
					const int z = __ffs(a) - 1;
					a &= a - 1;
					const uint32_t np = (uint32_t)(c_patch_pack[q_mask8(q, z)] >> 60);
					p0[k] |= (np & 1u) << z; p1[k] |= ((np >> 1) & 1u) << z; p2[k] |= ((np >> 2) & 1u) << z;
				}
			}
			sum[0] += cnt[k] & 0xFFFFu;
			sum[1] += cnt[k] >> 16;
			sum[2] += __popc(act[k]);
		}
		block_scan<3>(sum, tt);
		// the active cells go on a compact list (order irrelevant: every cell finds its own output positions)
		if (threadIdx.x == 0) s_lbase = tt[2] ? (uint32_t)atomicAdd(list_count, (unsigned long long)tt[2]) : 0u;
		__syncthreads();
		uint32_t li = s_lbase + sum[2];
		uint32_t rv = carry_v + sum[0], rq = carry_q + sum[1];
#pragma unroll
		for (int k = 0; k < Q_ITEMS; k++)
		{
			const int w = w0 + k;
			if (cnt[k])
			{
				const uint32_t gw = (uint32_t)((size_t)chunk * L.wc + w);
				wqv[gw] = make_uint4(rv, p0[k], p1[k], p2[k]);
				wqq[gw] = rq;
				uint32_t a = act[k];
				while (a)
				{
					const int z = __ffs(a) - 1;
					a &= a - 1;
					wlist[li++] = make_uint2(gw, (uint32_t)z);
				}
			}
			rv += cnt[k] & 0xFFFFu;
			rq += cnt[k] >> 16;
		}
		carry_v += tt[0];
		carry_q += tt[1];
		__syncthreads(); // s_lbase is rewritten in the next tile
	}
}

// chunk-local id of the dual vertex of cell (x, y, z) whose patch contains local edge e
__device__ __forceinline__ uint32_t q_vertex_id(const uint32_t* __restrict__ cb, const uint4* __restrict__ rec, const Layout& L, int x, int y, int z, int e)
{
	const int zb = z >> 5, bit = z & 31;
	const uint4 r = rec[((((size_t)x << L.ld) + y) << L.lzc) + zb];
	const uint32_t lt = (1u << bit) - 1u;
	uint32_t id = r.x + __popc(r.y & lt) + 2u * __popc(r.z & lt) + 4u * __popc(r.w & lt);
	if ((((r.z | r.w) >> bit) & 1u) == 0u) return id; // one patch (the common case): no need to look at the cell's corners
	const QWord q = q_load(cb, L, x, y, zb);
	const uint64_t pp = c_patch_pack[q_mask8(q, bit)];
	uint32_t p = 0;
#pragma unroll
	for (int k = 1; k < 4; k++)
		if ((pp >> (12 * k + e)) & 1ull) p = k;
	return id + p;
}

// ---- emit: one THREAD per listed cell (persistent grid).  A cell's first vertex and first quad follow from its word's
// record by popc prefixes (patch-count bit planes; X / Y / Z quad-owner masks), so every cell is independent.
__device__ __forceinline__ void q_emit_word(size_t gw, int bit, const uint32_t* __restrict__ bits, const Layout& L, const uint32_t* __restrict__ wq,
                                            const uint4* __restrict__ wqv, const uint32_t* __restrict__ wqq, const ChunkCounts* __restrict__ chunks,
                                            const SamplerDev& s, const DensitySource& src, const ChunkGeom* __restrict__ geom, float* __restrict__ pos,
                                            uint8_t* __restrict__ boundary, uint8_t* __restrict__ valence, uint32_t* __restrict__ inds)
{
	const int chunk = (int)(gw >> L.lwc);
	const ChunkCounts cc = chunks[chunk];
	const uint32_t cnt = wq[gw];
	const int w = (int)(gw & (size_t)(L.wc - 1));
	const int zb = w & (L.zc - 1), y = (w >> L.lzc) & (L.d - 1), x = w >> L.lwp;
	const int d = L.d;
	const uint32_t* cb = bits + (size_t)chunk * L.wc;
	const uint4* rec = wqv + (size_t)chunk * L.wc;
	const QWord q = q_load(cb, L, x, y, zb);
	const uint32_t valid = q_valid_bits(L, x, y, zb);
	const uint32_t any = q.r[0] | q.r[1] | q.r[2] | q.r[3] | q.s[0] | q.s[1] | q.s[2] | q.s[3];
	const uint32_t all = q.r[0] & q.r[1] & q.r[2] & q.r[3] & q.s[0] & q.s[1] & q.s[2] & q.s[3];
	const uint32_t act = any & ~all & valid;
	if (!((act >> bit) & 1u)) return;
	// quad owners of this word (the same masks k_q_count counted)
	const uint32_t zin = zb == 0 ? 0xFFFFFFFEu : 0xFFFFFFFFu;
	const uint32_t ox = y >= 1 ? (q.r[0] ^ q.r[2]) & valid & zin : 0u;
	const uint32_t oy = x >= 1 ? (q.r[0] ^ q.r[1]) & valid & zin : 0u;
	const uint32_t oz = (x >= 1 && y >= 1) ? (q.r[0] ^ q.s[0]) & valid : 0u;
	const uint32_t lt = (1u << bit) - 1u;
	const uint4 r = (cnt & 0xFFFFu) ? rec[w] : make_uint4(0, 0, 0, 0);
	size_t v = (size_t)cc.vert_base + r.x + __popc(r.y & lt) + 2u * __popc(r.z & lt) + 4u * __popc(r.w & lt);
	size_t qo = (size_t)cc.ind_base + 4 * ((size_t)wqq[(size_t)chunk * L.wc + w] + __popc(ox & lt) + __popc(oy & lt) + __popc(oz & lt));
	const ChunkGeom g = geom[chunk];
	const int z = zb * 32 + bit;
	const uint32_t m = q_mask8(q, bit);
	const uint64_t pp = c_patch_pack[m];
	const int np = (int)(pp >> 60);
	const uint8_t bd = (x == 0 || y == 0 || z == 0 || x == d - 2 || y == d - 2 || z == d - 2) ? 1 : 0;
	for (int p = 0; p < np; p++)
	{
		const uint32_t em = (uint32_t)(pp >> (12 * p)) & 0xFFFu;
		float sx = 0.0f, sy = 0.0f, sz = 0.0f;
		int n = 0;
		for (int e = 0; e < 12; e++)
		{
			if (!((em >> e) & 1u)) continue;
			const int axis = e >> 2, hi = (e >> 1) & 1, lo = e & 1;
			const int a = axis == 0 ? ((hi << 1) | lo) : axis == 1 ? ((hi << 2) | lo) : ((hi << 2) | (lo << 1));
			const int x0 = x + (a >> 2), y0 = y + ((a >> 1) & 1), z0 = z + (a & 1);
			const int x1 = x0 + (axis == 0), y1 = y0 + (axis == 1), z1 = z0 + (axis == 2);
			const float s0 = density_at(s, src, g, d, chunk, x0, y0, z0), s1 = density_at(s, src, g, d, chunk, x1, y1, z1);
			// _get_intersection (DMCChunk.cpp:657-662), grid units
			const float mu = (0.0f - s0) / (s1 - s0);
			sx += ((float)x1 - (float)x0) * mu + (float)x0;
			sy += ((float)y1 - (float)y0) * mu + (float)y0;
			sz += ((float)z1 - (float)z0) * mu + (float)z0;
			n++;
		}
		pos[3 * v] = sx / (float)n;
		pos[3 * v + 1] = sy / (float)n;
		pos[3 * v + 2] = sz / (float)n;
		boundary[v] = bd;
		// init_valence = quads on this vertex = edges of the patch that have all four cells around them
		int val = 0;
		for (int e = 0; e < 12; e++)
		{
			if (!((em >> e) & 1u)) continue;
			const int axis = e >> 2, hi = (e >> 1) & 1, lo = e & 1;
			// lattice coordinates of the edge across its axis
			const int u = axis == 0 ? y + hi : x + hi, w2 = axis == 2 ? y + lo : z + lo;
			val += (u >= 1 && u <= d - 2 && w2 >= 1 && w2 <= d - 2) ? 1 : 0;
		}
		valence[v] = (uint8_t)val;
		v++;
	}
	const uint32_t b0 = m & 1u;
	for (int axis = 0; axis < 3; axis++)
	{
		const uint32_t own = axis == 0 ? ox : axis == 1 ? oy : oz;
		if (!((own >> bit) & 1u)) continue;
		uint32_t id[4];
#pragma unroll
		for (int rr = 0; rr < 4; rr++)
		{
			// ring of the four cells around the edge, counter-clockwise about the +axis
			const int du = axis == 1 ? (rr >> 1) : ((rr == 1 || rr == 2) ? 1 : 0);
			const int dv = axis == 1 ? ((rr == 1 || rr == 2) ? 1 : 0) : (rr >> 1);
			const int cx = axis == 0 ? x : x - du, cy = axis == 0 ? y - du : (axis == 1 ? y : y - dv), cz = axis == 2 ? z : z - dv;
			id[rr] = q_vertex_id(cb, rec, L, cx, cy, cz, 4 * axis + ((du << 1) | dv));
		}
#pragma unroll
		for (int rr = 0; rr < 4; rr++)
		{
			const uint32_t vi = b0 ? id[rr] : id[3 - rr];
			inds[qo + rr] = vi;
		}
		qo += 4;
	}
}

__global__ void __launch_bounds__(CTA) k_q_emit(const uint32_t* __restrict__ bits, Layout L, const uint32_t* __restrict__ wq, const uint4* __restrict__ wqv,
                                                 const uint32_t* __restrict__ wqq, const uint2* __restrict__ wlist, const unsigned long long* __restrict__ list_count,
                                                 const ChunkCounts* __restrict__ chunks, SamplerDev s, DensitySource src, const ChunkGeom* __restrict__ geom,
                                                 float* __restrict__ pos, uint8_t* __restrict__ boundary, uint8_t* __restrict__ valence, uint32_t* __restrict__ inds,
                                                 const unsigned long long* __restrict__ tot)
{
	if (tot[7]) return;
	const uint32_t n_cells = (uint32_t)list_count[0];
	for (uint32_t e = blockIdx.x * CTA + threadIdx.x; e < n_cells; e += gridDim.x * CTA)
	{
		const uint2 c = wlist[e];
		q_emit_word((size_t)c.x, (int)c.y, bits, L, wq, wqv, wqq, chunks, s, src, geom, pos, boundary, valence, inds);
	}
}

// MeshProcessor<4>::flush_to_tris (MeshProcessor.cpp:73-91): (v0,v1,v2,v3) -> (v0,v1,v2),(v2,v3,v0)
__global__ void __launch_bounds__(CTA) k_quads_to_tris(const uint32_t* __restrict__ quads, size_t n_quads, uint32_t* __restrict__ tris)
{
	for (size_t q = (size_t)blockIdx.x * CTA + threadIdx.x; q < n_quads; q += (size_t)gridDim.x * CTA)
	{
		const uint4 v = reinterpret_cast<const uint4*>(quads)[q];
		uint32_t* o = tris + 6 * q;
		o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.z; o[4] = v.w; o[5] = v.x;
	}
}

// GLChunk::format_data(vertices, indexes, unwind_verts = true, smooth_normals) (GLChunk.cpp:278-335) over the whole quad
// batch: per quad the four corner positions / colours and one normal for all four -- the mean of the corner normals
// (smooth) or -normalize((n0 + n1) / 2) of the two triangle normals with the reference's NaN guards.  One thread per quad.
__global__ void __launch_bounds__(CTA) k_format_unwind(const float* __restrict__ pos, const float* __restrict__ normal, const float* __restrict__ color,
                                                        const uint32_t* __restrict__ inds, size_t n_quads, const ChunkCounts* __restrict__ chunks, int n_chunks,
                                                        int smooth_normals, float* __restrict__ p_out, float* __restrict__ n_out, float* __restrict__ c_out)
{
	for (size_t q = (size_t)blockIdx.x * CTA + threadIdx.x; q < n_quads; q += (size_t)gridDim.x * CTA)
	{
		const size_t vb = (size_t)chunks[chunk_of_index(chunks, n_chunks, 4 * (uint64_t)q)].vert_base;
		const uint4 iv = reinterpret_cast<const uint4*>(inds)[q];
		const size_t v[4] = { vb + iv.x, vb + iv.y, vb + iv.z, vb + iv.w };
		f3 p[4];
#pragma unroll
		for (int k = 0; k < 4; k++)
		{
			p[k] = ld3(pos, v[k]);
			st3(p_out, 4 * q + k, p[k]);
			st3(c_out, 4 * q + k, ld3(color, v[k]));
		}
		f3 n;
		if (smooth_normals)
		{
			const f3 a = ld3(normal, v[0]), b = ld3(normal, v[1]), c = ld3(normal, v[2]), d = ld3(normal, v[3]);
			n = mul3(add3(add3(add3(a, b), c), d), 0.25f);
		}
		else
		{
			f3 n0 = cross3(normalize3(sub3(p[0], p[1])), normalize3(sub3(p[0], p[2])));
			f3 n1 = cross3(normalize3(sub3(p[2], p[3])), normalize3(sub3(p[2], p[0])));
			if (isnan(n0.x)) n0 = n1;
			if (isnan(n1.x)) n1 = n0;
			const f3 h = normalize3(mul3(add3(n0, n1), 0.5f));
			n = { -h.x, -h.y, -h.z };
		}
#pragma unroll
		for (int k = 0; k < 4; k++) st3(n_out, 4 * q + k, n);
	}
}

} // namespace bmf

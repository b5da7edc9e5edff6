// post.cuh -- the two mesh post-processing steps around MeshProcessor that the reference ships but never calls:
// ColorMapper::generate_colors (ColorMapper.cpp:15-60) and MeshProcessor<4>::collapse_bad_quads (MeshProcessor.cpp:308-396).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "extract.cuh"

namespace bmf
{

// ---- ColorMapper ----------------------------------------------------------------------------------------------------
// hsl_to_rgb (ColorMapper.cpp:62-121): despite the name it is the HSV sextant formula; every operation in the reference's order
__device__ __forceinline__ void hsl_to_rgb_dev(float h, float s, float v, float& r, float& g, float& b)
{
	if (s <= 0.0f) { r = v; g = v; b = v; return; }
	float hh = h;
	hh = fmodf(fabsf(hh), 360.0f);
	hh = hh / 60.0f;
	const int i = (int)hh;
	const float ff = hh - (float)i;
	const float p = v * (1.0f - s);
	const float q = v * (1.0f - (s * ff));
	const float t = v * (1.0f - (s * (1.0f - ff)));
	switch (i)
	{
	case 0: r = v; g = t; b = p; break;
	case 1: r = q; g = v; b = p; break;
	case 2: r = p; g = v; b = t; break;
	case 3: r = p; g = q; b = v; break;
	case 4: r = t; g = p; b = v; break;
	default: r = v; g = p; b = q; break;
	}
}

// get_noise + map_noise (ColorMapper.cpp:27-60): one thread per vertex.  The vector set is the vertex positions (scale 1.0), the noise a
// fresh FastNoiseSIMD object set to SimplexFractal / 4 octaves / FBM (`ns`, built by the host); n = noise * 4,
// colour = hsl_to_rgb((n + 1) * 0.5 * 360, 0.72, 1).
__global__ void __launch_bounds__(CTA) k_color_map(NoiseState ns, const float* __restrict__ pos, size_t n, float* __restrict__ color)
{
	const float scale = 1.0f;
	for (size_t i = (size_t)blockIdx.x * CTA + threadIdx.x; i < n; i += (size_t)gridDim.x * CTA)
	{
		const float noise = noise_eval<NT_SIMPLEX>(ns, pos[3 * i] * scale, pos[3 * i + 1] * scale, pos[3 * i + 2] * scale);
		const float nn = noise * 4.0f;
		float r, g, b;
		hsl_to_rgb_dev((nn + 1.0f) * 0.5f * 360.0f, 0.72f, 1.0f, r, g, b);
		color[3 * i] = r; color[3 * i + 1] = g; color[3 * i + 2] = b;
	}
}

// ---- collapse_bad_quads ---------------------------------------------------------------------------------------------
// The reference's loop is serial and order-dependent: quad i is judged on the state the quads before it left behind (adj_next of its
// corners, its own -- possibly rewired -- corners, the corners of the quads around it).  What makes a parallel form exact:
//   * a collapse only ever LOWERS the number of valence-3 corners (`next`) of any quad: it sets adj_next 3 -> 4 on the kept corner
//     and rewires the opposite (valence-3) corner of the neighbours to that valence-4 vertex;
//   * a quad with next < 2 is a no-op whatever else the state is (MeshProcessor.cpp:359).
// So one CTA walks the quads in tiles: every thread evaluates `next` of one quad of the tile on the state at the START of the tile --
// an upper bound of what the serial loop will see -- and only the quads with next >= 2 (rare: pairs of valence-3 vertices) are then
// replayed by ONE thread in ascending order with the reference's exact logic on the live state.  Everything else was a no-op in
// the reference too.  Afterwards the same CTA flushes the surviving quads in order (MeshProcessor<4>::flush, :57-71).
static constexpr int COLLAPSE_CTA = 1024;

template <typename T>
__device__ __forceinline__ T ldv(const T* p) { return *reinterpret_cast<const volatile T*>(p); }
template <typename T>
__device__ __forceinline__ void stv(T* p, T v) { *reinterpret_cast<volatile T*>(p) = v; }

__global__ void __launch_bounds__(COLLAPSE_CTA, 1) k_collapse_bad_quads(uint32_t* quads, uint32_t n_quads, float* pos, uint8_t* adj_next, uint32_t* adj_off, uint32_t* adj,
                                                                          uint32_t adj_count0, uint8_t* destroyed, uint32_t* __restrict__ flushed,
                                                                          unsigned long long* __restrict__ out /* [0] bad_count, [1] surviving quads */)
{
	__shared__ uint32_t s_list[COLLAPSE_CTA];
	__shared__ uint32_t s_warp[COLLAPSE_CTA / 32];
	__shared__ uint32_t s_n, s_carry;
	const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
	uint32_t adj_count = adj_count0, bad = 0; // live in thread 0 only
	for (uint32_t base = 0; base < n_quads; base += COLLAPSE_CTA)
	{
		const uint32_t i = base + tid;
		bool cand = false;
		if (i < n_quads)
		{
			destroyed[i] = 0;
			int nx = 0;
#pragma unroll
			for (int k = 0; k < 4; k++) nx += ldv(adj_next + ldv(quads + 4 * (size_t)i + k)) == 3 ? 1 : 0;
			cand = nx >= 2;
		}
		// ordered list of the tile's candidates
		const uint32_t bal = __ballot_sync(0xffffffffu, cand);
		if (lane == 0) s_warp[wid] = __popc(bal);
		__syncthreads();
		if (wid == 0)
		{
			const uint32_t v = s_warp[lane];
			uint32_t inc = v;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1)
			{
				const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
				if (lane >= o) inc += u;
			}
			s_warp[lane] = inc - v;
			if (lane == 31) s_n = inc;
		}
		__syncthreads();
		if (cand) s_list[s_warp[wid] + __popc(bal & ((1u << lane) - 1u))] = i;
		__syncthreads();
		if (tid == 0)
		{
			const uint32_t nc = s_n;
			for (uint32_t c = 0; c < nc; c++)
			{
				// ---- MeshProcessor.cpp:315-390 for quad q, on the live state
				const uint32_t q = s_list[c];
				uint32_t* pv = quads + 4 * (size_t)q;
				uint32_t v4[4] = { ldv(pv), ldv(pv + 1), ldv(pv + 2), ldv(pv + 3) };
				uint32_t pair[4], p_out[12] = { 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0 };
				int next_p = 0, nx = 0;
				for (int k = 0; k < 4; k++)
				{
					const uint32_t dv = v4[k];
					if (ldv(adj_next + dv) == 3)
					{
						pair[nx++] = (uint32_t)k;
						const uint32_t o = ldv(adj_off + dv);
						for (int a = 0; a < 3; a++)
						{
							const uint32_t e = ldv(adj + o + a);
							if (e != 0xFFFFFFFFu && e != q) p_out[next_p++] = e;
						}
					}
				}
				if (nx == 4 && next_p == 8) continue;
				if (nx != 2 || next_p != 4 || pair[1] - pair[0] != 2) continue;
				float np3[3] = { 0.0f, 0.0f, 0.0f };
				const uint32_t new_index = v4[pair[0]];
				for (int k = 0; k < 4; k++)
					for (int a = 0; a < 3; a++) np3[a] = np3[a] + ldv(pos + 3 * (size_t)v4[k] + a);
				for (int a = 0; a < 3; a++) stv(pos + 3 * (size_t)new_index + a, np3[a] * 0.25f);
				stv(adj_next + new_index, (uint8_t)4);
				const uint32_t p_other = v4[pair[1]];
				for (int k = 0; k < 4; k++)
				{
					uint32_t* nv = quads + 4 * (size_t)p_out[k];
					const uint32_t n0 = ldv(nv), n1 = ldv(nv + 1), n2 = ldv(nv + 2), n3 = ldv(nv + 3);
					if (n0 == new_index || n1 == new_index || n2 == new_index || n3 == new_index) continue;
					if (n0 == p_other) stv(nv, new_index);
					else if (n1 == p_other) stv(nv + 1, new_index);
					else if (n2 == p_other) stv(nv + 2, new_index);
					else if (n3 == p_other) stv(nv + 3, new_index);
				}
				stv(adj_off + new_index, adj_count);
				for (int k = 0; k < 4; k++) stv(adj + adj_count++, p_out[k]);
				destroyed[q] = 1;
				bad++;
			}
			__threadfence_block();
		}
		__syncthreads();
	}
	// ---- flush (MeshProcessor.cpp:57-71): the corners of the quads that were not destroyed, in order
	if (tid == 0) s_carry = 0;
	__syncthreads();
	for (uint32_t base = 0; base < n_quads; base += COLLAPSE_CTA)
	{
		const uint32_t i = base + tid;
		const bool keep = i < n_quads && !ldv(destroyed + i);
		const uint32_t bal = __ballot_sync(0xffffffffu, keep);
		if (lane == 0) s_warp[wid] = __popc(bal);
		__syncthreads();
		if (wid == 0)
		{
			const uint32_t v = s_warp[lane];
			uint32_t inc = v;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1)
			{
				const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
				if (lane >= o) inc += u;
			}
			s_warp[lane] = inc - v;
			if (lane == 31) s_n = inc;
		}
		__syncthreads();
		if (keep)
		{
			const size_t o = (size_t)s_carry + s_warp[wid] + __popc(bal & ((1u << lane) - 1u));
#pragma unroll
			for (int k = 0; k < 4; k++) flushed[4 * o + k] = ldv(quads + 4 * (size_t)i + k);
		}
		__syncthreads();
		if (tid == 0) s_carry += s_n;
		__syncthreads();
	}
	if (tid == 0)
	{
		out[0] = bad;
		out[1] = s_carry;
	}
}

} // namespace bmf

// noise.cuh -- K1 device arithmetic: fractal value/perlin/simplex noise with gradient perturb, and the
// ImplicitSampler primitives.  One thread evaluates one point; everything stays in registers.
//
// The reference delegates noise to the external FastNoiseSIMD library (call sites
// NoiseSampler.cpp:113-263).  What is implemented here is that library's published algorithm at its
// FMA SIMD level: each fused multiply-add is an explicit __fmaf_rn, everything else is a separate
// IEEE op (the translation unit is built with -fmad=false so nvcc never contracts on its own).
// Integer hashing wraps mod 2^32 (unsigned arithmetic) and shifts arithmetically where the library does.
#pragma once
#include <cstdint>

namespace bmf
{

enum { NT_VALUE = 0, NT_PERLIN = 1, NT_SIMPLEX = 2 };
enum { FT_FBM = 0, FT_BILLOW = 1, FT_RIGIDMULTI = 2 };

struct NoiseState
{
	int32_t seed;
	float frequency;
	int32_t base;       // NT_*
	int32_t fractal;    // 0 = single octave, 1 = fractal
	int32_t octaves;
	float lacunarity, gain;
	int32_t fractal_type; // FT_*
	float fractal_bounding;
	int32_t perturb;      // 0 none, 1 gradient, 2 gradient fractal
	float perturb_amp;    // already / 511.5
	float perturb_frequency;
	int32_t perturb_octaves;
	float perturb_lacunarity, perturb_gain, perturb_bounding;
};

static constexpr int32_t XPRIME = 1619, YPRIME = 31337, ZPRIME = 6971;
static constexpr uint32_t HASHPRIME = 60493u;

__device__ __forceinline__ int32_t hash_hb(int32_t seed, int32_t x, int32_t y, int32_t z)
{
	uint32_t h = (uint32_t)seed ^ (uint32_t)x ^ (uint32_t)y ^ (uint32_t)z;
	h = ((h * h) * HASHPRIME) * h;
	return (int32_t)h;
}

__device__ __forceinline__ int32_t hash_full(int32_t seed, int32_t x, int32_t y, int32_t z)
{
	int32_t h = hash_hb(seed, x, y, z);
	return (h >> 13) ^ h;
}

__device__ __forceinline__ float val_coord(int32_t seed, int32_t x, int32_t y, int32_t z)
{
	return (1.0f / 2147483648.0f) * __int2float_rn(hash_hb(seed, x, y, z));
}

__device__ __forceinline__ float grad_coord(int32_t seed, int32_t xi, int32_t yi, int32_t zi, float x, float y, float z)
{
	int32_t hash = hash_full(seed, xi, yi, zi);
	int32_t h13 = hash & 13;
	float u = (h13 < 8) ? x : y;
	float v = (h13 < 2) ? y : ((h13 == 12) ? x : z);
	uint32_t h1 = (uint32_t)hash << 31;
	uint32_t h2 = ((uint32_t)hash & 2u) << 30;
	return __uint_as_float(__float_as_uint(u) ^ h1) + __uint_as_float(__float_as_uint(v) ^ h2);
}

__device__ __forceinline__ float lerpf(float a, float b, float t) { return __fmaf_rn(b - a, t, a); }

__device__ __forceinline__ float quintic(float t)
{
	float r = __fmaf_rn(t, 6.0f, -15.0f);
	r = __fmaf_rn(r, t, 10.0f);
	r = r * t;
	r = r * t;
	r = r * t;
	return r;
}

__device__ __forceinline__ float value_single(int32_t seed, float x, float y, float z)
{
	float xs = floorf(x), ys = floorf(y), zs = floorf(z);
	int32_t x0 = (int32_t)((uint32_t)(int32_t)xs * (uint32_t)XPRIME);
	int32_t y0 = (int32_t)((uint32_t)(int32_t)ys * (uint32_t)YPRIME);
	int32_t z0 = (int32_t)((uint32_t)(int32_t)zs * (uint32_t)ZPRIME);
	int32_t x1 = (int32_t)((uint32_t)x0 + (uint32_t)XPRIME);
	int32_t y1 = (int32_t)((uint32_t)y0 + (uint32_t)YPRIME);
	int32_t z1 = (int32_t)((uint32_t)z0 + (uint32_t)ZPRIME);
	xs = quintic(x - xs);
	ys = quintic(y - ys);
	zs = quintic(z - zs);
	return lerpf(
		lerpf(lerpf(val_coord(seed, x0, y0, z0), val_coord(seed, x1, y0, z0), xs),
		      lerpf(val_coord(seed, x0, y1, z0), val_coord(seed, x1, y1, z0), xs), ys),
		lerpf(lerpf(val_coord(seed, x0, y0, z1), val_coord(seed, x1, y0, z1), xs),
		      lerpf(val_coord(seed, x0, y1, z1), val_coord(seed, x1, y1, z1), xs), ys),
		zs);
}

__device__ __forceinline__ float perlin_single(int32_t seed, float x, float y, float z)
{
	float xs = floorf(x), ys = floorf(y), zs = floorf(z);
	int32_t x0 = (int32_t)((uint32_t)(int32_t)xs * (uint32_t)XPRIME);
	int32_t y0 = (int32_t)((uint32_t)(int32_t)ys * (uint32_t)YPRIME);
	int32_t z0 = (int32_t)((uint32_t)(int32_t)zs * (uint32_t)ZPRIME);
	int32_t x1 = (int32_t)((uint32_t)x0 + (uint32_t)XPRIME);
	int32_t y1 = (int32_t)((uint32_t)y0 + (uint32_t)YPRIME);
	int32_t z1 = (int32_t)((uint32_t)z0 + (uint32_t)ZPRIME);
	float xf0 = x - xs, yf0 = y - ys, zf0 = z - zs;
	float xf1 = xf0 - 1.0f, yf1 = yf0 - 1.0f, zf1 = zf0 - 1.0f;
	xs = quintic(xf0);
	ys = quintic(yf0);
	zs = quintic(zf0);
	return lerpf(
		lerpf(lerpf(grad_coord(seed, x0, y0, z0, xf0, yf0, zf0), grad_coord(seed, x1, y0, z0, xf1, yf0, zf0), xs),
		      lerpf(grad_coord(seed, x0, y1, z0, xf0, yf1, zf0), grad_coord(seed, x1, y1, z0, xf1, yf1, zf0), xs), ys),
		lerpf(lerpf(grad_coord(seed, x0, y0, z1, xf0, yf0, zf1), grad_coord(seed, x1, y0, z1, xf1, yf0, zf1), xs),
		      lerpf(grad_coord(seed, x0, y1, z1, xf0, yf1, zf1), grad_coord(seed, x1, y1, z1, xf1, yf1, zf1), xs), ys),
		zs);
}

__device__ __forceinline__ float simplex_single(int32_t seed, float x, float y, float z)
{
	const float F3 = 1.0f / 3.0f;
	const float G3 = 1.0f / 6.0f;
	const float G32 = (1.0f / 6.0f) * 2.0f;
	const float G33 = (1.0f / 6.0f) * 3.0f - 1.0f;

	float f = F3 * ((x + y) + z);
	float x0 = floorf(x + f);
	float y0 = floorf(y + f);
	float z0 = floorf(z + f);

	int32_t i = (int32_t)((uint32_t)(int32_t)x0 * (uint32_t)XPRIME);
	int32_t j = (int32_t)((uint32_t)(int32_t)y0 * (uint32_t)YPRIME);
	int32_t k = (int32_t)((uint32_t)(int32_t)z0 * (uint32_t)ZPRIME);

	float g = G3 * ((x0 + y0) + z0);
	x0 = x - (x0 - g);
	y0 = y - (y0 - g);
	z0 = z - (z0 - g);

	bool x0_ge_y0 = x0 >= y0;
	bool y0_ge_z0 = y0 >= z0;
	bool x0_ge_z0 = x0 >= z0;

	bool i1 = x0_ge_y0 && x0_ge_z0;
	bool j1 = (!x0_ge_y0) && y0_ge_z0;
	bool k1 = (!x0_ge_z0) && (!y0_ge_z0);

	bool i2 = x0_ge_y0 || x0_ge_z0;
	bool j2 = (!x0_ge_y0) || y0_ge_z0;
	bool k2 = !(x0_ge_z0 && y0_ge_z0);

	float x1 = (i1 ? x0 - 1.0f : x0) + G3;
	float y1 = (j1 ? y0 - 1.0f : y0) + G3;
	float z1 = (k1 ? z0 - 1.0f : z0) + G3;
	float x2 = (i2 ? x0 - 1.0f : x0) + G32;
	float y2 = (j2 ? y0 - 1.0f : y0) + G32;
	float z2 = (k2 ? z0 - 1.0f : z0) + G32;
	float x3 = x0 + G33;
	float y3 = y0 + G33;
	float z3 = z0 + G33;

	float t0 = __fmaf_rn(-z0, z0, __fmaf_rn(-y0, y0, __fmaf_rn(-x0, x0, 0.6f)));
	float t1 = __fmaf_rn(-z1, z1, __fmaf_rn(-y1, y1, __fmaf_rn(-x1, x1, 0.6f)));
	float t2 = __fmaf_rn(-z2, z2, __fmaf_rn(-y2, y2, __fmaf_rn(-x2, x2, 0.6f)));
	float t3 = __fmaf_rn(-z3, z3, __fmaf_rn(-y3, y3, __fmaf_rn(-x3, x3, 0.6f)));

	bool n0 = t0 >= 0.0f;
	bool n1 = t1 >= 0.0f;
	bool n2 = t2 >= 0.0f;
	bool n3 = t3 >= 0.0f;

	t0 = t0 * t0;
	t1 = t1 * t1;
	t2 = t2 * t2;
	t3 = t3 * t3;

	float v0 = (t0 * t0) * grad_coord(seed, i, j, k, x0, y0, z0);
	float v1 = (t1 * t1) * grad_coord(seed,
		(int32_t)((uint32_t)i + (i1 ? (uint32_t)XPRIME : 0u)),
		(int32_t)((uint32_t)j + (j1 ? (uint32_t)YPRIME : 0u)),
		(int32_t)((uint32_t)k + (k1 ? (uint32_t)ZPRIME : 0u)), x1, y1, z1);
	float v2 = (t2 * t2) * grad_coord(seed,
		(int32_t)((uint32_t)i + (i2 ? (uint32_t)XPRIME : 0u)),
		(int32_t)((uint32_t)j + (j2 ? (uint32_t)YPRIME : 0u)),
		(int32_t)((uint32_t)k + (k2 ? (uint32_t)ZPRIME : 0u)), x2, y2, z2);
	float v3 = (t3 * t3) * grad_coord(seed,
		(int32_t)((uint32_t)i + (uint32_t)XPRIME),
		(int32_t)((uint32_t)j + (uint32_t)YPRIME),
		(int32_t)((uint32_t)k + (uint32_t)ZPRIME), x3, y3, z3);

	float r = n0 ? v0 : 0.0f;
	r = r + (n1 ? v1 : 0.0f);
	r = r + (n2 ? v2 : 0.0f);
	r = r + (n3 ? v3 : 0.0f);
	return 32.0f * r;
}

template <int BASE>
__device__ __forceinline__ float noise_single(int32_t seed, float x, float y, float z)
{
	if (BASE == NT_VALUE) return value_single(seed, x, y, z);
	if (BASE == NT_PERLIN) return perlin_single(seed, x, y, z);
	return simplex_single(seed, x, y, z);
}

// one octave of the gradient perturb: displaces (x,y,z)
__device__ __forceinline__ void gradient_perturb_single(int32_t seed, float amp, float freq, float& x, float& y, float& z)
{
	float xf = x * freq, yf = y * freq, zf = z * freq;
	float xs = floorf(xf), ys = floorf(yf), zs = floorf(zf);
	int32_t x0 = (int32_t)((uint32_t)(int32_t)xs * (uint32_t)XPRIME);
	int32_t y0 = (int32_t)((uint32_t)(int32_t)ys * (uint32_t)YPRIME);
	int32_t z0 = (int32_t)((uint32_t)(int32_t)zs * (uint32_t)ZPRIME);
	int32_t x1 = (int32_t)((uint32_t)x0 + (uint32_t)XPRIME);
	int32_t y1 = (int32_t)((uint32_t)y0 + (uint32_t)YPRIME);
	int32_t z1 = (int32_t)((uint32_t)z0 + (uint32_t)ZPRIME);
	xs = quintic(xf - xs);
	ys = quintic(yf - ys);
	zs = quintic(zf - zs);

	// three 10-bit fields of the high-bit hash at each lattice corner
	int32_t h000 = hash_hb(seed, x0, y0, z0), h100 = hash_hb(seed, x1, y0, z0);
	int32_t h010 = hash_hb(seed, x0, y1, z0), h110 = hash_hb(seed, x1, y1, z0);
	int32_t h001 = hash_hb(seed, x0, y0, z1), h101 = hash_hb(seed, x1, y0, z1);
	int32_t h011 = hash_hb(seed, x0, y1, z1), h111 = hash_hb(seed, x1, y1, z1);

#define BMF_FIELD(h, sh) __int2float_rn(((h) >> (sh)) & 1023)
	float gx0 = lerpf(lerpf(BMF_FIELD(h000, 0), BMF_FIELD(h100, 0), xs), lerpf(BMF_FIELD(h010, 0), BMF_FIELD(h110, 0), xs), ys);
	float gy0 = lerpf(lerpf(BMF_FIELD(h000, 10), BMF_FIELD(h100, 10), xs), lerpf(BMF_FIELD(h010, 10), BMF_FIELD(h110, 10), xs), ys);
	float gz0 = lerpf(lerpf(BMF_FIELD(h000, 20), BMF_FIELD(h100, 20), xs), lerpf(BMF_FIELD(h010, 20), BMF_FIELD(h110, 20), xs), ys);
	float gx1 = lerpf(lerpf(BMF_FIELD(h001, 0), BMF_FIELD(h101, 0), xs), lerpf(BMF_FIELD(h011, 0), BMF_FIELD(h111, 0), xs), ys);
	float gy1 = lerpf(lerpf(BMF_FIELD(h001, 10), BMF_FIELD(h101, 10), xs), lerpf(BMF_FIELD(h011, 10), BMF_FIELD(h111, 10), xs), ys);
	float gz1 = lerpf(lerpf(BMF_FIELD(h001, 20), BMF_FIELD(h101, 20), xs), lerpf(BMF_FIELD(h011, 20), BMF_FIELD(h111, 20), xs), ys);
#undef BMF_FIELD

	x = __fmaf_rn(lerpf(gx0, gx1, zs) - 511.5f, amp, x);
	y = __fmaf_rn(lerpf(gy0, gy1, zs) - 511.5f, amp, y);
	z = __fmaf_rn(lerpf(gz0, gz1, zs) - 511.5f, amp, z);
}

// ---- the same octave, evaluated by a full converged warp.  A perturb lattice cell is many voxels wide at the low octaves (the world default
// perturbs at 0.585 x 2^o cells per noise unit, a 64^3 chunk of the finest LOD steps 0.0006 noise units per voxel), so the 32 voxels of a warp
// usually share ONE cell: then its 8 corner hashes and their 24 field conversions -- half of the octave's instructions -- are the same in every
// lane.  Lanes 0..7 hash one corner each and leave its three fields in shared memory; everybody reads the eight records back as broadcast
// 128-bit loads and does only the interpolation.  Same operations on the same operands as gradient_perturb_single, so the same bits; a warp
// that straddles a cell boundary (or holds a NaN) takes gradient_perturb_single itself.  s_corner: 8 float4 owned by this warp.
// `share` (warp-uniform): cleared by the first octave whose cell the warp does not share -- the octaves after it have twice the frequency each and
// will not share theirs either, so they skip the test.
__device__ __forceinline__ void gradient_perturb_warp(int32_t seed, float amp, float freq, float& x, float& y, float& z, float4* s_corner, int lane, bool& share)
{
	if (!share)
	{
		gradient_perturb_single(seed, amp, freq, x, y, z);
		return;
	}
	const float xf = x * freq, yf = y * freq, zf = z * freq;
	float xs = floorf(xf), ys = floorf(yf), zs = floorf(zf);
	const float xs0 = __shfl_sync(0xffffffffu, xs, 0), ys0 = __shfl_sync(0xffffffffu, ys, 0), zs0 = __shfl_sync(0xffffffffu, zs, 0);
	if (!__all_sync(0xffffffffu, xs == xs0 && ys == ys0 && zs == zs0))
	{
		share = false;
		gradient_perturb_single(seed, amp, freq, x, y, z);
		return;
	}
	if (lane < 8)
	{
		const int32_t x0 = (int32_t)((uint32_t)(int32_t)xs * (uint32_t)XPRIME);
		const int32_t y0 = (int32_t)((uint32_t)(int32_t)ys * (uint32_t)YPRIME);
		const int32_t z0 = (int32_t)((uint32_t)(int32_t)zs * (uint32_t)ZPRIME);
		const int32_t xc = (int32_t)((uint32_t)x0 + ((lane & 1) ? (uint32_t)XPRIME : 0u));
		const int32_t yc = (int32_t)((uint32_t)y0 + ((lane & 2) ? (uint32_t)YPRIME : 0u));
		const int32_t zc = (int32_t)((uint32_t)z0 + ((lane & 4) ? (uint32_t)ZPRIME : 0u));
		const int32_t h = hash_hb(seed, xc, yc, zc);
		s_corner[lane] = make_float4(__int2float_rn(h & 1023), __int2float_rn((h >> 10) & 1023), __int2float_rn((h >> 20) & 1023), 0.0f);
	}
	__syncwarp();
	const float4 c000 = s_corner[0], c100 = s_corner[1], c010 = s_corner[2], c110 = s_corner[3];
	const float4 c001 = s_corner[4], c101 = s_corner[5], c011 = s_corner[6], c111 = s_corner[7];
	__syncwarp(); // the next octave's corners may overwrite the records only after every lane has read them
	xs = quintic(xf - xs);
	ys = quintic(yf - ys);
	zs = quintic(zf - zs);
	const float gx0 = lerpf(lerpf(c000.x, c100.x, xs), lerpf(c010.x, c110.x, xs), ys);
	const float gy0 = lerpf(lerpf(c000.y, c100.y, xs), lerpf(c010.y, c110.y, xs), ys);
	const float gz0 = lerpf(lerpf(c000.z, c100.z, xs), lerpf(c010.z, c110.z, xs), ys);
	const float gx1 = lerpf(lerpf(c001.x, c101.x, xs), lerpf(c011.x, c111.x, xs), ys);
	const float gy1 = lerpf(lerpf(c001.y, c101.y, xs), lerpf(c011.y, c111.y, xs), ys);
	const float gz1 = lerpf(lerpf(c001.z, c101.z, xs), lerpf(c011.z, c111.z, xs), ys);
	x = __fmaf_rn(lerpf(gx0, gx1, zs) - 511.5f, amp, x);
	y = __fmaf_rn(lerpf(gy0, gy1, zs) - 511.5f, amp, y);
	z = __fmaf_rn(lerpf(gz0, gz1, zs) - 511.5f, amp, z);
}

// noise at raw vector-set coordinates (FillNoiseSet with sampleScale == 0 and zero offsets):
// coord = fma(v, frequency, 0), then perturb, then the (fractal) base noise
template <int BASE>
__device__ __forceinline__ float noise_fractal(const NoiseState& s, float xF, float yF, float zF);

template <int BASE>
__device__ __forceinline__ float noise_eval(const NoiseState& s, float vx, float vy, float vz)
{
	float xF = __fmaf_rn(vx, s.frequency, 0.0f);
	float yF = __fmaf_rn(vy, s.frequency, 0.0f);
	float zF = __fmaf_rn(vz, s.frequency, 0.0f);

	if (s.perturb == 1)
	{
		gradient_perturb_single(s.seed - 1, s.perturb_amp, s.perturb_frequency, xF, yF, zF);
	}
	else if (s.perturb == 2)
	{
		int32_t seedF = s.seed - 1;
		float freqF = s.perturb_frequency;
		float ampF = s.perturb_amp * s.perturb_bounding;
		gradient_perturb_single(seedF, ampF, freqF, xF, yF, zF);
		for (int o = 1; o < s.perturb_octaves; o++)
		{
			freqF = freqF * s.perturb_lacunarity;
			seedF = seedF - 1;
			ampF = ampF * s.perturb_gain;
			gradient_perturb_single(seedF, ampF, freqF, xF, yF, zF);
		}
	}
	return noise_fractal<BASE>(s, xF, yF, zF);
}

// noise_eval for the sampling kernels, where a FULL CONVERGED warp evaluates 32 neighbouring points: the perturb octaves share their lattice
// corners across the warp when they can (gradient_perturb_warp).  Bit-identical to noise_eval.
template <int BASE>
__device__ __forceinline__ float noise_eval_warp(const NoiseState& s, float vx, float vy, float vz, float4* s_corner, int lane)
{
	float xF = __fmaf_rn(vx, s.frequency, 0.0f);
	float yF = __fmaf_rn(vy, s.frequency, 0.0f);
	float zF = __fmaf_rn(vz, s.frequency, 0.0f);

	bool share = true;
	if (s.perturb == 1)
	{
		gradient_perturb_warp(s.seed - 1, s.perturb_amp, s.perturb_frequency, xF, yF, zF, s_corner, lane, share);
	}
	else if (s.perturb == 2)
	{
		int32_t seedF = s.seed - 1;
		float freqF = s.perturb_frequency;
		float ampF = s.perturb_amp * s.perturb_bounding;
		gradient_perturb_warp(seedF, ampF, freqF, xF, yF, zF, s_corner, lane, share);
		for (int o = 1; o < s.perturb_octaves; o++)
		{
			freqF = freqF * s.perturb_lacunarity;
			seedF = seedF - 1;
			ampF = ampF * s.perturb_gain;
			gradient_perturb_warp(seedF, ampF, freqF, xF, yF, zF, s_corner, lane, share);
		}
	}
	return noise_fractal<BASE>(s, xF, yF, zF);
}

template <int BASE>
__device__ __forceinline__ float noise_fractal(const NoiseState& s, float xF, float yF, float zF)
{
	if (!s.fractal)
		return noise_single<BASE>(s.seed, xF, yF, zF);

	int32_t seedF = s.seed;
	float ampF = 1.0f;
	float result;
	if (s.fractal_type == FT_FBM)
	{
		result = noise_single<BASE>(seedF, xF, yF, zF);
		for (int o = 1; o < s.octaves; o++)
		{
			xF = xF * s.lacunarity; yF = yF * s.lacunarity; zF = zF * s.lacunarity;
			seedF = seedF + 1;
			ampF = ampF * s.gain;
			result = __fmaf_rn(noise_single<BASE>(seedF, xF, yF, zF), ampF, result);
		}
		return result * s.fractal_bounding;
	}
	if (s.fractal_type == FT_BILLOW)
	{
		result = __fmaf_rn(fabsf(noise_single<BASE>(seedF, xF, yF, zF)), 2.0f, -1.0f);
		for (int o = 1; o < s.octaves; o++)
		{
			xF = xF * s.lacunarity; yF = yF * s.lacunarity; zF = zF * s.lacunarity;
			seedF = seedF + 1;
			ampF = ampF * s.gain;
			result = __fmaf_rn(__fmaf_rn(fabsf(noise_single<BASE>(seedF, xF, yF, zF)), 2.0f, -1.0f), ampF, result);
		}
		return result * s.fractal_bounding;
	}
	result = 1.0f - fabsf(noise_single<BASE>(seedF, xF, yF, zF));
	for (int o = 1; o < s.octaves; o++)
	{
		xF = xF * s.lacunarity; yF = yF * s.lacunarity; zF = zF * s.lacunarity;
		seedF = seedF + 1;
		ampF = ampF * s.gain;
		result = __fmaf_rn(-(1.0f - fabsf(noise_single<BASE>(seedF, xF, yF, zF))), ampF, result);
	}
	return result;
}

// ---- ImplicitSampler primitives (ImplicitSampler.cpp:27-65): density = -SDF, positive inside ----------

enum { IK_SPHERE = 0, IK_TORUS_Z = 1, IK_CUBOID = 2, IK_PLANE_Y = 3 };

__device__ __forceinline__ float implicit_value(int kind, float ws, float px, float py, float pz)
{
	switch (kind)
	{
	case IK_SPHERE:
	{
		float r = ws * 0.25f;
		float len = sqrtf((px * px + py * py) + pz * pz);
		return -(len - r);
	}
	case IK_TORUS_Z:
	{
		float r1 = ws / 4.0f;
		float r2 = ws / 10.0f;
		float qx = fabsf(sqrtf(px * px + py * py)) - r1;
		float len = sqrtf(qx * qx + pz * pz);
		return -(len - r2);
	}
	case IK_CUBOID:
	{
		float r = ws / 8.0f;
		float dx = fabsf(px) - r, dy = fabsf(py) - r, dz = fabsf(pz) - r;
		float m = fmaxf(dx, fmaxf(dy, dz));
		float len = sqrtf((dx * dx + dy * dy) + dz * dz);
		return -fminf(m, len);
	}
	default:
		return -py;
	}
}

} // namespace bmf

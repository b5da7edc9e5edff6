/*
 * fastnoise_ref.h -- scalar CPU restatement of the subset of Auburns/FastNoiseSIMD that
 * Lin20/BinaryMeshFitting calls.  TEST INFRASTRUCTURE ONLY (oracle): nothing in the product
 * path may include, link or call this file.
 *
 * PARITY UNPINNED.  The arithmetic of the reference's noise lives in the third-party library
 * Auburns/FastNoiseSIMD, which is neither vendored in /root/reference nor version-pinned
 * (only `find_package(FastNoiseSIMD)`: BinaryMeshFitting/CMakeLists.txt:21,32,40,
 * cmake/Modules/FindFastNoiseSIMD.cmake:1-25) and is not installable offline.  The reference
 * holds no golden vectors for it.  This file restates the library's *published* algorithm
 * (v0.7-era FastNoiseSIMD_internal.cpp, FMA-capable SIMD level: AVX2/AVX-512, i.e. every
 * SIMDf_MUL_ADD / MUL_SUB / NMUL_ADD is one fused operation) as recalled; it is the declared
 * oracle for the noise stage and is anchored on the reference's call sites:
 *   NoiseSampler.cpp:8-35    NOISE_BLOCK builds the FastNoiseVectorSet (sampleScale = 0)
 *   NoiseSampler.cpp:113-146 terrain2d_block        (ValueFractal, 12 oct, FBM)
 *   NoiseSampler.cpp:148-194 terrain2d_pert_block   (ValueFractal + GradientFractal perturb)
 *   NoiseSampler.cpp:196-227 terrain3d_block        (ValueFractal, 4 oct, RigidMulti)
 *   NoiseSampler.cpp:229-263 terrain3d_pert_block   (SimplexFractal 8 oct RigidMulti + perturb)
 *   NoiseSampler.hpp:78-139  NewFastNoiseSIMD() per OMP thread (seed 1337, library defaults)
 *
 * All arithmetic is IEEE binary32 with explicit fmaf() where the FMA build fuses; build this
 * file with -ffp-contract=off so nothing else is contracted.
 */
#ifndef FASTNOISE_REF_H
#define FASTNOISE_REF_H

#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { FNR_VALUE = 0, FNR_VALUE_FRACTAL, FNR_PERLIN, FNR_PERLIN_FRACTAL, FNR_SIMPLEX, FNR_SIMPLEX_FRACTAL };
enum { FNR_FBM = 0, FNR_BILLOW, FNR_RIGIDMULTI };
enum { FNR_PERTURB_NONE = 0, FNR_PERTURB_GRADIENT, FNR_PERTURB_GRADIENT_FRACTAL };

typedef struct fnr_state
{
	int32_t seed;
	float frequency;
	float x_scale, y_scale, z_scale;
	int noise_type;
	int octaves;
	float lacunarity;
	float gain;
	int fractal_type;
	float fractal_bounding;
	int perturb_type;
	float perturb_amp; /* stored already divided by 511.5, like the library's setter */
	float perturb_frequency;
	int perturb_octaves;
	float perturb_lacunarity;
	float perturb_gain;
	float perturb_bounding;
} fnr_state;

static inline float fnr_bounding(float gain, int octaves)
{
	float amp = gain;
	float amp_fractal = 1.0f;
	for (int i = 1; i < octaves; i++)
	{
		amp_fractal += amp;
		amp *= gain;
	}
	return 1.0f / amp_fractal;
}

/* library defaults: seed 1337, frequency 0.01, SimplexFractal, 3 octaves, lacunarity 2, gain 0.5, FBM,
 * no perturb, perturb amp 1.0 (/511.5), perturb frequency 0.5, 3 perturb octaves, lacunarity 2, gain 0.5 */
static inline void fnr_init(fnr_state* s, int32_t seed)
{
	s->seed = seed;
	s->frequency = 0.01f;
	s->x_scale = s->y_scale = s->z_scale = 1.0f;
	s->noise_type = FNR_SIMPLEX_FRACTAL;
	s->octaves = 3;
	s->lacunarity = 2.0f;
	s->gain = 0.5f;
	s->fractal_type = FNR_FBM;
	s->fractal_bounding = fnr_bounding(s->gain, s->octaves);
	s->perturb_type = FNR_PERTURB_NONE;
	s->perturb_amp = 1.0f / 511.5f;
	s->perturb_frequency = 0.5f;
	s->perturb_octaves = 3;
	s->perturb_lacunarity = 2.0f;
	s->perturb_gain = 0.5f;
	s->perturb_bounding = fnr_bounding(s->perturb_gain, s->perturb_octaves);
}

static inline void fnr_set_fractal_octaves(fnr_state* s, int o) { s->octaves = o; s->fractal_bounding = fnr_bounding(s->gain, s->octaves); }
static inline void fnr_set_fractal_gain(fnr_state* s, float g) { s->gain = g; s->fractal_bounding = fnr_bounding(s->gain, s->octaves); }
static inline void fnr_set_perturb_amp(fnr_state* s, float a) { s->perturb_amp = a / 511.5f; }
static inline void fnr_set_perturb_octaves(fnr_state* s, int o) { s->perturb_octaves = o; s->perturb_bounding = fnr_bounding(s->perturb_gain, s->perturb_octaves); }
static inline void fnr_set_perturb_gain(fnr_state* s, float g) { s->perturb_gain = g; s->perturb_bounding = fnr_bounding(s->perturb_gain, s->perturb_octaves); }

#define FNR_XPRIME 1619
#define FNR_YPRIME 31337
#define FNR_ZPRIME 6971
#define FNR_HASHPRIME 60493u

/* a*b + c, a*b - c, -(a*b) + c : single rounding (the FMA SIMD levels) */
#define FNR_MUL_ADD(a, b, c) fmaf((a), (b), (c))
#define FNR_MUL_SUB(a, b, c) fmaf((a), (b), -(c))
#define FNR_NMUL_ADD(a, b, c) fmaf(-(a), (b), (c))

/* hash without the final xor-shift ("high bits" hash): 32-bit wrapping integer arithmetic */
static inline int32_t fnr_hash_hb(int32_t seed, int32_t x, int32_t y, int32_t z)
{
	uint32_t h = (uint32_t)seed ^ (uint32_t)x ^ (uint32_t)y ^ (uint32_t)z;
	h = ((h * h) * FNR_HASHPRIME) * h;
	return (int32_t)h;
}

static inline int32_t fnr_hash(int32_t seed, int32_t x, int32_t y, int32_t z)
{
	int32_t h = fnr_hash_hb(seed, x, y, z);
	return (h >> 13) ^ h; /* arithmetic shift (srai) */
}

static inline float fnr_val_coord(int32_t seed, int32_t x, int32_t y, int32_t z)
{
	return (1.0f / 2147483648.0f) * (float)fnr_hash_hb(seed, x, y, z);
}

static inline float fnr_xor_sign(float v, uint32_t signbit)
{
	uint32_t u;
	memcpy(&u, &v, 4);
	u ^= signbit;
	memcpy(&v, &u, 4);
	return v;
}

static inline float fnr_grad_coord(int32_t seed, int32_t xi, int32_t yi, int32_t zi, float x, float y, float z)
{
	int32_t hash = fnr_hash(seed, xi, yi, zi);
	int32_t h13 = hash & 13;
	float u = (h13 < 8) ? x : y;
	float v = (h13 < 2) ? y : ((h13 == 12) ? x : z);
	uint32_t h1 = (uint32_t)hash << 31;
	uint32_t h2 = ((uint32_t)hash & 2u) << 30;
	return fnr_xor_sign(u, h1) + fnr_xor_sign(v, h2);
}

static inline float fnr_lerp(float a, float b, float t)
{
	float r = b - a;
	return FNR_MUL_ADD(r, t, a);
}

static inline float fnr_quintic(float t)
{
	float r = FNR_MUL_SUB(t, 6.0f, 15.0f);
	r = FNR_MUL_ADD(r, t, 10.0f);
	r = r * t;
	r = r * t;
	r = r * t;
	return r;
}

static inline float fnr_value_single(int32_t seed, float x, float y, float z)
{
	float xs = floorf(x), ys = floorf(y), zs = floorf(z);
	int32_t x0 = (int32_t)((uint32_t)(int32_t)xs * (uint32_t)FNR_XPRIME);
	int32_t y0 = (int32_t)((uint32_t)(int32_t)ys * (uint32_t)FNR_YPRIME);
	int32_t z0 = (int32_t)((uint32_t)(int32_t)zs * (uint32_t)FNR_ZPRIME);
	int32_t x1 = (int32_t)((uint32_t)x0 + (uint32_t)FNR_XPRIME);
	int32_t y1 = (int32_t)((uint32_t)y0 + (uint32_t)FNR_YPRIME);
	int32_t z1 = (int32_t)((uint32_t)z0 + (uint32_t)FNR_ZPRIME);
	xs = fnr_quintic(x - xs);
	ys = fnr_quintic(y - ys);
	zs = fnr_quintic(z - zs);
	return fnr_lerp(
		fnr_lerp(fnr_lerp(fnr_val_coord(seed, x0, y0, z0), fnr_val_coord(seed, x1, y0, z0), xs),
		         fnr_lerp(fnr_val_coord(seed, x0, y1, z0), fnr_val_coord(seed, x1, y1, z0), xs), ys),
		fnr_lerp(fnr_lerp(fnr_val_coord(seed, x0, y0, z1), fnr_val_coord(seed, x1, y0, z1), xs),
		         fnr_lerp(fnr_val_coord(seed, x0, y1, z1), fnr_val_coord(seed, x1, y1, z1), xs), ys),
		zs);
}

static inline float fnr_perlin_single(int32_t seed, float x, float y, float z)
{
	float xs = floorf(x), ys = floorf(y), zs = floorf(z);
	int32_t x0 = (int32_t)((uint32_t)(int32_t)xs * (uint32_t)FNR_XPRIME);
	int32_t y0 = (int32_t)((uint32_t)(int32_t)ys * (uint32_t)FNR_YPRIME);
	int32_t z0 = (int32_t)((uint32_t)(int32_t)zs * (uint32_t)FNR_ZPRIME);
	int32_t x1 = (int32_t)((uint32_t)x0 + (uint32_t)FNR_XPRIME);
	int32_t y1 = (int32_t)((uint32_t)y0 + (uint32_t)FNR_YPRIME);
	int32_t z1 = (int32_t)((uint32_t)z0 + (uint32_t)FNR_ZPRIME);
	float xf0 = x - xs, yf0 = y - ys, zf0 = z - zs;
	float xf1 = xf0 - 1.0f, yf1 = yf0 - 1.0f, zf1 = zf0 - 1.0f;
	xs = fnr_quintic(xf0);
	ys = fnr_quintic(yf0);
	zs = fnr_quintic(zf0);
	return fnr_lerp(
		fnr_lerp(fnr_lerp(fnr_grad_coord(seed, x0, y0, z0, xf0, yf0, zf0), fnr_grad_coord(seed, x1, y0, z0, xf1, yf0, zf0), xs),
		         fnr_lerp(fnr_grad_coord(seed, x0, y1, z0, xf0, yf1, zf0), fnr_grad_coord(seed, x1, y1, z0, xf1, yf1, zf0), xs), ys),
		fnr_lerp(fnr_lerp(fnr_grad_coord(seed, x0, y0, z1, xf0, yf0, zf1), fnr_grad_coord(seed, x1, y0, z1, xf1, yf0, zf1), xs),
		         fnr_lerp(fnr_grad_coord(seed, x0, y1, z1, xf0, yf1, zf1), fnr_grad_coord(seed, x1, y1, z1, xf1, yf1, zf1), xs), ys),
		zs);
}

static inline float fnr_simplex_single(int32_t seed, float x, float y, float z)
{
	const float F3 = 1.0f / 3.0f;
	const float G3 = 1.0f / 6.0f;
	const float G32 = (1.0f / 6.0f) * 2.0f;
	const float G33 = (1.0f / 6.0f) * 3.0f - 1.0f;

	float f = F3 * ((x + y) + z);
	float x0 = floorf(x + f);
	float y0 = floorf(y + f);
	float z0 = floorf(z + f);

	int32_t i = (int32_t)((uint32_t)(int32_t)x0 * (uint32_t)FNR_XPRIME);
	int32_t j = (int32_t)((uint32_t)(int32_t)y0 * (uint32_t)FNR_YPRIME);
	int32_t k = (int32_t)((uint32_t)(int32_t)z0 * (uint32_t)FNR_ZPRIME);

	float g = G3 * ((x0 + y0) + z0);
	x0 = x - (x0 - g);
	y0 = y - (y0 - g);
	z0 = z - (z0 - g);

	int x0_ge_y0 = x0 >= y0;
	int y0_ge_z0 = y0 >= z0;
	int x0_ge_z0 = x0 >= z0;

	int i1 = x0_ge_y0 & x0_ge_z0;
	int j1 = (!x0_ge_y0) & y0_ge_z0;
	int k1 = (!x0_ge_z0) & (!y0_ge_z0);

	int i2 = x0_ge_y0 | x0_ge_z0;
	int j2 = (!x0_ge_y0) | y0_ge_z0;
	int k2 = !(x0_ge_z0 & y0_ge_z0);

	float x1 = (i1 ? x0 - 1.0f : x0) + G3;
	float y1 = (j1 ? y0 - 1.0f : y0) + G3;
	float z1 = (k1 ? z0 - 1.0f : z0) + G3;
	float x2 = (i2 ? x0 - 1.0f : x0) + G32;
	float y2 = (j2 ? y0 - 1.0f : y0) + G32;
	float z2 = (k2 ? z0 - 1.0f : z0) + G32;
	float x3 = x0 + G33;
	float y3 = y0 + G33;
	float z3 = z0 + G33;

	float t0 = FNR_NMUL_ADD(z0, z0, FNR_NMUL_ADD(y0, y0, FNR_NMUL_ADD(x0, x0, 0.6f)));
	float t1 = FNR_NMUL_ADD(z1, z1, FNR_NMUL_ADD(y1, y1, FNR_NMUL_ADD(x1, x1, 0.6f)));
	float t2 = FNR_NMUL_ADD(z2, z2, FNR_NMUL_ADD(y2, y2, FNR_NMUL_ADD(x2, x2, 0.6f)));
	float t3 = FNR_NMUL_ADD(z3, z3, FNR_NMUL_ADD(y3, y3, FNR_NMUL_ADD(x3, x3, 0.6f)));

	int n0 = t0 >= 0.0f;
	int n1 = t1 >= 0.0f;
	int n2 = t2 >= 0.0f;
	int n3 = t3 >= 0.0f;

	t0 = t0 * t0;
	t1 = t1 * t1;
	t2 = t2 * t2;
	t3 = t3 * t3;

	float v0 = (t0 * t0) * fnr_grad_coord(seed, i, j, k, x0, y0, z0);
	float v1 = (t1 * t1) * fnr_grad_coord(seed,
		(int32_t)((uint32_t)i + (i1 ? (uint32_t)FNR_XPRIME : 0u)),
		(int32_t)((uint32_t)j + (j1 ? (uint32_t)FNR_YPRIME : 0u)),
		(int32_t)((uint32_t)k + (k1 ? (uint32_t)FNR_ZPRIME : 0u)), x1, y1, z1);
	float v2 = (t2 * t2) * fnr_grad_coord(seed,
		(int32_t)((uint32_t)i + (i2 ? (uint32_t)FNR_XPRIME : 0u)),
		(int32_t)((uint32_t)j + (j2 ? (uint32_t)FNR_YPRIME : 0u)),
		(int32_t)((uint32_t)k + (k2 ? (uint32_t)FNR_ZPRIME : 0u)), x2, y2, z2);
	float v3 = (t3 * t3) * fnr_grad_coord(seed,
		(int32_t)((uint32_t)i + (uint32_t)FNR_XPRIME),
		(int32_t)((uint32_t)j + (uint32_t)FNR_YPRIME),
		(int32_t)((uint32_t)k + (uint32_t)FNR_ZPRIME), x3, y3, z3);

	float r = n0 ? v0 : 0.0f;
	r = r + (n1 ? v1 : 0.0f);
	r = r + (n2 ? v2 : 0.0f);
	r = r + (n3 ? v3 : 0.0f);
	return 32.0f * r;
}

static inline float fnr_single(int base_type, int32_t seed, float x, float y, float z)
{
	switch (base_type)
	{
	case FNR_VALUE: return fnr_value_single(seed, x, y, z);
	case FNR_PERLIN: return fnr_perlin_single(seed, x, y, z);
	default: return fnr_simplex_single(seed, x, y, z);
	}
}

/* one octave of the gradient perturb: displaces (x,y,z) in place */
static inline void fnr_gradient_perturb_single(int32_t seed, float amp, float freq, float* x, float* y, float* z)
{
	float xf = *x * freq, yf = *y * freq, zf = *z * freq;
	float xs = floorf(xf), ys = floorf(yf), zs = floorf(zf);
	int32_t xi[2], yi[2], zi[2];
	xi[0] = (int32_t)((uint32_t)(int32_t)xs * (uint32_t)FNR_XPRIME);
	yi[0] = (int32_t)((uint32_t)(int32_t)ys * (uint32_t)FNR_YPRIME);
	zi[0] = (int32_t)((uint32_t)(int32_t)zs * (uint32_t)FNR_ZPRIME);
	xi[1] = (int32_t)((uint32_t)xi[0] + (uint32_t)FNR_XPRIME);
	yi[1] = (int32_t)((uint32_t)yi[0] + (uint32_t)FNR_YPRIME);
	zi[1] = (int32_t)((uint32_t)zi[0] + (uint32_t)FNR_ZPRIME);
	xs = fnr_quintic(xf - xs);
	ys = fnr_quintic(yf - ys);
	zs = fnr_quintic(zf - zs);

	/* per lattice corner: three 10-bit fields of the high-bit hash */
	float gx[2][2][2], gy[2][2][2], gz[2][2][2];
	for (int a = 0; a < 2; a++)
		for (int b = 0; b < 2; b++)
			for (int c = 0; c < 2; c++)
			{
				int32_t h = fnr_hash_hb(seed, xi[a], yi[b], zi[c]);
				gx[a][b][c] = (float)(h & 1023);
				gy[a][b][c] = (float)((h >> 10) & 1023);
				gz[a][b][c] = (float)((h >> 20) & 1023);
			}

	float x0y = fnr_lerp(fnr_lerp(gx[0][0][0], gx[1][0][0], xs), fnr_lerp(gx[0][1][0], gx[1][1][0], xs), ys);
	float y0y = fnr_lerp(fnr_lerp(gy[0][0][0], gy[1][0][0], xs), fnr_lerp(gy[0][1][0], gy[1][1][0], xs), ys);
	float z0y = fnr_lerp(fnr_lerp(gz[0][0][0], gz[1][0][0], xs), fnr_lerp(gz[0][1][0], gz[1][1][0], xs), ys);
	float x1y = fnr_lerp(fnr_lerp(gx[0][0][1], gx[1][0][1], xs), fnr_lerp(gx[0][1][1], gx[1][1][1], xs), ys);
	float y1y = fnr_lerp(fnr_lerp(gy[0][0][1], gy[1][0][1], xs), fnr_lerp(gy[0][1][1], gy[1][1][1], xs), ys);
	float z1y = fnr_lerp(fnr_lerp(gz[0][0][1], gz[1][0][1], xs), fnr_lerp(gz[0][1][1], gz[1][1][1], xs), ys);

	*x = FNR_MUL_ADD(fnr_lerp(x0y, x1y, zs) - 511.5f, amp, *x);
	*y = FNR_MUL_ADD(fnr_lerp(y0y, y1y, zs) - 511.5f, amp, *y);
	*z = FNR_MUL_ADD(fnr_lerp(z0y, z1y, zs) - 511.5f, amp, *z);
}

/* noise at one already frequency-scaled point (after perturb) */
static inline float fnr_eval_scaled(const fnr_state* s, float xF, float yF, float zF)
{
	switch (s->perturb_type)
	{
	case FNR_PERTURB_GRADIENT:
		fnr_gradient_perturb_single(s->seed - 1, s->perturb_amp, s->perturb_frequency, &xF, &yF, &zF);
		break;
	case FNR_PERTURB_GRADIENT_FRACTAL:
	{
		int32_t seedF = s->seed - 1;
		float freqF = s->perturb_frequency;
		float ampF = s->perturb_amp * s->perturb_bounding;
		fnr_gradient_perturb_single(seedF, ampF, freqF, &xF, &yF, &zF);
		int octave = 0;
		while (++octave < s->perturb_octaves)
		{
			freqF = freqF * s->perturb_lacunarity;
			seedF = seedF - 1;
			ampF = ampF * s->perturb_gain;
			fnr_gradient_perturb_single(seedF, ampF, freqF, &xF, &yF, &zF);
		}
		break;
	}
	default: break;
	}

	int fractal = (s->noise_type == FNR_VALUE_FRACTAL || s->noise_type == FNR_PERLIN_FRACTAL || s->noise_type == FNR_SIMPLEX_FRACTAL);
	int base = (s->noise_type == FNR_VALUE || s->noise_type == FNR_VALUE_FRACTAL) ? FNR_VALUE
	         : (s->noise_type == FNR_PERLIN || s->noise_type == FNR_PERLIN_FRACTAL) ? FNR_PERLIN : FNR_SIMPLEX;
	if (!fractal)
		return fnr_single(base, s->seed, xF, yF, zF);

	int32_t seedF = s->seed;
	float ampF = 1.0f;
	float result;
	int octave = 0;
	switch (s->fractal_type)
	{
	case FNR_FBM:
		result = fnr_single(base, seedF, xF, yF, zF);
		while (++octave < s->octaves)
		{
			xF = xF * s->lacunarity; yF = yF * s->lacunarity; zF = zF * s->lacunarity;
			seedF = seedF + 1;
			ampF = ampF * s->gain;
			result = FNR_MUL_ADD(fnr_single(base, seedF, xF, yF, zF), ampF, result);
		}
		return result * s->fractal_bounding;
	case FNR_BILLOW:
		result = FNR_MUL_SUB(fabsf(fnr_single(base, seedF, xF, yF, zF)), 2.0f, 1.0f);
		while (++octave < s->octaves)
		{
			xF = xF * s->lacunarity; yF = yF * s->lacunarity; zF = zF * s->lacunarity;
			seedF = seedF + 1;
			ampF = ampF * s->gain;
			result = FNR_MUL_ADD(FNR_MUL_SUB(fabsf(fnr_single(base, seedF, xF, yF, zF)), 2.0f, 1.0f), ampF, result);
		}
		return result * s->fractal_bounding;
	default: /* FNR_RIGIDMULTI */
		result = 1.0f - fabsf(fnr_single(base, seedF, xF, yF, zF));
		while (++octave < s->octaves)
		{
			xF = xF * s->lacunarity; yF = yF * s->lacunarity; zF = zF * s->lacunarity;
			seedF = seedF + 1;
			ampF = ampF * s->gain;
			result = FNR_NMUL_ADD(1.0f - fabsf(fnr_single(base, seedF, xF, yF, zF)), ampF, result);
		}
		return result;
	}
}

/* FillNoiseSet(out, vectorSet, xOffset, yOffset, zOffset) with sampleScale == 0:
 * coord = fma(set[i], frequency*scale, offset*frequency*scale) */
static inline void fnr_fill_noise_set(const fnr_state* s, float* out, const float* xs, const float* ys, const float* zs,
                                      int count, float x_off, float y_off, float z_off)
{
	float xFreq = s->frequency * s->x_scale;
	float yFreq = s->frequency * s->y_scale;
	float zFreq = s->frequency * s->z_scale;
	float xOff = x_off * xFreq, yOff = y_off * yFreq, zOff = z_off * zFreq;
	for (int i = 0; i < count; i++)
	{
		float xF = FNR_MUL_ADD(xs[i], xFreq, xOff);
		float yF = FNR_MUL_ADD(ys[i], yFreq, yOff);
		float zF = FNR_MUL_ADD(zs[i], zFreq, zOff);
		out[i] = fnr_eval_scaled(s, xF, yF, zF);
	}
}

#ifdef __cplusplus
}
#endif
#endif /* FASTNOISE_REF_H */

// API-shape stand-in for Auburns/FastNoiseSIMD (external to the reference, not vendored, not
// installable offline).  Only the members the reference calls are provided
// (NoiseSampler.cpp:120-125,160-167,203-206,236-242,280-293; ColorMapper.cpp:45-48;
// ChunkBlocks.hpp:103,126).  The arithmetic behind FillNoiseSet is oracle/fastnoise_ref.h -- the
// declared, UNPINNED restatement of the library.  Test infrastructure only (oracle/_ref build).
#pragma once
#include <cstdlib>
#include "../fastnoise_ref.h"

struct FastNoiseVectorSet
{
	int size = -1;
	float* xSet = nullptr;
	float* ySet = nullptr;
	float* zSet = nullptr;
	int sampleScale = 0;
	int sampleSizeX = -1, sampleSizeY = -1, sampleSizeZ = -1;

	FastNoiseVectorSet() {}
	~FastNoiseVectorSet() { free(xSet); }
	void SetSize(int _size)
	{
		free(xSet);
		size = _size;
		xSet = (float*)aligned_alloc(64, ((sizeof(float) * 3 * (size_t)_size + 63) / 64) * 64);
		ySet = xSet + _size;
		zSet = ySet + _size;
	}
};

class FastNoiseSIMD
{
public:
	enum NoiseType { Value, ValueFractal, Perlin, PerlinFractal, Simplex, SimplexFractal, WhiteNoise, Cellular, Cubic, CubicFractal };
	enum FractalType { FBM, Billow, RigidMulti };
	enum PerturbType { None, Gradient, GradientFractal, Normalise, Gradient_Normalise, GradientFractal_Normalise };

	static FastNoiseSIMD* NewFastNoiseSIMD(int seed = 1337) { return new FastNoiseSIMD(seed); }
	static int GetSIMDLevel() { return 3; /* the restatement mirrors the AVX2/FMA3 level */ }
	static float* GetEmptySet(int x, int y, int z) { return (float*)aligned_alloc(64, ((sizeof(float) * (size_t)x * y * z + 63) / 64) * 64); }
	static void FreeNoiseSet(float* p) { free(p); }

	explicit FastNoiseSIMD(int seed = 1337) { fnr_init(&st, seed); }
	virtual ~FastNoiseSIMD() {}

	void SetSeed(int seed) { st.seed = seed; }
	void SetFrequency(float f) { st.frequency = f; }
	void SetNoiseType(NoiseType t)
	{
		switch (t)
		{
		case Value: st.noise_type = FNR_VALUE; break;
		case ValueFractal: st.noise_type = FNR_VALUE_FRACTAL; break;
		case Perlin: st.noise_type = FNR_PERLIN; break;
		case PerlinFractal: st.noise_type = FNR_PERLIN_FRACTAL; break;
		case Simplex: st.noise_type = FNR_SIMPLEX; break;
		case SimplexFractal: st.noise_type = FNR_SIMPLEX_FRACTAL; break;
		default: abort(); /* not restated: never requested by the reference */
		}
	}
	void SetFractalOctaves(int o) { fnr_set_fractal_octaves(&st, o); }
	void SetFractalGain(float g) { fnr_set_fractal_gain(&st, g); }
	void SetFractalLacunarity(float l) { st.lacunarity = l; }
	void SetFractalType(FractalType t) { st.fractal_type = (t == FBM) ? FNR_FBM : (t == Billow) ? FNR_BILLOW : FNR_RIGIDMULTI; }
	void SetPerturbType(PerturbType t)
	{
		switch (t)
		{
		case None: st.perturb_type = FNR_PERTURB_NONE; break;
		case Gradient: st.perturb_type = FNR_PERTURB_GRADIENT; break;
		case GradientFractal: st.perturb_type = FNR_PERTURB_GRADIENT_FRACTAL; break;
		default: abort();
		}
	}
	void SetPerturbFractalOctaves(int o) { fnr_set_perturb_octaves(&st, o); }
	void SetPerturbAmp(float a) { fnr_set_perturb_amp(&st, a); }
	void SetPerturbFrequency(float f) { st.perturb_frequency = f; }
	void SetPerturbFractalGain(float g) { fnr_set_perturb_gain(&st, g); }
	void SetPerturbFractalLacunarity(float l) { st.perturb_lacunarity = l; }

	virtual void FillNoiseSet(float* out, FastNoiseVectorSet* vs, float xOffset = 0.0f, float yOffset = 0.0f, float zOffset = 0.0f)
	{
		fnr_fill_noise_set(&st, out, vs->xSet, vs->ySet, vs->zSet, vs->size, xOffset, yOffset, zOffset);
	}

	fnr_state st;
};

// bmf_host.hpp -- C++ host side above the C ABI (include/bmf_b200.h): the reference's operator interface for
// the extraction path -- Sampler / DMCChunk / ChunkGenerator / Processing::MeshProcessor<N> -- with the same
// names, argument meaning, public fields and (absent) error behaviour, re-hosted on libbmf_b200.so.
//
// What each class replaces (reference file:line under BinaryMeshFitting/):
//   SmartContainer<T>              SmartContainer.hpp:8-139      POD vector (count / size / elements)
//   DualVertex                     Vertices.hpp:5-24             84-byte vertex record (layout asserted below)
//   Sampler, SamplerProperties     Sampler.hpp:8-34              callback bundle + the device descriptor a GPU needs
//   ImplicitFunctions / NoiseSamplers factories  ImplicitSampler.hpp:51-59, NoiseSampler.hpp:82-140
//   BinaryBlock / DensityBlock / VerticesIndicesBlock / ...      ChunkBlocks.hpp:9-227
//   ResourceAllocator<T>           ResourceAllocator.hpp:8-69    mutex-guarded free list of blocks
//   DMCChunk                       DMCChunk.hpp:21-85, DMCChunk.cpp:58-166, 168-508, 514-591
//   WorldProperties / WorldOctreeNode / WorldOctree (the slice process_queue touches)  WorldOctree.hpp:18-33,
//                                  WorldOctreeNode.hpp:30-170, WorldOctree.cpp:20-33, 263-269
//   ChunkGenerator                 ChunkGenerator.hpp:15-57, ChunkGenerator.cpp:19-147
//   Processing::MeshProcessor<N>   MeshProcessor.hpp:55-84, MeshProcessor.cpp:25-306
//
// Differences a maintainer must know (also in INTEGRATION.md):
//   * Sampler carries `device` (bmf_sampler_desc).  The factories below fill it.  A Sampler whose `block`
//     is an arbitrary host callback keeps working through kind = BMF_SAMPLER_HOST_DENSITY: the shim calls
//     `block` on the host, uploads the density block and the GPU does everything after it.
//   * ChunkGenerator::process_queue runs the whole batch as ONE device submission (the reference loops
//     chunks under `#pragma omp parallel for`); results land in the same fields (chunk->vi, contains_mesh, ...).
//   * DMCChunk's staged calls stay valid, but label_grid already runs the fused device pipeline for the chunk;
//     label_edges / polygonize publish its results (vertices, then indices + init_valence) at the same points.
//   * There is no CPU fallback: without libbmf_b200.so / a CUDA device the constructors report the error and
//     every compute call returns false.
#pragma once

#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <algorithm>
#include <array>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/bmf_b200.h"

#if defined(__has_include)
#if __has_include(<glm/glm.hpp>)
#include <glm/glm.hpp>
#define BMF_HAVE_GLM 1
#endif
#endif
#ifndef BMF_HAVE_GLM
namespace glm
{
// the few glm types the interface exposes, for builds without GLM
template <typename T>
struct tvec3
{
	T x, y, z;
	tvec3() : x(0), y(0), z(0) {}
	tvec3(T a, T b, T c) : x(a), y(b), z(c) {}
	T& operator[](int i) { return (&x)[i]; }
	const T& operator[](int i) const { return (&x)[i]; }
};
typedef tvec3<float> vec3;
typedef tvec3<int> ivec3;
inline vec3 operator+(const vec3& a, float s) { return vec3(a.x + s, a.y + s, a.z + s); }
inline vec3 operator-(const vec3& a, float s) { return vec3(a.x - s, a.y - s, a.z - s); }
} // namespace glm
#endif

// ---- SmartContainer ------------------------------------------------------------------------------------
template <class T>
class SmartContainer
{
public:
	size_t size;
	size_t count;
	T* elements;
	int scale;

	SmartContainer() : size(0), count(0), elements(nullptr), scale(2) {}
	SmartContainer(const SmartContainer&) = delete;
	SmartContainer& operator=(const SmartContainer&) = delete;
	~SmartContainer() { reset(); }

	T& operator[](int index) { return elements[index]; }
	bool resize(size_t new_size)
	{
		T* p = static_cast<T*>(std::realloc((void*)elements, sizeof(T) * new_size));
		if (!p && new_size) return false;
		elements = p;
		size = new_size;
		return true;
	}
	bool prepare(size_t amount)
	{
		if (count + amount <= size) return true;
		size_t want = 32;
		while (want < size + amount) want *= 2;
		return resize(want);
	}
	bool prepare_exact(size_t amount) { return count + amount <= size ? true : resize(size + amount); }
	bool push_back(const T& v)
	{
		if (count >= size && !resize(size ? size * (size_t)scale : 32)) return false;
		elements[count++] = v;
		return true;
	}
	bool push_back(const T* other, size_t n)
	{
		if (!n) return true;
		if (!prepare(n)) return false;
		std::memcpy((void*)(elements + count), (const void*)other, sizeof(T) * n);
		count += n;
		return true;
	}
	bool push_back(const SmartContainer<T>& other) { return push_back(other.elements, other.count); }
	void zero()
	{
		if (elements && size) std::memset((void*)elements, 0, sizeof(T) * size);
	}
	void reset()
	{
		std::free((void*)elements);
		elements = nullptr;
		size = count = 0;
	}
};

// ---- vertex record -------------------------------------------------------------------------------------
struct DualVertex
{
	bool boundary;
	uint8_t mask;
	uint32_t index;
	uint8_t valence;
	uint8_t init_valence;
	uint8_t adj_next;
	uint32_t adj_offset;
	uint16_t edge_mask;
	float s;
	glm::ivec3 xyz;
	glm::vec3 p;
	glm::vec3 n;
	glm::vec3 avg;
	glm::vec3 color;
};
static_assert(sizeof(DualVertex) == 84, "DualVertex must keep the reference's 84-byte layout");
static_assert(offsetof(DualVertex, p) == 36 && offsetof(DualVertex, n) == 48 && offsetof(DualVertex, color) == 72, "DualVertex field offsets");

// ---- Sampler --------------------------------------------------------------------------------------------
struct FastNoiseVectorSet; // opaque here: the noise library is replaced by the device kernels
class FastNoiseSIMD;

class SamplerProperties
{
public:
	int thread_id;
	SamplerProperties() : thread_id(0) {}
	explicit SamplerProperties(int t) : thread_id(t) {}
	virtual ~SamplerProperties() {}
};

typedef const float (*SamplerValueFunction)(const float world_size, const glm::vec3& p);
typedef std::function<void(const float world_size, const glm::vec3& p, const glm::ivec3& size, const float scale, void** out, FastNoiseVectorSet* vectorset_out,
                           float* dest_noise, int offset, int stride, SamplerProperties* properties)>
	SamplerBlockFunction;
typedef std::function<glm::vec3(const float world_size, const glm::vec3& p, float h)> SamplerGradientFunction;

struct Sampler
{
	float world_size;
	SamplerValueFunction value;
	SamplerBlockFunction block;
	SamplerGradientFunction gradient;
	FastNoiseSIMD* noise_samplers[8];
	bmf_sampler_desc device; // what the GPU runs; kind == BMF_SAMPLER_HOST_DENSITY -> `block` is called on the host

	Sampler() : world_size(256.0f), value(nullptr)
	{
		for (int i = 0; i < 8; i++) noise_samplers[i] = nullptr;
		bmf_sampler_defaults(&device, BMF_SAMPLER_HOST_DENSITY);
	}
	virtual ~Sampler() {}
};

namespace NoiseSamplers
{
class NoiseSamplerProperties : public SamplerProperties
{
public:
	int level;
	float g_scale, height;
	int octaves;
	float amp, frequency, gain;
	int noise_type, fractal_type;
	NoiseSamplerProperties() : SamplerProperties(), level(0), g_scale(0.25f), height(75.0f), octaves(13), amp(0.87f), frequency(0.585f), gain(0.488f), noise_type(1), fractal_type(0) {}
};
inline const float noise3d(const float, const glm::vec3&) { return 0; }
inline void make(Sampler* s, int kind)
{
	s->value = noise3d;
	// s->gradient = std::bind(implicit_gradient, noise3d, _1, _2, _3) (NoiseSampler.hpp:86-136): differences of the constant 0
	s->gradient = [](const float, const glm::vec3&, float) { return glm::vec3(0.0f - 0.0f, 0.0f - 0.0f, 0.0f - 0.0f); };
	bmf_sampler_defaults(&s->device, kind);
	s->device.world_size = s->world_size;
}
inline void create_sampler_terrain_2d(Sampler* s) { make(s, BMF_SAMPLER_TERRAIN2D); }
inline void create_sampler_terrain_pert_2d(Sampler* s) { make(s, BMF_SAMPLER_TERRAIN2D_PERT); }
inline void create_sampler_terrain_3d(Sampler* s) { make(s, BMF_SAMPLER_TERRAIN3D); }
inline void create_sampler_terrain_pert_3d(Sampler* s) { make(s, BMF_SAMPLER_TERRAIN3D_PERT); }
} // namespace NoiseSamplers

namespace ImplicitFunctions
{
// value callbacks keep their reference meaning (density = -SDF, positive inside); they are only evaluated on the
// host by callers that use Sampler::value / gradient directly -- the device evaluates its own copy per voxel
inline const float sphere(const float r, const glm::vec3& p) { return -(std::sqrt(p.x * p.x + p.y * p.y + p.z * p.z) - r * 0.25f); }
inline const float torus_z(const float r, const glm::vec3& p)
{
	float qx = std::fabs(std::sqrt(p.x * p.x + p.y * p.y)) - r / 4.0f;
	return -(std::sqrt(qx * qx + p.z * p.z) - r / 10.0f);
}
inline const float cuboid(const float res, const glm::vec3& p)
{
	float r = res / 8.0f, dx = std::fabs(p.x) - r, dy = std::fabs(p.y) - r, dz = std::fabs(p.z) - r;
	return -std::fmin(std::fmax(dx, std::fmax(dy, dz)), std::sqrt(dx * dx + dy * dy + dz * dz));
}
inline const float plane_y(const float, const glm::vec3& p) { return -p.y; }
inline glm::vec3 implicit_gradient(SamplerValueFunction f, const float res, const glm::vec3& p, float h = 0.01f)
{
	return glm::vec3(f(res, glm::vec3(p.x + h, p.y, p.z)) - f(res, glm::vec3(p.x - h, p.y, p.z)), f(res, glm::vec3(p.x, p.y + h, p.z)) - f(res, glm::vec3(p.x, p.y - h, p.z)),
	                 f(res, glm::vec3(p.x, p.y, p.z + h)) - f(res, glm::vec3(p.x, p.y, p.z - h)));
}
inline Sampler create_sampler(const SamplerValueFunction& f)
{
	Sampler s;
	s.value = f;
	s.gradient = [f](const float ws, const glm::vec3& p, float h) { return implicit_gradient(f, ws, p, h); };
	int kind = f == sphere ? BMF_SAMPLER_SPHERE : f == torus_z ? BMF_SAMPLER_TORUS_Z : f == cuboid ? BMF_SAMPLER_CUBOID : f == plane_y ? BMF_SAMPLER_PLANE_Y : BMF_SAMPLER_HOST_DENSITY;
	bmf_sampler_defaults(&s.device, kind);
	if (kind == BMF_SAMPLER_HOST_DENSITY)
	{
		// unknown analytic function: evaluate it on the host like implicit_block (ImplicitSampler.hpp:14-36)
		s.block = [f](const float ws, const glm::vec3& p, const glm::ivec3& size, const float scale, void** out, FastNoiseVectorSet*, float*, int, int, SamplerProperties*) {
			float* o = (float*)*out;
			for (int x = 0; x < size.x; x++)
				for (int y = 0; y < size.y; y++)
					for (int z = 0; z < size.z; z++) o[((size_t)x * size.y + y) * size.z + z] = f(ws, glm::vec3(p.x + (float)x * scale, p.y + (float)y * scale, p.z + (float)z * scale));
		};
	}
	return s;
}
} // namespace ImplicitFunctions

// ---- pooled blocks ---------------------------------------------------------------------------------------
template <typename T>
struct PodBlock
{
	uint32_t size = 0;
	T* data = nullptr;
	bool initialized = false;
	~PodBlock() { std::free(data); }
	void init(uint32_t n)
	{
		if (initialized && n <= size) return;
		std::free(data);
		data = (T*)std::malloc(sizeof(T) * (size_t)n);
		size = n;
		initialized = true;
	}
};
struct BinaryBlock : PodBlock<uint32_t>
{
	void init(uint32_t /*raw*/, uint32_t binary_size) { PodBlock<uint32_t>::init(binary_size); }
};
struct DensityBlock : PodBlock<float> {};
struct MasksBlock : PodBlock<uint64_t> {};
struct IndexesBlock : PodBlock<uint32_t> {};
struct NoiseBlock : PodBlock<float> {};
struct DMC_Cell { uint16_t mask; };
struct DMC_CellsBlock
{
	SmartContainer<DMC_Cell> cells; // only the count is meaningful: the device path has no per-cell records
	void init() { cells.count = 0; }
};
struct VerticesIndicesBlock
{
	SmartContainer<DualVertex> vertices;
	SmartContainer<uint32_t> mesh_indexes;
	void init() { vertices.count = 0; mesh_indexes.count = 0; }
};

template <class T>
class ResourceAllocator
{
public:
	~ResourceAllocator()
	{
		for (T* e : all) delete e;
	}
	T* new_element(bool no_lock = false)
	{
		std::unique_lock<std::mutex> l(_mutex, std::defer_lock);
		if (!no_lock) l.lock();
		if (!free_list.empty())
		{
			T* e = free_list.back();
			free_list.pop_back();
			return e;
		}
		T* e = new T();
		all.push_back(e);
		return e;
	}
	void free_element(T* e, bool no_lock = false)
	{
		if (!e) return;
		std::unique_lock<std::mutex> l(_mutex, std::defer_lock);
		if (!no_lock) l.lock();
		free_list.push_back(e);
	}
	std::mutex _mutex;

private:
	std::vector<T*> all, free_list;
};

// ---- device context shared by the classes below ------------------------------------------------------------
class BmfDevice
{
public:
	// slot 0 = the default context on `device`; further slots are the extra contexts of the multi-GPU generator
	// (one per GPU, or several on one GPU for testing)
	static BmfDevice& get(int device = 0) { return slot(0, device); }
	static BmfDevice& slot(int index, int device)
	{
		static std::mutex m;
		static std::vector<BmfDevice*> slots;
		std::lock_guard<std::mutex> lock(m);
		if ((int)slots.size() <= index) slots.resize(index + 1, nullptr);
		if (!slots[index]) slots[index] = new BmfDevice(device); // lives for the process (contexts are cheap to keep)
		return *slots[index];
	}
	bmf_ctx* ctx = nullptr;
	int device = 0;
	// a bmf_ctx is single-threaded (bmf_b200.h); callers that share this context from several threads -- the reference runs
	// label_grid / label_edges / polygonize per chunk inside `#pragma omp parallel for` (ChunkGenerator.cpp:93-108) -- hold this lock
	// for one whole device transaction (sampler + submit + fetch)
	std::mutex lock;
	bool ok() const { return ctx != nullptr; }
	const char* error() const { return ctx ? bmf_last_error(ctx) : "no CUDA device / libbmf_b200 context (no CPU fallback)"; }

private:
	explicit BmfDevice(int dev) : device(dev)
	{
		if (bmf_ctx_create(dev, &ctx) != BMF_OK) ctx = nullptr;
	}
	~BmfDevice() { bmf_ctx_destroy(ctx); }
};

// Sampler::gradient (Sampler.hpp:29) for many points at once on the device: the same six evaluations and raw differences per point
// as implicit_gradient (ImplicitSampler.hpp:38-49), bit-identical to calling s.gradient(world_size, p[i], h) on the host for the
// samplers the device knows (s.device.kind); false for host-callback samplers (call s.gradient per point instead).
inline bool sampler_gradient_block(const Sampler& s, const glm::vec3* p, size_t m, float h, glm::vec3* out, int device = 0)
{
	static_assert(sizeof(glm::vec3) == 3 * sizeof(float), "glm::vec3 must be three packed floats");
	if (s.device.kind == BMF_SAMPLER_HOST_DENSITY) return false;
	BmfDevice& dev = BmfDevice::get(device);
	if (!dev.ok()) return false;
	std::lock_guard<std::mutex> guard(dev.lock);
	bmf_sampler_desc d = s.device;
	d.world_size = s.world_size;
	if (bmf_sampler_set(dev.ctx, &d) != BMF_OK) return false;
	return bmf_sampler_gradient(dev.ctx, reinterpret_cast<const float*>(p), (int64_t)m, h, reinterpret_cast<float*>(out)) == BMF_OK;
}

namespace bmf_detail
{
inline void fill_dual_vertices(SmartContainer<DualVertex>& out, const float* pos, const float* nrm, const float* col, const uint8_t* bnd, const uint8_t* val, size_t n,
                               bool with_valence, bool processed)
{
	out.count = 0;
	out.prepare(n);
	uint32_t off = 0;
	for (size_t i = 0; i < n; i++)
	{
		DualVertex v;
		std::memset((void*)&v, 0, sizeof(v));
		v.boundary = bnd[i] != 0;
		v.index = (uint32_t)i;
		v.init_valence = with_valence ? val[i] : 0;
		v.valence = processed ? val[i] : 0;
		v.adj_next = processed ? val[i] : 0;
		v.adj_offset = processed ? off : 0;
		v.p = glm::vec3(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
		if (nrm) v.n = glm::vec3(nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]);
		v.color = glm::vec3(col[3 * i], col[3 * i + 1], col[3 * i + 2]);
		out.elements[out.count++] = v;
		off += val[i];
	}
}
} // namespace bmf_detail

// ---- DMCChunk ----------------------------------------------------------------------------------------------
class DMCChunk
{
public:
	bool pem = false;
	int id = 0;
	int level = 0;
	uint32_t dim = 32; // RESOLUTION (DMCChunk.cpp:34)
	float size = 0;
	glm::vec3 pos;
	bool contains_mesh = false;
	uint32_t mesh_offset = 0;
	uint64_t parent_code = 0;
	Sampler sampler;
	VerticesIndicesBlock* vi = nullptr;
	DensityBlock* density_block = nullptr;
	BinaryBlock* binary_block = nullptr;
	DMC_CellsBlock* cell_block = nullptr;
	IndexesBlock* indexes_block = nullptr;
	glm::vec3 overlap_pos, bound_start;
	float bound_size = 0, scale = 0;

	DMCChunk() {}
	DMCChunk(glm::vec3 p, float s, int lvl, Sampler& smp, uint64_t code) { init(p, s, lvl, smp, code); }

	void init(glm::vec3 p, float s, int lvl, Sampler& smp, uint64_t code)
	{
		pos = p; dim = 32; size = s; level = lvl; pem = false; contains_mesh = false; mesh_offset = 0; sampler = smp;
		vi = nullptr; cell_block = nullptr; indexes_block = nullptr; density_block = nullptr; binary_block = nullptr; parent_code = code;
	}

	// sample + sign pack (+ the fused device pipeline of this one chunk).  Fills binary_block, density_block,
	// contains_mesh, overlap_pos, scale, bound_start, bound_size like DMCChunk.cpp:79-166.
	bool label_grid(ResourceAllocator<BinaryBlock>* binary_allocator, ResourceAllocator<DensityBlock>* density_allocator, ResourceAllocator<NoiseBlock>* /*noise_allocator*/, float overlap,
	                NoiseSamplers::NoiseSamplerProperties properties)
	{
		BmfDevice& dev = BmfDevice::get();
		if (!dev.ok()) return false;
		const size_t n = (size_t)dim * dim * dim;
		binary_block = binary_allocator->new_element();
		binary_block->init((uint32_t)n, (uint32_t)(n / 32));
		density_block = density_allocator->new_element();
		density_block->init((uint32_t)n);
		const float delta = size * (1.0f + overlap * 2.0f) / (float)(dim - 1);
		overlap_pos = pos - size * overlap;
		scale = delta;
		bound_size = size * (1.0f + overlap * 2.0f) * 0.5f;
		bound_start = overlap_pos + bound_size;

		bmf_sampler_desc d = sampler.device;
		d.world_size = sampler.world_size;
		if (d.kind == BMF_SAMPLER_TERRAIN2D_PERT)
		{
			d.g_scale = properties.g_scale; d.height = properties.height; d.octaves = properties.octaves;
			d.amp = properties.amp; d.frequency = properties.frequency; d.gain = properties.gain;
		}
		const float* host_density = nullptr;
		if (d.kind == BMF_SAMPLER_HOST_DENSITY)
		{
			if (!sampler.block) return false;
			void* out = density_block->data;
			sampler.block(sampler.world_size, overlap_pos, glm::ivec3((int)dim, (int)dim, (int)dim), delta, &out, nullptr, nullptr, 0, sizeof(float), &properties);
			host_density = density_block->data;
		}
		// One device transaction under the context lock: the whole pipeline of THIS chunk runs here and everything the later
		// stages publish (cell count, vertex records, indices) is fetched into the chunk itself, so label_edges / polygonize are
		// host-only and chunks may be interleaved (grid A, grid B, edges A) or driven from several threads like the reference's
		// OpenMP loop without ever reading another chunk's resident batch.
		std::lock_guard<std::mutex> guard(dev.lock);
		if (bmf_sampler_set(dev.ctx, &d) != BMF_OK) return false;
		bmf_chunk_desc c;
		c.pos[0] = pos.x; c.pos[1] = pos.y; c.pos[2] = pos.z; c.size = size; c.level = level; c.overlap = overlap; c.morton = parent_code;
		bmf_params p;
		std::memset(&p, 0, sizeof(p));
		p.dim = (int32_t)dim;
		p.keep_density = 1;
		if (bmf_batch_submit(dev.ctx, &c, 1, &p, host_density) != BMF_OK) return false;
		bmf_chunk_info info;
		if (bmf_batch_chunk_info(dev.ctx, 0, &info) != BMF_OK) return false;
		contains_mesh = info.contains_mesh != 0;
		staged_cells = (size_t)info.n_cells;
		staged_verts.assign((size_t)info.n_verts * 84, 0);
		staged_inds.assign((size_t)info.n_inds, 0u);
		staged = true;
		return bmf_batch_copy_chunk(dev.ctx, 0, staged_verts.empty() ? nullptr : staged_verts.data(), staged_inds.empty() ? nullptr : staged_inds.data(), binary_block->data, nullptr,
		                            d.kind == BMF_SAMPLER_HOST_DENSITY ? nullptr : density_block->data) == BMF_OK;
	}

	// publishes the iso-vertices (positions, boundary flags; valences still 0) -- DMCChunk.cpp:168-508
	bool label_edges(ResourceAllocator<VerticesIndicesBlock>* vi_allocator, ResourceAllocator<DMC_CellsBlock>* cell_allocator, ResourceAllocator<IndexesBlock>* /*inds_allocator*/,
	                 ResourceAllocator<DensityBlock>* /*density_allocator*/, ResourceAllocator<MasksBlock>* /*masks_allocator*/)
	{
		if (!contains_mesh) return true;
		if (!staged) return false; // label_grid of THIS chunk has not run
		if (!vi) { vi = vi_allocator->new_element(); vi->init(); }
		if (!cell_block) { cell_block = cell_allocator->new_element(); cell_block->init(); }
		cell_block->cells.count = 0;
		cell_block->cells.prepare(staged_cells);
		cell_block->cells.count = staged_cells;
		return publish_vertices(false);
	}

	// publishes the index buffer and init_valence -- DMCChunk.cpp:514-576
	bool polygonize()
	{
		if (!contains_mesh) return true;
		if (!staged || !vi) return false;
		vi->mesh_indexes.count = 0;
		vi->mesh_indexes.prepare(staged_inds.size());
		if (!staged_inds.empty()) std::memcpy((void*)vi->mesh_indexes.elements, staged_inds.data(), staged_inds.size() * sizeof(uint32_t));
		vi->mesh_indexes.count = staged_inds.size();
		return publish_vertices(true);
	}

	void copy_verts_and_inds(SmartContainer<DualVertex>& v_out, SmartContainer<uint32_t>& i_out)
	{
		if (!contains_mesh || !vi) return;
		mesh_offset = (uint32_t)v_out.count;
		size_t start = i_out.count;
		v_out.push_back(vi->vertices);
		i_out.push_back(vi->mesh_indexes);
		for (size_t i = start; i < i_out.count; i++) i_out.elements[i] += mesh_offset;
	}

private:
	// what label_grid fetched from the device for this chunk (84-byte DualVertex records, index buffer, active-cell count)
	bool staged = false;
	size_t staged_cells = 0;
	std::vector<uint8_t> staged_verts;
	std::vector<uint32_t> staged_inds;

	bool publish_vertices(bool with_valence)
	{
		const size_t nv = staged_verts.size() / 84;
		vi->vertices.count = 0;
		vi->vertices.prepare(nv);
		if (nv) std::memcpy((void*)vi->vertices.elements, staged_verts.data(), staged_verts.size());
		vi->vertices.count = nv;
		if (!with_valence)
			for (size_t i = 0; i < vi->vertices.count; i++) vi->vertices.elements[i].init_valence = 0;
		return true;
	}
};

// ---- the slice of the world the generator touches -------------------------------------------------------------
enum GENERATION_STAGES
{
	GENERATION_STAGES_UNHANDLED = 0, GENERATION_STAGES_WATCHER_QUEUED = 1, GENERATION_STAGES_GENERATOR_QUEUED = 2, GENERATION_STAGES_GENERATOR_ACKNOWLEDGED = 3,
	GENERATION_STAGES_GENERATING = 4, GENERATION_STAGES_NEEDS_FORMAT = 5, GENERATION_STAGES_NEEDS_UPLOAD = 6, GENERATION_STAGES_UPLOADING = 7,
	GENERATION_STAGES_AWAITING_STITCHING = 8, GENERATION_STAGES_AWAITING_STITCHING_UPLOAD = 9, GENERATION_STAGES_DONE = 10
};

struct WorldProperties
{
	float split_multiplier = 1.0f, group_multiplier = 2.0f, size_modifier = 0.0f;
	int max_level = 7, min_level = 1, num_threads = 4, process_iters = 0, chunk_resolution = 32;
	bool enable_stitching = false;
	float overlap = 0.035f;
	bool boundary_processing = false;
};

// GLChunk's CPU side: the SoA the renderer uploads (GLChunk.hpp:32-34, format_data GLChunk.cpp:278-296)
struct GLChunk
{
	SmartContainer<glm::vec3> p_data, n_data, c_data;
};

class WorldOctreeNode
{
public:
	float size = 0;
	uint8_t level = 0;
	glm::vec3 pos;
	uint64_t morton_code = 0;
	int generation_stage = GENERATION_STAGES_UNHANDLED;
	DMCChunk* chunk = nullptr;
	GLChunk* gl_chunk = nullptr;
	// octree links (OctreeNode::parent / children, WorldOctreeNode::world_leaf_flag) and the watcher's split / group mark
	WorldOctreeNode* parent = nullptr;
	WorldOctreeNode* children[8] = { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr };
	bool world_leaf_flag = true;
	int flags = 0; // NODE_FLAGS_SPLIT = 1, NODE_FLAGS_GROUP = 2 (the subset the synchronous watcher needs)
	WorldOctreeNode() {}
	WorldOctreeNode(float s, glm::vec3 p, uint8_t l, uint64_t code) : size(s), level(l), pos(p), morton_code(code) {}
};

// Depth-normalised Z-curve order of WorldOctreeNode::morton_code values (sentinel 1, then 3 bits per level).  The raw codes
// order level-major (deeper nodes are numerically larger), which scatters a device's share over the LODs; stripping the
// sentinel and left-aligning the digits orders leaves of mixed levels along one curve.  Ties break by level.
inline bool morton_less(uint64_t a, uint64_t b)
{
	auto level_of = [](uint64_t c) { int bits = 0; while (c >> bits) bits++; return (bits - 1) / 3; };
	const int la = a ? level_of(a) : 0, lb = b ? level_of(b) : 0;
	const uint64_t ka = (a ^ (a ? (uint64_t)1 << (3 * la) : 0)) << (3 * (20 - la)), kb = (b ^ (b ? (uint64_t)1 << (3 * lb) : 0)) << (3 * (20 - lb));
	return ka != kb ? ka < kb : la < lb;
}

class WorldOctree
{
public:
	Sampler sampler;
	WorldProperties properties;
	NoiseSamplers::NoiseSamplerProperties noise_properties;
	std::mutex chunk_mutex;
	int next_chunk_id = 0;
	std::vector<DMCChunk*> chunks;
	std::vector<WorldOctreeNode*> leaves; // split_leaves() result, in the reference's order (LIFO DFS, not Morton order)
	glm::vec3 focus_point;
	WorldOctreeNode octree; // root

	~WorldOctree()
	{
		for (DMCChunk* c : chunks) delete c;
		for (WorldOctreeNode* n : owned) delete n;
	}

	// WorldOctree::init (WorldOctree.cpp:117-127): root cube of twice the world size centred on the origin, code 1
	void init(uint32_t size)
	{
		size *= 2;
		const float p = 1.0f * (float)size * -0.5f;
		octree = WorldOctreeNode((float)size, glm::vec3(p, p, p), 0, 1);
	}

	// WorldOctree::node_needs_split (WorldOctree.cpp:239-249); middle = pos + size*0.5 (WorldOctreeNode.cpp:26)
	bool node_needs_split(const glm::vec3& center, const WorldOctreeNode* n) const
	{
		if (n->level >= properties.max_level) return false;
		if (n->level < properties.min_level) return true;
		const float half = n->size * 0.5f;
		const float mx = n->pos.x + half, my = n->pos.y + half, mz = n->pos.z + half;
		const float dx = center.x - mx, dy = center.y - my, dz = center.z - mz;
		const float t0 = dx * dx, t1 = dy * dy, t2 = dz * dz;
		const float d = std::sqrt(t0 + t1 + t2);
		return d < n->size * properties.split_multiplier + properties.size_modifier + n->size * 0.5f;
	}

	// WorldOctree::split_leaves (WorldOctree.cpp:129-173) + split_node (:175-210): static LOD build; children in
	// MC-corner order (Tables.hpp:7-9), Morton digit x | y<<1 | z<<2, chunk created for every leaf
	void split_leaves()
	{
		leaves.clear();
		std::vector<WorldOctreeNode*> stack;
		stack.push_back(&octree);
		while (!stack.empty())
		{
			WorldOctreeNode* n = stack.back();
			stack.pop_back();
			if (!node_needs_split(focus_point, n))
			{
				leaves.push_back(n);
				continue;
			}
			split_node(n);
			for (int i = 0; i < 8; i++) stack.push_back(n->children[i]);
		}
		next_chunk_id = 0;
		for (WorldOctreeNode* n : leaves) create_chunk(n);
	}

	// WorldOctree::split_node (WorldOctree.cpp:175-210): eight children in MC-corner order, Morton digit x | y<<1 | z<<2
	void split_node(WorldOctreeNode* n)
	{
		static const int MCDX[8] = { 0, 1, 1, 0, 0, 1, 1, 0 }, MCDY[8] = { 0, 0, 0, 0, 1, 1, 1, 1 }, MCDZ[8] = { 0, 0, 1, 1, 0, 0, 1, 1 };
		const float c_size = n->size * 0.5f;
		for (int i = 0; i < 8; i++)
		{
			glm::vec3 c_pos(n->pos.x + (float)MCDX[i] * c_size, n->pos.y + (float)MCDY[i] * c_size, n->pos.z + (float)MCDZ[i] * c_size);
			const uint64_t code = (n->morton_code << 3) | (uint64_t)(MCDX[i] | (MCDY[i] << 1) | (MCDZ[i] << 2));
			WorldOctreeNode* c = new WorldOctreeNode(c_size, c_pos, (uint8_t)(n->level + 1), code);
			c->parent = n;
			owned.push_back(c);
			n->children[i] = c;
		}
		n->world_leaf_flag = false;
	}

	// WorldOctree::group_node: the children go away, the node is a leaf again (the nodes stay owned until the world dies)
	void group_node(WorldOctreeNode* n)
	{
		for (int i = 0; i < 8; i++) n->children[i] = nullptr;
		n->world_leaf_flag = true;
	}

	// WorldOctree::node_needs_group (WorldOctree.cpp:251-261)
	bool node_needs_group(const glm::vec3& center, const WorldOctreeNode* n) const
	{
		if (n->level < properties.min_level) return false;
		if (n->level > properties.max_level) return true;
		const float half = n->size * 0.5f;
		const float mx = n->pos.x + half, my = n->pos.y + half, mz = n->pos.z + half;
		const float dx = center.x - mx, dy = center.y - my, dz = center.z - mz;
		const float t0 = dx * dx, t1 = dy * dy, t2 = dz * dz;
		const float d = std::sqrt(t0 + t1 + t2);
		return d > n->size * properties.group_multiplier + properties.size_modifier + n->size * 0.5f;
	}

	void create_chunk(WorldOctreeNode* n)
	{
		n->chunk = new DMCChunk();
		chunks.push_back(n->chunk);
		n->chunk->init(n->pos, n->size, n->level, sampler, n->morton_code);
		n->chunk->dim = (uint32_t)properties.chunk_resolution;
		n->chunk->id = next_chunk_id++;
	}

private:
	std::vector<WorldOctreeNode*> owned;
};

// ---- WorldStitcher ----------------------------------------------------------------------------------------------
// WorldStitcher.hpp:15-77.  stitch_all(root) (WorldStitcher.cpp:26-49) walks the world octree for dual cells between
// leaves and polygonises them into `vertices`, a non-indexed DualVertex triangle soup (:568-570) that format() turns
// into the SoA gl_chunk (GLChunk::format_data_tris).  Here the leaves go to the device as one batch and
// bmf_batch_stitch does the work (csrc/seam.cuh); behaviour is build-defined (the reference's version is
// non-functional as committed).  Runs on the first device of the generator; with several GPUs this is the pass after
// the host gather.
class WorldStitcher
{
public:
	SmartContainer<DualVertex> vertices;
	GLChunk gl_chunk;
	int stage = 0; // STITCHING_STAGES_READY
	float last_ms[2] = { 0, 0 }; // device time of the last pass: count+scan, emit

	void init() {}

	bool stitch_all(WorldOctree* world, int device_slot = 0, int device_id = 0)
	{
		vertices.count = 0;
		if (!world || world->leaves.empty()) return true;
		return stitch_batch_nodes(world, world->leaves.data(), world->leaves.size(), device_slot, device_id);
	}

	bool stitch_batch_nodes(WorldOctree* world, WorldOctreeNode* const* nodes, size_t count, int device_slot = 0, int device_id = 0)
	{
		vertices.count = 0;
		std::vector<float> tris;
		if (!seam_pass(world, nodes, count, nullptr, false, device_slot, device_id, tris, last_ms)) return false;
		return append_soup(tris);
	}

	// Several GPUs ("a final host gather ... and a WorldStitcher seam pass"): the leaves are dealt to the
	// devices in contiguous Morton-ordered ranges like ChunkGenerator::process_queue does; every device stitches the dual
	// cells that lie inside its own chunks (one host thread per device), then the first device stitches the cells that span
	// devices from the border chunks only (group ids + cross_group_only).  Chunks are pure functions of their descriptors, so
	// that device re-samples the border chunks instead of receiving them: no collective.  The union of the passes is exactly
	// the single-device seam (as a set of triangles; the order is per device, then the cross pass).
	bool stitch_all(WorldOctree* world, const std::vector<int>& devices)
	{
		vertices.count = 0;
		if (!world || world->leaves.empty()) return true;
		const int n_dev = (int)devices.size();
		if (n_dev <= 1) return stitch_all(world, 0, devices.empty() ? 0 : devices[0]);
		const std::vector<WorldOctreeNode*>& leaves = world->leaves;
		const size_t n = leaves.size();
		std::vector<int> order(n);
		for (size_t i = 0; i < n; i++) order[i] = (int)i;
		std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return morton_less(leaves[a]->morton_code, leaves[b]->morton_code); });
		std::vector<int32_t> group(n, 0);
		for (int k = 0; k < n_dev; k++)
			for (size_t q = n * k / n_dev; q < n * (k + 1) / n_dev; q++) group[order[q]] = k;
		std::vector<std::vector<WorldOctreeNode*>> part(n_dev);
		for (size_t i = 0; i < n; i++) part[group[i]].push_back(leaves[i]); // batch order inside a part
		std::vector<std::vector<float>> soup(n_dev + 1);
		std::vector<char> ok(n_dev, 1);
		std::vector<std::thread> th;
		for (int k = 0; k < n_dev; k++)
			th.emplace_back([&, k] {
				float ms[2];
				ok[k] = part[k].empty() ? 1 : (seam_pass(world, part[k].data(), part[k].size(), nullptr, false, k, devices[k], soup[k], ms) ? 1 : 0);
			});
		for (std::thread& t : th) t.join();
		for (char r : ok)
			if (!r) return false;
		// border chunks: those that touch (face, edge or corner) a chunk of another group
		std::vector<WorldOctreeNode*> border;
		std::vector<int32_t> border_group;
		{
			float smin = leaves[0]->size, org[3] = { leaves[0]->pos.x, leaves[0]->pos.y, leaves[0]->pos.z };
			for (const WorldOctreeNode* l : leaves)
			{
				smin = std::min(smin, l->size);
				org[0] = std::min(org[0], l->pos.x); org[1] = std::min(org[1], l->pos.y); org[2] = std::min(org[2], l->pos.z);
			}
			std::vector<std::array<long long, 4>> lat(n); // origin (slots), extent
			long long G[3] = { 0, 0, 0 };
			for (size_t i = 0; i < n; i++)
			{
				const WorldOctreeNode* l = leaves[i];
				const long long e = std::llround((double)l->size / smin);
				lat[i] = { std::llround(((double)l->pos.x - org[0]) / smin), std::llround(((double)l->pos.y - org[1]) / smin), std::llround(((double)l->pos.z - org[2]) / smin), e };
				for (int a = 0; a < 3; a++) G[a] = std::max(G[a], lat[i][a] + e);
			}
			std::vector<int32_t> grid((size_t)(G[0] * G[1] * G[2]), -1);
			for (size_t i = 0; i < n; i++)
				for (long long x = lat[i][0]; x < lat[i][0] + lat[i][3]; x++)
					for (long long y = lat[i][1]; y < lat[i][1] + lat[i][3]; y++)
						for (long long z = lat[i][2]; z < lat[i][2] + lat[i][3]; z++) grid[(size_t)((x * G[1] + y) * G[2] + z)] = group[i];
			for (size_t i = 0; i < n; i++)
			{
				bool touches = false;
				for (long long x = std::max(0LL, lat[i][0] - 1); x < std::min(G[0], lat[i][0] + lat[i][3] + 1) && !touches; x++)
					for (long long y = std::max(0LL, lat[i][1] - 1); y < std::min(G[1], lat[i][1] + lat[i][3] + 1) && !touches; y++)
						for (long long z = std::max(0LL, lat[i][2] - 1); z < std::min(G[2], lat[i][2] + lat[i][3] + 1) && !touches; z++)
						{
							const int32_t g = grid[(size_t)((x * G[1] + y) * G[2] + z)];
							touches = g >= 0 && g != group[i];
						}
				if (touches)
				{
					border.push_back(leaves[i]);
					border_group.push_back(group[i]);
				}
			}
		}
		if (!border.empty() && !seam_pass(world, border.data(), border.size(), border_group.data(), true, 0, devices[0], soup[n_dev], last_ms)) return false;
		for (const std::vector<float>& t : soup)
			if (!append_soup(t)) return false;
		return true;
	}

private:
	// one device: submit the nodes at voxel-node centres (signs and border samples only) and run the seam pass
	static bool seam_pass(WorldOctree* world, WorldOctreeNode* const* nodes, size_t count, const int32_t* group, bool cross_only, int device_slot, int device_id,
	                      std::vector<float>& tris, float ms[2])
	{
		BmfDevice& dev = BmfDevice::slot(device_slot, device_id);
		if (!dev.ok()) return false;
		std::vector<bmf_chunk_desc> descs(count);
		const int dim = world->properties.chunk_resolution;
		for (size_t k = 0; k < count; k++)
		{
			const WorldOctreeNode* n = nodes[k];
			bmf_chunk_desc& c = descs[k];
			c.pos[0] = n->pos.x; c.pos[1] = n->pos.y; c.pos[2] = n->pos.z; c.size = n->size; c.level = n->level; c.morton = n->morton_code;
			c.overlap = bmf_seam_overlap(dim);
		}
		bmf_sampler_desc d = world->sampler.device;
		if (d.kind == BMF_SAMPLER_HOST_DENSITY) return false;
		d.world_size = world->sampler.world_size;
		if (d.kind == BMF_SAMPLER_TERRAIN2D_PERT)
		{
			const NoiseSamplers::NoiseSamplerProperties& np = world->noise_properties;
			d.g_scale = np.g_scale; d.height = np.height; d.octaves = np.octaves; d.amp = np.amp; d.frequency = np.frequency; d.gain = np.gain;
		}
		std::lock_guard<std::mutex> guard(dev.lock);
		if (bmf_sampler_set(dev.ctx, &d) != BMF_OK) return false;
		bmf_params p;
		std::memset(&p, 0, sizeof(p));
		p.dim = dim; // signs and border samples only: no smoothing
		if (bmf_batch_submit(dev.ctx, descs.data(), (int)count, &p, nullptr) != BMF_OK) return false;
		int64_t nt = 0;
		if (bmf_batch_stitch(dev.ctx, group, cross_only ? 1 : 0, &nt) != BMF_OK) return false;
		bmf_seam_stage_ms(dev.ctx, ms);
		tris.assign(9 * (size_t)nt, 0.0f);
		if (nt && bmf_seam_download(dev.ctx, tris.data()) != BMF_OK) return false;
		return true;
	}

	bool append_soup(const std::vector<float>& tris)
	{
		const size_t nv = tris.size() / 3, v0 = vertices.count;
		if (!nv) return true;
		if (!vertices.prepare(nv)) return false;
		vertices.count = v0 + nv;
		for (size_t v = 0; v < nv; v++)
		{
			DualVertex& dv = vertices[(int)(v0 + v)];
			std::memset((void*)&dv, 0, sizeof(dv));
			dv.p = glm::vec3(tris[3 * v], tris[3 * v + 1], tris[3 * v + 2]);
			dv.color = glm::vec3(0.85f, 1.0f, 0.85f); // DUAL_VERTEX (WorldStitcher.cpp:484-487)
		}
		return true;
	}

public:
	// GLChunk::format_data_tris(vertices) (GLChunk.cpp:257-276): positions, normals (the stitcher never sets them: zero) and
	// colours of the soup, flat
	void format()
	{
		gl_chunk.p_data.count = gl_chunk.c_data.count = gl_chunk.n_data.count = 0;
		for (size_t v = 0; v < vertices.count; v++)
		{
			gl_chunk.p_data.push_back(vertices[(int)v].p);
			gl_chunk.n_data.push_back(vertices[(int)v].n);
			gl_chunk.c_data.push_back(vertices[(int)v].color);
		}
	}
};

// ---- ChunkGenerator ---------------------------------------------------------------------------------------------
class ChunkGenerator
{
public:
	WorldStitcher stitcher; // ChunkGenerator.hpp:38
	ResourceAllocator<GLChunk> gl_allocator;
	ResourceAllocator<DensityBlock> density_allocator;
	ResourceAllocator<BinaryBlock> binary_allocator;
	ResourceAllocator<MasksBlock> masks_allocator;
	ResourceAllocator<VerticesIndicesBlock> vi_allocator;
	ResourceAllocator<DMC_CellsBlock> cell_allocator;
	ResourceAllocator<IndexesBlock> inds_allocator;
	ResourceAllocator<NoiseBlock> noise_allocator;

	void init(WorldOctree* w) { world = w; }

	// Multi-GPU: chunks are independent, so a batch is dealt to the devices in contiguous
	// Morton-ordered, equal-count ranges (one context + one host thread per entry; the same GPU may be listed twice),
	// with no collective: each device fills the chunks of its own range and the host just joins the threads.
	void set_devices(const std::vector<int>& device_ids) { devices = device_ids; }

	// ChunkGenerator.cpp:27-60 + extract_chunk :80-147, the whole batch as one device submission per GPU
	bool process_queue(SmartContainer<WorldOctreeNode*>& batch)
	{
		if (!world) return false;
		const int count = (int)batch.count;
		{
			std::unique_lock<std::mutex> lock(world->chunk_mutex);
			for (int i = 0; i < count; i++)
				if (!batch[i]->chunk) world->create_chunk(batch[i]);
		}
		std::vector<int> who;
		for (int i = 0; i < count; i++)
			if (batch[i]->generation_stage == GENERATION_STAGES_GENERATING) who.push_back(i);
		bool ok = true;
		if (!who.empty())
		{
			if (world->sampler.device.kind == BMF_SAMPLER_HOST_DENSITY) return false; // host callbacks go through DMCChunk::label_grid one chunk at a time
			const int n_dev = devices.empty() ? 1 : (int)devices.size();
			if (n_dev == 1)
				ok = run_range(BmfDevice::slot(0, devices.empty() ? 0 : devices[0]), batch, who);
			else
			{
				std::vector<int> order(who);
				std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return morton_less(batch[a]->morton_code, batch[b]->morton_code); });
				std::vector<std::vector<int>> parts(n_dev);
				for (int k = 0; k < n_dev; k++)
					parts[k].assign(order.begin() + (size_t)order.size() * k / n_dev, order.begin() + (size_t)order.size() * (k + 1) / n_dev);
				std::vector<char> res(n_dev, 1);
				std::vector<std::thread> th;
				for (int k = 0; k < n_dev; k++)
					th.emplace_back([&, k] { res[k] = parts[k].empty() ? 1 : (run_range(BmfDevice::slot(k, devices[k]), batch, parts[k]) ? 1 : 0); });
				for (std::thread& t : th) t.join();
				for (char r : res) ok = ok && r;
			}
		}
		for (int i = 0; i < count; i++)
			batch[i]->generation_stage = (batch[i]->chunk && batch[i]->chunk->vi) ? GENERATION_STAGES_NEEDS_UPLOAD : GENERATION_STAGES_DONE; // ChunkGenerator.cpp:137-143
		return ok;
	}

private:
	// one device, one submission: the chunks batch[who[k]]
	bool run_range(BmfDevice& dev, SmartContainer<WorldOctreeNode*>& batch, const std::vector<int>& who)
	{
		if (!dev.ok()) return false;
		const int iters = world->properties.process_iters, max_level = world->properties.max_level;
		const bool pb = world->properties.boundary_processing;
		const float base_overlap = world->properties.overlap;
		std::vector<bmf_chunk_desc> descs(who.size());
		for (size_t k = 0; k < who.size(); k++)
		{
			WorldOctreeNode* n = batch[who[k]];
			bmf_chunk_desc& c = descs[k];
			c.pos[0] = n->pos.x; c.pos[1] = n->pos.y; c.pos[2] = n->pos.z; c.size = n->size; c.level = n->level; c.morton = n->morton_code;
			c.overlap = (n->level == max_level && (!pb || iters == 0)) ? 0.0f : base_overlap + 0.005f * (float)iters; // ChunkGenerator.cpp:98
			// with a working seam pass the chunks are sampled at voxel-node centres instead of being overlapped
			if (world->properties.enable_stitching) c.overlap = bmf_seam_overlap(world->properties.chunk_resolution);
		}
		bmf_sampler_desc d = world->sampler.device;
		d.world_size = world->sampler.world_size;
		if (d.kind == BMF_SAMPLER_TERRAIN2D_PERT)
		{
			const NoiseSamplers::NoiseSamplerProperties& np = world->noise_properties;
			d.g_scale = np.g_scale; d.height = np.height; d.octaves = np.octaves; d.amp = np.amp; d.frequency = np.frequency; d.gain = np.gain;
		}
		std::lock_guard<std::mutex> guard(dev.lock);
		if (bmf_sampler_set(dev.ctx, &d) != BMF_OK) return false;
		bmf_params p;
		std::memset(&p, 0, sizeof(p));
		p.dim = world->properties.chunk_resolution;
		p.iters = iters;
		p.process_boundary = pb ? 1 : 0;
		p.smooth_normals = 0; // SMOOTH_NORMALS 0 (DefaultOptions.h:8)
		if (bmf_batch_submit(dev.ctx, descs.data(), (int)descs.size(), &p, nullptr) != BMF_OK) return false;
		int64_t nc = 0, nv = 0, ni = 0;
		bmf_batch_totals(dev.ctx, &nc, &nv, &ni);
		std::vector<float> pos(3 * (size_t)nv + 3), col(3 * (size_t)nv + 3), nrm(3 * (size_t)nv + 3);
		std::vector<uint8_t> bnd((size_t)nv + 1), val((size_t)nv + 1);
		std::vector<uint32_t> idx((size_t)ni + 1);
		if (bmf_batch_download(dev.ctx, pos.data(), nrm.data(), col.data(), bnd.data(), val.data(), idx.data()) != BMF_OK) return false;
		std::vector<bmf_chunk_info> infos(descs.size());
		bmf_batch_chunk_infos(dev.ctx, infos.data());
		for (size_t k = 0; k < descs.size(); k++)
		{
			WorldOctreeNode* n = batch[who[k]];
			DMCChunk* c = n->chunk;
			const bmf_chunk_info& inf = infos[k];
			c->contains_mesh = inf.contains_mesh != 0;
			c->overlap_pos = glm::vec3(inf.overlap_pos[0], inf.overlap_pos[1], inf.overlap_pos[2]);
			c->scale = inf.scale;
			c->bound_size = c->size * (1.0f + descs[k].overlap * 2.0f) * 0.5f;
			c->bound_start = c->overlap_pos + c->bound_size;
			if (!c->contains_mesh) continue;
			if (!c->vi) { c->vi = vi_allocator.new_element(); c->vi->init(); }
			const size_t v0 = (size_t)inf.vert_offset, i0 = (size_t)inf.ind_offset;
			const bool processed = iters > 0 && inf.n_verts > 0 && inf.n_inds > 0;
			bmf_detail::fill_dual_vertices(c->vi->vertices, &pos[3 * v0], &nrm[3 * v0], &col[3 * v0], &bnd[v0], &val[v0], (size_t)inf.n_verts, true, processed);
			c->vi->mesh_indexes.count = 0;
			c->vi->mesh_indexes.push_back(&idx[i0], (size_t)inf.n_inds);
			// WorldOctreeNode::format -> GLChunk::format_data(vertices, indexes, false, false) (WorldOctreeNode.cpp:72-88)
			if (!n->gl_chunk) n->gl_chunk = gl_allocator.new_element();
			GLChunk* g = n->gl_chunk;
			g->p_data.count = g->n_data.count = g->c_data.count = 0;
			g->p_data.push_back((const glm::vec3*)&pos[3 * v0], (size_t)inf.n_verts);
			g->n_data.push_back((const glm::vec3*)&nrm[3 * v0], (size_t)inf.n_verts);
			g->c_data.push_back((const glm::vec3*)&col[3 * v0], (size_t)inf.n_verts);
		}
		return true;
	}

	WorldOctree* world = nullptr;
	std::vector<int> devices;
};

// ---- WorldWatcher -----------------------------------------------------------------------------------------------
// WorldWatcher::update (WorldWatcher.cpp:34-134) as one SYNCHRONOUS tick: check_leaves (:155-173) with handle_split_check /
// handle_group_check (:175-212), process_batch (:214-238) with split_node / group_node_1 (:354-392), the generator call
// (:69-73), optional restitch (:74-103) and post_process_batch (:240-352).  The watcher thread, the render-thread upload
// handshake and the draw / stitch flags are the renderer's business and are not mirrored.
class WorldWatcher
{
public:
	WorldOctree* world = nullptr;
	ChunkGenerator generator;
	glm::vec3 focus_pos;
	std::vector<WorldOctreeNode*> renderables; // the reference's linked list, in link order
	size_t last_generated = 0;

	// from_root = true: like WorldWatcher::init (WorldWatcher.cpp:21-31) only the root is drawable and the ticks build the
	// world (the sequence of batches and the final list then equal the reference's, tests/golden/watcher_golden.json);
	// false: start from the leaves of a previous WorldOctree::split_leaves
	void init(WorldOctree* w, const glm::vec3& focus, bool from_root = false)
	{
		world = w;
		focus_pos = focus;
		generator.init(w);
		if (from_root)
		{
			renderables.clear();
			renderables.push_back(&w->octree);
		}
		else
			renderables = w->leaves;
	}

	// returns false if the generator failed; last_generated = chunks handed to process_queue in this tick
	bool update(int max_gen = 400)
	{
		std::vector<WorldOctreeNode*> dirty;
		int counter = 0;
		for (WorldOctreeNode* n : renderables) // check_leaves
		{
			if (counter >= max_gen) break;
			if (world->node_needs_split(focus_pos, n))
			{
				if (n->world_leaf_flag && !n->flags) { n->flags = 1; dirty.push_back(n); }
				counter += 8;
			}
			else if (n->world_leaf_flag && n->parent && world->node_needs_group(focus_pos, n->parent))
			{
				WorldOctreeNode* par = n->parent;
				bool can = !par->flags;
				for (int i = 0; i < 8 && can; i++) can = par->children[i] && par->children[i]->world_leaf_flag && !par->children[i]->flags;
				if (can) { par->flags = 2; dirty.push_back(par); }
			}
		}
		SmartContainer<WorldOctreeNode*> generate_batch;
		std::vector<WorldOctreeNode*> gone;
		for (WorldOctreeNode* n : dirty) // process_batch
		{
			if (n->flags == 1)
			{
				world->split_node(n);
				for (int i = 0; i < 8; i++)
				{
					n->children[i]->generation_stage = GENERATION_STAGES_GENERATING;
					generate_batch.push_back(n->children[i]);
					renderables.push_back(n->children[i]);
				}
				gone.push_back(n);
			}
			else
			{
				for (int i = 0; i < 8; i++) gone.push_back(n->children[i]);
				world->group_node(n);
				n->generation_stage = GENERATION_STAGES_GENERATING;
				generate_batch.push_back(n);
				renderables.push_back(n);
			}
			n->flags = 0;
		}
		last_generated = generate_batch.count;
		bool ok = true;
		if (generate_batch.count) ok = generator.process_queue(generate_batch);
		if (ok && generate_batch.count && world->properties.enable_stitching) ok = generator.stitcher.stitch_batch_nodes(world, renderables_alive(gone).data(), renderables_alive(gone).size());
		if (!gone.empty()) renderables = renderables_alive(gone); // post_process_batch: unlink_renderable
		return ok;
	}

private:
	std::vector<WorldOctreeNode*> renderables_alive(const std::vector<WorldOctreeNode*>& gone) const
	{
		std::vector<WorldOctreeNode*> out;
		out.reserve(renderables.size());
		for (WorldOctreeNode* r : renderables)
			if (std::find(gone.begin(), gone.end(), r) == gone.end()) out.push_back(r);
		return out;
	}
};

// ---- MeshProcessor -------------------------------------------------------------------------------------------------
namespace Processing
{
template <int N>
class MeshProcessor
{
public:
	bool simple_quality;
	MeshProcessor(bool simple_quality_, bool smooth_normals_) : simple_quality(simple_quality_), smooth_normals(smooth_normals_) {}

	// copies the vertices and the index buffer (MeshProcessor.cpp:25-55); CSR adjacency is built on the device
	bool init(SmartContainer<DualVertex>& v, SmartContainer<uint32_t>& inds, Sampler& /*sampler*/)
	{
		if (v.count == 0 || inds.count < (size_t)N) return true;
		vertices.count = 0;
		vertices.push_back(v);
		indices.assign(inds.elements, inds.elements + inds.count / N * N);
		destroyed.clear();
		uint32_t a = 0;
		for (size_t i = 0; i < vertices.count; i++)
		{
			DualVertex& dv = vertices.elements[i];
			dv.adj_offset = a; dv.adj_next = dv.init_valence; dv.valence = dv.init_valence;
			a += dv.init_valence;
		}
		pending_iters = 0;
		return a != 0;
	}
	// recorded; executed by the following optimize_primal_grid (the reference driver's sequence,
	// ChunkGenerator.cpp:119-120) or, without it, by flush
	void optimize_dual_grid(int iterations, bool process_boundary = true)
	{
		pending_iters = iterations;
		pending_pb = process_boundary;
	}
	void optimize_primal_grid(bool /*qef: ignored by the reference too, MeshProcessor.cpp:239*/, bool /*set_colors*/, bool process_boundary = true)
	{
		(void)process_boundary; // the driver passes the same flag to both calls
		if (pending_iters > 0) run(pending_iters, pending_pb, 1);
		pending_iters = 0;
	}
	// MeshProcessor.cpp:308-396: quads only (N != 4 returns at once, :311-312); serial order kept by the device kernel (csrc/post.cuh)
	void collapse_bad_quads()
	{
		if (N != 4 || vertices.count == 0 || indices.size() < 4) return;
		BmfDevice& dev = BmfDevice::get();
		if (!dev.ok()) return;
		const size_t n = vertices.count, nq = indices.size() / 4;
		std::vector<float> pos(3 * n);
		std::vector<uint8_t> adj_next(n);
		for (size_t i = 0; i < n; i++)
		{
			const DualVertex& v = vertices.elements[i];
			pos[3 * i] = v.p.x; pos[3 * i + 1] = v.p.y; pos[3 * i + 2] = v.p.z;
		}
		destroyed.assign(nq, 0);
		int64_t bad = 0;
		{
			std::lock_guard<std::mutex> guard(dev.lock);
			if (bmf_mesh_collapse_bad_quads(dev.ctx, pos.data(), (int)n, indices.data(), (int64_t)nq, destroyed.data(), adj_next.data(), nullptr, nullptr, &bad) != BMF_OK)
			{
				destroyed.clear();
				return;
			}
		}
		for (size_t i = 0; i < n; i++)
		{
			DualVertex& v = vertices.elements[i];
			v.p = glm::vec3(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
			v.adj_next = adj_next[i];
		}
		bad_quads = (uint32_t)bad; // the reference prints "detected N bad quads..." (:395)
	}
	uint32_t bad_quads = 0;
	void flush(SmartContainer<DualVertex>& v_out, SmartContainer<uint32_t>& inds)
	{
		if (pending_iters > 0) run(pending_iters, pending_pb, 0);
		pending_iters = 0;
		v_out.push_back(vertices);
		if (destroyed.empty()) inds.push_back(indices.data(), indices.size());
		else
			for (size_t t = 0; t < indices.size() / N; t++) // `if (t.destroyed) continue;` (MeshProcessor.cpp:64-65)
				if (!destroyed[t]) inds.push_back(&indices[t * N], N);
	}
	void flush_to_tris(SmartContainer<DualVertex>& v_out, SmartContainer<uint32_t>& inds) // quad -> (0,1,2),(2,3,0), MeshProcessor.cpp:73-91
	{
		if (pending_iters > 0) run(pending_iters, pending_pb, 0);
		pending_iters = 0;
		v_out.push_back(vertices);
		for (size_t t = 0; t + 3 < indices.size() + 1 && N == 4; t += 4)
		{
			if (!destroyed.empty() && destroyed[t / 4]) continue;
			const uint32_t* q = &indices[t];
			const uint32_t tri[6] = { q[0], q[1], q[2], q[2], q[3], q[0] };
			inds.push_back(tri, 6);
		}
	}

private:
	bool run(int iters, bool pb, int final_primal)
	{
		BmfDevice& dev = BmfDevice::get();
		if (!dev.ok() || vertices.count == 0) return false;
		const size_t n = vertices.count;
		std::vector<float> pos(3 * n), col(3 * n), nrm(3 * n);
		std::vector<uint8_t> bnd(n);
		for (size_t i = 0; i < n; i++)
		{
			const DualVertex& v = vertices.elements[i];
			pos[3 * i] = v.p.x; pos[3 * i + 1] = v.p.y; pos[3 * i + 2] = v.p.z;
			col[3 * i] = v.color.x; col[3 * i + 1] = v.color.y; col[3 * i + 2] = v.color.z;
			nrm[3 * i] = v.n.x; nrm[3 * i + 1] = v.n.y; nrm[3 * i + 2] = v.n.z;
			bnd[i] = v.boundary ? 1 : 0;
		}
		{
			std::lock_guard<std::mutex> guard(dev.lock);
			if (bmf_mesh_process_steps(dev.ctx, pos.data(), col.data(), nrm.data(), bnd.data(), nullptr, (int)n, indices.data(), (int)indices.size(), N, iters, pb ? 1 : 0,
			                           smooth_normals ? 1 : 0, final_primal) != BMF_OK)
				return false;
		}
		for (size_t i = 0; i < n; i++)
		{
			DualVertex& v = vertices.elements[i];
			v.p = glm::vec3(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
			v.color = glm::vec3(col[3 * i], col[3 * i + 1], col[3 * i + 2]);
			// always: with smooth normals off the reference's set_colors step still turns the zero normal of every processed vertex
			// into NaN (MeshProcessor.cpp:229-232, 296-303) -- the batch path reproduces that, so the mirror must as well
			v.n = glm::vec3(nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2]);
			v.s = 0.0f;
		}
		return true;
	}

	SmartContainer<DualVertex> vertices;
	std::vector<uint32_t> indices;
	std::vector<uint8_t> destroyed; // Primitive::destroyed, only after collapse_bad_quads
	bool smooth_normals;
	int pending_iters = 0;
	bool pending_pb = true;
};
} // namespace Processing

// ---- ColorMapper (ColorMapper.hpp:7-24, ColorMapper.cpp:15-60) ----------------------------------------------------
class ColorMapper
{
public:
	FastNoiseSIMD* noise_context = nullptr; // the noise library is replaced by the device kernel
	ColorMapper() {}
	~ColorMapper() {}
	void generate_colors(SmartContainer<DualVertex>& verts)
	{
		if (!verts.count) return;
		BmfDevice& dev = BmfDevice::get();
		if (!dev.ok()) return;
		const size_t n = verts.count;
		std::vector<float> pos(3 * n), col(3 * n);
		for (size_t i = 0; i < n; i++)
		{
			const DualVertex& v = verts.elements[i];
			pos[3 * i] = v.p.x; pos[3 * i + 1] = v.p.y; pos[3 * i + 2] = v.p.z;
		}
		{
			std::lock_guard<std::mutex> guard(dev.lock);
			if (bmf_color_map(dev.ctx, pos.data(), (int64_t)n, col.data()) != BMF_OK) return;
		}
		for (size_t i = 0; i < n; i++) verts.elements[i].color = glm::vec3(col[3 * i], col[3 * i + 1], col[3 * i + 2]);
	}
};

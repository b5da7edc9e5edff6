// ref_export.cpp -- extern "C" surface over the UNMODIFIED reference translation units
// (compiled from /root/reference where they lie; see oracle/Makefile).  TEST INFRASTRUCTURE ONLY:
// it is the parity oracle for stages 2-5 and the CPU baseline of bench.py; nothing in the
// product path links it.  It contains no algorithm of its own: every function drives reference
// classes exactly like ChunkGenerator::extract_chunk (ChunkGenerator.cpp:80-147) does and copies
// their public fields out.
#include "PCH.h"
// the watcher's tick functions (check_leaves / process_batch / post_process_batch) are private members; this driver only
// CALLS them, in the order WorldWatcher::update does (g++ does not reorder members across access specifiers)
#include <atomic>
#include <condition_variable>
#include <list>
#include <map>
#include <mutex>
#include <sstream>
#include <thread>
#include <unordered_map>
#include "ThreadDebug.hpp"
#include "SmartContainer.hpp"
#include "ChunkGenerator.hpp"
#include "WorldOctreeNode.hpp"
#include "HashMap.hpp"
#include "sparsepp/spp.h"
#define private public
#include "WorldWatcher.hpp"
#undef private
#include "WorldOctree.hpp"
#include "ImplicitSampler.hpp"
#include "NoiseSampler.hpp"
#include "MeshProcessor.hpp"
#include "ColorMapper.hpp"
#include "DefaultOptions.h"
#include <omp.h>
#include <chrono>
#include <cstdio>
#include <iostream>
#include <sstream>
#include <vector>

float qef_solve_from_points_3d(const float* positions, const float* normals, const int count, float* solved_position);

namespace
{
const float host_value_stub(const float, const glm::vec3&) { return 0.0f; }

// sampler kinds understood by this driver (mirrors include/bmf_b200.h bmf_sampler_kind)
enum { K_SPHERE = 0, K_TORUS_Z = 1, K_CUBOID = 2, K_PLANE_Y = 3, K_TERRAIN2D = 10, K_TERRAIN2D_PERT = 11, K_TERRAIN3D = 12, K_TERRAIN3D_PERT = 13, K_HOST_DENSITY = 100 };

struct NullBuf : std::streambuf { int overflow(int c) override { return c; } };

struct CoutSilencer
{
	NullBuf nb;
	std::streambuf* old;
	CoutSilencer() { old = std::cout.rdbuf(&nb); }
	~CoutSilencer() { std::cout.rdbuf(old); }
};

bool make_sampler(Sampler* s, int kind, float world_size, const float* host_density, size_t host_count)
{
	switch (kind)
	{
	case K_SPHERE: *s = ImplicitFunctions::create_sampler(ImplicitFunctions::sphere); break;
	case K_TORUS_Z: *s = ImplicitFunctions::create_sampler(ImplicitFunctions::torus_z); break;
	case K_CUBOID: *s = ImplicitFunctions::create_sampler(ImplicitFunctions::cuboid); break;
	case K_PLANE_Y: *s = ImplicitFunctions::create_sampler(ImplicitFunctions::plane_y); break;
	case K_TERRAIN2D: NoiseSamplers::create_sampler_terrain_2d(s); break;
	case K_TERRAIN2D_PERT: NoiseSamplers::create_sampler_terrain_pert_2d(s); break;
	case K_TERRAIN3D: NoiseSamplers::create_sampler_terrain_3d(s); break;
	case K_TERRAIN3D_PERT: NoiseSamplers::create_sampler_terrain_pert_3d(s); break;
	case K_HOST_DENSITY:
		s->value = host_value_stub;
		s->block = [host_density, host_count](const float, const glm::vec3&, const glm::ivec3& size, const float, void** out, FastNoiseVectorSet*, float*, int, int, SamplerProperties*) {
			size_t n = (size_t)size.x * size.y * size.z;
			if (n > host_count) n = host_count;
			memcpy(*out, host_density, n * sizeof(float));
		};
		break;
	default: return false;
	}
	s->world_size = world_size;
	return true;
}

// props = {g_scale, height, octaves, amp, frequency, gain}; null -> WorldOctree.cpp:47-54 defaults
void fill_noise_props(NoiseSamplers::NoiseSamplerProperties* p, const float* props)
{
	p->noise_type = FastNoiseSIMD::NoiseType::ValueFractal;
	p->fractal_type = FastNoiseSIMD::FractalType::FBM;
	p->level = 0;
	p->g_scale = props ? props[0] : 0.25f;
	p->height = props ? props[1] : 75.0f;
	p->octaves = props ? (int)props[2] : 13;
	p->amp = props ? props[3] : 0.87f;
	p->frequency = props ? props[4] : 0.585f;
	p->gain = props ? props[5] : 0.488f;
}

struct Pools
{
	ResourceAllocator<DensityBlock> density;
	ResourceAllocator<BinaryBlock> binary;
	ResourceAllocator<MasksBlock> masks;
	ResourceAllocator<VerticesIndicesBlock> vi;
	ResourceAllocator<DMC_CellsBlock> cells;
	ResourceAllocator<IndexesBlock> inds;
	ResourceAllocator<NoiseBlock> noise;
};

struct RefChunk
{
	Pools pools;
	Sampler sampler;
	DMCChunk* chunk;
	std::vector<uint8_t> masks; // uint8[d][d][d] image of the MasksBlock
	std::vector<float> host_density;
};

struct RefBatch
{
	WorldOctree* world;
	SmartContainer<WorldOctreeNode*> batch;
	double last_ms;
	int kind;
};
}

extern "C" {

int ref_sizeof_dualvertex() { return (int)sizeof(DualVertex); }
int ref_sizeof_cell() { return (int)sizeof(DMC_Cell); }

// Offsets of the DualVertex fields, for the numpy structured dtype in tests.
void ref_dualvertex_offsets(int* out)
{
	DualVertex* v = 0;
	out[0] = (int)(size_t)&v->boundary; out[1] = (int)(size_t)&v->index; out[2] = (int)(size_t)&v->valence;
	out[3] = (int)(size_t)&v->init_valence; out[4] = (int)(size_t)&v->adj_next; out[5] = (int)(size_t)&v->adj_offset;
	out[6] = (int)(size_t)&v->s; out[7] = (int)(size_t)&v->p; out[8] = (int)(size_t)&v->n; out[9] = (int)(size_t)&v->color;
}

// One chunk through label_grid -> label_edges -> polygonize (SURVEY call stack C).
void* ref_chunk_create(int kind, float world_size, const float* noise_props, const float* pos, float size, int level, int dim,
                       float overlap, const float* host_density)
{
	RefChunk* rc = new RefChunk();
	size_t n = (size_t)dim * dim * dim;
	if (kind == K_HOST_DENSITY)
		rc->host_density.assign(host_density, host_density + n);
	if (!make_sampler(&rc->sampler, kind, world_size, rc->host_density.data(), n))
	{
		delete rc;
		return nullptr;
	}
	NoiseSamplers::NoiseSamplerProperties props;
	fill_noise_props(&props, noise_props);

	// NoiseBlock is sized dim*dim by label_grid (DMCChunk.cpp:107) but the 3-D samplers write dim^3
	// (NoiseSampler.cpp:201,234): pre-size the pooled block so the reference does not overrun it.
	// 2-D samplers keep the reference's own dim*dim sizing (FillNoiseSet walks vectorset.size points).
	if (kind == K_TERRAIN3D || kind == K_TERRAIN3D_PERT)
	{
		NoiseBlock* nb = rc->pools.noise.new_element();
		nb->init((uint32_t)n);
		rc->pools.noise.free_element(nb);
	}

	rc->chunk = new DMCChunk(glm::vec3(pos[0], pos[1], pos[2]), size, level, rc->sampler, 1);
	rc->chunk->dim = dim;
	rc->chunk->label_grid(&rc->pools.binary, &rc->pools.density, &rc->pools.noise, overlap, props);
	rc->chunk->label_edges(&rc->pools.vi, &rc->pools.cells, &rc->pools.inds, &rc->pools.density, &rc->pools.masks);
	if (rc->chunk->contains_mesh)
	{
		// label_edges returned its MasksBlock to the pool; take it back to copy the byte image out
		MasksBlock* mb = rc->pools.masks.new_element();
		rc->masks.assign((uint8_t*)mb->data, (uint8_t*)mb->data + n);
		rc->pools.masks.free_element(mb);
	}
	rc->chunk->polygonize();
	return rc;
}

// MeshProcessor<3> exactly as ChunkGenerator.cpp:110-124 drives it.
void ref_chunk_process(void* h, int iters, int process_boundary, int smooth_normals)
{
	RefChunk* rc = (RefChunk*)h;
	DMCChunk* c = rc->chunk;
	if (iters > 0 && c->contains_mesh && c->vi->vertices.count && c->vi->mesh_indexes.count)
	{
		auto& v_out = c->vi->vertices;
		auto& i_out = c->vi->mesh_indexes;
		Processing::MeshProcessor<3> mp(true, smooth_normals != 0);
		mp.init(c->vi->vertices, c->vi->mesh_indexes, rc->sampler);
		mp.optimize_dual_grid(iters, process_boundary != 0);
		mp.optimize_primal_grid(false, false, process_boundary != 0);
		v_out.count = 0;
		i_out.count = 0;
		mp.flush(v_out, i_out);
	}
}

void ref_chunk_info(void* h, int* contains_mesh, int* n_cells, int* n_verts, int* n_inds, float* geom /* overlap_pos[3], scale */)
{
	RefChunk* rc = (RefChunk*)h;
	DMCChunk* c = rc->chunk;
	*contains_mesh = c->contains_mesh ? 1 : 0;
	*n_cells = (c->contains_mesh && c->cell_block) ? (int)c->cell_block->cells.count : 0;
	*n_verts = (c->contains_mesh && c->vi) ? (int)c->vi->vertices.count : 0;
	*n_inds = (c->contains_mesh && c->vi) ? (int)c->vi->mesh_indexes.count : 0;
	if (geom)
	{
		geom[0] = c->overlap_pos.x; geom[1] = c->overlap_pos.y; geom[2] = c->overlap_pos.z; geom[3] = c->scale;
	}
}

// Any output pointer may be null.  cell_masks/cell_index: per active cell, in cell order.
void ref_chunk_copy(void* h, float* density, uint32_t* bits, uint8_t* masks, uint32_t* dense_inds, uint8_t* cell_masks,
                    uint32_t* cell_grid_index, void* vertices, uint32_t* indices)
{
	RefChunk* rc = (RefChunk*)h;
	DMCChunk* c = rc->chunk;
	size_t d = c->dim, n = d * d * d;
	if (density) memcpy(density, c->density_block->data, n * sizeof(float));
	if (bits) memcpy(bits, c->binary_block->data, (n / 32) * sizeof(uint32_t));
	if (!c->contains_mesh) return;
	if (masks) memcpy(masks, rc->masks.data(), n);
	if (dense_inds) memcpy(dense_inds, c->indexes_block->inds, n * sizeof(uint32_t));
	size_t nc = c->cell_block->cells.count;
	for (size_t i = 0; i < nc; i++)
	{
		if (cell_masks) cell_masks[i] = (uint8_t)c->cell_block->cells[(int)i].mask;
		if (cell_grid_index) cell_grid_index[i] = c->cell_block->cells[(int)i].edges[0].grid_v0;
	}
	if (vertices) memcpy(vertices, c->vi->vertices.elements, c->vi->vertices.count * sizeof(DualVertex));
	if (indices) memcpy(indices, c->vi->mesh_indexes.elements, c->vi->mesh_indexes.count * sizeof(uint32_t));
}

void ref_chunk_destroy(void* h)
{
	RefChunk* rc = (RefChunk*)h;
	// the chunk's blocks are owned by the pools; DMCChunk's destructor touches a never-constructed
	// MemoryPool when contains_mesh is set, so the chunk object is intentionally leaked (tests only)
	delete rc;
}

// Stand-alone MeshProcessor<N> over caller arrays (N = 3 or 4); vertices are DualVertex records.
void ref_mesh_process(void* vertices, int n_verts, uint32_t* indices, int n_inds, int prim_n, int iters, int process_boundary,
                      int smooth_normals)
{
	SmartContainer<DualVertex> v;
	SmartContainer<uint32_t> idx;
	v.push_back((DualVertex*)vertices, (size_t)n_verts);
	idx.push_back(indices, (size_t)n_inds);
	Sampler s = ImplicitFunctions::create_sampler(ImplicitFunctions::sphere);
	s.world_size = 256;
	if (prim_n == 3)
	{
		Processing::MeshProcessor<3> mp(true, smooth_normals != 0);
		mp.init(v, idx, s);
		mp.optimize_dual_grid(iters, process_boundary != 0);
		mp.optimize_primal_grid(false, false, process_boundary != 0);
		v.count = 0; idx.count = 0;
		mp.flush(v, idx);
	}
	else
	{
		Processing::MeshProcessor<4> mp(true, smooth_normals != 0);
		mp.init(v, idx, s);
		mp.optimize_dual_grid(iters, process_boundary != 0);
		mp.optimize_primal_grid(false, false, process_boundary != 0);
		v.count = 0; idx.count = 0;
		mp.flush(v, idx);
	}
	memcpy(vertices, v.elements, (size_t)n_verts * sizeof(DualVertex));
	memcpy(indices, idx.elements, (size_t)n_inds * sizeof(uint32_t));
}

// MeshProcessor<4>::init + collapse_bad_quads + flush (MeshProcessor.cpp:25-71, 308-396) on caller arrays; returns the number of
// surviving quads * 4 written back into indices.  (collapse_bad_quads prints its count to std::cout.)
int ref_collapse_bad_quads(void* vertices, int n_verts, uint32_t* indices, int n_inds)
{
	SmartContainer<DualVertex> v;
	SmartContainer<uint32_t> idx;
	v.push_back((DualVertex*)vertices, (size_t)n_verts);
	idx.push_back(indices, (size_t)n_inds);
	Sampler s = ImplicitFunctions::create_sampler(ImplicitFunctions::sphere);
	s.world_size = 256;
	Processing::MeshProcessor<4> mp(true, false);
	mp.init(v, idx, s);
	mp.collapse_bad_quads();
	v.count = 0; idx.count = 0;
	mp.flush(v, idx);
	memcpy(vertices, v.elements, (size_t)n_verts * sizeof(DualVertex));
	memcpy(indices, idx.elements, idx.count * sizeof(uint32_t));
	return (int)idx.count;
}

// ColorMapper::generate_colors (ColorMapper.cpp:15-25) on caller DualVertex records
void ref_color_map(void* vertices, int n_verts)
{
	SmartContainer<DualVertex> v;
	v.push_back((DualVertex*)vertices, (size_t)n_verts);
	ColorMapper cm;
	cm.generate_colors(v);
	memcpy(vertices, v.elements, (size_t)n_verts * sizeof(DualVertex));
}

// GLChunk::format_data(vertices, indexes, unwind_verts = true, smooth_normals) (GLChunk.cpp:278-335): the flat-quad SoA
void ref_format_unwind(void* vertices, int n_verts, uint32_t* indices, int n_inds, int smooth_normals, float* p_out, float* n_out, float* c_out)
{
	SmartContainer<DualVertex> v;
	SmartContainer<uint32_t> idx;
	v.push_back((DualVertex*)vertices, (size_t)n_verts);
	idx.push_back(indices, (size_t)n_inds);
	GLChunk g;
	g.format_data(v, idx, true, smooth_normals != 0);
	memcpy(p_out, g.p_data.elements, g.p_data.count * sizeof(glm::vec3));
	memcpy(n_out, g.n_data.elements, g.n_data.count * sizeof(glm::vec3));
	memcpy(c_out, g.c_data.elements, g.c_data.count * sizeof(glm::vec3));
}

float ref_qef_solve(const float* positions, const float* normals, int count, float* solved)
{
	return qef_solve_from_points_3d(positions, normals, count, solved);
}

float ref_implicit_value(int kind, float world_size, const float* p)
{
	glm::vec3 v(p[0], p[1], p[2]);
	switch (kind)
	{
	case K_SPHERE: return ImplicitFunctions::sphere(world_size, v);
	case K_TORUS_Z: return ImplicitFunctions::torus_z(world_size, v);
	case K_CUBOID: return ImplicitFunctions::cuboid(world_size, v);
	default: return ImplicitFunctions::plane_y(world_size, v);
	}
}

void ref_implicit_gradient(int kind, float world_size, const float* p, float h, float* out)
{
	SamplerValueFunction f = kind == K_SPHERE ? ImplicitFunctions::sphere : kind == K_TORUS_Z ? ImplicitFunctions::torus_z
	                       : kind == K_CUBOID ? ImplicitFunctions::cuboid : ImplicitFunctions::plane_y;
	glm::vec3 g = ImplicitFunctions::implicit_gradient(f, world_size, glm::vec3(p[0], p[1], p[2]), h);
	out[0] = g.x; out[1] = g.y; out[2] = g.z;
}

// ---- world / batch driver (SURVEY call stack B) -------------------------------------------------

// props6 -> noise properties; wp = {max_level, min_level, process_iters, chunk_resolution, boundary_processing}
// fp = {split_multiplier, size_modifier, overlap, focus.x, focus.y, focus.z}
void* ref_world_create(int kind, float world_size, const float* noise_props, const int* wp, const float* fp)
{
	CoutSilencer quiet;
	RefBatch* rb = new RefBatch();
	rb->kind = kind;
	rb->world = new WorldOctree();
	make_sampler(&rb->world->sampler, kind, world_size, nullptr, 0);
	fill_noise_props(&rb->world->noise_properties, noise_props);
	WorldProperties& p = rb->world->properties;
	p.max_level = wp[0]; p.min_level = wp[1]; p.process_iters = wp[2]; p.chunk_resolution = wp[3]; p.boundary_processing = wp[4] != 0;
	p.split_multiplier = fp[0]; p.group_multiplier = fp[0] * 2.0f; p.size_modifier = fp[1]; p.overlap = fp[2];
	rb->world->focus_point = glm::vec3(fp[3], fp[4], fp[5]);
	rb->world->init((uint32_t)world_size);
	rb->world->watcher.generator.init(rb->world); // WorldWatcher::init is NOT called: no thread is spawned
	rb->last_ms = 0;
	return rb;
}

// Static LOD build: WorldOctree::split_leaves (WorldOctree.cpp:129-173); leaves become the batch.
int ref_world_split_leaves(void* h)
{
	CoutSilencer quiet;
	RefBatch* rb = (RefBatch*)h;
	rb->world->split_leaves();
	rb->batch.count = 0;
	for (auto n : rb->world->leaves)
	{
		n->generation_stage = GENERATION_STAGES_GENERATING;
		rb->batch.push_back(n);
	}
	return (int)rb->batch.count;
}

// The watcher's tick without its thread: WorldWatcher::init minus std::thread (:21-31), then per tick the body of
// WorldWatcher::update (:56-112): check_leaves(dirty, 400) -> process_batch -> generator.process_queue -> post_process_batch.
// focus: [ticks][3]; codes_out: up to cap Morton codes of the renderables list (link order) after the last tick;
// gen_counts: chunks handed to the generator per tick.  Returns the number of renderables.
int ref_watcher_run(void* h, const float* focus, int ticks, uint64_t* codes_out, int cap, int* gen_counts)
{
	CoutSilencer quiet;
	RefBatch* rb = (RefBatch*)h;
	WorldWatcher& w = rb->world->watcher;
	w.world = rb->world;
	w.renderables_head = &rb->world->octree;
	w.renderables_tail = &rb->world->octree;
	w.renderables_count = 1;
	SmartContainer<WorldOctreeNode*> dirty, generate, stitch;
	for (int t = 0; t < ticks; t++)
	{
		w.focus_pos = glm::vec3(focus[3 * t], focus[3 * t + 1], focus[3 * t + 2]);
		dirty.count = 0; generate.count = 0; stitch.count = 0;
		w.check_leaves(dirty, 400);
		w.process_batch(dirty, generate, stitch);
		if (generate.count) w.generator.process_queue(generate);
		w.post_process_batch(dirty);
		if (gen_counts) gen_counts[t] = (int)generate.count;
	}
	int n = 0;
	for (WorldOctreeNode* r = w.renderables_head; r; r = r->renderable_next)
	{
		if (n < cap) codes_out[n] = r->morton_code.code;
		n++;
	}
	return n;
}

// Explicit chunk list (config 3): nodes made by hand, chunks by WorldOctree::create_chunk via process_queue.
int ref_world_add_chunks(void* h, const float* pos_size /* n x 4 */, const int* levels, int n)
{
	RefBatch* rb = (RefBatch*)h;
	for (int i = 0; i < n; i++)
	{
		const float* ps = pos_size + 4 * i;
		WorldOctreeNode* node = rb->world->node_pool.newElement(0, (WorldOctreeNode*)nullptr, ps[3], glm::vec3(ps[0], ps[1], ps[2]), (uint8_t)levels[i]);
		node->morton_code = (uint64_t)(i + 1);
		node->generation_stage = GENERATION_STAGES_GENERATING;
		rb->batch.push_back(node);
	}
	return (int)rb->batch.count;
}

int ref_world_count(void* h) { return (int)((RefBatch*)h)->batch.count; }

void ref_world_leaf(void* h, int i, float* pos_size, int* level, uint64_t* morton)
{
	RefBatch* rb = (RefBatch*)h;
	WorldOctreeNode* n = rb->batch[i];
	pos_size[0] = n->pos.x; pos_size[1] = n->pos.y; pos_size[2] = n->pos.z; pos_size[3] = n->size;
	*level = n->level;
	*morton = n->morton_code.code;
}

// ChunkGenerator::process_queue over the batch; returns wall milliseconds of that call alone.
// Chunks that already ran are reset to GENERATING and their vi returned to the pool first, so the
// call can be repeated for timing (pools warm, like the reference's steady state).
double ref_world_process(void* h, int threads)
{
	RefBatch* rb = (RefBatch*)h;
	if (threads > 8) threads = 8; // Sampler::noise_samplers[8] is indexed by omp_get_thread_num() (Sampler.hpp:30, DMCChunk.cpp:109)
	omp_set_num_threads(threads);
	ChunkGenerator& gen = rb->world->watcher.generator;
	size_t dim = rb->world->properties.chunk_resolution;
	if (rb->kind == K_TERRAIN3D || rb->kind == K_TERRAIN3D_PERT)
	{
		// pre-size one NoiseBlock per thread for the 3-D samplers (see ref_chunk_create)
		std::vector<NoiseBlock*> nbs;
		for (int t = 0; t < threads; t++) { NoiseBlock* nb = gen.noise_allocator.new_element(); nb->init((uint32_t)(dim * dim * dim)); nbs.push_back(nb); }
		for (auto nb : nbs) gen.noise_allocator.free_element(nb);
	}
	for (int i = 0; i < (int)rb->batch.count; i++)
	{
		WorldOctreeNode* n = rb->batch[i];
		if (n->chunk && n->chunk->vi)
		{
			gen.vi_allocator.free_element(n->chunk->vi);
			n->chunk->vi = 0;
		}
		if (n->chunk) n->chunk->contains_mesh = false;
		n->generation_stage = GENERATION_STAGES_GENERATING;
	}
	auto t0 = std::chrono::steady_clock::now();
	gen.process_queue(rb->batch);
	auto t1 = std::chrono::steady_clock::now();
	rb->last_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
	return rb->last_ms;
}

void ref_world_chunk_info(void* h, int i, int* contains_mesh, int* n_verts, int* n_inds)
{
	RefBatch* rb = (RefBatch*)h;
	DMCChunk* c = rb->batch[i]->chunk;
	bool m = c && c->contains_mesh && c->vi;
	*contains_mesh = (c && c->contains_mesh) ? 1 : 0;
	*n_verts = m ? (int)c->vi->vertices.count : 0;
	*n_inds = m ? (int)c->vi->mesh_indexes.count : 0;
}

// vertices as DualVertex records; p/c = the SoA the renderer receives (GLChunk::format_data, GLChunk.cpp:278-296)
void ref_world_chunk_copy(void* h, int i, void* vertices, uint32_t* indices, float* p_data, float* c_data)
{
	RefBatch* rb = (RefBatch*)h;
	WorldOctreeNode* n = rb->batch[i];
	DMCChunk* c = n->chunk;
	if (!(c && c->contains_mesh && c->vi)) return;
	if (vertices) memcpy(vertices, c->vi->vertices.elements, c->vi->vertices.count * sizeof(DualVertex));
	if (indices) memcpy(indices, c->vi->mesh_indexes.elements, c->vi->mesh_indexes.count * sizeof(uint32_t));
	if (n->gl_chunk)
	{
		if (p_data) memcpy(p_data, n->gl_chunk->p_data.elements, n->gl_chunk->p_data.count * sizeof(glm::vec3));
		if (c_data) memcpy(c_data, n->gl_chunk->c_data.elements, n->gl_chunk->c_data.count * sizeof(glm::vec3));
	}
}

} // extern "C"

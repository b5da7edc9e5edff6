// host_test.cpp -- drives the C++ mirror (bmf_host.hpp) exactly the way the reference's own callers do and prints
// counts + zlib-compatible CRC32s of the produced buffers; tests/test_gpu_host_shim.py compares them with the oracle.
//   host_test chunk <kind> <dim> <overlap> <iters>          DMCChunk staged calls + MeshProcessor<3> (ChunkGenerator.cpp:98-124)
//   host_test world <kind> <dim> <max_level> <iters> <file> ChunkGenerator::process_queue over the leaves listed in <file>
//   host_test hostfn <dim> [density.bin]                                a Sampler with an arbitrary host callback (HOST_DENSITY escape path)
#include "bmf_host.hpp"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>

static uint32_t crc32_of(const void* data, size_t n, uint32_t crc = 0)
{
	static uint32_t table[256];
	static bool init = false;
	if (!init)
	{
		for (uint32_t i = 0; i < 256; i++)
		{
			uint32_t c = i;
			for (int k = 0; k < 8; k++) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
			table[i] = c;
		}
		init = true;
	}
	crc = ~crc;
	const uint8_t* p = (const uint8_t*)data;
	for (size_t i = 0; i < n; i++) crc = table[(crc ^ p[i]) & 0xFF] ^ (crc >> 8);
	return ~crc;
}

static Sampler make_sampler(int kind)
{
	Sampler s;
	switch (kind)
	{
	case 0: s = ImplicitFunctions::create_sampler(ImplicitFunctions::sphere); break;
	case 1: s = ImplicitFunctions::create_sampler(ImplicitFunctions::torus_z); break;
	case 2: s = ImplicitFunctions::create_sampler(ImplicitFunctions::cuboid); break;
	case 3: s = ImplicitFunctions::create_sampler(ImplicitFunctions::plane_y); break;
	case 10: NoiseSamplers::create_sampler_terrain_2d(&s); break;
	case 11: NoiseSamplers::create_sampler_terrain_pert_2d(&s); break;
	case 12: NoiseSamplers::create_sampler_terrain_3d(&s); break;
	default: NoiseSamplers::create_sampler_terrain_pert_3d(&s); break;
	}
	s.world_size = 256;
	return s;
}

static void print_chunk(const char* tag, DMCChunk& c)
{
	std::vector<float> p;
	std::vector<uint8_t> bnd, val;
	size_t nv = c.vi ? c.vi->vertices.count : 0, ni = c.vi ? c.vi->mesh_indexes.count : 0;
	for (size_t i = 0; i < nv; i++)
	{
		const DualVertex& v = c.vi->vertices.elements[i];
		p.push_back(v.p.x); p.push_back(v.p.y); p.push_back(v.p.z);
		bnd.push_back(v.boundary ? 1 : 0);
		val.push_back(v.init_valence);
	}
	size_t n = (size_t)c.dim * c.dim * c.dim;
	printf("%s contains_mesh=%d cells=%zu verts=%zu inds=%zu bits_crc=%u density_crc=%u inds_crc=%u pos_crc=%u boundary_crc=%u valence_crc=%u scale=%.9g\n", tag, (int)c.contains_mesh,
	       c.cell_block ? c.cell_block->cells.count : (size_t)0, nv, ni, c.binary_block ? crc32_of(c.binary_block->data, n / 8) : 0u,
	       c.density_block ? crc32_of(c.density_block->data, n * 4) : 0u, ni ? crc32_of(c.vi->mesh_indexes.elements, ni * 4) : 0u, nv ? crc32_of(p.data(), p.size() * 4) : 0u,
	       nv ? crc32_of(bnd.data(), nv) : 0u, nv ? crc32_of(val.data(), nv) : 0u, c.scale);
}

static const float wavy(const float ws, const glm::vec3& p) { return 20.0f * std::sin(p.x * 0.05f) * std::cos(p.z * 0.07f) - p.y + ws * 0.0f; }

int main(int argc, char** argv)
{
	if (argc < 2) return 2;
	if (!BmfDevice::get().ok())
	{
		fprintf(stderr, "host_test: %s\n", BmfDevice::get().error());
		return 3;
	}
	ResourceAllocator<BinaryBlock> binary_allocator;
	ResourceAllocator<DensityBlock> density_allocator;
	ResourceAllocator<NoiseBlock> noise_allocator;
	ResourceAllocator<VerticesIndicesBlock> vi_allocator;
	ResourceAllocator<DMC_CellsBlock> cell_allocator;
	ResourceAllocator<IndexesBlock> inds_allocator;
	ResourceAllocator<MasksBlock> masks_allocator;
	NoiseSamplers::NoiseSamplerProperties props;

	if (!strcmp(argv[1], "chunk") && argc >= 6)
	{
		int kind = atoi(argv[2]), dim = atoi(argv[3]), iters = atoi(argv[5]);
		float overlap = (float)atof(argv[4]);
		Sampler s = make_sampler(kind);
		DMCChunk c(glm::vec3(-128, -128, -128), 256.0f, 0, s, 1);
		c.dim = dim;
		c.label_grid(&binary_allocator, &density_allocator, &noise_allocator, overlap, props);
		c.label_edges(&vi_allocator, &cell_allocator, &inds_allocator, &density_allocator, &masks_allocator);
		size_t valence_before = 0;
		if (c.vi)
			for (size_t i = 0; i < c.vi->vertices.count; i++) valence_before += c.vi->vertices.elements[i].init_valence;
		c.polygonize();
		print_chunk("extract", c);
		printf("valence_sum_before_polygonize=%zu\n", valence_before);
		if (iters > 0 && c.contains_mesh && c.vi->vertices.count && c.vi->mesh_indexes.count)
		{
			// verbatim shape of ChunkGenerator.cpp:112-123
			auto& v_out = c.vi->vertices;
			auto& i_out = c.vi->mesh_indexes;
			Processing::MeshProcessor<3> mp(true, false);
			mp.init(c.vi->vertices, c.vi->mesh_indexes, s);
			mp.optimize_dual_grid(iters, false);
			mp.optimize_primal_grid(false, false, false);
			v_out.count = 0;
			i_out.count = 0;
			mp.flush(v_out, i_out);
			print_chunk("processed", c);
		}
		return 0;
	}
	if (!strcmp(argv[1], "interleave") && argc >= 4)
	{
		// the reference drives label_grid / label_edges / polygonize per chunk inside `#pragma omp parallel for`
		// (ChunkGenerator.cpp:93-108): chunks interleave on the shared device context.  mode 0: grid A, grid B, edges A, poly A,
		// edges B, poly B on one thread; mode 1: one thread per chunk.  Every chunk must come out as if it had run alone.
		const int dim = atoi(argv[2]), mode = atoi(argv[3]);
		Sampler sa = make_sampler(BMF_SAMPLER_SPHERE), sb = make_sampler(BMF_SAMPLER_TORUS_Z);
		DMCChunk a(glm::vec3(-128, -128, -128), 256.0f, 0, sa, 1), b(glm::vec3(-128, -128, -128), 256.0f, 0, sb, 1);
		a.dim = b.dim = dim;
		auto stage12 = [&](DMCChunk& c) {
			c.label_edges(&vi_allocator, &cell_allocator, &inds_allocator, &density_allocator, &masks_allocator);
			c.polygonize();
		};
		if (mode == 0)
		{
			a.label_grid(&binary_allocator, &density_allocator, &noise_allocator, 0.0f, props);
			b.label_grid(&binary_allocator, &density_allocator, &noise_allocator, 0.0f, props);
			stage12(a);
			stage12(b);
		}
		else
		{
			std::thread ta([&] { a.label_grid(&binary_allocator, &density_allocator, &noise_allocator, 0.0f, props); stage12(a); });
			std::thread tb([&] { b.label_grid(&binary_allocator, &density_allocator, &noise_allocator, 0.0f, props); stage12(b); });
			ta.join();
			tb.join();
		}
		print_chunk("A", a);
		print_chunk("B", b);
		return 0;
	}
	if (!strcmp(argv[1], "gradient") && argc >= 4)
	{
		// Sampler::gradient per point on the host (the reference's way, ImplicitSampler.hpp:38-49) against the device block evaluation
		const int kind = atoi(argv[2]), m = atoi(argv[3]);
		Sampler s = make_sampler(kind);
		std::vector<glm::vec3> p(m), dev(m);
		uint32_t r = 12345u;
		auto rnd = [&] { r = r * 1664525u + 1013904223u; return ((float)(r >> 8) / 16777216.0f - 0.5f) * 300.0f; };
		for (int i = 0; i < m; i++) p[i] = glm::vec3(rnd(), rnd(), rnd());
		if (!sampler_gradient_block(s, p.data(), p.size(), 0.01f, dev.data())) { fprintf(stderr, "sampler_gradient_block failed: %s\n", BmfDevice::get().error()); return 4; }
		int mismatches = 0;
		for (int i = 0; i < m; i++)
		{
			const glm::vec3 h = s.gradient(s.world_size, p[i], 0.01f);
			if (memcmp(&h, &dev[i], sizeof(h))) mismatches++;
		}
		printf("gradient n=%d mismatches=%d crc=%u\n", m, mismatches, crc32_of(dev.data(), sizeof(glm::vec3) * (size_t)m));
		return 0;
	}
	if (!strcmp(argv[1], "quadpost") && argc >= 4)
	{
		// the sequence the reference's author left commented out (DebugScene.cpp:253-262 / ChunkGenerator.cpp:271-281):
		// MeshProcessor<4> mp; mp.init(v, i); mp.collapse_bad_quads(); mp.flush_to_tris(v_out, i_out); + ColorMapper::generate_colors
		// input: a quad mesh written by the test (n_verts, n_quads, then floats / uint32s)
		FILE* f = fopen(argv[2], "rb");
		if (!f) return 5;
		uint32_t hdr[2];
		if (fread(hdr, 4, 2, f) != 2) return 5;
		std::vector<float> pos(3 * (size_t)hdr[0]);
		std::vector<uint32_t> q(4 * (size_t)hdr[1]);
		if (fread(pos.data(), 4, pos.size(), f) != pos.size() || fread(q.data(), 4, q.size(), f) != q.size()) return 5;
		fclose(f);
		SmartContainer<DualVertex> v, v_out;
		SmartContainer<uint32_t> idx, i_out;
		std::vector<uint8_t> val(hdr[0], 0);
		for (uint32_t k : q) val[k]++;
		for (uint32_t i = 0; i < hdr[0]; i++)
		{
			DualVertex dv;
			memset((void*)&dv, 0, sizeof(dv));
			dv.index = i; dv.init_valence = val[i];
			dv.p = glm::vec3(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
			v.push_back(dv);
		}
		idx.push_back(q.data(), q.size());
		Sampler s = make_sampler(0);
		Processing::MeshProcessor<4> mp(true, false);
		mp.init(v, idx, s);
		mp.collapse_bad_quads();
		const bool tris = atoi(argv[3]) != 0;
		if (tris) mp.flush_to_tris(v_out, i_out); else mp.flush(v_out, i_out);
		ColorMapper cm;
		cm.generate_colors(v_out);
		std::vector<float> p, c;
		std::vector<uint8_t> an;
		for (size_t i = 0; i < v_out.count; i++)
		{
			const DualVertex& dv = v_out.elements[i];
			p.push_back(dv.p.x); p.push_back(dv.p.y); p.push_back(dv.p.z);
			c.push_back(dv.color.x); c.push_back(dv.color.y); c.push_back(dv.color.z);
			an.push_back(dv.adj_next);
		}
		printf("quadpost bad=%u verts=%zu inds=%zu inds_crc=%u pos_crc=%u color_crc=%u adj_next_crc=%u\n", mp.bad_quads, v_out.count, i_out.count,
		       crc32_of(i_out.elements, i_out.count * 4), crc32_of(p.data(), p.size() * 4), crc32_of(c.data(), c.size() * 4), crc32_of(an.data(), an.size()));
		return 0;
	}
	if (!strcmp(argv[1], "hostfn") && argc >= 3)
	{
		int dim = atoi(argv[2]);
		Sampler s = ImplicitFunctions::create_sampler(wavy); // not a known primitive -> host callback + HOST_DENSITY
		s.world_size = 256;
		DMCChunk c(glm::vec3(-64, -64, -64), 128.0f, 0, s, 1);
		c.dim = dim;
		c.label_grid(&binary_allocator, &density_allocator, &noise_allocator, 0.0f, props);
		c.label_edges(&vi_allocator, &cell_allocator, &inds_allocator, &density_allocator, &masks_allocator);
		c.polygonize();
		print_chunk("hostfn", c);
		if (argc >= 4)
		{
			FILE* f = fopen(argv[3], "wb"); // the host-evaluated density, so the test can feed the same samples to the oracle
			fwrite(c.density_block->data, sizeof(float), (size_t)dim * dim * dim, f);
			fclose(f);
		}
		return 0;
	}
	if (!strcmp(argv[1], "lod") && argc >= 6)
	{
		// SURVEY call stack B, all in C++: WorldOctree::init -> split_leaves -> ChunkGenerator::process_queue;
		// optional 6th argument: comma-separated device list for the multi-GPU partition, e.g. 0,1 or 0,0,0
		int kind = atoi(argv[2]), dim = atoi(argv[3]), max_level = atoi(argv[4]), iters = atoi(argv[5]);
		WorldOctree world;
		world.sampler = make_sampler(kind);
		world.properties.chunk_resolution = dim;
		world.properties.max_level = max_level;
		world.properties.process_iters = iters;
		world.init(256);
		world.split_leaves();
		ChunkGenerator gen;
		gen.init(&world);
		if (argc >= 7)
		{
			std::vector<int> devs;
			for (char* tok = strtok(argv[6], ","); tok; tok = strtok(nullptr, ",")) devs.push_back(atoi(tok));
			gen.set_devices(devs);
		}
		SmartContainer<WorldOctreeNode*> batch;
		for (WorldOctreeNode* n : world.leaves)
		{
			n->generation_stage = GENERATION_STAGES_GENERATING;
			batch.push_back(n);
		}
		if (!gen.process_queue(batch))
		{
			fprintf(stderr, "process_queue failed: %s\n", BmfDevice::get().error());
			return 5;
		}
		size_t nm = 0, nv = 0, ni = 0;
		uint32_t hi = 0, hl = 0;
		for (WorldOctreeNode* n : world.leaves)
		{
			float rec[4] = { n->pos.x, n->pos.y, n->pos.z, n->size };
			hl = crc32_of(rec, sizeof(rec), hl);
			DMCChunk* c = n->chunk;
			if (!(c->contains_mesh && c->vi)) continue;
			nm += c->vi->vertices.count ? 1 : 0;
			nv += c->vi->vertices.count;
			ni += c->vi->mesh_indexes.count;
			hi = crc32_of(c->vi->mesh_indexes.elements, c->vi->mesh_indexes.count * 4, hi);
		}
		printf("lod chunks=%zu with_mesh=%zu verts=%zu inds=%zu inds_crc=%u leaves_crc=%u\n", world.leaves.size(), nm, nv, ni, hi, hl);
		return 0;
	}
	if (!strcmp(argv[1], "stitch") && argc >= 8)
	{
		// LOD world with enable_stitching: process_queue samples at voxel-node centres, stitcher.stitch_all closes the seams
		int kind = atoi(argv[2]), dim = atoi(argv[3]), max_level = atoi(argv[4]);
		WorldOctree world;
		world.sampler = make_sampler(kind);
		world.properties.chunk_resolution = dim;
		world.properties.max_level = max_level;
		world.properties.enable_stitching = true;
		world.focus_point = glm::vec3((float)atof(argv[5]), (float)atof(argv[6]), (float)atof(argv[7]));
		world.init(256);
		world.split_leaves();
		ChunkGenerator gen;
		gen.init(&world);
		SmartContainer<WorldOctreeNode*> batch;
		for (WorldOctreeNode* n : world.leaves)
		{
			n->generation_stage = GENERATION_STAGES_GENERATING;
			batch.push_back(n);
		}
		// optional 9th argument: device list for the multi-GPU seam scheme, e.g. 0,0 (two contexts on GPU 0)
		std::vector<int> sdevs;
		if (argc >= 9)
			for (char* tok = strtok(argv[8], ","); tok; tok = strtok(nullptr, ",")) sdevs.push_back(atoi(tok));
		if (!gen.process_queue(batch) || !(sdevs.empty() ? gen.stitcher.stitch_all(&world) : gen.stitcher.stitch_all(&world, sdevs)))
		{
			fprintf(stderr, "stitch failed: %s\n", BmfDevice::get().error());
			return 5;
		}
		gen.stitcher.format();
		size_t nv = 0, ni = 0;
		for (WorldOctreeNode* n : world.leaves)
			if (n->chunk->contains_mesh && n->chunk->vi) { nv += n->chunk->vi->vertices.count; ni += n->chunk->vi->mesh_indexes.count; }
		const uint32_t hp = crc32_of(gen.stitcher.gl_chunk.p_data.elements, gen.stitcher.gl_chunk.p_data.count * 12, 0);
		// order-independent checksum over the triangles (the multi-device scheme emits them in another order)
		unsigned long long tsum = 0;
		for (size_t t = 0; t + 2 < gen.stitcher.gl_chunk.p_data.count; t += 3) tsum += crc32_of(gen.stitcher.gl_chunk.p_data.elements + t, 36, 0);
		printf("stitch chunks=%zu verts=%zu inds=%zu seam_verts=%zu seam_crc=%u tri_sum=%llu color_g=%g\n", world.leaves.size(), nv, ni, gen.stitcher.vertices.count, hp, tsum,
		       gen.stitcher.vertices.count ? (double)gen.stitcher.vertices[0].color.y : 0.0);
		return 0;
	}
	if (!strcmp(argv[1], "watch") && argc >= 6)
	{
		// the reference's own start: only the root is drawable; one tick per focus point read from a file ("x y z" per line)
		int kind = atoi(argv[2]), dim = atoi(argv[3]), max_level = atoi(argv[4]);
		WorldOctree world;
		world.sampler = make_sampler(kind);
		world.properties.chunk_resolution = dim;
		world.properties.max_level = max_level;
		world.init(256);
		WorldWatcher watcher;
		watcher.init(&world, glm::vec3(0, 0, 0), true);
		FILE* f = fopen(argv[5], "r");
		if (!f) return 4;
		float x, y, z;
		printf("watch gens=");
		while (fscanf(f, "%f %f %f", &x, &y, &z) == 3)
		{
			watcher.focus_pos = glm::vec3(x, y, z);
			if (!watcher.update()) return 5;
			printf("%zu,", watcher.last_generated);
		}
		fclose(f);
		uint32_t hc = 0;
		for (WorldOctreeNode* n : watcher.renderables)
		{
			unsigned long long code = n->morton_code;
			hc = crc32_of(&code, sizeof(code), hc);
		}
		printf(" leaves=%zu codes_crc=%u\n", watcher.renderables.size(), hc);
		return 0;
	}
	if (!strcmp(argv[1], "fly") && argc >= 9)
	{
		// WorldWatcher ticks while the focus moves from the origin to (fx, fy, fz) in `steps` equal steps, then until quiescent
		int kind = atoi(argv[2]), dim = atoi(argv[3]), max_level = atoi(argv[4]), steps = atoi(argv[8]);
		const glm::vec3 target((float)atof(argv[5]), (float)atof(argv[6]), (float)atof(argv[7]));
		WorldOctree world;
		world.sampler = make_sampler(kind);
		world.properties.chunk_resolution = dim;
		world.properties.max_level = max_level;
		world.init(256);
		world.split_leaves();
		WorldWatcher watcher;
		watcher.init(&world, glm::vec3(0, 0, 0));
		{
			// the static world first (WorldOctree::generate_outline / the first watcher batch)
			SmartContainer<WorldOctreeNode*> all;
			for (WorldOctreeNode* n : world.leaves)
			{
				n->generation_stage = GENERATION_STAGES_GENERATING;
				all.push_back(n);
			}
			if (!watcher.generator.process_queue(all)) return 5;
		}
		size_t generated = 0, ticks = 0;
		for (int k = 1; k <= steps + 64; k++)
		{
			const float t = k >= steps ? 1.0f : (float)k / (float)steps;
			watcher.focus_pos = glm::vec3(target.x * t, target.y * t, target.z * t);
			if (!watcher.update())
			{
				fprintf(stderr, "update failed: %s\n", BmfDevice::get().error());
				return 5;
			}
			generated += watcher.last_generated;
			ticks++;
			if (k >= steps && watcher.last_generated == 0) break;
		}
		uint32_t hc = 0;
		size_t nv = 0, ni = 0;
		for (WorldOctreeNode* n : watcher.renderables)
		{
			unsigned long long code = n->morton_code;
			hc = crc32_of(&code, sizeof(code), hc);
			if (n->chunk && n->chunk->contains_mesh && n->chunk->vi) { nv += n->chunk->vi->vertices.count; ni += n->chunk->vi->mesh_indexes.count; }
		}
		printf("fly leaves=%zu generated=%zu ticks=%zu codes_crc=%u verts=%zu inds=%zu\n", watcher.renderables.size(), generated, ticks, hc, nv, ni);
		return 0;
	}
	if (!strcmp(argv[1], "world") && argc >= 7)
	{
		int kind = atoi(argv[2]), dim = atoi(argv[3]), max_level = atoi(argv[4]), iters = atoi(argv[5]);
		WorldOctree world;
		world.sampler = make_sampler(kind);
		world.properties.chunk_resolution = dim;
		world.properties.max_level = max_level;
		world.properties.process_iters = iters;
		ChunkGenerator gen;
		gen.init(&world);
		FILE* f = fopen(argv[6], "r");
		if (!f) return 4;
		std::vector<WorldOctreeNode*> nodes;
		float x, y, z, sz;
		int lvl;
		unsigned long long code;
		SmartContainer<WorldOctreeNode*> batch;
		while (fscanf(f, "%f %f %f %f %d %llu", &x, &y, &z, &sz, &lvl, &code) == 6)
		{
			WorldOctreeNode* n = new WorldOctreeNode(sz, glm::vec3(x, y, z), (uint8_t)lvl, code);
			n->generation_stage = GENERATION_STAGES_GENERATING;
			nodes.push_back(n);
			batch.push_back(n);
		}
		fclose(f);
		if (!gen.process_queue(batch))
		{
			fprintf(stderr, "process_queue failed: %s\n", BmfDevice::get().error());
			return 5;
		}
		size_t nm = 0, nv = 0, ni = 0, upload = 0;
		uint32_t hi = 0, hp = 0;
		for (WorldOctreeNode* n : nodes)
		{
			DMCChunk* c = n->chunk;
			if (n->generation_stage == GENERATION_STAGES_NEEDS_UPLOAD) upload++;
			if (!(c->contains_mesh && c->vi)) continue;
			nm += c->vi->vertices.count ? 1 : 0;
			nv += c->vi->vertices.count;
			ni += c->vi->mesh_indexes.count;
			hi = crc32_of(c->vi->mesh_indexes.elements, c->vi->mesh_indexes.count * 4, hi);
			hp = crc32_of(n->gl_chunk->p_data.elements, n->gl_chunk->p_data.count * 12, hp);
		}
		printf("world chunks=%zu with_mesh=%zu verts=%zu inds=%zu inds_crc=%u pdata_crc=%u needs_upload=%zu\n", nodes.size(), nm, nv, ni, hi, hp, upload);
		return 0;
	}
	return 2;
}

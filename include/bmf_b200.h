/*
 * bmf_b200.h -- C ABI of the B200-native chunk-extraction path (libbmf_b200.so).
 *
 * The reference (Lin20/BinaryMeshFitting) has no FFI: the hot path sits behind C++ classes compiled
 * into one executable.  This header is the drop-in boundary those classes are re-hosted
 * on: plain pointers and sizes, no C++ or torch types.  Each entry point names the reference
 * interface it replaces (file:line under /root/reference/BinaryMeshFitting).  The C++ mirror of the
 * reference classes (Sampler / DMCChunk / ChunkGenerator / Processing::MeshProcessor) that sits on
 * top of this ABI is binarymeshfitting_b200/host/bmf_host.hpp; INTEGRATION.md shows the binding.
 *
 * Conventions: every function returns 0 on success, a negative bmf_status otherwise;
 * bmf_last_error(ctx) holds the message.  One bmf_ctx per GPU; a ctx is single-threaded (the
 * reference calls ChunkGenerator::process_queue from exactly one thread, WorldWatcher.cpp:71).
 * The caller owns host buffers; the library owns device arenas.  There is NO CPU fallback: without a
 * CUDA device every compute entry point fails with BMF_ERR_CUDA.
 */
#ifndef BMF_B200_H
#define BMF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum bmf_status
{
	BMF_OK = 0,
	BMF_ERR_INVALID = -1, /* bad argument (dim not in {32,64,128,256}, null pointer, count out of range ...) */
	BMF_ERR_CUDA = -2,    /* CUDA runtime error (message in bmf_last_error) */
	BMF_ERR_STATE = -3,   /* call order (no batch submitted, sampler not set ...) */
	BMF_ERR_NOMEM = -4
} bmf_status;

/* Sampler.hpp:24-34 -- a GPU can only run *known* samplers, so the callback bundle becomes a
 * descriptor.  HOST_DENSITY is the escape hatch for arbitrary `Sampler::block` callbacks: the host
 * fills the density block, the device does everything after it (also the bit-exact parity mode). */
typedef enum bmf_sampler_kind
{
	BMF_SAMPLER_SPHERE = 0,          /* ImplicitSampler.cpp:46-50 */
	BMF_SAMPLER_TORUS_Z = 1,         /* ImplicitSampler.cpp:27-34 */
	BMF_SAMPLER_CUBOID = 2,          /* ImplicitSampler.cpp:52-60 */
	BMF_SAMPLER_PLANE_Y = 3,         /* ImplicitSampler.cpp:62-65 */
	BMF_SAMPLER_CSG = 4,             /* two primitives combined (config 5; build-defined, unpinned) */
	BMF_SAMPLER_TERRAIN2D = 10,      /* NoiseSampler.cpp:113-146 */
	BMF_SAMPLER_TERRAIN2D_PERT = 11, /* NoiseSampler.cpp:148-194 (the reference world's default) */
	BMF_SAMPLER_TERRAIN3D = 12,      /* NoiseSampler.cpp:196-227 */
	BMF_SAMPLER_TERRAIN3D_PERT = 13, /* NoiseSampler.cpp:229-263 */
	BMF_SAMPLER_HOST_DENSITY = 100
} bmf_sampler_kind;

typedef enum bmf_csg_op { BMF_CSG_UNION = 0, BMF_CSG_INTERSECT = 1, BMF_CSG_SUBTRACT = 2 } bmf_csg_op;

typedef struct bmf_sampler_desc
{
	int32_t kind;     /* bmf_sampler_kind */
	float world_size; /* Sampler::world_size (Sampler.hpp:26) */
	/* NoiseSamplers::NoiseSamplerProperties (NoiseSampler.hpp:7-22); defaults WorldOctree.cpp:47-54 */
	float g_scale, height;
	int32_t octaves;
	float amp, frequency, gain;
	int32_t seed; /* FastNoiseSIMD::NewFastNoiseSIMD(seed), library default 1337 */
	/* BMF_SAMPLER_CSG: value = op(a(p - offset_a), b(p - offset_b)), positive = inside */
	int32_t csg_op, csg_kind_a, csg_kind_b;
	float csg_world_size_a, csg_world_size_b;
	float csg_offset_a[3], csg_offset_b[3];
} bmf_sampler_desc;

/* DMCChunk::init(pos, size, level, sampler, parent_code) (DMCChunk.cpp:58-77) + the per-chunk overlap
 * ChunkGenerator::extract_chunk derives (ChunkGenerator.cpp:98) */
typedef struct bmf_chunk_desc
{
	float pos[3];
	float size;
	int32_t level;
	float overlap;
	uint64_t morton; /* WorldOctreeNode::morton_code (carried through, not interpreted) */
} bmf_chunk_desc;

/* WorldProperties / DefaultOptions knobs that reach the hot path (WorldOctree.cpp:20-33,
 * DefaultOptions.h:7-8, ChunkGenerator.cpp:112-124) */
typedef struct bmf_params
{
	int32_t dim;              /* chunk_resolution: 32, 64, 128 or 256 */
	int32_t iters;            /* process_iters: MeshProcessor<3> iterations, 0 = none */
	int32_t process_boundary; /* boundary_processing */
	int32_t smooth_normals;   /* SMOOTH_NORMALS */
	int32_t qef;              /* after smoothing, re-place each vertex by the QEF (qef_simd.h:411-579) of its adjacent primitives' planes (build-defined policy):
	                           * 1 = (dual_p, face normal); 2 = (dual_p, normalised Sampler::gradient at dual_p, h = 0.01) -- the line the reference
	                           * left commented out at MeshProcessor.cpp:224; analytic samplers only */
	int32_t keep_density;     /* 1: materialise the f32 density block (DMCChunk::density_block) so it can be copied out */
	int32_t keep_masks;       /* 1: materialise the 8-bit cell masks (MasksBlock image) so they can be copied out */
	int32_t density_on_device; /* HOST_DENSITY only: the density pointer passed to submit is a device pointer */
	int32_t quads;            /* 1: emit QUADS instead of triangles (Nielson dual marching cubes: one vertex per surface patch of a cell,
	                             one quad = 4 indices per sign-changing edge; build-defined, the reference has no quad producer) and run
	                             MeshProcessor<4> for iters > 0; n_inds counts 4 per quad.  bmf_quads_to_tris = flush_to_tris */
} bmf_params;

/* bmf_chunk_info.flags: what a consumer may synthesise instead of receiving (bmf_batch_download_enqueue can skip these streams) */
enum
{
	BMF_CHUNK_COLOR_ONE = 1,   /* every vertex colour of the chunk is exactly (1,1,1) (calculate_dual_vertex, DMCChunk.cpp:681; smoothing keeps it) */
	BMF_CHUNK_NORMAL_ZERO = 2, /* every vertex normal of the chunk is exactly (0,0,0) (no step of this batch writes v.n) */
	BMF_CHUNK_INDEX16 = 4      /* n_verts < 65536: the chunk-local indices fit uint16 */
};

typedef struct bmf_chunk_info
{
	int32_t contains_mesh; /* DMCChunk::contains_mesh (DMCChunk.cpp:159-162) */
	int32_t n_cells, n_verts, n_inds;
	int64_t vert_offset, ind_offset; /* into the batch-wide SoA arrays */
	float overlap_pos[3];            /* DMCChunk::overlap_pos (DMCChunk.cpp:97) */
	float scale;                     /* DMCChunk::scale = delta (DMCChunk.cpp:94,98) */
	int32_t flags;                   /* BMF_CHUNK_* */
	int32_t reserved;
} bmf_chunk_info;

enum
{
	BMF_STAGE_SAMPLE = 0, /* K1 sampling + sign pack (or K2 pack from a supplied density block) */
	BMF_STAGE_COUNT = 1,  /* K3 cell-mask build + per-word counts */
	BMF_STAGE_SCAN = 2,   /* segment scan */
	BMF_STAGE_VERTS = 3,  /* K4 vertex emission */
	BMF_STAGE_INDS = 4,   /* K4 index emission + valence */
	BMF_STAGE_SMOOTH = 5, /* K5 CSR + dual/primal iterations (+K6 QEF placement) */
	BMF_STAGE_TOTAL = 6,  /* first kernel start to last kernel end */
	BMF_NUM_STAGES = 7
};

typedef struct bmf_ctx bmf_ctx;

/* library */
int bmf_ctx_create(int device, bmf_ctx** out);
void bmf_ctx_destroy(bmf_ctx* ctx);
const char* bmf_last_error(const bmf_ctx* ctx);
const char* bmf_version(void);

/* Sampler (Sampler.hpp:24-34; factories ImplicitSampler.hpp:51-59, NoiseSampler.hpp:82-140) */
void bmf_sampler_defaults(bmf_sampler_desc* desc, int kind);
int bmf_sampler_set(bmf_ctx* ctx, const bmf_sampler_desc* desc);

/* ChunkGenerator::process_queue / extract_chunk (ChunkGenerator.cpp:27-60, 80-147): the whole batch
 * through label_grid -> label_edges -> polygonize -> MeshProcessor<3>.  Asynchronous on the ctx stream.
 * `density` is required (n * dim^3 floats, chunk-major, [x][y][z] z fastest) for HOST_DENSITY, else NULL.
 * One batch per ctx is resident at a time; submitting again recycles the arenas. */
int bmf_batch_submit(bmf_ctx* ctx, const bmf_chunk_desc* chunks, int n, const bmf_params* params, const float* density);
int bmf_batch_wait(bmf_ctx* ctx);
int bmf_batch_totals(bmf_ctx* ctx, int64_t* n_cells, int64_t* n_verts, int64_t* n_inds);
int bmf_batch_chunk_info(bmf_ctx* ctx, int i, bmf_chunk_info* out);
int bmf_batch_chunk_infos(bmf_ctx* ctx, bmf_chunk_info* out /* [n] */);

/* The renderer-facing SoA of the whole batch (what GLChunk::format_data packs, GLChunk.cpp:278-296:
 * p_data / n_data / c_data + the index buffer), chunk i at [vert_offset, +n_verts) / [ind_offset, +n_inds).
 * Indices are chunk-local like DMCChunk::vi->mesh_indexes.  Any pointer may be NULL. */
int bmf_batch_download(bmf_ctx* ctx, float* pos, float* normal, float* color, uint8_t* boundary, uint8_t* valence, uint32_t* indices);
/* same, but only enqueues the copies on the ctx stream (use pinned host buffers); bmf_batch_wait completes them.
 * Two contexts on one GPU ping-pong this way: the copies of batch i overlap the kernels of batch i+1. */
int bmf_batch_download_async(bmf_ctx* ctx, float* pos, float* normal, float* color, uint8_t* boundary, uint8_t* valence, uint32_t* indices);

/* Host-asynchronous, device-driven download (csrc/download.cuh): a kernel on the ctx stream stores the batch's arrays straight
 * into the caller's PINNED host buffers, sized by the totals it reads on the device -- the call neither waits for the batch nor
 * for the copy; bmf_batch_wait (or any query) completes both and reports a buffer that was too small (BMF_ERR_NOMEM, nothing
 * written) or uint16 indices that do not fit (BMF_ERR_INVALID).  Buffers must come from bmf_host_alloc or be registered with
 * bmf_host_register (page-locked and device-mapped); they should be 16-byte aligned.  This is the OPT-IN compact form of
 * GLChunk::format_data's output (GLChunk.cpp:278-296): leave a stream NULL to skip it -- bmf_chunk_info.flags says which ones
 * are constant and can be synthesised by the consumer -- and pass indices16 instead of indices32 to receive the chunk-local
 * indices as uint16 (valid when every chunk of the batch has < 65536 vertices; bmf_chunk_info.flags & BMF_CHUNK_INDEX16).
 * bmf_batch_download stays the reference-layout default. */
typedef struct bmf_download_desc
{
	float* pos;          /* [cap_verts][3] or NULL */
	float* normal;       /* [cap_verts][3] or NULL */
	float* color;        /* [cap_verts][3] or NULL */
	uint8_t* boundary;   /* [cap_verts] or NULL */
	uint8_t* valence;    /* [cap_verts] or NULL */
	uint32_t* indices32; /* [cap_inds] or NULL */
	uint16_t* indices16; /* [cap_inds] or NULL (at most one of indices32 / indices16) */
	int64_t cap_verts, cap_inds; /* capacity of the buffers above, in vertices / indices */
} bmf_download_desc;
int bmf_batch_download_enqueue(bmf_ctx* ctx, const bmf_download_desc* desc);
/* The same streams through the COPY ENGINE: the call first completes the batch's kernels (the host waits for them -- that is the price: the
 * sizes of the transfers must be known to the host), then enqueues DMA transfers of exactly the batch's sizes on the ctx stream and returns;
 * bmf_batch_wait completes them.  uint16 indices are packed on the device first.  Buffers should be page-locked (bmf_host_alloc / register);
 * errors are reported at once (BMF_ERR_NOMEM: buffers too small; BMF_ERR_INVALID: a chunk has >= 65536 vertices and indices16 was asked).
 * Measured on B200 / PCIe 5: 56 GB/s and no slow-down of another context's kernels, against 51 GB/s for the kernel-driven form, whose
 * stores to host memory slow concurrent kernels down by ~2.7x -- so pipelines that overlap batch i's download with batch i+1's kernels
 * (several contexts round robin, bench.py's e2e) use this call, and bmf_batch_download_enqueue is for callers that must not block the host. */
int bmf_batch_download_dma(bmf_ctx* ctx, const bmf_download_desc* desc);
/* page-locked, device-mapped host memory for the call above (cudaHostAlloc portable + mapped) -- so that a host program needs no
 * CUDA headers -- and registration of memory the caller already owns (e.g. a shared-memory segment several ranks write into) */
int bmf_host_alloc(size_t bytes, void** out);
void bmf_host_free(void* p);
int bmf_host_register(void* p, size_t bytes);
int bmf_host_unregister(void* p);

/* One chunk in the reference's own layouts: DualVertex[n_verts] (84-byte records, Vertices.hpp:5-24),
 * mesh_indexes, BinaryBlock words (dim^3/32), MasksBlock byte image (dim^3, needs keep_masks),
 * DensityBlock (dim^3 floats, needs keep_density or HOST_DENSITY).  Any pointer may be NULL. */
int bmf_batch_copy_chunk(bmf_ctx* ctx, int i, void* dual_vertices, uint32_t* indices, uint32_t* bits, uint8_t* masks, float* density);

/* per-stage device times of the last batch (CUDA events on the ctx stream), milliseconds */
int bmf_batch_stage_ms(bmf_ctx* ctx, float* ms /* [BMF_NUM_STAGES] */);
/* kernels launched by this ctx since creation (bench.py's gpu_launches) */
int64_t bmf_ctx_launch_count(const bmf_ctx* ctx);
/* Pipelines that keep two contexts busy on one GPU (batch i+1 meshing while batch i's device-driven download runs) call this with a small n
 * (8 is what bench.py uses): the persistent kernels of the ctx then size their grids for (SMs - n) SMs, so the other context's kernels find
 * free registers and are not queued behind this context's longest kernels.  0 (default) = use every SM.  No reference counterpart: the
 * reference's generator and its GL upload never overlap (WorldWatcher.cpp:105-112). */
int bmf_ctx_set_reserved_sms(bmf_ctx* ctx, int n);

/* Scheduling hint for callers that keep SEVERAL batches in flight on this device (one ctx = one stream each, like bench.py does): with
 * n > 1 the ctx picks, among kernel sets that give identical results, the one with the least total SM time -- the per-chunk kernels
 * (csrc/fused.cuh) at any batch size -- instead of the one with the shortest latency for a batch that has the GPU to itself (per-segment
 * kernels below 8 x SMs chunks).  Measured on a 512-chunk batch: 0.32 / 0.38 ms alone, 0.163 / 0.146 ms per batch with four in flight.
 * n <= 1 (default): latency first.  No reference counterpart (the reference meshes one queue at a time, ChunkGenerator.cpp:27-60). */
int bmf_ctx_set_batches_in_flight(bmf_ctx* ctx, int n);

/* optional per-launch timing: when on, every kernel of the next batches is bracketed by its own event pair;
 * bmf_ctx_kernel_times returns the number of launches of the last batch and fills up to `cap` (name, ms) */
int bmf_ctx_set_kernel_timing(bmf_ctx* ctx, int on);
int bmf_ctx_kernel_times(bmf_ctx* ctx, int cap, const char** names, float* ms);
/* the cudaStream_t every kernel of this ctx is launched on (so a harness can record its own events there) */
void* bmf_ctx_stream(const bmf_ctx* ctx);
/* device pointers of the resident batch (for zero-copy consumers / profiling); valid until next submit */
int bmf_batch_device_ptrs(bmf_ctx* ctx, void** pos, void** indices, void** bits, void** density);

/* Processing::MeshProcessor<N> stand-alone (MeshProcessor.hpp:55-84): init + optimize_dual_grid(iters,
 * pb) + optimize_primal_grid(false,false,pb) + flush, on caller arrays (updated in place).  N = 3 or 4. */
int bmf_mesh_process(bmf_ctx* ctx, float* pos, float* color, float* normal, const uint8_t* boundary, const uint8_t* valence,
                     int n_verts, const uint32_t* indices, int n_inds, int prim_n, int iters, int process_boundary, int smooth_normals);

/* same, with the trailing optimize_primal_grid(false,false,pb) of ChunkGenerator.cpp:120 optional (final_primal = 0:
 * optimize_dual_grid alone, MeshProcessor.cpp:130-236) */
int bmf_mesh_process_steps(bmf_ctx* ctx, float* pos, float* color, float* normal, const uint8_t* boundary, const uint8_t* valence,
                           int n_verts, const uint32_t* indices, int n_inds, int prim_n, int iters, int process_boundary, int smooth_normals,
                           int final_primal);

/* qef_solve_from_points_3d (qef_simd.h:550-579), m independent systems: system j reads counts[j]
 * (2..12) planes from positions/normals[j*12*3 ...]; writes out_pos[3*j..], out_err[j]. */
int bmf_qef_solve(bmf_ctx* ctx, const float* positions, const float* normals, const int32_t* counts, int m, float* out_pos, float* out_err);

/* Sampler::gradient (Sampler.hpp:29; implicit_gradient, ImplicitSampler.hpp:38-49 = NoiseSampler.hpp:35-47) of the ctx's sampler at
 * m world-space points: six evaluations of the sampler's value callback per point, raw differences (not normalised, not divided
 * by 2h; the reference's default h is 0.01).  points / out: [m*3] host arrays.  CSG: differences of the combinator; noise kinds:
 * their value callback is NoiseSamplers::noise3d == 0 (NoiseSampler.cpp:99-102), so the result is (0,0,0); HOST_DENSITY: BMF_ERR_STATE. */
int bmf_sampler_gradient(bmf_ctx* ctx, const float* points, int64_t m, float h, float* out);

/* ColorMapper::generate_colors (ColorMapper.cpp:15-60): per vertex, 4-octave simplex FBM (fresh FastNoiseSIMD object, seed 1337) at the
 * vertex position, n = noise * 4, colour = hsl_to_rgb((n + 1) * 0.5 * 360, 0.72, 1) (ColorMapper.cpp:62-121).  pos / color: [n*3] host arrays. */
int bmf_color_map(bmf_ctx* ctx, const float* pos, int64_t n, float* color);

/* Processing::MeshProcessor<4>::init + collapse_bad_quads (MeshProcessor.cpp:25-55, 98-128, 308-396) [+ flush, :57-71] on caller arrays.
 * A quad whose two opposite corners have exactly three adjacent quads is collapsed: the first of them moves to the quad's centre and
 * becomes valence 4, the other is rewired to it in the four surrounding quads, the quad is marked destroyed.  The reference's loop is
 * serial and order-dependent; the result here is that of the serial order.
 *   pos [n_verts*3] in/out; quads [n_quads*4] in/out (Primitive::v after the rewiring, all quads);
 *   destroyed [n_quads] out (Primitive::destroyed), adj_next [n_verts] out (DualVertex::adj_next) -- either may be NULL;
 *   flushed [n_quads*4] out (NULL ok): the corners of the surviving quads in order, what flush() appends; *n_flushed = their number (quads);
 *   *bad_count = the count the reference prints. */
int bmf_mesh_collapse_bad_quads(bmf_ctx* ctx, float* pos, int n_verts, uint32_t* quads, int64_t n_quads, uint8_t* destroyed, uint8_t* adj_next,
                                uint32_t* flushed, int64_t* n_flushed, int64_t* bad_count);

/* MeshProcessor<4>::flush_to_tris (MeshProcessor.cpp:73-91): every quad (v0,v1,v2,v3) becomes (v0,v1,v2),(v2,v3,v0).
 * quads: [n_quads*4] indices in, tris: [n_quads*6] out (host arrays). */
int bmf_quads_to_tris(bmf_ctx* ctx, const uint32_t* quads, int64_t n_quads, uint32_t* tris);

/* Measured issue-rate denominators for the FP32 / INT32 bound stages: chain steps per second, in 1e9, of
 * [0] FP32 FMA, [1] INT32 multiply-add, [2] INT32 logic+shift+add step, [3] FMA and the INT32 step interleaved (both counted). */
int bmf_ubench_issue(bmf_ctx* ctx, float* gops /* [4] */);

/* GLChunk::format_data(vertices, indexes, unwind_verts = true, smooth_normals) (GLChunk.cpp:278-335; DebugScene.cpp:286
 * with FLAT_QUADS) for the resident QUAD batch: per quad corner position / normal / colour, [n_inds][3] floats each; the
 * normal of a quad is the mean of its corner normals (smooth_normals) or the reference's two-triangle face normal.
 * Any pointer may be NULL. */
int bmf_batch_download_flat_quads(bmf_ctx* ctx, int smooth_normals, float* p_data, float* n_data, float* c_data);

/* WorldStitcher::stitch_all(root) (WorldStitcher.cpp:26-49 -> stitch_cell :184-239 -> stitch_indexes :491-572): the seam
 * pass over the RESIDENT batch, whose chunks must be aligned leaves of one octree (any mix of levels; a missing
 * neighbour simply gets no seam).  Every dual cell formed by 8 voxel nodes that do not all belong to one chunk is
 * polygonised with the marching-cubes table from the nodes' sample positions and densities.  The reference's version is
 * non-functional as committed, so the behaviour is defined by this build (UNPINNED; see csrc/seam.cuh): chunks are
 * expected to be sampled at voxel-node centres, i.e. submitted with overlap = bmf_seam_overlap(dim) = -1/(2 dim), so
 * that the chunk meshes and the seam tile space without gaps or overlaps.
 * group / cross_group_only: with per-chunk group ids (e.g. the GPU that meshed the chunk) and cross_group_only = 1 only
 * the cells whose nodes span more than one group are emitted -- the final pass after a multi-GPU gather.
 * Output: a non-indexed triangle soup in world coordinates (stitch_indexes :568-570 pushes DualVertex triples),
 * ordered by (chunk, lattice point, table order); zero-area triangles with two coincident corners are dropped. */
float bmf_seam_overlap(int dim);
int bmf_batch_stitch(bmf_ctx* ctx, const int32_t* group /* [n] or NULL */, int cross_group_only, int64_t* n_tris);
/* positions: [n_tris][3 corners][xyz] floats.  The reference colours every seam vertex (0.85, 1, 0.85) (:484-487). */
int bmf_seam_download(bmf_ctx* ctx, float* positions);
/* device times of the last seam pass: ms[0] = count + scan, ms[1] = emit */
int bmf_seam_stage_ms(bmf_ctx* ctx, float* ms /* [2] */);

#ifdef __cplusplus
}
#endif
#endif /* BMF_B200_H */

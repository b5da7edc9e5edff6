/*
 * bmf_oracle.h -- plain-C CPU restatement of BinaryMeshFitting's per-chunk extraction path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg
 * may call it, and only as the checker.  The product (binarymeshfitting_b200/) never links it.
 *
 * Pinning: stages 2-5 and the implicit samplers are pinned bit-exactly against the reference itself
 * (oracle/_ref/libbmf_ref.so, built from /root/reference) by tests/test_oracle_vs_ref.py and against
 * the committed golden vectors in tests/golden/ (generated from that library by
 * tests/golden/make_golden.py).  The noise stage is PARITY UNPINNED (see fastnoise_ref.h).
 * The QEF solver is pinned to the reference's qef_simd.h up to its one _mm_rsqrt_ps (approximate,
 * CPU-specific); this restatement uses an exact 1/sqrt there.
 */
#ifndef BMF_ORACLE_H
#define BMF_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* same numbering as include/bmf_b200.h */
enum {
	ORC_SPHERE = 0, ORC_TORUS_Z = 1, ORC_CUBOID = 2, ORC_PLANE_Y = 3, ORC_CSG = 4,
	ORC_TERRAIN2D = 10, ORC_TERRAIN2D_PERT = 11, ORC_TERRAIN3D = 12, ORC_TERRAIN3D_PERT = 13,
	ORC_HOST_DENSITY = 100
};
enum { ORC_CSG_UNION = 0, ORC_CSG_INTERSECT = 1, ORC_CSG_SUBTRACT = 2 };

typedef struct orc_sampler
{
	int32_t kind;
	float world_size;
	/* NoiseSamplerProperties (NoiseSampler.hpp:7-22; defaults WorldOctree.cpp:47-54) */
	float g_scale, height;
	int32_t octaves;
	float amp, frequency, gain;
	int32_t seed; /* FastNoiseSIMD seed (library default 1337) */
	/* CSG of two reference primitives (config 5; unpinned: the reference has no combinators) */
	int32_t csg_op, csg_kind_a, csg_kind_b;
	float csg_world_size_a, csg_world_size_b;
	float csg_offset_a[3], csg_offset_b[3];
} orc_sampler;

void orc_sampler_defaults(orc_sampler* s, int kind);

/* DMCChunk::label_grid geometry (DMCChunk.cpp:94-101) */
void orc_chunk_geometry(const float pos[3], float size, int dim, float overlap, float overlap_pos[3], float* delta);

float orc_implicit_value(int kind, float world_size, const float p[3]);
void orc_implicit_gradient(int kind, float world_size, const float p[3], float h, float out[3]);
/* Sampler::gradient for any sampler kind (CSG: differences of the combinator; noise kinds: the value callback is the constant 0) */
void orc_sampler_gradient(const orc_sampler* s, const float p[3], float h, float out[3]);
float orc_sampler_value(const orc_sampler* s, const float p[3]); /* implicit kinds + CSG only */

/* sampler.block(...) : density[(x*d + y)*d + z] at overlap_pos + (x,y,z)*delta */
int orc_sample_block(const orc_sampler* s, const float overlap_pos[3], float delta, int dim, float* density);

/* label_grid's sign pack (DMCChunk.cpp:118-162): returns contains_mesh */
int orc_label_grid(const float* density, int dim, uint32_t* bits);

/* label_edges' four mask passes (DMCChunk.cpp:184-438): masks viewed as uint8[d][d][d] */
void orc_cell_masks(const uint32_t* bits, int dim, uint8_t* masks);

typedef struct orc_mesh
{
	int32_t n_cells, n_verts, n_inds;
	uint32_t* dense_inds;   /* [d^3] cell id or 0xFFFFFFFF (IndexesBlock) */
	uint8_t* cell_masks;    /* [n_cells] */
	uint32_t* cell_grid;    /* [n_cells] linear grid index of the cell */
	float* pos;             /* [n_verts*3] grid units */
	uint8_t* boundary;      /* [n_verts] */
	uint8_t* valence;       /* [n_verts] init_valence */
	uint32_t* inds;         /* [n_inds] */
} orc_mesh;

/* label_edges' cell scan + polygonize (DMCChunk.cpp:440-498, 514-576, 593-689) */
void orc_extract(const float* density, const uint8_t* masks, int dim, orc_mesh* out);
void orc_mesh_free(orc_mesh* m);

/* MeshProcessor<N>::init + optimize_dual_grid(iters, pb) + optimize_primal_grid(false,false,pb)
 * (MeshProcessor.cpp:25-55, 98-128, 130-236, 238-306), N = 3 or 4; arrays updated in place.
 * normal may be null when smooth_normals == 0. */
void orc_smooth(float* pos, float* color, float* normal, const uint8_t* boundary, const uint8_t* valence, int n_verts,
                const uint32_t* inds, int n_inds, int prim_n, int iters, int process_boundary, int smooth_normals);

/* qef_solve_from_points_3d (qef_simd.h:550-579), exact 1/sqrt in givens_coeffs_sym */
float orc_qef_solve(const float* positions, const float* normals, int count, float solved[3]);

/* build-defined QEF placement after smoothing (UNPINNED policy; N = 3) */
void orc_qef_place(float* pos, const uint8_t* boundary, const uint8_t* valence, int n_verts, const uint32_t* inds, int n_inds, int process_boundary);
/* ColorMapper::generate_colors (ColorMapper.cpp:15-60): 4-octave simplex FBM at the vertex positions -> HSL -> RGB; pos, color [n][3] */
void orc_color_map(const float* pos, int n, float* color);
/* MeshProcessor<4>::init + collapse_bad_quads (MeshProcessor.cpp:308-396); pos / quads updated in place, returns bad_count */
int orc_collapse_bad_quads(float* pos, int n_verts, uint32_t* quads, int n_quads, uint8_t* destroyed, uint8_t* adj_next_out);
/* qef = 2: plane normals = normalised sampler gradient at the triangle centroids (MeshProcessor.cpp:224, commented out there) */
void orc_qef_place_gradient(float* pos, const uint8_t* boundary, const uint8_t* valence, int n_verts, const uint32_t* inds, int n_inds, int process_boundary,
                            const orc_sampler* s, const float overlap_pos[3], float delta, float h);

/* whole chunk: geometry -> sample -> bits -> masks -> extract -> smooth; returns contains_mesh.
 * density_io: if kind == ORC_HOST_DENSITY it is the input, else (if non-null) receives the samples. */
int orc_chunk(const orc_sampler* s, const float pos[3], float size, int dim, float overlap, int iters, int process_boundary,
              int smooth_normals, float* density_io, uint32_t* bits_out, uint8_t* masks_out, orc_mesh* mesh, float* color_out,
              float* normal_out);

/* batch of chunks, OpenMP over chunks (the cpu_baseline "port" leg); returns total verts, fills counts[n][2] */
int64_t orc_batch(const orc_sampler* s, const float* pos_size /* n x 4 */, int n, int dim, const float* overlaps, int iters,
                  int process_boundary, int threads, int32_t* counts);

/* quad emission (UNPINNED, build-defined: Nielson dual marching cubes on the sign field; see bmf_oracle.c).  Fills
 * n_cells / n_verts / n_inds (4 per quad), pos (grid units), boundary, valence, inds; the other members stay null. */
void orc_quads(const float* density, const uint32_t* bits, int dim, orc_mesh* out);

/* GLChunk::format_data(vertices, indexes, true, smooth_normals) (GLChunk.cpp:278-335): per quad corner p / n / c
 * ([n_inds][3] each); pinned against the compiled reference */
void orc_format_unwind(const float* pos, const float* normal, const float* color, const uint32_t* inds, int n_inds, int smooth_normals,
                       float* p_out, float* n_out, float* c_out);

/* seam pass between chunks (UNPINNED, build-defined: the reference's WorldStitcher is non-functional as committed).
 * bits / density: the chunk's sign words and density block as orc_label_grid / orc_sample_block produce them.
 * Returns the number of triangles (or -1 if the chunks are not aligned octree leaves); *tris_out = malloc'd
 * [n_tris][3][3] world-space positions (orc_free). */
typedef struct orc_seam_chunk
{
	float pos[3];
	float size;
	float overlap;
	const uint32_t* bits;
	const float* density;
} orc_seam_chunk;
int64_t orc_seam(const orc_seam_chunk* chunks, int n, int dim, const int32_t* group, int cross_group_only, float** tris_out);
void orc_free(void* p);

#ifdef __cplusplus
}
#endif
#endif

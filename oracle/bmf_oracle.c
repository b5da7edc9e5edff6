/*
 * bmf_oracle.c -- plain-C CPU restatement of BinaryMeshFitting's per-chunk extraction path.
 * TEST INFRASTRUCTURE ONLY (see bmf_oracle.h).  Build with -ffp-contract=off: every float
 * expression below is evaluated as written, one IEEE binary32 rounding per operation, like the
 * reference's MSVC /fp:strict build.  All file:line citations are into /root/reference/BinaryMeshFitting.
 */
#include "bmf_oracle.h"
#include "fastnoise_ref.h"
#include "mc_tables_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static const uint64_t TRI_PACK[256] = ORACLE_TRI_PACK_INIT;

/* ---- samplers ---------------------------------------------------------------------------------- */

void orc_sampler_defaults(orc_sampler* s, int kind)
{
	memset(s, 0, sizeof(*s));
	s->kind = kind;
	s->world_size = 256.0f;
	/* WorldOctree.cpp:47-54 */
	s->g_scale = 0.25f;
	s->height = 75.0f;
	s->octaves = 13;
	s->amp = 0.87f;
	s->frequency = 0.585f;
	s->gain = 0.488f;
	s->seed = 1337;
	s->csg_op = ORC_CSG_UNION;
	s->csg_kind_a = ORC_SPHERE;
	s->csg_kind_b = ORC_TORUS_Z;
	s->csg_world_size_a = s->csg_world_size_b = 256.0f;
}

/* ImplicitSampler.cpp:27-34 torus_z, :46-50 sphere, :52-60 cuboid, :62-65 plane_y  (density = -SDF) */
float orc_implicit_value(int kind, float ws, const float p[3])
{
	switch (kind)
	{
	case ORC_SPHERE:
	{
		float r = ws * 0.25f;
		float t0 = p[0] * p[0], t1 = p[1] * p[1], t2 = p[2] * p[2];
		float len = sqrtf(t0 + t1 + t2); /* glm::length = sqrt(dot), dot = x*x + y*y + z*z left to right */
		return -(len - r);
	}
	case ORC_TORUS_Z:
	{
		float r1 = ws / 4.0f;
		float r2 = ws / 10.0f;
		float q_x = fabsf(sqrtf(p[0] * p[0] + p[1] * p[1])) - r1;
		float len = sqrtf(q_x * q_x + p[2] * p[2]);
		return -(len - r2);
	}
	case ORC_CUBOID:
	{
		float r = ws / 8.0f;
		float dx = fabsf(p[0]) - r, dy = fabsf(p[1]) - r, dz = fabsf(p[2]) - r;
		float m = fmaxf(dx, fmaxf(dy, dz));
		float t0 = dx * dx, t1 = dy * dy, t2 = dz * dz;
		float len = sqrtf(t0 + t1 + t2);
		return -fminf(m, len);
	}
	default: /* ORC_PLANE_Y */
		return -p[1];
	}
}

/* ImplicitSampler.hpp:38-49: six evaluations, raw differences (not normalised, not divided by 2h) */
void orc_implicit_gradient(int kind, float ws, const float p[3], float h, float out[3])
{
	for (int a = 0; a < 3; a++)
	{
		float pp[3] = { p[0], p[1], p[2] }, pm[3] = { p[0], p[1], p[2] };
		pp[a] = p[a] + h;
		pm[a] = p[a] - h;
		out[a] = orc_implicit_value(kind, ws, pp) - orc_implicit_value(kind, ws, pm);
	}
}

/* implicit kinds + the build-defined CSG combinators (union = max, intersect = min, subtract = min(a,-b)
 * in the reference's positive-inside convention) */
float orc_sampler_value(const orc_sampler* s, const float p[3])
{
	if (s->kind == ORC_CSG)
	{
		float pa[3] = { p[0] - s->csg_offset_a[0], p[1] - s->csg_offset_a[1], p[2] - s->csg_offset_a[2] };
		float pb[3] = { p[0] - s->csg_offset_b[0], p[1] - s->csg_offset_b[1], p[2] - s->csg_offset_b[2] };
		float a = orc_implicit_value(s->csg_kind_a, s->csg_world_size_a, pa);
		float b = orc_implicit_value(s->csg_kind_b, s->csg_world_size_b, pb);
		switch (s->csg_op)
		{
		case ORC_CSG_UNION: return fmaxf(a, b);
		case ORC_CSG_INTERSECT: return fminf(a, b);
		default: return fminf(a, -b);
		}
	}
	return orc_implicit_value(s->kind, s->world_size, p);
}

/* Sampler::gradient of any sampler kind: create_sampler binds implicit_gradient to the sampler's VALUE callback
 * (ImplicitSampler.hpp:57, NoiseSampler.hpp:76-136).  For the noise samplers that callback is NoiseSamplers::noise3d, which
 * returns 0 (NoiseSampler.cpp:99-102), so their gradient is (0-0, 0-0, 0-0); the CSG value is the build-defined combinator. */
void orc_sampler_gradient(const orc_sampler* s, const float p[3], float h, float out[3])
{
	const int analytic = s->kind == ORC_CSG || (s->kind >= 0 && s->kind <= 3);
	for (int a = 0; a < 3; a++)
	{
		float pp[3] = { p[0], p[1], p[2] }, pm[3] = { p[0], p[1], p[2] };
		pp[a] = p[a] + h;
		pm[a] = p[a] - h;
		const float vp = analytic ? orc_sampler_value(s, pp) : 0.0f, vm = analytic ? orc_sampler_value(s, pm) : 0.0f;
		out[a] = vp - vm;
	}
}

/* DMCChunk.cpp:94-101 */
void orc_chunk_geometry(const float pos[3], float size, int dim, float overlap, float overlap_pos[3], float* delta)
{
	*delta = size * (1.0f + overlap * 2.0f) / (float)(dim - 1);
	float so = size * overlap;
	overlap_pos[0] = pos[0] - so;
	overlap_pos[1] = pos[1] - so;
	overlap_pos[2] = pos[2] - so;
}

/* FastNoiseSIMD state after the setter sequence of each *_block function; a fresh per-thread sampler
 * object (NoiseSampler.hpp:78-139) carries library defaults for everything a block does not set. */
static void noise_state_for(const orc_sampler* s, fnr_state* st)
{
	fnr_init(st, s->seed);
	switch (s->kind)
	{
	case ORC_TERRAIN2D: /* NoiseSampler.cpp:120-125 */
		st->noise_type = FNR_VALUE_FRACTAL;
		fnr_set_fractal_octaves(st, 12);
		fnr_set_fractal_gain(st, 0.5f);
		st->lacunarity = 2.0f;
		st->fractal_type = FNR_FBM;
		break;
	case ORC_TERRAIN2D_PERT: /* NoiseSampler.cpp:160-167 */
		st->noise_type = FNR_VALUE_FRACTAL;
		st->perturb_type = FNR_PERTURB_GRADIENT_FRACTAL;
		fnr_set_perturb_octaves(st, s->octaves);
		fnr_set_perturb_amp(st, s->amp);
		st->perturb_frequency = s->frequency;
		fnr_set_perturb_gain(st, s->gain);
		st->fractal_type = FNR_FBM;
		break;
	case ORC_TERRAIN3D: /* NoiseSampler.cpp:203-206 */
		st->noise_type = FNR_VALUE_FRACTAL;
		fnr_set_fractal_octaves(st, 4);
		st->fractal_type = FNR_RIGIDMULTI;
		break;
	default: /* ORC_TERRAIN3D_PERT, NoiseSampler.cpp:236-242 */
		st->noise_type = FNR_SIMPLEX_FRACTAL;
		st->perturb_type = FNR_PERTURB_GRADIENT_FRACTAL;
		fnr_set_fractal_octaves(st, 8);
		fnr_set_perturb_amp(st, 1.0f);
		st->perturb_frequency = 0.05f;
		st->fractal_type = FNR_RIGIDMULTI;
		break;
	}
}

int orc_sample_block(const orc_sampler* s, const float op[3], float scale, int dim, float* density)
{
	const int d = dim;
	if (s->kind <= ORC_CSG)
	{
		/* implicit_block (ImplicitSampler.hpp:14-36): coordinate = p + (float)i * scale */
		for (int x = 0; x < d; x++)
		{
			float px = op[0] + (float)x * scale;
			for (int y = 0; y < d; y++)
			{
				float py = op[1] + (float)y * scale;
				for (int z = 0; z < d; z++)
				{
					float p[3] = { px, py, op[2] + (float)z * scale };
					density[((size_t)x * d + y) * d + z] = orc_sampler_value(s, p);
				}
			}
		}
		return 0;
	}
	if (s->kind < ORC_TERRAIN2D || s->kind > ORC_TERRAIN3D_PERT)
		return -1;

	fnr_state st;
	noise_state_for(s, &st);
	const int two_d = (s->kind == ORC_TERRAIN2D || s->kind == ORC_TERRAIN2D_PERT);
	float g, nm;
	switch (s->kind)
	{
	case ORC_TERRAIN2D: g = 1.0f; nm = 64.0f; break;             /* :115-116 */
	case ORC_TERRAIN2D_PERT: g = s->g_scale; nm = s->height; break; /* :150-151 */
	case ORC_TERRAIN3D: g = 0.15f; nm = 1.0f; break;            /* :198 */
	default: g = 0.15f; nm = 48.0f; break;                      /* :231-232 */
	}
	/* NOISE_BLOCK (NoiseSampler.cpp:8-35): d = (float)i * (scale*g) + p*g ; the y coordinate of the
	 * 2-D variants is (float)0 * s + 0 = 0 */
	const float sg = scale * g;
	const float pxg = op[0] * g, pyg = op[1] * g, pzg = op[2] * g;
	const size_t count = two_d ? (size_t)d * d : (size_t)d * d * d;
	float* xs = (float*)malloc(sizeof(float) * count * 4);
	if (!xs)
		return -2;
	float* ys = xs + count;
	float* zs = ys + count;
	float* noise = zs + count;
	size_t idx = 0;
	for (int ix = 0; ix < d; ix++)
	{
		float dx = (float)ix * sg + pxg;
		for (int iy = 0; iy < (two_d ? 1 : d); iy++)
		{
			float dy = two_d ? ((float)iy * sg + 0.0f) : ((float)iy * sg + pyg);
			for (int iz = 0; iz < d; iz++)
			{
				xs[idx] = dx;
				ys[idx] = dy;
				zs[idx] = (float)iz * sg + pzg;
				idx++;
			}
		}
	}
	fnr_fill_noise_set(&st, noise, xs, ys, zs, (int)count, 0.0f, 0.0f, 0.0f);

	for (int ix = 0; ix < d; ix++)
		for (int iy = 0; iy < d; iy++)
		{
			float dy;
			if (s->kind == ORC_TERRAIN3D)
				dy = ((float)iy * scale + op[1]) * g * 0.5f; /* :216, ym = 0.5 */
			else
				dy = ((float)iy * scale + op[1]) * g;         /* :136, :180, :250 */
			for (int iz = 0; iz < d; iz++)
			{
				float n = two_d ? noise[(size_t)ix * d + iz] : noise[((size_t)ix * d + iy) * d + iz];
				float v = (s->kind == ORC_TERRAIN3D) ? (-dy - n) : (-dy - n * nm);
				density[((size_t)ix * d + iy) * d + iz] = v;
			}
		}
	free(xs);
	return 0;
}

/* ---- stage 2: sign pack (DMCChunk.cpp:118-162) --------------------------------------------------- */

int orc_label_grid(const float* density, int dim, uint32_t* bits)
{
	const int d = dim, zc = (d + 31) / 32;
	int mesh = 0, negative = 0, positive = 0;
	for (int x = 0; x < d; x++)
		for (int y = 0; y < d; y++)
			for (int zb = 0; zb < zc; zb++)
			{
				const float* row = density + ((size_t)x * d + y) * d + zb * 32;
				int zmax = d - zb * 32;
				if (zmax > 32) zmax = 32;
				uint32_t m = 0;
				for (int z = 0; z < zmax; z++)
					if (row[z] < 0.0f)
						m |= 1u << z;
				bits[((size_t)x * d + y) * zc + zb] = m;
				if (m != 0)
				{
					if (m != 0xFFFFFFFFu) mesh = 1;
					else if (!mesh) negative = 1;
				}
				else if (!mesh)
					positive = 1;
			}
	return mesh ? 1 : (negative && positive);
}

/* ---- stage 3: cell masks (closed form of DMCChunk.cpp:184-438) ----------------------- */

static inline int bit_at(const uint32_t* bits, int d, int zc, int x, int y, int z)
{
	if (x >= d || y >= d || z >= d) return 0;
	return (bits[((size_t)x * d + y) * zc + (z >> 5)] >> (z & 31)) & 1;
}

void orc_cell_masks(const uint32_t* bits, int dim, uint8_t* masks)
{
	const int d = dim, zc = (d + 31) / 32;
	for (int x = 0; x < d; x++)
		for (int y = 0; y < d; y++)
			for (int z = 0; z < d; z++)
			{
				unsigned m = 0;
				for (int i = 0; i < 8; i++)
					m |= (unsigned)bit_at(bits, d, zc, x + (i >> 2), y + ((i >> 1) & 1), z + (i & 1)) << i;
				masks[((size_t)x * d + y) * d + z] = (uint8_t)m;
			}
}

/* ---- stage 4: cell scan + polygonize ------------------------------------------------------------- */

void orc_mesh_free(orc_mesh* m)
{
	free(m->dense_inds); free(m->cell_masks); free(m->cell_grid); free(m->pos);
	free(m->boundary); free(m->valence); free(m->inds);
	memset(m, 0, sizeof(*m));
}

/* _get_intersection (DMCChunk.cpp:657-662) with isolevel 0; calculate_isovertex (:664-674) */
static void iso_vertex(const float* D, int d, int x0, int y0, int z0, int x1, int y1, int z1, float* p, uint8_t* boundary)
{
	float s0 = D[((size_t)x0 * d + y0) * d + z0];
	float s1 = D[((size_t)x1 * d + y1) * d + z1];
	float mu = (0.0f - s0) / (s1 - s0);
	float p0[3] = { (float)x0, (float)y0, (float)z0 };
	float p1[3] = { (float)x1, (float)y1, (float)z1 };
	for (int a = 0; a < 3; a++)
	{
		float delta = (p1[a] - p0[a]) * mu;
		p[a] = delta + p0[a];
	}
	*boundary = (x0 == 0 || y0 == 0 || z0 == 0 || x0 == d - 1 || y0 == d - 1 || z0 == d - 1 || x1 == d - 1 || y1 == d - 1 || z1 == d - 1) ? 1 : 0;
}

void orc_extract(const float* D, const uint8_t* masks, int dim, orc_mesh* out)
{
	const int d = dim;
	const size_t n = (size_t)d * d * d;
	memset(out, 0, sizeof(*out));
	/* pass 1: count (serial x->y->z scan order, DMCChunk.cpp:449-498) */
	size_t nc = 0, nv = 0;
	for (size_t i = 0; i < n; i++)
	{
		unsigned m = masks[i];
		if (m == 0 || m == 255) continue;
		int x = (int)(i / ((size_t)d * d)), y = (int)(i / d % d), z = (int)(i % d);
		nc++;
		nv += (((m ^ (m >> 4)) & 1) && x + 1 < d) + (((m ^ (m >> 2)) & 1) && y + 1 < d) + (((m ^ (m >> 1)) & 1) && z + 1 < d);
	}
	out->dense_inds = (uint32_t*)malloc(sizeof(uint32_t) * n);
	out->cell_masks = (uint8_t*)malloc(nc ? nc : 1);
	out->cell_grid = (uint32_t*)malloc(sizeof(uint32_t) * (nc ? nc : 1));
	out->pos = (float*)malloc(sizeof(float) * 3 * (nv ? nv : 1));
	out->boundary = (uint8_t*)malloc(nv ? nv : 1);
	out->valence = (uint8_t*)calloc(nv ? nv : 1, 1);
	uint32_t* vx = (uint32_t*)malloc(sizeof(uint32_t) * 3 * (nc ? nc : 1)); /* per cell: ids of its X/Y/Z edge vertices */

	size_t c = 0, v = 0;
	for (size_t i = 0; i < n; i++)
	{
		unsigned m = masks[i];
		if (m == 0 || m == 255)
		{
			out->dense_inds[i] = 0xFFFFFFFFu;
			continue;
		}
		int x = (int)(i / ((size_t)d * d)), y = (int)(i / d % d), z = (int)(i % d);
		out->dense_inds[i] = (uint32_t)c;
		out->cell_masks[c] = (uint8_t)m;
		out->cell_grid[c] = (uint32_t)i;
		/* calculate_cell (DMCChunk.cpp:593-655): X, Y, Z edge in that order */
		vx[3 * c + 0] = vx[3 * c + 1] = vx[3 * c + 2] = 0xFFFFFFFFu;
		if (((m ^ (m >> 4)) & 1) && x + 1 < d)
		{
			iso_vertex(D, d, x, y, z, x + 1, y, z, out->pos + 3 * v, out->boundary + v);
			vx[3 * c + 0] = (uint32_t)v++;
		}
		if (((m ^ (m >> 2)) & 1) && y + 1 < d)
		{
			iso_vertex(D, d, x, y, z, x, y + 1, z, out->pos + 3 * v, out->boundary + v);
			vx[3 * c + 1] = (uint32_t)v++;
		}
		if (((m ^ (m >> 1)) & 1) && z + 1 < d)
		{
			iso_vertex(D, d, x, y, z, x, y, z + 1, out->pos + 3 * v, out->boundary + v);
			vx[3 * c + 2] = (uint32_t)v++;
		}
		c++;
	}
	out->n_cells = (int32_t)nc;
	out->n_verts = (int32_t)nv;

	/* polygonize (DMCChunk.cpp:514-576): cells in cell order with x,y,z < d-1 */
	size_t ni = 0;
	for (size_t k = 0; k < nc; k++)
	{
		size_t i = out->cell_grid[k];
		int x = (int)(i / ((size_t)d * d)), y = (int)(i / d % d), z = (int)(i % d);
		if (x >= d - 1 || y >= d - 1 || z >= d - 1) continue;
		ni += (size_t)(TRI_PACK[out->cell_masks[k]] >> 60);
	}
	out->inds = (uint32_t*)malloc(sizeof(uint32_t) * (ni ? ni : 1));
	out->n_inds = (int32_t)ni;
	ni = 0;
	/* EDGE_V (DMCChunk.cpp:32, 543-565): neighbour cell offsets (dx,dy,dz) and which of its edges */
	static const int EOFF[12][4] = {
		{0,0,0,0},{0,0,1,0},{0,1,0,0},{0,1,1,0},
		{0,0,0,1},{0,0,1,1},{1,0,0,1},{1,0,1,1},
		{0,0,0,2},{0,1,0,2},{1,0,0,2},{1,1,0,2} };
	for (size_t k = 0; k < nc; k++)
	{
		size_t i = out->cell_grid[k];
		int x = (int)(i / ((size_t)d * d)), y = (int)(i / d % d), z = (int)(i % d);
		if (x >= d - 1 || y >= d - 1 || z >= d - 1) continue;
		uint64_t tp = TRI_PACK[out->cell_masks[k]];
		int cnt = (int)(tp >> 60);
		for (int t = 0; t < cnt; t++)
		{
			int e = (int)((tp >> (4 * t)) & 15);
			size_t nb = ((size_t)(x + EOFF[e][0]) * d + (y + EOFF[e][1])) * d + (z + EOFF[e][2]);
			uint32_t vid = vx[3 * (size_t)out->dense_inds[nb] + EOFF[e][3]];
			out->inds[ni++] = vid;
			out->valence[vid]++;
		}
	}
	free(vx);
}

/* ---- stage 5: dual/primal smoothing -------------------------------------------------- */

static inline void v3_normalize(float* v)
{
	float t0 = v[0] * v[0], t1 = v[1] * v[1], t2 = v[2] * v[2];
	float inv = 1.0f / sqrtf(t0 + t1 + t2);
	v[0] = v[0] * inv; v[1] = v[1] * inv; v[2] = v[2] * inv;
}

static inline void v3_cross(const float* x, const float* y, float* o)
{
	o[0] = x[1] * y[2] - y[1] * x[2];
	o[1] = x[2] * y[0] - y[2] * x[0];
	o[2] = x[0] * y[1] - y[0] * x[1];
}

typedef struct
{
	int N, nv, np, smooth;
	const uint32_t* inds;
	const uint8_t* boundary;
	uint32_t *adj_off, *adj;
	uint8_t* adj_cnt;
	float *dp, *dc, *dn;
	float *p, *c, *nrm;
} smooth_ctx;

/* optimize_primal_grid (MeshProcessor.cpp:238-306) */
static void primal_step(smooth_ctx* s, int set_colors, int pb)
{
	for (int i = 0; i < s->nv; i++)
	{
		if (s->adj_cnt[i] == 0 || (!pb && s->boundary[i])) continue;
		float p[3] = { 0, 0, 0 }, n[3] = { 0, 0, 0 }, c[3] = { 0, 0, 0 };
		int count = 0;
		const uint32_t* adj = s->adj + s->adj_off[i];
		for (int k = 0; k < s->adj_cnt[i]; k++)
		{
			uint32_t t = adj[k];
			count++;
			for (int a = 0; a < 3; a++)
			{
				p[a] += s->dp[3 * (size_t)t + a];
				c[a] += s->dc[3 * (size_t)t + a];
				if (s->smooth) n[a] += s->dn[3 * (size_t)t + a];
			}
		}
		float fc = (float)count;
		for (int a = 0; a < 3; a++)
		{
			p[a] /= fc;
			c[a] /= fc;
			if (s->smooth) n[a] /= fc;
		}
		if (set_colors) v3_normalize(n);
		for (int a = 0; a < 3; a++)
		{
			s->p[3 * (size_t)i + a] = p[a];
			s->c[3 * (size_t)i + a] = c[a];
		}
		if (n[1] != 0 && s->nrm)
			for (int a = 0; a < 3; a++) s->nrm[3 * (size_t)i + a] = n[a];
	}
}

void orc_smooth(float* pos, float* color, float* normal, const uint8_t* boundary, const uint8_t* valence, int n_verts,
                const uint32_t* inds, int n_inds, int prim_n, int iters, int pb, int smooth)
{
	const int N = prim_n;
	if (n_verts == 0 || n_inds < N || iters <= 0) return;
	smooth_ctx s;
	memset(&s, 0, sizeof(s));
	s.N = N; s.nv = n_verts; s.np = n_inds / N; s.smooth = smooth; s.inds = inds; s.boundary = boundary;
	s.p = pos; s.c = color; s.nrm = normal;
	/* init (MeshProcessor.cpp:25-55): adj_offset = exclusive prefix of init_valence */
	s.adj_off = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)n_verts);
	s.adj_cnt = (uint8_t*)calloc((size_t)n_verts, 1);
	uint32_t a_count = 0;
	for (int i = 0; i < n_verts; i++)
	{
		s.adj_off[i] = a_count;
		a_count += valence[i];
	}
	if (!a_count)
	{
		free(s.adj_off); free(s.adj_cnt);
		return;
	}
	s.adj = (uint32_t*)malloc(sizeof(uint32_t) * a_count);
	/* init_primitives (:98-128): adj filled in (primitive, corner) order */
	for (int t = 0; t < s.np; t++)
		for (int k = 0; k < N; k++)
		{
			uint32_t v = inds[(size_t)t * N + k];
			s.adj[s.adj_off[v] + s.adj_cnt[v]++] = (uint32_t)t;
		}
	s.dp = (float*)calloc((size_t)s.np * 3, sizeof(float));
	s.dc = (float*)calloc((size_t)s.np * 3, sizeof(float));
	s.dn = (float*)calloc((size_t)s.np * 3, sizeof(float));

	/* optimize_dual_grid (:130-236) */
	const int hard_norm_max = 10;
	int max_norms = (iters / 2 - 3 < hard_norm_max ? iters / 2 - 3 : hard_norm_max);
	for (int m = 0; m < iters; m++)
	{
		for (int t = 0; t < s.np; t++)
		{
			const uint32_t* tv = inds + (size_t)t * N;
			float sp[3] = { 0, 0, 0 }, sc[3] = { 0, 0, 0 };
			for (int k = 0; k < N; k++)
				for (int a = 0; a < 3; a++)
				{
					sp[a] += pos[3 * (size_t)tv[k] + a];
					sc[a] += color[3 * (size_t)tv[k] + a];
				}
			for (int a = 0; a < 3; a++)
			{
				s.dp[3 * (size_t)t + a] = sp[a] / (float)N;
				s.dc[3 * (size_t)t + a] = sc[a] / (float)N;
			}
			if (smooth)
			{
				float* dn = s.dn + 3 * (size_t)t;
				if (m == 0 || m < max_norms || m < 3)
				{
					const float* p0 = pos + 3 * (size_t)tv[0];
					if (N == 3)
					{
						/* :165-185 (the duplicate-vertex branch leaves a,b uninitialised in the reference;
						 * MC output never has duplicates) */
						const float *p1 = pos + 3 * (size_t)tv[1], *p2 = pos + 3 * (size_t)tv[2];
						float a[3] = { p0[0] - p1[0], p0[1] - p1[1], p0[2] - p1[2] };
						float b[3] = { p0[0] - p2[0], p0[1] - p2[1], p0[2] - p2[2] };
						v3_normalize(a); v3_normalize(b);
						float cr[3];
						v3_cross(a, b, cr);
						dn[0] = -cr[0]; dn[1] = -cr[1]; dn[2] = -cr[2];
					}
					else
					{
						/* :186-207 quad: average of two triangle normals with NaN guards */
						const float *p1 = pos + 3 * (size_t)tv[1], *p2 = pos + 3 * (size_t)tv[2], *p3 = pos + 3 * (size_t)tv[3];
						float a[3] = { p0[0] - p1[0], p0[1] - p1[1], p0[2] - p1[2] };
						float b[3] = { p0[0] - p2[0], p0[1] - p2[1], p0[2] - p2[2] };
						float n1[3], n2[3];
						v3_normalize(a); v3_normalize(b);
						v3_cross(a, b, n1);
						float a2[3] = { p2[0] - p3[0], p2[1] - p3[1], p2[2] - p3[2] };
						float b2[3] = { p2[0] - p0[0], p2[1] - p0[1], p2[2] - p0[2] };
						v3_normalize(a2); v3_normalize(b2);
						v3_cross(a2, b2, n2);
						if (isnan(n1[0]))
						{
							memcpy(n1, n2, sizeof(n1));
							if (isnan(n1[0])) { n1[0] = 0; n1[1] = 1; n1[2] = 0; }
						}
						if (isnan(n2[0])) memcpy(n2, n1, sizeof(n2));
						float h[3] = { (n1[0] + n2[0]) * 0.5f, (n1[1] + n2[1]) * 0.5f, (n1[2] + n2[2]) * 0.5f };
						v3_normalize(h);
						dn[0] = -h[0]; dn[1] = -h[1]; dn[2] = -h[2];
					}
				}
				else
				{
					float sn[3] = { 0, 0, 0 };
					for (int k = 0; k < N; k++)
						for (int a = 0; a < 3; a++) sn[a] += normal[3 * (size_t)tv[k] + a];
					dn[0] = sn[0]; dn[1] = sn[1]; dn[2] = sn[2]; /* not averaged (:211-213) */
				}
			}
		}
		if (m < iters - 1)
			primal_step(&s, (m == 3) || (m == 0 && iters <= 3), pb);
	}
	/* the driver's extra call (ChunkGenerator.cpp:120) */
	primal_step(&s, 0, pb);
	free(s.adj_off); free(s.adj_cnt); free(s.adj); free(s.dp); free(s.dc); free(s.dn);
}

/* ---- QEF (qef_simd.h:411-579; scalar form SURVEY C.4) --------------------------------------------- */

static inline float dot4(const float* a, const float* b)
{
	/* vec4_dot shuffle order (qef_simd.h:100-109) */
	return (a[0] * b[0] + a[1] * b[1]) + (a[3] * b[3] + a[2] * b[2]);
}

static inline void vmul44(const float* a, float B[4][4], float* r)
{
	/* vec4_mul_m4x4 (:113-122) */
	for (int j = 0; j < 4; j++)
		r[j] = ((a[0] * B[0][j] + a[1] * B[1][j]) + a[2] * B[2][j]) + a[3] * B[3][j];
}

float orc_qef_solve(const float* positions, const float* normals, int count, float solved[3])
{
	if (count < 2 || count > 12)
	{
		solved[0] = solved[1] = solved[2] = 0.0f;
		return 0.0f;
	}
	float ATA[4][4], ATb[4] = { 0, 0, 0, 0 }, acc[4] = { 0, 0, 0, 0 };
	memset(ATA, 0, sizeof(ATA));
	for (int i = 0; i < count; i++)
	{
		/* qef_simd_add (:411-430) */
		float p[4] = { positions[3 * i], positions[3 * i + 1], positions[3 * i + 2], 1.0f };
		float n[4] = { normals[3 * i], normals[3 * i + 1], normals[3 * i + 2], 0.0f };
		for (int r = 0; r < 3; r++)
			for (int j = 0; j < 4; j++) ATA[r][j] += n[r] * n[j];
		float d = dot4(p, n);
		float dv[4] = { d, d, d, 0.0f };
		for (int j = 0; j < 4; j++)
		{
			ATb[j] += dv[j] * n[j];
			acc[j] += p[j];
		}
	}
	/* qef_simd_solve (:444-461) */
	float mp[4], b[4], tmp[4];
	for (int j = 0; j < 4; j++) mp[j] = acc[j] / acc[3];
	vmul44(mp, ATA, tmp);
	for (int j = 0; j < 4; j++) b[j] = ATb[j] - tmp[j];

	/* svd_solve_sym (:304-342) */
	float A[4][4], V[4][4];
	memcpy(A, ATA, sizeof(A));
	memset(V, 0, sizeof(V));
	V[0][0] = V[1][1] = V[2][2] = 1.0f;
	static const int PAIRS[3][2] = { {0, 1}, {0, 2}, {1, 2} };
	for (int sweep = 0; sweep < 5; sweep++)
		for (int pi = 0; pi < 3; pi++)
		{
			int a = PAIRS[pi][0], q = PAIRS[pi][1];
			if (A[a][q] == 0.0f) continue;
			/* givens_coeffs_sym (:150-211) */
			float pp = A[a][a], pq = A[a][q], qq = A[q][q];
			float tau = (qq - pp) / (pq * 2.0f);
			float stt = sqrtf(tau * tau + 1.0f);
			float tn = 1.0f / ((tau >= 0.0f) ? (tau + stt) : (tau - stt));
			float c = 1.0f / sqrtf(1.0f + tn * tn); /* the reference uses _mm_rsqrt_ps (12-bit approximation) here */
			float s = tn * c;
			if (pq == 0.0f) { c = 1.0f; s = 0.0f; }
			/* rotateq_xy (:215-260) */
			float cc = c * c, ss = s * s;
			float mx = ((2.0f * c) * s) * pq;
			float xx = (cc * pp - mx) + ss * qq;
			float yy = (ss * pp + mx) + cc * qq;
			A[a][a] = xx;
			A[q][q] = yy;
			/* rotate_xy (:264-300): three rows of V plus the two other off-diagonals */
			float* o1 = &A[0][3 - q];
			float* o2 = &A[1 - a][2];
			float u[4] = { V[0][a], V[1][a], V[2][a], *o1 };
			float w[4] = { V[0][q], V[1][q], V[2][q], *o2 };
			float xr[4], yr[4];
			for (int k = 0; k < 4; k++)
			{
				xr[k] = c * u[k] - s * w[k];
				yr[k] = s * u[k] + c * w[k];
			}
			V[0][a] = xr[0]; V[1][a] = xr[1]; V[2][a] = xr[2]; *o1 = xr[3];
			V[0][q] = yr[0]; V[1][q] = yr[1]; V[2][q] = yr[2]; *o2 = yr[3];
			A[a][q] = 0.0f;
		}
	/* svd_invdet (:346-357) + svd_pseudoinverse (:361-387) */
	float sigma[4] = { A[0][0], A[1][1], A[2][2], 0.0f }, inv[4];
	for (int j = 0; j < 4; j++)
	{
		float one_over = 1.0f / sigma[j];
		float mn = fminf(fabsf(sigma[j]), fabsf(one_over));
		inv[j] = (mn >= 0.001f) ? one_over : 0.0f;
	}
	float M[3][4], P[4][4];
	for (int r = 0; r < 3; r++)
		for (int j = 0; j < 4; j++) M[r][j] = V[r][j] * inv[j];
	memset(P, 0, sizeof(P));
	for (int i = 0; i < 3; i++)
		for (int j = 0; j < 3; j++) P[i][j] = dot4(M[j], V[i]);
	float x[4];
	vmul44(b, P, x);
	/* qef_simd_calc_error (:434-440) */
	vmul44(x, ATA, tmp);
	float e[4];
	for (int j = 0; j < 4; j++) e[j] = ATb[j] - tmp[j];
	float err = dot4(e, e);
	solved[0] = x[0] + mp[0];
	solved[1] = x[1] + mp[1];
	solved[2] = x[2] + mp[2];
	return err;
}

/* Build-defined QEF placement (bmf_params.qef, config 5; UNPINNED: the reference never calls its solver,
 * MeshProcessor.cpp:239).  Restated here only so the GPU policy has a CPU twin: after smoothing, every processed
 * vertex with >= 2 adjacent triangles is moved to the QEF minimiser of the planes (centroid, face normal) of its
 * first <= 12 adjacent triangles (ascending triangle id), clamped to the bounding box of those centroids. */
static void qef_place_impl(float* pos, const uint8_t* boundary, const uint8_t* valence, int n_verts, const uint32_t* inds, int n_inds, int pb,
                           const orc_sampler* gs, const float* gop, float gdelta, float gh);

void orc_qef_place(float* pos, const uint8_t* boundary, const uint8_t* valence, int n_verts, const uint32_t* inds, int n_inds, int pb)
{
	qef_place_impl(pos, boundary, valence, n_verts, inds, n_inds, pb, NULL, NULL, 0.0f, 0.0f);
}

/* bmf_params.qef = 2 (config 5 "with gradients"): the same placement, but the plane normal of a triangle is the sampler's
 * gradient at the triangle's centroid, normalised -- what the reference's author left commented out in optimize_dual_grid,
 * `t.dual_n = normalize(sampler.gradient(sampler.world_size, t.dual_p, ...))` (MeshProcessor.cpp:224).  The centroid is in grid
 * units; its world position is overlap_pos + p * scale (DMCChunk.cpp:94-98).  UNPINNED policy, pinned gradient. */
void orc_qef_place_gradient(float* pos, const uint8_t* boundary, const uint8_t* valence, int n_verts, const uint32_t* inds, int n_inds, int pb,
                            const orc_sampler* s, const float overlap_pos[3], float delta, float h)
{
	qef_place_impl(pos, boundary, valence, n_verts, inds, n_inds, pb, s, overlap_pos, delta, h);
}

static void qef_place_impl(float* pos, const uint8_t* boundary, const uint8_t* valence, int n_verts, const uint32_t* inds, int n_inds, int pb,
                           const orc_sampler* gs, const float* gop, float gdelta, float gh)
{
	const int np = n_inds / 3;
	if (n_verts == 0 || np == 0) return;
	uint32_t* off = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)n_verts);
	uint8_t* cnt = (uint8_t*)calloc((size_t)n_verts, 1);
	uint32_t a = 0;
	for (int i = 0; i < n_verts; i++) { off[i] = a; a += valence[i]; }
	uint32_t* adj = (uint32_t*)malloc(sizeof(uint32_t) * (a ? a : 1));
	for (int t = 0; t < np; t++)
		for (int k = 0; k < 3; k++) { uint32_t v = inds[3 * (size_t)t + k]; adj[off[v] + cnt[v]++] = (uint32_t)t; }
	float* dp = (float*)malloc(sizeof(float) * 3 * (size_t)np);
	float* dn = (float*)malloc(sizeof(float) * 3 * (size_t)np);
	for (int t = 0; t < np; t++)
	{
		const float *p0 = pos + 3 * (size_t)inds[3 * (size_t)t], *p1 = pos + 3 * (size_t)inds[3 * (size_t)t + 1], *p2 = pos + 3 * (size_t)inds[3 * (size_t)t + 2];
		for (int c = 0; c < 3; c++) dp[3 * (size_t)t + c] = (((0.0f + p0[c]) + p1[c]) + p2[c]) / 3.0f;
		if (gs)
		{
			float wp[3], g[3];
			for (int c = 0; c < 3; c++) wp[c] = gop[c] + dp[3 * (size_t)t + c] * gdelta;
			orc_sampler_gradient(gs, wp, gh, g);
			v3_normalize(g);
			dn[3 * (size_t)t] = g[0]; dn[3 * (size_t)t + 1] = g[1]; dn[3 * (size_t)t + 2] = g[2];
			continue;
		}
		float u[3] = { p0[0] - p1[0], p0[1] - p1[1], p0[2] - p1[2] }, w[3] = { p0[0] - p2[0], p0[1] - p2[1], p0[2] - p2[2] }, cr[3];
		v3_normalize(u); v3_normalize(w);
		v3_cross(u, w, cr);
		dn[3 * (size_t)t] = -cr[0]; dn[3 * (size_t)t + 1] = -cr[1]; dn[3 * (size_t)t + 2] = -cr[2];
	}
	float* out = (float*)malloc(sizeof(float) * 3 * (size_t)n_verts);
	memcpy(out, pos, sizeof(float) * 3 * (size_t)n_verts);
	for (int v = 0; v < n_verts; v++)
	{
		int c = valence[v];
		if (c < 2 || (!pb && boundary[v])) continue;
		if (c > 12) c = 12;
		float P[36], Nn[36], lo[3] = { 3.0e38f, 3.0e38f, 3.0e38f }, hi[3] = { -3.0e38f, -3.0e38f, -3.0e38f }, x[3];
		for (int k = 0; k < c; k++)
		{
			uint32_t t = adj[off[v] + k];
			for (int q = 0; q < 3; q++)
			{
				P[3 * k + q] = dp[3 * (size_t)t + q];
				Nn[3 * k + q] = dn[3 * (size_t)t + q];
				lo[q] = fminf(lo[q], P[3 * k + q]);
				hi[q] = fmaxf(hi[q], P[3 * k + q]);
			}
		}
		orc_qef_solve(P, Nn, c, x);
		if (isnan(x[0]) || isnan(x[1]) || isnan(x[2])) continue;
		for (int q = 0; q < 3; q++) out[3 * (size_t)v + q] = fminf(fmaxf(x[q], lo[q]), hi[q]);
	}
	memcpy(pos, out, sizeof(float) * 3 * (size_t)n_verts);
	free(off); free(cnt); free(adj); free(dp); free(dn); free(out);
}

/* ---- ColorMapper::generate_colors (ColorMapper.cpp:15-60) -------------------------------------------- */
/* hsl_to_rgb (ColorMapper.cpp:62-121) */
static void hsl_to_rgb(float h, float s, float v, float out[3])
{
	float r = 0, g = 0, b = 0;
	if (s <= 0.0f) { out[0] = v; out[1] = v; out[2] = v; return; }
	float hh = h;
	hh = fmodf(fabsf(hh), 360.0f);
	hh /= 60.0f;
	int i = (int)hh;
	float ff = hh - i;
	float p = v * (1.0f - s);
	float q = v * (1.0f - (s * ff));
	float t = v * (1.0f - (s * (1.0f - ff)));
	switch (i)
	{
	case 0: r = v; g = t; b = p; break;
	case 1: r = q; g = v; b = p; break;
	case 2: r = p; g = v; b = t; break;
	case 3: r = p; g = q; b = v; break;
	case 4: r = t; g = p; b = v; break;
	default: r = v; g = p; b = q; break;
	}
	out[0] = r; out[1] = g; out[2] = b;
}

/* get_noise (:27-48): a fresh FastNoiseSIMD (seed 1337), SimplexFractal, 4 octaves, FBM, vector set = vertex positions * 1.0;
 * map_noise (:50-60): n = noise * 4; colour = hsl_to_rgb((n + 1) * 0.5 * 360, 0.72, 1).  pos, color: [n][3] */
void orc_color_map(const float* pos, int n, float* color)
{
	if (n <= 0) return;
	fnr_state st;
	fnr_init(&st, 1337);
	st.noise_type = FNR_SIMPLEX_FRACTAL;
	fnr_set_fractal_octaves(&st, 4);
	st.fractal_type = FNR_FBM;
	const float scale = 1.0f;
	float* xs = (float*)malloc(sizeof(float) * 4 * (size_t)n);
	float *ys = xs + n, *zs = ys + n, *noise = zs + n;
	for (int i = 0; i < n; i++)
	{
		xs[i] = pos[3 * (size_t)i] * scale;
		ys[i] = pos[3 * (size_t)i + 1] * scale;
		zs[i] = pos[3 * (size_t)i + 2] * scale;
	}
	fnr_fill_noise_set(&st, noise, xs, ys, zs, n, 0.0f, 0.0f, 0.0f);
	for (int i = 0; i < n; i++)
	{
		float nn = noise[i] * 4.0f;
		hsl_to_rgb((nn + 1.0f) * 0.5f * 360.0f, 0.72f, 1.0f, color + 3 * (size_t)i);
	}
	free(xs);
}

/* ---- MeshProcessor<4>::init + collapse_bad_quads (MeshProcessor.cpp:25-55, 98-128, 308-396) ------------
 * Serial and order-dependent, restated in the reference's order.  pos [n_verts][3] and quads [n_quads][4] are updated in
 * place (a collapsed quad's kept corner moves to the quad centre, its opposite corner is rewired to it in the neighbouring
 * quads); destroyed [n_quads] and adj_next [n_verts] (optional) receive Primitive::destroyed / DualVertex::adj_next.
 * Returns the reference's bad_count. */
int orc_collapse_bad_quads(float* pos, int n_verts, uint32_t* quads, int n_quads, uint8_t* destroyed, uint8_t* adj_next_out)
{
	if (n_verts <= 0 || n_quads <= 0) return 0;
	uint32_t* cnt = (uint32_t*)calloc((size_t)n_verts, sizeof(uint32_t));
	for (size_t k = 0; k < 4 * (size_t)n_quads; k++) cnt[quads[k]]++;
	uint32_t* off = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)n_verts);
	uint32_t a = 0;
	for (int i = 0; i < n_verts; i++) { off[i] = a; a += cnt[i]; }
	/* init_primitives: adj_block[adj_offset + adj_next++] = i in (prim, corner) order; room for one 4-entry list per collapse */
	uint32_t* adj = (uint32_t*)malloc(sizeof(uint32_t) * ((size_t)a + 4 * (size_t)n_quads + 4));
	uint32_t* next = (uint32_t*)calloc((size_t)n_verts, sizeof(uint32_t));
	for (int i = 0; i < n_quads; i++)
		for (int k = 0; k < 4; k++) { uint32_t v = quads[4 * (size_t)i + k]; adj[off[v] + next[v]++] = (uint32_t)i; }
	uint32_t adj_count = a;
	memset(destroyed, 0, (size_t)n_quads);
	int bad = 0;
	for (uint32_t i = 0; i < (uint32_t)n_quads; i++)
	{
		uint32_t* pv = quads + 4 * (size_t)i;
		uint32_t pair[4], p_out[12] = { 0, 0, 0, 0 };
		int next_p = 0, nx = 0;
		for (int k = 0; k < 4; k++)
		{
			const uint32_t dv = pv[k];
			if (next[dv] == 3)
			{
				pair[nx++] = (uint32_t)k;
				for (int q = 0; q < 3; q++)
				{
					const uint32_t e = adj[off[dv] + q];
					if (e != 0xFFFFFFFFu && e != i) p_out[next_p++] = e;
				}
			}
		}
		if (nx == 4 && next_p == 8) continue;
		if (nx != 2 || next_p != 4 || pair[1] - pair[0] != 2) continue;
		float np3[3] = { 0.0f, 0.0f, 0.0f };
		const uint32_t new_index = pv[pair[0]];
		for (int k = 0; k < 4; k++)
			for (int c = 0; c < 3; c++) np3[c] = np3[c] + pos[3 * (size_t)pv[k] + c];
		for (int c = 0; c < 3; c++) pos[3 * (size_t)new_index + c] = np3[c] * 0.25f;
		next[new_index] = 4;
		const uint32_t p_other = pv[pair[1]];
		for (int k = 0; k < 4; k++)
		{
			uint32_t* nv = quads + 4 * (size_t)p_out[k];
			if (nv[0] == new_index || nv[1] == new_index || nv[2] == new_index || nv[3] == new_index) continue;
			for (int j = 0; j < 4; j++)
				if (nv[j] == p_other) { nv[j] = new_index; break; }
		}
		off[new_index] = adj_count;
		for (int k = 0; k < 4; k++) adj[adj_count++] = p_out[k];
		destroyed[i] = 1;
		bad++;
	}
	if (adj_next_out)
		for (int i = 0; i < n_verts; i++) adj_next_out[i] = (uint8_t)next[i];
	free(cnt); free(off); free(adj); free(next);
	return bad;
}

/* ---- whole chunk / batch --------------------------------------------------------------------------- */

int orc_chunk(const orc_sampler* s, const float pos[3], float size, int dim, float overlap, int iters, int pb, int smooth,
              float* density_io, uint32_t* bits_out, uint8_t* masks_out, orc_mesh* mesh, float* color_out, float* normal_out)
{
	const size_t n = (size_t)dim * dim * dim;
	float op[3], delta;
	orc_chunk_geometry(pos, size, dim, overlap, op, &delta);
	float* D = density_io;
	int own_d = 0;
	if (!D)
	{
		D = (float*)malloc(sizeof(float) * n);
		own_d = 1;
	}
	if (s->kind != ORC_HOST_DENSITY)
		orc_sample_block(s, op, delta, dim, D);
	uint32_t* bits = bits_out ? bits_out : (uint32_t*)malloc(sizeof(uint32_t) * (n / 32 + 1));
	int contains = orc_label_grid(D, dim, bits);
	memset(mesh, 0, sizeof(*mesh));
	if (contains)
	{
		uint8_t* masks = masks_out ? masks_out : (uint8_t*)malloc(n);
		orc_cell_masks(bits, dim, masks);
		orc_extract(D, masks, dim, mesh);
		if (!masks_out) free(masks);
		if (mesh->n_verts)
		{
			float* color = color_out ? color_out : (float*)malloc(sizeof(float) * 3 * (size_t)mesh->n_verts);
			for (int i = 0; i < 3 * mesh->n_verts; i++) color[i] = 1.0f; /* calculate_dual_vertex :681 */
			float* nrm = normal_out;
			if (!nrm && smooth) nrm = (float*)calloc(3 * (size_t)mesh->n_verts, sizeof(float));
			else if (nrm) memset(nrm, 0, sizeof(float) * 3 * (size_t)mesh->n_verts);
			/* ChunkGenerator.cpp:110-124 */
			if (iters > 0 && mesh->n_inds)
				orc_smooth(mesh->pos, color, nrm, mesh->boundary, mesh->valence, mesh->n_verts, mesh->inds, mesh->n_inds, 3, iters, pb, smooth);
			if (!color_out) free(color);
			if (!normal_out) free(nrm);
		}
	}
	if (!bits_out) free(bits);
	if (own_d) free(D);
	return contains;
}

int64_t orc_batch(const orc_sampler* s, const float* pos_size, int n, int dim, const float* overlaps, int iters, int pb,
                  int threads, int32_t* counts)
{
	int64_t total = 0;
#ifdef _OPENMP
	if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : total)
	for (int i = 0; i < n; i++)
	{
		orc_mesh m;
		orc_chunk(s, pos_size + 4 * i, pos_size[4 * i + 3], dim, overlaps ? overlaps[i] : 0.0f, iters, pb, 0, 0, 0, 0, &m, 0, 0);
		if (counts)
		{
			counts[2 * i] = m.n_verts;
			counts[2 * i + 1] = m.n_inds;
		}
		total += m.n_verts;
		orc_mesh_free(&m);
	}
	return total;
}

/* ---- seam pass (build-defined, UNPINNED: WorldStitcher.cpp:26-49, 184-239, 491-572 is non-functional as committed) ----
 * Every chunk is d^3 voxel nodes; node (x,y,z) of a chunk sits at overlap_pos + (x,y,z)*delta and carries the chunk's
 * density sample there.  Around every corner point P of the voxel lattice that lies on a chunk's boundary shell, the 8
 * voxels (of whatever chunk / LOD) containing P's octants form a dual cell, polygonised like stitch_indexes does:
 * corner o = 4*dx + 2*dy + dz, crossing points by _get_intersection (:476-481), triangles from the MC table.
 * P is visited from the lowest-index chunk among the finest chunks touching it; cells reaching outside the chunk set
 * are skipped; triangles with two coincident corners are dropped.  Order: chunk, shell point (x faces, y faces
 * without x borders, z faces without x/y borders; low face first; row-major), table order. */
typedef struct seam_lat { int o[3]; int lg; } seam_lat;

static int seam_ilog2(long long v) { int l = 0; while ((1LL << l) < v) l++; return l; }

int64_t orc_seam(const orc_seam_chunk* ch, int n, int dim, const int32_t* group, int cross_group_only, float** tris_out)
{
	const int d = dim, zc = d / 32;
	*tris_out = 0;
	if (n <= 0) return 0;
	/* lattice origin anchored to the octree (the largest chunk is congruent to the root modulo its own size), not to the
	   batch's minimum corner: valid sub-ranges of a leaf set may have a fine chunk at their minimum corner */
	float smin = ch[0].size, smax = ch[0].size;
	double lo[3] = { ch[0].pos[0], ch[0].pos[1], ch[0].pos[2] }, org[3];
	int largest = 0;
	for (int i = 0; i < n; i++)
	{
		if (ch[i].size < smin) smin = ch[i].size;
		if (ch[i].size > smax) { smax = ch[i].size; largest = i; }
		for (int a = 0; a < 3; a++) if (ch[i].pos[a] < lo[a]) lo[a] = ch[i].pos[a];
	}
	for (int a = 0; a < 3; a++)
	{
		const double pl = (double)ch[largest].pos[a];
		org[a] = pl - ceil((pl - lo[a]) / (double)smax - 1e-6) * (double)smax;
	}
	seam_lat* lat = (seam_lat*)malloc(sizeof(seam_lat) * (size_t)n);
	float (*geo)[4] = (float (*)[4])malloc(sizeof(float) * 4 * (size_t)n);
	int G[3] = { 0, 0, 0 };
	for (int i = 0; i < n; i++)
	{
		long long e = llround((double)ch[i].size / (double)smin);
		if (e < 1 || (e & (e - 1))) { free(lat); free(geo); return -1; }
		lat[i].lg = seam_ilog2(e);
		for (int a = 0; a < 3; a++)
		{
			const double qf = ((double)ch[i].pos[a] - org[a]) / (double)smin;
			long long q = llround(qf);
			if (fabs(qf - (double)q) > 1e-3 || q < 0 || (q % e)) { free(lat); free(geo); return -1; }
			lat[i].o[a] = (int)q;
			if (q + e > G[a]) G[a] = (int)(q + e);
		}
		orc_chunk_geometry(ch[i].pos, ch[i].size, d, ch[i].overlap, geo[i], &geo[i][3]);
	}
	int32_t* map = (int32_t*)malloc(sizeof(int32_t) * (size_t)G[0] * G[1] * G[2]);
	for (size_t i = 0; i < (size_t)G[0] * G[1] * G[2]; i++) map[i] = -1;
	for (int i = 0; i < n; i++)
	{
		const int e = 1 << lat[i].lg;
		for (int x = 0; x < e; x++)
			for (int y = 0; y < e; y++)
				for (int z = 0; z < e; z++)
					map[((size_t)(lat[i].o[0] + x) * G[1] + lat[i].o[1] + y) * G[2] + lat[i].o[2] + z] = i;
	}
	size_t cap = 1024, cnt = 0;
	float* out = (float*)malloc(sizeof(float) * 9 * cap);
	const int npts = 6 * d * d + 2;
	for (int c = 0; c < n; c++)
	{
		const int vs = 1 << lat[c].lg; /* voxel size of chunk c in finest-voxel units */
		for (int t = 0; t < npts; t++)
		{
			/* shell point t -> (i,j,k) */
			int i, j, k, r = t;
			const int fa = (d + 1) * (d + 1), fb = (d - 1) * (d + 1), fc = (d - 1) * (d - 1);
			if (r < 2 * fa) { i = r < fa ? 0 : d; r %= fa; j = r / (d + 1); k = r % (d + 1); }
			else if ((r -= 2 * fa) < 2 * fb) { j = r < fb ? 0 : d; r %= fb; i = 1 + r / (d + 1); k = r % (d + 1); }
			else { r -= 2 * fb; k = r < fc ? 0 : d; r %= fc; i = 1 + r / (d - 1); j = 1 + r % (d - 1); }
			const long long P[3] = { (long long)lat[c].o[0] * d + (long long)i * vs, (long long)lat[c].o[1] * d + (long long)j * vs,
			                         (long long)lat[c].o[2] * d + (long long)k * vs };
			int node_c[8], node_v[8][3], mask = 0, ok = 1, owner = c, multi = 0;
			for (int o = 0; o < 8 && ok; o++)
			{
				const int off[3] = { o >> 2, (o >> 1) & 1, o & 1 };
				long long q[3];
				int m;
				for (int a = 0; a < 3; a++) q[a] = P[a] - 1 + off[a];
				if (q[0] < 0 || q[1] < 0 || q[2] < 0 || q[0] >= (long long)G[0] * d || q[1] >= (long long)G[1] * d || q[2] >= (long long)G[2] * d) { ok = 0; break; }
				m = map[((size_t)(q[0] / d) * G[1] + (size_t)(q[1] / d)) * G[2] + (size_t)(q[2] / d)];
				if (m < 0 || lat[m].lg < lat[c].lg) { ok = 0; break; }
				if (lat[m].lg == lat[c].lg && m < owner) owner = m;
				if (group && group[m] != group[c]) multi = 1;
				node_c[o] = m;
				for (int a = 0; a < 3; a++) node_v[o][a] = (int)((q[a] - (long long)lat[m].o[a] * d) >> lat[m].lg);
				const int bit = (ch[m].bits[((size_t)node_v[o][0] * d + node_v[o][1]) * zc + (node_v[o][2] >> 5)] >> (node_v[o][2] & 31)) & 1;
				mask |= bit << o;
			}
			if (!ok || owner != c || mask == 0 || mask == 255) continue;
			if (cross_group_only && !multi) continue;
			float np[8][3], ns[8];
			for (int o = 0; o < 8; o++)
			{
				const int m = node_c[o];
				for (int a = 0; a < 3; a++) np[o][a] = geo[m][a] + (float)node_v[o][a] * geo[m][3];
				ns[o] = ch[m].density[((size_t)node_v[o][0] * d + node_v[o][1]) * d + node_v[o][2]];
			}
			const uint64_t tp = TRI_PACK[mask];
			const int ni = (int)(tp >> 60);
			for (int q0 = 0; q0 < ni; q0 += 3)
			{
				float v[3][3];
				for (int rr = 0; rr < 3; rr++)
				{
					const int e = (int)((tp >> (4 * (q0 + rr))) & 15);
					int a, b;
					if (e < 4) { a = (((e >> 1) & 1) << 1) | (e & 1); b = a | 4; }
					else if (e < 8) { a = ((((e - 4) >> 1) & 1) << 2) | ((e - 4) & 1); b = a | 2; }
					else { a = ((((e - 8) >> 1) & 1) << 2) | (((e - 8) & 1) << 1); b = a | 1; }
					const float mu = (0.0f - ns[a]) / (ns[b] - ns[a]);
					for (int x = 0; x < 3; x++) v[rr][x] = (np[b][x] - np[a][x]) * mu + np[a][x];
				}
				if ((v[0][0] == v[1][0] && v[0][1] == v[1][1] && v[0][2] == v[1][2]) || (v[1][0] == v[2][0] && v[1][1] == v[2][1] && v[1][2] == v[2][2]) ||
				    (v[0][0] == v[2][0] && v[0][1] == v[2][1] && v[0][2] == v[2][2]))
					continue;
				if (cnt == cap) { cap *= 2; out = (float*)realloc(out, sizeof(float) * 9 * cap); }
				memcpy(out + 9 * cnt, v, sizeof(float) * 9);
				cnt++;
			}
		}
	}
	free(lat); free(geo); free(map);
	*tris_out = out;
	return (int64_t)cnt;
}

void orc_free(void* p) { free(p); }

/* ---- quad emission (build-defined, UNPINNED: the reference has the MeshProcessor<4> consumer but no producer) --------
 * Nielson's dual marching cubes on the chunk's (d-1)^3 cells: one dual vertex per surface patch of a cell (patch_pack:
 * the connected components of the cell's tri_table triangles), placed at the mean of the patch's edge crossing points
 * (edges in ascending id, crossing = the triangle emitter's iso-vertex formula); one quad per sign-changing grid edge that
 * has all four cells around it, joining the patch vertices that contain the edge.  Vertex ids follow the serial x->y->z
 * cell scan (patches in table order); quads are emitted in the same scan by the cell at the edge's lower end, X then Y
  * then Z edge; the ring runs counter-clockwise about the +axis and is reversed when the edge's lower endpoint is solid, so
 * the winding agrees with the triangle emitter's (MCTable.h's triangles are clockwise seen from the air side).  boundary = the cell touches the chunk border. */
static const uint64_t PATCH_PACK[256] = ORACLE_PATCH_PACK_INIT;

static inline unsigned q_mask(const uint32_t* bits, int d, int zc, int x, int y, int z)
{
	unsigned m = 0;
	for (int o = 0; o < 8; o++) m |= (unsigned)bit_at(bits, d, zc, x + (o >> 2), y + ((o >> 1) & 1), z + (o & 1)) << o;
	return m;
}

static inline int q_patch_of(unsigned m, int e)
{
	const uint64_t pp = PATCH_PACK[m];
	const int np = (int)(pp >> 60);
	for (int p = 0; p < np; p++)
		if ((pp >> (12 * p)) & (1ull << e)) return p;
	return -1;
}

void orc_quads(const float* D, const uint32_t* bits, int dim, orc_mesh* out)
{
	const int d = dim, zc = (d + 31) / 32, dc = d - 1;
	memset(out, 0, sizeof(*out));
	uint32_t* vbase = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)dc * dc * dc);
	size_t nc = 0, nv = 0, nq = 0;
	for (int x = 0; x < dc; x++)
		for (int y = 0; y < dc; y++)
			for (int z = 0; z < dc; z++)
			{
				const unsigned m = q_mask(bits, d, zc, x, y, z);
				vbase[((size_t)x * dc + y) * dc + z] = (uint32_t)nv;
				if (m == 0 || m == 255) continue;
				nc++;
				nv += (size_t)(PATCH_PACK[m] >> 60);
				nq += (((m ^ (m >> 4)) & 1) && y >= 1 && z >= 1) + (((m ^ (m >> 2)) & 1) && x >= 1 && z >= 1) + (((m ^ (m >> 1)) & 1) && x >= 1 && y >= 1);
			}
	out->n_cells = (int32_t)nc; out->n_verts = (int32_t)nv; out->n_inds = (int32_t)(4 * nq);
	out->pos = (float*)malloc(sizeof(float) * 3 * (nv ? nv : 1));
	out->boundary = (uint8_t*)malloc(nv ? nv : 1);
	out->valence = (uint8_t*)calloc(nv ? nv : 1, 1);
	out->inds = (uint32_t*)malloc(sizeof(uint32_t) * (nq ? 4 * nq : 1));
	size_t v = 0, q = 0;
	for (int x = 0; x < dc; x++)
		for (int y = 0; y < dc; y++)
			for (int z = 0; z < dc; z++)
			{
				const unsigned m = q_mask(bits, d, zc, x, y, z);
				if (m == 0 || m == 255) continue;
				const uint64_t pp = PATCH_PACK[m];
				const int np = (int)(pp >> 60);
				for (int p = 0; p < np; p++)
				{
					const unsigned em = (unsigned)((pp >> (12 * p)) & 0xFFF);
					float sx = 0, sy = 0, sz = 0;
					int cnt = 0;
					for (int e = 0; e < 12; e++)
					{
						if (!((em >> e) & 1)) continue;
						int a, b;
						if (e < 4) { a = (((e >> 1) & 1) << 1) | (e & 1); b = a | 4; }
						else if (e < 8) { a = ((((e - 4) >> 1) & 1) << 2) | ((e - 4) & 1); b = a | 2; }
						else { a = ((((e - 8) >> 1) & 1) << 2) | (((e - 8) & 1) << 1); b = a | 1; }
						float pt[3];
						uint8_t bd;
						iso_vertex(D, d, x + (a >> 2), y + ((a >> 1) & 1), z + (a & 1), x + (b >> 2), y + ((b >> 1) & 1), z + (b & 1), pt, &bd);
						sx += pt[0]; sy += pt[1]; sz += pt[2];
						cnt++;
					}
					out->pos[3 * v] = sx / (float)cnt; out->pos[3 * v + 1] = sy / (float)cnt; out->pos[3 * v + 2] = sz / (float)cnt;
					out->boundary[v] = (uint8_t)(x == 0 || y == 0 || z == 0 || x == dc - 1 || y == dc - 1 || z == dc - 1);
					v++;
				}
				const int b0 = m & 1;
				/* ring of (du, dv) offsets of the four cells around an edge, counter-clockwise about the +axis */
				static const int RX[4][2] = { { 0, 0 }, { 1, 0 }, { 1, 1 }, { 0, 1 } }; /* (dy, dz) */
				static const int RY[4][2] = { { 0, 0 }, { 0, 1 }, { 1, 1 }, { 1, 0 } }; /* (dx, dz) */
				static const int RZ[4][2] = { { 0, 0 }, { 1, 0 }, { 1, 1 }, { 0, 1 } }; /* (dx, dy) */
				for (int axis = 0; axis < 3; axis++)
				{
					const int crossed = axis == 0 ? ((m ^ (m >> 4)) & 1) : axis == 1 ? ((m ^ (m >> 2)) & 1) : ((m ^ (m >> 1)) & 1);
					const int interior = axis == 0 ? (y >= 1 && z >= 1) : axis == 1 ? (x >= 1 && z >= 1) : (x >= 1 && y >= 1);
					if (!crossed || !interior) continue;
					uint32_t id[4];
					for (int r = 0; r < 4; r++)
					{
						const int du = axis == 0 ? RX[r][0] : axis == 1 ? RY[r][0] : RZ[r][0], dv = axis == 0 ? RX[r][1] : axis == 1 ? RY[r][1] : RZ[r][1];
						const int cx = axis == 0 ? x : x - du, cy = axis == 0 ? y - du : (axis == 1 ? y : y - dv), cz = axis == 2 ? z : z - dv;
						const int e = 4 * axis + ((du << 1) | dv);
						const unsigned cm = q_mask(bits, d, zc, cx, cy, cz);
						id[r] = vbase[((size_t)cx * dc + cy) * dc + cz] + (uint32_t)q_patch_of(cm, e);
					}
					for (int r = 0; r < 4; r++)
					{
						const uint32_t vi = b0 ? id[r] : id[3 - r];
						out->inds[4 * q + r] = vi;
						out->valence[vi]++;
					}
					q++;
				}
			}
	free(vbase);
}

/* ---- GLChunk::format_data(vertices, indexes, unwind_verts = true, smooth_normals) (GLChunk.cpp:278-335): flat quads.
 * Per quad the four corner positions and colours are copied out; the normal of all four is the mean of the corner
 * normals (smooth) or -normalize((n0 + n1) / 2) of the two triangle normals with the reference's NaN guards.
 * Pinned against the compiled reference (tests/test_oracle_vs_ref.py). */
static void v3_sub(const float* a, const float* b, float* o) { o[0] = a[0] - b[0]; o[1] = a[1] - b[1]; o[2] = a[2] - b[2]; }

void orc_format_unwind(const float* pos, const float* normal, const float* color, const uint32_t* inds, int n_inds, int smooth_normals,
                       float* p_out, float* n_out, float* c_out)
{
	for (int i = 0; i + 3 < n_inds; i += 4)
	{
		const float* p[4];
		for (int k = 0; k < 4; k++)
		{
			p[k] = pos + 3 * (size_t)inds[i + k];
			memcpy(p_out + 3 * (size_t)(i + k), p[k], 12);
			memcpy(c_out + 3 * (size_t)(i + k), color + 3 * (size_t)inds[i + k], 12);
		}
		float n[3];
		if (smooth_normals)
		{
			const float* a = normal + 3 * (size_t)inds[i], *b = normal + 3 * (size_t)inds[i + 1], *c = normal + 3 * (size_t)inds[i + 2], *d = normal + 3 * (size_t)inds[i + 3];
			for (int x = 0; x < 3; x++) n[x] = (((a[x] + b[x]) + c[x]) + d[x]) * 0.25f;
		}
		else
		{
			float e0[3], e1[3], n0[3], n1[3];
			v3_sub(p[0], p[1], e0); v3_normalize(e0);
			v3_sub(p[0], p[2], e1); v3_normalize(e1);
			v3_cross(e0, e1, n0);
			v3_sub(p[2], p[3], e0); v3_normalize(e0);
			v3_sub(p[2], p[0], e1); v3_normalize(e1);
			v3_cross(e0, e1, n1);
			if (isnan(n0[0])) memcpy(n0, n1, 12);
			if (isnan(n1[0])) memcpy(n1, n0, 12);
			for (int x = 0; x < 3; x++) n[x] = (n0[x] + n1[x]) * 0.5f;
			v3_normalize(n);
			n[0] = -n[0]; n[1] = -n[1]; n[2] = -n[2];
		}
		for (int k = 0; k < 4; k++) memcpy(n_out + 3 * (size_t)(i + k), n, 12);
	}
}

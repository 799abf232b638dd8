/* TEST INFRASTRUCTURE ONLY -- see rt_oracle.h.  CPU oracle ("port") of the reference's per-tile
 * tracing hot path, restated in plain C from the reference's behaviour.  Every function cites
 * the reference file:line it follows.  Build: gcc -O2 -std=c99 -ffp-contract=off (oracle/Makefile).
 */
#define _GNU_SOURCE
#include "rt_oracle.h"

#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * small helpers with the reference's operation order
 * ---------------------------------------------------------------------------------------- */

/* lin_alg.h:138-144 -- Dot() starts from T() and accumulates left to right */
static float dot3(const float *a, const float *b)
{
    float r = 0.0f;
    r += a[0] * b[0];
    r += a[1] * b[1];
    r += a[2] * b[2];
    return r;
}

/* lin_alg.h:145-156 -- Normalize = vec * (1 / sqrt(Dot(vec, vec))) */
static void normalize3(const float *v, float *out)
{
    const float inv_len = 1.0f / sqrtf(dot3(v, v));
    out[0] = v[0] * inv_len;
    out[1] = v[1] * inv_len;
    out[2] = v[2] * inv_len;
}

/* std::min / std::max as ComponentMin / ComponentMax use them (lin_alg.h:157-170) */
static float std_min(float a, float b) { return (b < a) ? b : a; }
static float std_max(float a, float b) { return (a < b) ? b : a; }

static void cnt_add(rto_counters *dst, const rto_counters *src)
{
    dst->rays += src->rays;           dst->cells += src->cells;
    dst->tri_tests += src->tri_tests; dst->hits += src->hits;
    dst->rej_det += src->rej_det;     dst->rej_u += src->rej_u;
    dst->rej_v += src->rej_v;         dst->full += src->full;
    dst->box_miss += src->box_miss;   dst->nonempty += src->nonempty;
}

/* ------------------------------------------------------------------------------------------
 * a2 -- sample table
 * ---------------------------------------------------------------------------------------- */

/* sampling.cpp:194-210 with perm == nullptr */
double rto_radical_inverse(uint32_t n, uint32_t base)
{
    const double inv_base = 1.0 / (double) base;
    double inv_base_i = inv_base;
    double val = 0.0;
    while (n > 0)
    {
        const unsigned int digit = n % base;
        val += digit * inv_base_i;
        inv_base_i *= inv_base;
        n /= base;
    }
    return val;
}

/* renderer.cpp:53-57: x = Hammersley(smp, 0, N) - 0.5f, y = Hammersley(smp, 1, N) - 0.5f with
 * sampling.h:113-120: dim 0 -> double(n)/double(N), dim 1 -> RadicalInverse(n, prime[0] = 2).
 * "double - 0.5f" is a double subtraction, rounded to float by the store. */
void rto_sample_table(uint32_t spp, float *xy)
{
    for (uint32_t s = 0; s < spp; s++)
    {
        xy[2 * s + 0] = (float) ((double) s / (double) spp - 0.5f);
        xy[2 * s + 1] = (float) (rto_radical_inverse(s, 2) - 0.5f);
    }
}

/* ------------------------------------------------------------------------------------------
 * a3 -- camera
 * ---------------------------------------------------------------------------------------- */

/* camera.h:24 and camera.h:41-42; DegToRad = deg * float(0.0174532925) (lin_alg.h:232); the
 * unqualified tan() resolves to the double overload under libstdc++ */
void rto_camera_constants(float fov_deg, uint32_t width, uint32_t height, float *fov_xs, float *aspect)
{
    const float hfov = fov_deg * 0.0174532925f;
    *fov_xs = (float) tan((double) (hfov / 2));
    *aspect = (float) width / (float) height;
}

/* camera.h:20-21 (NDC), :43 (origin = Transf4x4(0), lin_alg.h:518-535),
 * :44-45 (dir = Transf3x3(Normalize(...)), lin_alg.h:495-509).  cam16 = Matrix44f::m_mat flat */
void rto_generate_ray(const float *m, uint32_t px, uint32_t py, uint32_t width, uint32_t height,
                      float off_x, float off_y, float fov_xs, float aspect, float *origin, float *dir)
{
    const float ndc_x = (px + off_x) / (float) width * 2.0f - 1.0f;
    const float ndc_y = (py + off_y) / (float) height * 2.0f - 1.0f;
    const float zero = 0.0f;
    float d[3], n[3];

    origin[0] = zero * m[0] + zero * m[4] + zero * m[8] + m[12];
    origin[1] = zero * m[1] + zero * m[5] + zero * m[9] + m[13];
    origin[2] = zero * m[2] + zero * m[6] + zero * m[10] + m[14];

    d[0] = ndc_x * fov_xs;
    d[1] = ndc_y * fov_xs / aspect;
    d[2] = -1.0f;
    normalize3(d, n);
    dir[0] = n[0] * m[0] + n[1] * m[4] + n[2] * m[8];
    dir[1] = n[0] * m[1] + n[1] * m[5] + n[2] * m[9];
    dir[2] = n[0] * m[2] + n[1] * m[6] + n[2] * m[10];
}

/* camera.h:25-36: origin = Transf4x4((ndc.x * w/2, ndc.y * h/2, 0)), dir = Transf3x3((0, 0, -1)), every
 * product and sum of lin_alg.h:495-535 spelled out (a zero factor still yields -0 / NaN where IEEE says so) */
void rto_generate_ray_ortho(const float *m, uint32_t px, uint32_t py, uint32_t width, uint32_t height,
                            float off_x, float off_y, float ortho_width, float aspect, float *origin, float *dir)
{
    const float ndc_x = (px + off_x) / (float) width * 2.0f - 1.0f;
    const float ndc_y = (py + off_y) / (float) height * 2.0f - 1.0f;
    const float w = ortho_width;
    const float h = (float) w / aspect;
    const float p[3] = { (float) (ndc_x * ((float) w / 2.0)), (float) (ndc_y * ((float) h / 2.0)), (float) 0.0 };
    const float f[3] = { 0.0f, 0.0f, -1.0f };
    for (int j = 0; j < 3; j++)
    {
        origin[j] = p[0] * m[0 + j] + p[1] * m[4 + j] + p[2] * m[8 + j] + m[12 + j];
        dir[j] = f[0] * m[0 + j] + f[1] * m[4 + j] + f[2] * m[8 + j];
    }
}

/* ------------------------------------------------------------------------------------------
 * a4 -- box tests
 * ---------------------------------------------------------------------------------------- */

/* aabb.h:9-13, inclusive on all six faces */
int rto_point_in_aabb(const float *p, const float *mn, const float *mx)
{
    return p[0] >= mn[0] && p[1] >= mn[1] && p[2] >= mn[2] &&
           p[0] <= mx[0] && p[1] <= mx[1] && p[2] <= mx[2];
}

/* aabb.h:34-83, Williams et al. slab test; inv = 1/dir may be +-inf, no t >= 0 check */
int rto_ray_aabb(const float *o, const float *d, const float *mn, const float *mx, float *tmin_out,
                 float *tmax_out)
{
    const float inv[3] = { 1.0f / d[0], 1.0f / d[1], 1.0f / d[2] };
    const float *box[2] = { mn, mx };
    const int sx = inv[0] < 0.0f ? 1 : 0, sy = inv[1] < 0.0f ? 1 : 0, sz = inv[2] < 0.0f ? 1 : 0;
    float tmin = (box[sx][0] - o[0]) * inv[0];
    float tmax = (box[1 - sx][0] - o[0]) * inv[0];
    const float tymin = (box[sy][1] - o[1]) * inv[1];
    const float tymax = (box[1 - sy][1] - o[1]) * inv[1];
    *tmin_out = tmin;
    *tmax_out = tmax;
    if ((tmin > tymax) || (tymin > tmax))
        return 0;
    if (tymin > tmin) tmin = tymin;
    if (tymax < tmax) tmax = tymax;
    const float tzmin = (box[sz][2] - o[2]) * inv[2];
    const float tzmax = (box[1 - sz][2] - o[2]) * inv[2];
    *tmin_out = tmin;
    *tmax_out = tmax;
    if ((tmin > tzmax) || (tzmin > tmax))
        return 0;
    if (tzmin > tmin) tmin = tzmin;
    if (tzmax < tmax) tmax = tzmax;
    *tmin_out = tmin;
    *tmax_out = tmax;
    return 1;
}

/* ------------------------------------------------------------------------------------------
 * a6 / a6' -- ray / triangle
 * ---------------------------------------------------------------------------------------- */

/* triangle.h:15-107, non-culling branch (:77-98).  The products are summed a0*b0+a1*b1+a2*b2
 * (no leading zero, unlike Dot()), the cross products are as the macro writes them (:27-31) */
int rto_ray_tri(const float *o, const float *d, const float *v0, const float *v1, const float *v2,
                float *t, float *u, float *v, rto_counters *cnt)
{
    const float eps = 0.00000001f;
    const float e1[3] = { v1[0] - v0[0], v1[1] - v0[1], v1[2] - v0[2] };
    const float e2[3] = { v2[0] - v0[0], v2[1] - v0[1], v2[2] - v0[2] };
    const float p[3] = { d[1] * e2[2] - d[2] * e2[1], d[2] * e2[0] - d[0] * e2[2],
                         d[0] * e2[1] - d[1] * e2[0] };
    const float det = e1[0] * p[0] + e1[1] * p[1] + e1[2] * p[2];
    if (cnt) cnt->tri_tests++;
    if (det > -eps && det < eps)
    {
        if (cnt) cnt->rej_det++;
        return 0;
    }
    const float inv_det = 1.0f / det;
    const float tv[3] = { o[0] - v0[0], o[1] - v0[1], o[2] - v0[2] };
    *u = (tv[0] * p[0] + tv[1] * p[1] + tv[2] * p[2]) * inv_det;
    if (*u < 0.0f || *u > 1.0f)
    {
        if (cnt) cnt->rej_u++;
        return 0;
    }
    const float q[3] = { tv[1] * e1[2] - tv[2] * e1[1], tv[2] * e1[0] - tv[0] * e1[2],
                         tv[0] * e1[1] - tv[1] * e1[0] };
    *v = (d[0] * q[0] + d[1] * q[1] + d[2] * q[2]) * inv_det;
    if (*v < 0.0f || *u + *v > 1.0f)
    {
        if (cnt) cnt->rej_v++;
        return 0;
    }
    *t = (e2[0] * q[0] + e2[1] * q[1] + e2[2] * q[2]) * inv_det;
    if (cnt) cnt->full++;
    return *t >= 0.0f;
}

/* triangle.h:210-226 = IntersectRayPlane (:200-208) then ComputeBarycentric (:133-156).
 * n is the face normal Mesh::Triangle::n.  Uses Dot() (leading zero). */
int rto_ray_tri_bary(const float *o, const float *d, const float *v0, const float *v1, const float *v2,
                     const float *n, float *t, float *u, float *v, rto_counters *cnt)
{
    if (cnt) cnt->tri_tests++;
    const float denom = dot3(n, d);
    if (fabsf(denom) < 0.00000001f)
    {
        if (cnt) cnt->rej_det++;
        return 0;
    }
    const float dd = dot3(n, v0);
    *t = (dd - dot3(n, o)) / denom;
    if (!(*t >= 0.0))
    {
        if (cnt) cnt->rej_det++;
        return 0;
    }
    const float pos[3] = { o[0] + d[0] * *t, o[1] + d[1] * *t, o[2] + d[2] * *t };
    const float e0[3] = { v2[0] - v0[0], v2[1] - v0[1], v2[2] - v0[2] };
    const float e1[3] = { v1[0] - v0[0], v1[1] - v0[1], v1[2] - v0[2] };
    const float e2[3] = { pos[0] - v0[0], pos[1] - v0[1], pos[2] - v0[2] };
    const float dot00 = dot3(e0, e0);
    const float dot01 = dot3(e0, e1);
    const float dot02 = dot3(e0, e2);
    const float dot11 = dot3(e1, e1);
    const float dot12 = dot3(e1, e2);
    const float inv_denom = 1 / (dot00 * dot11 - dot01 * dot01);
    *u = (dot00 * dot12 - dot01 * dot02) * inv_denom;
    *v = (dot11 * dot02 - dot01 * dot12) * inv_denom;
    if (cnt) cnt->full++;
    return (*u >= 0) && (*v >= 0) && (*u + *v < 1);
}

/* ------------------------------------------------------------------------------------------
 * a9 -- grid construction
 * ---------------------------------------------------------------------------------------- */

/* aabb_tri_internal.h:42-62 */
static int plane_box_overlap(const double normal[3], double d, const double maxbox[3])
{
    double vmin[3], vmax[3];
    for (int q = 0; q < 3; q++)
    {
        if (normal[q] > 0.0f) { vmin[q] = -maxbox[q]; vmax[q] = maxbox[q]; }
        else                  { vmin[q] = maxbox[q];  vmax[q] = -maxbox[q]; }
    }
    if (normal[0] * vmin[0] + normal[1] * vmin[1] + normal[2] * vmin[2] + d > 0.0f) return 0;
    if (normal[0] * vmax[0] + normal[1] * vmax[1] + normal[2] * vmax[2] + d >= 0.0f) return 1;
    return 0;
}

/* aabb_tri_internal.h:112-186 (Akenine-Moller SAT, fp64).  The nine edge-cross-axis tests are
 * table driven here: for triangle edge i and box axis a the reference projects the two vertices
 * listed in PAIR[i][a] (the AXISTEST_* macro it picks, :65-110) with
 *     X: p = e.z*v.y - e.y*v.z   rad = |e.z|*h.y + |e.y|*h.z
 *     Y: p = e.x*v.z - e.z*v.x   rad = |e.z|*h.x + |e.x|*h.z   (written -e.z*v.x + e.x*v.z)
 *     Z: p = e.y*v.x - e.x*v.y   rad = |e.y|*h.x + |e.x|*h.y
 * and rejects when min(p) > rad or max(p) < -rad. */
int rto_tri_box_overlap(const double center[3], const double half[3], const double tri[3][3])
{
    static const int PAIR[3][3][2] = {
        { { 0, 2 }, { 0, 2 }, { 1, 2 } },   /* edge 0 = v1 - v0: X01, Y02, Z12 */
        { { 0, 2 }, { 0, 2 }, { 0, 1 } },   /* edge 1 = v2 - v1: X01, Y02, Z0  */
        { { 0, 1 }, { 0, 1 }, { 1, 2 } } }; /* edge 2 = v0 - v2: X2,  Y1,  Z12 */
    double v[3][3], e[3][3];
    for (int i = 0; i < 3; i++)
        for (int k = 0; k < 3; k++)
            v[i][k] = tri[i][k] - center[k];
    for (int k = 0; k < 3; k++)
    {
        e[0][k] = v[1][k] - v[0][k];
        e[1][k] = v[2][k] - v[1][k];
        e[2][k] = v[0][k] - v[2][k];
    }
    for (int i = 0; i < 3; i++)
    {
        const double fex = fabs(e[i][0]), fey = fabs(e[i][1]), fez = fabs(e[i][2]);
        for (int a = 0; a < 3; a++)
        {
            const double *va = v[PAIR[i][a][0]], *vb = v[PAIR[i][a][1]];
            double pa, pb, rad;
            if (a == 0)
            {
                pa = e[i][2] * va[1] - e[i][1] * va[2];
                pb = e[i][2] * vb[1] - e[i][1] * vb[2];
                rad = fez * half[1] + fey * half[2];
            }
            else if (a == 1)
            {
                pa = -e[i][2] * va[0] + e[i][0] * va[2];
                pb = -e[i][2] * vb[0] + e[i][0] * vb[2];
                rad = fez * half[0] + fex * half[2];
            }
            else
            {
                pa = e[i][1] * va[0] - e[i][0] * va[1];
                pb = e[i][1] * vb[0] - e[i][0] * vb[1];
                rad = fey * half[0] + fex * half[1];
            }
            const double mn = pa < pb ? pa : pb, mx = pa < pb ? pb : pa;
            if (mn > rad || mx < -rad)
                return 0;
        }
    }
    for (int k = 0; k < 3; k++)
    {
        double mn = v[0][k], mx = v[0][k];
        if (v[1][k] < mn) mn = v[1][k];
        if (v[1][k] > mx) mx = v[1][k];
        if (v[2][k] < mn) mn = v[2][k];
        if (v[2][k] > mx) mx = v[2][k];
        if (mn > half[k] || mx < -half[k])
            return 0;
    }
    double normal[3];
    normal[0] = e[0][1] * e[1][2] - e[0][2] * e[1][1];
    normal[1] = e[0][2] * e[1][0] - e[0][0] * e[1][2];
    normal[2] = e[0][0] * e[1][1] - e[0][1] * e[1][0];
    const double d = -(normal[0] * v[0][0] + normal[1] * v[0][1] + normal[2] * v[0][2]);
    return plane_box_overlap(normal, d, half) ? 1 : 0;
}

/* aabb.h:15-32 -- centre / half size are computed in fp32 and only then widened */
static int tri_aabb_overlap(const float *v0, const float *v1, const float *v2, const float *mn,
                            const float *mx)
{
    double center[3], half[3], tri[3][3];
    for (int k = 0; k < 3; k++)
    {
        center[k] = (mn[k] + mx[k]) * 0.5f;
        half[k] = (mx[k] - mn[k]) * 0.5f;
        tri[0][k] = v0[k];
        tri[1][k] = v1[k];
        tri[2][k] = v2[k];
    }
    return rto_tri_box_overlap(center, half, tri);
}

int rto_grid_build_tight_ranges = 0;

typedef struct pair_list
{
    uint64_t *cell;
    uint32_t *tri;
    size_t n, cap;
} pair_list;

static void pairs_push(pair_list *pl, uint64_t cell, uint32_t tri)
{
    if (pl->n == pl->cap)
    {
        pl->cap = pl->cap ? pl->cap * 2 : 4096;
        pl->cell = (uint64_t *) realloc(pl->cell, pl->cap * sizeof(uint64_t));
        pl->tri = (uint32_t *) realloc(pl->tri, pl->cap * sizeof(uint32_t));
    }
    pl->cell[pl->n] = cell;
    pl->tri[pl->n] = tri;
    pl->n++;
}

typedef struct build_job
{
    const rto_scene *scene;
    uint32_t tri_begin, tri_end;
    pair_list pairs;
} build_job;

/* grid.cpp:65-129 for triangles [tri_begin, tri_end): candidate range from the triangle's AABB
 * (triangle.h:116-131 -- note the max is seeded with numeric_limits<float>::min(), a tiny
 * POSITIVE number, :123), then the exact SAT per candidate cell, x -> y -> z loop order */
static void *build_worker(void *arg)
{
    build_job *job = (build_job *) arg;
    const rto_scene *sc = job->scene;
    const rto_grid *g = &sc->grid;
    for (uint32_t ti = job->tri_begin; ti < job->tri_end; ti++)
    {
        const uint32_t *tr = sc->tri + (size_t) ti * 6;
        const float *v0 = sc->vtx + (size_t) tr[0] * 6;
        const float *v1 = sc->vtx + (size_t) tr[1] * 6;
        const float *v2 = sc->vtx + (size_t) tr[2] * 6;
        float tmin[3], tmax[3];
        uint32_t start[3], end[3];
        for (int k = 0; k < 3; k++)
        {
            float mn = FLT_MAX, mx = FLT_MIN;
            mn = std_min(mn, v0[k]); mx = std_max(mx, v0[k]);
            mn = std_min(mn, v1[k]); mx = std_max(mx, v1[k]);
            mn = std_min(mn, v2[k]); mx = std_max(mx, v2[k]);
            tmin[k] = mn - g->aabb_min[k];
            tmax[k] = mx - g->aabb_min[k];
            start[k] = (uint32_t) (tmin[k] / g->cell_wdh);
            end[k] = (uint32_t) (tmax[k] / g->cell_wdh);
            if (rto_grid_build_tight_ranges)
            {
                /* Optional (off by default, so the port stays the literal algorithm): cut the
                 * candidate range at the triangle's true maximum + 1 guard cell.  The lists cannot
                 * change -- the SAT alone decides membership -- but the reference's FLT_MIN-seeded
                 * maximum makes the literal loop take minutes on large all-negative-octant scenes.
                 * tests/test_oracle_vs_ref.py checks both modes give identical grids. */
                const float true_mx = std_max(std_max(v0[k], v1[k]), v2[k]);
                const uint32_t e_true = (uint32_t) ((true_mx - g->aabb_min[k]) / g->cell_wdh) + 1u;
                if (end[k] > e_true) end[k] = e_true;
            }
        }
        for (uint32_t x = start[0]; x <= end[0]; x++)
            for (uint32_t y = start[1]; y <= end[1]; y++)
                for (uint32_t z = start[2]; z <= end[2]; z++)
                {
                    const float cmin[3] = { g->aabb_min[0] + x * g->cell_wdh,
                                            g->aabb_min[1] + y * g->cell_wdh,
                                            g->aabb_min[2] + z * g->cell_wdh };
                    const float cmax[3] = { g->aabb_min[0] + (x + 1) * g->cell_wdh,
                                            g->aabb_min[1] + (y + 1) * g->cell_wdh,
                                            g->aabb_min[2] + (z + 1) * g->cell_wdh };
                    if (tri_aabb_overlap(v0, v1, v2, cmin, cmax))
                    {
                        /* grid.h:41-42; the reference asserts cell_idx < cells (grid.cpp:121) */
                        const uint64_t cell = x + (uint64_t) z * g->dim[0] +
                                              (uint64_t) y * g->dim[0] * g->dim[2];
                        if (x < g->dim[0] && y < g->dim[1] && z < g->dim[2])
                            pairs_push(&job->pairs, cell, ti);
                    }
                }
    }
    return NULL;
}

int rto_grid_build(rto_scene *sc, uint32_t grid_res, uint32_t n_threads)
{
    rto_grid *g = &sc->grid;
    memset(g, 0, sizeof(*g));
    if (sc->num_tri == 0 || sc->num_vtx == 0 || grid_res == 0)
        return 1;

    /* mesh.cpp:72-94 (only vertices referenced by triangles; max seeded with FLT_MIN > 0) */
    float mn[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, mx[3] = { FLT_MIN, FLT_MIN, FLT_MIN };
    for (uint32_t ti = 0; ti < sc->num_tri; ti++)
        for (int c = 0; c < 3; c++)
        {
            const float *p = sc->vtx + (size_t) sc->tri[(size_t) ti * 6 + c] * 6;
            for (int k = 0; k < 3; k++)
            {
                mn[k] = std_min(mn[k], p[k]);
                mx[k] = std_max(mx[k], p[k]);
            }
        }
    /* grid.cpp:29-38 */
    float ext[3];
    for (int k = 0; k < 3; k++)
    {
        g->aabb_min[k] = mn[k] - 0.0001f;
        g->aabb_max[k] = mx[k] + 0.0001f;
        ext[k] = g->aabb_max[k] - g->aabb_min[k];
    }
    const float largest = std_max(std_max(ext[0], ext[1]), ext[2]);
    g->cell_wdh = largest / (float) grid_res;
    g->inv_cell_wdh = 1.0f / g->cell_wdh;
    for (int k = 0; k < 3; k++)
        g->dim[k] = (uint32_t) ceilf(ext[k] / g->cell_wdh);
    g->num_cells = (uint64_t) g->dim[0] * g->dim[1] * g->dim[2];

    if (n_threads < 1) n_threads = 1;
    if (n_threads > sc->num_tri) n_threads = sc->num_tri;
    build_job *jobs = (build_job *) calloc(n_threads, sizeof(build_job));
    pthread_t *th = (pthread_t *) calloc(n_threads, sizeof(pthread_t));
    for (uint32_t i = 0; i < n_threads; i++)
    {
        jobs[i].scene = sc;
        jobs[i].tri_begin = (uint32_t) ((uint64_t) sc->num_tri * i / n_threads);
        jobs[i].tri_end = (uint32_t) ((uint64_t) sc->num_tri * (i + 1) / n_threads);
        pthread_create(&th[i], NULL, build_worker, &jobs[i]);
    }
    for (uint32_t i = 0; i < n_threads; i++)
        pthread_join(th[i], NULL);

    /* Stable counting sort by cell: jobs are contiguous ascending triangle ranges, so visiting
     * them in order reproduces the push_back order of grid.cpp:122 (ascending tri_idx) */
    g->cell_offset = (uint64_t *) calloc(g->num_cells + 1, sizeof(uint64_t));
    for (uint32_t i = 0; i < n_threads; i++)
        for (size_t k = 0; k < jobs[i].pairs.n; k++)
            g->cell_offset[jobs[i].pairs.cell[k] + 1]++;
    for (uint64_t c = 0; c < g->num_cells; c++)
        g->cell_offset[c + 1] += g->cell_offset[c];
    g->num_refs = g->cell_offset[g->num_cells];
    g->tri_index = (uint32_t *) malloc((g->num_refs ? g->num_refs : 1) * sizeof(uint32_t));
    uint64_t *cursor = (uint64_t *) malloc((g->num_cells ? g->num_cells : 1) * sizeof(uint64_t));
    memcpy(cursor, g->cell_offset, g->num_cells * sizeof(uint64_t));
    for (uint32_t i = 0; i < n_threads; i++)
    {
        for (size_t k = 0; k < jobs[i].pairs.n; k++)
            g->tri_index[cursor[jobs[i].pairs.cell[k]]++] = jobs[i].pairs.tri[k];
        free(jobs[i].pairs.cell);
        free(jobs[i].pairs.tri);
    }
    free(cursor);
    free(jobs);
    free(th);
    return 0;
}

void rto_grid_free(rto_grid *g)
{
    free(g->cell_offset);
    free(g->tri_index);
    g->cell_offset = NULL;
    g->tri_index = NULL;
}

/* ------------------------------------------------------------------------------------------
 * a5 -- 3D-DDA traversal
 * ---------------------------------------------------------------------------------------- */

/* grid.h:44-48 and grid.h:50-51 */
static int to_voxel(const rto_grid *g, const float *p, int axis)
{
    int v = (int) ((p[axis] - g->aabb_min[axis]) * g->inv_cell_wdh);
    const int hi = (int) g->dim[axis] - 1;
    if (v < 0) return 0;
    if (v > hi) return hi;
    return v;
}
static float to_pos(const rto_grid *g, int vox, int axis) { return g->aabb_min[axis] + vox * g->cell_wdh; }

/* grid.cpp:159-281 */
int rto_grid_intersect(const rto_scene *sc, const float *origin, const float *dir, int variant,
                       float *t, float *u, float *v, uint32_t *tri_idx, rto_counters *cnt)
{
    const rto_grid *g = &sc->grid;
    float enter_t, leave_t, gi[3];
    if (cnt) cnt->rays++;

    /* :175-185 */
    if (rto_point_in_aabb(origin, g->aabb_min, g->aabb_max))
    {
        enter_t = 0.0f;
        gi[0] = origin[0]; gi[1] = origin[1]; gi[2] = origin[2];
    }
    else if (rto_ray_aabb(origin, dir, g->aabb_min, g->aabb_max, &enter_t, &leave_t))
    {
        gi[0] = origin[0] + dir[0] * enter_t;
        gi[1] = origin[1] + dir[1] * enter_t;
        gi[2] = origin[2] + dir[2] * enter_t;
    }
    else
    {
        if (cnt) cnt->box_miss++;
        return 0;
    }

    /* :188-216.  For dir == 0 the reference leaves delta/step/out uninitialised; such an axis
     * can never win the argmin below while another axis is finite, so benign values do */
    float next_t[3], delta_t[3] = { 0.0f, 0.0f, 0.0f };
    int step[3] = { 1, 1, 1 }, out[3], pos[3];
    for (int a = 0; a < 3; a++)
    {
        out[a] = (int) g->dim[a];
        pos[a] = to_voxel(g, gi, a);
        if (dir[a] == 0.0f)
            next_t[a] = FLT_MAX;
        else if (dir[a] > 0.0f)
        {
            next_t[a] = enter_t + (to_pos(g, pos[a] + 1, a) - gi[a]) / dir[a];
            delta_t[a] = g->cell_wdh / dir[a];
            step[a] = 1;
            out[a] = (int) g->dim[a];
        }
        else
        {
            next_t[a] = enter_t + (to_pos(g, pos[a], a) - gi[a]) / dir[a];
            delta_t[a] = -g->cell_wdh / dir[a];
            step[a] = -1;
            out[a] = -1;
        }
    }

    /* :219-278 */
    *t = FLT_MAX;
    for (;;)
    {
        const int sa = (next_t[0] < next_t[1]) ? ((next_t[0] < next_t[2]) ? 0 : 2)
                                               : ((next_t[1] < next_t[2]) ? 1 : 2);
        const uint64_t cell = (uint64_t) pos[0] + (uint64_t) pos[2] * g->dim[0] +
                              (uint64_t) pos[1] * g->dim[0] * g->dim[2];
        if (cnt) cnt->cells++;
        if (cnt && g->cell_offset[cell] != g->cell_offset[cell + 1]) cnt->nonempty++;
        for (uint64_t k = g->cell_offset[cell]; k < g->cell_offset[cell + 1]; k++)
        {
            const uint32_t ci = g->tri_index[k];
            const uint32_t *tr = sc->tri + (size_t) ci * 6;
            const float *v0 = sc->vtx + (size_t) tr[0] * 6;
            const float *v1 = sc->vtx + (size_t) tr[1] * 6;
            const float *v2 = sc->vtx + (size_t) tr[2] * 6;
            float ct, cu, cv;
            int hit;
            if (variant == RTO_VARIANT_BARY)
                hit = rto_ray_tri_bary(origin, dir, v0, v1, v2, (const float *) (tr + 3), &ct, &cu, &cv, cnt);
            else
                hit = rto_ray_tri(origin, dir, v0, v1, v2, &ct, &cu, &cv, cnt);
            if (hit && ct < *t && ct < next_t[sa])
            {
                *t = ct; *u = cu; *v = cv; *tri_idx = ci;
            }
        }
        if (*t != FLT_MAX)
        {
            if (cnt) cnt->hits++;
            return 1;
        }
        pos[sa] += step[sa];
        if (pos[sa] == out[sa])
            break;
        next_t[sa] += delta_t[sa];
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * a7 / a8 -- shading and resolve
 * ---------------------------------------------------------------------------------------- */

/* renderer.cpp:109-117: n = Normalize(n1*u + n2*v + n0*(1-u-v)) (triangle.h:158-161), then
 * colour = (n + 1) * 0.5 */
void rto_shade_hit(const rto_scene *sc, uint32_t tri_idx, float u, float v, float *rgb)
{
    const uint32_t *tr = sc->tri + (size_t) tri_idx * 6;
    const float *n0 = sc->vtx + (size_t) tr[0] * 6 + 3;
    const float *n1 = sc->vtx + (size_t) tr[1] * 6 + 3;
    const float *n2 = sc->vtx + (size_t) tr[2] * 6 + 3;
    const float w = 1 - u - v;
    float n[3], nn[3];
    for (int k = 0; k < 3; k++)
        n[k] = n1[k] * u + n2[k] * v + n0[k] * w;
    normalize3(n, nn);
    for (int k = 0; k < 3; k++)
        rgb[k] = (nn[k] + 1.0f) * 0.5f;
}

/* renderer.cpp:124-133 and lin_alg.h:125-132 */
uint32_t rto_resolve_pixel(const float *rgb_sum, uint32_t spp, int gamma)
{
    float c[3];
    unsigned char b[3];
    for (int k = 0; k < 3; k++)
    {
        c[k] = rgb_sum[k] / (float) spp;
        if (gamma)
            c[k] = powf(c[k], 1.0f / 2.0f);
        b[k] = c[k] > 1.0f ? 255 : (unsigned char) (c[k] * 255.0f);
    }
    return (uint32_t) (b[0] << 16 | b[1] << 8 | b[2] << 0);
}

/* ------------------------------------------------------------------------------------------
 * a1 -- the tile kernel over rows
 * ---------------------------------------------------------------------------------------- */

typedef struct render_job
{
    const rto_scene *scene;
    const float *cam16;
    const float *smp;
    float fov_xs, aspect;
    uint32_t width, height, spp, y_begin, y_end;
    int variant, gamma;
    rto_render_options opt;
    uint32_t *next_row;
    uint32_t *bgra, *hit_tri;
    float *hit_t, *hit_u, *hit_v;
    rto_counters cnt;
} render_job;

static void *render_worker(void *arg)
{
    render_job *j = (render_job *) arg;
    for (;;)
    {
        const uint32_t y = __atomic_fetch_add(j->next_row, 1u, __ATOMIC_RELAXED);
        if (y >= j->y_end)
            break;
        for (uint32_t x = 0; x < j->width; x++)
        {
            float col[3] = { 0.0f, 0.0f, 0.0f };
            for (uint32_t s = 0; s < j->spp; s++)
            {
                float o[3], d[3], t = 0, u = 0, v = 0, rgb[3];
                uint32_t idx = RTO_MISS;
                if (j->opt.ortho)
                    rto_generate_ray_ortho(j->cam16, x, y, j->width, j->height, j->smp[2 * s], j->smp[2 * s + 1],
                                           j->opt.ortho_width, j->aspect, o, d);
                else
                    rto_generate_ray(j->cam16, x, y, j->width, j->height, j->smp[2 * s], j->smp[2 * s + 1],
                                     j->fov_xs, j->aspect, o, d);
                const int hit = rto_grid_intersect(j->scene, o, d, j->variant, &t, &u, &v, &idx, &j->cnt);
                if (hit && j->opt.shade_mode == RTO_SHADE_FACE_NORMAL)
                {
                    /* renderer.cpp:116 "Vec3f n = tri.n;" then :117 */
                    const float *n = (const float *) (j->scene->tri + (size_t) idx * 6 + 3);
                    for (int k = 0; k < 3; k++)
                        rgb[k] = (n[k] + 1.0f) * 0.5f;
                }
                else if (hit && j->opt.shade_mode == RTO_SHADE_DEPTH)
                    rgb[0] = rgb[1] = rgb[2] = t / 3; /* renderer.cpp:118 "col += Vec3f(t / 3);" */
                else if (hit)
                    rto_shade_hit(j->scene, idx, u, v, rgb);
                else
                    rgb[0] = rgb[1] = rgb[2] = (float) y / (float) j->height; /* renderer.cpp:121 */
                col[0] += rgb[0]; col[1] += rgb[1]; col[2] += rgb[2];
                const size_t k = ((size_t) (y - j->y_begin) * j->width + x) * j->spp + s;
                if (j->hit_tri) j->hit_tri[k] = hit ? idx : RTO_MISS;
                if (j->hit_t) j->hit_t[k] = hit ? t : 0.0f;
                if (j->hit_u) j->hit_u[k] = hit ? u : 0.0f;
                if (j->hit_v) j->hit_v[k] = hit ? v : 0.0f;
            }
            if (j->bgra)
                j->bgra[(size_t) (y - j->y_begin) * j->width + x] = rto_resolve_pixel(col, j->spp, j->gamma);
        }
    }
    return NULL;
}

void rto_render_rows(const rto_scene *scene, const float *cam16, float fov_deg, uint32_t width,
                     uint32_t height, uint32_t spp, int variant, int gamma, uint32_t y_begin,
                     uint32_t y_end, uint32_t n_threads, uint32_t *bgra, uint32_t *hit_tri,
                     float *hit_t, float *hit_u, float *hit_v, rto_counters *cnt)
{
    rto_render_rows_ex(scene, cam16, fov_deg, width, height, spp, variant, gamma, y_begin, y_end, n_threads, NULL, bgra,
                       hit_tri, hit_t, hit_u, hit_v, cnt);
}

void rto_render_rows_ex(const rto_scene *scene, const float *cam16, float fov_deg, uint32_t width,
                        uint32_t height, uint32_t spp, int variant, int gamma, uint32_t y_begin,
                        uint32_t y_end, uint32_t n_threads, const rto_render_options *opt, uint32_t *bgra,
                        uint32_t *hit_tri, float *hit_t, float *hit_u, float *hit_v, rto_counters *cnt)
{
    float *smp = (float *) malloc(sizeof(float) * 2 * (spp ? spp : 1));
    float fov_xs, aspect;
    uint32_t next_row = y_begin;
    rto_sample_table(spp, smp);
    rto_camera_constants(fov_deg, width, height, &fov_xs, &aspect);
    if (n_threads < 1) n_threads = 1;
    render_job *jobs = (render_job *) calloc(n_threads, sizeof(render_job));
    pthread_t *th = (pthread_t *) calloc(n_threads, sizeof(pthread_t));
    for (uint32_t i = 0; i < n_threads; i++)
    {
        render_job *j = &jobs[i];
        j->scene = scene; j->cam16 = cam16; j->smp = smp; j->fov_xs = fov_xs; j->aspect = aspect;
        j->width = width; j->height = height; j->spp = spp; j->y_begin = y_begin; j->y_end = y_end;
        j->variant = variant; j->gamma = gamma; j->next_row = &next_row;
        if (opt) j->opt = *opt;
        j->bgra = bgra; j->hit_tri = hit_tri; j->hit_t = hit_t; j->hit_u = hit_u; j->hit_v = hit_v;
        pthread_create(&th[i], NULL, render_worker, j);
    }
    if (cnt) memset(cnt, 0, sizeof(*cnt));
    for (uint32_t i = 0; i < n_threads; i++)
    {
        pthread_join(th[i], NULL);
        if (cnt) cnt_add(cnt, &jobs[i].cnt);
    }
    free(jobs);
    free(th);
    free(smp);
}

void rto_intersect_rays(const rto_scene *scene, uint32_t n, const float *origins, const float *dirs,
                        int variant, uint32_t *tri_idx, float *t, float *u, float *v)
{
    for (uint32_t i = 0; i < n; i++)
    {
        float ct = 0, cu = 0, cv = 0;
        uint32_t idx = RTO_MISS;
        const int hit = rto_grid_intersect(scene, origins + 3 * (size_t) i, dirs + 3 * (size_t) i, variant,
                                           &ct, &cu, &cv, &idx, NULL);
        tri_idx[i] = hit ? idx : RTO_MISS;
        t[i] = hit ? ct : 0.0f;
        u[i] = hit ? cu : 0.0f;
        v[i] = hit ? cv : 0.0f;
    }
}

/* ------------------------------------------------------------------------------------------
 * gamma: glibc powf(x, 0.5f) vs IEEE sqrtf(x)
 * ---------------------------------------------------------------------------------------- */

typedef struct pow_job
{
    uint32_t lo, hi;
    uint64_t diff, byte_diff;
} pow_job;

static void *pow_worker(void *arg)
{
    pow_job *j = (pow_job *) arg;
    for (uint32_t b = j->lo; b < j->hi; b++)
    {
        float x, a, s;
        memcpy(&x, &b, 4);
        a = powf(x, 0.5f);
        s = sqrtf(x);
        if (memcmp(&a, &s, 4) != 0 && !(a != a && s != s))
        {
            j->diff++;
            const unsigned char ba = a > 1.0f ? 255 : (unsigned char) (a * 255.0f);
            const unsigned char bs = s > 1.0f ? 255 : (unsigned char) (s * 255.0f);
            if (ba != bs)
                j->byte_diff++;
        }
    }
    return NULL;
}

uint64_t rto_powf_vs_sqrtf(uint32_t lo_bits, uint32_t hi_bits, uint32_t n_threads, uint64_t *byte_diff)
{
    if (n_threads < 1) n_threads = 1;
    pow_job *jobs = (pow_job *) calloc(n_threads, sizeof(pow_job));
    pthread_t *th = (pthread_t *) calloc(n_threads, sizeof(pthread_t));
    const uint64_t span = (uint64_t) hi_bits - lo_bits;
    uint64_t diff = 0, bd = 0;
    for (uint32_t i = 0; i < n_threads; i++)
    {
        jobs[i].lo = lo_bits + (uint32_t) (span * i / n_threads);
        jobs[i].hi = lo_bits + (uint32_t) (span * (i + 1) / n_threads);
        pthread_create(&th[i], NULL, pow_worker, &jobs[i]);
    }
    for (uint32_t i = 0; i < n_threads; i++)
    {
        pthread_join(th[i], NULL);
        diff += jobs[i].diff;
        bd += jobs[i].byte_diff;
    }
    if (byte_diff) *byte_diff = bd;
    free(jobs);
    free(th);
    return diff;
}

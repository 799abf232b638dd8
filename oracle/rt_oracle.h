/* TEST INFRASTRUCTURE ONLY -- the CPU oracle ("port") for the per-tile tracing hot path.
 *
 * Plain-C restatement of the reference algorithm.  It is the
 * checker for tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg; nothing in the
 * product path may include, link or call it.  Parity status: PINNED -- tests/test_oracle_vs_ref.py
 * checks every function here against the unmodified reference (oracle/_ref/libref_oracle.so)
 * and tests/golden/ holds reference-generated vectors for boxes where only the port exists.
 *
 * Arithmetic contract: IEEE-754 binary32, no FMA contraction (-ffp-contract=off), the
 * reference's left-to-right operation order; fp64 for the radical inverse and the
 * triangle/box SAT, exactly where the reference uses double.
 */
#ifndef RT_ORACLE_H
#define RT_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RTO_MISS 0xFFFFFFFFu

typedef struct rto_grid
{
    uint32_t  dim[3];
    float     aabb_min[3];
    float     aabb_max[3];
    float     cell_wdh;
    float     inv_cell_wdh;
    uint64_t  num_cells;
    uint64_t  num_refs;
    uint64_t *cell_offset; /* num_cells + 1, cell order x + z*dimx + y*dimx*dimz (grid.h:41-42) */
    uint32_t *tri_index;   /* num_refs, ascending triangle index inside each cell */
} rto_grid;

typedef struct rto_scene
{
    const float    *vtx; /* V x {px,py,pz,nx,ny,nz}          = Mesh::Vertex   (mesh.h:20-24) */
    const uint32_t *tri; /* T x {v0,v1,v2, n.x,n.y,n.z bits} = Mesh::Triangle (mesh.h:12-18) */
    uint32_t        num_vtx;
    uint32_t        num_tri;
    rto_grid        grid;
} rto_scene;

/* Work counters of the reference algorithm (feed SURVEY.md section 8d's flop/byte formula) */
typedef struct rto_counters
{
    uint64_t rays;
    uint64_t cells;      /* cells visited                                   */
    uint64_t tri_tests;  /* ray/triangle tests executed                     */
    uint64_t hits;       /* rays that returned a hit                        */
    uint64_t rej_det;    /* tests rejected at the determinant / plane stage */
    uint64_t rej_u;      /* ... at the u stage                              */
    uint64_t rej_v;      /* ... at the v stage                              */
    uint64_t full;       /* tests that computed t                           */
    uint64_t box_miss;   /* rays that missed the grid's box                 */
    uint64_t nonempty;   /* visited cells with a non-empty list */
} rto_counters;

enum { RTO_VARIANT_MT = 0 /* IntersectRayTri */, RTO_VARIANT_BARY = 1 /* IntersectRayTriBarycentric */ };

/* a2: renderer.cpp:49-60, sampling.h:113-120, sampling.cpp:194-210.  xy = spp x {x,y} */
void rto_sample_table(uint32_t spp, float *xy);
double rto_radical_inverse(uint32_t n, uint32_t base);

/* a3: camera.h:24,41-42 frame constants and camera.h:8-47 perspective branch */
void rto_camera_constants(float fov_deg, uint32_t width, uint32_t height, float *fov_xs, float *aspect);
void rto_generate_ray(const float *cam16, uint32_t px, uint32_t py, uint32_t width, uint32_t height,
                      float off_x, float off_y, float fov_xs, float aspect, float *origin, float *dir);

/* camera.h:25-36, the orthographic branch (never taken by RenderTile, which passes ortho = false).  The
 * reference forms ndc * (float(width) / 2.0) in double and narrows it to float: the product of two floats is
 * exact in double, so this is the correctly rounded float product */
void rto_generate_ray_ortho(const float *cam16, uint32_t px, uint32_t py, uint32_t width, uint32_t height,
                            float off_x, float off_y, float ortho_width, float aspect, float *origin, float *dir);

/* a4: aabb.h:9-13 and aabb.h:34-83 */
int rto_point_in_aabb(const float *p, const float *mn, const float *mx);
int rto_ray_aabb(const float *o, const float *d, const float *mn, const float *mx, float *tmin, float *tmax);

/* a6 / a6': triangle.h:15-107 (non-culling branch) and triangle.h:200-226,133-156 */
int rto_ray_tri(const float *o, const float *d, const float *v0, const float *v1, const float *v2,
                float *t, float *u, float *v, rto_counters *cnt);
int rto_ray_tri_bary(const float *o, const float *d, const float *v0, const float *v1, const float *v2,
                     const float *n, float *t, float *u, float *v, rto_counters *cnt);

/* a9: grid.cpp:12-154 (+ triangle.h:116-131, aabb.h:15-32, aabb_tri_internal.h:42-186,
 * mesh.cpp:72-94).  Fills scene->grid; returns 0 on success */
int  rto_grid_build(rto_scene *scene, uint32_t grid_res, uint32_t n_threads);
extern int rto_grid_build_tight_ranges; /* 1: cut candidate ranges at the true triangle maximum (same lists, faster) */
void rto_grid_free(rto_grid *grid);
int  rto_tri_box_overlap(const double center[3], const double half[3], const double tri[3][3]);

/* a5: grid.cpp:159-281.  Returns 1 on hit */
int rto_grid_intersect(const rto_scene *scene, const float *origin, const float *dir, int variant,
                       float *t, float *u, float *v, uint32_t *tri_idx, rto_counters *cnt);

/* a7 + a8 helpers: renderer.cpp:107-121, triangle.h:158-161, renderer.cpp:124-133, lin_alg.h:125-132 */
void     rto_shade_hit(const rto_scene *scene, uint32_t tri_idx, float u, float v, float *rgb);
uint32_t rto_resolve_pixel(const float *rgb_sum, uint32_t spp, int gamma);

/* a1: renderer.cpp:43-136 over rows [y_begin, y_end) of a width x height frame, multi-threaded
 * over rows.  bgra: (y_end-y_begin) x width.  hit_tri/hit_t/hit_u/hit_v (each optional):
 * ((y-y_begin)*width + x)*spp + smp.  cnt optional (summed over threads). */
void rto_render_rows(const rto_scene *scene, const float *cam16, float fov_deg, uint32_t width,
                     uint32_t height, uint32_t spp, int variant, int gamma, uint32_t y_begin,
                     uint32_t y_end, uint32_t n_threads, uint32_t *bgra, uint32_t *hit_tri,
                     float *hit_t, float *hit_u, float *hit_v, rto_counters *cnt);

/* The alternates the reference keeps next to its live lines: the orthographic camera (camera.h:25-36) and
 * the two commented-out shading lines of the pixel loop, "Vec3f n = tri.n" (renderer.cpp:116) and
 * "col += Vec3f(t / 3)" (renderer.cpp:118) */
enum { RTO_SHADE_NORMAL = 0 /* interpolated vertex normal (live) */, RTO_SHADE_FACE_NORMAL = 1, RTO_SHADE_DEPTH = 2 };
typedef struct rto_render_options
{
    int   ortho;        /* 1: orthographic camera of width ortho_width (fov_deg ignored) */
    float ortho_width;
    int   shade_mode;   /* RTO_SHADE_* */
} rto_render_options;
void rto_render_rows_ex(const rto_scene *scene, const float *cam16, float fov_deg, uint32_t width,
                        uint32_t height, uint32_t spp, int variant, int gamma, uint32_t y_begin,
                        uint32_t y_end, uint32_t n_threads, const rto_render_options *opt, uint32_t *bgra,
                        uint32_t *hit_tri, float *hit_t, float *hit_u, float *hit_v, rto_counters *cnt);

void rto_intersect_rays(const rto_scene *scene, uint32_t n, const float *origins, const float *dirs,
                        int variant, uint32_t *tri_idx, float *t, float *u, float *v);

/* How often does glibc powf(x, 0.5f) (renderer.cpp:125-131) differ from IEEE sqrtf(x), which the
 * CUDA kernel uses?  Scans the float bit patterns [lo_bits, hi_bits); returns the number of x
 * whose two results differ, and in *byte_diff how many of those change (uchar)(c*255.0f). */
uint64_t rto_powf_vs_sqrtf(uint32_t lo_bits, uint32_t hi_bits, uint32_t n_threads, uint64_t *byte_diff);

#ifdef __cplusplus
}
#endif

#endif /* RT_ORACLE_H */

// TEST / BENCH-BASELINE INFRASTRUCTURE ONLY -- never linked into the product library.
//
// Headless C-ABI driver around the UNMODIFIED reference translation units, which are
// compiled in place from /root/reference by oracle/Makefile into oracle/_ref/libref_oracle.so.
// This file is our own code: it only *calls* the reference's public classes
//   Mesh (mesh.h:10-37), Matrix44f (lin_alg.h:235-690), Scene (scene.h:13-27),
//   Grid (grid.h:11-52), Renderer (renderer.h:11-35), Framebuffer (framebuffer.h:16-102),
//   GenerateRay (camera.h:8-47), SAMP::HammersleySequence (sampling.h:113-120)
// so that tests/ and bench.py's cpu_baseline / --impl reference legs can
//   (1) render a frame with the reference's own worker-thread tile pool and time it,
//   (2) dump the per-sample hit record (tri_idx,t,u,v) of Grid::Intersect,
//   (3) dump the reference's grid (cells -> triangle lists) and post-transform meshes,
// all of which pin oracle/rt_oracle.c (the CPU restatement) and the CUDA path.
//
// Private members are reached with the usual "#define private public" trick in THIS TU only
// (needed for: joining the worker threads without cancelling them, reading tile buffers
// without going through a BMP file, reading Grid::m_cells).  Class layout is unaffected.

#include <algorithm>
#include <array>
#include <atomic>
#include <cassert>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <limits>
#include <memory>
#include <mutex>
#include <random>
#include <sstream>
#include <string>
#include <thread>
#include <vector>
#include <unistd.h>

#define private public
#define protected public
#include "renderer.h"
#include "scene.h"
#include "grid.h"
#include "mesh.h"
#include "camera.h"
#include "sampling.h"
#include "triangle.h"
#include "timer.h"
#undef private
#undef protected

namespace
{

// The reference logs through printf (trace.cpp:11-26).  Keep the host process' stdout clean
// (bench.py must print exactly one JSON line) by pointing fd 1 at fd 2 while we are inside.
struct StdoutToStderr
{
    int saved;
    StdoutToStderr()
    {
        std::fflush(stdout);
        saved = dup(1);
        dup2(2, 1);
    }
    ~StdoutToStderr()
    {
        std::fflush(stdout);
        dup2(saved, 1);
        close(saved);
    }
};

Matrix44f MatFrom16(const float *m)
{
    Matrix44f r;
    std::memcpy(&r.m_mat[0][0], m, sizeof(float) * 16);
    return r;
}

void MatTo16(const Matrix44f& r, float *m) { std::memcpy(m, &r.m_mat[0][0], sizeof(float) * 16); }

struct RefRenderer
{
    std::unique_ptr<Renderer> renderer;
    Scene *scene; // owned by renderer
};

void JoinWorkers(Framebuffer *fb)
{
    // Like Framebuffer::KillAllWorkerThreads (framebuffer.cpp:29-41) but WITHOUT raising the
    // stop flag, i.e. wait for the frame to finish
    for (auto& th : fb->m_threads)
        if (th.joinable())
            th.join();
    fb->m_threads.clear();
}

void SampleTable(uint spp, std::vector<Vec2f>& smp_loc)
{
    // Same two calls as renderer.cpp:55-56
    smp_loc.resize(spp);
    for (uint smp = 0; smp < spp; smp++)
    {
        smp_loc[smp].x = SAMP::HammersleySequence<SAMP::ScrambleNone>(smp, 0, spp) - 0.5f;
        smp_loc[smp].y = SAMP::HammersleySequence<SAMP::ScrambleNone>(smp, 1, spp) - 0.5f;
    }
}

} // namespace

extern "C"
{

int ref_hardware_threads() { return int(std::max(1u, std::thread::hardware_concurrency())); }

// ----------------------------------------------------------------------------- matrices
void ref_mat_identity(float *out) { Matrix44f m; MatTo16(m, out); }
void ref_mat_translation(float x, float y, float z, float *out)
    { Matrix44f m; m.Translation(x, y, z); MatTo16(m, out); }
void ref_mat_scaling(float f, float *out) { Matrix44f m; m.Scaling(f); MatTo16(m, out); }
void ref_mat_rotation_x(float deg, float *out) { Matrix44f m; m.RotationX(deg); MatTo16(m, out); }
void ref_mat_rotation_y(float deg, float *out) { Matrix44f m; m.RotationY(deg); MatTo16(m, out); }
void ref_mat_rotation_z(float deg, float *out) { Matrix44f m; m.RotationZ(deg); MatTo16(m, out); }
void ref_mat_multiply(const float *a, const float *b, float *out)
    { Matrix44f m = MatFrom16(a) * MatFrom16(b); MatTo16(m, out); }
int ref_mat_invert(const float *a, float *out)
    { Matrix44f m = MatFrom16(a); bool ok = m.Invert(); MatTo16(m, out); return ok ? 1 : 0; }
void ref_mat_look_at(const float *eye, const float *at, float *out)
{
    Matrix44f m;
    m.BuildLookAtMatrix(Vec3f(eye), Vec3f(at));
    MatTo16(m, out);
}

// ----------------------------------------------------------------------------- meshes
void * ref_mesh_new() { return new Mesh(); }
void ref_mesh_free(void *m) { delete static_cast<Mesh *>(m); }
int ref_mesh_read(void *m, const char *path, int flip_winding)
{
    StdoutToStderr quiet;
    return static_cast<Mesh *>(m)->Read(path, flip_winding != 0) ? 1 : 0;
}
void ref_mesh_set(void *m, const float *vtx6, uint32 num_vtx, const uint32 *tri6, uint32 num_tri)
{
    // Fill the public members directly (mesh.h:26-27); both records are 24 bytes
    static_assert(sizeof(Mesh::Vertex) == 24 && sizeof(Mesh::Triangle) == 24, "layout");
    Mesh *mesh = static_cast<Mesh *>(m);
    mesh->m_vertices.resize(num_vtx);
    mesh->m_triangles.resize(num_tri);
    std::memcpy(static_cast<void *>(mesh->m_vertices.data()), vtx6, size_t(num_vtx) * 24);
    std::memcpy(static_cast<void *>(mesh->m_triangles.data()), tri6, size_t(num_tri) * 24);
}
uint32 ref_mesh_num_vertices(void *m) { return uint32(static_cast<Mesh *>(m)->m_vertices.size()); }
uint32 ref_mesh_num_triangles(void *m) { return uint32(static_cast<Mesh *>(m)->m_triangles.size()); }
void ref_mesh_get(void *m, float *vtx6, uint32 *tri6)
{
    Mesh *mesh = static_cast<Mesh *>(m);
    std::memcpy(vtx6, mesh->m_vertices.data(), mesh->m_vertices.size() * 24);
    std::memcpy(tri6, mesh->m_triangles.data(), mesh->m_triangles.size() * 24);
}
void ref_mesh_cornell_box(void *m) { static_cast<Mesh *>(m)->CornellBox(); }
void ref_mesh_normalize_dimensions(void *m) { static_cast<Mesh *>(m)->NormalizeDimensions(); }
void ref_mesh_transform(void *m, const float *mat16) { static_cast<Mesh *>(m)->Transform(MatFrom16(mat16)); }
void ref_mesh_add_mesh(void *m, void *other)
    { static_cast<Mesh *>(m)->AddMesh(* static_cast<Mesh *>(other)); }
void ref_mesh_add_quad(void *m, const float *quad12) { static_cast<Mesh *>(m)->AddQuad(quad12); }
void ref_mesh_compute_aabb(void *m, float *mn, float *mx)
{
    Vec3f a, b;
    static_cast<Mesh *>(m)->ComputeAABB(a, b);
    for (int i = 0; i < 3; i++) { mn[i] = a[i]; mx[i] = b[i]; }
}

// Instanced soup (config C5): base mesh copied n times, each copy transformed by
// Scaling(s) * RotationY(ry) * RotationX(rx) * Translation(tx,ty,tz) through the reference's
// own Matrix44f / Mesh::Transform / Mesh::AddMesh.  params = n x {s, ry, rx, tx, ty, tz}
void ref_mesh_add_instances(void *m, void *base, uint32 n, const float *params)
{
    Mesh *mesh = static_cast<Mesh *>(m);
    const Mesh *b = static_cast<Mesh *>(base);
    mesh->m_vertices.reserve(mesh->m_vertices.size() + size_t(n) * b->m_vertices.size());
    mesh->m_triangles.reserve(mesh->m_triangles.size() + size_t(n) * b->m_triangles.size());
    for (uint32 i = 0; i < n; i++)
    {
        const float *p = params + size_t(i) * 6;
        Mesh inst = * b;
        Matrix44f sc, ry, rx, tr;
        sc.Scaling(p[0]);
        ry.RotationY(p[1]);
        rx.RotationX(p[2]);
        tr.Translation(p[3], p[4], p[5]);
        inst.Transform(sc * ry * rx * tr);
        mesh->AddMesh(inst);
    }
}

// ----------------------------------------------------------------------------- scene / grid
// Takes ownership of the mesh handle.  grid_res == 64 goes through the reference's own
// Scene ctor (scene.cpp:6-10, hard-coded 64); any other value builds the Scene around a
// one-triangle placeholder and then move-assigns Grid(mesh, grid_res) (grid.cpp:12) into it,
// so that every reference TU stays unmodified.
void * ref_renderer_new(void *mesh_handle, float fov, const float *cam16, uint32 grid_res)
{
    StdoutToStderr quiet;
    std::unique_ptr<Mesh> mesh(static_cast<Mesh *>(mesh_handle));
    std::unique_ptr<Scene> scene;
    if (grid_res == 64)
        scene.reset(new Scene(std::move(mesh), fov, MatFrom16(cam16)));
    else
    {
        std::unique_ptr<Mesh> dummy(new Mesh());
        const float quad[12] = { 0, 0, 0,  1, 0, 0,  1, 1, 0,  0, 1, 0 };
        dummy->AddQuad(quad);
        scene.reset(new Scene(std::move(dummy), fov, MatFrom16(cam16)));
        scene->m_grid = Grid(std::move(mesh), grid_res);
    }
    RefRenderer *r = new RefRenderer();
    r->scene = scene.get();
    r->renderer.reset(new Renderer(std::move(scene)));
    return r;
}

void ref_renderer_free(void *h)
{
    StdoutToStderr quiet;
    delete static_cast<RefRenderer *>(h);
}

void ref_renderer_set_threads(void *h, uint32 n)
{
    // m_num_cpus is a const member initialised at run time (framebuffer.cpp:9-10)
    Framebuffer *fb = static_cast<RefRenderer *>(h)->renderer.get();
    const_cast<uint&>(fb->m_num_cpus) = std::max(1u, n);
}

uint32 ref_renderer_get_threads(void *h)
    { return static_cast<RefRenderer *>(h)->renderer->m_num_cpus; }

// Render one frame through the reference's own tile pool: SetSampleCount -> Resize /
// StartRendering -> (join) -> copy tiles like SaveToBMP (framebuffer.cpp:195-221).
// bgra (optional) receives width*height pixels, row 0 = y 0.  Returns seconds from the
// Resize/StartRendering call to the last worker thread being done.
double ref_renderer_render(void *h, uint32 width, uint32 height, uint32 spp, uint32 *bgra)
{
    StdoutToStderr quiet;
    Renderer *r = static_cast<RefRenderer *>(h)->renderer.get();
    JoinWorkers(r);
    r->SetSampleCount(spp);

    const auto t0 = std::chrono::steady_clock::now();
    if (r->m_width != width || r->m_height != height)
        r->Resize(width, height); // starts rendering (framebuffer.cpp:94-122)
    else
        r->StartRendering();      // framebuffer.cpp:124-134
    JoinWorkers(r);
    const auto t1 = std::chrono::steady_clock::now();

    if (bgra != nullptr)
        for (auto& tile : r->m_tiles)
        {
            uint x0, y0, x1, y1;
            tile.GetPosition(x0, y0, x1, y1);
            const uint32 *buf = tile.GetBuffer();
            for (uint y = 0; y < tile.GetHeight(); y++)
                std::memcpy(&bgra[x0 + size_t(y0 + y) * width], &buf[size_t(y) * tile.GetWidth()],
                            sizeof(uint32) * tile.GetWidth());
        }
    return std::chrono::duration<double>(t1 - t0).count();
}

// Render only every `stride`-th tile (tile_idx % stride == offset) of the 12x9 layout with
// n_threads host threads calling the reference's RenderTile directly.  Used to bound the CPU
// time of very large frames.  Returns seconds; *tiles_done = tiles rendered,
// *pixels_done = pixels in those tiles.
double ref_renderer_render_tile_subset(void *h, uint32 width, uint32 height, uint32 spp,
                                       uint32 stride, uint32 offset, uint32 n_threads,
                                       uint32 *tiles_done, uint64 *pixels_done)
{
    StdoutToStderr quiet;
    Renderer *r = static_cast<RefRenderer *>(h)->renderer.get();
    JoinWorkers(r);
    r->SetSampleCount(spp);
    if (r->m_width != width || r->m_height != height)
    {
        // Resize() would start the full pool; park it immediately by having zero work left
        r->Resize(width, height);
        r->m_threads_stop = true;
        JoinWorkers(r);
        r->m_threads_stop = false;
    }
    std::vector<uint> queue;
    uint64 pix = 0;
    for (uint i = 0; i < uint(r->m_tiles.size()); i++)
        if (i % stride == offset)
        {
            queue.push_back(i);
            pix += uint64(r->m_tiles[i].GetWidth()) * r->m_tiles[i].GetHeight();
        }
    std::atomic<uint> next(0);
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> pool;
    for (uint t = 0; t < std::max(1u, n_threads); t++)
        pool.emplace_back([&]()
        {
            for (;;)
            {
                const uint i = next.fetch_add(1);
                if (i >= queue.size())
                    break;
                r->RenderTile(r->m_tiles[queue[i]]);
            }
        });
    for (auto& th : pool)
        th.join();
    const auto t1 = std::chrono::steady_clock::now();
    if (tiles_done) * tiles_done = uint32(queue.size());
    if (pixels_done) * pixels_done = pix;
    return std::chrono::duration<double>(t1 - t0).count();
}

// Per-sample hit records for rows [y_begin, y_end): GenerateRay (camera.h:8-47) with the
// renderer.cpp:49-60 sample table, then the public Grid::Intersect (grid.cpp:159-281).
// Arrays are indexed ((y - y_begin) * width + x) * spp + smp; tri_idx = 0xFFFFFFFF on a miss
// (t,u,v = 0 then).  t/u/v may be null.
void ref_trace_hits(void *h, uint32 width, uint32 height, uint32 spp, uint32 y_begin, uint32 y_end,
                    uint32 n_threads, uint32 *tri_idx, float *t_out, float *u_out, float *v_out)
{
    RefRenderer *rr = static_cast<RefRenderer *>(h);
    Matrix44f cam_mat;
    float fov;
    rr->scene->GetCameraParameters(fov, cam_mat);
    const Grid *grid = rr->scene->GetGrid();
    std::vector<Vec2f> smp_loc;
    SampleTable(spp, smp_loc);

    std::atomic<uint> next_row(y_begin);
    std::vector<std::thread> pool;
    for (uint th = 0; th < std::max(1u, n_threads); th++)
        pool.emplace_back([&]()
        {
            for (;;)
            {
                const uint y = next_row.fetch_add(1);
                if (y >= y_end)
                    break;
                for (uint x = 0; x < width; x++)
                    for (uint smp = 0; smp < spp; smp++)
                    {
                        Vec3f origin, dir;
                        GenerateRay(cam_mat, Vec2ui(x, y), width, height, smp_loc[smp], false, fov,
                                    origin, dir);
                        float t, u, v;
                        uint32 idx;
                        const bool hit = grid->Intersect(origin, dir, t, u, v, idx);
                        const size_t o = (size_t(y - y_begin) * width + x) * spp + smp;
                        tri_idx[o] = hit ? idx : 0xFFFFFFFFu;
                        if (t_out) t_out[o] = hit ? t : 0.0f;
                        if (u_out) u_out[o] = hit ? u : 0.0f;
                        if (v_out) v_out[o] = hit ? v : 0.0f;
                    }
            }
        });
    for (auto& t : pool)
        t.join();
}

// Arbitrary rays through Grid::Intersect.  origins/dirs are n x 3 floats
void ref_intersect_rays(void *h, uint32 n, const float *origins, const float *dirs,
                        uint32 *tri_idx, float *t_out, float *u_out, float *v_out)
{
    const Grid *grid = static_cast<RefRenderer *>(h)->scene->GetGrid();
    for (uint32 i = 0; i < n; i++)
    {
        float t, u, v;
        uint32 idx;
        const bool hit = grid->Intersect(Vec3f(origins + 3 * i), Vec3f(dirs + 3 * i), t, u, v, idx);
        tri_idx[i] = hit ? idx : 0xFFFFFFFFu;
        t_out[i] = hit ? t : 0.0f;
        u_out[i] = hit ? u : 0.0f;
        v_out[i] = hit ? v : 0.0f;
    }
}

// Renderer::IntersectBruteForce (renderer.cpp:157-197): the author's own cross-check of the grid
void ref_intersect_rays_brute_force(void *h, uint32 n, const float *origins, const float *dirs,
                                    uint32 *tri_idx, float *t_out, float *u_out, float *v_out)
{
    Renderer *r = static_cast<RefRenderer *>(h)->renderer.get();
    for (uint32 i = 0; i < n; i++)
    {
        float t = 0.0f, u = 0.0f, v = 0.0f;
        uint32 idx = 0xFFFFFFFFu;
        const bool hit = r->IntersectBruteForce(Vec3f(origins + 3 * i), Vec3f(dirs + 3 * i), t, u, v, idx);
        tri_idx[i] = hit ? idx : 0xFFFFFFFFu;
        t_out[i] = hit ? t : 0.0f;
        u_out[i] = hit ? u : 0.0f;
        v_out[i] = hit ? v : 0.0f;
    }
}

// Renderer::RayMarch (renderer.cpp:24-41): sphere tracing against DistanceBruteForce (:138-155)
void ref_ray_march(void *h, uint32 n, const float *origins, const float *dirs, uint32 *hit_out, float *t_out)
{
    Renderer *r = static_cast<RefRenderer *>(h)->renderer.get();
    for (uint32 i = 0; i < n; i++)
    {
        float t = 0.0f;
        hit_out[i] = r->RayMarch(Vec3f(origins + 3 * i), Vec3f(dirs + 3 * i), t) ? 1u : 0u;
        t_out[i] = t;
    }
}

// Primary rays exactly as RenderTile generates them (for ray-generation parity)
void ref_generate_rays(void *h, uint32 width, uint32 height, uint32 spp, uint32 y_begin,
                       uint32 y_end, float *origins, float *dirs)
{
    RefRenderer *rr = static_cast<RefRenderer *>(h);
    Matrix44f cam_mat;
    float fov;
    rr->scene->GetCameraParameters(fov, cam_mat);
    std::vector<Vec2f> smp_loc;
    SampleTable(spp, smp_loc);
    for (uint y = y_begin; y < y_end; y++)
        for (uint x = 0; x < width; x++)
            for (uint smp = 0; smp < spp; smp++)
            {
                Vec3f origin, dir;
                GenerateRay(cam_mat, Vec2ui(x, y), width, height, smp_loc[smp], false, fov, origin, dir);
                const size_t o = ((size_t(y - y_begin) * width + x) * spp + smp) * 3;
                for (int i = 0; i < 3; i++) { origins[o + i] = origin[i]; dirs[o + i] = dir[i]; }
            }
}

// The same through the orthographic branch of GenerateRay (camera.h:25-36), which RenderTile never takes
void ref_generate_rays_ortho(void *h, uint32 width, uint32 height, uint32 spp, uint32 y_begin, uint32 y_end,
                             float ortho_width, float *origins, float *dirs)
{
    RefRenderer *rr = static_cast<RefRenderer *>(h);
    Matrix44f cam_mat;
    float fov;
    rr->scene->GetCameraParameters(fov, cam_mat);
    std::vector<Vec2f> smp_loc;
    SampleTable(spp, smp_loc);
    for (uint y = y_begin; y < y_end; y++)
        for (uint x = 0; x < width; x++)
            for (uint smp = 0; smp < spp; smp++)
            {
                Vec3f origin, dir;
                GenerateRay(cam_mat, Vec2ui(x, y), width, height, smp_loc[smp], true, ortho_width, origin, dir);
                const size_t o = ((size_t(y - y_begin) * width + x) * spp + smp) * 3;
                for (int i = 0; i < 3; i++) { origins[o + i] = origin[i]; dirs[o + i] = dir[i]; }
            }
}

void ref_sample_table(uint32 spp, float *xy)
{
    std::vector<Vec2f> smp_loc;
    SampleTable(spp, smp_loc);
    for (uint i = 0; i < spp; i++) { xy[2 * i] = smp_loc[i].x; xy[2 * i + 1] = smp_loc[i].y; }
}

// Grid parameters: dim[3], aabb_min[3], aabb_max[3], cell_wdh, inv_cell_wdh; returns the
// total number of (cell, triangle) references
uint64 ref_grid_info(void *h, uint32 *dim, float *aabb_min, float *aabb_max, float *cell_wdh,
                     float *inv_cell_wdh)
{
    const Grid *g = static_cast<RefRenderer *>(h)->scene->GetGrid();
    for (int i = 0; i < 3; i++)
    {
        dim[i] = g->m_grid_dim[i];
        aabb_min[i] = g->m_aabb_min[i];
        aabb_max[i] = g->m_aabb_max[i];
    }
    * cell_wdh = g->m_cell_wdh;
    * inv_cell_wdh = g->m_inv_cell_wdh;
    uint64 refs = 0;
    for (const auto& c : g->m_cells)
        refs += c.m_isect_tri_idx.size();
    return refs;
}

// Flatten Grid::m_cells (grid.h:33-39) to CSR in the reference's own cell order
// (GridIdx = x + z*dimx + y*dimx*dimz, grid.h:41-42): cell_offset has cells+1 entries
void ref_grid_dump(void *h, uint64 *cell_offset, uint32 *tri_index)
{
    const Grid *g = static_cast<RefRenderer *>(h)->scene->GetGrid();
    uint64 o = 0;
    for (size_t c = 0; c < g->m_cells.size(); c++)
    {
        cell_offset[c] = o;
        for (uint32 idx : g->m_cells[c].m_isect_tri_idx)
            tri_index[o++] = idx;
    }
    cell_offset[g->m_cells.size()] = o;
}

// The (post-transform) mesh the reference grid owns
uint32 ref_scene_num_vertices(void *h)
    { return uint32(static_cast<RefRenderer *>(h)->scene->GetGrid()->GetMesh()->m_vertices.size()); }
uint32 ref_scene_num_triangles(void *h)
    { return uint32(static_cast<RefRenderer *>(h)->scene->GetGrid()->GetMesh()->m_triangles.size()); }
void ref_scene_get_mesh(void *h, float *vtx6, uint32 *tri6)
{
    const Mesh *mesh = static_cast<RefRenderer *>(h)->scene->GetGrid()->GetMesh();
    std::memcpy(vtx6, mesh->m_vertices.data(), mesh->m_vertices.size() * 24);
    std::memcpy(tri6, mesh->m_triangles.data(), mesh->m_triangles.size() * 24);
}

// The frame constants GenerateRay derives per ray (camera.h:24,41-42), computed by this
// toolchain: aspect and fov_xs = float(tan(double(DegToRad(fov) / 2)))
void ref_camera_constants(float fov, uint32 width, uint32 height, float *fov_xs, float *aspect)
{
    const float hfov = DegToRad(fov);
    * fov_xs = tan(hfov / 2);
    * aspect = float(width) / float(height);
}

// One ray/triangle test through the reference's own functions: variant 0 = IntersectRayTri
// (triangle.h:15-107), 1 = IntersectRayTriBarycentric (triangle.h:210-226, needs the face
// normal n).  tuv receives t,u,v as the function left them
int ref_tri_test(int variant, const float *o, const float *d, const float *v0, const float *v1,
                 const float *v2, const float *n, float *tuv)
{
    float t = 0.0f, u = 0.0f, v = 0.0f;
    bool hit;
    if (variant == 0)
        hit = IntersectRayTri(Vec3f(o), Vec3f(d), Vec3f(v0), Vec3f(v1), Vec3f(v2), t, u, v);
    else
        hit = IntersectRayTriBarycentric(Vec3f(o), Vec3f(d), Vec3f(v0), Vec3f(v1), Vec3f(v2), Vec3f(n),
                                         t, u, v);
    tuv[0] = t; tuv[1] = u; tuv[2] = v;
    return hit ? 1 : 0;
}

// ----------------------------------------------------------------------------- sampling module
// kind: 0 Halton, 1 Hammersley, 2 HaltonZaremba, 3 HammersleyZaremba, 4 RadicalInverseBase2,
//       5 SobolRadicalInverseBase2, 6 LarcherPillichshammerRadicalInverseBase2
// scramble (kinds 0,1): 0 none, 1 Braaten-Weller, 2 Faure, 3 reverse, 4 randomized
// out[i * dim_count + j] = sequence(n_begin + i, dim_begin + j)     (sampling.h:91-120, sampling.cpp:194-281)
void ref_qmc_sequence(uint32 kind, uint32 scramble, uint32 n_begin, uint32 count, uint32 dim_begin,
                      uint32 dim_count, uint32 num_smp, uint32 bits, double *out)
{
    static bool init = false;
    if (!init)
    {
        SAMP::Initialize();
        init = true;
    }
    for (uint32 i = 0; i < count; i++)
        for (uint32 j = 0; j < dim_count; j++)
        {
            const uint32 n = n_begin + i, dim = dim_begin + j;
            double v = 0.0;
            switch (kind)
            {
                case 0:
                    v = scramble == 0 ? SAMP::HaltonSequence<SAMP::ScrambleNone>(n, dim)
                      : scramble == 1 ? SAMP::HaltonSequence<SAMP::ScrambleBraatenWeller>(n, dim)
                      : scramble == 2 ? SAMP::HaltonSequence<SAMP::ScrambleFaure>(n, dim)
                      : scramble == 3 ? SAMP::HaltonSequence<SAMP::ScrambleReverse>(n, dim)
                      :                 SAMP::HaltonSequence<SAMP::ScrambleRandomized>(n, dim);
                    break;
                case 1:
                    v = scramble == 0 ? SAMP::HammersleySequence<SAMP::ScrambleNone>(n, dim, num_smp)
                      : scramble == 1 ? SAMP::HammersleySequence<SAMP::ScrambleBraatenWeller>(n, dim, num_smp)
                      : scramble == 2 ? SAMP::HammersleySequence<SAMP::ScrambleFaure>(n, dim, num_smp)
                      : scramble == 3 ? SAMP::HammersleySequence<SAMP::ScrambleReverse>(n, dim, num_smp)
                      :                 SAMP::HammersleySequence<SAMP::ScrambleRandomized>(n, dim, num_smp);
                    break;
                case 2: v = SAMP::HaltonZarembaSequence(n, dim); break;
                case 3: v = SAMP::HammersleyZarembaSequence(n, dim, num_smp); break;
                case 4: v = SAMP::RadicalInverseBase2(n, bits); break;
                case 5: v = SAMP::SobolRadicalInverseBase2(n, bits); break;
                default: v = SAMP::LarcherPillichshammerRadicalInverseBase2(n, bits); break;
            }
            out[size_t(i) * dim_count + j] = v;
        }
}

// Permutation table of one prime index: scramble 1 BW (16 primes), 2 Faure, 3 reverse, 4 randomized
// (128 primes each).  Returns the table length (= the prime) or 0 if there is none.
uint32 ref_qmc_table(uint32 scramble, uint32 prime_idx, uint32 *out)
{
    static bool init = false;
    if (!init)
    {
        SAMP::Initialize();
        init = true;
    }
    const std::vector<uint> *tbl = nullptr;
    if (scramble == 1 && prime_idx < SAMP::BW_TBL_SIZE) tbl = &SAMP::g_braaten_weller_table[prime_idx];
    if (scramble == 2 && prime_idx < SAMP::FAURE_TBL_SIZE) tbl = &SAMP::g_faure_table[prime_idx];
    if (scramble == 3 && prime_idx < SAMP::REVERSE_TBL_SIZE) tbl = &SAMP::g_reverse_table[prime_idx];
    if (scramble == 4 && prime_idx < SAMP::RANDOMIZED_TBL_SIZE) tbl = &SAMP::g_randomized_table[prime_idx];
    if (!tbl)
        return 0;
    for (size_t i = 0; i < tbl->size(); i++)
        out[i] = (*tbl)[i];
    return uint32(tbl->size());
}

uint32 ref_qmc_prime(uint32 idx) { return idx < SAMP::PRIME_TBL_SIZE ? SAMP::g_prime_table[idx] : 0; }
double ref_qmc_cranley_patterson(double x, double e) { return SAMP::CranleyPattersonRotation(x, e); }

void ref_save_bmp(void *h, const char *filename)
{
    StdoutToStderr quiet;
    static_cast<RefRenderer *>(h)->renderer->SaveToBMP(filename);
}

} // extern "C"

// Flat C wrappers over the host classes (Mesh, Matrix44f, Scene, Renderer) so that Python tests,
// bench.py and other FFI hosts can drive the drop-in API.  Function names and signatures mirror
// oracle/ref_driver.cpp's ref_* set one to one (prefix rtm_), which lets one scene recipe run
// against either implementation.
#include <chrono>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>

#include "../../include/cuda_trace.h"
#include "bmp_writer.h"
#include "camera.h"
#include "mesh.h"
#include "renderer.h"
#include "scene.h"

namespace
{

std::string g_error;

Matrix44f From16(const float *m)
{
    Matrix44f r;
    std::memcpy(r.m_mat, m, sizeof(r.m_mat));
    return r;
}
void To16(const Matrix44f& r, float *m) { std::memcpy(m, r.m_mat, sizeof(r.m_mat)); }

struct HostRenderer
{
    std::unique_ptr<Renderer> renderer;
};

} // namespace

extern "C"
{

const char *rtm_last_error() { return g_error.c_str(); }

// ---- matrices
void rtm_mat_identity(float *out) { Matrix44f m; To16(m, out); }
void rtm_mat_translation(float x, float y, float z, float *out) { Matrix44f m; m.Translation(x, y, z); To16(m, out); }
void rtm_mat_scaling(float f, float *out) { Matrix44f m; m.Scaling(f); To16(m, out); }
void rtm_mat_rotation_x(float deg, float *out) { Matrix44f m; m.RotationX(deg); To16(m, out); }
void rtm_mat_rotation_y(float deg, float *out) { Matrix44f m; m.RotationY(deg); To16(m, out); }
void rtm_mat_rotation_z(float deg, float *out) { Matrix44f m; m.RotationZ(deg); To16(m, out); }
void rtm_mat_multiply(const float *a, const float *b, float *out) { To16(From16(a) * From16(b), out); }
int rtm_mat_invert(const float *a, float *out)
{
    Matrix44f m = From16(a);
    const bool ok = m.Invert();
    To16(m, out);
    return ok ? 1 : 0;
}
void rtm_mat_look_at(const float *eye, const float *at, float *out)
{
    Matrix44f m;
    m.BuildLookAtMatrix(Vec3f(eye), Vec3f(at));
    To16(m, out);
}
void rtm_camera_constants(float fov, uint32 width, uint32 height, float *fov_xs, float *aspect)
{
    CameraFrameConstants(fov, width, height, *fov_xs, *aspect);
}

// ---- meshes
void *rtm_mesh_new() { return new Mesh(); }
void rtm_mesh_free(void *m) { delete static_cast<Mesh *>(m); }
int rtm_mesh_read(void *m, const char *path, int flip) { return static_cast<Mesh *>(m)->Read(path, flip != 0) ? 1 : 0; }
int rtm_mesh_read_binary(void *m, const char *path) { return static_cast<Mesh *>(m)->ReadBinary(path) ? 1 : 0; }
void rtm_mesh_set(void *m, const float *vtx6, uint32 nv, const uint32 *tri6, uint32 nt)
{
    static_cast<Mesh *>(m)->SetArrays(vtx6, nv, tri6, nt);
}
uint32 rtm_mesh_num_vertices(void *m) { return uint32(static_cast<Mesh *>(m)->m_vertices.size()); }
uint32 rtm_mesh_num_triangles(void *m) { return uint32(static_cast<Mesh *>(m)->m_triangles.size()); }
void rtm_mesh_get(void *m, float *vtx6, uint32 *tri6)
{
    const Mesh *mesh = static_cast<Mesh *>(m);
    std::memcpy(vtx6, mesh->m_vertices.data(), mesh->m_vertices.size() * sizeof(Mesh::Vertex));
    std::memcpy(tri6, mesh->m_triangles.data(), mesh->m_triangles.size() * sizeof(Mesh::Triangle));
}
void rtm_mesh_cornell_box(void *m) { static_cast<Mesh *>(m)->CornellBox(); }
void rtm_mesh_normalize_dimensions(void *m) { static_cast<Mesh *>(m)->NormalizeDimensions(); }
void rtm_mesh_transform(void *m, const float *mat16) { static_cast<Mesh *>(m)->Transform(From16(mat16)); }
void rtm_mesh_add_mesh(void *m, void *other) { static_cast<Mesh *>(m)->AddMesh(*static_cast<Mesh *>(other)); }
void rtm_mesh_add_quad(void *m, const float *quad12) { static_cast<Mesh *>(m)->AddQuad(quad12); }
void rtm_mesh_compute_aabb(void *m, float *mn, float *mx)
{
    Vec3f a, b;
    static_cast<Mesh *>(m)->ComputeAABB(a, b);
    for (int i = 0; i < 3; i++)
    {
        mn[i] = a[i];
        mx[i] = b[i];
    }
}
// n copies of `base`, copy i transformed by Scaling(s) * RotationY(ry) * RotationX(rx) *
// Translation(t) with params = n x {s, ry, rx, tx, ty, tz} (the synthetic soup of config C5)
void rtm_mesh_add_instances(void *m, void *base, uint32 n, const float *params)
{
    Mesh *mesh = static_cast<Mesh *>(m);
    const Mesh *b = static_cast<Mesh *>(base);
    mesh->m_vertices.reserve(mesh->m_vertices.size() + size_t(n) * b->m_vertices.size());
    mesh->m_triangles.reserve(mesh->m_triangles.size() + size_t(n) * b->m_triangles.size());
    for (uint32 i = 0; i < n; i++)
    {
        const float *p = params + size_t(i) * 6;
        Matrix44f sc, ry, rx, tr;
        sc.Scaling(p[0]);
        ry.RotationY(p[1]);
        rx.RotationX(p[2]);
        tr.Translation(p[3], p[4], p[5]);
        Mesh inst = *b;
        inst.Transform(sc * ry * rx * tr);
        mesh->AddMesh(inst);
    }
}

int rtm_write_bitmap(const char *filename, uint32 width, uint32 height, const uint32 *bgra)
{
    return WriteBitmap(filename, width, height, bgra) ? 1 : 0;
}

// ---- renderer (Mesh -> Scene -> Renderer; takes ownership of the mesh handle)
void *rtm_renderer_new(void *mesh_handle, float fov, const float *cam16, uint32 grid_res, int n_gpus)
{
    std::unique_ptr<Mesh> mesh(static_cast<Mesh *>(mesh_handle));
    try
    {
        Grid::SetDeviceCount(n_gpus > 0 ? n_gpus : 1);
        std::unique_ptr<Scene> scene(new Scene(std::move(mesh), fov, From16(cam16), grid_res));
        HostRenderer *h = new HostRenderer();
        h->renderer.reset(new Renderer(std::move(scene)));
        return h;
    }
    catch (const std::exception& e)
    {
        g_error = e.what();
        return nullptr;
    }
}
void rtm_renderer_free(void *h) { delete static_cast<HostRenderer *>(h); }

// SetSampleCount -> Resize / StartRendering -> WaitRendering -> copy tiles.  Returns seconds from
// the Resize/StartRendering call until the frame is in the tiles, < 0 on error
double rtm_renderer_render(void *h, uint32 width, uint32 height, uint32 spp, uint32 variant, int gamma, uint32 *bgra)
{
    Renderer *r = static_cast<HostRenderer *>(h)->renderer.get();
    try
    {
        r->WaitRendering();
        r->SetSampleCount(spp);
        r->SetIntersectVariant(variant);
        r->SetGammaCorrection(gamma != 0);
        const auto t0 = std::chrono::steady_clock::now();
        if (r->GetWidth() != width || r->GetHeight() != height)
            r->Resize(width, height);
        else
            r->StartRendering();
        if (!r->WaitRendering())
        {
            g_error = r->GetLastError();
            return -1.0;
        }
        const auto t1 = std::chrono::steady_clock::now();
        if (bgra)
            r->CopyToBitmap(bgra);
        return std::chrono::duration<double>(t1 - t0).count();
    }
    catch (const std::exception& e)
    {
        g_error = e.what();
        return -1.0;
    }
}
// The asynchronous form a viewer uses: start (returns at once), poll / save screenshots, wait
int rtm_renderer_start(void *h, uint32 width, uint32 height, uint32 spp)
{
    Renderer *r = static_cast<HostRenderer *>(h)->renderer.get();
    try
    {
        r->WaitRendering();
        r->SetSampleCount(spp);
        if (r->GetWidth() != width || r->GetHeight() != height)
            r->Resize(width, height);
        else
            r->StartRendering();
        return 0;
    }
    catch (const std::exception& e)
    {
        g_error = e.what();
        return -1;
    }
}
int rtm_renderer_wait(void *h)
{
    Renderer *r = static_cast<HostRenderer *>(h)->renderer.get();
    if (r->WaitRendering())
        return 0;
    g_error = r->GetLastError();
    return -1;
}
void rtm_renderer_stop(void *h) { static_cast<HostRenderer *>(h)->renderer->StopRendering(); }
uint32 rtm_renderer_finished_tiles(void *h) { return static_cast<HostRenderer *>(h)->renderer->CountFinishedTiles(); }
double rtm_renderer_last_render_seconds(void *h) { return static_cast<HostRenderer *>(h)->renderer->GetLastRenderSeconds(); }
void rtm_renderer_copy_bitmap(void *h, uint32 *bgra) { static_cast<HostRenderer *>(h)->renderer->CopyToBitmap(bgra); }
void rtm_renderer_set_alternates(void *h, float ortho_width, uint32 shade_mode)
{
    Renderer *r = static_cast<HostRenderer *>(h)->renderer.get();
    r->WaitRendering();
    r->SetOrthographicWidth(ortho_width);
    r->SetShadingMode(shade_mode);
}
float rtm_renderer_last_kernel_ms(void *h) { return static_cast<HostRenderer *>(h)->renderer->GetLastKernelMilliseconds(); }
void rtm_renderer_save_bmp(void *h, const char *filename) { static_cast<HostRenderer *>(h)->renderer->SaveToBMP(filename); }
void *rtm_renderer_device_context(void *h)
{
    return static_cast<HostRenderer *>(h)->renderer->GetScene()->GetGrid()->GetDeviceContext();
}
int rtm_renderer_intersect(void *h, const float *origin, const float *dir, float *tuv, uint32 *tri_idx)
{
    try
    {
        return static_cast<HostRenderer *>(h)->renderer->GetScene()->GetGrid()->Intersect(
                   Vec3f(origin), Vec3f(dir), tuv[0], tuv[1], tuv[2], *tri_idx) ? 1 : 0;
    }
    catch (const std::exception& e)
    {
        g_error = e.what();
        return -1;
    }
}
int rtm_renderer_ray_march(void *h, const float *origin, const float *dir, float *t)
{
    return static_cast<HostRenderer *>(h)->renderer->RayMarch(Vec3f(origin), Vec3f(dir), *t) ? 1 : 0;
}
void rtm_renderer_grid_info(void *h, uint32 *dim, float *aabb_min, float *aabb_max, float *cell_wdh, uint64 *refs)
{
    const Grid *g = static_cast<HostRenderer *>(h)->renderer->GetScene()->GetGrid();
    uint d[3];
    Vec3f mn, mx;
    g->GetDimensions(d);
    g->GetAABB(mn, mx);
    for (int i = 0; i < 3; i++)
    {
        dim[i] = d[i];
        aabb_min[i] = mn[i];
        aabb_max[i] = mx[i];
    }
    *cell_wdh = g->GetCellWidth();
    *refs = g->GetReferenceCount();
}

} // extern "C"

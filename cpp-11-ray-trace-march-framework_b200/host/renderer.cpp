#include "renderer.h"

#include <algorithm>
#include <cstring>
#include <stdexcept>
#include <string>

#include "../../include/cuda_trace.h"
#include "camera.h"
#include "trace.h"

Renderer::Renderer(std::unique_ptr<Scene> scene) : m_scene(std::move(scene)) { }

Renderer::~Renderer()
{
    KillAllWorkerThreads();
    WaitRendering();
}

void Renderer::SetSampleCount(uint cnt)
{
    // the reference asserts that nothing is rendering (renderer.cpp:20); wait instead
    WaitRendering();
    m_sample_count = std::max(1u, cnt);
}

void Renderer::OnCancel()
{
    cuda_trace_cancel(m_scene->GetGrid()->GetDeviceContext());
}

void Renderer::RenderTile(Tile& tile)
{
    Tile *one = &tile;
    RenderTiles(&one, 1);
}

void Renderer::RenderTiles(Tile * const *tiles, uint count)
{
    cuda_trace_ctx *ctx = m_scene->GetGrid()->GetDeviceContext();

    // Per-frame inputs of the kernel (reference renderer.cpp:63-72,91-99)
    cuda_trace_frame frame;
    std::memset(&frame, 0, sizeof(frame));
    float fov;
    Matrix44f cam_mat;
    m_scene->GetCameraParameters(fov, cam_mat);
    frame.width = m_width;
    frame.height = m_height;
    frame.spp = m_sample_count;
    frame.variant = m_variant;
    frame.flags = m_gamma ? CUDA_TRACE_FLAG_GAMMA : 0u;
    CameraFrameConstants(fov, m_width, m_height, frame.fov_xs, frame.aspect);
    if (m_ortho_width > 0.0f)
    {
        frame.flags |= CUDA_TRACE_FLAG_ORTHO;
        frame.fov_xs = m_ortho_width; // width_or_hfov of camera.h:13
    }
    if (m_shade_mode == 1u) frame.flags |= CUDA_TRACE_FLAG_SHADE_FACE_NORMAL;
    if (m_shade_mode == 2u) frame.flags |= CUDA_TRACE_FLAG_SHADE_DEPTH;
    std::memcpy(frame.cam_mat, cam_mat.m_mat, sizeof(frame.cam_mat));

    std::vector<cuda_trace_tile_rect> rects(count);
    std::vector<uint32 *> buffers(count);
    for (uint i = 0; i < count; i++)
    {
        tiles[i]->GetPosition(rects[i].x0, rects[i].y0, rects[i].x1, rects[i].y1);
        buffers[i] = tiles[i]->GetBuffer(); // index x + y * tile_width, as the reference's kernel writes it (renderer.cpp:133)
    }

    // One launch for the whole list; every tile is copied from the device framebuffer straight into its own
    // (page-locked) buffer as soon as the rows it lies in are traced, and handed back -- unlocked -- right then
    struct Done
    {
        Renderer *self;
        Tile * const *tiles;
        static void Call(const uint32_t *idx, uint32_t n, void *user)
        {
            Done *d = static_cast<Done *>(user);
            for (uint32_t k = 0; k < n; k++)
                d->self->TileFinished(*d->tiles[idx[k]]);
        }
    } done = { this, tiles };
    const int rc = cuda_trace_tiles_into(ctx, &frame, rects.data(), count, buffers.data(), &Done::Call, &done);
    if (rc == CUDA_TRACE_ERR_CANCELLED)
        return;
    if (rc)
        throw std::runtime_error(std::string("Renderer: cuda_trace_tiles_into failed: ") + cuda_trace_last_error(ctx));
    cuda_trace_last_kernel_ms(ctx, &m_last_kernel_ms);
}

bool Renderer::RayMarch(Vec3f origin, Vec3f dir, float& t)
{
    const float o[3] = { origin.x, origin.y, origin.z }, d[3] = { dir.x, dir.y, dir.z };
    uint32 hit = 0;
    if (cuda_trace_ray_march(m_scene->GetGrid()->GetDeviceContext(), 1, o, d, &hit, &t) != 0)
        return false;
    return hit != 0;
}

#include "grid.h"

#include <cstdlib>
#include <stdexcept>
#include <string>

#include "../../include/cuda_trace.h"
#include "trace.h"

namespace
{
int g_device_count = 0; // 0 = not set: use $RTM_NUM_GPUS or 1
}

void Grid::SetDeviceCount(int n) { g_device_count = n; }

int Grid::GetDeviceCount()
{
    if (g_device_count > 0)
        return g_device_count;
    const char *env = std::getenv("RTM_NUM_GPUS");
    const int n = env ? std::atoi(env) : 1;
    return n > 0 ? n : 1;
}

Grid::Grid(std::unique_ptr<Mesh> mesh, uint grid_res)
    : m_mesh(std::move(mesh)), m_cell_wdh(0.0f), m_inv_cell_wdh(0.0f), m_num_refs(0), m_ctx(nullptr)
{
    m_grid_dim[0] = m_grid_dim[1] = m_grid_dim[2] = 0;
    if (!m_mesh || m_mesh->m_vertices.empty() || m_mesh->m_triangles.empty())
        throw std::runtime_error("Grid: empty mesh"); // reference asserts (grid.cpp:15)
    if (grid_res == kAutoResolution)
        grid_res = cuda_trace_suggest_grid_res(uint32(m_mesh->m_triangles.size()));
    if (grid_res == 0)
        throw std::runtime_error("Grid: zero resolution"); // reference asserts (grid.cpp:16)

    const double t0 = TimerGetTick();
    int rc = cuda_trace_init(GetDeviceCount(), &m_ctx);
    if (rc)
        throw std::runtime_error(std::string("Grid: cuda_trace_init failed: ") + cuda_trace_last_error(nullptr));
    rc = cuda_trace_upload_scene(m_ctx, &m_mesh->m_vertices[0].p.x, uint32(m_mesh->m_vertices.size()),
                                 &m_mesh->m_triangles[0].v0, uint32(m_mesh->m_triangles.size()), grid_res);
    cuda_trace_grid_desc desc;
    if (!rc)
        rc = cuda_trace_download_grid(m_ctx, &desc, nullptr, nullptr);
    if (rc)
    {
        const std::string msg = cuda_trace_last_error(m_ctx);
        cuda_trace_destroy(m_ctx);
        m_ctx = nullptr;
        throw std::runtime_error("Grid: device grid build failed: " + msg);
    }
    for (int k = 0; k < 3; k++)
    {
        m_grid_dim[k] = desc.dim[k];
        m_aabb_min[k] = desc.aabb_min[k];
        m_aabb_max[k] = desc.aabb_max[k];
    }
    m_cell_wdh = desc.cell_wdh;
    m_inv_cell_wdh = desc.inv_cell_wdh;
    m_num_refs = desc.num_refs;

    // same facts the reference logs (grid.cpp:43-59,142-151)
    Trace("Built %ix%ix%i grid (%llu total cells, %.3f cell width) at (%.3f, %.3f, %.3f) - (%.3f, %.3f, %.3f) "
          "for mesh with %i triangles and %i vertices on %i GPU(s) in %.3fs, %llu cell references (%.3f per cell)",
          m_grid_dim[0], m_grid_dim[1], m_grid_dim[2], (unsigned long long) desc.num_cells, m_cell_wdh,
          m_aabb_min.x, m_aabb_min.y, m_aabb_min.z, m_aabb_max.x, m_aabb_max.y, m_aabb_max.z,
          int(m_mesh->m_triangles.size()), int(m_mesh->m_vertices.size()), GetDeviceCount(),
          TimerGetTick() - t0, (unsigned long long) desc.num_refs, double(desc.num_refs) / double(desc.num_cells));
}

Grid::~Grid()
{
    if (m_ctx)
        cuda_trace_destroy(m_ctx);
}

bool Grid::Intersect(Vec3f origin, Vec3f dir, float& t, float& u, float& v, uint32& tri_idx) const
{
    uint32 idx = CUDA_TRACE_MISS;
    float ht = 0.0f, hu = 0.0f, hv = 0.0f;
    const int rc = cuda_trace_intersect_rays(m_ctx, 1, &origin.x, &dir.x, CUDA_TRACE_VARIANT_MT, &idx, &ht, &hu, &hv);
    if (rc)
        throw std::runtime_error(std::string("Grid::Intersect: ") + cuda_trace_last_error(m_ctx));
    if (idx == CUDA_TRACE_MISS)
        return false;
    t = ht;
    u = hu;
    v = hv;
    tri_idx = idx;
    return true;
}

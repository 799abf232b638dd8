// Uniform grid acceleration structure -- host handle.  Same public surface as the reference's
// Grid (grid.h:11-52: ctor from a Mesh + resolution, GetMesh, Intersect), but the structure
// itself lives in GPU memory: the constructor uploads the mesh and has the devices build the
// CSR grid and the cell-major triangle records (cuda_trace_upload_scene).  There is no host
// copy of the cells and no CPU traversal; Intersect() is a one-ray GPU query.
#ifndef RTM_HOST_GRID_H
#define RTM_HOST_GRID_H

#include <memory>

#include "lin_alg.h"
#include "mesh.h"
#include "types.h"

struct cuda_trace_ctx;

class Grid
{
public:
    // grid_res: cells along the longest axis (reference: always 64), or kAutoResolution for about
    // three cells per triangle.  Throws std::runtime_error if no CUDA device is usable / the build fails
    static const uint kAutoResolution = 0xFFFFFFFFu;
    Grid(std::unique_ptr<Mesh> mesh, uint grid_res);
    ~Grid();
    Grid(const Grid&) = delete;
    Grid& operator = (const Grid&) = delete;

    inline const Mesh * GetMesh() const { return m_mesh.get(); }
    bool Intersect(Vec3f origin, Vec3f dir, float& t, float& u, float& v, uint32& tri_idx) const;

    // --- additions -----------------------------------------------------------------------
    // Number of GPUs the next Grid (= device context) spans; default 1, or $RTM_NUM_GPUS
    static void SetDeviceCount(int n);
    static int GetDeviceCount();
    cuda_trace_ctx * GetDeviceContext() const { return m_ctx; }
    void GetDimensions(uint dim[3]) const { dim[0] = m_grid_dim[0]; dim[1] = m_grid_dim[1]; dim[2] = m_grid_dim[2]; }
    float GetCellWidth() const { return m_cell_wdh; }
    void GetAABB(Vec3f& mn, Vec3f& mx) const { mn = m_aabb_min; mx = m_aabb_max; }
    uint64 GetReferenceCount() const { return m_num_refs; }

protected:
    std::unique_ptr<Mesh> m_mesh;
    uint  m_grid_dim[3];
    float m_cell_wdh;
    float m_inv_cell_wdh;
    Vec3f m_aabb_min;
    Vec3f m_aabb_max;
    uint64 m_num_refs;
    cuda_trace_ctx *m_ctx;
};

#endif

// Scene = geometry (through its Grid) + camera, as in the reference (scene.h:13-27).  The grid
// resolution, hard-coded to 64 in the reference (scene.cpp:7), is an optional fourth argument.
#ifndef RTM_HOST_SCENE_H
#define RTM_HOST_SCENE_H

#include <memory>

#include "grid.h"
#include "lin_alg.h"
#include "types.h"

struct Mesh;

class Scene
{
public:
    Scene(std::unique_ptr<Mesh> mesh, float fov, Matrix44f cam_mat, uint grid_res = 64)
        : m_grid(std::move(mesh), grid_res), m_fov(fov), m_cam_mat(cam_mat) { }

    void GetCameraParameters(float& fov, Matrix44f& cam_mat) { fov = m_fov; cam_mat = m_cam_mat; }
    void SetCameraParameters(float fov, const Matrix44f& cam_mat) { m_fov = fov; m_cam_mat = cam_mat; }
    inline const Grid * GetGrid() const { return &m_grid; }

protected:
    Grid      m_grid;
    float     m_fov;     // horizontal field of view, degrees
    Matrix44f m_cam_mat;
};

#endif

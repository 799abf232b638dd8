// Scene = geometry (through its Grid) + camera, the object a Renderer is built from (the reference's
// scene.h:13-27 interface: constructor taking ownership of the mesh, GetCameraParameters, GetGrid).
// The grid resolution, hard-coded to 64 in the reference (scene.cpp:7), is an optional fourth argument
// (Grid::kAutoResolution picks it from the triangle density), and the camera can be replaced between frames.
#ifndef RTM_HOST_SCENE_H
#define RTM_HOST_SCENE_H

#include <memory>

#include "grid.h"
#include "lin_alg.h"
#include "types.h"

struct Mesh;

class Scene
{
    // what GenerateRay needs besides the frame size: horizontal field of view in degrees + camera-to-world matrix
    struct Camera
    {
        float     fov_deg;
        Matrix44f to_world;
    };

public:
    Scene(std::unique_ptr<Mesh> mesh, float fov, Matrix44f cam_mat, uint grid_res = 64);

    void GetCameraParameters(float& fov, Matrix44f& cam_mat);
    void SetCameraParameters(float fov, const Matrix44f& cam_mat); // addition: move the camera, keep the grid
    const Grid * GetGrid() const;

protected:
    Grid   m_grid;   // owns the mesh and its device-resident uniform grid
    Camera m_camera;
};

inline Scene::Scene(std::unique_ptr<Mesh> mesh, float fov, Matrix44f cam_mat, uint grid_res)
    : m_grid(std::move(mesh), grid_res), m_camera{ fov, cam_mat }
{
}

inline void Scene::GetCameraParameters(float& fov, Matrix44f& cam_mat)
{
    fov = m_camera.fov_deg;
    cam_mat = m_camera.to_world;
}

inline void Scene::SetCameraParameters(float fov, const Matrix44f& cam_mat)
{
    m_camera.fov_deg = fov;
    m_camera.to_world = cam_mat;
}

inline const Grid * Scene::GetGrid() const
{
    return &m_grid;
}

#endif

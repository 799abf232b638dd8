// Renderer = Framebuffer + Scene, the reference's top-level rendering object (renderer.h:11-35):
// Renderer(unique_ptr<Scene>), SetSampleCount / GetSampleCount, RenderTile override.  The per-tile
// kernel itself (reference renderer.cpp:43-136) runs on the GPU: RenderTile() is a one-tile
// cuda_trace_tiles_into launch, RenderTiles() a single launch for the whole frame that hands tiles
// back as they complete.
#ifndef RTM_HOST_RENDERER_H
#define RTM_HOST_RENDERER_H

#include <memory>
#include <vector>

#include "framebuffer.h"
#include "lin_alg.h"
#include "scene.h"

class Renderer : public Framebuffer
{
public:
    Renderer(std::unique_ptr<Scene> scene);
    ~Renderer();

    void SetSampleCount(uint cnt);
    uint GetSampleCount() const { return m_sample_count; }

    // --- additions ---------------------------------------------------------------------------
    // 0: Moeller-Trumbore (reference triangle.h:15-107, the live one), 1: plane + barycentric
    // (triangle.h:210-226).  Gamma 1/2 is on by default like the reference's GAMMA_CORRECTION.
    void SetIntersectVariant(uint variant) { m_variant = variant ? 1u : 0u; }
    void SetGammaCorrection(bool on) { m_gamma = on; }
    // The reference's other alternates (none reachable through its own API): GenerateRay's orthographic branch
    // (camera.h:25-36; width of the viewing volume, 0 = perspective) and the two commented-out shading lines,
    // 1: "Vec3f n = tri.n" (renderer.cpp:116), 2: "col += Vec3f(t / 3)" (renderer.cpp:118), 0: the live one
    void SetOrthographicWidth(float width) { m_ortho_width = width; }
    void SetShadingMode(uint mode) { m_shade_mode = mode <= 2u ? mode : 0u; }
    Scene * GetScene() { return m_scene.get(); }
    float GetLastKernelMilliseconds() const { return m_last_kernel_ms; }
    // The reference's sphere tracer (renderer.h:21 / renderer.cpp:24-41; protected and unused there): a one-ray
    // GPU query through cuda_trace_ray_march.
    bool RayMarch(Vec3f origin, Vec3f dir, float& t);

protected:
    void RenderTile(Tile& tile) override;
    void RenderTiles(Tile * const *tiles, uint count) override;
    void OnCancel() override;

    std::unique_ptr<Scene> m_scene;
    uint m_sample_count = 16;
    uint m_variant = 0;
    bool m_gamma = true;
    float m_ortho_width = 0.0f;
    uint m_shade_mode = 0;
    float m_last_kernel_ms = 0.0f;
};

#endif

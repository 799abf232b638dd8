// Renderer = Framebuffer + Scene, the reference's top-level rendering object (renderer.h:11-35):
// Renderer(unique_ptr<Scene>), SetSampleCount / GetSampleCount, RenderTile override.  The per-tile
// kernel itself (reference renderer.cpp:43-136) runs on the GPU: RenderTile() is a one-tile
// cuda_trace_tiles launch, RenderTiles() a single launch for the whole frame.
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
    float m_last_kernel_ms = 0.0f;
    uint32 *m_frame = nullptr;   // page-locked full-frame staging the device framebuffer is copied into
    size_t m_frame_pixels = 0;
};

#endif

// Host copy of the primary-ray set-up (reference camera.h:8-47) -- kept for API parity and to
// compute the per-frame constants the device needs.  The device evaluates the per-ray part
// itself (csrc/rt_device.cuh generate_ray).
#ifndef RTM_HOST_CAMERA_H
#define RTM_HOST_CAMERA_H

#include <cmath>

#include "lin_alg.h"
#include "types.h"

// fov_xs = tan(hfov / 2) evaluated in DOUBLE and rounded to float, aspect = w / h: the two values
// the reference derives per ray (camera.h:24,41-42; the unqualified tan() is the double overload
// under libstdc++).  Frame constants: computed once here and passed to cuda_trace_tiles.
inline void CameraFrameConstants(float hfov_degrees, uint width, uint height, float& fov_xs, float& aspect)
{
    const float hfov = DegToRad(hfov_degrees);
    fov_xs = float(::tan(double(hfov / 2)));
    aspect = float(width) / float(height);
}

inline void GenerateRay(const Matrix44f& camera, Vec2ui pixel, uint width, uint height, Vec2f sample_offs,
                        bool ortho, float width_or_hfov, Vec3f& origin, Vec3f& dir)
{
    const float ndc_x = (pixel.x + sample_offs.x) / float(width) * 2.0f - 1.0f;
    const float ndc_y = (pixel.y + sample_offs.y) / float(height) * 2.0f - 1.0f;
    const float aspect = float(width) / float(height);
    if (ortho)
    {
        // frame [-w/2, w/2] horizontally, keep the aspect vertically (reference camera.h:25-36)
        const float ow = width_or_hfov, oh = float(ow) / aspect;
        camera.Transf4x4(Vec3f(float(ndc_x * (float(ow) / 2.0)), float(ndc_y * (float(oh) / 2.0)), 0.0f), origin);
        camera.Transf3x3(Vec3f(0.0f, 0.0f, -1.0f), dir);
    }
    else
    {
        float fov_xs, unused;
        CameraFrameConstants(width_or_hfov, width, height, fov_xs, unused);
        camera.Transf4x4(Vec3f(0.0f), origin);
        camera.Transf3x3(Normalize(Vec3f(ndc_x * fov_xs, ndc_y * fov_xs / aspect, -1.0f)), dir);
    }
}

#endif

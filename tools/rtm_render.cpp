// Headless command-line renderer on the drop-in host API (Mesh -> Scene -> Renderer -> BMP): the
// replacement for the reference's GLUT viewer loop for boxes without a display.
//
//   rtm_render --scene killeroo|<file.dat|file.meshbin> [--width W] [--height H] [--spp N]
//              [--grid-res R] [--variant 0|1] [--gpus G] [--frames F] [--eye x y z] [--at x y z]
//              [--fov deg] [--ortho width] [--shade 0|1|2] [--out image.bmp]
//
// Built by <package>/build.py as <package>/rtm_render.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>

#include "renderer.h"
#include "trace.h"

namespace
{
bool ends_with(const std::string& s, const char *suffix)
{
    const size_t n = std::strlen(suffix);
    return s.size() >= n && s.compare(s.size() - n, n, suffix) == 0;
}
}

int main(int argc, char **argv)
{
    std::string scene = "cornell", out = "out.bmp", assets = "assets/meshes";
    uint width = 1920, height = 1080, spp = 16, grid_res = 64, variant = 0, gpus = 1, frames = 1, shade = 0;
    float ortho_width = 0.0f; // > 0: orthographic camera of that width (camera.h:25-36)
    float eye[3] = { -1.6f, 1.2f, -1.0f }, at[3] = { 0.0f, 0.0f, -0.1f }, fov = 30.0f;
    bool camera_given = false;
    for (int i = 1; i < argc; i++)
    {
        const std::string a = argv[i];
        auto next = [&]() -> const char * { if (i + 1 >= argc) { std::fprintf(stderr, "missing value for %s\n", a.c_str()); std::exit(2); } return argv[++i]; };
        if (a == "--scene") scene = next();
        else if (a == "--assets") assets = next();
        else if (a == "--width") width = uint(std::atoi(next()));
        else if (a == "--height") height = uint(std::atoi(next()));
        else if (a == "--spp") spp = uint(std::atoi(next()));
        else if (a == "--grid-res") grid_res = uint(std::atoi(next()));
        else if (a == "--variant") variant = uint(std::atoi(next()));
        else if (a == "--gpus") gpus = uint(std::atoi(next()));
        else if (a == "--frames") frames = uint(std::atoi(next()));
        else if (a == "--fov") { fov = float(std::atof(next())); camera_given = true; }
        else if (a == "--eye") { for (float& v : eye) v = float(std::atof(next())); camera_given = true; }
        else if (a == "--at") { for (float& v : at) v = float(std::atof(next())); camera_given = true; }
        else if (a == "--ortho") ortho_width = float(std::atof(next()));
        else if (a == "--shade") shade = uint(std::atoi(next())); // 1 face normals, 2 depth (renderer.cpp:116,118)
        else if (a == "--out") out = next();
        else { std::fprintf(stderr, "unknown option %s\n", a.c_str()); return 2; }
    }

    try
    {
        std::unique_ptr<Mesh> mesh(new Mesh());
        if (scene == "cornell")
        {
            // preset 1 without the cube (reference application.cpp:319-341)
            mesh->CornellBox();
            mesh->NormalizeDimensions();
            if (!camera_given) { eye[0] = 0; eye[1] = 0; eye[2] = -2; at[0] = at[1] = at[2] = 0; fov = 51.0f; }
        }
        else if (scene == "killeroo")
        {
            // preset 8 (reference application.cpp:443-459)
            if (!mesh->ReadBinary((assets + "/killeroo.meshbin").c_str()))
                throw std::runtime_error("cannot read " + assets + "/killeroo.meshbin");
            mesh->NormalizeDimensions();
            const float quad[12] = { -0.75f, -0.229267f, 0.75f, 0.75f, -0.229267f, 0.75f,
                                     0.75f, -0.229267f, -0.75f, -0.75f, -0.229267f, -0.75f };
            mesh->AddQuad(quad);
        }
        else
        {
            const bool ok = ends_with(scene, ".meshbin") ? mesh->ReadBinary(scene.c_str()) : mesh->Read(scene.c_str());
            if (!ok)
                throw std::runtime_error("cannot read mesh " + scene);
            mesh->NormalizeDimensions();
        }
        Matrix44f cam;
        cam.BuildLookAtMatrix(Vec3f(eye), Vec3f(at));
        Grid::SetDeviceCount(int(gpus));
        std::unique_ptr<Scene> sc(new Scene(std::move(mesh), fov, cam, grid_res));
        Renderer renderer(std::move(sc));
        renderer.SetSampleCount(spp);
        renderer.SetIntersectVariant(variant);
        renderer.SetOrthographicWidth(ortho_width);
        renderer.SetShadingMode(shade);
        for (uint f = 0; f < frames; f++)
        {
            if (f == 0)
                renderer.Resize(width, height);
            else
                renderer.StartRendering();
            if (!renderer.WaitRendering())
                throw std::runtime_error(renderer.GetLastError());
            const double rays = double(width) * height * spp;
            std::printf("frame %u: %.3f ms wall, %.3f ms kernel, %.1f Mrays/s\n", f, renderer.GetLastRenderSeconds() * 1e3,
                        renderer.GetLastKernelMilliseconds(), rays / (renderer.GetLastKernelMilliseconds() * 1e3));
        }
        renderer.SaveToBMP(out.c_str());
    }
    catch (const std::exception& e)
    {
        std::fprintf(stderr, "rtm_render: %s\n", e.what());
        return 1;
    }
    return 0;
}

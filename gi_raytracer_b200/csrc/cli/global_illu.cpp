// global-illu — the reference's executable (main.cpp:19-46) run headless: no Qt window, the frame is written as PNG or PPM (by extension).
// usage: global-illu [scene.scn] [width height] [out.png|out.ppm] [--device N] [--max-depth D] [--spp N] [--photons P] [--seed S]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "../../../include/gi_api.h"

extern "C" int gih_render_scene(const char* path, int w, int h, int device, int max_depth, int spp_override, int photons_override, uint64_t seed,
                                const char* out_ppm, uint8_t* rgb_out, gi_stats* frame_stats, gi_stats* photon_stats, double* photon_ms, double* frame_ms);

int main(int argc, char** argv)
{
    std::string scene = "scenes/cornell/cornell.scn", out = "render.ppm";
    int w = 1000, h = 1000, device = 0, max_depth = -1, spp = 0, photons = -1, pos = 0;
    unsigned long long seed = 1;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        if (a == "--device" && i + 1 < argc) device = atoi(argv[++i]);
        else if (a == "--max-depth" && i + 1 < argc) max_depth = atoi(argv[++i]);
        else if (a == "--spp" && i + 1 < argc) spp = atoi(argv[++i]);
        else if (a == "--photons" && i + 1 < argc) photons = atoi(argv[++i]);
        else if (a == "--seed" && i + 1 < argc) seed = strtoull(argv[++i], nullptr, 10);
        else {
            if (pos == 0) scene = a; else if (pos == 1) w = atoi(a.c_str()); else if (pos == 2) h = atoi(a.c_str()); else if (pos == 3) out = a;
            pos++;
        }
    }
    gi_stats fs, ps;
    double pms = 0, fms = 0;
    int rc = gih_render_scene(scene.c_str(), w, h, device, max_depth, spp, photons, seed, out.c_str(), nullptr, &fs, &ps, &pms, &fms);
    if (rc != GI_OK) { std::fprintf(stderr, "global-illu: failed with code %d\n", rc); return 1; }
    double rays = (double)fs.closest_rays + (double)fs.shadow_rays;
    std::printf("frame %dx%d: %.1f ms, %.3f Mrays (closest %.3f M + shadow %.3f M) -> %.2f Mrays/s, %.3f M gathers; photons: %.1f ms, %llu stored\n", w, h, fms,
                rays / 1e6, fs.closest_rays / 1e6, fs.shadow_rays / 1e6, rays / 1e3 / fms, fs.gathers / 1e6, pms, (unsigned long long)ps.photons_stored);
    std::printf("wrote %s\n", out.c_str());
    return 0;
}

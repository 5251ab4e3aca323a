// global-illu — the reference's executable (main.cpp:19-46) run headless: the same calls on the same classes (Camera, RayTracer, Octree,
// loadScene, setScene), then — instead of the Qt window, whose Viewer calls start() + run(w, h) on a worker thread and whose Gui saves
// the image through QImage (viewer.h:48-62, gui.h:39-45) — start(), run(w, h) and the frame written as PNG or PPM (by extension).
// usage: global-illu [scene.scn] [width height] [out.png|out.ppm] [--gpus N] [--device D] [--max-depth D] [--spp N] [--photons P] [--seed S]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>

#include "../host/gi_scene.hpp"

int main(int argc, char** argv)
{
    std::string path = "scenes/foliage/foliage.scn", out = "render.png";   // main.cpp:39: the default scene
    int w = 1000, h = 1000, device = 0, gpus = 1, max_depth = -1, spp = 0, photons = -1, pos = 0;   // gui: Gui window(1000, 1000, raytracer)
    unsigned long long seed = 1;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        if (a == "--device" && i + 1 < argc) device = atoi(argv[++i]);
        else if (a == "--gpus" && i + 1 < argc) gpus = atoi(argv[++i]);
        else if (a == "--max-depth" && i + 1 < argc) max_depth = atoi(argv[++i]);
        else if (a == "--spp" && i + 1 < argc) spp = atoi(argv[++i]);
        else if (a == "--photons" && i + 1 < argc) photons = atoi(argv[++i]);
        else if (a == "--seed" && i + 1 < argc) seed = strtoull(argv[++i], nullptr, 10);
        else {
            if (pos == 0) path = a; else if (pos == 1) w = atoi(a.c_str()); else if (pos == 2) h = atoi(a.c_str()); else if (pos == 3) out = a;
            pos++;
        }
    }
    Camera camera({ 10, 5, 0 }, { 0, 0, 0 });        // main.cpp:30
    RayTracer raytracer(camera);                     // main.cpp:32
    Octree* scene = new Octree();                    // main.cpp:36
    loadScene(scene, raytracer, path.c_str());       // main.cpp:38-41
    raytracer.setScene(scene);                       // main.cpp:43
    raytracer.device = device;
    raytracer.gpus = gpus < 1 ? 1 : gpus;
    raytracer.seed = seed;
    if (max_depth >= 0) raytracer.max_depth = max_depth;
    if (spp > 0) { raytracer.min_samples = raytracer.max_samples = spp; }
    if (photons >= 0) raytracer.photons = photons;
    raytracer.start();                               // viewer.h:48-54
    const int rc = raytracer.run(w, h);
    if (rc != GI_OK) { std::fprintf(stderr, "global-illu: failed with code %d\n", rc); return 1; }
    const bool png = out.size() > 4 && (out.compare(out.size() - 4, 4, ".png") == 0 || out.compare(out.size() - 4, 4, ".PNG") == 0);
    const bool ok = png ? raytracer.getImage()->writePNG(out.c_str()) : raytracer.getImage()->writePPM(out.c_str());
    const gi_stats& fs = raytracer.last_frame_stats;
    const double rays = (double)fs.closest_rays + (double)fs.shadow_rays;
    std::printf("frame %dx%d on %d GPU(s): %.1f ms, %.3f Mrays (closest %.3f M + shadow %.3f M) -> %.2f Mrays/s, %.3f M gathers; photons: %.1f ms, %llu stored\n", w, h, raytracer.gpus,
                raytracer.last_frame_ms, rays / 1e6, fs.closest_rays / 1e6, fs.shadow_rays / 1e6, rays / 1e3 / raytracer.last_frame_ms, fs.gathers / 1e6, raytracer.last_photon_ms,
                (unsigned long long)raytracer.last_photon_stats.photons_stored);
    if (!ok) { std::fprintf(stderr, "global-illu: cannot write %s\n", out.c_str()); return 1; }
    std::printf("wrote %s\n", out.c_str());
    return 0;
}

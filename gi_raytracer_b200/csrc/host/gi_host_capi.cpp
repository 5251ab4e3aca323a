// gi_host_capi.cpp — a small C facade over the host scene classes for the Python harness (tests, bench.py):
// load a .scn through loadScene, rebuild the octree, flatten, and expose the arrays.  No device work here.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <thread>
#include <cstring>
#include <iostream>
#include <sstream>
#include <string>

#include "gi_scene.hpp"

struct gih_scene {
    Octree* octree;
    RayTracer* rt;
    FlatScene flat;
    gi_scene_desc desc;
};

extern "C" {

// Load + rebuild + flatten.  quiet != 0 silences the loader's std::cout chatter.
int gih_scene_load(const char* path, int quiet, gih_scene** out)
{
    if (!path || !out) return GI_ERR_INVALID;
    std::streambuf* old = nullptr;
    std::ostringstream sink;
    if (quiet) old = std::cout.rdbuf(sink.rdbuf());
    gih_scene* s = new gih_scene();
    Camera camera(gi::dvec3(10, 5, 0), gi::dvec3(0, 0, 0));  // main.cpp:30 default camera
    s->rt = new RayTracer(camera);
    s->octree = new Octree();
    std::string spath = path;
    int api_variant = 0;   // "<file>.scn#api" / "#api2": add the API-built primitives of api_scene.inc
    if (spath.size() > 4 && spath.compare(spath.size() - 4, 4, "#api") == 0) { api_variant = 1; spath.resize(spath.size() - 4); }
    else if (spath.size() > 5 && spath.compare(spath.size() - 5, 5, "#api2") == 0) { api_variant = 2; spath.resize(spath.size() - 5); }
    const bool api_scene = api_variant != 0;
    loadScene(s->octree, *s->rt, spath.c_str());
    if (api_scene) {
        Octree* o = s->octree;
#define V3(x, y, z) gi::dvec3(x, y, z)
#include "api_scene.inc"
#undef V3
    }
    if (!s->octree->entities().empty()) s->octree->rebuild(); else s->octree->valid = true;
    s->octree->flatten(s->rt->_camera, s->rt->ambient, s->flat);
    s->desc = s->flat.desc();
    if (quiet) std::cout.rdbuf(old);
    *out = s;
    return GI_OK;
}

// Entity::boundingBox() of every primitive, [n_prims][6] (the input gi_octree_build wants next to prim_type / prim_geom)
void gih_scene_prim_bbox(const gih_scene* s, double* out6)
{
    std::vector<double> b;
    s->octree->entity_boxes(b);
    std::memcpy(out6, b.data(), b.size() * sizeof(double));
}
// rebuild the loaded scene's octree on the device and flatten again; returns a GI_* code, *build_ms = device time
int gih_scene_rebuild_device(gih_scene* s, gi_ctx* ctx, double* build_ms)
{
    if (!s || !ctx) return GI_ERR_INVALID;
    int rc = s->octree->rebuild(ctx);
    if (rc != GI_OK) return rc;
    if (build_ms) *build_ms = s->octree->last_build_ms;
    s->octree->flatten(s->rt->_camera, s->rt->ambient, s->flat);
    s->desc = s->flat.desc();
    return GI_OK;
}

const gi_scene_desc* gih_scene_desc(const gih_scene* s) { return s ? &s->desc : nullptr; }

// photons, photon_depth, min_samples, max_samples, noise_thresh
void gih_scene_knobs(const gih_scene* s, int* photons, int* photon_depth, int* min_samples, int* max_samples, double* noise_thresh)
{
    if (photons) *photons = s->rt->photons;
    if (photon_depth) *photon_depth = s->rt->photon_depth;
    if (min_samples) *min_samples = s->rt->min_samples;
    if (max_samples) *max_samples = s->rt->max_samples;
    if (noise_thresh) *noise_thresh = s->rt->noise_thresh;
}

void gih_scene_free(gih_scene* s)
{
    if (!s) return;
    delete s->rt;
    delete s->octree;  // entities/materials/textures are never freed, like the reference (SURVEY §8b ownership)
    delete s;
}

// Headless equivalent of the reference's main(): load, run(w,h) on `device`, write a PPM.  Returns 0 or GI_ERR_*.
int gih_render_scene(const char* path, int w, int h, int device, int max_depth, int spp_override, int photons_override, uint64_t seed,
                     const char* out_ppm, uint8_t* rgb_out, gi_stats* frame_stats, gi_stats* photon_stats, double* photon_ms, double* frame_ms)
{
    Camera camera(gi::dvec3(10, 5, 0), gi::dvec3(0, 0, 0));
    RayTracer rt(camera);
    Octree* scene = new Octree();
    loadScene(scene, rt, path);
    rt.setScene(scene);
    rt.device = device;
    rt.seed = seed;
    if (max_depth >= 0) rt.max_depth = max_depth;
    if (spp_override > 0) { rt.min_samples = rt.max_samples = spp_override; }
    if (photons_override >= 0) rt.photons = photons_override;
    rt.start();
    int rc = rt.run(w, h);
    if (rc != GI_OK) { delete scene; return rc; }
    if (out_ppm && *out_ppm) {   // by extension: .png -> PNG, anything else -> binary PPM
        const std::string o = out_ppm;
        const bool png = o.size() > 4 && (o.compare(o.size() - 4, 4, ".png") == 0 || o.compare(o.size() - 4, 4, ".PNG") == 0);
        if (png) rt.getImage()->writePNG(out_ppm); else rt.getImage()->writePPM(out_ppm);
    }
    if (rgb_out) std::memcpy(rgb_out, rt.getImage()->rgb.data(), rt.getImage()->rgb.size());
    if (frame_stats) *frame_stats = rt.last_frame_stats;
    if (photon_stats) *photon_stats = rt.last_photon_stats;
    if (photon_ms) *photon_ms = rt.last_photon_ms;
    if (frame_ms) *frame_ms = rt.last_frame_ms;
    delete scene;
    return GI_OK;
}

// The viewer's use of the tracer (viewer.h:29-62): run() on a worker thread with progressive bands, stop() from the calling thread
// once `stop_after_rows` rows have been published (< 0: never), join.  rgb_out = the image as the viewer would see it;
// *rows_done = rows published; *stop_latency_ms = time from stop() to run() returning.  Returns run()'s code.
int gih_render_progressive(const char* path, int w, int h, int device, int max_depth, int spp_override, int photons_override, uint64_t seed, int rows_per_band, int stop_after_rows,
                           uint8_t* rgb_out, int* rows_done, double* stop_latency_ms)
{
    Camera camera(gi::dvec3(10, 5, 0), gi::dvec3(0, 0, 0));
    RayTracer rt(camera);
    Octree* scene = new Octree();
    loadScene(scene, rt, path);
    rt.setScene(scene);
    rt.device = device;
    rt.seed = seed;
    rt.progressive_rows = rows_per_band;
    if (max_depth >= 0) rt.max_depth = max_depth;
    if (spp_override > 0) { rt.min_samples = rt.max_samples = spp_override; }
    if (photons_override >= 0) rt.photons = photons_override;
    rt.start();
    int rc = GI_OK;
    std::atomic<bool> done{ false };
    std::thread worker([&]() { rc = rt.run(w, h); done = true; });
    auto t_stop = std::chrono::steady_clock::now();
    bool stopped = false;
    while (!done) {
        if (!stopped && stop_after_rows >= 0 && rt.rows_done() >= std::max(stop_after_rows, 1)) { rt.stop(); t_stop = std::chrono::steady_clock::now(); stopped = true; }
        std::this_thread::sleep_for(std::chrono::microseconds(200));
    }
    worker.join();
    if (stop_latency_ms) *stop_latency_ms = stopped ? std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_stop).count() : 0.0;
    if (rows_done) *rows_done = rt.rows_done();
    if (rgb_out && rt.getImage()->rgb.size() == (size_t)w * h * 3) std::memcpy(rgb_out, rt.getImage()->rgb.data(), rt.getImage()->rgb.size());
    delete scene;
    return rc;
}

}  // extern "C"

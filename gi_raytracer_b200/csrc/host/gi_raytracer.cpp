// gi_raytracer.cpp — RayTracer::run and PhotonMap::rebuild: the host-side callers of the C ABI (raytracer.h:41-165).
#include <cstdlib>
#include <chrono>
#include <cstdio>
#include <iostream>

#include "gi_scene.hpp"

RayTracer::~RayTracer()
{
    if (_ctx) gi_destroy(_ctx);
}

gi_ctx* RayTracer::context()
{
    if (!_ctx) {
        int rc = gi_create(device, &_ctx);
        if (rc != GI_OK) { std::cout << "gi_create failed (" << rc << "): no usable CUDA device — there is no CPU fallback\n"; _ctx = nullptr; }
    }
    return _ctx;
}

void RayTracer::setScene(Octree* scene)  // raytracer.h:35-39: the photon map takes the scene's root box
{
    _scene = scene;
    delete _photon_map;
    _photon_map = new PhotonMap(_scene->_root._bbox.min, _scene->_root._bbox.max);
    _uploaded = false;
}

void PhotonMap::rebuild(gi_ctx* ctx)  // photonMap.cpp:33-47
{
    if (!staged.empty()) {
        std::vector<double> buf(staged.size() * 9);
        for (size_t i = 0; i < staged.size(); i++) {
            const Photon& p = staged[i];
            double* o = &buf[i * 9];
            o[0] = p.origin.x; o[1] = p.origin.y; o[2] = p.origin.z; o[3] = p.dir.x; o[4] = p.dir.y; o[5] = p.dir.z; o[6] = p.col.x; o[7] = p.col.y; o[8] = p.col.z;
        }
        if (gi_photon_upload(ctx, staged.size(), buf.data()) != GI_OK) return;
    }
    double box[6] = { min.x, min.y, min.z, max.x, max.y, max.z };
    if (gi_photon_map_build(ctx, box) == GI_OK) valid = true;
}

int RayTracer::run(int w, int h)
{
    std::cout << "starting raytracer with frame size: " << w << ", " << h << "\n";
    _image = std::make_shared<Image>(w, h);
    if (!_running) return GI_OK;  // like the reference: nothing is rendered unless start() was called (raytracer.h:98)
    gi_ctx* ctx = context();
    if (!ctx) return GI_ERR_NO_DEVICE;
    if (!_scene) return GI_ERR_NO_SCENE;
    int rc;
    if (!_scene->valid) {                                                        // raytracer.h:56-59
        // Node::partition runs on the device (identical tree, tested); GI_HOST_BUILD=1 keeps the host build
        if (std::getenv("GI_HOST_BUILD") || (rc = _scene->rebuild(ctx)) != GI_OK) _scene->rebuild();
        _uploaded = false;
    }
    if (!_uploaded) {
        FlatScene flat;
        _scene->flatten(_camera, ambient, flat);
        gi_scene_desc d = flat.desc();
        if ((rc = gi_scene_upload(ctx, &d)) != GI_OK) { std::cout << "gi_scene_upload: " << gi_last_error(ctx) << "\n"; return rc; }
        _uploaded = true;
    }
    if (!_photon_map->valid) {                                                   // raytracer.h:61-72
        auto t0 = std::chrono::high_resolution_clock::now();
        std::cout << "emitting photons...\n";
        uint64_t stored = 0;
        if ((rc = gi_photon_trace(ctx, photons, 5, seed, &stored, &last_photon_stats)) != GI_OK) { std::cout << "gi_photon_trace: " << gi_last_error(ctx) << "\n"; return rc; }
        _photon_map->rebuild(ctx);
        gi_synchronize(ctx);
        auto t1 = std::chrono::high_resolution_clock::now();
        last_photon_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
        std::cout << "photon time: " << last_photon_ms / 1000.0 << " s\n";
        std::cout << "total photons: " << stored << "\n";
        if (!_photon_map->valid) { std::cout << "gi_photon_map_build: " << gi_last_error(ctx) << "\n"; return GI_ERR_CUDA; }
    }
    // `samples N N t`: a fixed count, rendered as one sample range.  `samples min max t` with min != max: the reference's
    // variance-driven per-pixel loop (raytracer.h:100-148) -> gi_render_adaptive; the result is the running-mean colour.
    gi_render_params p;
    p.width = w; p.height = h; p.max_depth = max_depth; p.min_depth = min_depth; p.spp = max_samples; p.k_photons = 32; p.caustic_max_depth = 10; p._pad = 0; p.seed = seed;
    std::vector<double> accum((size_t)w * h * 3);
    auto f0 = std::chrono::high_resolution_clock::now();
    int resolve_spp = p.spp;
    if (min_samples != max_samples) {
        if ((rc = gi_render_adaptive(ctx, &p, min_samples, max_samples, noise_thresh, 0, 0, w, h, accum.data(), nullptr, &last_frame_stats)) != GI_OK) { std::cout << "gi_render_adaptive: " << gi_last_error(ctx) << "\n"; return rc; }
        resolve_spp = 1;
    } else if ((rc = gi_render_tile(ctx, &p, 0, 0, w, h, 0, p.spp, accum.data(), &last_frame_stats)) != GI_OK) { std::cout << "gi_render_tile: " << gi_last_error(ctx) << "\n"; return rc; }
    if ((rc = gi_resolve(ctx, (size_t)w * h, accum.data(), resolve_spp, _image->rgb.data())) != GI_OK) { std::cout << "gi_resolve: " << gi_last_error(ctx) << "\n"; return rc; }
    auto f1 = std::chrono::high_resolution_clock::now();
    last_frame_ms = std::chrono::duration<double, std::milli>(f1 - f0).count();
    return GI_OK;
}

// gi_raytracer.cpp — RayTracer::run and PhotonMap::rebuild: the host-side callers of the C ABI (raytracer.h:41-165).
#include <algorithm>
#include <cstddef>
#include <cstring>
#include <cstdlib>
#include <chrono>
#include <cstdio>
#include <iostream>

#include <thread>

#include "gi_scene.hpp"

RayTracer::~RayTracer()
{
    for (gi_ctx* p : _peers) gi_destroy(p);
    if (gi_ctx* c = _ctx.load()) gi_destroy(c);
}

RayTracer::RayTracer(const RayTracer& o)
    : photons(o.photons), photon_depth(o.photon_depth), min_samples(o.min_samples), max_samples(o.max_samples), noise_thresh(o.noise_thresh), ambient(o.ambient), _camera(o._camera),
      max_depth(o.max_depth), min_depth(o.min_depth), seed(o.seed), device(o.device), progressive_rows(o.progressive_rows), _running(o._running.load()), _rows_done(o._rows_done.load()),
      _scene(o._scene), _photon_map(o._photon_map), _image(o._image)
{
    // the copy shares the scene, the photon map object and the image (gui.h:19, viewer.h:16) but gets its own device context, created
    // lazily: that context holds neither the scene nor the map yet (_uploaded / _map_in_ctx start false), so the first run()
    // of the copy uploads and builds them there instead of rendering without caustics
}

void RayTracer::stop()   // may be called from another thread while run() is in flight (viewer.h:29-34)
{
    _running = false;
    if (gi_ctx* c = _ctx.load(std::memory_order_acquire)) gi_cancel(c, 1);
}
void RayTracer::start()
{
    _running = true;
    if (gi_ctx* c = _ctx.load(std::memory_order_acquire)) gi_cancel(c, 0);
}

// The frame on several GPUs of this machine: one thread per GPU drives its context; everything between the GPUs is the C ABI's own
// NCCL (gi_comm_init, gi_photon_map_bcast, the gather inside gi_render_rows_image).
int RayTracer::run_multi(int w, int h)
{
    gi_ctx* ctx0 = context();
    const int n = gpus;
    if ((int)_peers.size() != n - 1) {
        for (gi_ctx* p : _peers) gi_destroy(p);
        _peers.clear();
        _peers_have_scene = false;
        for (int r = 1; r < n; r++) {
            gi_ctx* c = nullptr;
            int rc = gi_create(device + r, &c);
            if (rc != GI_OK) { std::cout << "gi_create failed on device " << device + r << " (" << rc << ")\n"; for (gi_ctx* p : _peers) gi_destroy(p); _peers.clear(); return rc; }
            _peers.push_back(c);
        }
        unsigned char id[GI_COMM_ID_BYTES];
        if (int rc = gi_comm_unique_id(id, sizeof(id))) { std::cout << "gi_comm_unique_id failed (" << rc << ")\n"; return rc; }
        std::vector<int> rcs(n, GI_OK);
        std::vector<std::thread> th;
        for (int r = 0; r < n; r++) th.emplace_back([&, r]() { rcs[r] = gi_comm_init(r == 0 ? ctx0 : _peers[r - 1], id, sizeof(id), r, n); });   // collective
        for (auto& t : th) t.join();
        for (int r = 0; r < n; r++) if (rcs[r] != GI_OK) { std::cout << "gi_comm_init: " << gi_last_error(r == 0 ? ctx0 : _peers[r - 1]) << "\n"; return rcs[r]; }
    }
    FlatScene flat;
    _scene->flatten(_camera, ambient, flat);
    const gi_scene_desc d = flat.desc();
    const bool need_map = !_photon_map->valid || !_map_in_ctx || !_peers_have_scene;
    if (need_map && (!_photon_map->valid || !_map_in_ctx)) {                       // raytracer.h:61-72, on the first GPU
        int rc;
        if (!_uploaded) { if ((rc = gi_scene_upload(ctx0, &d)) != GI_OK) { std::cout << "gi_scene_upload: " << gi_last_error(ctx0) << "\n"; return rc; } _uploaded = true; }
        auto t0 = std::chrono::high_resolution_clock::now();
        std::cout << "emitting photons...\n";
        uint64_t stored = 0;
        if ((rc = gi_photon_trace(ctx0, photons, 5, seed, &stored, &last_photon_stats)) != GI_OK) { std::cout << "gi_photon_trace: " << gi_last_error(ctx0) << "\n"; return rc; }
        _photon_map->rebuild(ctx0);
        gi_synchronize(ctx0);
        last_photon_ms = std::chrono::duration<double, std::milli>(std::chrono::high_resolution_clock::now() - t0).count();
        std::cout << "photon time: " << last_photon_ms / 1000.0 << " s\n";
        std::cout << "total photons: " << stored << "\n";
        if (!_photon_map->valid) { std::cout << "gi_photon_map_build: " << gi_last_error(ctx0) << "\n"; return GI_ERR_CUDA; }
        _map_in_ctx = true;
    }
    gi_render_params p;
    p.width = w; p.height = h; p.max_depth = max_depth; p.min_depth = min_depth; p.spp = max_samples; p.k_photons = 32; p.caustic_max_depth = 10; p._pad = 0; p.seed = seed;
    std::vector<int> rcs(n, GI_OK);
    std::vector<gi_stats> st(n);
    auto f0 = std::chrono::high_resolution_clock::now();
    {
        std::vector<std::thread> th;
        for (int r = 0; r < n; r++) th.emplace_back([&, r]() {
            gi_ctx* c = r == 0 ? ctx0 : _peers[r - 1];
            int rc = GI_OK;
            if ((r > 0 && !_peers_have_scene) || (r == 0 && !_uploaded)) rc = gi_scene_upload(c, &d);
            if (rc == GI_OK && need_map) rc = gi_photon_map_bcast(c, 0);               // "built once and broadcast": one slab
            if (rc == GI_OK) rc = gi_render_rows_image(c, &p, 16, 0, p.spp, r == 0 ? _image->rgb.data() : nullptr, 0, &st[r]);
            rcs[r] = rc;
        });
        for (auto& t : th) t.join();
    }
    _uploaded = true; _peers_have_scene = true;
    last_frame_ms = std::chrono::duration<double, std::milli>(std::chrono::high_resolution_clock::now() - f0).count();
    for (int r = 0; r < n; r++) if (rcs[r] != GI_OK) { std::cout << "gi_render (gpu " << r << "): " << gi_last_error(r == 0 ? ctx0 : _peers[r - 1]) << "\n"; return rcs[r]; }
    std::memset(&last_frame_stats, 0, sizeof(last_frame_stats));
    for (int r = 0; r < n; r++) {   // totals over the GPUs; times: the slowest GPU
        uint64_t* dst = reinterpret_cast<uint64_t*>(&last_frame_stats); const uint64_t* src = reinterpret_cast<const uint64_t*>(&st[r]);
        for (size_t k = 0; k < offsetof(gi_stats, trace_ms) / 8; k++) dst[k] += src[k];
        last_frame_stats.total_ms = std::max(last_frame_stats.total_ms, st[r].total_ms);
        last_frame_stats.trace_ms = std::max(last_frame_stats.trace_ms, st[r].trace_ms); last_frame_stats.shadow_ms = std::max(last_frame_stats.shadow_ms, st[r].shadow_ms);
        last_frame_stats.gather_ms = std::max(last_frame_stats.gather_ms, st[r].gather_ms); last_frame_stats.shade_ms = std::max(last_frame_stats.shade_ms, st[r].shade_ms);
    }
    _rows_done = h;
    return GI_OK;
}

gi_ctx* RayTracer::context()
{
    gi_ctx* c = _ctx.load(std::memory_order_acquire);
    if (!c) {
        int rc = gi_create(device, &c);
        if (rc != GI_OK) { std::cout << "gi_create failed (" << rc << "): no usable CUDA device — there is no CPU fallback\n"; return nullptr; }
        _ctx.store(c, std::memory_order_release);
        if (!_running) gi_cancel(c, 1);   // a stop() that arrived before the context existed
    }
    return c;
}

void RayTracer::setScene(Octree* scene)  // raytracer.h:35-39: the photon map takes the scene's root box
{
    _scene = scene;
    // a NEW map object: copies of this tracer made earlier keep theirs alive through their own shared_ptr (the reference never
    // frees its map, raytracer.h:38)
    _photon_map = std::make_shared<PhotonMap>(_scene->_root._bbox.min, _scene->_root._bbox.max);
    _uploaded = false;
    _map_in_ctx = false;
    _peers_have_scene = false;
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
    if (gi_photon_map_build(ctx, box) == GI_OK) { valid = true; _ctx = ctx; _host.clear(); }
}

std::vector<Photon*> PhotonMap::getInRange(gi::dvec3& pos, double& scale, double dist) const   // photonMap.cpp:50-68 (scale / dist are unused there too)
{
    (void)scale; (void)dist;
    std::vector<Photon*> res;
    if (!valid || !_ctx) return res;
    if (_host.empty()) {
        size_t n = 0;
        if (gi_photon_count(_ctx, &n) != GI_OK || !n) return res;
        std::vector<double> buf(n * 9);
        if (gi_photon_download(_ctx, n, buf.data()) != GI_OK) return res;
        _host.reserve(n);
        for (size_t i = 0; i < n; i++) { const double* p = &buf[9 * i]; _host.emplace_back(gi::dvec3(p[0], p[1], p[2]), gi::dvec3(p[3], p[4], p[5]), gi::dvec3(p[6], p[7], p[8])); }
    }
    const double q[3] = { pos.x, pos.y, pos.z };
    uint32_t cap = 256, count = 0;
    std::vector<uint32_t> ids(cap);
    if (gi_photon_in_range(_ctx, 1, q, cap, ids.data(), &count) != GI_OK) return res;
    if (count > cap) { cap = count; ids.resize(cap); if (gi_photon_in_range(_ctx, 1, q, cap, ids.data(), &count) != GI_OK) return res; }
    res.reserve(count);
    for (uint32_t k = 0; k < count; k++) if (ids[k] < _host.size()) res.push_back(&_host[ids[k]]);
    return res;
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
        _scene->attach(ctx, flat);   // Octree::intersect / intersectSorted and Entity::intersect now answer from this device
        _uploaded = true;
    }
    if (!_running) return GI_OK;                                                 // stop() arrived while the context was being created
    if (!_photon_map) return GI_ERR_NO_SCENE;
    if (!_photon_map->valid || !_map_in_ctx) {                                   // raytracer.h:61-72; also when the map lives in another copy's context
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
        _map_in_ctx = true;
    }
    if (gpus > 1 && min_samples == max_samples && progressive_rows == 0) return run_multi(w, h);
    // `samples N N t`: a fixed count, rendered as one sample range.  `samples min max t` with min != max: the reference's
    // variance-driven per-pixel loop (raytracer.h:100-148) -> gi_render_adaptive; the result is the running-mean colour.
    gi_render_params p;
    p.width = w; p.height = h; p.max_depth = max_depth; p.min_depth = min_depth; p.spp = max_samples; p.k_photons = 32; p.caustic_max_depth = 10; p._pad = 0; p.seed = seed;
    auto f0 = std::chrono::high_resolution_clock::now();
    const bool adaptive = min_samples != max_samples;
    const int resolve_spp = adaptive ? 1 : p.spp;
    // bands of rows, top to bottom (one band = the whole frame unless progressive_rows is set); a band that has not started
    // when stop() arrives is skipped, like the rows of the reference's loop (raytracer.h:93-98)
    const int band = progressive_rows > 0 ? progressive_rows : h;
    std::vector<double> accum;
    std::vector<uint8_t> band_rgb((size_t)w * std::min(band, h) * 3);   // a cancelled band must not touch the image: staged, then published
    _rows_done = 0;
    std::memset(&last_frame_stats, 0, sizeof(last_frame_stats));
    for (int y0 = 0; y0 < h; y0 += band) {
        if (!_running) break;
        const int y1 = std::min(h, y0 + band);
        gi_stats bs;
        uint8_t* rows = _image->rgb.data() + (size_t)y0 * w * 3;
        if (adaptive) {
            accum.resize((size_t)w * (y1 - y0) * 3);
            rc = gi_render_adaptive(ctx, &p, min_samples, max_samples, noise_thresh, 0, y0, w, y1, accum.data(), nullptr, &bs);
            if (rc == GI_OK) rc = gi_resolve(ctx, (size_t)w * (y1 - y0), accum.data(), resolve_spp, band_rgb.data());
        } else rc = gi_render_image(ctx, &p, 0, y0, w, y1, 0, p.spp, band_rgb.data(), nullptr, &bs);   // rendered and resolved on the device
        if (rc == GI_ERR_CANCELLED) break;   // stop() while the band was on the device: its pixels are not published
        if (rc != GI_OK) { std::cout << "gi_render: " << gi_last_error(ctx) << "\n"; return rc; }
        std::memcpy(rows, band_rgb.data(), (size_t)w * (y1 - y0) * 3);
        _rows_done = y1;
        // totals over the bands
        uint64_t* dst = reinterpret_cast<uint64_t*>(&last_frame_stats); const uint64_t* src = reinterpret_cast<const uint64_t*>(&bs);
        for (size_t k = 0; k < offsetof(gi_stats, trace_ms) / 8; k++) dst[k] += src[k];
        last_frame_stats.trace_ms += bs.trace_ms; last_frame_stats.shadow_ms += bs.shadow_ms; last_frame_stats.gather_ms += bs.gather_ms; last_frame_stats.shade_ms += bs.shade_ms;
        last_frame_stats.total_ms += bs.total_ms; last_frame_stats.bin_ms += bs.bin_ms;
        for (uint64_t *d2 = &last_frame_stats.tail_closest_rays, *e2 = &last_frame_stats.tail_gather_selected + 1, *s2 = &bs.tail_closest_rays; d2 != e2; ++d2, ++s2) *d2 += *s2;
    }
    auto f1 = std::chrono::high_resolution_clock::now();
    last_frame_ms = std::chrono::duration<double, std::milli>(f1 - f0).count();
    return GI_OK;
}

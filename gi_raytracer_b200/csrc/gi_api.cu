// gi_api.cu — the C ABI of libgi_b200.so (include/gi_api.h): context, scene upload, and the host side of every kernel
// launch.  All work of a context is enqueued on one CUDA stream; kernel families are timed with CUDA events on that
// stream (gi_last_kernel_ms).  There is no CPU fallback: every compute entry point needs a CUDA device.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <atomic>
#include <chrono>
#include <string>
#include <vector>

#include "gi_kernels.cuh"
#include "gi_octree_build.cuh"
#include "gi_query.cuh"

#define GI_VERSION "gi_b200 0.1.0 (sm_100a)"
#define GI_MAX_PATHS (1u << 23)   // paths in flight per chunk of the wavefront

// ---- small RAII-free device buffer helper -------------------------------------------------------------------------------------
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct FamStat { double ms = 0; uint64_t launches = 0; };
struct TimedLaunch { std::string fam; cudaEvent_t a, b; };

#define GI_NHL 8   // most hit lists the ring can hold (gi_ctx::ring are used; sched_mode 0 alternates between the first two)
struct gi_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    // scene
    bool has_scene = false;
    DScene S{};
    double root_box[6] = { 0, 0, 0, 0, 0, 0 };
    DevBuf b_nodes, b_refs, b_geom, b_nrm, b_uv, b_fnorm, b_pmat, b_ptype, b_mats, b_tex, b_texpx, b_lights, b_htab, b_hdims, b_fogs, b_foggrid;
    // photons + photon map
    DevBuf b_photons; size_t n_photons = 0;
    bool has_map = false;
    DevBuf b_slab;          // compact map: header | nodes | pos | dircol | pid
    size_t slab_bytes = 0;
    uint32_t pm_nodes = 0, pm_kept = 0, pm_leaves = 0, pm_depth = 0;
    DGatherMap G{};
    // workspaces
    DevBuf w0, w1, w2, w3, w4, w5, w6, w7, w8, w9;           // API staging
    DevBuf q_a[5], q_b[5], hl[7], ps[5], tq[5], ad[6], b_cnt, b_accum, b_scan0, b_scan1, b_misc, b_work, b_tail, b_binkey, b_binperm, b_binhist, b_bincur, b_gnode, b_gperm, b_ghist, b_gcur, b_gheavy;
    // device octree build (gi_octree_build): inputs, outputs (gi_scene_desc layout), per-level work buffers
    DevBuf ob_type, ob_geom, ob_bbox, ob_nodebox, ob_child, ob_mask, ob_poff, ob_pcnt, ob_leaf, ob_list[2], ob_owner[2], ob_flags, ob_pos, ob_abox[2], ob_anode[2], ob_astart[2], ob_acount[2],
        ob_slotactive[2], ob_slot[5], ob_rank[3], ob_tot;
    uint32_t ob_n_nodes = 0, ob_n_refs = 0; bool ob_valid = false;
    std::atomic<int> cancel{ 0 };  // gi_cancel: polled at launch boundaries
    // side streams: k_direct (side[0]) and the gather pipeline (side[1]) of a bounce depth run behind the next depth's bounce
    // kernel once the hit list is short (< overlap_threshold): those launches no longer fill the machine
    cudaStream_t main_stream = nullptr, side[2] = { nullptr, nullptr };   // side: the pair in use — one of side_pairs
    cudaStream_t side_pairs[2][2] = {};      // [0] at the main stream's priority (sched_mode 0), [1] below it (the deferred schedule)
    cudaEvent_t side_done[GI_NHL][2] = {};   // [hit list of the ring][side stream]
    cudaEvent_t fork_ev = nullptr;           // main stream -> side streams (the tail's queued shadow rays and gathers)
    cudaStream_t gather_aux = nullptr;       // k_gather_heavy beside the rest of a gather run (run_gather)
    cudaEvent_t gather_ev[2] = { nullptr, nullptr };   // [0] locate done (calling stream -> aux), [1] heavy done (aux -> calling stream)
    uint32_t overlap_threshold = 1u << 20;   // GI_OVERLAP_THRESHOLD, 0 = off
    // sched_mode (GI_SCHED_MODE): how a chunk's kernels are laid over the three streams when overlap is on.
    //   0  shadow rays of a depth beside its gather pipeline; both behind the NEXT depth's bounce kernel only when the depth is short
    //      (< overlap_threshold hits); everything is drained before the tail starts.
    //   1  "deferred": the main stream carries only what decides how paths go on — generate, bounce kernels, binning, the tail
    //      kernel — at the higher stream priority; EVERY depth's shadow rays go to side stream 0 and every gather pipeline run to side
    //      stream 1 (the tail's queued ones included), reading hit lists out of a ring of GI_NHL (memory: 132 B per path of a chunk and list).  The short, latency-bound launches
    //      of the deep bounces and the tail's one-warp-per-path walk then run underneath the long depth-0 / depth-1 launches
    //      instead of after them.  Per path the sums keep their order: L is only touched on the main stream, Ld only on side 0,
    //      Lc only on side 1, each in bounce order.
    //      (Measured and dropped, profiles/r02/ab_t23: holding the side work of a depth back while ANOTHER LONG bounce launch follows, so that
    //      long traversal launches never run side by side — slower on every scene, caustics 19.14 -> 19.59 ms, glass 135.7 -> 140.5.)
    //      ring: hit lists in use (GI_RING, 2..GI_NHL): with four, bounce d+4 waited for the shadow rays / gathers of depth d —
    //      cornell 46.0 ms against 43.9 with six; glass 140.8 / 135.6 / 132.8 ms with four / six / eight (profiles/r02/ab_t24).
    //   2  (default) 1, unless the last large frame of this scene had nothing short in it — no tail, no depth below 2^20
    //      hits (cornell at MAX_DEPTH 4: five long depths): then there is nothing latency-bound to put underneath the long launches,
    //      and running them side by side only costs (frames 43.9 or 45.8 ms from one run to the next, against a steady 43.9 under 0).
    int sched_mode = 2;
    int ring = 8;
    int sched_hint = 1;              // sched_mode 2: 1 = the last large frame of this scene had short launches or a tail (or none was rendered yet)
    DevBuf tsh[7];                           // the tail's deferred shadow rays (DTailQ::sh_*)
    DevBuf hl2[7], b_scan1s;                 // second hit list (depth parity), scan scratch of the gather side stream
    DevBuf hlr[GI_NHL - 2][7];               // hit lists 3 .. GI_NHL of the ring (sched_mode 1)
    bool no_implicit = false;      // GI_NO_IMPLICIT_BOXES at gi_create: always load child boxes (for A/B tests)
    int trace_mode = 0;            // 0: thread per ray, 1: warp per ray (API batch kernels; GI_TRACE_MODE)
    uint32_t tail_threshold = 32768; // queues smaller than this finish in the tail megakernel (GI_TAIL_THRESHOLD, 0 = off)
    int tail_mode = 0;               // tail gathers: 0 queued and served by one gather pipeline run, 1 inline in the tail kernel (GI_TAIL_MODE)
    int bounce_mode = 0;             // 0: pick per scene, 1: always one ray per thread, 2: always persistent warps with refetch (GI_BOUNCE_MODE)
    double nodes_per_ray = 0;        // node tests per closest-hit ray of the last frame rendered with the current scene
    double prims_per_ray = 0;        // primitive tests per closest-hit ray, likewise
    // per-scene choice between the two bounce kernels: the first full-size frame runs the form guessed from the tree, the
    // second the other one, later frames the faster of the two (bounce + direct ms per closest-hit ray)
    uint64_t tune_sig = 0;           // signature of the scene the statistic above belongs to
    uint32_t bin_threshold = 65536;  // queues at least this long are binned by origin cell / direction octant before the next bounce (GI_BIN_THRESHOLD, 0 = off)
    struct gi_comm* comm = nullptr;  // gi_comm_init: the NCCL communicator of this context (gi_comm.inc)
    DevBuf b_comm, b_stage;          // collective scratch: size word; the root's staging area of gi_framebuffer_gather
    DevBuf b_err;                    // DScene::err: the sticky device error word
    uint32_t dev_err = 0;            // its last read-back (collect_timers)
    int n_sm = 148;                  // multiprocessors of the device (persistent kernels are sized from it)
    unsigned long long work_host[16] = { 0 };   // [0,1] closest nodes/prims, [2,3] any-hit, [4..6] gather depth/cand/sel, [8] rays, [9] shadow rays, [10] queries
    // timing
    std::vector<TimedLaunch> pending;
    std::vector<cudaEvent_t> event_pool;
    std::map<std::string, FamStat> fam;
};

static int fail(gi_ctx* c, int code, const std::string& msg)
{
    if (c) c->err = msg;
    return code;
}
#define CK(call)                                                                                                   \
    do {                                                                                                           \
        cudaError_t e_ = (call);                                                                                   \
        if (e_ != cudaSuccess) return fail(ctx, e_ == cudaErrorMemoryAllocation ? GI_ERR_OOM : GI_ERR_CUDA,       \
                                           std::string(#call) + ": " + cudaGetErrorString(e_));                    \
    } while (0)

// launch KERNEL<FULL, IMPL> for the scene's traversal variant
#define GI_LAUNCH(KERNEL, GRID, BLOCK, ...)                                                                                  \
    do {                                                                                                                     \
        if (ctx->S.full) {                                                                                                   \
            if (ctx->S.implicit_boxes) KERNEL<true, true><<<(GRID), (BLOCK), 0, ctx->stream>>>(__VA_ARGS__);                \
            else KERNEL<true, false><<<(GRID), (BLOCK), 0, ctx->stream>>>(__VA_ARGS__);                                     \
        } else {                                                                                                             \
            if (ctx->S.implicit_boxes) KERNEL<false, true><<<(GRID), (BLOCK), 0, ctx->stream>>>(__VA_ARGS__);               \
            else KERNEL<false, false><<<(GRID), (BLOCK), 0, ctx->stream>>>(__VA_ARGS__);                                    \
        }                                                                                                                    \
    } while (0)

// launch KERNEL<MODE, IMPL> for the wavefront / photon kernels: MODE 2 (atmosphere) when the scene holds fog volumes
#define GI_LAUNCH_M(KERNEL, GRID, BLOCK, ...)                                                                                \
    do {                                                                                                                     \
        if (ctx->S.n_fog) {                                                                                                  \
            if (ctx->S.implicit_boxes) KERNEL<2, true><<<(GRID), (BLOCK), 0, ctx->stream>>>(__VA_ARGS__);                   \
            else KERNEL<2, false><<<(GRID), (BLOCK), 0, ctx->stream>>>(__VA_ARGS__);                                        \
        } else if (ctx->S.full) {                                                                                            \
            if (ctx->S.implicit_boxes) KERNEL<1, true><<<(GRID), (BLOCK), 0, ctx->stream>>>(__VA_ARGS__);                   \
            else KERNEL<1, false><<<(GRID), (BLOCK), 0, ctx->stream>>>(__VA_ARGS__);                                        \
        } else {                                                                                                             \
            if (ctx->S.implicit_boxes) KERNEL<0, true><<<(GRID), (BLOCK), 0, ctx->stream>>>(__VA_ARGS__);                   \
            else KERNEL<0, false><<<(GRID), (BLOCK), 0, ctx->stream>>>(__VA_ARGS__);                                        \
        }                                                                                                                    \
    } while (0)

static inline unsigned grid_for(size_t n, unsigned block) { return (unsigned)((n + block - 1) / block); }

static cudaEvent_t get_event(gi_ctx* ctx)
{
    if (!ctx->event_pool.empty()) { cudaEvent_t e = ctx->event_pool.back(); ctx->event_pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
}
struct ScopedTimer {   // records an event pair around the launches issued in its scope
    gi_ctx* ctx; TimedLaunch t;
    ScopedTimer(gi_ctx* c, const char* fam) : ctx(c) { t.fam = fam; t.a = get_event(c); t.b = get_event(c); cudaEventRecord(t.a, c->stream); }
    ~ScopedTimer() { cudaEventRecord(t.b, ctx->stream); ctx->pending.push_back(t); }
};
struct EventPair {     // a start / stop event pair that goes back to the pool on every exit path (early returns included)
    gi_ctx* ctx; cudaEvent_t a, b;
    explicit EventPair(gi_ctx* c) : ctx(c), a(get_event(c)), b(get_event(c)) {}
    ~EventPair() { ctx->event_pool.push_back(a); ctx->event_pool.push_back(b); }
    EventPair(const EventPair&) = delete;
    EventPair& operator=(const EventPair&) = delete;
    float ms() const { float v = 0; cudaEventElapsedTime(&v, a, b); return v; }
};
static void fam_reset(gi_ctx* ctx, const char* fam) { ctx->fam[fam] = FamStat(); }
static unsigned long long* work_ptr(gi_ctx* ctx, int slot) { return ctx->b_work.as<unsigned long long>() + slot; }
static void collect_timers(gi_ctx* ctx)   // after a stream sync
{
    if (ctx->b_work.p) cudaMemcpy(ctx->work_host, ctx->b_work.p, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    if (ctx->b_err.p) {
        uint32_t e = 0;
        if (cudaMemcpy(&e, ctx->b_err.p, 4, cudaMemcpyDeviceToHost) == cudaSuccess && e) { ctx->dev_err |= e; cudaMemset(ctx->b_err.p, 0, 4); }
    }
    static const bool trace_launches = getenv("GI_TRACE_LAUNCHES") != nullptr;
    for (auto& t : ctx->pending) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, t.a, t.b) == cudaSuccess) { ctx->fam[t.fam].ms += ms; ctx->fam[t.fam].launches++; }
        if (trace_launches) fprintf(stderr, "[gi] %-14s %9.4f ms\n", t.fam.c_str(), ms);
        ctx->event_pool.push_back(t.a); ctx->event_pool.push_back(t.b);
    }
    ctx->pending.clear();
}

// ---- Halton tables (host): Faure permutations + digit-block tables, the rule behind halton_sampler.h:573-603, 890-1414 ----------
static void build_halton(std::vector<uint16_t>& tab, std::vector<DHaltonDim>& dims)
{
    const unsigned max_base = 1619;
    std::vector<std::vector<uint16_t>> perms(max_base + 1);
    for (unsigned k = 1; k <= 3; ++k) { perms[k].resize(k); for (unsigned i = 0; i < k; ++i) perms[k][i] = (uint16_t)i; }
    for (unsigned base = 4; base <= max_base; ++base) {
        perms[base].resize(base);
        const unsigned b = base / 2;
        if (base & 1) {
            for (unsigned i = 0; i < base - 1; ++i) perms[base][i + (i >= b)] = (uint16_t)(perms[base - 1][i] + (perms[base - 1][i] >= b));
            perms[base][b] = (uint16_t)b;
        } else {
            for (unsigned i = 0; i < b; ++i) { perms[base][i] = (uint16_t)(2 * perms[b][i]); perms[base][b + i] = (uint16_t)(2 * perms[b][i] + 1); }
        }
    }
    dims.assign(256, DHaltonDim());
    tab.clear();
    unsigned found = 0;
    for (unsigned c = 2; found < 256; c++) {
        bool prime = true;
        for (unsigned d = 2; d * d <= c; d++) if (c % d == 0) { prime = false; break; }
        if (!prime) continue;
        unsigned digits = 1; uint64_t bk = c;
        while (bk * c <= 500) { bk *= c; digits++; }          // digits per table block (3^5, 5^3, 7^3, 11^2 .. 19^2, then 1)
        unsigned nb = 1; uint64_t pw = bk;
        while (pw * bk <= 0xFFFFFFFFull) { pw *= bk; nb++; }   // blocks so that base^(digits*nb) < 2^32
        DHaltonDim& D = dims[found];
        D.base = c; D.block = (uint32_t)bk; D.nblocks = nb; D.table_off = (uint32_t)tab.size();
        D.scale = (float)(0.9999998807907104 / (double)pw);
        if (found > 0)
            for (uint32_t i = 0; i < bk; i++) {
                uint32_t idx = i, r = 0;
                for (unsigned k = 0; k < digits; k++) { r = r * c + perms[c][idx % c]; idx /= c; }
                tab.push_back((uint16_t)r);
            }
        found++;
    }
}

static DHEnum make_henum(uint32_t width, uint32_t height)   // Halton_enum::Halton_enum (halton_enum.h:69-104)
{
    DHEnum he{};
    uint32_t w = 1; while (w < width) { ++he.p2; w *= 2; }
    uint32_t h = 1; while (h < height) { ++he.p3; h *= 3; }
    he.scale_x = (float)w; he.scale_y = (float)h; he.inc = w * h;
    // extended Euclid for the two modular inverses
    long long a = h, b = w, s0 = 1, s1 = 0, t0 = 0, t1 = 1;
    while (b) { long long q = a / b, r = a % b; a = b; b = r; long long s2 = s0 - q * s1; s0 = s1; s1 = s2; long long t2 = t0 - q * t1; t0 = t1; t1 = t2; }
    long long i1 = s0, i2 = t0;   // h*i1 + w*i2 = 1
    uint32_t inv2 = (i1 < 0) ? (uint32_t)(i1 + (long long)w) : (uint32_t)(i1 % (long long)w);
    uint32_t inv3 = (i2 < 0) ? (uint32_t)(i2 + (long long)h) : (uint32_t)(i2 % (long long)h);
    he.mx = h * inv2; he.my = w * inv3;
    return he;
}

static DFrame make_frame(const gi_ctx* ctx, int w, int h, int x0, int y0, int x1, int y1)
{
    DFrame F{};
    const gi_camera& c = ctx->S.cam;
    F.w = w; F.h = h; F.x0 = x0; F.y0 = y0; F.tw = x1 - x0; F.th = y1 - y0; F.rb = 0; F.rstride = 0;
    F.halfW = (c.sensor_diag * w) / (std::sqrt((double)w * w + h * h));   // raytracer.h:74-75
    F.halfH = F.halfW * ((double)h / w);
    auto v = [](const double* p) { d3 r; r.x = p[0]; r.y = p[1]; r.z = p[2]; return r; };
    F.pos = v(c.pos); F.up = v(c.up);
    d3 fw = v(c.forward);
    F.center.x = F.pos.x + fw.x * c.focal_dist; F.center.y = F.pos.y + fw.y * c.focal_dist; F.center.z = F.pos.z + fw.z * c.focal_dist;   // :77
    d3 cr; cr.x = fw.y * F.up.z - F.up.y * fw.z; cr.y = fw.z * F.up.x - F.up.z * fw.x; cr.z = fw.x * F.up.y - F.up.x * fw.y;               // cross(forward, up)
    double il = 1.0 / std::sqrt(cr.x * cr.x + cr.y * cr.y + cr.z * cr.z);
    F.right.x = cr.x * il; F.right.y = cr.y * il; F.right.z = cr.z * il;                                                                       // :78
    F.he = make_henum((uint32_t)w, (uint32_t)h);
    return F;
}

// ---- lifetime -----------------------------------------------------------------------------------------------------------------------
extern "C" const char* gi_version(void) { return GI_VERSION; }

extern "C" void gi_destroy(gi_ctx* ctx);
extern "C" int gi_comm_destroy(gi_ctx* ctx);
extern "C" int gi_create(int device, gi_ctx** out)
{
    if (!out) return GI_ERR_INVALID;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) return GI_ERR_NO_DEVICE;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return GI_ERR_NO_DEVICE;
    if (prop.major != 10) return GI_ERR_NO_DEVICE;   // the kernels are built for sm_100a only
    if (cudaSetDevice(device) != cudaSuccess) return GI_ERR_NO_DEVICE;
    gi_ctx* ctx = new gi_ctx();
    ctx->device = device;
    // every failure below goes through gi_destroy, which releases whatever exists so far (streams, events, buffers)
    // the main stream carries the critical chain of a frame (bounce kernels, tail) and gets the higher priority: its blocks are placed
    // before the waiting blocks of the shadow / gather launches on the side streams (GI_STREAM_PRIO=0: all equal)
    int prio_least = 0, prio_greatest = 0;
    cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest);
    // GI_STREAM_PRIO: 0 all equal, 1 main above both side streams, 2 main > shadow rays > gathers, 3 main > gathers > shadow rays.
    // Measured (profiles/r02/ab_t22, ab_t24; caustics / glass / cornell frame ms): 1: 19.11 / 135.6 / 44.0, 2: 19.05 / 134.5 / 43.9,
    // 3: 20.33 / 148.3 / 49.8; without priorities the deferred schedule LOSES to the old one (20.98 against 20.54 on caustics).
    int prio_mode = 2;
    if (const char* e = getenv("GI_STREAM_PRIO")) prio_mode = atoi(e);
    if (prio_mode == 0) prio_greatest = prio_least;
    const int prio_mid = prio_least - prio_greatest >= 2 ? prio_least - 1 : prio_least;   // numerically lower = higher priority
    const int side_prio[2] = { prio_mode == 2 ? prio_mid : prio_least, prio_mode == 3 ? prio_mid : prio_least };
    if (cudaStreamCreateWithPriority(&ctx->stream, cudaStreamNonBlocking, prio_greatest) != cudaSuccess) { ctx->stream = nullptr; gi_destroy(ctx); return GI_ERR_CUDA; }
    ctx->main_stream = ctx->stream;
    for (int k = 0; k < 2; k++) {
        // two pairs of side streams: the round-1 schedule wants its shadow rays at the priority of the gather pipeline beside them (below
        // it they starve and spill into the next bounce kernel: cornell 43.9 -> 47.0 ms), the deferred schedule wants both below the main stream
        if (cudaStreamCreateWithPriority(&ctx->side_pairs[0][k], cudaStreamNonBlocking, prio_greatest) != cudaSuccess) { ctx->side_pairs[0][k] = nullptr; gi_destroy(ctx); return GI_ERR_CUDA; }
        if (cudaStreamCreateWithPriority(&ctx->side_pairs[1][k], cudaStreamNonBlocking, side_prio[k]) != cudaSuccess) { ctx->side_pairs[1][k] = nullptr; gi_destroy(ctx); return GI_ERR_CUDA; }
        ctx->side[k] = ctx->side_pairs[1][k];
        for (int q = 0; q < GI_NHL; q++) if (cudaEventCreateWithFlags(&ctx->side_done[q][k], cudaEventDisableTiming) != cudaSuccess) { ctx->side_done[q][k] = nullptr; gi_destroy(ctx); return GI_ERR_CUDA; }
    }
    if (cudaEventCreateWithFlags(&ctx->fork_ev, cudaEventDisableTiming) != cudaSuccess) { ctx->fork_ev = nullptr; gi_destroy(ctx); return GI_ERR_CUDA; }
    if (cudaStreamCreateWithPriority(&ctx->gather_aux, cudaStreamNonBlocking, prio_least) != cudaSuccess) { ctx->gather_aux = nullptr; gi_destroy(ctx); return GI_ERR_CUDA; }
    for (int k = 0; k < 2; k++) if (cudaEventCreateWithFlags(&ctx->gather_ev[k], cudaEventDisableTiming) != cudaSuccess) { ctx->gather_ev[k] = nullptr; gi_destroy(ctx); return GI_ERR_CUDA; }
    if (getenv("GI_NO_GATHER_AUX")) { cudaStreamDestroy(ctx->gather_aux); ctx->gather_aux = nullptr; }   // A/B: the long lists after the thread-per-query kernel, on its stream
    if (const char* e = getenv("GI_OVERLAP_THRESHOLD")) ctx->overlap_threshold = (uint32_t)strtoul(e, nullptr, 10);
    if (const char* e = getenv("GI_SCHED_MODE")) ctx->sched_mode = atoi(e);
    if (const char* e = getenv("GI_RING")) ctx->ring = atoi(e);
    // Halton tables are scene independent
    std::vector<uint16_t> tab; std::vector<DHaltonDim> dims;
    build_halton(tab, dims);
    if (ctx->b_htab.reserve(tab.size() * 2) != cudaSuccess || ctx->b_hdims.reserve(dims.size() * sizeof(DHaltonDim)) != cudaSuccess) { gi_destroy(ctx); return GI_ERR_OOM; }
    cudaMemcpy(ctx->b_htab.p, tab.data(), tab.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(ctx->b_hdims.p, dims.data(), dims.size() * sizeof(DHaltonDim), cudaMemcpyHostToDevice);
    if (ctx->b_work.reserve(16 * sizeof(unsigned long long)) != cudaSuccess || ctx->b_err.reserve(16) != cudaSuccess) { gi_destroy(ctx); return GI_ERR_OOM; }
    cudaMemset(ctx->b_work.p, 0, 16 * sizeof(unsigned long long));
    cudaMemset(ctx->b_err.p, 0, 16);
    ctx->S.err = ctx->b_err.as<uint32_t>();
    ctx->n_sm = prop.multiProcessorCount > 0 ? prop.multiProcessorCount : 148;
    // k_gather_sorted keeps 24 KB of heaps per 64-thread block: nine blocks per SM need the largest shared-memory carve-out
    cudaFuncSetAttribute(k_gather_sorted, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    ctx->S.halton_tab = ctx->b_htab.as<uint16_t>();
    ctx->S.halton_dims = ctx->b_hdims.as<DHaltonDim>();
    if (getenv("GI_NO_IMPLICIT_BOXES")) ctx->no_implicit = true;
    if (const char* e = getenv("GI_TRACE_MODE")) ctx->trace_mode = atoi(e);
    if (const char* e = getenv("GI_TAIL_THRESHOLD")) ctx->tail_threshold = (uint32_t)strtoul(e, nullptr, 10);
    if (const char* e = getenv("GI_BIN_THRESHOLD")) ctx->bin_threshold = (uint32_t)strtoul(e, nullptr, 10);
    if (const char* e = getenv("GI_BOUNCE_MODE")) ctx->bounce_mode = atoi(e);
    if (const char* e = getenv("GI_TAIL_MODE")) ctx->tail_mode = atoi(e);
    *out = ctx;
    return GI_OK;
}

extern "C" void gi_destroy(gi_ctx* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->main_stream) cudaStreamSynchronize(ctx->main_stream);
    gi_comm_destroy(ctx);
    ctx->b_err.release(); ctx->b_comm.release(); ctx->b_stage.release();
    DevBuf* all[] = { &ctx->b_nodes, &ctx->b_refs, &ctx->b_geom, &ctx->b_nrm, &ctx->b_uv, &ctx->b_fnorm, &ctx->b_pmat, &ctx->b_ptype, &ctx->b_mats, &ctx->b_tex, &ctx->b_texpx,
                      &ctx->b_lights, &ctx->b_htab, &ctx->b_hdims, &ctx->b_fogs, &ctx->b_foggrid, &ctx->b_photons, &ctx->b_slab, &ctx->w0, &ctx->w1, &ctx->w2, &ctx->w3, &ctx->w4, &ctx->w5, &ctx->w6, &ctx->w7,
                      &ctx->w8, &ctx->w9, &ctx->b_cnt, &ctx->b_accum, &ctx->b_scan0, &ctx->b_scan1, &ctx->b_misc, &ctx->b_work, &ctx->b_tail, &ctx->b_binkey, &ctx->b_binperm,
                      &ctx->b_binhist, &ctx->b_bincur, &ctx->b_gnode, &ctx->b_gperm, &ctx->b_ghist, &ctx->b_gcur, &ctx->b_gheavy };
    for (DevBuf* b : all) b->release();
    {
        DevBuf* ob[] = { &ctx->ob_type, &ctx->ob_geom, &ctx->ob_bbox, &ctx->ob_nodebox, &ctx->ob_child, &ctx->ob_mask, &ctx->ob_poff, &ctx->ob_pcnt, &ctx->ob_leaf, &ctx->ob_flags, &ctx->ob_pos, &ctx->ob_tot };
        for (DevBuf* b : ob) b->release();
        for (int k = 0; k < 2; k++) { ctx->ob_list[k].release(); ctx->ob_owner[k].release(); ctx->ob_abox[k].release(); ctx->ob_anode[k].release(); ctx->ob_astart[k].release(); ctx->ob_acount[k].release(); ctx->ob_slotactive[k].release(); }
        for (auto& b : ctx->ob_slot) b.release();
        for (auto& b : ctx->ob_rank) b.release();
    }
    for (auto& b : ctx->q_a) b.release();
    for (auto& b : ctx->q_b) b.release();
    for (auto& b : ctx->hl) b.release();
    for (auto& b : ctx->ps) b.release();
    for (auto& b : ctx->tq) b.release();
    for (auto& b : ctx->ad) b.release();
    for (auto& t : ctx->pending) { cudaEventDestroy(t.a); cudaEventDestroy(t.b); }
    for (auto& e : ctx->event_pool) cudaEventDestroy(e);
    for (int k = 0; k < 2; k++) {
        for (int pr = 0; pr < 2; pr++) if (ctx->side_pairs[pr][k]) { cudaStreamSynchronize(ctx->side_pairs[pr][k]); cudaStreamDestroy(ctx->side_pairs[pr][k]); }
        for (int q = 0; q < GI_NHL; q++) if (ctx->side_done[q][k]) cudaEventDestroy(ctx->side_done[q][k]);
    }
    if (ctx->fork_ev) cudaEventDestroy(ctx->fork_ev);
    if (ctx->gather_aux) { cudaStreamSynchronize(ctx->gather_aux); cudaStreamDestroy(ctx->gather_aux); }
    for (int k = 0; k < 2; k++) if (ctx->gather_ev[k]) cudaEventDestroy(ctx->gather_ev[k]);
    for (auto& b : ctx->hl2) b.release();
    for (auto& l : ctx->hlr) for (auto& b : l) b.release();
    for (auto& b : ctx->tsh) b.release();
    ctx->b_scan1s.release();
    if (ctx->main_stream) cudaStreamDestroy(ctx->main_stream);
    delete ctx;
}

// run-time forms of the GI_* environment knobs read at gi_create
extern "C" int gi_configure(gi_ctx* ctx, const char* key, long long value)
{
    if (!ctx || !key) return GI_ERR_INVALID;
    const std::string k = key;
    if (k == "overlap_threshold") ctx->overlap_threshold = (uint32_t)value;
    else if (k == "tail_threshold") ctx->tail_threshold = (uint32_t)value;
    else if (k == "bin_threshold") ctx->bin_threshold = (uint32_t)value;
    else if (k == "bounce_mode") ctx->bounce_mode = (int)value;
    else if (k == "trace_mode") ctx->trace_mode = (int)value;
    else if (k == "tail_mode") ctx->tail_mode = (int)value;
    else if (k == "sched_mode") ctx->sched_mode = (int)value;
    else if (k == "ring") ctx->ring = (int)value;
    else return fail(ctx, GI_ERR_INVALID, "gi_configure: unknown key " + k);
    return GI_OK;
}

extern "C" int gi_cancel(gi_ctx* ctx, int raise)
{
    if (!ctx) return GI_ERR_INVALID;
    ctx->cancel.store(raise ? 1 : 0, std::memory_order_release);
    return GI_OK;
}
// at a launch boundary: drain the stream and give up when the flag is raised
#define GI_POLL_CANCEL(what)                                                                                     \
    do {                                                                                                         \
        if (ctx->cancel.load(std::memory_order_acquire)) {                                                       \
            ctx->stream = ctx->main_stream;                                                                      \
            cudaStreamSynchronize(ctx->side[0]); cudaStreamSynchronize(ctx->side[1]); cudaStreamSynchronize(ctx->stream);                                       \
            collect_timers(ctx);                                                                                 \
            return fail(ctx, GI_ERR_CANCELLED, what " cancelled");                                               \
        }                                                                                                        \
    } while (0)

extern "C" const char* gi_last_error(const gi_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
extern "C" void* gi_stream(gi_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
extern "C" int gi_synchronize(gi_ctx* ctx)
{
    if (!ctx) return GI_ERR_INVALID;
    CK(cudaStreamSynchronize(ctx->stream));
    collect_timers(ctx);
    if (ctx->dev_err & GI_DEV_ERR_STACK) {   // a traversal ran out of stack: the results of the calls since the last synchronisation are not to be trusted
        ctx->dev_err = 0;
        return fail(ctx, GI_ERR_INVALID, "traversal stack overflow (octree too deep for GI_STACK_MAX): results discarded");
    }
    return GI_OK;
}
extern "C" int gi_last_work(gi_ctx* ctx, const char* family, uint64_t out[4])
{
    if (!ctx || !family || !out) return GI_ERR_INVALID;
    std::string f = family;
    const unsigned long long* w = ctx->work_host;
    if (f == "trace_closest") { out[0] = w[8]; out[1] = w[0]; out[2] = w[1]; out[3] = 0; }
    else if (f == "trace_any") { out[0] = w[9]; out[1] = w[2]; out[2] = w[3]; out[3] = 0; }
    else if (f == "gather") { out[0] = w[10]; out[1] = w[4]; out[2] = w[5]; out[3] = w[6]; }
    else return GI_ERR_INVALID;
    return GI_OK;
}
extern "C" int gi_last_kernel_ms(gi_ctx* ctx, const char* family, double* ms, uint64_t* launches)
{
    if (!ctx || !family) return GI_ERR_INVALID;
    auto it = ctx->fam.find(family);
    if (it == ctx->fam.end() || it->second.launches == 0) { if (ms) *ms = 0; if (launches) *launches = 0; return GI_OK; }
    if (ms) *ms = it->second.ms / (double)it->second.launches;
    if (launches) *launches = it->second.launches;
    return GI_OK;
}

// ---- scene upload ---------------------------------------------------------------------------------------------------------------------
template <typename T> static cudaError_t upload(DevBuf& b, const T* src, size_t n, cudaStream_t st)
{
    cudaError_t e = b.reserve(std::max<size_t>(n, 1) * sizeof(T));
    if (e != cudaSuccess) return e;
    if (n) e = cudaMemcpyAsync(b.p, src, n * sizeof(T), cudaMemcpyHostToDevice, st);
    return e;
}

extern "C" int gi_scene_upload(gi_ctx* ctx, const gi_scene_desc* sc)
{
    if (!ctx || !sc) return GI_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    if (sc->n_nodes == 0 || !sc->node_box) return fail(ctx, GI_ERR_INVALID, "scene has no octree nodes (an empty scene still has a root)");
    // validate topology so that a malformed description cannot make a kernel read out of bounds
    for (uint32_t i = 0; i < sc->n_nodes; i++) {
        uint32_t m = sc->node_mask[i];
        if (m) { uint32_t nc = (uint32_t)__builtin_popcount(m); if ((uint64_t)sc->node_child[i] + nc > sc->n_nodes || sc->node_child[i] <= i) return fail(ctx, GI_ERR_INVALID, "node child range out of bounds"); }
        if ((uint64_t)sc->node_prim_off[i] + sc->node_prim_cnt[i] > sc->n_refs) return fail(ctx, GI_ERR_INVALID, "leaf primitive range out of bounds");
    }
    {   // tree depth (children come after their parent, checked above): the traversal stack holds <= 3 entries per level for an
        // ordinary ray (at most four children of a node are met, one is walked into); a tree beyond that bound is refused here,
        // and a ray that still overflows (zero direction components can meet all eight children) raises the sticky device error
        std::vector<uint8_t> depth(sc->n_nodes, 0);
        uint32_t max_depth = 0;
        for (uint32_t i = 0; i < sc->n_nodes; i++) {
            const uint32_t m = sc->node_mask[i];
            if (!m) continue;
            if (depth[i] >= 250) return fail(ctx, GI_ERR_INVALID, "octree deeper than 250 levels");
            const uint32_t nc = (uint32_t)__builtin_popcount(m);
            for (uint32_t c = 0; c < nc; c++) depth[sc->node_child[i] + c] = (uint8_t)(depth[i] + 1);
            max_depth = std::max<uint32_t>(max_depth, depth[i] + 1u);
        }
        if (3 * max_depth + 8 > GI_STACK_MAX) return fail(ctx, GI_ERR_INVALID, "octree too deep for the traversal stack (3 * depth + 8 > GI_STACK_MAX)");
    }
    for (uint32_t i = 0; i < sc->n_refs; i++) if (sc->leaf_prims[i] >= sc->n_prims) return fail(ctx, GI_ERR_INVALID, "leaf primitive id out of bounds");
    for (uint32_t i = 0; i < sc->n_prims; i++) if (sc->prim_mat[i] >= sc->n_mats || sc->prim_type[i] > GI_PRIM_CONE) return fail(ctx, GI_ERR_INVALID, "primitive material/type out of bounds");
    for (uint32_t i = 0; i < sc->n_mats; i++) if (sc->mats[i].diffuse_tex >= sc->n_tex || sc->mats[i].emissive_tex >= sc->n_tex) return fail(ctx, GI_ERR_INVALID, "material texture out of bounds");
    for (uint32_t i = 0; i < sc->n_tex; i++) {
        const gi_texture& t = sc->tex[i];
        if (t.kind == GI_TEX_IMAGE && (t.width <= 0 || t.height <= 0 || t.pixel_offset + (uint64_t)t.width * t.height * 4 > sc->tex_pixel_bytes)) return fail(ctx, GI_ERR_INVALID, "image texture out of bounds or empty");
        if (t.kind < 0 || t.kind > GI_TEX_IMAGE) return fail(ctx, GI_ERR_INVALID, "unknown texture kind");
    }
    for (uint32_t i = 0; i < sc->n_fog; i++) {
        const gi_fog& g = sc->fogs[i];
        if (g.grid_offset + g.grid_count > sc->fog_grid_count) return fail(ctx, GI_ERR_INVALID, "fog noise grid out of bounds");
        if (!(g.size[1] != 0)) return fail(ctx, GI_ERR_INVALID, "fog volume with zero height");
    }
    // nodes: SoA description -> 64-byte records
    std::vector<DNode> nodes(sc->n_nodes);
    for (uint32_t i = 0; i < sc->n_nodes; i++) {
        DNode& n = nodes[i];
        for (int k = 0; k < 3; k++) { n.bmin[k] = sc->node_box[6 * (size_t)i + k]; n.bmax[k] = sc->node_box[6 * (size_t)i + 3 + k]; }
        n.child = sc->node_child[i]; n.mask = sc->node_mask[i]; n.prim_off = sc->node_prim_off[i]; n.prim_cnt = sc->node_prim_cnt[i];
    }
    // do the child boxes follow the partition formula of their parent (octree.cpp:318-328)?  Then traversal derives them.
    bool implicit = true;
    for (uint32_t i = 0; i < sc->n_nodes && implicit; i++) {
        const DNode& n = nodes[i];
        if (!n.mask) continue;
        const double* lo = n.bmin; const double* hi = n.bmax;
        double mid[3], h[3];
        for (int k = 0; k < 3; k++) { mid[k] = lo[k] + .5 * (hi[k] - lo[k]); h[k] = .5 * (hi[k] - lo[k]); }
        // the sequence walk (children_sequence) needs the planes in order: min <= mid <= mid + h and mid <= max
        for (int k = 0; k < 3; k++) if (!(lo[k] <= mid[k] && mid[k] <= mid[k] + h[k] && mid[k] <= hi[k])) implicit = false;
        uint32_t c = n.child;
        for (int ci = 0; ci < 8; ci++) {
            if (!(n.mask & (1u << ci))) continue;
            double cmin[3], cmax[3];
            const int bit[3] = { ci & 1, (ci >> 2) & 1, (ci >> 1) & 1 };   // x: bit0, y: bit2, z: bit1
            for (int k = 0; k < 3; k++) {
                if (ci == 7) { cmin[k] = mid[k]; cmax[k] = hi[k]; }
                else if (bit[k]) { cmin[k] = lo[k] + h[k]; cmax[k] = mid[k] + h[k]; }
                else { cmin[k] = lo[k]; cmax[k] = mid[k]; }
            }
            if (std::memcmp(cmin, nodes[c].bmin, 24) != 0 || std::memcmp(cmax, nodes[c].bmax, 24) != 0) { implicit = false; break; }
            c++;
        }
    }
    if (ctx->no_implicit) implicit = false;
    // leaf references: the primitive's geometry is replicated per (leaf, primitive) occurrence, in leaf order, so that a
    // leaf is one contiguous run of 80-byte records
    bool full = false;
    std::vector<uint8_t> mat_alpha(sc->n_mats, 0);
    for (uint32_t i = 0; i < sc->n_mats; i++) {
        const gi_material& m = sc->mats[i];
        const gi_texture& t = sc->tex[m.diffuse_tex];
        bool may_fail = (m.ior == 1) && (m.opacity < 1.0 || (t.kind == GI_TEX_IMAGE && t.has_alpha));
        mat_alpha[i] = may_fail ? 1 : 0;
    }
    std::vector<uint32_t> pflags(sc->n_prims);
    for (uint32_t i = 0; i < sc->n_prims; i++) {
        uint32_t f = sc->prim_type[i];
        bool writes_uv = false;
        if (sc->prim_type[i] == GI_PRIM_TRIANGLE) {
            const double* nn = sc->prim_nrm + 9 * (size_t)i;
            auto l2 = [](const double* v) { return v[0] * v[0] + v[1] * v[1] + v[2] * v[2]; };
            writes_uv = l2(nn) > 0 && l2(nn + 3) > 0 && l2(nn + 6) > 0;
        } else if (sc->prim_type[i] == GI_PRIM_SPHERE) writes_uv = true;
        if (writes_uv) f |= LF_WRITES_UV; else full = true;
        if (mat_alpha[sc->prim_mat[i]]) { f |= LF_ALPHA; full = true; }
        pflags[i] = f;
    }
    cudaStream_t st = ctx->stream;
    CK(upload(ctx->b_nodes, nodes.data(), nodes.size(), st));
    CK(upload(ctx->b_geom, sc->prim_geom, (size_t)sc->n_prims * 9, st));
    CK(upload(ctx->b_nrm, sc->prim_nrm, (size_t)sc->n_prims * 9, st));
    CK(upload(ctx->b_uv, sc->prim_uv, (size_t)sc->n_prims * 6, st));
    CK(upload(ctx->b_fnorm, sc->prim_fnorm, (size_t)sc->n_prims * 3, st));
    CK(upload(ctx->b_pmat, sc->prim_mat, (size_t)sc->n_prims, st));
    CK(upload(ctx->b_ptype, sc->prim_type, (size_t)sc->n_prims, st));
    CK(upload(ctx->b_mats, sc->mats, (size_t)sc->n_mats, st));
    CK(upload(ctx->b_tex, sc->tex, (size_t)sc->n_tex, st));
    CK(upload(ctx->b_texpx, sc->tex_pixels, (size_t)sc->tex_pixel_bytes, st));
    CK(upload(ctx->b_lights, sc->lights, (size_t)sc->n_lights, st));
    CK(upload(ctx->b_fogs, sc->fogs, (size_t)sc->n_fog, st));
    CK(upload(ctx->b_foggrid, sc->fog_grid, (size_t)(sc->n_fog ? sc->fog_grid_count : 0), st));
    // leaf records: expanded on the device (k_build_leafrefs) from what was just uploaded + the leaf index list and the per-primitive flags
    CK(ctx->b_refs.reserve(std::max<size_t>(sc->n_refs, 1) * sizeof(DLeafRef)));
    if (sc->n_refs) {
        CK(upload(ctx->w8, sc->leaf_prims, (size_t)sc->n_refs, st));
        CK(upload(ctx->w9, pflags.data(), pflags.size(), st));
        k_build_leafrefs<<<grid_for(sc->n_refs, 256), 256, 0, st>>>(sc->n_refs, ctx->w8.as<uint32_t>(), ctx->b_geom.as<double>(), ctx->b_ptype.as<uint8_t>(), ctx->w9.as<uint32_t>(),
                                                                   ctx->b_refs.as<DLeafRef>());
        CK(cudaGetLastError());
    }
    CK(cudaStreamSynchronize(st));   // the host staging vectors die at return
    DScene& S = ctx->S;
    S.nodes = ctx->b_nodes.as<DNode>(); S.refs = ctx->b_refs.as<DLeafRef>();
    S.n_nodes = sc->n_nodes; S.n_refs = sc->n_refs; S.n_prims = sc->n_prims;
    S.prim_geom = ctx->b_geom.as<double>(); S.prim_nrm = ctx->b_nrm.as<double>(); S.prim_uv = ctx->b_uv.as<double>(); S.prim_fnorm = ctx->b_fnorm.as<double>();
    S.prim_mat = ctx->b_pmat.as<uint32_t>(); S.prim_type = ctx->b_ptype.as<uint8_t>();
    S.mats = ctx->b_mats.as<gi_material>(); S.tex = ctx->b_tex.as<gi_texture>(); S.tex_pixels = ctx->b_texpx.as<uint8_t>();
    S.lights = ctx->b_lights.as<gi_light>(); S.n_lights = sc->n_lights; S.n_mats = sc->n_mats; S.n_tex = sc->n_tex;
    S.n_fog = sc->n_fog; S.fogs = ctx->b_fogs.as<gi_fog>(); S.fog_grid = ctx->b_foggrid.as<double>();
    S.cam = sc->camera;
    for (int k = 0; k < 3; k++) S.ambient[k] = sc->ambient[k];
    S.full = full ? 1u : 0u;
    S.implicit_boxes = implicit ? 1u : 0u;
    for (int k = 0; k < 6; k++) ctx->root_box[k] = sc->node_box[k];
    {
        uint64_t sig = gi_mix64(((uint64_t)sc->n_nodes << 32) ^ sc->n_refs) ^ gi_mix64(((uint64_t)sc->n_prims << 20) ^ sc->n_lights);
        for (int k = 0; k < 6; k++) { uint64_t b; std::memcpy(&b, &sc->node_box[k], 8); sig = gi_mix64(sig ^ b); }
        if (sig != ctx->tune_sig) { ctx->tune_sig = sig; ctx->nodes_per_ray = 0; ctx->prims_per_ray = 0; ctx->sched_hint = 1; }   // a new scene: forget the previous one's traversal statistics
    }
    ctx->has_scene = true;   // photons / photon map are independent state and survive a re-upload (the reference keeps its
    return GI_OK;            // map across run() calls, raytracer.h:61); rebuild it explicitly when the geometry changed
}

extern "C" int gi_scene_info(gi_ctx* ctx, uint32_t out[4])
{
    if (!ctx || !out) return GI_ERR_INVALID;
    if (!ctx->has_scene) return fail(ctx, GI_ERR_NO_SCENE, "no scene");
    out[0] = ctx->S.full; out[1] = ctx->S.implicit_boxes; out[2] = ctx->S.n_nodes; out[3] = ctx->S.n_refs;
    return GI_OK;
}

// ---- materials, batch form (host pointers) ------------------------------------------------------------------------------------------
extern "C" int gi_material_eval(gi_ctx* ctx, size_t n, const uint32_t* prim, const double* uv, double* diffuse, double* emissive, double* alpha)
{
    if (!ctx || (n && (!prim || !uv || !diffuse || !emissive || !alpha))) return GI_ERR_INVALID;
    if (!ctx->has_scene) return fail(ctx, GI_ERR_NO_SCENE, "gi_material_eval needs gi_scene_upload");
    if (!n) return GI_OK;
    CK(cudaSetDevice(ctx->device));
    CK(ctx->w0.reserve(n * 4)); CK(ctx->w1.reserve(n * 16)); CK(ctx->w2.reserve(n * 24)); CK(ctx->w3.reserve(n * 24)); CK(ctx->w4.reserve(n * 8));
    CK(cudaMemcpyAsync(ctx->w0.p, prim, n * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->w1.p, uv, n * 16, cudaMemcpyHostToDevice, ctx->stream));
    k_material_eval<<<grid_for(n, 256), 256, 0, ctx->stream>>>(ctx->S, n, ctx->w0.as<uint32_t>(), ctx->w1.as<double>(), ctx->w2.as<double>(), ctx->w3.as<double>(), ctx->w4.as<double>());
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(diffuse, ctx->w2.p, n * 24, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(emissive, ctx->w3.p, n * 24, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(alpha, ctx->w4.p, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    return gi_synchronize(ctx);
}

// ---- atmosphere, batch forms (host pointers) -------------------------------------------------------------------------------------------
extern "C" int gi_fog_density(gi_ctx* ctx, size_t n, const double* pos, double* dens, double* col)
{
    if (!ctx || (n && (!pos || !dens))) return GI_ERR_INVALID;
    if (!ctx->has_scene) return fail(ctx, GI_ERR_NO_SCENE, "gi_fog_density needs gi_scene_upload");
    if (!n) return GI_OK;
    CK(cudaSetDevice(ctx->device));
    CK(ctx->w0.reserve(n * 24)); CK(ctx->w1.reserve(n * 8)); CK(ctx->w2.reserve(n * 24));
    CK(cudaMemcpyAsync(ctx->w0.p, pos, n * 24, cudaMemcpyHostToDevice, ctx->stream));
    k_fog_density<<<grid_for(n, 128), 128, 0, ctx->stream>>>(ctx->S, n, ctx->w0.as<double>(), ctx->w1.as<double>(), ctx->w2.as<double>());
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(dens, ctx->w1.p, n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (col) CK(cudaMemcpyAsync(col, ctx->w2.p, n * 24, cudaMemcpyDeviceToHost, ctx->stream));
    return gi_synchronize(ctx);
}
extern "C" int gi_raymarch(gi_ctx* ctx, size_t n, const double* org, const double* dir, const double* tmax, uint64_t seed, int march, uint8_t* hit, double* t0, double* t1, double* pos, double* col)
{
    if (!ctx || (n && (!org || !dir || !tmax || !hit))) return GI_ERR_INVALID;
    if (!ctx->has_scene) return fail(ctx, GI_ERR_NO_SCENE, "gi_raymarch needs gi_scene_upload");
    if (!n) return GI_OK;
    CK(cudaSetDevice(ctx->device));
    CK(ctx->w0.reserve(n * 24)); CK(ctx->w1.reserve(n * 24)); CK(ctx->w2.reserve(n * 8)); CK(ctx->w3.reserve(n)); CK(ctx->w4.reserve(n * 16)); CK(ctx->w5.reserve(n * 24)); CK(ctx->w6.reserve(n * 24));
    CK(cudaMemcpyAsync(ctx->w0.p, org, n * 24, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->w1.p, dir, n * 24, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->w2.p, tmax, n * 8, cudaMemcpyHostToDevice, ctx->stream));
    k_raymarch<<<grid_for(n, 128), 128, 0, ctx->stream>>>(ctx->S, n, ctx->w0.as<double>(), ctx->w1.as<double>(), ctx->w2.as<double>(), seed, march, ctx->w3.as<uint8_t>(), ctx->w4.as<double>(), ctx->w5.as<double>(),
                                                          ctx->w6.as<double>());
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(hit, ctx->w3.p, n, cudaMemcpyDeviceToHost, ctx->stream));
    if (t0) CK(cudaMemcpy2DAsync(t0, 8, ctx->w4.p, 16, 8, n, cudaMemcpyDeviceToHost, ctx->stream));
    if (t1) CK(cudaMemcpy2DAsync(t1, 8, (const char*)ctx->w4.p + 8, 16, 8, n, cudaMemcpyDeviceToHost, ctx->stream));
    if (pos) CK(cudaMemcpyAsync(pos, ctx->w5.p, n * 24, cudaMemcpyDeviceToHost, ctx->stream));
    if (col) CK(cudaMemcpyAsync(col, ctx->w6.p, n * 24, cudaMemcpyDeviceToHost, ctx->stream));
    return gi_synchronize(ctx);
}

// ---- Halton entry points --------------------------------------------------------------------------------------------------------------
extern "C" int gi_halton_sample(gi_ctx* ctx, size_t n, const uint32_t* dim, const uint32_t* index, float* out)
{
    if (!ctx || (n && (!dim || !index || !out))) return GI_ERR_INVALID;
    if (!n) return GI_OK;
    CK(cudaSetDevice(ctx->device));
    CK(ctx->w0.reserve(n * 4)); CK(ctx->w1.reserve(n * 4)); CK(ctx->w2.reserve(n * 4));
    CK(cudaMemcpyAsync(ctx->w0.p, dim, n * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->w1.p, index, n * 4, cudaMemcpyHostToDevice, ctx->stream));
    k_halton_sample<<<grid_for(n, 256), 256, 0, ctx->stream>>>(ctx->S, n, ctx->w0.as<uint32_t>(), ctx->w1.as<uint32_t>(), ctx->w2.as<float>());
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, ctx->w2.p, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return GI_OK;
}

extern "C" int gi_halton_index(gi_ctx* ctx, int width, int height, size_t n, const uint32_t* s, const uint32_t* x, const uint32_t* y, uint32_t* out)
{
    if (!ctx || width <= 0 || height <= 0 || (n && (!s || !x || !y || !out))) return GI_ERR_INVALID;
    if (!n) return GI_OK;
    CK(cudaSetDevice(ctx->device));
    CK(ctx->w0.reserve(n * 4)); CK(ctx->w1.reserve(n * 4)); CK(ctx->w2.reserve(n * 4)); CK(ctx->w3.reserve(n * 4));
    CK(cudaMemcpyAsync(ctx->w0.p, s, n * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->w1.p, x, n * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->w2.p, y, n * 4, cudaMemcpyHostToDevice, ctx->stream));
    k_halton_index<<<grid_for(n, 256), 256, 0, ctx->stream>>>(make_henum((uint32_t)width, (uint32_t)height), n, ctx->w0.as<uint32_t>(), ctx->w1.as<uint32_t>(), ctx->w2.as<uint32_t>(), ctx->w3.as<uint32_t>());
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, ctx->w3.p, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return GI_OK;
}

extern "C" int gi_camera_rays(gi_ctx* ctx, int width, int height, int x0, int y0, int x1, int y1, int s0, int s1, double* org, double* dir, uint32_t* index)
{
    if (!ctx || width <= 0 || height <= 0 || x1 <= x0 || y1 <= y0 || s1 <= s0 || !org || !dir) return GI_ERR_INVALID;
    if (!ctx->has_scene) return fail(ctx, GI_ERR_NO_SCENE, "gi_camera_rays needs gi_scene_upload (camera)");
    CK(cudaSetDevice(ctx->device));
    size_t n = (size_t)(x1 - x0) * (y1 - y0) * (s1 - s0);
    CK(ctx->w0.reserve(n * 24)); CK(ctx->w1.reserve(n * 24)); CK(ctx->w2.reserve(n * 4));
    DFrame F = make_frame(ctx, width, height, x0, y0, x1, y1);
    k_camera_rays<<<grid_for(n, 256), 256, 0, ctx->stream>>>(ctx->S, F, s0, n, ctx->w0.as<double>(), ctx->w1.as<double>(), ctx->w2.as<uint32_t>());
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(org, ctx->w0.p, n * 24, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(dir, ctx->w1.p, n * 24, cudaMemcpyDeviceToHost, ctx->stream));
    if (index) CK(cudaMemcpyAsync(index, ctx->w2.p, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return GI_OK;
}

// ---- closest / any hit ------------------------------------------------------------------------------------------------------------------
extern "C" int gi_trace_closest_dev(gi_ctx* ctx, size_t n, const double* org, const double* dir, uint64_t alpha_seed, uint32_t* prim, double* hit, double* normal, double* uv)
{
    if (!ctx || (n && (!org || !dir))) return GI_ERR_INVALID;
    if (!ctx->has_scene) return fail(ctx, GI_ERR_NO_SCENE, "gi_trace_closest needs gi_scene_upload");
    if (!n) return GI_OK;
    CK(cudaSetDevice(ctx->device));
    fam_reset(ctx, "trace_closest");
    CK(cudaMemsetAsync(work_ptr(ctx, 0), 0, 16, ctx->stream));
    ctx->work_host[8] = n;
    {
        ScopedTimer t(ctx, "trace_closest");
        if (ctx->trace_mode == 1) {
            GI_LAUNCH(k_trace_closest_w, grid_for(n, GI_WPB), GI_WPB * 32, ctx->S, n, org, dir, alpha_seed, prim, hit, normal, uv, work_ptr(ctx, 0));
            
        } else GI_LAUNCH(k_trace_closest, grid_for(n, GI_BLOCK), GI_BLOCK, ctx->S, n, org, dir, alpha_seed, prim, hit, normal, uv, work_ptr(ctx, 0));
    }
    CK(cudaGetLastError());
    return GI_OK;
}
extern "C" int gi_trace_closest(gi_ctx* ctx, size_t n, const double* org, const double* dir, uint64_t alpha_seed, uint32_t* prim, double* hit, double* normal, double* uv)
{
    if (!ctx || (n && (!org || !dir))) return GI_ERR_INVALID;
    if (!ctx->has_scene) return fail(ctx, GI_ERR_NO_SCENE, "gi_trace_closest needs gi_scene_upload");
    if (!n) return GI_OK;
    CK(cudaSetDevice(ctx->device));
    CK(ctx->w0.reserve(n * 24)); CK(ctx->w1.reserve(n * 24)); CK(ctx->w2.reserve(n * 4)); CK(ctx->w3.reserve(n * 24)); CK(ctx->w4.reserve(n * 24)); CK(ctx->w5.reserve(n * 16));
    CK(cudaMemcpyAsync(ctx->w0.p, org, n * 24, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->w1.p, dir, n * 24, cudaMemcpyHostToDevice, ctx->stream));
    int rc = gi_trace_closest_dev(ctx, n, ctx->w0.as<double>(), ctx->w1.as<double>(), alpha_seed, ctx->w2.as<uint32_t>(), hit ? ctx->w3.as<double>() : nullptr,
                                  normal ? ctx->w4.as<double>() : nullptr, uv ? ctx->w5.as<double>() : nullptr);
    if (rc != GI_OK) return rc;
    if (prim) CK(cudaMemcpyAsync(prim, ctx->w2.p, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (hit) CK(cudaMemcpyAsync(hit, ctx->w3.p, n * 24, cudaMemcpyDeviceToHost, ctx->stream));
    if (normal) CK(cudaMemcpyAsync(normal, ctx->w4.p, n * 24, cudaMemcpyDeviceToHost, ctx->stream));
    if (uv) CK(cudaMemcpyAsync(uv, ctx->w5.p, n * 16, cudaMemcpyDeviceToHost, ctx->stream));
    return gi_synchronize(ctx);
}

extern "C" int gi_trace_any_dev(gi_ctx* ctx, size_t n, const double* org, const double* dir, const double* maxt2, uint64_t alpha_seed, uint8_t* vis)
{
    if (!ctx || (n && (!org || !dir || !maxt2 || !vis))) return GI_ERR_INVALID;
    if (!ctx->has_scene) return fail(ctx, GI_ERR_NO_SCENE, "gi_trace_any needs gi_scene_upload");
    if (!n) return GI_OK;
    CK(cudaSetDevice(ctx->device));
    fam_reset(ctx, "trace_any");
    CK(cudaMemsetAsync(work_ptr(ctx, 2), 0, 16, ctx->stream));
    ctx->work_host[9] = n;
    {
        ScopedTimer t(ctx, "trace_any");
        if (ctx->trace_mode == 1) {
            GI_LAUNCH(k_trace_any_w, grid_for(n, GI_WPB), GI_WPB * 32, ctx->S, n, org, dir, maxt2, alpha_seed, vis, work_ptr(ctx, 2));
            
        } else GI_LAUNCH(k_trace_any, grid_for(n, GI_BLOCK), GI_BLOCK, ctx->S, n, org, dir, maxt2, alpha_seed, vis, work_ptr(ctx, 2));
    }
    CK(cudaGetLastError());
    return GI_OK;
}
extern "C" int gi_trace_any(gi_ctx* ctx, size_t n, const double* org, const double* dir, const double* maxt2, uint64_t alpha_seed, uint8_t* vis)
{
    if (!ctx || (n && (!org || !dir || !maxt2 || !vis))) return GI_ERR_INVALID;
    if (!ctx->has_scene) return fail(ctx, GI_ERR_NO_SCENE, "gi_trace_any needs gi_scene_upload");
    if (!n) return GI_OK;
    CK(cudaSetDevice(ctx->device));
    CK(ctx->w0.reserve(n * 24)); CK(ctx->w1.reserve(n * 24)); CK(ctx->w2.reserve(n * 8)); CK(ctx->w3.reserve(n));
    CK(cudaMemcpyAsync(ctx->w0.p, org, n * 24, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->w1.p, dir, n * 24, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->w2.p, maxt2, n * 8, cudaMemcpyHostToDevice, ctx->stream));
    int rc = gi_trace_any_dev(ctx, n, ctx->w0.as<double>(), ctx->w1.as<double>(), ctx->w2.as<double>(), alpha_seed, ctx->w3.as<uint8_t>());
    if (rc != GI_OK) return rc;
    CK(cudaMemcpyAsync(vis, ctx->w3.p, n, cudaMemcpyDeviceToHost, ctx->stream));
    return gi_synchronize(ctx);
}

// ---- device-side exclusive scan helper (k_scan_*) -----------------------------------------------------------------------------------------
static int scan_exclusive(gi_ctx* ctx, const uint32_t* in, int stride_words, uint32_t n, uint32_t* out, uint32_t* total_dev)
{
    uint32_t nb = (n + GI_SCAN_BLOCK - 1) / GI_SCAN_BLOCK;
    DevBuf& scratch = ctx->stream == ctx->main_stream ? ctx->b_scan1 : ctx->b_scan1s;   // a scan on a side stream may run beside one on the main stream
    CK(scratch.reserve((size_t)std::max<uint32_t>(nb, 1) * 4));
    if (n) {
        k_scan_block<<<nb, GI_SCAN_BLOCK, 0, ctx->stream>>>(in, n, out, scratch.as<uint32_t>(), stride_words);
        k_scan_sums<<<1, GI_SCAN_BLOCK, 0, ctx->stream>>>(scratch.as<uint32_t>(), nb, total_dev);
        k_scan_apply<<<nb, GI_SCAN_BLOCK, 0, ctx->stream>>>(out, n, scratch.as<uint32_t>());
    } else if (total_dev) CK(cudaMemsetAsync(total_dev, 0, 4, ctx->stream));
    CK(cudaGetLastError());
    return GI_OK;
}

// ---- scene octree build on the device (gi_octree_build.cuh) ---------------------------------------------------------------------------------
// grow a device buffer and keep its first `keep` bytes
static cudaError_t grow_keep(gi_ctx* ctx, DevBuf& b, size_t bytes, size_t keep)
{
    if (bytes <= b.cap) return cudaSuccess;
    void* np = nullptr;
    const size_t want = bytes * 2 + 256;
    cudaError_t e = cudaMalloc(&np, want);
    if (e != cudaSuccess) return e;
    if (keep && b.p) { e = cudaMemcpyAsync(np, b.p, keep, cudaMemcpyDeviceToDevice, ctx->stream); if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream); }
    if (b.p) cudaFree(b.p);
    b.p = np; b.cap = want;
    return e;
}
extern "C" int gi_octree_build(gi_ctx* ctx, uint32_t n_prims, const uint8_t* prim_type, const double* prim_geom, const double* prim_bbox, const double* root_box6, uint32_t* n_nodes_out,
                               uint32_t* n_refs_out, double* build_ms)
{
    if (!ctx || !root_box6 || (n_prims && (!prim_type || !prim_geom || !prim_bbox))) return GI_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    ctx->ob_valid = false;
    CK(ctx->ob_type.reserve(std::max<size_t>(n_prims, 1))); CK(ctx->ob_geom.reserve(std::max<size_t>(n_prims, 1) * 72)); CK(ctx->ob_bbox.reserve(std::max<size_t>(n_prims, 1) * 48));
    if (n_prims) {
        CK(cudaMemcpyAsync(ctx->ob_type.p, prim_type, n_prims, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(ctx->ob_geom.p, prim_geom, (size_t)n_prims * 72, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(ctx->ob_bbox.p, prim_bbox, (size_t)n_prims * 48, cudaMemcpyHostToDevice, st));
    }
    EventPair ev(ctx); const cudaEvent_t e0 = ev.a, e1 = ev.b;
    cudaEventRecord(e0, st);
    // root (octree.cpp:25-38 grew its box; :106: it is split only when it holds more than 16 entities)
    uint32_t n_nodes = 1, leaf_base = 0;
    CK(grow_keep(ctx, ctx->ob_nodebox, 48, 0)); CK(grow_keep(ctx, ctx->ob_child, 4, 0)); CK(grow_keep(ctx, ctx->ob_mask, 1, 0)); CK(grow_keep(ctx, ctx->ob_poff, 4, 0)); CK(grow_keep(ctx, ctx->ob_pcnt, 4, 0));
    CK(grow_keep(ctx, ctx->ob_leaf, std::max<size_t>(n_prims, 1) * 4, 0));
    if (n_prims > GI_OB_MAX_LEAF) {
        // first guess at the final size (grown when a level needs more): ~4 nodes and ~16 leaf references per primitive
        const size_t gn = (size_t)n_prims * 4, gr = (size_t)n_prims * 16;
        CK(grow_keep(ctx, ctx->ob_nodebox, gn * 24, 0)); CK(grow_keep(ctx, ctx->ob_child, gn * 2, 0)); CK(grow_keep(ctx, ctx->ob_mask, gn / 2, 0)); CK(grow_keep(ctx, ctx->ob_poff, gn * 2, 0));
        CK(grow_keep(ctx, ctx->ob_pcnt, gn * 2, 0)); CK(grow_keep(ctx, ctx->ob_leaf, gr * 2, 0));
        for (int k = 0; k < 2; k++) { CK(ctx->ob_list[k].reserve(gr * 2)); CK(ctx->ob_owner[k].reserve(gr * 2)); }
        CK(ctx->ob_flags.reserve(gr * 16)); CK(ctx->ob_pos.reserve(gr * 16));
    }
    static const bool trace_build = getenv("GI_TRACE_BUILD") != nullptr;
    auto wall0 = std::chrono::steady_clock::now();
    CK(cudaMemcpyAsync(ctx->ob_nodebox.p, root_box6, 48, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(ctx->ob_child.p, 0, 4, st)); CK(cudaMemsetAsync(ctx->ob_mask.p, 0, 1, st)); CK(cudaMemsetAsync(ctx->ob_poff.p, 0, 4, st));
    CK(cudaMemcpyAsync(ctx->ob_pcnt.p, &n_prims, 4, cudaMemcpyHostToDevice, st));
    if (n_prims) k_ob_iota<<<grid_for(n_prims, 256), 256, 0, st>>>(n_prims, ctx->ob_leaf.as<uint32_t>());   // a root that stays a leaf owns every entity in insertion order
    uint32_t n_items = n_prims, n_active = n_prims > GI_OB_MAX_LEAF ? 1u : 0u;
    int cur = 0;
    if (n_active) {
        CK(ctx->ob_list[0].reserve((size_t)n_items * 4));
        k_ob_iota<<<grid_for(n_items, 256), 256, 0, st>>>(n_items, ctx->ob_list[0].as<uint32_t>());
        CK(ctx->ob_abox[0].reserve(48)); CK(ctx->ob_anode[0].reserve(4)); CK(ctx->ob_astart[0].reserve(4)); CK(ctx->ob_acount[0].reserve(4));
        const uint32_t zero = 0;
        CK(cudaMemcpyAsync(ctx->ob_abox[0].p, root_box6, 48, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(ctx->ob_anode[0].p, &zero, 4, cudaMemcpyHostToDevice, st)); CK(cudaMemcpyAsync(ctx->ob_astart[0].p, &zero, 4, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(ctx->ob_acount[0].p, &n_items, 4, cudaMemcpyHostToDevice, st));
        leaf_base = 0;
    } else leaf_base = n_prims;
    CK(ctx->ob_tot.reserve(16));
    int level = 0;
    while (n_active) {
        if ((uint64_t)n_items * 8ull + 1ull > 0xFFFFFFF0ull) return fail(ctx, GI_ERR_INVALID, "octree level too large for 32-bit positions");
        DOBLevel L{};
        L.n_items = n_items; L.n_active = n_active;
        L.list = ctx->ob_list[cur].as<uint32_t>(); L.owner = level ? ctx->ob_owner[cur].as<uint32_t>() : nullptr; L.slot_active = level ? ctx->ob_slotactive[cur].as<uint32_t>() : nullptr;
        L.a_box = ctx->ob_abox[cur].as<double>(); L.a_node = ctx->ob_anode[cur].as<uint32_t>(); L.a_start = ctx->ob_astart[cur].as<uint32_t>(); L.a_count = ctx->ob_acount[cur].as<uint32_t>();
        const size_t nflag = (size_t)n_items * 8 + 1;
        CK(ctx->ob_flags.reserve(nflag * 4)); CK(ctx->ob_pos.reserve(nflag * 4));
        k_ob_classify<<<grid_for(n_items, 128), 128, 0, st>>>(L, ctx->ob_type.as<uint8_t>(), ctx->ob_geom.as<double>(), ctx->ob_bbox.as<double>(), ctx->ob_flags.as<uint32_t>());
        uint32_t* tot = ctx->ob_tot.as<uint32_t>();
        int rc = scan_exclusive(ctx, ctx->ob_flags.as<uint32_t>(), 1, (uint32_t)nflag, ctx->ob_pos.as<uint32_t>(), tot);
        if (rc != GI_OK) return rc;
        const size_t nslot = (size_t)n_active * 8;
        for (auto& b : ctx->ob_slot) CK(b.reserve(nslot * 4));
        for (auto& b : ctx->ob_rank) CK(b.reserve(nslot * 4));
        DOBSlots S{ ctx->ob_slot[0].as<uint32_t>(), ctx->ob_slot[1].as<uint32_t>(), ctx->ob_slot[2].as<uint32_t>(), ctx->ob_slot[3].as<uint32_t>(), ctx->ob_slot[4].as<uint32_t>() };
        k_ob_slots<<<grid_for(n_active, 128), 128, 0, st>>>(L, ctx->ob_pos.as<uint32_t>(), S);
        if ((rc = scan_exclusive(ctx, S.exists, 1, (uint32_t)nslot, ctx->ob_rank[0].as<uint32_t>(), tot + 1)) != GI_OK) return rc;
        if ((rc = scan_exclusive(ctx, S.cont, 1, (uint32_t)nslot, ctx->ob_rank[1].as<uint32_t>(), tot + 2)) != GI_OK) return rc;
        if ((rc = scan_exclusive(ctx, S.final_cnt, 1, (uint32_t)nslot, ctx->ob_rank[2].as<uint32_t>(), tot + 3)) != GI_OK) return rc;
        uint32_t h[4];
        CK(cudaMemcpyAsync(h, tot, 16, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        const uint32_t next_items = h[0], new_nodes = h[1], next_active = h[2], final_refs = h[3];
        if ((uint64_t)n_nodes + new_nodes > 0xFFFFFFF0ull || (uint64_t)leaf_base + final_refs > 0xFFFFFFF0ull) return fail(ctx, GI_ERR_INVALID, "octree too large");
        const size_t nn = (size_t)n_nodes + new_nodes;
        CK(grow_keep(ctx, ctx->ob_nodebox, nn * 48, (size_t)n_nodes * 48)); CK(grow_keep(ctx, ctx->ob_child, nn * 4, (size_t)n_nodes * 4)); CK(grow_keep(ctx, ctx->ob_mask, nn, n_nodes));
        CK(grow_keep(ctx, ctx->ob_poff, nn * 4, (size_t)n_nodes * 4)); CK(grow_keep(ctx, ctx->ob_pcnt, nn * 4, (size_t)n_nodes * 4));
        CK(grow_keep(ctx, ctx->ob_leaf, ((size_t)leaf_base + final_refs + 1) * 4, (size_t)leaf_base * 4));
        const int nxt = cur ^ 1;
        CK(ctx->ob_list[nxt].reserve(std::max<size_t>(next_items, 1) * 4)); CK(ctx->ob_owner[nxt].reserve(std::max<size_t>(next_items, 1) * 4));
        CK(ctx->ob_abox[nxt].reserve(std::max<size_t>(next_active, 1) * 48)); CK(ctx->ob_anode[nxt].reserve(std::max<size_t>(next_active, 1) * 4));
        CK(ctx->ob_astart[nxt].reserve(std::max<size_t>(next_active, 1) * 4)); CK(ctx->ob_acount[nxt].reserve(std::max<size_t>(next_active, 1) * 4));
        CK(ctx->ob_slotactive[nxt].reserve(nslot * 4));
        DOBOut O{ ctx->ob_nodebox.as<double>(), ctx->ob_child.as<uint32_t>(), ctx->ob_mask.as<uint8_t>(), ctx->ob_poff.as<uint32_t>(), ctx->ob_pcnt.as<uint32_t>(), ctx->ob_leaf.as<uint32_t>() };
        DOBNext N{ ctx->ob_abox[nxt].as<double>(), ctx->ob_anode[nxt].as<uint32_t>(), ctx->ob_astart[nxt].as<uint32_t>(), ctx->ob_acount[nxt].as<uint32_t>(), ctx->ob_slotactive[nxt].as<uint32_t>() };
        k_ob_emit<<<grid_for(nslot, 128), 128, 0, st>>>(L, S, ctx->ob_rank[0].as<uint32_t>(), ctx->ob_rank[1].as<uint32_t>(), ctx->ob_rank[2].as<uint32_t>(), n_nodes, leaf_base, O, N);
        k_ob_scatter<<<grid_for(n_items, 128), 128, 0, st>>>(L, ctx->ob_flags.as<uint32_t>(), ctx->ob_pos.as<uint32_t>(), S, ctx->ob_rank[2].as<uint32_t>(), leaf_base, ctx->ob_list[nxt].as<uint32_t>(),
                                                              ctx->ob_owner[nxt].as<uint32_t>(), ctx->ob_leaf.as<uint32_t>());
        CK(cudaGetLastError());
        if (trace_build) {
            cudaStreamSynchronize(st);
            auto w1 = std::chrono::steady_clock::now();
            fprintf(stderr, "[gi] octree level %2d: %9u items %8u active -> %8u new nodes, %9u final refs, %9u items next | %.3f ms\n", level, n_items, n_active, new_nodes, final_refs, next_items,
                    std::chrono::duration<double, std::milli>(w1 - wall0).count());
            wall0 = w1;
        }
        n_nodes += new_nodes; leaf_base += final_refs; n_items = next_items; n_active = next_active; cur = nxt; level++;
        if (level > 64) return fail(ctx, GI_ERR_INVALID, "octree deeper than 64 levels");
    }
    cudaEventRecord(e1, st);
    CK(cudaStreamSynchronize(st));
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    ctx->ob_n_nodes = n_nodes; ctx->ob_n_refs = leaf_base; ctx->ob_valid = true;
    if (n_nodes_out) *n_nodes_out = n_nodes;
    if (n_refs_out) *n_refs_out = leaf_base;
    if (build_ms) *build_ms = ms;
    return GI_OK;
}
extern "C" int gi_octree_download(gi_ctx* ctx, double* node_box, uint32_t* node_child, uint8_t* node_mask, uint32_t* node_prim_off, uint32_t* node_prim_cnt, uint32_t* leaf_prims)
{
    if (!ctx) return GI_ERR_INVALID;
    if (!ctx->ob_valid) return fail(ctx, GI_ERR_INVALID, "gi_octree_download needs gi_octree_build");
    CK(cudaSetDevice(ctx->device));
    const size_t n = ctx->ob_n_nodes, r = ctx->ob_n_refs;
    if (node_box) CK(cudaMemcpyAsync(node_box, ctx->ob_nodebox.p, n * 48, cudaMemcpyDeviceToHost, ctx->stream));
    if (node_child) CK(cudaMemcpyAsync(node_child, ctx->ob_child.p, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (node_mask) CK(cudaMemcpyAsync(node_mask, ctx->ob_mask.p, n, cudaMemcpyDeviceToHost, ctx->stream));
    if (node_prim_off) CK(cudaMemcpyAsync(node_prim_off, ctx->ob_poff.p, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (node_prim_cnt) CK(cudaMemcpyAsync(node_prim_cnt, ctx->ob_pcnt.p, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (leaf_prims && r) CK(cudaMemcpyAsync(leaf_prims, ctx->ob_leaf.p, r * 4, cudaMemcpyDeviceToHost, ctx->stream));
    return gi_synchronize(ctx);
}

// ---- photons ------------------------------------------------------------------------------------------------------------------------------
extern "C" int gi_photon_upload(gi_ctx* ctx, size_t n, const double* photons9)
{
    if (!ctx || (n && !photons9)) return GI_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    CK(ctx->b_photons.reserve(std::max<size_t>(n, 1) * 72));
    if (n) CK(cudaMemcpyAsync(ctx->b_photons.p, photons9, n * 72, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->n_photons = n;
    ctx->has_map = false;
    return GI_OK;
}
extern "C" int gi_photon_count(gi_ctx* ctx, size_t* n)
{
    if (!ctx || !n) return GI_ERR_INVALID;
    *n = ctx->n_photons;
    return GI_OK;
}
extern "C" int gi_photon_download(gi_ctx* ctx, size_t n, double* photons9)
{
    if (!ctx || (n && !photons9) || n > ctx->n_photons) return GI_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    if (n) CK(cudaMemcpyAsync(photons9, ctx->b_photons.p, n * 72, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return GI_OK;
}

extern "C" int gi_photon_trace(gi_ctx* ctx, int count, int max_depth, uint64_t seed, uint64_t* n_stored, gi_stats* stats)
{
    if (!ctx || count < 0 || max_depth < 0) return GI_ERR_INVALID;
    if (!ctx->has_scene) return fail(ctx, GI_ERR_NO_SCENE, "gi_photon_trace needs gi_scene_upload");
    CK(cudaSetDevice(ctx->device));
    size_t slots = (size_t)count * ctx->S.n_lights;
    ctx->has_map = false;
    ctx->n_photons = 0;
    if (stats) std::memset(stats, 0, sizeof(*stats));
    if (n_stored) *n_stored = 0;
    if (!slots) return GI_OK;
    if (slots > 0xFFFFFFF0ull) return fail(ctx, GI_ERR_INVALID, "too many photons");
    CK(ctx->w0.reserve(slots * 72)); CK(ctx->w1.reserve(slots)); CK(ctx->w2.reserve(slots * 4)); CK(ctx->b_scan0.reserve(slots * 4)); CK(ctx->b_misc.reserve(64));
    CK(ctx->b_photons.reserve(slots * 72));
    CK(cudaMemsetAsync(ctx->w1.p, 0, slots, ctx->stream));
    CK(cudaMemsetAsync(ctx->b_misc.p, 0, 64, ctx->stream));
    DPhotonOut O{ ctx->w0.as<double>(), ctx->w1.as<uint8_t>(), ctx->b_misc.as<unsigned long long>(), ctx->b_misc.as<unsigned long long>() + 1, work_ptr(ctx, 0) };
    CK(cudaMemsetAsync(work_ptr(ctx, 0), 0, 16, ctx->stream));
    fam_reset(ctx, "photon_trace");
    EventPair ev(ctx); const cudaEvent_t e0 = ev.a, e1 = ev.b;
    cudaEventRecord(e0, ctx->stream);
    uint64_t launches = 4;
    {
        // rounds of `width` speculative tries per active slot (k_photon_round); width grows as the survivor list shrinks
        ScopedTimer t(ctx, "photon_trace");
        uint32_t* lists[2] = { ctx->w2.as<uint32_t>(), ctx->b_scan0.as<uint32_t>() };   // survivor lists (both buffers are reused by the compaction below)
        uint32_t* n_out = reinterpret_cast<uint32_t*>(ctx->b_misc.as<unsigned long long>() + 4);
        const uint32_t* in = nullptr;
        uint32_t n_active = (uint32_t)slots;
        int base_try = 0;
        for (int round = 0; base_try < 500 && n_active > 0; round++) {
            GI_POLL_CANCEL("photon trace");
            int width = 1;
            while (width < 32 && (uint64_t)n_active * (uint64_t)(2 * width) <= (1u << 19)) width *= 2;
            CK(cudaMemsetAsync(n_out, 0, 4, ctx->stream));
            uint32_t* out = lists[round & 1];
            GI_LAUNCH_M(k_photon_round, grid_for((size_t)n_active * width, GI_BLOCK), GI_BLOCK, ctx->S, count, max_depth, seed, base_try, width, n_active, in, out, n_out, O);
            launches++;
            CK(cudaGetLastError());
            const uint32_t was = n_active;
            CK(cudaMemcpyAsync(&n_active, n_out, 4, cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            if (getenv("GI_TRACE_LAUNCHES")) fprintf(stderr, "[gi] photon round %3d tries %3d..%3d active %9u -> %9u\n", round, base_try, base_try + width - 1, was, n_active);
            in = out;
            base_try += width;
        }
    }
    CK(cudaGetLastError());
    // canonical (i, light) order: flags -> exclusive scan -> scatter
    k_flags_to_u32<<<grid_for(slots, 256), 256, 0, ctx->stream>>>(ctx->w1.as<uint8_t>(), (uint32_t)slots, ctx->w2.as<uint32_t>());
    uint32_t* total_dev = reinterpret_cast<uint32_t*>(ctx->b_misc.as<unsigned long long>() + 2);
    int rc = scan_exclusive(ctx, ctx->w2.as<uint32_t>(), 1, (uint32_t)slots, ctx->b_scan0.as<uint32_t>(), total_dev);
    if (rc != GI_OK) return rc;
    k_photon_compact<<<grid_for(slots, 256), 256, 0, ctx->stream>>>(ctx->w0.as<double>(), ctx->w1.as<uint8_t>(), ctx->b_scan0.as<uint32_t>(), (uint32_t)slots, ctx->b_photons.as<double>());
    CK(cudaGetLastError());
    cudaEventRecord(e1, ctx->stream);
    unsigned long long host[3];
    CK(cudaMemcpyAsync(host, ctx->b_misc.p, 24, cudaMemcpyDeviceToHost, ctx->stream));
    rc = gi_synchronize(ctx);
    if (rc != GI_OK) return rc;
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    ctx->n_photons = (uint32_t)host[2];
    if (n_stored) *n_stored = ctx->n_photons;
    ctx->work_host[8] = host[1];
    if (stats) {
        stats->photon_tries = host[0]; stats->closest_rays = host[1]; stats->photons_stored = ctx->n_photons; stats->kernel_launches = launches; stats->total_ms = ms;
        stats->trace_ms = ctx->fam["photon_trace"].ms; stats->closest_node_tests = ctx->work_host[0]; stats->closest_prim_tests = ctx->work_host[1];
    }
    return GI_OK;
}

// ---- photon map ---------------------------------------------------------------------------------------------------------------------------
struct SlabHeader { uint32_t magic, n_nodes, n_kept, n_leaves, max_depth, n_cand, pad[2]; uint64_t off_nodes, off_pos, off_dircol, off_pid, off_cand_off, off_cand_rec, total; };
#define GI_SLAB_MAGIC 0x47495035u   // "GIP5"
static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

static void bind_slab(gi_ctx* ctx, const SlabHeader& h)
{
    char* base = ctx->b_slab.as<char>();
    ctx->G.nodes = reinterpret_cast<const DNode*>(base + h.off_nodes);
    ctx->G.pos4 = reinterpret_cast<const double*>(base + h.off_pos);
    ctx->G.dircol = reinterpret_cast<const double*>(base + h.off_dircol);
    ctx->G.pid = reinterpret_cast<const uint32_t*>(base + h.off_pid);
    ctx->G.cand_off = reinterpret_cast<const uint32_t*>(base + h.off_cand_off);
    ctx->G.cand_rec = reinterpret_cast<const double*>(base + h.off_cand_rec);
    ctx->G.n_nodes = h.n_nodes;
    ctx->pm_nodes = h.n_nodes; ctx->pm_kept = h.n_kept; ctx->pm_leaves = h.n_leaves; ctx->pm_depth = h.max_depth;
    ctx->slab_bytes = h.total;
    ctx->has_map = true;
}

extern "C" int gi_photon_map_build(gi_ctx* ctx, const double* box6)
{
    if (!ctx) return GI_ERR_INVALID;
    if (!box6 && !ctx->has_scene) return fail(ctx, GI_ERR_NO_SCENE, "gi_photon_map_build needs a root box (scene or box6)");
    CK(cudaSetDevice(ctx->device));
    ctx->has_map = false;   // set again by bind_slab at the very end: a failure below must not leave ctx->G pointing at a freed slab
    const uint32_t n = (uint32_t)ctx->n_photons;
    double box[6];
    for (int k = 0; k < 6; k++) box[k] = box6 ? box6[k] : ctx->root_box[k];
    const uint32_t cap_nodes = std::max<uint32_t>(4u * n + 1024u, 1024u);
    // build workspace
    CK(ctx->w0.reserve((size_t)cap_nodes * sizeof(DNode)));    // nodes
    CK(ctx->w1.reserve(std::max<size_t>(n, 1) * 4));           // pnode
    CK(ctx->w2.reserve(std::max<size_t>(n, 1) * 4));           // pid (leaf order)
    CK(ctx->w3.reserve(64));                                   // counters: n_nodes, n_kept, overflow
    CK(ctx->w4.reserve(48));                                   // box
    CK(ctx->b_photons.reserve(72));
    CK(cudaMemcpyAsync(ctx->w4.p, box, 48, cudaMemcpyHostToDevice, ctx->stream));
    DPMap M{};
    M.nodes = ctx->w0.as<DNode>(); M.n_nodes = ctx->w3.as<uint32_t>(); M.cap_nodes = cap_nodes; M.ph = ctx->b_photons.as<double>(); M.n_photons = n;
    M.pnode = ctx->w1.as<uint32_t>(); M.pid = ctx->w2.as<uint32_t>(); M.n_kept = ctx->w3.as<uint32_t>() + 1; M.overflow = ctx->w3.as<uint32_t>() + 2;
    fam_reset(ctx, "pm_build");
    EventPair ev(ctx); const cudaEvent_t e0 = ev.a, e1 = ev.b;
    cudaEventRecord(e0, ctx->stream);
    k_pm_init<<<grid_for(std::max<uint32_t>(n, 1), 256), 256, 0, ctx->stream>>>(M, ctx->w4.as<double>());
    CK(cudaGetLastError());
    uint32_t lb = 0, le = 1, depth = 0;
    uint32_t host_cnt[3] = { 1, 0, 0 };
    while (le > lb && depth < 128) {
        k_pm_split<<<grid_for(le - lb, 128), 128, 0, ctx->stream>>>(M, lb, le);
        if (n) k_pm_assign<<<grid_for(n, 256), 256, 0, ctx->stream>>>(M, lb, le);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(host_cnt, ctx->w3.p, 12, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        if (host_cnt[2]) return fail(ctx, GI_ERR_OOM, "photon map node pool exhausted");
        lb = le; le = host_cnt[0];
        if (le > lb) depth++;
    }
    const uint32_t n_nodes = host_cnt[0];
    // leaf offsets = exclusive scan of the per-node counts (interior nodes hold 0)
    CK(ctx->b_scan0.reserve((size_t)n_nodes * 4)); CK(ctx->w5.reserve((size_t)n_nodes * 4));
    int rc = scan_exclusive(ctx, &M.nodes[0].prim_cnt, (int)(sizeof(DNode) / 4), n_nodes, ctx->b_scan0.as<uint32_t>(), M.n_kept);
    if (rc != GI_OK) return rc;
    k_pm_set_offsets<<<grid_for(n_nodes, 256), 256, 0, ctx->stream>>>(M, n_nodes, ctx->b_scan0.as<uint32_t>(), ctx->w5.as<uint32_t>());
    if (n) k_pm_scatter<<<grid_for(n, 256), 256, 0, ctx->stream>>>(M, ctx->w5.as<uint32_t>());
    k_pm_order_leaf<<<grid_for(n_nodes, 128), 128, 0, ctx->stream>>>(M, n_nodes);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(host_cnt, ctx->w3.p, 12, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    const uint32_t n_kept = host_cnt[1];
    // per-leaf candidate lists (Node::get run once per leaf): count -> scan -> fill
    CK(ctx->w6.reserve((size_t)(n_nodes + 1) * 4)); CK(ctx->w7.reserve((size_t)(n_nodes + 1) * 4));
    uint32_t* cand_cnt = ctx->w6.as<uint32_t>();
    uint32_t* cand_off = ctx->w7.as<uint32_t>();
    k_pm_cands<4, false><<<grid_for(n_nodes, 4), 128, 0, ctx->stream>>>(M.nodes, n_nodes, cand_cnt, nullptr, nullptr, M.overflow);
    CK(cudaGetLastError());
    rc = scan_exclusive(ctx, cand_cnt, 1, n_nodes, cand_off, cand_off + n_nodes);
    if (rc != GI_OK) return rc;
    CK(cudaMemcpyAsync(host_cnt, cand_off + n_nodes, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(host_cnt + 2, M.overflow, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (host_cnt[2]) return fail(ctx, GI_ERR_OOM, "photon map candidate traversal stack overflow");
    const uint32_t n_cand = host_cnt[0];
    // compact slab: header | nodes | pos4 | dircol | pid | cand_off | cand_rec
    SlabHeader h{};
    h.magic = GI_SLAB_MAGIC; h.n_nodes = n_nodes; h.n_kept = n_kept; h.max_depth = depth; h.n_cand = n_cand;
    h.off_nodes = 256; h.off_pos = align256(h.off_nodes + (size_t)n_nodes * sizeof(DNode)); h.off_dircol = align256(h.off_pos + (size_t)n_kept * 32);
    h.off_pid = align256(h.off_dircol + (size_t)n_kept * 48); h.off_cand_off = align256(h.off_pid + (size_t)n_kept * 4);
    h.off_cand_rec = align256(h.off_cand_off + (size_t)(n_nodes + 1) * 4); h.total = align256(h.off_cand_rec + (size_t)n_cand * 32);
    CK(ctx->w8.reserve(std::max<size_t>(n_cand, 1) * 4));   // candidate slots in DFS order (scratch; the slab holds the ordered records)
    CK(ctx->b_slab.reserve(h.total));
    char* base = ctx->b_slab.as<char>();
    CK(cudaMemcpyAsync(base + h.off_nodes, M.nodes, (size_t)n_nodes * sizeof(DNode), cudaMemcpyDeviceToDevice, ctx->stream));
    if (n_kept) CK(cudaMemcpyAsync(base + h.off_pid, M.pid, (size_t)n_kept * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    CK(cudaMemcpyAsync(base + h.off_cand_off, cand_off, (size_t)(n_nodes + 1) * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    k_pm_cands<4, true><<<grid_for(n_nodes, 4), 128, 0, ctx->stream>>>(M.nodes, n_nodes, nullptr, cand_off, ctx->w8.as<uint32_t>(), M.overflow);
    M.pos4 = reinterpret_cast<double*>(base + h.off_pos); M.dircol = reinterpret_cast<double*>(base + h.off_dircol);
    if (n_kept) k_pm_payload<<<grid_for(n_kept, 256), 256, 0, ctx->stream>>>(M, n_kept);
    CK(cudaGetLastError());
    // candidate records.  Lists the gather scans per thread (<= GI_GS_MAX_CANDS entries) are ordered by distance from their
    // leaf centre (early stop); longer lists are streamed by whole warps and keep the DFS order: along a centre-ordered list
    // of a large, nearly empty leaf the distances to an off-centre query keep FALLING, so every candidate would beat the
    // running k-th and force a merge, while in DFS order only ~k ln(C/k) do
    if (n_cand) {
        const uint32_t SHORT_LEN = GI_GS_MAX_CANDS, LONG_LEN = GI_GS_MAX_CANDS;
        if (getenv("GI_TRACE_LAUNCHES")) {
            std::vector<uint32_t> hc(n_nodes);
            cudaMemcpy(hc.data(), cand_cnt, (size_t)n_nodes * 4, cudaMemcpyDeviceToHost);
            uint64_t n512 = 0, n8k = 0, n64k = 0, mx = 0, s8k = 0;
            for (uint32_t v : hc) { n512 += v > 512; n8k += v > 8192; n64k += v > 65536; mx = std::max<uint64_t>(mx, v); if (v > 8192) s8k += v; }
            fprintf(stderr, "[gi] candidate lists: %u total entries, max %llu, >512: %llu, >8192: %llu (sum %llu), >65536: %llu\n", n_cand, (unsigned long long)mx, (unsigned long long)n512,
                    (unsigned long long)n8k, (unsigned long long)s8k, (unsigned long long)n64k);
        }
        const uint32_t* cslot = ctx->w8.as<uint32_t>();
        double* crec = reinterpret_cast<double*>(base + h.off_cand_rec);
        k_pm_cand_order<<<n_nodes, 64, SHORT_LEN * 12, ctx->stream>>>(M.nodes, n_nodes, M.pos4, cand_off, cslot, crec, 0u, SHORT_LEN, LONG_LEN);
        k_pm_cand_order<<<n_nodes, 256, 0, ctx->stream>>>(M.nodes, n_nodes, M.pos4, cand_off, cslot, crec, SHORT_LEN, 0xFFFFFFFFu, LONG_LEN);
        CK(cudaGetLastError());
    }
    // leaves are counted on the host from the node records (also validates the build)
    std::vector<DNode> hn(n_nodes);
    CK(cudaMemcpyAsync(hn.data(), M.nodes, (size_t)n_nodes * sizeof(DNode), cudaMemcpyDeviceToHost, ctx->stream));
    cudaEventRecord(e1, ctx->stream);
    CK(cudaStreamSynchronize(ctx->stream));
    uint32_t leaves = 0;
    for (auto& nd : hn) if (nd.mask == 0) leaves++;
    h.n_leaves = leaves;
    CK(cudaMemcpyAsync(base, &h, sizeof(h), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    ctx->fam["pm_build"].ms = ms; ctx->fam["pm_build"].launches = 1;
    bind_slab(ctx, h);
    return GI_OK;
}

extern "C" int gi_photon_map_info(gi_ctx* ctx, uint32_t* n_nodes, uint32_t* n_leaves, uint32_t* n_kept, uint32_t* max_depth)
{
    if (!ctx) return GI_ERR_INVALID;
    if (!ctx->has_map) return fail(ctx, GI_ERR_NO_PHOTONS, "no photon map");
    if (n_nodes) *n_nodes = ctx->pm_nodes;
    if (n_leaves) *n_leaves = ctx->pm_leaves;
    if (n_kept) *n_kept = ctx->pm_kept;
    if (max_depth) *max_depth = ctx->pm_depth;
    return GI_OK;
}

extern "C" int gi_photon_map_download(gi_ctx* ctx, double* node_box6, uint8_t* node_is_leaf, uint32_t* node_count, uint32_t* photon_ids)
{
    if (!ctx) return GI_ERR_INVALID;
    if (!ctx->has_map) return fail(ctx, GI_ERR_NO_PHOTONS, "no photon map");
    CK(cudaSetDevice(ctx->device));
    std::vector<DNode> hn(ctx->pm_nodes);
    std::vector<uint32_t> pid(std::max<uint32_t>(ctx->pm_kept, 1));
    CK(cudaMemcpy(hn.data(), ctx->G.nodes, (size_t)ctx->pm_nodes * sizeof(DNode), cudaMemcpyDeviceToHost));
    if (ctx->pm_kept) CK(cudaMemcpy(pid.data(), ctx->G.pid, (size_t)ctx->pm_kept * 4, cudaMemcpyDeviceToHost));
    // DFS pre-order, children 0..7, like the reference's recursion
    std::vector<uint32_t> st = { 0 };
    size_t ni = 0, pi = 0;
    while (!st.empty()) {
        uint32_t i = st.back(); st.pop_back();
        const DNode& nd = hn[i];
        if (node_box6) for (int k = 0; k < 3; k++) { node_box6[6 * ni + k] = nd.bmin[k]; node_box6[6 * ni + 3 + k] = nd.bmax[k]; }
        if (node_is_leaf) node_is_leaf[ni] = nd.mask == 0;
        if (node_count) node_count[ni] = nd.prim_cnt;
        ni++;
        if (nd.mask == 0) { if (photon_ids) for (uint32_t k = 0; k < nd.prim_cnt; k++) photon_ids[pi + k] = pid[nd.prim_off + k]; pi += nd.prim_cnt; }
        else for (int c = 7; c >= 0; c--) st.push_back(nd.child + c);
    }
    return GI_OK;
}

extern "C" int gi_photon_map_slab_size(gi_ctx* ctx, size_t* bytes)
{
    if (!ctx || !bytes) return GI_ERR_INVALID;
    if (!ctx->has_map) return fail(ctx, GI_ERR_NO_PHOTONS, "no photon map");
    *bytes = ctx->slab_bytes;
    return GI_OK;
}
extern "C" int gi_photon_map_slab_ptr(gi_ctx* ctx, void** dev_ptr)
{
    if (!ctx || !dev_ptr) return GI_ERR_INVALID;
    if (!ctx->has_map) return fail(ctx, GI_ERR_NO_PHOTONS, "no photon map");
    *dev_ptr = ctx->b_slab.p;
    return GI_OK;
}
extern "C" int gi_photon_map_reserve_slab(gi_ctx* ctx, size_t bytes, void** dev_ptr)
{
    if (!ctx || !dev_ptr || bytes < 256) return GI_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    ctx->has_map = false;
    CK(ctx->b_slab.reserve(bytes));
    *dev_ptr = ctx->b_slab.p;
    return GI_OK;
}
extern "C" int gi_photon_map_adopt_slab(gi_ctx* ctx, size_t bytes)
{
    if (!ctx || bytes < 256 || bytes > ctx->b_slab.cap) return GI_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    SlabHeader h;
    CK(cudaMemcpy(&h, ctx->b_slab.p, sizeof(h), cudaMemcpyDeviceToHost));
    ctx->has_map = false;
    // every section must lie inside the slab, in the builder's order, 256-byte aligned (a truncated or mismatched broadcast must
    // not become out-of-bounds reads in the gather)
    auto sect = [&](uint64_t off, uint64_t len, uint64_t next) { return (off & 255u) == 0 && off >= 256 && off + len <= next && next <= bytes; };
    const bool ok = h.magic == GI_SLAB_MAGIC && h.total == bytes && h.n_nodes >= 1 && h.n_leaves <= h.n_nodes
                    && sect(h.off_nodes, (uint64_t)h.n_nodes * sizeof(DNode), h.off_pos) && sect(h.off_pos, (uint64_t)h.n_kept * 32, h.off_dircol)
                    && sect(h.off_dircol, (uint64_t)h.n_kept * 48, h.off_pid) && sect(h.off_pid, (uint64_t)h.n_kept * 4, h.off_cand_off)
                    && sect(h.off_cand_off, ((uint64_t)h.n_nodes + 1) * 4, h.off_cand_rec) && sect(h.off_cand_rec, (uint64_t)h.n_cand * 32, bytes);
    if (!ok) return fail(ctx, GI_ERR_INVALID, "bad photon map slab");
    bind_slab(ctx, h);
    return GI_OK;
}

// ---- gather ---------------------------------------------------------------------------------------------------------------------------------
// the gather pipeline on device pointers: locate -> order by leaf (counting sort) -> one thread per query
#define GI_GATHER_SORT_MIN 2048u   // shorter queues skip the ordering
static int run_gather(gi_ctx* ctx, uint32_t n, const double* pos, const double* dir, int k, double* rgb, uint32_t* knn, uint32_t* n_cand, const double* weight, double* accum,
                      const uint32_t* accum_idx, uint64_t* launches)
{
    const uint32_t nk = ctx->G.n_nodes + 1;   // keys: leaf node id, n_nodes = "in no leaf"
    const bool order = n >= GI_GATHER_SORT_MIN;
    CK(ctx->b_gnode.reserve((size_t)n * 4)); CK(ctx->b_gheavy.reserve((size_t)n * 4 + 16));
    uint32_t* heavy_cnt = ctx->b_gheavy.as<uint32_t>();      // [0] queued, [1] next; the queue starts at [4]
    CK(cudaMemsetAsync(heavy_cnt, 0, 16, ctx->stream));
    if (order) { CK(ctx->b_gperm.reserve((size_t)n * 4)); CK(ctx->b_ghist.reserve((size_t)nk * 4)); CK(ctx->b_gcur.reserve((size_t)nk * 4)); CK(cudaMemsetAsync(ctx->b_ghist.p, 0, (size_t)nk * 4, ctx->stream)); }
    k_gather_locate<<<grid_for(n, 256), 256, 0, ctx->stream>>>(ctx->G, n, pos, ctx->b_gnode.as<uint32_t>(), order ? ctx->b_ghist.as<uint32_t>() : nullptr, work_ptr(ctx, 4), heavy_cnt + 4, heavy_cnt);
    // the long lists (queued by the locate kernel) go to persistent warps on the auxiliary stream, beside the counting sort and the
    // thread-per-query kernel; the calling stream takes them back in before anything that follows (the next depth's gather adds to the same Lc sums)
    // (also when overlap_threshold = 0 keeps the FAMILIES of a frame on one stream: the auxiliary stream is the gather pipeline's own)
    const bool aux = ctx->gather_aux != nullptr;
    const cudaStream_t hs = aux ? ctx->gather_aux : ctx->stream;
    auto launch_heavy = [&]() {
        k_gather_heavy<<<ctx->n_sm * 4, GI_WPB * 32, 0, hs>>>(ctx->G, heavy_cnt + 4, heavy_cnt, heavy_cnt + 1, ctx->b_gnode.as<uint32_t>(), pos, dir, k, rgb, knn, n_cand, weight, accum, accum_idx);
    };
    if (aux) {
        cudaEventRecord(ctx->gather_ev[0], ctx->stream); cudaStreamWaitEvent(hs, ctx->gather_ev[0], 0);
        launch_heavy();
        cudaEventRecord(ctx->gather_ev[1], hs);
    }
    if (order) {
        int rc = scan_exclusive(ctx, ctx->b_ghist.as<uint32_t>(), 1, nk, ctx->b_gcur.as<uint32_t>(), nullptr);
        if (rc != GI_OK) return rc;
        k_bin_scatter<<<grid_for(n, 256), 256, 0, ctx->stream>>>(n, ctx->b_gnode.as<uint32_t>(), ctx->b_gcur.as<uint32_t>(), ctx->b_gperm.as<uint32_t>());
    }
    k_gather_sorted<<<grid_for(n, GI_GS_BLOCK), GI_GS_BLOCK, 0, ctx->stream>>>(ctx->G, n, order ? ctx->b_gperm.as<uint32_t>() : nullptr, ctx->b_gnode.as<uint32_t>(), pos, dir, k, rgb, knn, n_cand,
                                                                              weight, accum, accum_idx, work_ptr(ctx, 4));
    if (aux) cudaStreamWaitEvent(ctx->stream, ctx->gather_ev[1], 0);
    else launch_heavy();
    CK(cudaGetLastError());
    if (launches) *launches += order ? 7 : 3;
    return GI_OK;
}

extern "C" int gi_photon_gather_dev(gi_ctx* ctx, size_t n, const double* pos, const double* dir, int k, double* rgb, uint32_t* knn, uint32_t* n_cand)
{
    if (!ctx || (n && (!pos || !dir)) || k < 1 || k > 32) return GI_ERR_INVALID;
    if (!ctx->has_map) return fail(ctx, GI_ERR_NO_PHOTONS, "gi_photon_gather needs gi_photon_map_build");
    if (!n) return GI_OK;
    CK(cudaSetDevice(ctx->device));
    fam_reset(ctx, "gather");
    CK(cudaMemsetAsync(work_ptr(ctx, 4), 0, 24, ctx->stream));
    ctx->work_host[10] = n;
    {
        ScopedTimer t(ctx, "gather");
        if (n > 0xFFFFFFF0ull) return fail(ctx, GI_ERR_INVALID, "too many queries in one call");
        int rc = run_gather(ctx, (uint32_t)n, pos, dir, k, rgb, knn, n_cand, nullptr, nullptr, nullptr, nullptr);
        if (rc != GI_OK) return rc;
    }
    CK(cudaGetLastError());
    return GI_OK;
}
extern "C" int gi_photon_gather(gi_ctx* ctx, size_t n, const double* pos, const double* dir, int k, double* rgb, uint32_t* knn, uint32_t* n_cand)
{
    if (!ctx || (n && (!pos || !dir)) || k < 1 || k > 32) return GI_ERR_INVALID;
    if (!ctx->has_map) return fail(ctx, GI_ERR_NO_PHOTONS, "gi_photon_gather needs gi_photon_map_build");
    if (!n) return GI_OK;
    CK(cudaSetDevice(ctx->device));
    CK(ctx->w0.reserve(n * 24)); CK(ctx->w1.reserve(n * 24)); CK(ctx->w2.reserve(n * 24)); CK(ctx->w3.reserve(n * 4 * (size_t)k)); CK(ctx->w4.reserve(n * 4));
    CK(cudaMemcpyAsync(ctx->w0.p, pos, n * 24, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->w1.p, dir, n * 24, cudaMemcpyHostToDevice, ctx->stream));
    int rc = gi_photon_gather_dev(ctx, n, ctx->w0.as<double>(), ctx->w1.as<double>(), k, ctx->w2.as<double>(), knn ? ctx->w3.as<uint32_t>() : nullptr, n_cand ? ctx->w4.as<uint32_t>() : nullptr);
    if (rc != GI_OK) return rc;
    if (rgb) CK(cudaMemcpyAsync(rgb, ctx->w2.p, n * 24, cudaMemcpyDeviceToHost, ctx->stream));
    if (knn) CK(cudaMemcpyAsync(knn, ctx->w3.p, n * 4 * (size_t)k, cudaMemcpyDeviceToHost, ctx->stream));
    if (n_cand) CK(cudaMemcpyAsync(n_cand, ctx->w4.p, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    return gi_synchronize(ctx);
}

// ---- frame ------------------------------------------------------------------------------------------------------------------------------------
// Fixed sample range [s0, s1) into accum_dev (sums), or — with `adapt` — the reference's adaptive per-pixel loop: accum_dev then
// receives the final running-mean colour of each pixel and adapt->s_done the samples taken.
// rows of part `part` of `nparts` under the tile split's plan: blocks part, part + nparts, .. of block_rows rows
static inline int rows_of_part(int height, int block_rows, int nparts, int part)
{
    int n = 0;
    for (int b = part, y0; (y0 = b * block_rows) < height; b += nparts) n += std::min(block_rows, height - y0);
    return n;
}
struct RowPlan { int block_rows, nparts, part; };
static int render_device(gi_ctx* ctx, const gi_render_params* P, int x0, int y0, int x1, int y1, int s0, int s1, double* accum_dev, gi_stats* stats, DAdapt* adapt = nullptr,
                         const RowPlan* rows = nullptr)
{
    if (rows) { x0 = 0; x1 = P->width; y0 = rows->part * rows->block_rows; y1 = y0 + rows_of_part(P->height, rows->block_rows, rows->nparts, rows->part); }   // y1 - y0 = local rows
    const size_t npx = (size_t)(x1 - x0) * (y1 - y0);
    ctx->stream = ctx->main_stream;
    const uint64_t total_paths = adapt ? (uint64_t)npx : (uint64_t)npx * (uint64_t)(s1 - s0);
    const uint32_t chunk_cap = (uint32_t)std::min<uint64_t>(total_paths, GI_MAX_PATHS);
    // queues (double buffered), hit list, per-path state
    for (int b = 0; b < 4; b++) { CK(ctx->q_a[b].reserve((size_t)chunk_cap * 24)); CK(ctx->q_b[b].reserve((size_t)chunk_cap * 24)); }
    CK(ctx->q_a[4].reserve((size_t)chunk_cap * 4)); CK(ctx->q_b[4].reserve((size_t)chunk_cap * 4));
    for (int b = 0; b < 5; b++) CK(ctx->hl[b].reserve((size_t)chunk_cap * 24));
    CK(ctx->hl[5].reserve((size_t)chunk_cap * 8)); CK(ctx->hl[6].reserve((size_t)chunk_cap * 4));
    const bool overlap = ctx->overlap_threshold > 0;
    const bool deferred = overlap && (ctx->sched_mode == 1 || (ctx->sched_mode == 2 && ctx->sched_hint != 0));   // see gi_ctx::sched_mode
    bool had_short = false;   // this frame ran a tail or a depth of fewer than 2^20 hits
    for (int k = 0; k < 2; k++) ctx->side[k] = ctx->side_pairs[deferred ? 1 : 0][k];   // (both pairs are idle between calls: every call drains them)
    const int nhl = deferred ? std::min(std::max(ctx->ring, 2), GI_NHL) : (overlap ? 2 : 1);   // hit lists in use: depths take them in turn
    DevBuf* hls[GI_NHL] = { ctx->hl, ctx->hl2 };
    for (int r = 2; r < GI_NHL; r++) hls[r] = ctx->hlr[r - 2];
    for (int r = 1; r < nhl; r++) {
        for (int b = 0; b < 5; b++) CK(hls[r][b].reserve((size_t)chunk_cap * 24));
        CK(hls[r][5].reserve((size_t)chunk_cap * 8)); CK(hls[r][6].reserve((size_t)chunk_cap * 4));
    }
    CK(ctx->ps[0].reserve((size_t)chunk_cap * 4)); CK(ctx->ps[1].reserve((size_t)chunk_cap * 8)); CK(ctx->ps[2].reserve((size_t)chunk_cap * 24)); CK(ctx->ps[3].reserve((size_t)chunk_cap * 24)); CK(ctx->ps[4].reserve((size_t)chunk_cap * 24));
    CK(ctx->b_cnt.reserve(sizeof(DCounters))); CK(ctx->b_misc.reserve(64));
    CK(ctx->b_tail.reserve(sizeof(DTailCounters)));
    CK(cudaMemsetAsync(ctx->b_tail.p, 0, sizeof(DTailCounters), ctx->stream));
    const bool binning = ctx->bin_threshold > 0 && chunk_cap >= ctx->bin_threshold;
    if (binning) { CK(ctx->b_binkey.reserve((size_t)chunk_cap * 4)); CK(ctx->b_binperm.reserve((size_t)chunk_cap * 4)); CK(ctx->b_binhist.reserve((size_t)GI_SORT_BINS * 4)); CK(ctx->b_bincur.reserve((size_t)GI_SORT_BINS * 4)); }
    d3 bin_min, bin_inv;
    bin_min.x = ctx->root_box[0]; bin_min.y = ctx->root_box[1]; bin_min.z = ctx->root_box[2];
    bin_inv.x = (double)(1 << GI_BIN_AXIS_BITS) / std::max(ctx->root_box[3] - ctx->root_box[0], 1e-300); bin_inv.y = (double)(1 << GI_BIN_AXIS_BITS) / std::max(ctx->root_box[4] - ctx->root_box[1], 1e-300); bin_inv.z = (double)(1 << GI_BIN_AXIS_BITS) / std::max(ctx->root_box[5] - ctx->root_box[2], 1e-300);
    DQueue qa{ ctx->q_a[0].as<double>(), ctx->q_a[1].as<double>(), ctx->q_a[2].as<double>(), ctx->q_a[3].as<double>(), ctx->q_a[4].as<uint32_t>() };
    DQueue qb{ ctx->q_b[0].as<double>(), ctx->q_b[1].as<double>(), ctx->q_b[2].as<double>(), ctx->q_b[3].as<double>(), ctx->q_b[4].as<uint32_t>() };
    DHitList Hs[GI_NHL];
    for (int r = 0; r < GI_NHL; r++) Hs[r] = DHitList{ hls[r][0].as<double>(), hls[r][1].as<double>(), hls[r][2].as<double>(), hls[r][3].as<double>(), hls[r][4].as<double>(), hls[r][5].as<double>(), hls[r][6].as<uint32_t>() };
    DPathState PS{ ctx->ps[0].as<uint32_t>(), ctx->ps[1].as<uint64_t>(), ctx->ps[2].as<double>(), ctx->ps[3].as<double>(), ctx->ps[4].as<double>() };
    DCounters* C = ctx->b_cnt.as<DCounters>();
    DFrame F = make_frame(ctx, P->width, P->height, x0, y0, x1, y1);
    if (rows) { F.rb = rows->block_rows; F.rstride = rows->block_rows * rows->nparts; }
    const bool have_map = ctx->has_map && ctx->pm_kept > 0;
    // long, uneven walks (deep trees): persistent warps with ray refetch; short ones: one ray per thread, launch per queue.
    // Decided — deterministically — from the node tests per closest-hit ray that the previous large frame of this scene tallied
    // (first frame: from the tree size).  Measured: refetch wins at 139 (sponza stand-in, 4483 -> 3092 ms) and loses ~8 % at 28
    // (caustics) and 60 (glass).  (A timing-based choice between the two forms flipped from run to run on scenes where they
    // are within a few per cent of each other, which made frame times bimodal.)
    const bool tunable = total_paths >= (1u << 20);   // small calls do not update the statistic
    // (round 2, 64-register kernels, before the pruned walk; frame ms classic -> refetch: atrium 264 -> 224, foliage 212 -> 227, glass 49.2 -> 52.4,
    //  caustics 24.6 -> 25.8.  With the pruned walk (rules R1-R3 of gi_device.cuh) the long tails inside a warp are what is left of the walks on
    //  the two stand-ins, and refetch wins on both: atrium 164 -> 153, foliage 133 -> 124; glass 157 -> 169 still loses (profiles/r02/ab_t14.txt).
    //  Node tests per ray after pruning: caustics 24, glass 45, foliage ~85, atrium 105.)
    const bool long_walks = ctx->nodes_per_ray > 0 ? ctx->nodes_per_ray > 64.0 : ctx->S.n_nodes > 200000u;
    const bool persistent = ctx->bounce_mode == 2 || (ctx->bounce_mode == 0 && long_walks);
    uint64_t n_closest = 0, n_shadow = 0, n_gather = 0, launches = 0;
    for (const char* f : { "bounce", "direct", "gather", "tail", "bin" }) fam_reset(ctx, f);
    CK(cudaMemsetAsync(work_ptr(ctx, 0), 0, 8 * sizeof(unsigned long long), ctx->stream));
    EventPair ev(ctx); const cudaEvent_t e0 = ev.a, e1 = ev.b;
    cudaEventRecord(e0, ctx->stream);
    // the bounce loop over one batch of n camera paths sitting in qa / PS
    auto run_depths = [&](uint32_t n) -> int {
        DQueue in = qa, out = qb;
        uint32_t n_active = n;
        const uint32_t* perm = nullptr;
        bool pending[GI_NHL] = {};   // side-stream work still reading hit list r of the ring
        auto wait_side = [&](int par) {       // the main stream waits for the side-stream work that reads hit list `par`
            if (!pending[par]) return;
            cudaStreamWaitEvent(ctx->main_stream, ctx->side_done[par][0], 0); cudaStreamWaitEvent(ctx->main_stream, ctx->side_done[par][1], 0);
            pending[par] = false;
        };
        auto wait_all_sides = [&]() { for (int r = 0; r < GI_NHL; r++) wait_side(r); };
        int hpar = 0;                         // hit list the next bounce writes
        for (int depth = 0; n_active > 0 && depth <= P->max_depth; depth++) {
            GI_POLL_CANCEL("render");
            // depths alternate between the two hit lists, so that a depth's shadow rays and gathers can run behind the next
            // depth's bounce kernel when they are few (decided below, once the hit count is known)
            hpar = (hpar + 1) % nhl;
            const DHitList& H = Hs[hpar];
            wait_side(hpar);
            if (depth > 0 && n_active < ctx->tail_threshold) {
                had_short = true;
                // few paths left: one warp per path runs them to the end inside one kernel; their gathers are queued and served
                // by one gather pipeline run afterwards (GI_TAIL_MODE=1: gathers inline)
                DTailQ Q{};
                const int last_gather_depth = std::min(P->max_depth, P->caustic_max_depth);
                Q.qmax = ctx->tail_mode == 0 && have_map && last_gather_depth >= depth ? (uint32_t)(last_gather_depth - depth + 1) : 0u;
                // The tail continues the paths' L / Ld / Lc sums, every earlier term must be in: L is the main stream's own; the tail's
                // Ld terms are added by k_tail_direct on side stream 0 behind every k_direct, its Lc terms by k_tail_caustic behind every
                // gather run — on side stream 1 in the deferred schedule (which then needs no wait here, unless the tail kernel makes
                // its gathers inline and so touches Lc itself: k_tail's `lc_rw`), on the main stream otherwise.
                const bool tail_lc_rw = have_map && Q.qmax == 0 && depth <= P->caustic_max_depth;
                if (!deferred || tail_lc_rw) wait_all_sides();
                const bool tail_sides = deferred && !tail_lc_rw;   // the tail's queued work goes behind the side streams' earlier work
                const int tp = tail_sides ? hpar : 0;              // the (free) ring slot whose events mark that work
                const size_t slots = (size_t)n_active * Q.qmax;
                if (Q.qmax) {
                    if (slots > 0xFFFFFFF0ull) return fail(ctx, GI_ERR_INVALID, "too many tail gather slots");
                    for (int b = 0; b < 4; b++) CK(ctx->tq[b].reserve(slots * 24));
                    CK(ctx->tq[4].reserve((size_t)n_active * 4));
                    Q.pos = ctx->tq[0].as<double>(); Q.dir = ctx->tq[1].as<double>(); Q.w = ctx->tq[2].as<double>(); Q.rgb = ctx->tq[3].as<double>(); Q.count = ctx->tq[4].as<uint32_t>();
                }
                if (ctx->S.n_lights) {
                    Q.smax = (uint32_t)(P->max_depth - depth + 1) * ctx->S.n_lights;
                    const size_t ss = (size_t)n_active * Q.smax;
                    if (ss > 0xFFFFFFF0ull) return fail(ctx, GI_ERR_INVALID, "too many tail shadow slots");
                    CK(ctx->tsh[0].reserve(ss * 24)); CK(ctx->tsh[1].reserve(ss * 24)); CK(ctx->tsh[2].reserve(ss * 8)); CK(ctx->tsh[3].reserve(ss * 24)); CK(ctx->tsh[4].reserve(ss * 4));
                    CK(ctx->tsh[5].reserve((size_t)n_active * 4)); CK(ctx->tsh[6].reserve(ss));
                    Q.sh_o = ctx->tsh[0].as<double>(); Q.sh_d = ctx->tsh[1].as<double>(); Q.sh_mt = ctx->tsh[2].as<double>(); Q.sh_w = ctx->tsh[3].as<double>();
                    Q.sh_depth = ctx->tsh[4].as<uint32_t>(); Q.sh_count = ctx->tsh[5].as<uint32_t>(); Q.sh_vis = ctx->tsh[6].as<uint8_t>();
                }
                {
                    ScopedTimer t(ctx, "tail");
                    CK(cudaMemsetAsync(&ctx->b_tail.as<DTailCounters>()->next, 0, 4, ctx->stream));
                    const unsigned tail_grid = std::min<unsigned>(grid_for(n_active, GI_WPB), (unsigned)ctx->n_sm * (unsigned)GI_TAIL_MINB);   // persistent warps: one resident wave
                    GI_LAUNCH_M(k_tail, tail_grid, GI_WPB * 32, ctx->S, ctx->G, have_map ? 1 : 0, *P, depth, n_active, in, PS, ctx->b_tail.as<DTailCounters>(), Q);
                    launches++;
                }
                if (overlap) cudaEventRecord(ctx->fork_ev, ctx->main_stream);   // k_tail has filled the queues
                if (Q.smax) {
                    // the queued shadow rays, one thread each, beside the tail's gather run (side stream 0); k_tail_direct then adds the
                    // unshadowed terms to Ld in bounce order
                    if (overlap) { ctx->stream = ctx->side[0]; cudaStreamWaitEvent(ctx->side[0], ctx->fork_ev, 0); }
                    {
                        ScopedTimer t(ctx, "direct");
                        GI_LAUNCH_M(k_tail_shadow, grid_for((size_t)n_active * Q.smax, GI_BLOCK), GI_BLOCK, ctx->S, *P, n_active, in, PS, Q, work_ptr(ctx, 2));
                    }
                    {
                        ScopedTimer t(ctx, "tail");
                        k_tail_direct<<<grid_for(n_active, 256), 256, 0, ctx->stream>>>(n_active, ctx->S.n_lights, in, PS, Q);
                    }
                    ctx->stream = ctx->main_stream;
                    launches += 2;
                }
                if (Q.qmax) {
                    if (tail_sides) { ctx->stream = ctx->side[1]; cudaStreamWaitEvent(ctx->side[1], ctx->fork_ev, 0); }
                    int rcg;
                    {
                        ScopedTimer t(ctx, "gather");
                        rcg = run_gather(ctx, (uint32_t)slots, Q.pos, Q.dir, P->k_photons, Q.rgb, nullptr, nullptr, nullptr, nullptr, nullptr, &launches);
                    }
                    if (rcg == GI_OK) {
                        ScopedTimer t(ctx, "tail");
                        k_tail_caustic<<<grid_for(n_active, 256), 256, 0, ctx->stream>>>(n_active, in, PS, Q);
                        launches++;
                    }
                    ctx->stream = ctx->main_stream;
                    if (rcg != GI_OK) return rcg;
                }
                if (overlap && (Q.smax || (tail_sides && Q.qmax))) {
                    cudaEventRecord(ctx->side_done[tp][0], ctx->side[0]); cudaEventRecord(ctx->side_done[tp][1], ctx->side[1]); pending[tp] = true;
                }
                CK(cudaGetLastError());
                break;
            }
            CK(cudaMemsetAsync(C, 0, sizeof(DCounters), ctx->stream));
            CK(cudaMemsetAsync(ctx->b_misc.p, 0, 4, ctx->stream));   // the persistent warps' ray counter
            {
                ScopedTimer t(ctx, "bounce");
                if (persistent) {
                    const unsigned grid = std::min<unsigned>(grid_for(n_active, GI_BLOCK), (unsigned)ctx->n_sm * GI_MINB);
                    GI_LAUNCH_M(k_bounce_p, grid, GI_BLOCK, ctx->S, *P, depth, n_active, in, perm, out, H, PS, C, work_ptr(ctx, 0), ctx->b_misc.as<uint32_t>());
                } else GI_LAUNCH_M(k_bounce, grid_for(n_active, GI_BLOCK), GI_BLOCK, ctx->S, *P, depth, n_active, in, perm, out, H, PS, C, work_ptr(ctx, 0));
            }
            CK(cudaGetLastError());
            DCounters hc;
            CK(cudaMemcpyAsync(&hc, C, sizeof(DCounters), cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            launches++;
            n_closest += n_active;
            if (hc.n_hits) {
                if (hc.n_hits < (1u << 20)) had_short = true;   // (the hint does not follow the overlap_threshold knob: bench.py turns that to 0 for its single-stream frame)
                // the host has just synchronised the main stream (counters), so the side streams need no event to start
                const bool side = deferred || (overlap && hc.n_hits < ctx->overlap_threshold);
                static const bool big_direct = getenv("GI_NO_BIG_DIRECT_OVERLAP") == nullptr;   // long shadow launches run beside the gather pipeline (C2 25.85 -> 25.43 ms, glass 51.0 -> 50.7)
                const bool side_d = side || (overlap && big_direct);
                if (ctx->S.n_lights) {
                    if (side_d) ctx->stream = ctx->side[0];
                    {
                        ScopedTimer t(ctx, "direct");
                        GI_LAUNCH_M(k_direct, grid_for(hc.n_hits, GI_BLOCK), GI_BLOCK, ctx->S, *P, depth, hc.n_hits, H, PS, work_ptr(ctx, 2));
                    }
                    ctx->stream = ctx->main_stream;
                    launches++;
                    n_shadow += (uint64_t)hc.n_hits * ctx->S.n_lights;
                }
                if (depth <= P->caustic_max_depth) {
                    n_gather += hc.n_hits;   // samplePhotons is called whether or not photons exist (raytracer.h:258)
                    if (have_map) {
                        if (side) ctx->stream = ctx->side[1];
                        int rcg;
                        {
                            ScopedTimer t(ctx, "gather");
                            rcg = run_gather(ctx, hc.n_hits, H.p, H.refdir, P->k_photons, nullptr, nullptr, nullptr, H.wcaustic, PS.Lc, H.path, &launches);
                        }
                        ctx->stream = ctx->main_stream;
                        if (rcg != GI_OK) return rcg;
                    }
                }
                if (side || side_d) { cudaEventRecord(ctx->side_done[hpar][0], ctx->side[0]); cudaEventRecord(ctx->side_done[hpar][1], ctx->side[1]); pending[hpar] = true; }
                CK(cudaGetLastError());
            }
            if (getenv("GI_TRACE_LAUNCHES")) {
                unsigned long long wk[8];
                cudaMemcpy(wk, ctx->b_work.p, sizeof(wk), cudaMemcpyDeviceToHost);
                fprintf(stderr, "[gi] depth %2d active %9u hits %9u next %9u | cum closest nodes %llu prims %llu | shadow nodes %llu prims %llu\n", depth, n_active, hc.n_hits, hc.n_next, wk[0], wk[1], wk[2], wk[3]);
            }
            n_active = hc.n_next;
            std::swap(in, out);
            perm = nullptr;
            if (binning && n_active >= ctx->bin_threshold && depth + 1 <= P->max_depth) {
                // bin the scattered rays by origin cell and direction octant (k_bin_*): the next bounce reads them through `perm`
                ScopedTimer t(ctx, "bin");
                CK(cudaMemsetAsync(ctx->b_binhist.p, 0, (size_t)GI_SORT_BINS * 4, ctx->stream));
                k_bin_keys<<<grid_for(n_active, 256), 256, 0, ctx->stream>>>(n_active, in.o, in.d, bin_min, bin_inv, ctx->b_binkey.as<uint32_t>(), ctx->b_binhist.as<uint32_t>());
                int rcs = scan_exclusive(ctx, ctx->b_binhist.as<uint32_t>(), 1, GI_SORT_BINS, ctx->b_bincur.as<uint32_t>(), nullptr);
                if (rcs != GI_OK) return rcs;
                k_bin_scatter<<<grid_for(n_active, 256), 256, 0, ctx->stream>>>(n_active, ctx->b_binkey.as<uint32_t>(), ctx->b_bincur.as<uint32_t>(), ctx->b_binperm.as<uint32_t>());
                CK(cudaGetLastError());
                launches += 5;
                perm = ctx->b_binperm.as<uint32_t>();
            }
        }
        wait_all_sides();   // whatever follows on the main stream (accumulate, the next chunk) sees complete sums
        return GI_OK;
    };
    if (adapt) {
        // passes over the sample index; pass s renders sample s of the pixels that are still active (raytracer.h:108)
        if (npx > chunk_cap) return fail(ctx, GI_ERR_INVALID, "adaptive tiles are limited to 2^23 pixels");
        k_adapt_init<<<grid_for(npx, 256), 256, 0, ctx->stream>>>(npx, *adapt);
        for (int s = 0; s < adapt->max_samples; s++) {
            GI_POLL_CANCEL("adaptive render");
            CK(cudaMemsetAsync(adapt->n_list, 0, 4, ctx->stream));
            k_adapt_select<<<grid_for(npx, 256), 256, 0, ctx->stream>>>((uint32_t)npx, s, *adapt);
            uint32_t n = 0;
            CK(cudaMemcpyAsync(&n, adapt->n_list, 4, cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            launches += 2;
            if (!n) break;
            k_generate_list<<<grid_for(n, 256), 256, 0, ctx->stream>>>(ctx->S, F, s, n, adapt->list, qa, PS);
            launches++;
            int rcd = run_depths(n);
            if (rcd != GI_OK) return rcd;
            k_adapt_update<<<grid_for(n, 256), 256, 0, ctx->stream>>>(n, s, adapt->list, PS.L, PS.Ld, PS.Lc, *adapt);
            launches++;
            CK(cudaGetLastError());
        }
        CK(cudaMemcpyAsync(accum_dev, adapt->color, npx * 24, cudaMemcpyDeviceToDevice, ctx->stream));
    } else {
    CK(cudaMemsetAsync(accum_dev, 0, npx * 24, ctx->stream));
    for (uint64_t c0 = 0; c0 < total_paths; c0 += chunk_cap) {
        uint32_t n = (uint32_t)std::min<uint64_t>(chunk_cap, total_paths - c0);
        GI_POLL_CANCEL("render");
        k_generate<<<grid_for(n, 256), 256, 0, ctx->stream>>>(ctx->S, F, s0, c0, n, qa, PS);
        launches++;
        int rcd = run_depths(n);
        if (rcd != GI_OK) return rcd;
        k_accumulate<<<grid_for(npx, 256), 256, 0, ctx->stream>>>(c0, n, npx, F.tw, F.th, PS.L, PS.Ld, PS.Lc, accum_dev);
        launches++;
        CK(cudaGetLastError());
    }
    }
    cudaEventRecord(e1, ctx->stream);
    DTailCounters tc;
    CK(cudaMemcpyAsync(&tc, ctx->b_tail.p, sizeof(tc), cudaMemcpyDeviceToHost, ctx->stream));
    int rc = gi_synchronize(ctx);
    if (rc != GI_OK) return rc;
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    n_closest += tc.closest; n_shadow += tc.shadow; n_gather += tc.gathers;
    if (tunable) ctx->sched_hint = had_short ? 1 : 0;
    if (n_closest && tunable) {
        ctx->nodes_per_ray = (double)(ctx->work_host[0] + tc.nodes_c) / (double)n_closest;
        ctx->prims_per_ray = (double)(ctx->work_host[1] + tc.prims_c) / (double)n_closest;
    }
    ctx->work_host[0] += tc.nodes_c; ctx->work_host[1] += tc.prims_c; ctx->work_host[2] += tc.nodes_s; ctx->work_host[3] += tc.prims_s;
    ctx->work_host[4] += tc.g_depth; ctx->work_host[5] += tc.g_cand; ctx->work_host[6] += tc.g_sel;
    if (stats) {
        std::memset(stats, 0, sizeof(*stats));
        stats->closest_rays = n_closest; stats->shadow_rays = n_shadow; stats->gathers = n_gather; stats->kernel_launches = launches;
        stats->trace_ms = ctx->fam["bounce"].ms; stats->shadow_ms = ctx->fam["direct"].ms; stats->gather_ms = ctx->fam["gather"].ms; stats->shade_ms = ctx->fam["tail"].ms; stats->total_ms = ms;
        const unsigned long long* w = ctx->work_host;
        stats->closest_node_tests = w[0]; stats->closest_prim_tests = w[1]; stats->shadow_node_tests = w[2]; stats->shadow_prim_tests = w[3];
        stats->gather_leaf_depth = w[4]; stats->gather_candidates = w[5]; stats->gather_selected = w[6];
        stats->tail_closest_rays = tc.closest;
        stats->tail_shadow_rays = 0;   // the tail's shadow rays are queued, then traced (and tallied) by the batched any-hit kernel k_tail_shadow
        stats->tail_gathers = (ctx->tail_mode == 1 || !have_map) ? tc.gathers : 0;   // queued tail gathers are served (and tallied) by the gather pipeline
        stats->tail_closest_node_tests = tc.nodes_c; stats->tail_closest_prim_tests = tc.prims_c; stats->tail_shadow_node_tests = tc.nodes_s; stats->tail_shadow_prim_tests = tc.prims_s;
        stats->tail_gather_leaf_depth = tc.g_depth; stats->tail_gather_candidates = tc.g_cand; stats->tail_gather_selected = tc.g_sel;
        stats->bin_ms = ctx->fam["bin"].ms;
    }
    return GI_OK;
}

static int check_render_args(gi_ctx* ctx, const gi_render_params* P, int x0, int y0, int x1, int y1, int s0, int s1, const double* accum)
{
    if (!ctx || !P || !accum) return GI_ERR_INVALID;
    if (P->width <= 0 || P->height <= 0 || x0 < 0 || y0 < 0 || x1 > P->width || y1 > P->height || x1 <= x0 || y1 <= y0 || s1 <= s0 || s0 < 0) return fail(ctx, GI_ERR_INVALID, "bad tile / sample range");
    if (P->k_photons < 1 || P->k_photons > 32 || P->max_depth < 0 || P->max_depth > 126) return fail(ctx, GI_ERR_INVALID, "k_photons must be 1..32 and max_depth 0..126");
    if (!ctx->has_scene) return fail(ctx, GI_ERR_NO_SCENE, "gi_render_tile needs gi_scene_upload");
    return GI_OK;
}

extern "C" int gi_render_tile_dev(gi_ctx* ctx, const gi_render_params* P, int x0, int y0, int x1, int y1, int s0, int s1, double* accum, gi_stats* stats)
{
    int rc = check_render_args(ctx, P, x0, y0, x1, y1, s0, s1, accum);
    if (rc != GI_OK) return rc;
    CK(cudaSetDevice(ctx->device));
    return render_device(ctx, P, x0, y0, x1, y1, s0, s1, accum, stats);
}
extern "C" int gi_render_tile(gi_ctx* ctx, const gi_render_params* P, int x0, int y0, int x1, int y1, int s0, int s1, double* accum, gi_stats* stats)
{
    int rc = check_render_args(ctx, P, x0, y0, x1, y1, s0, s1, accum);
    if (rc != GI_OK) return rc;
    CK(cudaSetDevice(ctx->device));
    size_t npx = (size_t)(x1 - x0) * (y1 - y0);
    CK(ctx->b_accum.reserve(npx * 24));
    rc = render_device(ctx, P, x0, y0, x1, y1, s0, s1, ctx->b_accum.as<double>(), stats);
    if (rc != GI_OK) return rc;
    CK(cudaMemcpyAsync(accum, ctx->b_accum.p, npx * 24, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return GI_OK;
}

// The frame as the reference's run() leaves it: rendered, resolved to 8-bit on the device, only the image (and, if asked for,
// the fp64 sums) copied back — no host round trip of the accumulator between render and resolve.
extern "C" int gi_render_image(gi_ctx* ctx, const gi_render_params* P, int x0, int y0, int x1, int y1, int s0, int s1, uint8_t* rgb8, double* accum, gi_stats* stats)
{
    if (!rgb8) return GI_ERR_INVALID;
    int rc = check_render_args(ctx, P, x0, y0, x1, y1, s0, s1, reinterpret_cast<const double*>(rgb8));
    if (rc != GI_OK) return rc;
    CK(cudaSetDevice(ctx->device));
    const size_t npx = (size_t)(x1 - x0) * (y1 - y0);
    CK(ctx->b_accum.reserve(npx * 24)); CK(ctx->w1.reserve(npx * 3));
    rc = render_device(ctx, P, x0, y0, x1, y1, s0, s1, ctx->b_accum.as<double>(), stats);
    if (rc != GI_OK) return rc;
    rc = gi_resolve_dev(ctx, npx, ctx->b_accum.as<double>(), s1 - s0, ctx->w1.as<uint8_t>());
    if (rc != GI_OK) return rc;
    CK(cudaMemcpyAsync(rgb8, ctx->w1.p, npx * 3, cudaMemcpyDeviceToHost, ctx->stream));
    if (accum) CK(cudaMemcpyAsync(accum, ctx->b_accum.p, npx * 24, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return GI_OK;
}

// ---- tile split: the rows of one part of the frame under the interleaved row-block plan, rendered in ONE wavefront.
// accum [local rows][width][3] compact; a part's pixels are exactly those of the one-GPU frame (same Halton indices, same PRNG keys).
extern "C" int gi_rows_of_part(int height, int block_rows, int nparts, int part)
{
    if (height <= 0 || block_rows <= 0 || nparts <= 0 || part < 0 || part >= nparts) return GI_ERR_INVALID;
    return rows_of_part(height, block_rows, nparts, part);
}
static int check_rows_args(gi_ctx* ctx, const gi_render_params* P, int block_rows, int nparts, int part, int s0, int s1, const void* out)
{
    if (!ctx || !P || !out) return GI_ERR_INVALID;
    if (block_rows <= 0 || nparts <= 0 || part < 0 || part >= nparts) return fail(ctx, GI_ERR_INVALID, "bad row plan");
    int rc = check_render_args(ctx, P, 0, 0, P->width, P->height, s0, s1, reinterpret_cast<const double*>(out));
    if (rc != GI_OK) return rc;
    if (rows_of_part(P->height, block_rows, nparts, part) == 0) return fail(ctx, GI_ERR_INVALID, "this part of the row plan is empty");
    return GI_OK;
}
extern "C" int gi_render_rows_dev(gi_ctx* ctx, const gi_render_params* P, int block_rows, int nparts, int part, int s0, int s1, double* accum, gi_stats* stats)
{
    int rc = check_rows_args(ctx, P, block_rows, nparts, part, s0, s1, accum);
    if (rc != GI_OK) return rc;
    CK(cudaSetDevice(ctx->device));
    const RowPlan rp{ block_rows, nparts, part };
    return render_device(ctx, P, 0, 0, 0, 0, s0, s1, accum, stats, nullptr, &rp);
}
extern "C" int gi_render_rows(gi_ctx* ctx, const gi_render_params* P, int block_rows, int nparts, int part, int s0, int s1, double* accum, gi_stats* stats)
{
    int rc = check_rows_args(ctx, P, block_rows, nparts, part, s0, s1, accum);
    if (rc != GI_OK) return rc;
    CK(cudaSetDevice(ctx->device));
    const size_t npx = (size_t)rows_of_part(P->height, block_rows, nparts, part) * (size_t)P->width;
    CK(ctx->b_accum.reserve(npx * 24));
    const RowPlan rp{ block_rows, nparts, part };
    rc = render_device(ctx, P, 0, 0, 0, 0, s0, s1, ctx->b_accum.as<double>(), stats, nullptr, &rp);
    if (rc != GI_OK) return rc;
    CK(cudaMemcpyAsync(accum, ctx->b_accum.p, npx * 24, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return GI_OK;
}

extern "C" int gi_render_adaptive_dev(gi_ctx* ctx, const gi_render_params* P, int min_samples, int max_samples, double noise_thresh, int x0, int y0, int x1, int y1, double* color,
                                      uint32_t* samples, gi_stats* stats)
{
    int rc = check_render_args(ctx, P, x0, y0, x1, y1, 0, 1, color);
    if (rc != GI_OK) return rc;
    if (min_samples < 0 || max_samples < 0) return fail(ctx, GI_ERR_INVALID, "negative sample counts");
    CK(cudaSetDevice(ctx->device));
    const size_t npx = (size_t)(x1 - x0) * (y1 - y0);
    CK(ctx->ad[0].reserve(npx * 24)); CK(ctx->ad[1].reserve(npx * 8)); CK(ctx->ad[2].reserve(npx * 4)); CK(ctx->ad[3].reserve(npx * 4)); CK(ctx->ad[4].reserve(npx * 4)); CK(ctx->ad[5].reserve(16));
    DAdapt A{ ctx->ad[0].as<double>(), ctx->ad[1].as<double>(), ctx->ad[2].as<int>(), ctx->ad[3].as<int>(), ctx->ad[4].as<uint32_t>(), ctx->ad[5].as<uint32_t>(), min_samples, max_samples, noise_thresh };
    rc = render_device(ctx, P, x0, y0, x1, y1, 0, 1, color, stats, &A);
    if (rc != GI_OK) return rc;
    if (samples) CK(cudaMemcpyAsync(samples, A.s_done, npx * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return GI_OK;
}
extern "C" int gi_render_adaptive(gi_ctx* ctx, const gi_render_params* P, int min_samples, int max_samples, double noise_thresh, int x0, int y0, int x1, int y1, double* color,
                                  uint32_t* samples, gi_stats* stats)
{
    int rc = check_render_args(ctx, P, x0, y0, x1, y1, 0, 1, color);
    if (rc != GI_OK) return rc;
    CK(cudaSetDevice(ctx->device));
    const size_t npx = (size_t)(x1 - x0) * (y1 - y0);
    CK(ctx->b_accum.reserve(npx * 24)); CK(ctx->w9.reserve(npx * 4));
    rc = gi_render_adaptive_dev(ctx, P, min_samples, max_samples, noise_thresh, x0, y0, x1, y1, ctx->b_accum.as<double>(), ctx->w9.as<uint32_t>(), stats);
    if (rc != GI_OK) return rc;
    CK(cudaMemcpyAsync(color, ctx->b_accum.p, npx * 24, cudaMemcpyDeviceToHost, ctx->stream));
    if (samples) CK(cudaMemcpyAsync(samples, ctx->w9.p, npx * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return GI_OK;
}

extern "C" int gi_resolve_dev(gi_ctx* ctx, size_t n_pixels, const double* accum, int spp, uint8_t* rgb8)
{
    if (!ctx || !accum || !rgb8 || spp <= 0) return GI_ERR_INVALID;
    if (!n_pixels) return GI_OK;
    CK(cudaSetDevice(ctx->device));
    k_resolve<<<grid_for(n_pixels * 3, 256), 256, 0, ctx->stream>>>(n_pixels * 3, accum, spp, rgb8);
    CK(cudaGetLastError());
    return GI_OK;
}
extern "C" int gi_resolve(gi_ctx* ctx, size_t n_pixels, const double* accum, int spp, uint8_t* rgb8)
{
    if (!ctx || !accum || !rgb8 || spp <= 0) return GI_ERR_INVALID;
    if (!n_pixels) return GI_OK;
    CK(cudaSetDevice(ctx->device));
    CK(ctx->w0.reserve(n_pixels * 24)); CK(ctx->w1.reserve(n_pixels * 3));
    CK(cudaMemcpyAsync(ctx->w0.p, accum, n_pixels * 24, cudaMemcpyHostToDevice, ctx->stream));
    int rc = gi_resolve_dev(ctx, n_pixels, ctx->w0.as<double>(), spp, ctx->w1.as<uint8_t>());
    if (rc != GI_OK) return rc;
    CK(cudaMemcpyAsync(rgb8, ctx->w1.p, n_pixels * 3, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return GI_OK;
}

// ---- scene-API queries, batch forms (host pointers; gi_query.cuh) ---------------------------------------------------------------------------
static int upload_rays(gi_ctx* ctx, size_t n, const double* org, const double* dir, const double* tmin, const double* tmax)
{
    CK(ctx->w0.reserve(n * 24)); CK(ctx->w1.reserve(n * 24)); CK(ctx->w2.reserve(n * 8)); CK(ctx->w3.reserve(n * 8));
    CK(cudaMemcpyAsync(ctx->w0.p, org, n * 24, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->w1.p, dir, n * 24, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->w2.p, tmin, n * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->w3.p, tmax, n * 8, cudaMemcpyHostToDevice, ctx->stream));
    return GI_OK;
}
extern "C" int gi_octree_intersect(gi_ctx* ctx, size_t n, const double* org, const double* dir, const double* tmin, const double* tmax, uint32_t cap, uint32_t* prim_ids, uint32_t* counts)
{
    if (!ctx || (n && (!org || !dir || !tmin || !tmax || !counts || (cap && !prim_ids)))) return GI_ERR_INVALID;
    if (!ctx->has_scene) return fail(ctx, GI_ERR_NO_SCENE, "gi_octree_intersect needs gi_scene_upload");
    if (!n) return GI_OK;
    CK(cudaSetDevice(ctx->device));
    int rc = upload_rays(ctx, n, org, dir, tmin, tmax);
    if (rc != GI_OK) return rc;
    CK(ctx->w4.reserve(std::max<size_t>(n * cap, 1) * 4)); CK(ctx->w5.reserve(n * 4));
    k_octree_intersect<<<grid_for(n, 128), 128, 0, ctx->stream>>>(ctx->S, n, ctx->w0.as<double>(), ctx->w1.as<double>(), ctx->w2.as<double>(), ctx->w3.as<double>(), cap, ctx->w4.as<uint32_t>(),
                                                                 ctx->w5.as<uint32_t>());
    CK(cudaGetLastError());
    if (cap) CK(cudaMemcpyAsync(prim_ids, ctx->w4.p, n * cap * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(counts, ctx->w5.p, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    return gi_synchronize(ctx);
}
extern "C" int gi_octree_intersect_sorted(gi_ctx* ctx, size_t n, const double* org, const double* dir, const double* tmin, const double* tmax, uint32_t cap, uint32_t* node_ids, double* t0,
                                          uint32_t* counts)
{
    if (!ctx || (n && (!org || !dir || !tmin || !tmax || !counts || !cap || !node_ids || !t0))) return GI_ERR_INVALID;
    if (!ctx->has_scene) return fail(ctx, GI_ERR_NO_SCENE, "gi_octree_intersect_sorted needs gi_scene_upload");
    if (!n) return GI_OK;
    CK(cudaSetDevice(ctx->device));
    int rc = upload_rays(ctx, n, org, dir, tmin, tmax);
    if (rc != GI_OK) return rc;
    CK(ctx->w4.reserve(n * cap * 4)); CK(ctx->w5.reserve(n * 4)); CK(ctx->w6.reserve(n * cap * 8));
    k_octree_intersect_sorted<<<grid_for(n, 128), 128, 0, ctx->stream>>>(ctx->S, n, ctx->w0.as<double>(), ctx->w1.as<double>(), ctx->w2.as<double>(), ctx->w3.as<double>(), cap,
                                                                        ctx->w4.as<uint32_t>(), ctx->w6.as<double>(), ctx->w5.as<uint32_t>());
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(node_ids, ctx->w4.p, n * cap * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(t0, ctx->w6.p, n * cap * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(counts, ctx->w5.p, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    return gi_synchronize(ctx);
}
extern "C" int gi_photon_in_range(gi_ctx* ctx, size_t n, const double* pos, uint32_t cap, uint32_t* photon_ids, uint32_t* counts)
{
    if (!ctx || (n && (!pos || !counts || (cap && !photon_ids)))) return GI_ERR_INVALID;
    if (!ctx->has_map) return fail(ctx, GI_ERR_NO_PHOTONS, "gi_photon_in_range needs gi_photon_map_build");
    if (!n) return GI_OK;
    CK(cudaSetDevice(ctx->device));
    CK(ctx->w0.reserve(n * 24)); CK(ctx->w4.reserve(std::max<size_t>(n * cap, 1) * 4)); CK(ctx->w5.reserve(n * 4)); CK(ctx->w3.reserve(16));
    CK(cudaMemcpyAsync(ctx->w0.p, pos, n * 24, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemsetAsync(ctx->w3.p, 0, 4, ctx->stream));
    k_photon_in_range<<<grid_for(n, GI_WPB), GI_WPB * 32, 0, ctx->stream>>>(ctx->G, n, ctx->w0.as<double>(), cap, ctx->w4.as<uint32_t>(), ctx->w5.as<uint32_t>(), ctx->w3.as<uint32_t>());
    CK(cudaGetLastError());
    uint32_t ovf = 0;
    if (cap) CK(cudaMemcpyAsync(photon_ids, ctx->w4.p, n * cap * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(counts, ctx->w5.p, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(&ovf, ctx->w3.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
    int rc = gi_synchronize(ctx);
    if (rc != GI_OK) return rc;
    if (ovf) return fail(ctx, GI_ERR_INVALID, "photon map too deep for the candidate walk");
    return GI_OK;
}
extern "C" int gi_prim_intersect(gi_ctx* ctx, size_t n, const uint32_t* prim, const double* org, const double* dir, uint8_t* ok, double* hit, double* normal, double* uv, uint8_t* wrote_uv)
{
    if (!ctx || (n && (!prim || !org || !dir || !ok || !hit || !normal || !uv || !wrote_uv))) return GI_ERR_INVALID;
    if (!ctx->has_scene) return fail(ctx, GI_ERR_NO_SCENE, "gi_prim_intersect needs gi_scene_upload");
    if (!n) return GI_OK;
    CK(cudaSetDevice(ctx->device));
    CK(ctx->w0.reserve(n * 24)); CK(ctx->w1.reserve(n * 24)); CK(ctx->w2.reserve(n * 4)); CK(ctx->w3.reserve(n * 2)); CK(ctx->w4.reserve(n * 24)); CK(ctx->w5.reserve(n * 24)); CK(ctx->w6.reserve(n * 16));
    CK(cudaMemcpyAsync(ctx->w0.p, org, n * 24, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->w1.p, dir, n * 24, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->w2.p, prim, n * 4, cudaMemcpyHostToDevice, ctx->stream));
    uint8_t* flags = ctx->w3.as<uint8_t>();
    k_prim_intersect<<<grid_for(n, 128), 128, 0, ctx->stream>>>(ctx->S, n, ctx->w2.as<uint32_t>(), ctx->w0.as<double>(), ctx->w1.as<double>(), flags, ctx->w4.as<double>(), ctx->w5.as<double>(),
                                                               ctx->w6.as<double>(), flags + n);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(ok, flags, n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(wrote_uv, flags + n, n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(hit, ctx->w4.p, n * 24, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(normal, ctx->w5.p, n * 24, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(uv, ctx->w6.p, n * 16, cudaMemcpyDeviceToHost, ctx->stream));
    return gi_synchronize(ctx);
}

#include "gi_comm.inc"

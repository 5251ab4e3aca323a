// gi_kernels.cuh — the sm_100a kernels of the hot path (K1..K8 of SURVEY §2.2).  One thread per ray / path / photon,
// one warp per gather query.  All kernels are grid-stride free: the host sizes the grid to the queue length.
#pragma once
#include <math_constants.h>

#include "gi_device.cuh"

#define GI_BLOCK 128
#ifndef GI_MINB
#define GI_MINB 1   // minimum resident blocks per SM requested from ptxas for the thread-per-ray traversal kernels
#endif
#define GI_PM_LEAF_MAX 16   // MAX_PHOTONS_PER_LEAF (util.h:15)

// ---- K0: Halton known-answer entry points -----------------------------------------------------------------------------------
__global__ void k_halton_sample(DScene S, size_t n, const uint32_t* dim, const uint32_t* index, float* out)
{
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) out[i] = halton_sample(S, dim[i], index[i]);
}
__global__ void k_halton_index(DHEnum he, size_t n, const uint32_t* s, const uint32_t* x, const uint32_t* y, uint32_t* out)
{
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) out[i] = henum_index(he, s[i], x[i], y[i]);
}

// ---- K1: camera rays (raytracer.h:74-78, 112-129) ------------------------------------------------------------------------------
struct DFrame { int w, h, x0, y0, tw, th; double halfW, halfH; d3 center, right, up, pos; DHEnum he; };

__device__ __forceinline__ DRay camera_ray(const DScene& S, const DFrame& F, int x, int y, int s, uint32_t& idx_out)
{
    int idx = (int)henum_index(F.he, (uint32_t)s, (uint32_t)x, (uint32_t)y);
    double xr = halton_sample(S, 0, (uint32_t)idx);
    double yr = halton_sample(S, 1, (uint32_t)idx);
    double dx = (double)__fmul_rn((float)xr, F.he.scale_x);   // Halton_enum::scale_x/y are float -> float (halton_enum.h:116-124)
    double dy = (double)__fmul_rn((float)yr, F.he.scale_y);
    d3 pixelPos = (F.center + F.right * (F.halfW * (dx / F.w - .5))) - F.up * (F.halfH * (dy / F.h - .5));
    d3 eye = (F.pos + F.right * (0 * (xr - .5))) + F.up * (0 * (yr - .5));   // FOCAL_BLUR = 0 (util.h:30)
    idx_out = (uint32_t)idx;
    return make_ray(eye, normalize3(pixelPos - eye));
}

__global__ void k_camera_rays(DScene S, DFrame F, int s0, size_t n, double* org, double* dir, uint32_t* index)
{
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    size_t npx = (size_t)F.tw * F.th;
    int s = s0 + (int)(i / npx);
    size_t p = i % npx;
    int y = F.y0 + (int)(p / F.tw), x = F.x0 + (int)(p % F.tw);
    uint32_t idx;
    DRay r = camera_ray(S, F, x, y, s, idx);
    st3(org + 3 * i, r.o); st3(dir + 3 * i, r.d);
    if (index) index[i] = idx;
}

// ---- K2/K3 (API form): batch closest hit / any hit over caller-supplied rays --------------------------------------------------------
template <bool FULL, bool IMPL>
__global__ void __launch_bounds__(GI_BLOCK, GI_MINB) k_trace_closest(DScene S, size_t n, const double* __restrict__ org, const double* __restrict__ dir, uint64_t seed,
                                                           uint32_t* __restrict__ prim, double* __restrict__ hit, double* __restrict__ normal, double* __restrict__ uv,
                                                           unsigned long long* work)
{
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    uint32_t nn_ = 0, np_ = 0;
    if (i < n) {
        DRay r = ray_as_stored(ld3(org + 3 * i), ld3(dir + 3 * i));
        DHit h;
        trace_closest<FULL, IMPL>(S, r, seed, (uint64_t)i, 0, h, nn_, np_);
        if (prim) prim[i] = h.prim;
        d3 p = mk3(0, 0, 0), nn = mk3(0, 0, 0); double tu = 0, tv = 0;
        if (h.prim != GI_NO_HIT) hit_surface(S, r, h, FULL, p, nn, tu, tv);
        if (hit) st3(hit + 3 * i, p);
        if (normal) st3(normal + 3 * i, nn);
        if (uv) { uv[2 * i] = tu; uv[2 * i + 1] = tv; }
    }
    tally2(work, nn_, np_);
}

template <bool FULL, bool IMPL>
__global__ void __launch_bounds__(GI_BLOCK, GI_MINB) k_trace_any(DScene S, size_t n, const double* __restrict__ org, const double* __restrict__ dir, const double* __restrict__ maxt2,
                                                       uint64_t seed, uint8_t* __restrict__ vis, unsigned long long* work)
{
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    uint32_t nn_ = 0, np_ = 0;
    if (i < n) {
        DRay r = ray_as_stored(ld3(org + 3 * i), ld3(dir + 3 * i));
        vis[i] = trace_visible<FULL, IMPL>(S, r, maxt2[i], seed, (uint64_t)i, 0, 0, nn_, np_) ? 1 : 0;
    }
    tally2(work, nn_, np_);
}

// warp-per-ray forms of the two batch kernels (one ray per warp, see trace_closest_warp)
#define GI_WPB 4   // warps per block in the warp-per-ray kernels
template <bool FULL, bool IMPL>
__global__ void __launch_bounds__(GI_WPB * 32) k_trace_closest_w(DScene S, size_t n, const double* __restrict__ org, const double* __restrict__ dir, uint64_t seed,
                                                                uint32_t* __restrict__ prim, double* __restrict__ hit, double* __restrict__ normal, double* __restrict__ uv,
                                                                unsigned long long* work)
{
    __shared__ uint32_t s_stack[GI_WPB][GI_STACK_MAX];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    size_t i = blockIdx.x * (size_t)GI_WPB + wib;
    if (i >= n) return;
    uint32_t nn_ = 0, np_ = 0;
    DRay r = ray_as_stored(ld3(org + 3 * i), ld3(dir + 3 * i));
    DHit h;
    trace_closest_warp<FULL, IMPL>(S, r, seed, (uint64_t)i, 0, h, s_stack[wib], lane, nn_, np_);
    if (lane == 0) {
        if (prim) prim[i] = h.prim;
        d3 p = mk3(0, 0, 0), nn = mk3(0, 0, 0); double tu = 0, tv = 0;
        if (h.prim != GI_NO_HIT) hit_surface(S, r, h, FULL, p, nn, tu, tv);
        if (hit) st3(hit + 3 * i, p);
        if (normal) st3(normal + 3 * i, nn);
        if (uv) { uv[2 * i] = tu; uv[2 * i + 1] = tv; }
        if (work) { atomicAdd(work, (unsigned long long)nn_); atomicAdd(work + 1, (unsigned long long)np_); }
    }
}
template <bool FULL, bool IMPL>
__global__ void __launch_bounds__(GI_WPB * 32) k_trace_any_w(DScene S, size_t n, const double* __restrict__ org, const double* __restrict__ dir, const double* __restrict__ maxt2,
                                                            uint64_t seed, uint8_t* __restrict__ vis, unsigned long long* work)
{
    __shared__ uint32_t s_stack[GI_WPB][GI_STACK_MAX];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    size_t i = blockIdx.x * (size_t)GI_WPB + wib;
    if (i >= n) return;
    uint32_t nn_ = 0, np_ = 0;
    DRay r = ray_as_stored(ld3(org + 3 * i), ld3(dir + 3 * i));
    bool v = trace_visible_warp<FULL, IMPL>(S, r, maxt2[i], seed, (uint64_t)i, 0, 0, s_stack[wib], lane, nn_, np_);
    if (lane == 0) {
        vis[i] = v ? 1 : 0;
        if (work) { atomicAdd(work, (unsigned long long)nn_); atomicAdd(work + 1, (unsigned long long)np_); }
    }
}

// ---- photon map (K6): level-synchronous build of PhotonMap::Node::partition (photonMap.cpp:137-192) --------------------------------
// Nodes reuse DNode: mask != 0 marks an interior node whose 8 children start at `child`; prim_off/prim_cnt = photon range.
struct DPMap {
    DNode* nodes;
    uint32_t* n_nodes;          // device counter
    uint32_t cap_nodes;
    const double* ph;           // original photons [n][9]
    uint32_t n_photons;
    uint32_t* pnode;            // node of each photon during the build (0xFFFFFFFF = dropped)
    // leaf-ordered photon storage (after finalize)
    double* pos;                // [kept][3]
    double* dircol;             // [kept][6]
    uint32_t* pid;              // [kept] original photon index
    uint32_t* n_kept;
    uint32_t* overflow;
};

__device__ __forceinline__ void child_box(const double* bmin, const double* bmax, int i, double* cmin, double* cmax)
{
    // the eight child boxes exactly as written in photonMap.cpp:139-149 / octree.cpp:318-328
    double mx = bmin[0] + .5 * (bmax[0] - bmin[0]), my = bmin[1] + .5 * (bmax[1] - bmin[1]), mz = bmin[2] + .5 * (bmax[2] - bmin[2]);
    double hx = .5 * (bmax[0] - bmin[0]), hy = .5 * (bmax[1] - bmin[1]), hz = .5 * (bmax[2] - bmin[2]);
    if (i == 7) { cmin[0] = mx; cmin[1] = my; cmin[2] = mz; cmax[0] = bmax[0]; cmax[1] = bmax[1]; cmax[2] = bmax[2]; return; }
    cmin[0] = bmin[0]; cmin[1] = bmin[1]; cmin[2] = bmin[2]; cmax[0] = mx; cmax[1] = my; cmax[2] = mz;
    if (i & 1) { cmin[0] = bmin[0] + hx; cmax[0] = mx + hx; }
    if (i & 2) { cmin[2] = bmin[2] + hz; cmax[2] = mz + hz; }
    if (i & 4) { cmin[1] = bmin[1] + hy; cmax[1] = my + hy; }
}

__global__ void k_pm_init(DPMap M, const double* box6)
{
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i == 0) {
        DNode r;
        for (int k = 0; k < 3; k++) { r.bmin[k] = box6[k]; r.bmax[k] = box6[3 + k]; }
        r.child = 0; r.prim_off = 0; r.prim_cnt = M.n_photons; r.mask = 0;
        M.nodes[0] = r;
        *M.n_nodes = 1; *M.n_kept = 0; *M.overflow = 0;
    }
    if (i < M.n_photons) M.pnode[i] = 0;
}

// split every node of the level that holds more than 16 photons (photonMap.cpp:186-190; root: :37)
__global__ void k_pm_split(DPMap M, uint32_t lb, uint32_t le)
{
    uint32_t i = lb + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= le) return;
    DNode nd = M.nodes[i];
    if (nd.prim_cnt <= GI_PM_LEAF_MAX) return;
    uint32_t base = atomicAdd(M.n_nodes, 8u);
    if (base + 8 > M.cap_nodes) { atomicExch(M.overflow, 1u); atomicSub(M.n_nodes, 8u); return; }
    for (int c = 0; c < 8; c++) {
        DNode ch;
        child_box(nd.bmin, nd.bmax, c, ch.bmin, ch.bmax);
        ch.child = 0; ch.prim_off = 0; ch.prim_cnt = 0; ch.mask = 0;
        M.nodes[base + c] = ch;
    }
    M.nodes[i].child = base; M.nodes[i].mask = 0xFFu; M.nodes[i].prim_cnt = 0;   // interior nodes keep no photons (:181-182)
}

// move every photon of a freshly split node into the child whose half-open box contains it (photonMap.cpp:153-166);
// a photon in no child is dropped, as in the reference
__global__ void k_pm_assign(DPMap M, uint32_t lb, uint32_t le)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M.n_photons) return;
    uint32_t nd = M.pnode[i];
    if (nd < lb || nd >= le) return;
    const DNode& N = M.nodes[nd];
    if (N.mask == 0) return;
    d3 p = ld3(M.ph + 9 * (size_t)i);
    uint32_t base = N.child, dst = 0xFFFFFFFFu;
    for (int c = 0; c < 8; c++) {
        const DNode& ch = M.nodes[base + c];
        if (box_contains(ch.bmin, ch.bmax, p)) { dst = base + c; break; }   // sibling boxes are disjoint half-open cells
    }
    M.pnode[i] = dst;
    if (dst != 0xFFFFFFFFu) atomicAdd(&M.nodes[dst].prim_cnt, 1u);
}

// exclusive scan of the per-node photon counts -> prim_off (three small kernels: block sums, scan of sums, apply)
#define GI_SCAN_BLOCK 1024
__global__ void k_scan_block(const uint32_t* in_stride16, uint32_t n, uint32_t* out, uint32_t* block_sums, int stride_words)
{
    __shared__ uint32_t sh[GI_SCAN_BLOCK];
    uint32_t i = blockIdx.x * GI_SCAN_BLOCK + threadIdx.x;
    uint32_t v = i < n ? in_stride16[(size_t)i * stride_words] : 0u;
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int off = 1; off < GI_SCAN_BLOCK; off <<= 1) {
        uint32_t t = threadIdx.x >= off ? sh[threadIdx.x - off] : 0u;
        __syncthreads();
        sh[threadIdx.x] += t;
        __syncthreads();
    }
    if (i < n) out[i] = sh[threadIdx.x] - v;
    if (threadIdx.x == GI_SCAN_BLOCK - 1) block_sums[blockIdx.x] = sh[threadIdx.x];
}
__global__ void k_scan_sums(uint32_t* block_sums, uint32_t nb, uint32_t* total)
{
    // single thread: nb is at most a few thousand
    uint32_t acc = 0;
    for (uint32_t b = 0; b < nb; b++) { uint32_t v = block_sums[b]; block_sums[b] = acc; acc += v; }
    if (total) *total = acc;
}
__global__ void k_scan_apply(uint32_t* out, uint32_t n, const uint32_t* block_sums)
{
    uint32_t i = blockIdx.x * GI_SCAN_BLOCK + threadIdx.x;
    if (i < n) out[i] += block_sums[blockIdx.x];
}

__global__ void k_pm_set_offsets(DPMap M, uint32_t n_nodes, const uint32_t* offs, uint32_t* fill)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    M.nodes[i].prim_off = offs[i];
    fill[i] = 0;
}
__global__ void k_pm_scatter(DPMap M, uint32_t* fill)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M.n_photons) return;
    uint32_t nd = M.pnode[i];
    if (nd == 0xFFFFFFFFu) return;
    uint32_t slot = M.nodes[nd].prim_off + atomicAdd(&fill[nd], 1u);
    M.pid[slot] = i;
}
// restore insertion order inside each leaf (the reference appends photons in index order) and gather the payload
__global__ void k_pm_order_leaf(DPMap M, uint32_t n_nodes)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    DNode nd = M.nodes[i];
    if (nd.mask != 0 || nd.prim_cnt == 0) return;
    uint32_t* ids = M.pid + nd.prim_off;
    for (uint32_t a = 1; a < nd.prim_cnt; a++) {
        uint32_t v = ids[a]; int b = (int)a - 1;
        while (b >= 0 && ids[b] > v) { ids[b + 1] = ids[b]; b--; }
        ids[b + 1] = v;
    }
}
__global__ void k_pm_payload(DPMap M, uint32_t n_kept)
{
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_kept) return;
    const double* src = M.ph + 9 * (size_t)M.pid[s];
    for (int k = 0; k < 3; k++) M.pos[3 * (size_t)s + k] = src[k];
    for (int k = 0; k < 6; k++) M.dircol[6 * (size_t)s + k] = src[3 + k];
}

// ---- K7: gather — RayTracer::samplePhotons (raytracer.h:532-579) over PhotonMap::getInRange (photonMap.cpp:50-92,115-134) ----
// The candidate set of a query is a function of the LEAF that contains it: every photon of every leaf whose closed box
// touches (leaf box +- EPSILON).  So the map carries, per leaf, the precomputed candidate list (photon slots in the
// reference's DFS order) — built once on the device by running Node::get for every leaf (k_pm_cand_*).  A query is then:
// descend to the leaf (lanes 0..7 test the eight children in parallel), stream the leaf's candidate list 32 at a time,
// and keep the k <= 32 nearest in a warp-wide sorted list (one entry per lane, ordered by (distance^2, photon id)).
struct DGatherMap {
    const DNode* nodes; const double* pos; const double* dircol; const uint32_t* pid; uint32_t n_nodes;
    const uint32_t* cand_off;    // [n_nodes + 1] start of each node's candidate list (only leaves have entries)
    const uint32_t* cand_slot;   // photon slots, concatenated per leaf
};

// ordering of candidates: (distance^2, photon slot).  Exact distance ties between different photons are unordered in
// the reference (std::partial_sort is unstable); the slot makes them deterministic here.
__device__ __forceinline__ bool kv_less(double a, uint32_t ai, double b, uint32_t bi) { return a < b || (a == b && ai < bi); }

// bitonic compare-exchange across lanes on (key, slot) pairs
__device__ __forceinline__ void cmpx(double& d, uint32_t& sl, int lane, int j, bool up)
{
    double od = __shfl_xor_sync(0xffffffffu, d, j);
    uint32_t osl = __shfl_xor_sync(0xffffffffu, sl, j);
    bool lower = (lane & j) == 0;
    bool mine_less = kv_less(d, sl, od, osl);
    bool take = (lower == up) ? !mine_less : mine_less;
    if (take && !(d == od && sl == osl)) { d = od; sl = osl; }
}
__device__ __forceinline__ void warp_sort32(double& d, uint32_t& sl, int lane)
{
    for (int k = 2; k <= 32; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) cmpx(d, sl, lane, j, (lane & k) == 0 || k == 32);
}
// merge a sorted-ascending batch (bd) into the sorted-ascending best list (d): keep the 32 smallest of the 64
__device__ __forceinline__ void warp_merge32(double& d, uint32_t& sl, double bd, uint32_t bsl, int lane)
{
    // reverse the batch, take the element-wise minimum -> bitonic sequence holding the 32 smallest; then bitonic merge
    double rd = __shfl_sync(0xffffffffu, bd, 31 - lane);
    uint32_t rsl = __shfl_sync(0xffffffffu, bsl, 31 - lane);
    if (kv_less(rd, rsl, d, sl)) { d = rd; sl = rsl; }
    for (int j = 16; j > 0; j >>= 1) cmpx(d, sl, lane, j, true);
}

// getBounds (photonMap.cpp:115-134): the leaf whose half-open box contains p; false when p is in no child on the way down
__device__ __forceinline__ bool pm_find_leaf(const DGatherMap& M, d3 p, int lane, uint32_t& node, uint32_t& depth)
{
    node = 0; depth = 0;
    if (M.n_nodes == 0) return false;
    uint4 top = __ldg(reinterpret_cast<const uint4*>(M.nodes) + 3);   // root topology: child, prim_off, prim_cnt, mask
    uint32_t child = top.x, mask = top.w;
    while (mask != 0) {
        depth++;
        bool in = false;
        uint32_t cchild = 0, cmask = 0;
        if (lane < 8) { DNode ch = load_node(M.nodes, child + lane); in = box_contains(ch.bmin, ch.bmax, p); cchild = ch.child; cmask = ch.mask; }
        uint32_t b = __ballot_sync(0xffffffffu, in) & 0xffu;
        if (!b) return false;   // box(-inf,-inf): nothing overlaps (photonMap.cpp:132)
        int src = __ffs(b) - 1;
        node = child + src;
        // the lane that owns the containing child already holds its topology: one dependent load per level, not two
        child = __shfl_sync(0xffffffffu, cchild, src);
        mask = __shfl_sync(0xffffffffu, cmask, src);
    }
    return true;
}

struct GatherOut { d3 rgb; uint32_t total, depth; int count; uint32_t best_slot; };

// one query, executed by a full warp; every lane returns the same rgb/total/depth; lane l holds the slot of the l-th
// nearest.  `sm` = 96 doubles of shared memory per warp (the ordered sum of the estimate).
__device__ __forceinline__ GatherOut gather_warp(const DGatherMap& M, d3 p, d3 dq, int k, int lane, double* sm)
{
    GatherOut out; out.rgb = mk3(0, 0, 0); out.total = 0; out.depth = 0; out.count = 0; out.best_slot = 0xFFFFFFFFu;
    uint32_t node, depth;
    bool found = pm_find_leaf(M, p, lane, node, depth);
    out.depth = depth;
    double best_d = CUDART_INF; uint32_t best_sl = 0xFFFFFFFFu;   // sorted ascending across lanes
    uint32_t total = 0;
    if (found) {
        const uint32_t off = __ldg(M.cand_off + node);
        total = __ldg(M.cand_off + node + 1) - off;
        double bat_d = CUDART_INF; uint32_t bat_sl = 0xFFFFFFFFu;   // pending batch, filled from lane 0 up
        int nb = 0;
        double kth_d = CUDART_INF; uint32_t kth_sl = 0xFFFFFFFFu;
        for (uint32_t base = 0; base < total; base += 32) {
            bool have = base + lane < total;
            double cd = CUDART_INF; uint32_t csl = 0xFFFFFFFFu;
            if (have) {
                csl = __ldg(M.cand_slot + off + base + lane);
                const double* pp = M.pos + 3 * (size_t)csl;
                cd = len2(mk3(__ldg(pp), __ldg(pp + 1), __ldg(pp + 2)) - p);
            }
            // only candidates that beat the current k-th entry can enter the result
            bool useful = have && kv_less(cd, csl, kth_d, kth_sl);
            uint32_t um = __ballot_sync(0xffffffffu, useful);
            int cnt = __popc(um);
            int taken = 0;
            while (taken < cnt) {
                int room = 32 - nb, put = cnt - taken < room ? cnt - taken : room;
                // lane j in [nb, nb+put) pulls the (taken + j - nb)-th useful candidate
                int want = lane - nb;
                int src = (want >= 0 && want < put) ? (int)__fns(um, 0, taken + want + 1) : lane;
                double sd = __shfl_sync(0xffffffffu, cd, src); uint32_t ssl = __shfl_sync(0xffffffffu, csl, src);
                if (want >= 0 && want < put) { bat_d = sd; bat_sl = ssl; }
                nb += put; taken += put;
                if (nb == 32) {
                    warp_sort32(bat_d, bat_sl, lane);
                    warp_merge32(best_d, best_sl, bat_d, bat_sl, lane);
                    kth_d = __shfl_sync(0xffffffffu, best_d, 31); kth_sl = __shfl_sync(0xffffffffu, best_sl, 31);
                    bat_d = CUDART_INF; bat_sl = 0xFFFFFFFFu; nb = 0;
                }
            }
        }
        if (nb > 0) {
            warp_sort32(bat_d, bat_sl, lane);
            warp_merge32(best_d, best_sl, bat_d, bat_sl, lane);
        }
    }
    // radiance estimate (raytracer.h:545-576): sum over the count = min(k, total) nearest in ascending distance order;
    // lanes 0..2 each add up one colour channel in that order
    int count = (int)total < k ? (int)total : k;
    d3 term = mk3(0, 0, 0);
    if (lane < count && best_sl != 0xFFFFFFFFu) {
        const double* dc = M.dircol + 6 * (size_t)best_sl;
        term = ld3(dc + 3) * dot3(ld3(dc), dq);
    }
    __syncwarp();
    sm[lane] = term.x; sm[32 + lane] = term.y; sm[64 + lane] = term.z;
    __syncwarp();
    double acc = 0;
    if (lane < 3) for (int i = 0; i < count; i++) acc += sm[32 * lane + i];
    d3 res = mk3(__shfl_sync(0xffffffffu, acc, 0), __shfl_sync(0xffffffffu, acc, 1), __shfl_sync(0xffffffffu, acc, 2));
    if (total > 0) {
        double md = __shfl_sync(0xffffffffu, best_d, count - 1);
        double den = GI_D_PI * md;
        res = mk3(res.x / den, res.y / den, res.z / den);
    }
    out.rgb = res; out.total = total; out.count = count; out.best_slot = best_sl;
    return out;
}

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k_gather(DGatherMap M, size_t n, const double* __restrict__ qpos, const double* __restrict__ qdir, int k,
                                                     double* __restrict__ rgb, uint32_t* __restrict__ knn, uint32_t* __restrict__ ncand,
                                                     const double* __restrict__ weight, double* __restrict__ accum, const uint32_t* __restrict__ accum_idx,
                                                     unsigned long long* work)
{
    __shared__ double s_sum[WARPS][96];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    size_t q = blockIdx.x * (size_t)WARPS + wib;
    if (q >= n) return;
    d3 p = ld3(qpos + 3 * q), dq = ld3(qdir + 3 * q);
    GatherOut g = gather_warp(M, p, dq, k, lane, s_sum[wib]);
    if (knn && lane < k) knn[q * (size_t)k + lane] = (lane < g.count && g.best_slot != 0xFFFFFFFFu) ? __ldg(M.pid + g.best_slot) : GI_NO_HIT;
    if (lane == 0) {
        if (work) { atomicAdd(work, (unsigned long long)g.depth); atomicAdd(work + 1, (unsigned long long)g.total); atomicAdd(work + 2, (unsigned long long)g.count); }
        if (rgb) st3(rgb + 3 * q, g.rgb);
        if (ncand) ncand[q] = g.total;
        if (accum) {   // render pipeline: L[path] += weight * caustic
            size_t a = accum_idx[q];
            accum[3 * a] += weight[3 * q] * g.rgb.x; accum[3 * a + 1] += weight[3 * q + 1] * g.rgb.y; accum[3 * a + 2] += weight[3 * q + 2] * g.rgb.z;
        }
    }
}

// ---- candidate lists: Node::get (photonMap.cpp:71-92) run once per leaf with that leaf's query box -------------------------
// One warp per node (interior nodes exit).  FILL = false: count candidates into cnt[node]; FILL = true: write the photon
// slots in the reference's DFS order (children 0..7, insertion order inside a leaf) starting at off[node].
#define GI_GATHER_STACK 1024   // a degenerate map (coincident photons) can be ~128 levels deep
template <int WARPS, bool FILL>
__global__ void __launch_bounds__(WARPS * 32) k_pm_cands(const DNode* __restrict__ nodes, uint32_t n_nodes, uint32_t* __restrict__ cnt, const uint32_t* __restrict__ off,
                                                       uint32_t* __restrict__ slots, uint32_t* __restrict__ overflow)
{
    __shared__ uint32_t s_stack[WARPS][GI_GATHER_STACK];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    uint32_t leaf = blockIdx.x * WARPS + wib;
    if (leaf >= n_nodes) return;
    DNode nd = load_node(nodes, leaf);
    if (nd.mask != 0) { if (!FILL && lane == 0) cnt[leaf] = 0; return; }
    double qmin[3], qmax[3];
    for (int a = 0; a < 3; a++) { qmin[a] = nd.bmin[a] - GI_D_EPSILON; qmax[a] = nd.bmax[a] + GI_D_EPSILON; }   // photonMap.cpp:119
    uint32_t total = 0;
    uint32_t wpos = FILL ? off[leaf] : 0;
    int sp = 0;
    if ((qmax[0] - qmin[0]) > 0) { if (lane == 0) s_stack[wib][0] = 0; sp = 1; }   // `if (bbox.dx() <= 0) return` (photonMap.cpp:73)
    __syncwarp();
    while (sp > 0) {
        uint32_t ni = s_stack[wib][sp - 1]; sp--;
        __syncwarp();
        DNode cur = load_node(nodes, ni);
        if (cur.mask == 0) {
            if (FILL) for (uint32_t b = lane; b < cur.prim_cnt; b += 32) slots[wpos + b] = cur.prim_off + b;
            wpos += cur.prim_cnt; total += cur.prim_cnt;
        } else {
            bool ov = false;
            if (lane < 8) {
                DNode ch = load_node(nodes, cur.child + lane);
                ov = (ch.bmin[0] <= qmax[0] && ch.bmax[0] >= qmin[0]) && (ch.bmin[1] <= qmax[1] && ch.bmax[1] >= qmin[1]) && (ch.bmin[2] <= qmax[2] && ch.bmax[2] >= qmin[2]);
            }
            uint32_t b = __ballot_sync(0xffffffffu, ov) & 0xffu;
            if (lane < 8 && ov) {   // push in reverse child order so that children pop in DFS order
                int rank = __popc(b >> (lane + 1));
                if (sp + rank < GI_GATHER_STACK) s_stack[wib][sp + rank] = cur.child + lane;
            }
            sp += __popc(b);
            if (sp > GI_GATHER_STACK) { sp = GI_GATHER_STACK; if (lane == 0) atomicExch(overflow, 1u); }
            __syncwarp();
        }
    }
    if (!FILL && lane == 0) cnt[leaf] = total;
}

// ---- K4: wavefront bounce = closest hit + shade + scatter (raytracer.h:167-276, 321-379, 481-506) -------------------------------------
// Queue entry (SoA): ray origin/dir, throughput T, Russian-roulette weight contrib, path id.  Per path: Halton index, PRNG
// key, radiance sum L.  Per hit (compacted "hit list"): what the shadow and gather kernels need.
struct DQueue { double* o; double* d; double* T; double* contrib; uint32_t* path; };
struct DHitList {
    double* p;        // hit point
    double* n;        // shading normal after the flip of secondaryRay
    double* wdirect;  // T * color          (weight of the direct term)
    double* wcaustic; // cont ? T * color : 0 (weight of the caustic term)
    double* refdir;   // outgoing direction (gather uses it)
    double* rough;    // roughness
    uint32_t* path;
};
struct DPathState { uint32_t* sample; uint64_t* key; double* L; };
struct DCounters { uint32_t n_next, n_hits; unsigned long long closest, shadow, gathers; };

template <bool FULL, bool IMPL>
__global__ void __launch_bounds__(GI_BLOCK, GI_MINB) k_bounce(DScene S, gi_render_params P, int depth, uint32_t n, DQueue in, DQueue out, DHitList H, DPathState PS, DCounters* C,
                                                     unsigned long long* work)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    bool active = i < n;
    bool is_hit = false, cont = false;
    d3 hp, hn, refDir, wdir, wcau, Tn, contrib;
    double rough = 1, offset = GI_D_SHADOW_BIAS;
    uint32_t path = 0, wn = 0, wp = 0;
    if (active) {
        path = in.path[i];
        DRay r = ray_as_stored(ld3(in.o + 3 * (size_t)i), ld3(in.d + 3 * (size_t)i));
        d3 T = ld3(in.T + 3 * (size_t)i);
        contrib = ld3(in.contrib + 3 * (size_t)i);
        uint64_t key = PS.key[path]; uint32_t sample = PS.sample[path];
        float sx = halton_sample(S, (uint32_t)(2 + 2 * depth), sample);     // raytracer.h:172-173
        float sy = halton_sample(S, (uint32_t)(3 + 2 * depth), sample);
        DHit h;
        trace_closest<FULL, IMPL>(S, r, P.seed, key, (uint64_t)depth, h, wn, wp);  // :190
        double* L = PS.L + 3 * (size_t)path;
        if (h.prim == GI_NO_HIT) {
            d3 a = T * ld3(S.ambient);                                       // :275
            L[0] += a.x; L[1] += a.y; L[2] += a.z;
        } else {
            is_hit = true;
            double tu, tv;
            hit_surface(S, r, h, FULL, hp, hn, tu, tv);
            const gi_material& m = S.mats[S.prim_mat[h.prim]];
            d3 color = tex_get(S, m.diffuse_tex, tu, tv);                    // :200
            rough = m.roughness;
            d3 f = mk3(1, 1, 1);
            refDir = secondary_ray(S, m, r, hn, tu, tv, sx, sy, color, f, contrib, offset, P.seed, key, (uint64_t)depth);   // :207
            double q = contrib.x < contrib.y ? contrib.y : contrib.x; q = q < contrib.z ? contrib.z : q;                   // compMax :263
            cont = depth <= P.min_depth || gi_rand(P.seed, key, (uint64_t)depth, SITE(SITE_RR, 0)) < q;                    // :265
            wdir = T * color;
            wcau = cont ? wdir : mk3(0, 0, 0);
            if (cont) {
                f = f * (depth <= P.min_depth ? 1.0 : (1.0 / q));            // :267
                d3 em = tex_get(S, m.emissive_tex, tu, tv);
                d3 e = T * em;                                               // emissive only on continued paths (:269 vs :272)
                L[0] += e.x; L[1] += e.y; L[2] += e.z;
                Tn = T * f;
                if (depth + 1 > P.max_depth) cont = false;                   // :169 — the next level would return 0
            }
        }
    }
    tally2(work, wn, wp);
    // queue compaction: warp ballot + prefix popcount, one atomic per warp and list
    const unsigned lane = threadIdx.x & 31u;
    unsigned mh = __ballot_sync(0xffffffffu, is_hit), mc = __ballot_sync(0xffffffffu, cont);
    uint32_t bh = 0, bc = 0;
    if (lane == 0) {
        if (mh) bh = atomicAdd(&C->n_hits, (uint32_t)__popc(mh));
        if (mc) bc = atomicAdd(&C->n_next, (uint32_t)__popc(mc));
    }
    bh = __shfl_sync(0xffffffffu, bh, 0); bc = __shfl_sync(0xffffffffu, bc, 0);
    if (is_hit) {
        size_t s = bh + __popc(mh & ((1u << lane) - 1u));
        st3(H.p + 3 * s, hp); st3(H.n + 3 * s, hn); st3(H.wdirect + 3 * s, wdir); st3(H.wcaustic + 3 * s, wcau); st3(H.refdir + 3 * s, refDir);
        H.rough[s] = rough; H.path[s] = path;
    }
    if (cont) {
        size_t s = bc + __popc(mc & ((1u << lane) - 1u));
        DRay nr = make_ray(hp + hn * offset, refDir);                        // Ray(minHit + offset*minNorm, refDir) :269
        st3(out.o + 3 * s, nr.o); st3(out.d + 3 * s, nr.d); st3(out.T + 3 * s, Tn); st3(out.contrib + 3 * s, contrib);
        out.path[s] = path;
    }
}

// ---- K3 in the pipeline: direct light with one shadow ray per light (raytracer.h:230-256) -----------------------------------------
template <bool FULL, bool IMPL>
__global__ void __launch_bounds__(GI_BLOCK, GI_MINB) k_direct(DScene S, gi_render_params P, int depth, uint32_t n, DHitList H, DPathState PS, unsigned long long* work)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t wn = 0, wp = 0;
    if (i >= n) { tally2(work, 0, 0); return; }
    uint32_t path = H.path[i];
    uint64_t key = PS.key[path];
    d3 p = ld3(H.p + 3 * (size_t)i), nn = ld3(H.n + 3 * (size_t)i);
    double rough = H.rough[i];
    d3 li = mk3(0, 0, 0);
    for (uint32_t l = 0; l < S.n_lights; l++) {
        const gi_light& light = S.lights[l];
        d3 sp = p + nn * GI_D_SHADOW_BIAS;
        d3 lightDir = light_point(light, gi_rand(P.seed, key, (uint64_t)depth, SITE(SITE_LIGHT_U, l)), gi_rand(P.seed, key, (uint64_t)depth, SITE(SITE_LIGHT_V, l))) - sp;   // :233
        double maxt = len2(lightDir);
        double hfrac = 1 / (GI_D_PI * len2(ld3(light.pos) - p));                                                             // :238
        DRay sr = make_ray(sp, lightDir);                                                                                     // :241
        if (trace_visible<FULL, IMPL>(S, sr, maxt, P.seed, key, (uint64_t)depth, l, wn, wp)) {                                      // :243
            double d = dot3(nn, normalize3(ld3(light.pos) - p));
            if (d < 0) d = 0;
            double lv = pow_like_libm(d, (1.0 / rough));                                                                      // :252
            li = (ld3(light.col) * lv) * hfrac;                                                                               // assignment, not += (:254)
        }
    }
    d3 w = ld3(H.wdirect + 3 * (size_t)i) * li;
    double* L = PS.L + 3 * (size_t)path;
    L[0] += w.x; L[1] += w.y; L[2] += w.z;
    tally2(work, wn, wp);
}

// ---- the tail: once few paths are left, one warp takes one path to its end (closest hit, shading, shadow rays, gather and
// the next bounce all inside the kernel), so a frame does not pay ~65 x 3 nearly empty launches with a host round trip
// each.  Arithmetic and the order in which terms are added to L[path] are those of k_bounce / k_direct / k_gather.
struct DTailCounters { unsigned long long closest, shadow, gathers, nodes_c, prims_c, nodes_s, prims_s, g_depth, g_cand, g_sel; unsigned int next; unsigned int pad; };

template <bool FULL, bool IMPL>
__global__ void __launch_bounds__(GI_WPB * 32) k_tail(DScene S, DGatherMap G, int have_map, gi_render_params P, int depth0, uint32_t n, DQueue in, DPathState PS, DTailCounters* TC)
{
    __shared__ uint32_t s_stack[GI_WPB][GI_STACK_MAX];
    __shared__ double s_sum[GI_WPB][96];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    uint32_t* stack = s_stack[wib];
  // persistent warps: paths are handed out one at a time, so a 60-bounce path does not hold back a wave of short ones
  for (;;) {
    uint32_t i = 0;
    if (lane == 0) i = atomicAdd(&TC->next, 1u);
    i = __shfl_sync(0xffffffffu, i, 0);
    if (i >= n) return;
    const uint32_t path = in.path[i];
    DRay r = ray_as_stored(ld3(in.o + 3 * (size_t)i), ld3(in.d + 3 * (size_t)i));
    d3 T = ld3(in.T + 3 * (size_t)i), contrib = ld3(in.contrib + 3 * (size_t)i);
    const uint64_t key = PS.key[path]; const uint32_t sample = PS.sample[path];
    d3 L = ld3(PS.L + 3 * (size_t)path);
    unsigned long long c_closest = 0, c_shadow = 0, c_gather = 0, g_depth = 0, g_cand = 0, g_sel = 0;
    uint32_t nc = 0, pc = 0, ns = 0, ps = 0;
    for (int depth = depth0; depth <= P.max_depth; depth++) {
        float sx = halton_sample(S, (uint32_t)(2 + 2 * depth), sample);
        float sy = halton_sample(S, (uint32_t)(3 + 2 * depth), sample);
        DHit h;
        trace_closest_warp<FULL, IMPL>(S, r, P.seed, key, (uint64_t)depth, h, stack, lane, nc, pc);
        c_closest++;
        if (h.prim == GI_NO_HIT) { L = L + T * ld3(S.ambient); break; }
        d3 hp, hn; double tu, tv;
        hit_surface(S, r, h, FULL, hp, hn, tu, tv);
        const gi_material& m = S.mats[S.prim_mat[h.prim]];
        d3 color = tex_get(S, m.diffuse_tex, tu, tv);
        double rough = m.roughness, offset = GI_D_SHADOW_BIAS;
        d3 f = mk3(1, 1, 1);
        d3 refDir = secondary_ray(S, m, r, hn, tu, tv, sx, sy, color, f, contrib, offset, P.seed, key, (uint64_t)depth);
        double q = contrib.x < contrib.y ? contrib.y : contrib.x; q = q < contrib.z ? contrib.z : q;
        bool cont = depth <= P.min_depth || gi_rand(P.seed, key, (uint64_t)depth, SITE(SITE_RR, 0)) < q;
        d3 wdir = T * color;
        d3 Tn = T;
        if (cont) {
            f = f * (depth <= P.min_depth ? 1.0 : (1.0 / q));
            L = L + T * tex_get(S, m.emissive_tex, tu, tv);
            Tn = T * f;
        }
        // direct light (k_direct)
        if (S.n_lights) {
            d3 li = mk3(0, 0, 0);
            for (uint32_t l = 0; l < S.n_lights; l++) {
                const gi_light& light = S.lights[l];
                d3 sp = hp + hn * GI_D_SHADOW_BIAS;
                d3 lightDir = light_point(light, gi_rand(P.seed, key, (uint64_t)depth, SITE(SITE_LIGHT_U, l)), gi_rand(P.seed, key, (uint64_t)depth, SITE(SITE_LIGHT_V, l))) - sp;
                double maxt = len2(lightDir);
                double hfrac = 1 / (GI_D_PI * len2(ld3(light.pos) - hp));
                DRay sr = make_ray(sp, lightDir);
                c_shadow++;
                if (trace_visible_warp<FULL, IMPL>(S, sr, maxt, P.seed, key, (uint64_t)depth, l, stack, lane, ns, ps)) {
                    double d = dot3(hn, normalize3(ld3(light.pos) - hp));
                    if (d < 0) d = 0;
                    li = (ld3(light.col) * pow_like_libm(d, (1.0 / rough))) * hfrac;
                }
            }
            L = L + wdir * li;
        }
        // caustic estimate (k_gather)
        if (depth <= P.caustic_max_depth) {
            c_gather++;
            if (have_map) {
                GatherOut g = gather_warp(G, hp, refDir, P.k_photons, lane, s_sum[wib]);
                g_depth += g.depth; g_cand += g.total; g_sel += (unsigned long long)g.count;
                d3 wc = cont ? wdir : mk3(0, 0, 0);
                L = L + wc * g.rgb;
            }
        }
        if (!cont) break;
        T = Tn;
        r = make_ray(hp + hn * offset, refDir);
    }
    if (lane == 0) {
        st3(PS.L + 3 * (size_t)path, L);
        atomicAdd(&TC->closest, c_closest); atomicAdd(&TC->shadow, c_shadow); atomicAdd(&TC->gathers, c_gather);
        atomicAdd(&TC->nodes_c, (unsigned long long)nc); atomicAdd(&TC->prims_c, (unsigned long long)pc); atomicAdd(&TC->nodes_s, (unsigned long long)ns); atomicAdd(&TC->prims_s, (unsigned long long)ps);
        atomicAdd(&TC->g_depth, g_depth); atomicAdd(&TC->g_cand, g_cand); atomicAdd(&TC->g_sel, g_sel);
    }
  }
}

// generate the camera paths of one chunk (path-linear range [c0, c0+n) of the tile's sample-major path space)
__global__ void k_generate(DScene S, DFrame F, int s0, uint64_t c0, uint32_t n, DQueue q, DPathState PS)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t lin = c0 + i;
    size_t npx = (size_t)F.tw * F.th;
    int s = s0 + (int)(lin / npx);
    size_t pix = lin % npx;
    int y = F.y0 + (int)(pix / F.tw), x = F.x0 + (int)(pix % F.tw);
    uint32_t idx;
    DRay r = camera_ray(S, F, x, y, s, idx);
    st3(q.o + 3 * (size_t)i, r.o); st3(q.d + 3 * (size_t)i, r.d);
    st3(q.T + 3 * (size_t)i, mk3(1, 1, 1)); st3(q.contrib + 3 * (size_t)i, mk3(1, 1, 1));
    q.path[i] = i;
    PS.sample[i] = idx;
    PS.key[i] = ((uint64_t)((uint64_t)y * (uint64_t)F.w + (uint64_t)x) << 24) | (uint64_t)s;
    PS.L[3 * (size_t)i] = 0; PS.L[3 * (size_t)i + 1] = 0; PS.L[3 * (size_t)i + 2] = 0;
}

// add the chunk's per-path radiance into the tile accumulator, samples in ascending order per pixel
__global__ void k_accumulate(uint64_t c0, uint32_t n, size_t npx, const double* L, double* accum)
{
    size_t pix = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (pix >= npx) return;
    uint64_t c1 = c0 + n;
    // paths of this pixel inside the chunk: lin = k*npx + pix
    uint64_t k0 = c0 > pix ? (c0 - pix + npx - 1) / npx : 0;
    double a0 = accum[3 * pix], a1 = accum[3 * pix + 1], a2 = accum[3 * pix + 2];
    for (uint64_t k = k0;; k++) {
        uint64_t lin = k * npx + pix;
        if (lin >= c1) break;
        size_t i = (size_t)(lin - c0);
        a0 += L[3 * i]; a1 += L[3 * i + 1]; a2 += L[3 * i + 2];
    }
    accum[3 * pix] = a0; accum[3 * pix + 1] = a1; accum[3 * pix + 2] = a2;
}

// ---- K8: resolve (raytracer.h:150-156, util.h:94-97, image.h:14-16) ------------------------------------------------------------------
__global__ void k_resolve(size_t n3, const double* accum, int spp, uint8_t* rgb8)
{
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n3) return;
    double c = accum[i] * (1.0 / spp);
    c = pow(c, 1.0 / 2.2);
    c = c < 0.0 ? 0.0 : (c > 1.0 ? 1.0 : c);
    if (!(c == c)) c = 0.0;
    rgb8[i] = (uint8_t)(int)(255 * c);
}

// ---- K5: photon emission and tracing (raytracer.h:582-715) ----------------------------------------------------------------------------
struct DPhotonOut { double* ph; uint8_t* stored; unsigned long long* tries; unsigned long long* traces; unsigned long long* work; };

template <bool FULL, bool IMPL>
__global__ void __launch_bounds__(GI_BLOCK) k_photon_trace(DScene S, int count, int max_depth, uint64_t seed, DPhotonOut O)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long my_tries = 0, my_traces = 0;
    uint32_t wn = 0, wp = 0;
    for (uint32_t li = 0; li < S.n_lights && i < count; li++) {
        const gi_light& l = S.lights[li];
        int tries = 0; bool stored = false;
        while (!stored && tries < 500) {                                                        // :602
            uint64_t path = PHOTON_PATH_BIT | ((uint64_t)li << 48) | (uint64_t)((uint64_t)i * 500u + (uint64_t)tries);
            float sx = halton_sample(S, 0, (uint32_t)(i * 500 + tries));                        // :604-605
            float sy = halton_sample(S, 1, (uint32_t)(i * 500 + tries));
            d3 pos = light_point_in_range(l, sx, sy);                                           // :612
            float du = (float)fmod(gi_rand(seed, path, 0, SITE(SITE_PH_DIR_U, 0)) + 5 * i, 1.0);
            float dv = (float)fmod(gi_rand(seed, path, 0, SITE(SITE_PH_DIR_V, 0)) + 13 * i, 1.0);
            d3 dir = sphere_cap_cos(normalize3(pos - ld3(l.pos)), du, dv, 2, l.angle);          // :613
            DRay r = make_ray(pos, dir);
            d3 col = ld3(l.col) * ((1.0 / count) * .5 * l.angle);                               // :618
            int depth = 0; bool term = false, isCaustic = false;
            DHit h;
            trace_closest<FULL, IMPL>(S, r, seed, path, 0, h, wn, wp); my_traces++;
            if (h.prim == GI_NO_HIT) { tries++; continue; }                                     // :626-630
            uint32_t cur = h.prim;
            d3 hit = mk3(0, 0, 0);
            while (depth < max_depth && !term) {                                                // :633
                double roughness = S.mats[S.prim_mat[cur]].roughness;
                if (roughness < 0.1) {
                    trace_closest<FULL, IMPL>(S, r, seed, path, (uint64_t)(depth + 1), h, wn, wp); my_traces++;   // :640
                    if (h.prim == GI_NO_HIT) { term = true; continue; }
                    cur = h.prim;
                    d3 norm; double tu, tv;
                    hit_surface(S, r, h, FULL, hit, norm, tu, tv);
                    const gi_material& m = S.mats[S.prim_mat[cur]];
                    roughness = m.roughness;
                    d3 f = mk3(0, 0, 0), contrib = mk3(0, 0, 0); double offset = GI_D_SHADOW_BIAS;
                    double su = fmod(gi_rand(seed, path, (uint64_t)(depth + 1), SITE(SITE_PH_SEC_U, 0)) + 5 * i, 1.0);
                    double sv = fmod(gi_rand(seed, path, (uint64_t)(depth + 1), SITE(SITE_PH_SEC_V, 0)) + 13 * i, 1.0);
                    d3 color = tex_get(S, m.diffuse_tex, tu, tv);
                    d3 refDir = secondary_ray(S, m, r, norm, tu, tv, su, sv, color, f, contrib, offset, seed, path, (uint64_t)(depth + 1));   // :656
                    col = col * f;                                                              // :677
                    r = make_ray(hit + norm * offset, refDir);                                  // :679-680
                    isCaustic = true;
                }
                if (depth > 0 && isCaustic && roughness >= 0.1) {                               // :685-692
                    size_t slot = (size_t)i * S.n_lights + li;
                    double* out = O.ph + 9 * slot;
                    st3(out, hit); st3(out + 3, r.d); st3(out + 6, col);
                    O.stored[slot] = 1;
                    term = true; stored = true;
                }
                depth++;
            }
            tries++;
        }
        my_tries += (unsigned long long)tries;
    }
    // warp-aggregated tallies
    for (int o = 16; o > 0; o >>= 1) { my_tries += __shfl_down_sync(0xffffffffu, my_tries, o); my_traces += __shfl_down_sync(0xffffffffu, my_traces, o); }
    if ((threadIdx.x & 31) == 0) { atomicAdd(O.tries, my_tries); atomicAdd(O.traces, my_traces); }
    tally2(O.work, wn, wp);
}

// stable compaction of the stored photons into (i, light) order: flags -> exclusive scan (k_scan_*) -> scatter
__global__ void k_flags_to_u32(const uint8_t* flags, uint32_t n, uint32_t* out)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = flags[i];
}
__global__ void k_photon_compact(const double* src, const uint8_t* flags, const uint32_t* offs, uint32_t n, double* dst)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !flags[i]) return;
    for (int k = 0; k < 9; k++) dst[9 * (size_t)offs[i] + k] = src[9 * (size_t)i + k];
}

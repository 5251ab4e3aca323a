// gi_kernels.cuh — the sm_100a kernels of the hot path (K1..K8 of SURVEY §2.2).  One thread per ray / path / photon,
// one warp per gather query.  All kernels are grid-stride free: the host sizes the grid to the queue length.
#pragma once
#include <math_constants.h>

#include "gi_device.cuh"

#ifndef GI_MINB
#define GI_MINB 8   // resident blocks of 128 threads per SM asked of ptxas for the thread-per-ray traversal kernels: 64 registers.
#endif              // round 1 (C2 bounce + direct ms): 1-3 blocks (136-142 regs) 23.1, 4 (128) 19.4, 5 (96) 18.8, 6 (80) 17.9, 8 (64) 19.1; round 2, after the
                    // by-value cone test and the later T / contrib loads took pressure off the walk: 8 beats 6 on every config (C2 frame 24.7 vs 25.2 ms,
                    // glass 49.4 vs 51.6, foliage 488 vs 512, atrium 461 vs 472; profiles/r02/ab_t7.txt)
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

// ---- atmosphere, batch forms: Octree::atmosphereDensity at points; Octree::atmosphereBounds (+ RayTracer::raymarch) on rays --------------
__global__ void k_fog_density(DScene S, size_t n, const double* __restrict__ pos, double* __restrict__ dens, double* __restrict__ col)
{
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    d3 c = mk3(0, 0, 0);
    dens[i] = atmosphere_density(S, ld3(pos + 3 * i), c);
    st3(col + 3 * i, c);
}
__global__ void k_raymarch(DScene S, size_t n, const double* __restrict__ org, const double* __restrict__ dir, const double* __restrict__ tmax, uint64_t seed, int march, uint8_t* __restrict__ hit,
                           double* __restrict__ t01, double* __restrict__ pos, double* __restrict__ col)
{
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    DRay r = ray_as_stored(ld3(org + 3 * i), ld3(dir + 3 * i));
    double a = 0, b = tmax[i];
    d3 h = mk3(0, 0, 0), c = mk3(0, 0, 0);
    bool ok = atmosphere_bounds(S, r, a, b);
    t01[2 * i] = a; t01[2 * i + 1] = b;
    if (march) { ok = ok && raymarch(S, r, h, c, a, b, seed, (uint64_t)i, 0, SITE_FOG_RAD, 0); if (!ok) { h = mk3(0, 0, 0); c = mk3(0, 0, 0); } }
    hit[i] = ok ? 1 : 0;
    st3(pos + 3 * i, h); st3(col + 3 * i, c);
}

// ---- materials, batch form: Material::diffuse->get(uv), emissive->get(uv), Material::getAlpha(uv) (material.h:18-26, 39-45, 63-81, 90-93)
// for the material of primitive prim[i] at uv[i] — what radiance() reads at raytracer.h:200 and the alpha test at :455
__global__ void k_material_eval(DScene S, size_t n, const uint32_t* __restrict__ prim, const double* __restrict__ uv, double* __restrict__ dif, double* __restrict__ em, double* __restrict__ alpha)
{
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t p = prim[i];
    if (p >= S.n_prims) { st3(dif + 3 * i, mk3(0, 0, 0)); st3(em + 3 * i, mk3(0, 0, 0)); alpha[i] = 0; return; }
    const gi_material& m = S.mats[S.prim_mat[p]];
    const double u = uv[2 * i], v = uv[2 * i + 1];
    st3(dif + 3 * i, tex_get(S, m.diffuse_tex, u, v));
    st3(em + 3 * i, tex_get(S, m.emissive_tex, u, v));
    alpha[i] = m.opacity * tex_alpha(S, m.diffuse_tex, u, v);
}

// ---- scene upload: the leaf records are expanded ON THE DEVICE from the primitive arrays and the leaf index list — the host sends
// 9 doubles per PRIMITIVE and 4 bytes per leaf reference instead of an 80-byte record per reference (the atrium stand-in: 72 MB instead
// of 1.06 GB).  A triangle's record holds v0, v1 - v0, v2 - v0: the edges the reference recomputes per test (entities.h:447-448),
// subtracted here once in the same fp64 arithmetic.
__global__ void k_build_leafrefs(uint32_t n_refs, const uint32_t* __restrict__ leaf_prims, const double* __restrict__ prim_geom, const uint8_t* __restrict__ prim_type,
                                 const uint32_t* __restrict__ prim_flags, DLeafRef* __restrict__ refs)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_refs) return;
    const uint32_t p = leaf_prims[i];
    const double* g = prim_geom + 9 * (size_t)p;
    DLeafRef r;
#pragma unroll
    for (int k = 0; k < 9; k++) r.g[k] = g[k];
    if (prim_type[p] == GI_PRIM_TRIANGLE) {
#pragma unroll
        for (int k = 0; k < 3; k++) { r.g[3 + k] = r.g[3 + k] - r.g[k]; r.g[6 + k] = r.g[6 + k] - r.g[k]; }
    }
    r.prim = p; r.flags = prim_flags[p];
    refs[i] = r;
}

// ---- K1: camera rays (raytracer.h:74-78, 112-129) ------------------------------------------------------------------------------
// tile = local pixel space tw x th.  A rectangle maps local row ly to image row y0 + ly; the tile split's row plan (blocks of `rb`
// rows every `rstride` rows, gi_render_rows) maps it to y0 + (ly / rb) * rstride + ly % rb.  rb = 0 means a plain rectangle.
struct DFrame { int w, h, x0, y0, tw, th, rb, rstride; double halfW, halfH; d3 center, right, up, pos; DHEnum he; };
__device__ __forceinline__ int frame_row(const DFrame& F, int ly) { return F.rb ? F.y0 + (ly / F.rb) * F.rstride + ly % F.rb : F.y0 + ly; }

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
    int y = frame_row(F, (int)(p / F.tw)), x = F.x0 + (int)(p % F.tw);
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
    double* pos4;               // [kept][4]: x, y, z, pad — 32-byte records, two aligned 16-byte loads per candidate
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
    // one block: exclusive scan of the block sums, GI_SCAN_BLOCK at a time with a running carry
    __shared__ uint32_t sh[GI_SCAN_BLOCK];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < nb; base += GI_SCAN_BLOCK) {
        uint32_t i = base + threadIdx.x;
        uint32_t v = i < nb ? block_sums[i] : 0u;
        sh[threadIdx.x] = v;
        __syncthreads();
        for (int off = 1; off < GI_SCAN_BLOCK; off <<= 1) {
            uint32_t t = threadIdx.x >= off ? sh[threadIdx.x - off] : 0u;
            __syncthreads();
            sh[threadIdx.x] += t;
            __syncthreads();
        }
        uint32_t c = carry;
        if (i < nb) block_sums[i] = c + sh[threadIdx.x] - v;
        __syncthreads();
        if (threadIdx.x == GI_SCAN_BLOCK - 1) carry = c + sh[threadIdx.x];
        __syncthreads();
    }
    if (threadIdx.x == 0 && total) *total = carry;
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
    for (int k = 0; k < 3; k++) M.pos4[4 * (size_t)s + k] = src[k];
    M.pos4[4 * (size_t)s + 3] = 0.0;
    for (int k = 0; k < 6; k++) M.dircol[6 * (size_t)s + k] = src[3 + k];
}

// ---- K7: gather — RayTracer::samplePhotons (raytracer.h:532-579) over PhotonMap::getInRange (photonMap.cpp:50-92,115-134) ----
// The candidate set of a query is a function of the LEAF that contains it: every photon of every leaf whose closed box
// touches (leaf box +- EPSILON).  So the map carries, per leaf, the precomputed candidate list — built once on the device by
// running Node::get for every leaf (k_pm_cands, k_pm_cand_order).  Two forms of the query follow: the warp-cooperative one
// right below (one query per warp: tail kernel and long lists) and the sorted wavefront form (one thread per query in leaf
// order: k_gather_locate / k_gather_sorted / k_gather_heavy, further down).
struct DGatherMap {
    const DNode* nodes; const double* pos4; const double* dircol; const uint32_t* pid; uint32_t n_nodes;
    const uint32_t* cand_off;    // [n_nodes + 1] start of each node's candidate list (only leaves have entries)
    // the candidate lists, concatenated per leaf, each ordered by distance from the leaf centre: one 32-byte record per
    // candidate = position x, y, z | photon slot (u32) | squared distance from the centre of the leaf, rounded down (f32; 0 =
    // list not ordered).  A list is read front to back with two aligned 16-byte loads per candidate and no indirection.
    const double* cand_rec;
};
struct Cand { d3 pos; uint32_t slot; float key; };
__device__ __forceinline__ Cand load_cand(const DGatherMap& M, uint32_t idx)
{
    const double2* r = reinterpret_cast<const double2*>(M.cand_rec) + 2 * (size_t)idx;
    const double2 a = __ldg(r), b = __ldg(r + 1);
    Cand c; c.pos = mk3(a.x, a.y, b.x); c.slot = (uint32_t)__double2loint(b.y); c.key = __int_as_float(__double2hiint(b.y));
    return c;
}
// ordering of candidates: (distance^2, photon slot).  Exact distance ties between different photons are unordered in
// the reference (std::partial_sort is unstable); the slot makes them deterministic here.
__device__ __forceinline__ bool kv_less(double a, uint32_t ai, double b, uint32_t bi) { return a < b || (a == b && ai < bi); }

// bitonic compare-exchange across lanes on (key, slot) pairs.  Squared distances are >= +0 (or +inf padding), so their
// bit patterns order like unsigned integers: the three-word comparison (hi, lo, slot) needs no fp64 pipe.
__device__ __forceinline__ void cmpx(double& d, uint32_t& sl, int lane, int j, bool up)
{
    const uint32_t hi = (uint32_t)__double2hiint(d), lo = (uint32_t)__double2loint(d);
    const uint32_t ohi = __shfl_xor_sync(0xffffffffu, hi, j), olo = __shfl_xor_sync(0xffffffffu, lo, j), osl = __shfl_xor_sync(0xffffffffu, sl, j);
    const bool mine_less = hi < ohi || (hi == ohi && (lo < olo || (lo == olo && sl < osl)));
    const bool keep_min = ((lane & j) == 0) == up;
    if (mine_less != keep_min) { d = __hiloint2double((int)ohi, (int)olo); sl = osl; }
}
__device__ __forceinline__ void warp_sort32(double& d, uint32_t& sl, int lane)
{
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1)
#pragma unroll
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

// getBounds (photonMap.cpp:115-134), warp form: the leaf whose half-open box contains p; false when p is in no child
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

// getBounds, thread form: one query per thread.  The map's child boxes ARE the partition formula of the parent box
// (k_pm_split writes child_box()), so the walk keeps the current box in registers and derives the child that contains p
// with the same doubles: per axis the lower half is [min, mid), the upper half [min + h, mid + h) (min + h == mid bit for
// bit), and child 7 is [mid, max).  Children are disjoint, so "first containing child in 0..7 order" is "the containing
// child".  One 16-byte topology load per level.
__device__ __forceinline__ bool pm_find_leaf_thread(const DGatherMap& M, d3 p, uint32_t& node, uint32_t& depth)
{
    node = 0; depth = 0;
    if (M.n_nodes == 0 || !(p.x < CUDART_INF)) return false;   // +inf marks an unused query slot (k_tail_t)
    DNode root = load_node(M.nodes, 0);
    double lo[3] = { root.bmin[0], root.bmin[1], root.bmin[2] }, hi[3] = { root.bmax[0], root.bmax[1], root.bmax[2] };
    const double pp[3] = { p.x, p.y, p.z };
    uint32_t child = root.child, mask = root.mask;
    while (mask != 0) {
        depth++;
        double m[3], mh[3];
        bool low[3], up[3], up7[3];
#pragma unroll
        for (int a = 0; a < 3; a++) {
            double h = .5 * (hi[a] - lo[a]);
            m[a] = lo[a] + .5 * (hi[a] - lo[a]);
            mh[a] = m[a] + h;
            low[a] = pp[a] >= lo[a] && pp[a] < m[a];
            up[a] = pp[a] >= lo[a] + h && pp[a] < mh[a];
            up7[a] = pp[a] >= m[a] && pp[a] < hi[a];
        }
        int c;
        const bool all_up = up[0] && up[1] && up[2];
        if ((low[0] || up[0]) && (low[1] || up[1]) && (low[2] || up[2]) && !all_up) c = (up[0] ? 1 : 0) | (up[2] ? 2 : 0) | (up[1] ? 4 : 0);   // bit0 = +x, bit1 = +z, bit2 = +y
        else if (up7[0] && up7[1] && up7[2]) c = 7;
        else return false;   // box(-inf,-inf): nothing overlaps (photonMap.cpp:132)
#pragma unroll
        for (int a = 0; a < 3; a++) {
            const bool u = a == 0 ? (c & 1) : (a == 1 ? (c & 4) : (c & 2));
            if (c == 7) lo[a] = m[a];
            else if (u) { lo[a] = m[a]; hi[a] = mh[a]; }   // lo + h == mid
            else hi[a] = m[a];
        }
        node = child + c;
        uint4 top = __ldg(reinterpret_cast<const uint4*>(M.nodes + node) + 3);
        child = top.x; mask = top.w;
    }
    return true;
}

struct GatherOut { d3 rgb; uint32_t total, depth; int count; uint32_t best_slot; };

__device__ __forceinline__ void warp_order32(double& d, uint32_t& sl, int count, int lane);

// streaming top-k (any candidate count, total order (distance, slot)): candidates 32 at a time; those that beat the current
// 32nd entry are appended to a batch in shared memory (ballot + prefix popcount); a full batch is ordered and merged into the
// warp-wide sorted list (one entry per lane).  `delta` >= 0: half diagonal of the query's leaf — the list is ordered by
// distance from the leaf centre and the scan stops once no later candidate can beat the current 32nd (see k_gather_sorted);
// delta < 0: scan everything.  `sm` = 48 doubles of shared memory of this warp.
__device__ __noinline__ void gather_topk_stream(const DGatherMap& M, uint32_t off, uint32_t total, d3 p, int lane, double delta, double* sm, double& best_d, uint32_t& best_sl)
{
    float bound = CUDART_INF_F;
    uint32_t* sms = reinterpret_cast<uint32_t*>(sm + 32);
    best_d = CUDART_INF; best_sl = 0xFFFFFFFFu;
    int nb = 0;   // entries in the pending batch
    double kth_d = CUDART_INF; uint32_t kth_sl = 0xFFFFFFFFu;
    const uint32_t lt = (1u << lane) - 1u;
    __syncwarp();
    for (uint32_t base = 0; base < total; base += 32) {
        const bool have = base + lane < total;
        double cd = CUDART_INF; uint32_t csl = 0xFFFFFFFFu; float key = 0.f;
        if (have) { const Cand c = load_cand(M, off + base + lane); csl = c.slot; cd = len2(c.pos - p); key = c.key; }
        if (delta >= 0 && __shfl_sync(0xffffffffu, key, 0) > bound) break;   // lane 0 holds the smallest key of the step
        const bool useful = have && kv_less(cd, csl, kth_d, kth_sl);
        const uint32_t um = __ballot_sync(0xffffffffu, useful);
        if (!um) continue;
        const int pos = nb + __popc(um & lt), cnt = __popc(um);
        if (useful && pos < 32) { sm[pos] = cd; sms[pos] = csl; }
        if (nb + cnt >= 32) {
            __syncwarp();
            double bd = sm[lane]; uint32_t bs = sms[lane];
            __syncwarp();
            warp_order32(bd, bs, 32, lane);
            warp_merge32(best_d, best_sl, bd, bs, lane);
            kth_d = __shfl_sync(0xffffffffu, best_d, 31); kth_sl = __shfl_sync(0xffffffffu, best_sl, 31);
            if (delta >= 0 && kth_d < CUDART_INF) { const double b = sqrt(kth_d) + delta; bound = __double2float_ru(b * b * (1.0 + 1e-9)); }
            if (useful && pos >= 32) { sm[pos - 32] = cd; sms[pos - 32] = csl; }   // the overflow opens the next batch
            nb = nb + cnt - 32;
        } else nb += cnt;
    }
    if (nb > 0) {
        __syncwarp();
        double bd = lane < nb ? sm[lane] : CUDART_INF; uint32_t bs = lane < nb ? sms[lane] : 0xFFFFFFFFu;
        __syncwarp();
        warp_order32(bd, bs, nb, lane);
        warp_merge32(best_d, best_sl, bd, bs, lane);
    }
}

// ---- order: sort <= 32 (distance, slot) pairs held one per lane (lanes >= count hold +inf) -------------------------------------------
// The bitonic network runs on ONE 32-bit word per lane: the distance bits, rebased to the smallest binade present and
// shifted down to 26 bits (monotone in the distance), with the lane number in the low 5 bits — SHFL + min/max per stage.
// The pairs are then fetched from their source lanes.  Two distances that fall into the same 26-bit cell come out in lane
// order instead of (distance, slot) order; a neighbour check detects that (an unsorted run always has an inverted adjacent
// pair) and the exact three-word network below redoes the sort.
__device__ __forceinline__ void warp_order32(double& d, uint32_t& sl, int count, int lane)
{
    const bool valid = lane < count;
    const uint32_t hw = (uint32_t)__double2hiint(d);
    const uint32_t hmn = __reduce_min_sync(0xffffffffu, valid ? hw : 0xFFFFFFFFu), hmx = __reduce_max_sync(0xffffffffu, valid ? hw : 0u);
    const unsigned long long kbase = (unsigned long long)hmn << 32, range = ((unsigned long long)(hmx - hmn) + 1ull) << 32;
    int sh = 38 - __clzll((long long)range); sh = sh < 0 ? 0 : sh;   // (range >> sh) < 2^26
    uint32_t key = valid ? ((uint32_t)(((unsigned long long)__double_as_longlong(d) - kbase) >> sh) << 5) | (uint32_t)lane : 0xFFFFFFFFu;
#pragma unroll
    for (int k2 = 2; k2 <= 32; k2 <<= 1)
#pragma unroll
        for (int j = k2 >> 1; j > 0; j >>= 1) {
            const uint32_t o = __shfl_xor_sync(0xffffffffu, key, j);
            const bool keep_min = ((lane & j) == 0) == ((lane & k2) == 0 || k2 == 32);
            key = keep_min ? min(key, o) : max(key, o);
        }
    const int src = (int)(key & 31u);
    double sd = __shfl_sync(0xffffffffu, d, src); uint32_t ssl = __shfl_sync(0xffffffffu, sl, src);
    if (!valid) { sd = CUDART_INF; ssl = 0xFFFFFFFFu; }
    const double nd = __shfl_down_sync(0xffffffffu, sd, 1); const uint32_t nsl = __shfl_down_sync(0xffffffffu, ssl, 1);
    const bool inverted = (lane + 1 < count) & !kv_less(sd, ssl, nd, nsl);
    d = sd; sl = ssl;
    if (__any_sync(0xffffffffu, inverted)) warp_sort32(d, sl, lane);
}

// One query at a known leaf, executed by a full warp (tail kernel, long candidate lists): the list is streamed by
// gather_topk_stream, then the estimate (raytracer.h:545-576) is summed over the count = min(k, total) nearest in ascending
// distance order.  Every lane returns the same rgb/total; lane l holds the slot of the l-th nearest.  `sm` = 96 doubles of
// shared memory per warp (batch scratch, then the ordered sum).
__device__ __forceinline__ GatherOut gather_at_leaf(const DGatherMap& M, bool found, uint32_t node, d3 p, d3 dq, int k, int lane, double* sm, double delta = -1.0)
{
    GatherOut out; out.rgb = mk3(0, 0, 0); out.total = 0; out.depth = 0; out.count = 0; out.best_slot = 0xFFFFFFFFu;
    double best_d = CUDART_INF; uint32_t best_sl = 0xFFFFFFFFu;   // sorted ascending across lanes
    uint32_t total = 0, off = 0;
    if (found) { off = __ldg(M.cand_off + node); total = __ldg(M.cand_off + node + 1) - off; }
    const int count = (int)total < k ? (int)total : k;
    double acc = 0;
    if (total > 0) {
        gather_topk_stream(M, off, total, p, lane, delta, sm, best_d, best_sl);
        d3 term = mk3(0, 0, 0);
        if (lane < count && best_sl != 0xFFFFFFFFu) {
            const double2* dc = reinterpret_cast<const double2*>(M.dircol + 6 * (size_t)best_sl);
            double2 a = __ldg(dc), b = __ldg(dc + 1), c = __ldg(dc + 2);
            term = mk3(b.y, c.x, c.y) * dot3(mk3(a.x, a.y, b.x), dq);
        }
        __syncwarp();
        sm[lane] = term.x; sm[32 + lane] = term.y; sm[64 + lane] = term.z;
        __syncwarp();
        // lanes 0..2 each add up one colour channel; terms past `count` are +0.0 and leave the sum unchanged, so the loop is a
        // fixed 32 steps of 16-byte loads; the division by pi*r^2 is done once, by the lane that owns the channel
        const double2* s2 = reinterpret_cast<const double2*>(sm + 32 * (lane < 3 ? lane : 0));
#pragma unroll
        for (int i = 0; i < 16; i++) { double2 v = s2[i]; acc += v.x; acc += v.y; }
        const double md = __shfl_sync(0xffffffffu, best_d, count - 1);
        acc = acc / (GI_D_PI * md);
    }
    out.rgb = mk3(__shfl_sync(0xffffffffu, acc, 0), __shfl_sync(0xffffffffu, acc, 1), __shfl_sync(0xffffffffu, acc, 2));
    out.total = total; out.count = count; out.best_slot = best_sl;
    return out;
}

// one query per warp, leaf located by the warp (tail kernel)
__device__ __forceinline__ GatherOut gather_warp(const DGatherMap& M, d3 p, d3 dq, int k, int lane, double* sm)
{
    uint32_t node, depth;
    bool found = pm_find_leaf(M, p, lane, node, depth);
    GatherOut g = gather_at_leaf(M, found, node, p, dq, k, lane, sm);
    g.depth = depth;
    return g;
}

// ---- K7, sorted form: queries ordered by leaf, one THREAD per query ------------------------------------------------------------
// The warp-cooperative form above spends its time waiting: every step of a query (shuffle network, REDUX, ordered sum) hangs
// on the one before it, and a warp carries a single such chain (ncu: issue slots 35-55 % busy, stalls spread evenly over the
// whole instruction stream).  Queries are independent, so the wavefront form turns the problem around:
//   k_gather_locate   one thread per query walks to its leaf (thread form of getBounds) and counts the leaf in a histogram;
//   scan + k_bin_scatter order the queries by leaf (counting sort, permutation only);
//   k_gather_sorted   one thread per query, in leaf order.  The 32 lanes of a warp now sit in the same leaf (or two): they
//                     run through the SAME candidate list in lockstep — the slot and position loads are warp-wide broadcasts
//                     served by L1 — each against its own query point, with 32 independent dependency chains per warp.
//                     The k nearest so far are a max-heap in a per-thread column of shared memory; the scan of the
//                     (centre-ordered) list stops early; a heap sort orders the k for the sum (details at the kernel).
//   k_gather_heavy    lists longer than GI_GS_MAX_CANDS: persistent warps stream them, 32 candidates per step.
#ifndef GI_GS_MAX_CANDS
#define GI_GS_MAX_CANDS 256u
#endif
// GI_GS_MAX_CANDS: longer candidate lists are streamed by a whole warp (k_gather_heavy) instead of by one thread.  The locate kernel
// already knows the leaf and so the length of the list: it queues those queries, and the host starts k_gather_heavy — persistent
// warps, latency-bound, 16 warps per SM — on an auxiliary stream right away, BESIDE the counting sort and k_gather_sorted instead of
// after them.  (Until round 2 k_gather_sorted queued them; a rule that kept a long list with its thread when enough lanes of the warp
// shared it was measured and never paid: profiles/r02/ab_t20_heavy_thresholds.txt.)
__global__ void k_gather_locate(DGatherMap M, uint32_t n, const double* __restrict__ qpos, uint32_t* __restrict__ qnode, uint32_t* __restrict__ hist, unsigned long long* work,
                                uint32_t* __restrict__ heavy, uint32_t* __restrict__ heavy_n)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t depth = 0;
    const unsigned live = __ballot_sync(0xffffffffu, i < n);
    if (i < n) {
        uint32_t node;
        bool found = pm_find_leaf_thread(M, ld3(qpos + 3 * (size_t)i), node, depth);
        uint32_t key = found ? node : M.n_nodes;   // queries outside every leaf go last
        qnode[i] = key;
        if (hist) {   // one atomic per distinct leaf of the warp (neighbouring queries mostly share it), like k_bin_keys
            const unsigned peers = __match_any_sync(live, key);
            if ((threadIdx.x & 31u) == (unsigned)(__ffs(peers) - 1)) atomicAdd(hist + key, (uint32_t)__popc(peers));
        }
        // long lists: queue the query for k_gather_heavy (one warp per query, taken from a shared counter)
        const bool hard = found && __ldg(M.cand_off + node + 1) - __ldg(M.cand_off + node) > GI_GS_MAX_CANDS;
        const unsigned hm = __ballot_sync(live, hard);
        if (hm) {
            const unsigned lane = threadIdx.x & 31u;
            const int leader = __ffs(hm) - 1;
            uint32_t base = 0;
            if (lane == (unsigned)leader) base = atomicAdd(heavy_n, (uint32_t)__popc(hm));
            base = __shfl_sync(live, base, leader);
            if (hard) heavy[base + __popc(hm & ((1u << lane) - 1u))] = i;
        }
    }
    if (work) {
        unsigned long long wd = depth;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) wd += __shfl_down_sync(0xffffffffu, wd, o);
        if ((threadIdx.x & 31) == 0 && wd) atomicAdd(work, wd);
    }
}

#define GI_GS_BLOCK 64   // 64 columns x 32 rows x 12 B = 24 KB of shared memory per block
// (measured and dropped in round 2: the k nearest as an ascending sorted column with insertion from the end instead of the max-heap +
//  final heap sort — no sort afterwards, but 27 % slower: 1.18 vs 0.93 ms for the 776 666 C2 queries, profiles/r02/ab_gather.txt)

// max-heap of (distance^2, slot) pairs in one thread's column of shared memory ([row][thread]: conflict-free for any mix of rows);
// sift(n, i, d, sl) puts (d, sl) into the hole at row i of a heap of n rows and sifts it down.
// Measured against this form in round 2 and dropped, all with identical index sets (profiles/r02/ab_t13_bin_atomics_heap.txt,
// ab_t15_heap_regtop.txt; isolated C2 gather, 776 666 queries, this form 0.88 ms):
//   * sibling pairs as one 16-byte word (one LDS.128 per level, but single-row stores become 2-way bank conflicts): 0.95 ms;
//   * bottom-up extraction in the final sort (one comparison per level, but always the full height): 0.90 ms;
//   * rows 0..2 in registers (no shared-memory access on the first level, 29 rows -> a tenth block per SM; but every access by a
//     run-time row number turns into a branch ladder, 96 registers): 1.16 ms;
//   * an ascending sorted column with insertion from the end instead of heap + heap sort (ab_gather.txt): 1.18 ms.
#define GI_GS_ROWS 32
struct GHeap {
    double (*sd)[GI_GS_BLOCK]; uint32_t (*ss)[GI_GS_BLOCK]; int t;
    __device__ __forceinline__ void get(int r, double& d, uint32_t& s) const { d = sd[r][t]; s = ss[r][t]; }
    __device__ __forceinline__ void set(int r, double d, uint32_t s) { sd[r][t] = d; ss[r][t] = s; }
    __device__ __forceinline__ void sift(int n, int i, double d, uint32_t sl)
    {
        for (;;) {
            int c = 2 * i + 1;
            if (c >= n) break;
            double cd = sd[c][t]; uint32_t cs = ss[c][t];
            if (c + 1 < n) { const double e = sd[c + 1][t]; const uint32_t es = ss[c + 1][t]; if (kv_less(cd, cs, e, es)) { c++; cd = e; cs = es; } }
            if (!kv_less(d, sl, cd, cs)) break;
            sd[i][t] = cd; ss[i][t] = cs; i = c;
        }
        sd[i][t] = d; ss[i][t] = sl;
    }
};
#define GI_GHEAP_DECL(h) __shared__ double s_hd[GI_GS_ROWS][GI_GS_BLOCK]; __shared__ uint32_t s_hs[GI_GS_ROWS][GI_GS_BLOCK]; GHeap h; h.sd = s_hd; h.ss = s_hs; h.t = threadIdx.x

// Each leaf's candidate list is stored in ascending distance from the CENTRE of the leaf box (k_pm_cand_order), with that
// squared distance (rounded down) beside it.  A query point q lies inside the leaf box, so for a candidate c
//     |q - c| >= |centre - c| - delta,     delta = half diagonal of the leaf box  >= |q - centre|.
// The thread keeps the k nearest so far as a MAX-HEAP in its column of shared memory ([row][thread]: conflict-free for any mix
// of rows); tau = the root = the k-th distance.  A candidate below tau replaces the root and sifts down (<= 5 levels, whatever
// the arrival order — an insertion-sorted list degenerates to k shifts per candidate when distances keep falling along the
// list).  Once |centre - c| - delta > sqrt(tau) no later candidate of the (ordered) list can enter, and the scan stops.  At the
// end an in-place heap sort leaves the column in ascending (distance^2, slot) order for the sum.  The bound is inflated by
// 1e-9 relative against rounding.
__global__ void __launch_bounds__(GI_GS_BLOCK) k_gather_sorted(DGatherMap M, uint32_t n, const uint32_t* __restrict__ perm, const uint32_t* __restrict__ qnode,
                                                                 const double* __restrict__ qpos, const double* __restrict__ qdir, int k, double* __restrict__ rgb,
                                                                 uint32_t* __restrict__ knn, uint32_t* __restrict__ ncand, const double* __restrict__ weight,
                                                                 double* __restrict__ accum, const uint32_t* __restrict__ accum_idx, unsigned long long* work)
{
    GI_GHEAP_DECL(H);                            // per-thread heap, then sorted list: row = rank, column = thread
    const int t = threadIdx.x, lane = t & 31;
    const uint32_t i = blockIdx.x * GI_GS_BLOCK + t;
    const bool have = i < n;
    uint32_t q = 0, total = 0, off = 0, node = 0xFFFFFFFFu;
    d3 p = mk3(0, 0, 0), dq = mk3(0, 0, 0);
    double delta = 0;
    if (have) {
        q = perm ? perm[i] : i;
        node = qnode[q];
        p = ld3(qpos + 3 * (size_t)q); dq = ld3(qdir + 3 * (size_t)q);
        if (node < M.n_nodes) {
            off = __ldg(M.cand_off + node); total = __ldg(M.cand_off + node + 1) - off;
            const DNode nd = load_node(M.nodes, node);
            const double ex = nd.bmax[0] - nd.bmin[0], ey = nd.bmax[1] - nd.bmin[1], ez = nd.bmax[2] - nd.bmin[2];
            delta = .5 * sqrt(ex * ex + ey * ey + ez * ez) * (1.0 + 1e-9);
        }
    }
    const int count = (int)total < k ? (int)total : k;
    const float delta_f = __double2float_ru(delta);
    int m = 0;                       // rows in use
    double tau = CUDART_INF; uint32_t tau_sl = 0xFFFFFFFFu;   // the k-th (distance^2, slot) once m == k
    float bound = CUDART_INF_F;      // scan stops at the first key above it
    // long lists (a large leaf next to a dense region touches thousands of small ones) are streamed by a whole warp, 32 candidates
    // per step, instead of by one thread: k_gather_locate has queued those queries for k_gather_heavy, this kernel leaves them alone
    const bool hard = total > GI_GS_MAX_CANDS;
    // the list is read four records ahead of their use (eight independent 16-byte loads in flight per thread); the stop test
    // is made once per group, so the scan may run up to three candidates past the bound — they are simply rejected
    const uint32_t n_scan = hard ? 0u : total;
    for (uint32_t j0 = 0; j0 < n_scan; j0 += 4) {
        Cand cs[4];
#pragma unroll
        for (int u = 0; u < 4; u++) cs[u] = load_cand(M, off + (j0 + u < n_scan ? j0 + u : n_scan - 1));
        if (cs[0].key > bound) break;
#pragma unroll
        for (int u = 0; u < 4; u++) {
            if (j0 + u >= n_scan) break;
            const uint32_t sl = cs[u].slot;
            const double d = len2(cs[u].pos - p);
            if (m == k) {
                if (!kv_less(d, sl, tau, tau_sl)) continue;
                H.sift(k, 0, d, sl);                            // replaces the root
            } else {
                H.set(m, d, sl); m++;
                if (m < k) continue;
                for (int i = k / 2 - 1; i >= 0; i--) { double hd; uint32_t hs; H.get(i, hd, hs); H.sift(k, i, hd, hs); }   // k rows filled: heapify
            }
            H.get(0, tau, tau_sl);
            // (sqrt(tau) + delta)^2 from above, in fp32 with every step rounded up (the bound only decides where the scan may stop;
            // an fp64 sqrt after every heap change was 7 % of the kernel's instructions)
            const float bf = __fadd_ru(__fsqrt_ru(__double2float_ru(tau)), delta_f);
            bound = __fmul_ru(__fmul_ru(bf, bf), 1.000001f);
        }
    }
    if (!hard && m > 1) {
        if (m < k) for (int i = m / 2 - 1; i >= 0; i--) { double hd; uint32_t hs; H.get(i, hd, hs); H.sift(m, i, hd, hs); }   // fewer than k candidates: not a heap yet
        for (int n2 = m - 1; n2 > 0; n2--) {   // heap sort: the largest goes to the end, the rest is re-heaped
            double ld, rd; uint32_t ls, rs;
            H.get(n2, ld, ls); H.get(0, rd, rs);
            H.set(n2, rd, rs);
            H.sift(n2, 0, ld, ls);
        }
    }
    d3 res = mk3(0, 0, 0);
    if (total > 0 && !hard) {
        // radiance estimate (raytracer.h:545-576), ascending distance
        for (int r = 0; r < count; r++) {
            double rd_; uint32_t rs_; H.get(r, rd_, rs_);
            const double2* dc = reinterpret_cast<const double2*>(M.dircol + 6 * (size_t)rs_);
            double2 a = __ldg(dc), b = __ldg(dc + 1), c = __ldg(dc + 2);
            res = res + mk3(b.y, c.x, c.y) * dot3(mk3(a.x, a.y, b.x), dq);
        }
        double kd_; uint32_t ks_; H.get(count - 1, kd_, ks_);
        const double den = GI_D_PI * kd_;
        res = mk3(res.x / den, res.y / den, res.z / den);
    }
    if (have && knn && !hard) for (int r = 0; r < k; r++) {
        double rd_ = 0; uint32_t rs_ = 0;
        if (r < count) H.get(r, rd_, rs_);
        knn[(size_t)q * k + r] = r < count ? __ldg(M.pid + rs_) : GI_NO_HIT;
    }
    if (have && !hard) {
        if (rgb) st3(rgb + 3 * (size_t)q, res);
        if (ncand) ncand[q] = total;
        if (accum) {   // render pipeline: L[path] += weight * caustic
            const size_t a = accum_idx[q];
            accum[3 * a] += weight[3 * (size_t)q] * res.x; accum[3 * a + 1] += weight[3 * (size_t)q + 1] * res.y; accum[3 * a + 2] += weight[3 * (size_t)q + 2] * res.z;
        }
    }
    if (work) {
        unsigned long long wt = total, wc = (unsigned long long)count;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { wt += __shfl_down_sync(0xffffffffu, wt, o); wc += __shfl_down_sync(0xffffffffu, wc, o); }
        if (lane == 0) { if (wt) atomicAdd(work + 1, wt); if (wc) atomicAdd(work + 2, wc); }
    }
}

// the long candidate lists: persistent warps take one queued query at a time and stream its list 32 candidates per step
// (gather_at_leaf -> gather_topk_stream, with the ordered-list early stop)
__global__ void __launch_bounds__(GI_WPB * 32) k_gather_heavy(DGatherMap M, const uint32_t* __restrict__ heavy, const uint32_t* __restrict__ heavy_n, uint32_t* __restrict__ next,
                                                             const uint32_t* __restrict__ qnode, const double* __restrict__ qpos, const double* __restrict__ qdir, int k,
                                                             double* __restrict__ rgb, uint32_t* __restrict__ knn, uint32_t* __restrict__ ncand, const double* __restrict__ weight,
                                                             double* __restrict__ accum, const uint32_t* __restrict__ accum_idx)
{
    __shared__ __align__(16) double s_sum[GI_WPB][96];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint32_t nh = *heavy_n;
    for (;;) {
        uint32_t i = 0;
        if (lane == 0) i = atomicAdd(next, 1u);
        i = __shfl_sync(0xffffffffu, i, 0);
        if (i >= nh) return;
        const uint32_t q = heavy[i], node = qnode[q];
        const d3 p = ld3(qpos + 3 * (size_t)q), dq = ld3(qdir + 3 * (size_t)q);
        const DNode nd = load_node(M.nodes, node);
        const double ex = nd.bmax[0] - nd.bmin[0], ey = nd.bmax[1] - nd.bmin[1], ez = nd.bmax[2] - nd.bmin[2];
        const double delta = .5 * sqrt(ex * ex + ey * ey + ez * ez) * (1.0 + 1e-9);
        GatherOut g = gather_at_leaf(M, true, node, p, dq, k, lane, s_sum[wib], delta);
        if (knn && lane < k) knn[(size_t)q * k + lane] = (lane < g.count && g.best_slot != 0xFFFFFFFFu) ? __ldg(M.pid + g.best_slot) : GI_NO_HIT;
        if (lane == 0) {
            if (rgb) st3(rgb + 3 * (size_t)q, g.rgb);
            if (ncand) ncand[q] = g.total;
            if (accum) {
                const size_t a = accum_idx[q];
                accum[3 * a] += weight[3 * (size_t)q] * g.rgb.x; accum[3 * a + 1] += weight[3 * (size_t)q + 1] * g.rgb.y; accum[3 * a + 2] += weight[3 * (size_t)q + 2] * g.rgb.z;
            }
        }
    }
}

// order every leaf's candidate list by distance from the centre of the leaf box and write the keys (see k_gather_sorted).
// One block per leaf; lists of LO < length <= HI entries are sorted by a bitonic network in shared memory ((distance^2,
// slot) pairs, padded to a power of two); longer lists keep the DFS order with key 0, which never stops a scan.
__global__ void k_pm_cand_order(const DNode* __restrict__ nodes, uint32_t n_nodes, const double* __restrict__ pos4, const uint32_t* __restrict__ cand_off, const uint32_t* __restrict__ cand_slot,
                                double* __restrict__ cand_rec, uint32_t lo_len, uint32_t hi_len, uint32_t max_sortable)
{
    extern __shared__ __align__(16) unsigned char s_raw[];
    const uint32_t leaf = blockIdx.x;
    if (leaf >= n_nodes) return;
    const uint32_t off = cand_off[leaf], total = cand_off[leaf + 1] - off;
    if (total <= lo_len || total > hi_len) return;
    auto put = [&](uint32_t j, uint32_t sl, float key) {
        const double* pp = pos4 + 4 * (size_t)sl;
        double* r = cand_rec + 4 * (size_t)(off + j);
        r[0] = pp[0]; r[1] = pp[1]; r[2] = pp[2]; r[3] = __hiloint2double(__float_as_int(key), (int)sl);
    };
    if (total > max_sortable) { for (uint32_t j = threadIdx.x; j < total; j += blockDim.x) put(j, cand_slot[off + j], 0.f); return; }
    uint32_t np2 = 1; while (np2 < total) np2 <<= 1;
    double* kd = reinterpret_cast<double*>(s_raw);
    uint32_t* ks = reinterpret_cast<uint32_t*>(s_raw + (size_t)np2 * 8);
    const DNode nd = load_node(nodes, leaf);
    const d3 c = mk3(nd.bmin[0] + .5 * (nd.bmax[0] - nd.bmin[0]), nd.bmin[1] + .5 * (nd.bmax[1] - nd.bmin[1]), nd.bmin[2] + .5 * (nd.bmax[2] - nd.bmin[2]));
    for (uint32_t j = threadIdx.x; j < np2; j += blockDim.x) {
        if (j < total) { const uint32_t sl = cand_slot[off + j]; const double* pp = pos4 + 4 * (size_t)sl; kd[j] = len2(mk3(pp[0], pp[1], pp[2]) - c); ks[j] = sl; }
        else { kd[j] = CUDART_INF; ks[j] = 0xFFFFFFFFu; }
    }
    __syncthreads();
    for (uint32_t k2 = 2; k2 <= np2; k2 <<= 1)
        for (uint32_t j = k2 >> 1; j > 0; j >>= 1) {
            for (uint32_t x = threadIdx.x; x < np2; x += blockDim.x) {
                const uint32_t y = x ^ j;
                if (y > x) {
                    const bool up = (x & k2) == 0;
                    const double a = kd[x], b = kd[y]; const uint32_t sa = ks[x], sb = ks[y];
                    if (kv_less(b, sb, a, sa) == up) { kd[x] = b; kd[y] = a; ks[x] = sb; ks[y] = sa; }
                }
            }
            __syncthreads();
        }
    for (uint32_t j = threadIdx.x; j < total; j += blockDim.x) put(j, ks[j], __double2float_rd(kd[j]));
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

// ---- ray binning: restore coherence after a scattering bounce ---------------------------------------------------------------------
// Rays leave a diffuse bounce in random directions; a warp of 32 unrelated rays walks 32 different parts of the octree
// (measured: 5.8 of 32 lanes active in k_bounce at depth 1).  Before the next bounce the queue is therefore binned by
// (Morton cell of the origin in the root box, 5 bits per axis | direction octant): a counting sort that only builds a
// permutation — histogram with REDs, exclusive scan, scatter of indices — which the next k_bounce reads its rays through.
// Order inside a bin is arbitrary (atomics); no result depends on queue order: every path owns its accumulator and its
// PRNG key.
#ifndef GI_BIN_AXIS_BITS
#define GI_BIN_AXIS_BITS 5   // cells per axis = 2^bits (C2 frame with octant-only keys: 4 bits 30.8 ms, 5 29.8, 6 29.5, 7 29.9)
#endif
// GI_BIN_DIR: 0 = direction octant in the LOW key bits (cell-major order: a warp holds one or two cells with all their
// octants), 1 = octant in the HIGH bits (direction-major: a warp holds one octant over a run of adjacent Morton cells),
// 2 = octant + dominant axis (24 direction classes, 5 bits) in the high bits
// Measured (frame ms, caustics 1024^2 x 8 / glass 1024^2 x 4; bits per axis in brackets): 0[6] 29.75 / 62.66, 1[6] 30.24 / 60.76,
// 2[5] 29.70 / 59.47, 2[6] 29.27 / 59.91, 3[5] 29.54 / 58.61, 3[6] 29.26 / 60.03, 4[6] 28.88 / 62.41, 2[7] 31.22 (8 M bins: the
// histogram scan costs 2.7 ms).  Finer direction classes pay more than finer cells; 3[5] keeps the 2 M-bin histogram.
#ifndef GI_BIN_DIR
#define GI_BIN_DIR 3
#endif
// 3 = octant + the full ordering of |dx|,|dy|,|dz| (48 classes, 6 bits) in the high bits, 4 = like 2 but in the LOW bits
#define GI_BIN_DIR_BITS (GI_BIN_DIR == 3 ? 6 : ((GI_BIN_DIR == 2 || GI_BIN_DIR == 4) ? 5 : 3))
#define GI_SORT_BITS (3 * GI_BIN_AXIS_BITS + GI_BIN_DIR_BITS)
#define GI_SORT_BINS (1u << GI_SORT_BITS)
__device__ __forceinline__ uint32_t spread3(uint32_t v)   // ...cba -> ..c00b00a (Morton interleave of up to 10 bits)
{
    v &= 0x3FFu;
    v = (v | (v << 16)) & 0x030000FFu;
    v = (v | (v << 8)) & 0x0300F00Fu;
    v = (v | (v << 4)) & 0x030C30C3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}
// Both kernels aggregate their atomics over the lanes of a warp that hold the same key (__match_any_sync): neighbouring queue entries
// come from neighbouring rays of the previous bounce and mostly share a cell, and same-address atomics from one warp are served one
// after the other (the atrium frame spent 58 of 152 ms in these two kernels with one atomic per ray; profiles/r02/README.md).
__global__ void k_bin_keys(uint32_t n, const double* __restrict__ org, const double* __restrict__ dir, d3 bmin, d3 inv_ext, uint32_t* __restrict__ key, uint32_t* __restrict__ hist)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned live = __ballot_sync(0xffffffffu, i < n);
    if (i >= n) return;
    const int C = (1 << GI_BIN_AXIS_BITS) - 1;
    d3 o = ld3(org + 3 * (size_t)i), d = ld3(dir + 3 * (size_t)i);
    int cx = (int)((o.x - bmin.x) * inv_ext.x), cy = (int)((o.y - bmin.y) * inv_ext.y), cz = (int)((o.z - bmin.z) * inv_ext.z);
    cx = cx < 0 ? 0 : (cx > C ? C : cx); cy = cy < 0 ? 0 : (cy > C ? C : cy); cz = cz < 0 ? 0 : (cz > C ? C : cz);
    const uint32_t cell = spread3((uint32_t)cx) | (spread3((uint32_t)cy) << 1) | (spread3((uint32_t)cz) << 2);
    uint32_t dc = (d.x < 0 ? 1u : 0u) | (d.y < 0 ? 2u : 0u) | (d.z < 0 ? 4u : 0u);
#if GI_BIN_DIR == 2 || GI_BIN_DIR == 4
    { const double ax = fabs(d.x), ay = fabs(d.y), az = fabs(d.z); dc |= (ax >= ay && ax >= az ? 0u : (ay >= az ? 1u : 2u)) << 3; }
#endif
#if GI_BIN_DIR == 3
    { const double ax = fabs(d.x), ay = fabs(d.y), az = fabs(d.z); dc |= ((ax >= ay ? 1u : 0u) | (ay >= az ? 2u : 0u) | (ax >= az ? 4u : 0u)) << 3; }
#endif
#if GI_BIN_DIR == 0 || GI_BIN_DIR == 4
    uint32_t k = (cell << GI_BIN_DIR_BITS) | dc;
#else
    uint32_t k = (dc << (3 * GI_BIN_AXIS_BITS)) | cell;
#endif
    key[i] = k;
    const unsigned peers = __match_any_sync(live, k);
    if ((threadIdx.x & 31u) == (unsigned)(__ffs(peers) - 1)) atomicAdd(hist + k, (uint32_t)__popc(peers));
}
__global__ void k_bin_scatter(uint32_t n, const uint32_t* __restrict__ key, uint32_t* __restrict__ cursor, uint32_t* __restrict__ perm)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned live = __ballot_sync(0xffffffffu, i < n);
    if (i >= n) return;
    const uint32_t k = key[i];
    const unsigned lane = threadIdx.x & 31u;
    const unsigned peers = __match_any_sync(live, k);
    const int leader = __ffs(peers) - 1;
    uint32_t base = 0;
    if (lane == (unsigned)leader) base = atomicAdd(cursor + k, (uint32_t)__popc(peers));
    base = __shfl_sync(peers, base, leader);
    perm[base + (uint32_t)__popc(peers & ((1u << lane) - 1u))] = i;
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
// per path: Halton index, PRNG key, and the radiance sum in two parts — L takes the ambient / emissive / direct terms in bounce
// order, Lc the caustic terms in bounce order; the path's radiance is L + Lc.  (Keeping the caustic terms apart lets the tail
// hand its gathers to one batched gather run without changing the order of any sum.)
// Ld takes the direct-light terms in bounce order on their own: k_direct (and the gather pipeline) of depth d then touch
// nothing the bounce kernel of depth d+1 touches, so the host may run them on side streams behind it; radiance = (L + Ld) + Lc.
struct DPathState { uint32_t* sample; uint64_t* key; double* L; double* Lc; double* Ld; };
struct DCounters { uint32_t n_next, n_hits; unsigned long long closest, shadow, gathers; };

template <int MODE, bool IMPL>
__global__ void __launch_bounds__(GI_BLOCK, GI_MINB) k_bounce(DScene S, gi_render_params P, int depth, uint32_t n, DQueue in, const uint32_t* __restrict__ perm, DQueue out, DHitList H,
                                                     DPathState PS, DCounters* C, unsigned long long* work)
{
    constexpr bool FULL = MODE != 0, FOG = MODE == 2;   // MODE 0: uv-writing opaque primitives only, 1: any scene, 2: any scene + atmosphere
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    bool active = i < n;
    if (active && perm) i = perm[i];   // binned order (k_bin_*)
    bool is_hit = false, cont = false;
    d3 hp, hn, refDir, wdir, wcau, Tn, contrib;
    double rough = 1, offset = GI_D_SHADOW_BIAS;
    uint32_t path = 0, wn = 0, wp = 0;
    if (active) {
        path = in.path[i];
        DRay r = ray_as_stored(ld3(in.o + 3 * (size_t)i), ld3(in.d + 3 * (size_t)i));
        uint64_t key = PS.key[path];
        DHit h;
        trace_closest<FULL, IMPL>(S, r, P.seed, key, (uint64_t)depth, h, wn, wp);  // :190
        d3 T = ld3(in.T + 3 * (size_t)i);               // read after the walk: twelve registers less across it
        contrib = ld3(in.contrib + 3 * (size_t)i);
        double* L = PS.L + 3 * (size_t)path;
        if (h.prim == GI_NO_HIT) {
            d3 a = T * ld3(S.ambient);                                       // :275
            L[0] += a.x; L[1] += a.y; L[2] += a.z;
        } else {
            is_hit = true;
            const uint32_t sample = PS.sample[path];
            const float sx = halton_sample(S, (uint32_t)(2 + 2 * depth), sample);     // raytracer.h:172-173 (only a hit uses them)
            const float sy = halton_sample(S, (uint32_t)(3 + 2 * depth), sample);
            double tu, tv;
            hit_surface(S, r, h, FULL, hp, hn, tu, tv);
            const gi_material& m = S.mats[S.prim_mat[h.prim]];
            d3 color = tex_get(S, m.diffuse_tex, tu, tv);                    // :200
            rough = m.roughness;
            d3 f = mk3(1, 1, 1);
            refDir = secondary_ray(S, m, r, hn, tu, tv, sx, sy, color, f, contrib, offset, P.seed, key, (uint64_t)depth);   // :207
            if (FOG) {   // raytracer.h:209-228: the segment scattered in a volume before reaching the surface
                d3 fh, fc;
                if (fog_scatter(S, r, hp, fh, fc, P.seed, key, (uint64_t)depth, SITE_FOG_RAD)) { hp = fh; refDir = random_unit_vec(sx, sy); f = fc; color = fc; contrib = fc; rough = 1; }
            }
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

// Persistent warps with ray refetch.  Rays of one warp differ widely in length (ncu, sponza stand-in at 4K: 6.3 of 32 lanes
// active through the whole walk, 139 node tests per ray on average with a long tail): a warp that keeps its 32 rays until the
// last one is done spends most of its issue slots on a handful of lanes.  Here a lane whose ray is done is shaded and handed
// the next ray of the queue (one atomicAdd per warp and refill) as soon as fewer than GI_REFILL_MIN lanes are still walking;
// the walk itself is the resumable trace_walk.  Per-ray arithmetic is unchanged; the order in which rays reach the output
// queue / hit list changes, which no result depends on.  Measured (bounce ms per frame, classic -> persistent): sponza stand-in
// 4483 -> 3092, but caustics 11.8 -> 12.9 and glass 1118 -> 1208 (their warps are fuller to begin with and the kept-alive ray
// state costs spills), so the host picks the form per scene from the node tests per ray of the previous frame.
#ifndef GI_REFILL_MIN
#define GI_REFILL_MIN 20
#endif
template <int MODE, bool IMPL>
__global__ void __launch_bounds__(GI_BLOCK, GI_MINB) k_bounce_p(DScene S, gi_render_params P, int depth, uint32_t n, DQueue in, const uint32_t* __restrict__ perm, DQueue out, DHitList H,
                                                       DPathState PS, DCounters* C, unsigned long long* work, uint32_t* next_ray)
{
    constexpr bool FULL = MODE != 0, FOG = MODE == 2;   // MODE 0: uv-writing opaque primitives only, 1: any scene, 2: any scene + atmosphere
    __shared__ int s_walking[GI_BLOCK / 32];
    const unsigned lane = threadIdx.x & 31u;
    const int wib = threadIdx.x >> 5;
    GI_TSTACK_DECL(stack);
    TraceState st; st.sp = 0; st.term = false; st.best_d2 = 0; st.cur_tu = 0; st.cur_tv = 0; st.frac_t = CUDART_INF;
    DHit h; h.prim = GI_NO_HIT;
    DRay r = ray_as_stored(mk3(0, 0, 0), mk3(1, 0, 0));
    uint32_t qi = 0, path = 0, wn = 0, wp = 0;
    uint64_t key = 0;
    bool walking = false, finished = false, more = true;
    for (;;) {
        // ---- refill: lanes without a ray take the next ones of the queue
        const unsigned want = __ballot_sync(0xffffffffu, !walking);
        if (more && want) {
            uint32_t base = 0;
            if (lane == (unsigned)(__ffs(want) - 1)) base = atomicAdd(next_ray, (uint32_t)__popc(want));
            base = __shfl_sync(0xffffffffu, base, __ffs(want) - 1);
            more = base + (uint32_t)__popc(want) < n;
            if (!walking) {
                const uint32_t j = base + __popc(want & ((1u << lane) - 1u));
                if (j < n) {
                    qi = perm ? perm[j] : j;   // binned order (k_bin_*)
                    path = in.path[qi];
                    key = PS.key[path];
                    r = ray_as_stored(ld3(in.o + 3 * (size_t)qi), ld3(in.d + 3 * (size_t)qi));
                    walking = trace_begin(S, r, h, st, stack, wn);
                    finished = !walking;   // missed the root box: shaded as a miss right away
                }
            }
        }
        const unsigned wm = __ballot_sync(0xffffffffu, walking);
        if (!wm && !__any_sync(0xffffffffu, finished)) break;   // queue drained and every lane idle
        if (lane == 0) s_walking[wib] = __popc(wm);
        __syncwarp();
        // ---- walk until done, or until the warp has thinned out and there are rays left to fetch
        if (walking) {
            trace_walk<FULL, IMPL, true>(S, r, P.seed, key, (uint64_t)depth, h, st, stack, wn, wp, &s_walking[wib], more ? GI_REFILL_MIN : 0);   // :190
            if (st.sp == 0 || st.term) { walking = false; finished = true; }
        }
        __syncwarp();
        // ---- shade the finished rays (raytracer.h:167-276), emit hit-list entries and continuation rays
        bool is_hit = false, cont = false;
        d3 hp, hn, refDir, wdir, wcau, Tn, contrib;
        double rough = 1, offset = GI_D_SHADOW_BIAS;
        if (finished) {
            d3 T = ld3(in.T + 3 * (size_t)qi);
            contrib = ld3(in.contrib + 3 * (size_t)qi);
            double* L = PS.L + 3 * (size_t)path;
            if (h.prim == GI_NO_HIT) {
                d3 a = T * ld3(S.ambient);                                       // :275
                L[0] += a.x; L[1] += a.y; L[2] += a.z;
            } else {
                is_hit = true;
                const uint32_t sample = PS.sample[path];
                const float sx = halton_sample(S, (uint32_t)(2 + 2 * depth), sample);     // raytracer.h:172-173 (only a hit uses them)
                const float sy = halton_sample(S, (uint32_t)(3 + 2 * depth), sample);
                double tu, tv;
                hit_surface(S, r, h, FULL, hp, hn, tu, tv);
                const gi_material& m = S.mats[S.prim_mat[h.prim]];
                d3 color = tex_get(S, m.diffuse_tex, tu, tv);                    // :200
                rough = m.roughness;
                d3 f = mk3(1, 1, 1);
                refDir = secondary_ray(S, m, r, hn, tu, tv, sx, sy, color, f, contrib, offset, P.seed, key, (uint64_t)depth);   // :207
                if (FOG) {   // raytracer.h:209-228: the segment scattered in a volume before reaching the surface
                    d3 fh, fc;
                    if (fog_scatter(S, r, hp, fh, fc, P.seed, key, (uint64_t)depth, SITE_FOG_RAD)) { hp = fh; refDir = random_unit_vec(sx, sy); f = fc; color = fc; contrib = fc; rough = 1; }
                }
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
            finished = false;
        }
        // queue compaction: warp ballot + prefix popcount, one atomic per warp and list
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
    tally2(work, wn, wp);
}

// ---- K3 in the pipeline: direct light with one shadow ray per light (raytracer.h:230-256) -----------------------------------------
template <int MODE, bool IMPL>
__global__ void __launch_bounds__(GI_BLOCK, GI_MINB) k_direct(DScene S, gi_render_params P, int depth, uint32_t n, DHitList H, DPathState PS, unsigned long long* work)
{
    constexpr bool FULL = MODE != 0, FOG = MODE == 2;   // MODE 0: uv-writing opaque primitives only, 1: any scene, 2: any scene + atmosphere
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t wn = 0, wp = 0;
    if (i >= n) { tally2(work, 0, 0); return; }
    uint32_t path = H.path[i];
    uint64_t key = PS.key[path];
    d3 li = mk3(0, 0, 0);
    for (uint32_t l = 0; l < S.n_lights; l++) {
        const gi_light& light = S.lights[l];
        DRay sr; double maxt;
        {
            const d3 p = ld3(H.p + 3 * (size_t)i), nn = ld3(H.n + 3 * (size_t)i);
            d3 sp = p + nn * GI_D_SHADOW_BIAS;
            d3 lightDir = light_point(light, gi_rand(P.seed, key, (uint64_t)depth, SITE(SITE_LIGHT_U, l)), gi_rand(P.seed, key, (uint64_t)depth, SITE(SITE_LIGHT_V, l))) - sp;   // :233
            maxt = len2(lightDir);
            sr = make_ray(sp, lightDir);                                                                                      // :241
        }
        bool vis = trace_visible<FULL, IMPL>(S, sr, maxt, P.seed, key, (uint64_t)depth, l, wn, wp);                                 // :243
        if (FOG && vis && fog_blocks(S, sr, maxt, P.seed, key, (uint64_t)depth, l)) vis = false;                              // :308-316
        if (vis) {
            // the hit point, normal and roughness are read AGAIN here instead of kept across the walk: eleven live doubles next to the
            // ray made ptxas spill nine of them inside the interior loop of the traversal (9 STL.64 + 14 LDL.64 per node step, three
            // times the bytes of the node record itself).  The pointers go through an empty asm so that the loads are not merged.
            const double* hp = H.p; const double* hn = H.n; const double* hr = H.rough;
            asm volatile("" : "+l"(hp), "+l"(hn), "+l"(hr));
            const d3 p = ld3(hp + 3 * (size_t)i), nn = ld3(hn + 3 * (size_t)i);
            const double rough = hr[i];
            const double hfrac = 1 / (GI_D_PI * len2(ld3(light.pos) - p));                                                   // :238
            double d = dot3(nn, normalize3(ld3(light.pos) - p));
            if (d < 0) d = 0;
            double lv = pow_like_libm(d, (1.0 / rough));                                                                      // :252
            li = (ld3(light.col) * lv) * hfrac;                                                                               // assignment, not += (:254)
        }
    }
    d3 w = ld3(H.wdirect + 3 * (size_t)i) * li;
    double* L = PS.Ld + 3 * (size_t)path;
    L[0] += w.x; L[1] += w.y; L[2] += w.z;
    tally2(work, wn, wp);
}

// ---- the tail: once few paths are left, one warp takes one path to its end (closest hit, shading, shadow rays, gather and
// the next bounce all inside the kernel), so a frame does not pay ~65 x 3 nearly empty launches with a host round trip
// each.  Arithmetic and the order in which terms are added to L[path] are those of k_bounce / k_direct / k_gather.
// ---- batched tail gathers ---------------------------------------------------------------------------------------------------------
// The tail does not make its gathers inline: every gather query of a tail path goes into the path's own run of slots (one
// slot per remaining gather depth), ONE gather pipeline run then serves all tail queries (leaf-ordered, thread per query —
// instead of a warp streaming one candidate list at a time in the middle of a path), and k_tail_caustic adds weight x
// estimate to Lc[path] in bounce order.  This keeps the latency chain of a deep path to closest hit + shading + shadow ray and
// takes the gather code out of the tail kernel (ncu showed it starved for instructions: 45 KB of code, ~10 warps per SM).
// (A one-THREAD-per-path tail was measured too: 14.6 ms instead of 4.2 — the tail is bound by the latency of its deepest
// paths, and a warp walking one ray cooperatively has a third of the per-bounce latency of a thread.)
struct DTailQ { double* pos; double* dir; double* w; double* rgb; uint32_t* count; uint32_t qmax;
                // deferred shadow rays (smax > 0): one slot per (bounce of the path, light), in bounce order — origin, stored direction,
                // squared distance bound, weight x unshadowed light term, bounce depth; sh_count[path] = slots used, sh_vis = result
                double* sh_o; double* sh_d; double* sh_mt; double* sh_w; uint32_t* sh_depth; uint32_t* sh_count; uint8_t* sh_vis; uint32_t smax; };

struct DTailCounters { unsigned long long closest, shadow, gathers, nodes_c, prims_c, nodes_s, prims_s, g_depth, g_cand, g_sel; unsigned int next; unsigned int pad; };

#ifndef GI_TAIL_MINB
#define GI_TAIL_MINB 4   // resident blocks of 4 warps per SM asked of ptxas for the tail kernel (128 registers).  With the shadow rays
                         // deferred the kernel is small enough for it: tail 2.78 -> 2.70 ms on C2, 2.93 -> 2.62 on glass (3 -> 4; 5 loses)
#endif
template <int MODE, bool IMPL>
__global__ void __launch_bounds__(GI_WPB * 32, GI_TAIL_MINB) k_tail(DScene S, DGatherMap G, int have_map, gi_render_params P, int depth0, uint32_t n, DQueue in, DPathState PS, DTailCounters* TC, DTailQ Q)
{
    constexpr bool FULL = MODE != 0, FOG = MODE == 2;   // MODE 0: uv-writing opaque primitives only, 1: any scene, 2: any scene + atmosphere
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
    // Ld: k_tail_direct, from the queued shadow rays.  Lc is this kernel's only when it makes its gathers inline (lc_rw); with queued
    // gathers it must not even be read and written back: gather runs of earlier depths may still be adding to it on a side stream
    const bool lc_rw = have_map && Q.qmax == 0 && depth0 <= P.caustic_max_depth;
    d3 L = ld3(PS.L + 3 * (size_t)path), Lc = lc_rw ? ld3(PS.Lc + 3 * (size_t)path) : mk3(0, 0, 0);
    uint32_t nq = 0;   // gather queries queued by this path
    uint32_t nsh = 0;  // shadow rays queued by this path
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
        if (FOG) {   // raytracer.h:209-228: the segment scattered in a volume before reaching the surface
            d3 fh, fc;
            if (fog_scatter(S, r, hp, fh, fc, P.seed, key, (uint64_t)depth, SITE_FOG_RAD)) { hp = fh; refDir = random_unit_vec(sx, sy); f = fc; color = fc; contrib = fc; rough = 1; }
        }
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
        {
            for (uint32_t l = 0; l < S.n_lights; l++) {
                const gi_light& light = S.lights[l];
                d3 sp = hp + hn * GI_D_SHADOW_BIAS;
                d3 lightDir = light_point(light, gi_rand(P.seed, key, (uint64_t)depth, SITE(SITE_LIGHT_U, l)), gi_rand(P.seed, key, (uint64_t)depth, SITE(SITE_LIGHT_V, l))) - sp;
                double maxt = len2(lightDir);
                double hfrac = 1 / (GI_D_PI * len2(ld3(light.pos) - hp));
                DRay sr = make_ray(sp, lightDir);
                c_shadow++;
                // deferred: the shadow ray does not decide how the path goes on, so it leaves the latency chain of the path —
                // k_tail_shadow traces all of them at once, k_tail_direct adds the terms to Ld in bounce order
                double d = dot3(hn, normalize3(ld3(light.pos) - hp));
                if (d < 0) d = 0;
                const d3 lv = (ld3(light.col) * pow_like_libm(d, (1.0 / rough))) * hfrac;
                if (lane == 0 && nsh < Q.smax) {
                    const size_t s = (size_t)i * Q.smax + nsh;
                    st3(Q.sh_o + 3 * s, sr.o); st3(Q.sh_d + 3 * s, sr.d); Q.sh_mt[s] = maxt; st3(Q.sh_w + 3 * s, wdir * lv); Q.sh_depth[s] = (uint32_t)depth;
                }
                nsh++;
            }
        }
        // caustic estimate (k_gather)
        if (depth <= P.caustic_max_depth) {
            c_gather++;
            if (have_map && Q.qmax) {   // queued for the batched gather run
                if (lane == 0 && nq < Q.qmax) {
                    const size_t s = (size_t)i * Q.qmax + nq;
                    st3(Q.pos + 3 * s, hp); st3(Q.dir + 3 * s, refDir); st3(Q.w + 3 * s, cont ? wdir : mk3(0, 0, 0));
                }
                nq++;
            } else if (have_map) {
                GatherOut g = gather_warp(G, hp, refDir, P.k_photons, lane, s_sum[wib]);
                g_depth += g.depth; g_cand += g.total; g_sel += (unsigned long long)g.count;
                d3 wc = cont ? wdir : mk3(0, 0, 0);
                Lc = Lc + wc * g.rgb;
            }
        }
        if (!cont) break;
        T = Tn;
        r = make_ray(hp + hn * offset, refDir);
    }
    if (lane == 0) {
        st3(PS.L + 3 * (size_t)path, L);
        if (lc_rw) st3(PS.Lc + 3 * (size_t)path, Lc);
        if (Q.smax) Q.sh_count[i] = nsh < Q.smax ? nsh : Q.smax;
        if (have_map && Q.qmax) {
            Q.count[i] = nq < Q.qmax ? nq : Q.qmax;
            for (uint32_t k = nq; k < Q.qmax; k++) st3(Q.pos + 3 * ((size_t)i * Q.qmax + k), mk3(CUDART_INF, CUDART_INF, CUDART_INF));   // unused slots: in no leaf
        }
        atomicAdd(&TC->closest, c_closest); atomicAdd(&TC->shadow, c_shadow); atomicAdd(&TC->gathers, c_gather);
        atomicAdd(&TC->nodes_c, (unsigned long long)nc); atomicAdd(&TC->prims_c, (unsigned long long)pc); atomicAdd(&TC->nodes_s, (unsigned long long)ns); atomicAdd(&TC->prims_s, (unsigned long long)ps);
        atomicAdd(&TC->g_depth, g_depth); atomicAdd(&TC->g_cand, g_cand); atomicAdd(&TC->g_sel, g_sel);
    }
  }
}

__global__ void k_tail_caustic(uint32_t n, DQueue in, DPathState PS, DTailQ Q)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t path = in.path[i];
    d3 Lc = ld3(PS.Lc + 3 * (size_t)path);
    const uint32_t nq = Q.count[i];
    for (uint32_t k = 0; k < nq; k++) {
        const size_t s = (size_t)i * Q.qmax + k;
        Lc = Lc + ld3(Q.w + 3 * s) * ld3(Q.rgb + 3 * s);
    }
    st3(PS.Lc + 3 * (size_t)path, Lc);
}

// Order of the camera paths of a tile: pixel <-> slot.  Row-major order puts 32 pixels of ONE row into a warp; here the tile is
// cut into strips of GI_TILE_ROWS rows walked column by column, so a warp covers an (32 / GI_TILE_ROWS) x GI_TILE_ROWS block of
// pixels: closer rays, closer hit points, closer shadow rays.  Only the order in which paths sit in the queues changes — every
// path owns its state, frames are bit-identical (tested: tiles / chunks compose).  The last strip may be shorter.
// Measured (frame ms, caustics / glass as above, with GI_BIN_DIR 3[5]): rows 1: 29.51 / 58.67, 2: 29.08, 4: 30.91 (the bounce-form
// autotune flipped) / 57.54, 8: 29.12 / 57.85.
#ifndef GI_TILE_ROWS
#define GI_TILE_ROWS 8
#endif
__device__ __forceinline__ void slot_to_pixel(size_t slot, int tw, int th, int& lx, int& ly)
{
    const size_t per = (size_t)GI_TILE_ROWS * tw;
    const int strip = (int)(slot / per);
    const size_t j = slot - (size_t)strip * per;
    const int hh = min(GI_TILE_ROWS, th - strip * GI_TILE_ROWS);
    lx = (int)(j / hh); ly = strip * GI_TILE_ROWS + (int)(j % hh);
}
__device__ __forceinline__ size_t pixel_to_slot(int lx, int ly, int tw, int th)
{
    const int strip = ly / GI_TILE_ROWS;
    const int hh = min(GI_TILE_ROWS, th - strip * GI_TILE_ROWS);
    return (size_t)strip * GI_TILE_ROWS * tw + (size_t)lx * hh + (size_t)(ly - strip * GI_TILE_ROWS);
}

// the tail's deferred shadow rays: one thread per slot (RayTracer::visible, raytracer.h:280-319), then one thread per path adds the
// unshadowed terms to Ld in bounce order — per bounce the LAST visible light's term, like the reference's `i = ...` (raytracer.h:254)
template <int MODE, bool IMPL>
__global__ void __launch_bounds__(GI_BLOCK, GI_MINB) k_tail_shadow(DScene S, gi_render_params P, uint32_t n, DQueue in, DPathState PS, DTailQ Q, unsigned long long* work)
{
    constexpr bool FULL = MODE != 0, FOG = MODE == 2;
    const size_t s = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    uint32_t wn = 0, wp = 0;
    const uint32_t i = (uint32_t)(s / Q.smax), k = (uint32_t)(s % Q.smax);
    if (i < n && k < Q.sh_count[i]) {
        const uint64_t key = PS.key[in.path[i]];
        const uint32_t depth = Q.sh_depth[s], l = k % S.n_lights;
        const DRay sr = ray_as_stored(ld3(Q.sh_o + 3 * s), ld3(Q.sh_d + 3 * s));
        const double maxt = Q.sh_mt[s];
        bool vis = trace_visible<FULL, IMPL>(S, sr, maxt, P.seed, key, (uint64_t)depth, l, wn, wp);
        if (FOG && vis && fog_blocks(S, sr, maxt, P.seed, key, (uint64_t)depth, l)) vis = false;
        Q.sh_vis[s] = vis ? 1 : 0;
    }
    tally2(work, wn, wp);
}
__global__ void k_tail_direct(uint32_t n, uint32_t n_lights, DQueue in, DPathState PS, DTailQ Q)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t path = in.path[i];
    d3 Ld = ld3(PS.Ld + 3 * (size_t)path);
    const uint32_t ns = Q.sh_count[i];
    for (uint32_t b = 0; b + n_lights <= ns; b += n_lights) {   // one bounce = n_lights consecutive slots
        d3 term = mk3(0, 0, 0);
        for (uint32_t l = 0; l < n_lights; l++) {
            const size_t s = (size_t)i * Q.smax + b + l;
            if (Q.sh_vis[s]) term = ld3(Q.sh_w + 3 * s);
        }
        Ld = Ld + term;
    }
    st3(PS.Ld + 3 * (size_t)path, Ld);
}

// generate the camera paths of one chunk (path-linear range [c0, c0+n) of the tile's sample-major path space)
__global__ void k_generate(DScene S, DFrame F, int s0, uint64_t c0, uint32_t n, DQueue q, DPathState PS)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t lin = c0 + i;
    size_t npx = (size_t)F.tw * F.th;
    int s = s0 + (int)(lin / npx);
    size_t slot = lin % npx;
    int lx, ly;
    slot_to_pixel(slot, F.tw, F.th, lx, ly);
    int y = frame_row(F, ly), x = F.x0 + lx;
    uint32_t idx;
    DRay r = camera_ray(S, F, x, y, s, idx);
    st3(q.o + 3 * (size_t)i, r.o); st3(q.d + 3 * (size_t)i, r.d);
    st3(q.T + 3 * (size_t)i, mk3(1, 1, 1)); st3(q.contrib + 3 * (size_t)i, mk3(1, 1, 1));
    q.path[i] = i;
    PS.sample[i] = idx;
    PS.key[i] = ((uint64_t)((uint64_t)y * (uint64_t)F.w + (uint64_t)x) << 24) | (uint64_t)s;
    PS.L[3 * (size_t)i] = 0; PS.L[3 * (size_t)i + 1] = 0; PS.L[3 * (size_t)i + 2] = 0;
    PS.Lc[3 * (size_t)i] = 0; PS.Lc[3 * (size_t)i + 1] = 0; PS.Lc[3 * (size_t)i + 2] = 0;
    PS.Ld[3 * (size_t)i] = 0; PS.Ld[3 * (size_t)i + 1] = 0; PS.Ld[3 * (size_t)i + 2] = 0;
}

// ---- adaptive sampling: the per-pixel loop of RayTracer::run (raytracer.h:100-148) as passes over the sample index --------------------
// Per pixel the reference keeps (color = running mean, var = smoothed change of the mean, s = samples taken, samps) and goes on
// while s < max_samples && samps < min_samples; every sample adds 1 to samps, a sample that leaves var above the threshold takes
// 2 away again.  A pixel that stops never resumes, so pass s renders sample s of the pixels still active: k_adapt_select lists
// them (ballot + prefix popcount), k_generate_list makes their camera paths, the wavefront runs as usual, k_adapt_update folds
// the path radiance into the pixel state with the reference's arithmetic.
struct DAdapt { double* color; double* var; int* samps; int* s_done; uint32_t* list; uint32_t* n_list; int min_samples, max_samples; double noise_thresh; };

__global__ void k_adapt_init(size_t npx, DAdapt A)
{
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= npx) return;
    A.color[3 * i] = 0.5; A.color[3 * i + 1] = 0.5; A.color[3 * i + 2] = 0.5;   // raytracer.h:102
    A.var[i] = 0; A.samps[i] = 0; A.s_done[i] = 0;
}
__global__ void k_adapt_select(uint32_t npx, int s, DAdapt A)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool on = i < npx && A.s_done[i] == s && s < A.max_samples && A.samps[i] < A.min_samples;   // :108
    const unsigned lane = threadIdx.x & 31u, m = __ballot_sync(0xffffffffu, on);
    uint32_t base = 0;
    if (lane == 0 && m) base = atomicAdd(A.n_list, (uint32_t)__popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (on) A.list[base + __popc(m & ((1u << lane) - 1u))] = i;
}
// camera paths for sample s of the listed pixels (tile-linear pixel indices)
__global__ void k_generate_list(DScene S, DFrame F, int s, uint32_t n, const uint32_t* __restrict__ list, DQueue q, DPathState PS)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t pix = list[i];
    int y = frame_row(F, (int)(pix / F.tw)), x = F.x0 + (int)(pix % F.tw);
    uint32_t idx;
    DRay r = camera_ray(S, F, x, y, s, idx);
    st3(q.o + 3 * (size_t)i, r.o); st3(q.d + 3 * (size_t)i, r.d);
    st3(q.T + 3 * (size_t)i, mk3(1, 1, 1)); st3(q.contrib + 3 * (size_t)i, mk3(1, 1, 1));
    q.path[i] = i;
    PS.sample[i] = idx;
    PS.key[i] = ((uint64_t)((uint64_t)y * (uint64_t)F.w + (uint64_t)x) << 24) | (uint64_t)s;
    PS.L[3 * (size_t)i] = 0; PS.L[3 * (size_t)i + 1] = 0; PS.L[3 * (size_t)i + 2] = 0;
    PS.Lc[3 * (size_t)i] = 0; PS.Lc[3 * (size_t)i + 1] = 0; PS.Lc[3 * (size_t)i + 2] = 0;
    PS.Ld[3 * (size_t)i] = 0; PS.Ld[3 * (size_t)i + 1] = 0; PS.Ld[3 * (size_t)i + 2] = 0;
}
__global__ void k_adapt_update(uint32_t n, int s, const uint32_t* __restrict__ list, const double* __restrict__ L, const double* __restrict__ Ld, const double* __restrict__ Lc, DAdapt A)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t pix = list[i];
    const d3 rad = (ld3(L + 3 * (size_t)i) + ld3(Ld + 3 * (size_t)i)) + ld3(Lc + 3 * (size_t)i);
    const d3 last = ld3(A.color + 3 * (size_t)pix);                                      // lastCol = color (:110)
    d3 color = s == 0 ? rad : (last * (1.0 * s) + rad) * (1.0 / (s + 1));              // :131-134
    double var = A.var[pix];
    int samps = A.samps[pix];
    if (s > 0) {
        const d3 dc = color - last;
        var = (1.0 * 5 * var + sqrt(dot3(dc, dc))) * (1.0 / (5 + 1));                   // :138 (glm::length)
        if (var > A.noise_thresh) samps -= 2;                                            // :143-144
    }
    st3(A.color + 3 * (size_t)pix, color);
    A.var[pix] = var; A.samps[pix] = samps + 1; A.s_done[pix] = s + 1;                   // :146-147
}

// add the chunk's per-path radiance into the tile accumulator, samples in ascending order per pixel
__global__ void k_accumulate(uint64_t c0, uint32_t n, size_t npx, int tw, int th, const double* L, const double* Ld, const double* Lc, double* accum)
{
    size_t pix = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (pix >= npx) return;
    uint64_t c1 = c0 + n;
    // paths of this pixel inside the chunk: lin = k*npx + slot(pixel)
    const size_t slot = pixel_to_slot((int)(pix % tw), (int)(pix / tw), tw, th);
    uint64_t k0 = c0 > slot ? (c0 - slot + npx - 1) / npx : 0;
    double a0 = accum[3 * pix], a1 = accum[3 * pix + 1], a2 = accum[3 * pix + 2];
    for (uint64_t k = k0;; k++) {
        uint64_t lin = k * npx + slot;
        if (lin >= c1) break;
        size_t i = (size_t)(lin - c0);
        a0 += (L[3 * i] + Ld[3 * i]) + Lc[3 * i]; a1 += (L[3 * i + 1] + Ld[3 * i + 1]) + Lc[3 * i + 1]; a2 += (L[3 * i + 2] + Ld[3 * i + 2]) + Lc[3 * i + 2];
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

// one emission try of photon index i at light li (raytracer.h:604-695): true when a photon was stored into `out`
template <bool FULL, bool IMPL, bool FOG>
__device__ __forceinline__ bool photon_try(const DScene& S, int count, int max_depth, uint64_t seed, int i, uint32_t li, int tries, d3& ph_pos, d3& ph_dir, d3& ph_col, unsigned long long& n_traces,
                                           uint32_t& wn, uint32_t& wp)
{
    const gi_light& l = S.lights[li];
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
    trace_closest<FULL, IMPL>(S, r, seed, path, 0, h, wn, wp); n_traces++;
    if (h.prim == GI_NO_HIT) return false;                                              // :626-630
    uint32_t cur = h.prim;
    d3 hit = mk3(0, 0, 0);
    while (depth < max_depth && !term) {                                                // :633
        double roughness = S.mats[S.prim_mat[cur]].roughness;
        if (roughness < 0.1) {
            // :640 traces the ray again.  At depth 0 it is the very ray traced above (:620); without stochastic alpha (FULL) the
            // closest hit is a pure function of the ray, so the first answer stands (the trace is still counted)
            if (FULL || depth > 0) trace_closest<FULL, IMPL>(S, r, seed, path, (uint64_t)(depth + 1), h, wn, wp);
            n_traces++;
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
            if (FOG) {                                                                  // :658-675
                d3 ah, ac;
                if (fog_scatter(S, r, hit, ah, ac, seed, path, (uint64_t)(depth + 1), SITE_FOG_PHOTON)) {
                    const double fu = fmod(gi_rand(seed, path, (uint64_t)(depth + 1), SITE(SITE_PH_FOG_U, 0)) + 13 * i, 1.0);
                    const double fv = fmod(gi_rand(seed, path, (uint64_t)(depth + 1), SITE(SITE_PH_FOG_V, 0)) + 7 * i, 1.0);
                    hit = ah; refDir = random_unit_vec(fu, fv); f = ac; roughness = 1;
                }
            }
            col = col * f;                                                              // :677
            r = make_ray(hit + norm * offset, refDir);                                  // :679-680
            isCaustic = true;
        }
        if (depth > 0 && isCaustic && roughness >= 0.1) {                               // :685-692
            ph_pos = hit; ph_dir = r.d; ph_col = col;
            return true;
        }
        depth++;
    }
    return false;
}

// The reference gives every photon index up to 500 emission tries (raytracer.h:602) and stores the FIRST that lands on a
// caustic caster (one try in ten on caustics, one in forty on glass).  A thread that loops over its tries keeps 31 finished
// lanes waiting for the unluckiest one; and one try can be slow (a photon ray that misses everything walks every leaf along
// the whole line: the closest-hit rule has no early exit for a miss), so a kernel per try pays that latency 500 times.
// Tries are therefore run in ROUNDS of `width` speculative tries per active (photon index, light) slot: `width` adjacent
// lanes make tries base .. base+width-1 of the same slot at once, the lowest successful try wins (ballot + ffs) and writes
// the photon, slots without a winner are compacted (ballot + prefix popcount) into the next round's list.  All active slots
// have consumed the same number of tries, so the survivor list is all the state there is.  The host grows `width` as the
// list shrinks (1, 2, 4 .. 32) and keeps ~half a million lanes busy: ~25 rounds instead of 500.  Tries and traces are
// tallied as the sequential loop would have made them (up to and including the winner); same try, same counter-PRNG keys,
// same photons.
template <int MODE, bool IMPL>
__global__ void __launch_bounds__(GI_BLOCK, GI_MINB) k_photon_round(DScene S, int count, int max_depth, uint64_t seed, int base_try, int width, uint32_t n_active, const uint32_t* __restrict__ in_list,
                                                                  uint32_t* __restrict__ out_list, uint32_t* __restrict__ n_out, DPhotonOut O)
{
    constexpr bool FULL = MODE != 0, FOG = MODE == 2;   // MODE 0: uv-writing opaque primitives only, 1: any scene, 2: any scene + atmosphere
    const uint32_t gt = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned lane = threadIdx.x & 31u;
    const uint32_t j = gt / (uint32_t)width;       // active-list entry
    const int k = (int)(gt % (uint32_t)width);      // which of the round's tries
    const int tries = base_try + k;
    unsigned long long my_traces = 0;
    uint32_t wn = 0, wp = 0, slot = 0;
    bool stored = false;
    d3 pp = mk3(0, 0, 0), pd = pp, pc = pp;
    const bool live = j < n_active && tries < 500;
    if (j < n_active) slot = in_list ? in_list[j] : j;
    if (live) {
        const int i = (int)(slot / S.n_lights); const uint32_t li = slot % S.n_lights;
        stored = photon_try<FULL, IMPL, FOG>(S, count, max_depth, seed, i, li, tries, pp, pd, pc, my_traces, wn, wp);
    }
    // the slot's lanes are adjacent: lowest successful try wins
    const unsigned sm = __ballot_sync(0xffffffffu, stored);
    const unsigned gshift = lane & ~(unsigned)(width - 1), gmask = width == 32 ? 0xffffffffu : ((1u << width) - 1u);
    const unsigned gbits = (sm >> gshift) & gmask;
    const int winner = gbits ? __ffs(gbits) - 1 : -1;
    if (stored && k == winner) { double* out = O.ph + 9 * (size_t)slot; st3(out, pp); st3(out + 3, pd); st3(out + 6, pc); O.stored[slot] = 1; }
    const bool counted = live && (winner < 0 || k <= winner);       // the tries the sequential loop would have made
    const bool retry = j < n_active && k == 0 && winner < 0 && base_try + width < 500;
    const unsigned rm = __ballot_sync(0xffffffffu, retry);
    uint32_t base = 0;
    if (lane == 0 && rm) base = atomicAdd(n_out, (uint32_t)__popc(rm));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (retry) out_list[base + __popc(rm & ((1u << lane) - 1u))] = slot;
    unsigned long long my_tries = counted ? 1ull : 0ull;
    if (!counted) { my_traces = 0; wn = 0; wp = 0; }
    for (int o = 16; o > 0; o >>= 1) { my_tries += __shfl_down_sync(0xffffffffu, my_tries, o); my_traces += __shfl_down_sync(0xffffffffu, my_traces, o); }
    if (lane == 0 && my_tries) { atomicAdd(O.tries, my_tries); atomicAdd(O.traces, my_traces); }
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

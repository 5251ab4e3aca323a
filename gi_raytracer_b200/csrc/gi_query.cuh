// gi_query.cuh — batch forms of the scene-API queries that the reference exposes next to its renderer: Octree::intersect /
// intersectSorted (octree.cpp:150-211, 256-313), PhotonMap::getInRange (photonMap.cpp:50-92, 115-134) and Entity::intersect
// (entities.h:60-101, 158-258, 443-490).  RayTracer::trace / visible / samplePhotons are built on them in the reference; here the
// renderer has its own fused kernels, and these entry points exist so that the preserved C++ members (csrc/host) keep answering — on
// the device, one thread (or warp) per query — and so that each of them can be compared with the reference in isolation.
#pragma once
#include "gi_kernels.cuh"

#define GI_Q_STACK 256   // DFS with every existing child pushed: <= 7 per level + 1

// BoundingBox::intersectSimple (bbox.h:117-138)
__device__ __forceinline__ bool box_hit_simple(const double* bmin, const double* bmax, const DRay& r, double tmin, double tmax)
{
    const double o[3] = { r.o.x, r.o.y, r.o.z }, inv[3] = { r.inv.x, r.inv.y, r.inv.z };
#pragma unroll
    for (int i = 0; i < 3; i++) {
        double t0 = (bmin[i] - o[i]) * inv[i], t1 = (bmax[i] - o[i]) * inv[i];
        if (inv[i] < 0.0) { const double tmp = t0; t0 = t1; t1 = tmp; }
        tmin = t0 > tmin ? t0 : tmin;
        tmax = t1 < tmax ? t1 : tmax;
        if (tmax <= tmin) return false;
    }
    return true;
}
// BoundingBox::intersect(ray, tmin, tmax, t0, t1) (bbox.h:47-73) with the caller's tmin (box_entry's -1 marker needs tmin >= 0)
__device__ __forceinline__ bool box_hit_entry(const double* bmin, const double* bmax, const DRay& r, double tmin, double tmax, double& t0_out)
{
    const double o[3] = { r.o.x, r.o.y, r.o.z }, inv[3] = { r.inv.x, r.inv.y, r.inv.z };
#pragma unroll
    for (int i = 0; i < 3; i++) {
        double t0 = (bmin[i] - o[i]) * inv[i], t1 = (bmax[i] - o[i]) * inv[i];
        if (inv[i] < 0.0) { const double tmp = t0; t0 = t1; t1 = tmp; }
        tmin = t0 > tmin ? t0 : tmin;
        tmax = t1 < tmax ? t1 : tmax;
        if (tmax <= tmin) return false;
    }
    t0_out = tmin;
    return true;
}

// Octree::intersect: the entities of every non-empty leaf the segment [tmin, tmax] meets, in the recursion's order (children 0..7,
// stored order inside a leaf, a primitive once per leaf it sits in).  ids [n][cap]; counts = the full count (may exceed cap).
__global__ void k_octree_intersect(DScene S, size_t n, const double* __restrict__ org, const double* __restrict__ dir, const double* __restrict__ tmin, const double* __restrict__ tmax,
                                   uint32_t cap, uint32_t* __restrict__ ids, uint32_t* __restrict__ counts)
{
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const DRay r = ray_as_stored(ld3(org + 3 * i), ld3(dir + 3 * i));
    uint32_t stack[GI_Q_STACK];
    int sp = 0;
    uint32_t cnt = 0;
    if (S.n_nodes) stack[sp++] = 0;
    while (sp > 0) {
        const DNode nd = load_node(S.nodes, stack[--sp]);
        if (!box_hit_simple(nd.bmin, nd.bmax, r, tmin[i], tmax[i])) continue;
        if (nd.mask == 0) {
            for (uint32_t k = 0; k < nd.prim_cnt; k++, cnt++) if (cnt < cap) ids[i * cap + cnt] = S.refs[nd.prim_off + k].prim;
        } else {
            const int nc = __popc(nd.mask);
            for (int c = nc - 1; c >= 0; c--) if (sp < GI_Q_STACK) stack[sp++] = nd.child + c; else *reinterpret_cast<volatile uint32_t*>(S.err) = GI_DEV_ERR_STACK;
        }
    }
    counts[i] = cnt;
}

// Octree::intersectSorted: the non-empty leaves the ray meets in [tmin, tmax] with their entry distance, kept sorted the reference's way
// (std::partition_point on `t0 >= key`: a new leaf goes behind every entry that is not later, octree.cpp:297-300).
__global__ void k_octree_intersect_sorted(DScene S, size_t n, const double* __restrict__ org, const double* __restrict__ dir, const double* __restrict__ tmin,
                                          const double* __restrict__ tmax, uint32_t cap, uint32_t* __restrict__ nodes, double* __restrict__ t0s, uint32_t* __restrict__ counts)
{
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const DRay r = ray_as_stored(ld3(org + 3 * i), ld3(dir + 3 * i));
    uint32_t stack[GI_Q_STACK];
    int sp = 0;
    uint32_t cnt = 0;     // entries held (<= cap)
    uint32_t total = 0;   // leaves met
    uint32_t* on = nodes + i * cap; double* ot = t0s + i * cap;
    if (S.n_nodes) stack[sp++] = 0;
    while (sp > 0) {
        const uint32_t ni = stack[--sp];
        const DNode nd = load_node(S.nodes, ni);
        double t0;
        if (!box_hit_entry(nd.bmin, nd.bmax, r, tmin[i], tmax[i], t0)) continue;
        if (nd.mask == 0) {
            if (nd.prim_cnt == 0) continue;
            total++;
            uint32_t pos = cnt;                                  // partition point: first entry with key > t0
            while (pos > 0 && !(t0 >= ot[pos - 1])) pos--;
            if (pos >= cap) continue;                            // beyond what the caller can hold
            const uint32_t last = cnt < cap ? cnt : cap - 1;
            for (uint32_t k = last; k > pos; k--) { on[k] = on[k - 1]; ot[k] = ot[k - 1]; }
            on[pos] = ni; ot[pos] = t0;
            if (cnt < cap) cnt++;
        } else {
            const int nc = __popc(nd.mask);
            for (int c = nc - 1; c >= 0; c--) if (sp < GI_Q_STACK) stack[sp++] = nd.child + c; else *reinterpret_cast<volatile uint32_t*>(S.err) = GI_DEV_ERR_STACK;
        }
    }
    counts[i] = total;
}

// PhotonMap::getInRange: the photons of the leaf that contains pos (getBounds) and of every leaf whose closed box touches that leaf's
// box grown by EPSILON, in Node::get's order (children 0..7, insertion order inside a leaf).  One warp per query; ids = original
// photon indices [n][cap]; counts = the full count.
__global__ void __launch_bounds__(GI_WPB * 32) k_photon_in_range(DGatherMap M, size_t n, const double* __restrict__ pos, uint32_t cap, uint32_t* __restrict__ ids, uint32_t* __restrict__ counts,
                                                                uint32_t* __restrict__ overflow)
{
    __shared__ uint32_t s_stack[GI_WPB][GI_GATHER_STACK];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const size_t i = blockIdx.x * (size_t)GI_WPB + wib;
    if (i >= n) return;
    uint32_t leaf, depth;
    if (!pm_find_leaf(M, ld3(pos + 3 * i), lane, leaf, depth)) { if (lane == 0) counts[i] = 0; return; }
    const DNode ln = load_node(M.nodes, leaf);
    double qmin[3], qmax[3];
    for (int a = 0; a < 3; a++) { qmin[a] = ln.bmin[a] - GI_D_EPSILON; qmax[a] = ln.bmax[a] + GI_D_EPSILON; }   // photonMap.cpp:119
    uint32_t total = 0;
    int sp = 0;
    if ((qmax[0] - qmin[0]) > 0) { if (lane == 0) s_stack[wib][0] = 0; sp = 1; }   // photonMap.cpp:73
    __syncwarp();
    while (sp > 0) {
        const uint32_t ni = s_stack[wib][sp - 1]; sp--;
        __syncwarp();
        const DNode cur = load_node(M.nodes, ni);
        if (cur.mask == 0) {
            for (uint32_t b = lane; b < cur.prim_cnt; b += 32) if (total + b < cap) ids[i * cap + total + b] = __ldg(M.pid + cur.prim_off + b);
            total += cur.prim_cnt;
        } else {
            bool ov = false;
            if (lane < 8) {
                const DNode ch = load_node(M.nodes, cur.child + lane);
                ov = (ch.bmin[0] <= qmax[0] && ch.bmax[0] >= qmin[0]) && (ch.bmin[1] <= qmax[1] && ch.bmax[1] >= qmin[1]) && (ch.bmin[2] <= qmax[2] && ch.bmax[2] >= qmin[2]);
            }
            const uint32_t b = __ballot_sync(0xffffffffu, ov) & 0xffu;
            if (lane < 8 && ov) { const int rank = __popc(b >> (lane + 1)); if (sp + rank < GI_GATHER_STACK) s_stack[wib][sp + rank] = cur.child + lane; }
            sp += __popc(b);
            if (sp > GI_GATHER_STACK) { sp = GI_GATHER_STACK; if (lane == 0) atomicExch(overflow, 1u); }
            __syncwarp();
        }
    }
    if (lane == 0) counts[i] = total;
}

// Entity::intersect(ray, hit, norm, uv): one (primitive, ray) pair per thread.  hit / normal as the reference returns them (the
// barycentric normal un-normalised); uv is written only where the reference writes it (wrote_uv), a cone or a triangle without vertex
// normals leaves the caller's value alone (entities.h:480-487).
__global__ void k_prim_intersect(DScene S, size_t n, const uint32_t* __restrict__ prim, const double* __restrict__ org, const double* __restrict__ dir, uint8_t* __restrict__ ok,
                                 double* __restrict__ hit, double* __restrict__ normal, double* __restrict__ uv, uint8_t* __restrict__ wrote_uv)
{
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t p = prim[i];
    ok[i] = 0; wrote_uv[i] = 0;
    st3(hit + 3 * i, mk3(0, 0, 0)); st3(normal + 3 * i, mk3(0, 0, 0)); uv[2 * i] = 0; uv[2 * i + 1] = 0;
    if (p >= S.n_prims) return;
    const DRay r = ray_as_stored(ld3(org + 3 * i), ld3(dir + 3 * i));
    const double* gsrc = S.prim_geom + 9 * (size_t)p;
    double g[9];
    for (int k = 0; k < 9; k++) g[k] = gsrc[k];
    const uint32_t kind = S.prim_type[p];
    DHit h; h.prim = p; h.t = 0; h.u = 0; h.v = 0; h.tu = 0; h.tv = 0; h.n = mk3(0, 0, 0);
    bool good;
    if (kind == GI_PRIM_TRIANGLE) {
        for (int k = 0; k < 3; k++) { g[3 + k] = g[3 + k] - g[k]; g[6 + k] = g[6 + k] - g[k]; }   // v0, edge1, edge2: what tri_hit expects
        h.t = tri_hit(g, r, h.u, h.v); good = h.t > 0;
    } else if (kind == GI_PRIM_SPHERE) good = sphere_hit(g, r, h.t);
    else good = cone_hit(g, S.prim_nrm + 9 * (size_t)p, r, h.t, h.n);
    if (!good) return;
    d3 hp, hn; double tu, tv;
    hit_surface(S, r, h, false, hp, hn, tu, tv);
    ok[i] = 1;
    st3(hit + 3 * i, hp); st3(normal + 3 * i, hn);
    bool w = kind == GI_PRIM_SPHERE;
    if (kind == GI_PRIM_TRIANGLE) {
        const double* nn = S.prim_nrm + 9 * (size_t)p;
        w = len2(ld3(nn)) > 0 && len2(ld3(nn + 3)) > 0 && len2(ld3(nn + 6)) > 0;
    }
    if (w) { uv[2 * i] = tu; uv[2 * i + 1] = tv; wrote_uv[i] = 1; }
}

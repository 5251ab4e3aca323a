// gi_device.cuh — device-side building blocks of the B200 (sm_100a) hot path: fp64 vector math in the reference's
// operation order, the counter PRNG, Halton sampling, slab / primitive tests, materials, samplers and the two octree
// traversals.  Compiled with -fmad=false: no FMA contraction anywhere, so every fp64 result that the reference
// computes without libm is reproduced bit-for-bit (SURVEY §7 "hard parts").  Reference file:line cited per function.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/gi_api.h"

#define GI_D_EPSILON 0.00001     // util.h:18
#define GI_D_SHADOW_BIAS 0.0001  // util.h:20
#define GI_D_PI 3.14159265358979323846
#define GI_STACK_MAX 96          // >= 3*depth+8 entries; octree depth is bounded by MIN_LEAF_SIZE (<= ~20 levels)

// ---- device scene ------------------------------------------------------------------------------------------------------
struct __align__(16) DNode {   // 64 B: one record per octree node (SURVEY §8d bytes model: 6 x fp64 box + 16 B topology)
    double bmin[3], bmax[3];
    uint32_t child;            // first existing child (children contiguous, reference child order)
    uint32_t prim_off;         // first leaf reference
    uint32_t prim_cnt;
    uint32_t mask;             // bit i = child i exists; 0 = leaf
};
struct __align__(16) DLeafRef {  // 80 B: one record per (leaf, primitive) occurrence, stored in leaf order
    double g[9];               // triangle v0,v1,v2 | sphere c,r | cone pos,rad,height
    uint32_t prim;             // primitive id (insertion order)
    uint32_t flags;            // bits 0-1 kind, bit 2 writes uv, bit 3 needs the stochastic alpha test
};
#define LF_KIND(f) ((f)&3u)
#define LF_WRITES_UV 4u
#define LF_ALPHA 8u

struct DHaltonDim { uint32_t base, block, nblocks, table_off; float scale; };

struct DScene {
    const DNode* nodes;
    const DLeafRef* refs;
    uint32_t n_nodes, n_refs, n_prims;
    const double* prim_geom;   // [n][9]
    const double* prim_nrm;    // [n][9]
    const double* prim_uv;     // [n][6]
    const double* prim_fnorm;  // [n][3]
    const uint32_t* prim_mat;
    const uint8_t* prim_type;
    const gi_material* mats;
    const gi_texture* tex;
    const uint8_t* tex_pixels;
    const gi_light* lights;
    uint32_t n_lights, n_mats, n_tex;
    gi_camera cam;
    double ambient[3];
    const uint16_t* halton_tab;
    const DHaltonDim* halton_dims;   // [256]
    uint32_t full;                   // scene needs the FULL traversal (alpha materials or primitives that do not write uv)
    uint32_t implicit_boxes;         // every child box equals the partition formula of its parent box (checked at upload)
    uint32_t n_fog;                  // Octree::at: HeightFog volumes (atmosphere.h), in push_back order
    const gi_fog* fogs;
    const double* fog_grid;          // noise grids of all volumes, concatenated
    uint32_t* err;                   // sticky device error word (bit 0: a traversal stack overflowed) — read back at every synchronisation
};
#define GI_DEV_ERR_STACK 1u

// ---- fp64 vectors in glm's evaluation order -------------------------------------------------------------
struct d3 { double x, y, z; };
__device__ __forceinline__ d3 mk3(double x, double y, double z) { d3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ d3 ld3(const double* p) { return mk3(p[0], p[1], p[2]); }
__device__ __forceinline__ void st3(double* p, d3 a) { p[0] = a.x; p[1] = a.y; p[2] = a.z; }
__device__ __forceinline__ d3 operator+(d3 a, d3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ d3 operator-(d3 a, d3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ d3 operator*(d3 a, d3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ d3 operator*(d3 a, double s) { return mk3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ double dot3(d3 a, d3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ d3 cross3(d3 x, d3 y) { return mk3(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y); }
__device__ __forceinline__ d3 normalize3(d3 a) { return a * (1.0 / sqrt(dot3(a, a))); }
__device__ __forceinline__ d3 reflect3(d3 I, d3 N) { double d = dot3(N, I); return I - (N * d) * 2.0; }
__device__ __forceinline__ double len2(d3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }  // vecLengthSquared, util.h:35-38

struct DRay { d3 o, d, inv; };
// Ray::Ray -> setDir (ray.h:7-17)
__device__ __forceinline__ DRay make_ray(d3 o, d3 dir_raw)
{
    DRay r; r.o = o; r.d = normalize3(dir_raw);
    r.inv = mk3(1.0 / r.d.x, 1.0 / r.d.y, 1.0 / r.d.z);
    return r;
}
__device__ __forceinline__ DRay ray_as_stored(d3 o, d3 dir)
{
    DRay r; r.o = o; r.d = dir;
    r.inv = mk3(1.0 / dir.x, 1.0 / dir.y, 1.0 / dir.z);
    return r;
}

// ---- counter PRNG (the specification the CPU checker restates too; replaces util.h:52-80) ----------------------------
__host__ __device__ __forceinline__ uint64_t gi_mix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__host__ __device__ __forceinline__ double gi_rand(uint64_t seed, uint64_t path, uint64_t depth, uint64_t site)
{
    uint64_t h = gi_mix64(seed ^ gi_mix64(path ^ gi_mix64(depth ^ gi_mix64(site))));
    return (double)(h >> 11) * (1.0 / 9007199254740992.0);
}
#define SITE_LIGHT_U 0ull
#define SITE_LIGHT_V 1ull
#define SITE_TYPE_A 2ull
#define SITE_TYPE_B 3ull
#define SITE_RR 4ull
#define SITE_ALPHA_TRACE 5ull
#define SITE_ALPHA_SHADOW 6ull
#define SITE_PH_DIR_U 7ull
#define SITE_PH_DIR_V 8ull
#define SITE_PH_SEC_U 9ull
#define SITE_PH_SEC_V 10ull
#define SITE_FOG_RAD 11ull      // raymarch draws: counter = step (radiance), light << 32 | step (visible), step (photons)
#define SITE_FOG_SHADOW 12ull
#define SITE_FOG_PHOTON 13ull
#define SITE_PH_FOG_U 14ull
#define SITE_PH_FOG_V 15ull
#define SITE(s, c) (((uint64_t)(s) << 56) | (uint64_t)(c))
#define PHOTON_PATH_BIT (1ull << 63)

// ---- Halton (halton_sampler.h:626-888, 1417-3286) -----------------------------------------------------------------------
__device__ __forceinline__ float halton_sample(const DScene& S, uint32_t dim, uint32_t index)
{
    if (dim == 0) {  // halton2: bit reversal into the mantissa (halton_sampler.h:1417-1431)
        uint32_t rev = __brev(index);
        return __uint_as_float(0x3f800000u | (rev >> 9)) - 1.f;
    }
    if (dim >= 256) return 0.f;
    DHaltonDim D = S.halton_dims[dim];
    const uint16_t* T = S.halton_tab + D.table_off;
    uint32_t sum = 0, idx = index;
    // digit block j (least significant first) is weighted by block^(n-1-j): accumulate with Horner from the top instead
    // of the reference's explicit constants — identical u32 arithmetic (no overflow: the total is < 2^32)
    uint32_t blk[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        if (j < (int)D.nblocks) { blk[j] = T[idx % D.block]; idx /= D.block; } else blk[j] = 0;
    }
#pragma unroll
    for (int j = 0; j < 8; j++)
        if (j < (int)D.nblocks) sum = sum * D.block + blk[j];
    return __fmul_rn(__uint2float_rn(sum), D.scale);
}

// Halton_enum (halton_enum.h:69-155); the constants are computed on the host
struct DHEnum { uint32_t p2, p3, mx, my, inc; float scale_x, scale_y; };
__device__ __forceinline__ uint32_t henum_index(const DHEnum& he, uint32_t s, uint32_t x, uint32_t y)
{
    uint64_t hx = he.p2 ? (uint64_t)(__brev(x) >> (32 - he.p2)) : 0ull;   // halton2_inverse (halton_enum.h:136-144)
    uint32_t r3 = 0, yy = y;
    for (uint32_t d = 0; d < he.p3; ++d) { r3 = r3 * 3 + yy % 3; yy /= 3; }  // halton3_inverse (halton_enum.h:146-155)
    uint64_t hy = r3;
    uint32_t offset = (uint32_t)((hx * he.mx + hy * he.my) % he.inc);
    return offset + s * he.inc;  // u32 wrap-around kept
}

// ---- boxes (bbox.h) ----------------------------------------------------------------------------------------------------------
// BoundingBox::intersect(ray, tmin, tmax, t0, t1) (bbox.h:47-73).  Returns entry t (>= tmin) or -1 when rejected.
__device__ __forceinline__ double box_entry(const double* bmin, const double* bmax, const DRay& r, double tmin, double tmax)
{
    {
        double t0 = (bmin[0] - r.o.x) * r.inv.x, t1 = (bmax[0] - r.o.x) * r.inv.x;
        if (r.inv.x < 0.0) { double tmp = t0; t0 = t1; t1 = tmp; }
        tmin = t0 > tmin ? t0 : tmin; tmax = t1 < tmax ? t1 : tmax;
        if (tmax <= tmin) return -1.0;
    }
    {
        double t0 = (bmin[1] - r.o.y) * r.inv.y, t1 = (bmax[1] - r.o.y) * r.inv.y;
        if (r.inv.y < 0.0) { double tmp = t0; t0 = t1; t1 = tmp; }
        tmin = t0 > tmin ? t0 : tmin; tmax = t1 < tmax ? t1 : tmax;
        if (tmax <= tmin) return -1.0;
    }
    {
        double t0 = (bmin[2] - r.o.z) * r.inv.z, t1 = (bmax[2] - r.o.z) * r.inv.z;
        if (r.inv.z < 0.0) { double tmp = t0; t0 = t1; t1 = tmp; }
        tmin = t0 > tmin ? t0 : tmin; tmax = t1 < tmax ? t1 : tmax;
        if (tmax <= tmin) return -1.0;
    }
    return tmin;
}
// BoundingBox::contains (bbox.h:41-44), half-open
__device__ __forceinline__ bool box_contains(const double* bmin, const double* bmax, d3 p)
{
    return p.x >= bmin[0] && p.y >= bmin[1] && p.z >= bmin[2] && p.x < bmax[0] && p.y < bmax[1] && p.z < bmax[2];
}

// ---- implicit child boxes ----------------------------------------------------------------------------------------------------
// Octree::Node::partition (octree.cpp:318-328) derives the eight child boxes from the parent box alone: per axis the planes
// are P0 = min, P1 = mid = min + .5*(max-min) (bit-identical to min + .5*d), P2 = mid + .5*d and P3 = max; the lower half is
// [P0,P1], the upper half [P1,P2], and child 7 is [mid,max] = [P1,P3] on every axis.  When the uploaded tree obeys this
// (verified bit-for-bit in gi_scene_upload) a traversal step needs only the parent's record: the slab parameter of each of
// the 12 planes is computed once — the same (plane - o) * inv the reference computes per child — and shared by the children,
// instead of loading eight 64-byte child records.
struct PlaneT { double x[4], y[4], z[4]; };
__device__ __forceinline__ void plane_params(const DNode& nd, const DRay& r, PlaneT& T)
{
    double mx = nd.bmin[0] + .5 * (nd.bmax[0] - nd.bmin[0]), my = nd.bmin[1] + .5 * (nd.bmax[1] - nd.bmin[1]), mz = nd.bmin[2] + .5 * (nd.bmax[2] - nd.bmin[2]);
    double hx = .5 * (nd.bmax[0] - nd.bmin[0]), hy = .5 * (nd.bmax[1] - nd.bmin[1]), hz = .5 * (nd.bmax[2] - nd.bmin[2]);
    T.x[0] = (nd.bmin[0] - r.o.x) * r.inv.x; T.x[1] = (mx - r.o.x) * r.inv.x; T.x[2] = ((mx + hx) - r.o.x) * r.inv.x; T.x[3] = (nd.bmax[0] - r.o.x) * r.inv.x;
    T.y[0] = (nd.bmin[1] - r.o.y) * r.inv.y; T.y[1] = (my - r.o.y) * r.inv.y; T.y[2] = ((my + hy) - r.o.y) * r.inv.y; T.y[3] = (nd.bmax[1] - r.o.y) * r.inv.y;
    T.z[0] = (nd.bmin[2] - r.o.z) * r.inv.z; T.z[1] = (mz - r.o.z) * r.inv.z; T.z[2] = ((mz + hz) - r.o.z) * r.inv.z; T.z[3] = (nd.bmax[2] - r.o.z) * r.inv.z;
}
// slab test of child i (bit0 = +x, bit1 = +z, bit2 = +y) from the shared plane parameters; same comparisons as bbox.h:47-73
// (the per-axis early rejects of the reference are equivalent to one final test: tmin only grows, tmax only shrinks)
__device__ __forceinline__ double child_entry(const PlaneT& T, int i, const DRay& r, double tmin, double tmax)
{
    const bool ux = i & 1, uz = i & 2, uy = i & 4, last = i == 7;
    double a0 = ux ? T.x[1] : T.x[0], a1 = ux ? (last ? T.x[3] : T.x[2]) : T.x[1];
    double b0 = uy ? T.y[1] : T.y[0], b1 = uy ? (last ? T.y[3] : T.y[2]) : T.y[1];
    double c0 = uz ? T.z[1] : T.z[0], c1 = uz ? (last ? T.z[3] : T.z[2]) : T.z[1];
    if (r.inv.x < 0.0) { double t = a0; a0 = a1; a1 = t; }
    if (r.inv.y < 0.0) { double t = b0; b0 = b1; b1 = t; }
    if (r.inv.z < 0.0) { double t = c0; c0 = c1; c1 = t; }
    tmin = a0 > tmin ? a0 : tmin; tmax = a1 < tmax ? a1 : tmax;
    if (tmax <= tmin) return -1.0;
    tmin = b0 > tmin ? b0 : tmin; tmax = b1 < tmax ? b1 : tmax;
    if (tmax <= tmin) return -1.0;
    tmin = c0 > tmin ? c0 : tmin; tmax = c1 < tmax ? c1 : tmax;
    if (tmax <= tmin) return -1.0;
    return tmin;
}

// entry distances of all eight implicit children at once.  The reference folds, per child, tmin = ((tmin0 (+) x) (+) y) (+) z with
// a (+) b = b > a ? b : a, and tmax likewise with b < a (bbox.h:47-73; the per-axis early rejects equal one final test, see
// child_entry).  Children share those prefixes: the near/far plane of each half is picked once per axis by the sign of the
// inverse direction, the x-then-y prefix exists for 4 combinations, the full fold for 8 — the same operations in the same
// order as eight separate slab tests, at a third of the instructions.  Child 7 = [mid, max] has its own far (or, for a
// negative direction, near) planes.  t0c[i] = entry distance, or -1 when the child is absent or missed.
#define GI_UP(a, b) ((b) > (a) ? (b) : (a))
#define GI_DN(a, b) ((b) < (a) ? (b) : (a))
// Returns the mask of children that exist and are hit; t0c[i] = entry distance of child i (meaningful for the hit ones).
__device__ __forceinline__ uint32_t children_entry(const DNode& nd, const DRay& r, double tmin0, double tmax0, double (&t0c)[8])
{
    PlaneT T;
    plane_params(nd, r, T);
    const bool sx = r.inv.x < 0.0, sy = r.inv.y < 0.0, sz = r.inv.z < 0.0;
    // [0] lower half, [1] upper half, [2] child 7
    const double nx[3] = { sx ? T.x[1] : T.x[0], sx ? T.x[2] : T.x[1], sx ? T.x[3] : T.x[1] }, fx[3] = { sx ? T.x[0] : T.x[1], sx ? T.x[1] : T.x[2], sx ? T.x[1] : T.x[3] };
    const double ny[3] = { sy ? T.y[1] : T.y[0], sy ? T.y[2] : T.y[1], sy ? T.y[3] : T.y[1] }, fy[3] = { sy ? T.y[0] : T.y[1], sy ? T.y[1] : T.y[2], sy ? T.y[1] : T.y[3] };
    const double nz[3] = { sz ? T.z[1] : T.z[0], sz ? T.z[2] : T.z[1], sz ? T.z[3] : T.z[1] }, fz[3] = { sz ? T.z[0] : T.z[1], sz ? T.z[1] : T.z[2], sz ? T.z[1] : T.z[3] };
    double ex[2], lx[2];   // after the x axis
#pragma unroll
    for (int a = 0; a < 2; a++) { ex[a] = GI_UP(tmin0, nx[a]); lx[a] = GI_DN(tmax0, fx[a]); }
    double exy[2][2], lxy[2][2];   // [x half][y half]
#pragma unroll
    for (int a = 0; a < 2; a++)
#pragma unroll
        for (int b = 0; b < 2; b++) { exy[a][b] = GI_UP(ex[a], ny[b]); lxy[a][b] = GI_DN(lx[a], fy[b]); }
    uint32_t miss = 0;
#pragma unroll
    for (int i = 0; i < 7; i++) {   // bit0 = +x, bit1 = +z, bit2 = +y
        const int bx = i & 1, bz = (i >> 1) & 1, by = (i >> 2) & 1;
        const double en = GI_UP(exy[bx][by], nz[bz]), lv = GI_DN(lxy[bx][by], fz[bz]);
        t0c[i] = en;
        miss |= (lv <= en ? 1u : 0u) << i;
    }
    {
        const double en = GI_UP(GI_UP(GI_UP(tmin0, nx[2]), ny[2]), nz[2]), lv = GI_DN(GI_DN(GI_DN(tmax0, fx[2]), fy[2]), fz[2]);
        t0c[7] = en;
        miss |= (lv <= en ? 1u : 0u) << 7;
    }
    return nd.mask & ~miss;
}

// ---- one interior step -------------------------------------------------------------------------------------------------------------
// (Measured and dropped in round 2: deriving the <= 4 children a ray can meet as the SEQUENCE of octants along it — sort the three
//  mid-plane crossings, test the four segments — instead of slab-testing all eight children and ranking them.  It is exact and needs
//  fewer fp64 comparisons on paper, but fp64 min / max / select are 3-4 instructions each on sm_100 and the computation is one long
//  dependent chain, where the eight independent slab tests + integer ranking below overlap: closest-hit kernels ran 30-50 % SLOWER
//  (profiles/r02/README.md, variants v4 / v20 against v21).)
template <bool IMPL, bool ORDERED>
__device__ __forceinline__ uint32_t children_general(const DNode* __restrict__ nodes, const DNode& nd, const DRay& r, double tmax0, uint32_t& seq, double mark_t, uint32_t& n_near);

// ---- per-thread traversal stack: GI_STACK_MAX node indices in local memory (its top lines stay in L1; a shared-memory stack was
// measured 5-8 % slower: it shrinks the L1 the node and leaf records live in).  An overflow raises the sticky error word — a plain
// store of a constant, every writer writes the same bit: an atomicOr here (warp-aggregated REDUX + elected REDG) in the middle of the
// hottest loop of every traversal kernel cost 40-60 % of their speed even though it never executed.
struct TStack { uint32_t e[GI_STACK_MAX]; };
__device__ __forceinline__ void stack_push(TStack& st, int& sp, uint32_t v, uint32_t* err)
{
    if (sp < GI_STACK_MAX) st.e[sp++] = v;
    else *reinterpret_cast<volatile uint32_t*>(err) = GI_DEV_ERR_STACK;
}
__device__ __forceinline__ uint32_t stack_pop(TStack& st, int& sp) { return st.e[--sp]; }
#ifndef GI_BLOCK
#define GI_BLOCK 128
#endif

// (Measured and dropped in round 2: a per-ray "mailbox" of the primitives it has already missed — the reference's octree stores a
//  triangle in every leaf it overlaps, ~45-50 references per triangle in the foliage / atrium stand-ins, and a ray re-tests it in each —
//  as an 8-entry ring in shared memory looked up before the fp64 test.  Even there the look-ups cost more than the skipped tests saved:
//  frames 8 % slower on both stand-ins, 2-3 % on caustics / glass; profiles/r02/README.md.)
#define GI_TSTACK_DECL(name) TStack name

__device__ __forceinline__ DNode load_node(const DNode* nodes, uint32_t i)
{
    // 4 x 16-byte vector loads through the read-only path
    const double2* p = reinterpret_cast<const double2*>(nodes + i);
    double2 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
    uint4 t = __ldg(reinterpret_cast<const uint4*>(p + 3));
    DNode n;
    n.bmin[0] = a.x; n.bmin[1] = a.y; n.bmin[2] = b.x; n.bmax[0] = b.y; n.bmax[1] = c.x; n.bmax[2] = c.y;
    n.child = t.x; n.prim_off = t.y; n.prim_cnt = t.z; n.mask = t.w;
    return n;
}


// n_near = how many of the hit children are entered before mark_t (they come first in `seq`; ORDERED only)
template <bool IMPL, bool ORDERED>
__device__ __forceinline__ uint32_t children_general(const DNode* __restrict__ nodes, const DNode& nd, const DRay& r, double tmax0, uint32_t& seq, double mark_t, uint32_t& n_near)
{
    double t0c[8];
    uint32_t hm = 0;
    if (IMPL) hm = children_entry(nd, r, 0.0, tmax0, t0c);
    else {
        uint32_t c = nd.child;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            t0c[i] = -1.0;
            if (nd.mask & (1u << i)) {
                DNode ch = load_node(nodes, c);
                t0c[i] = box_entry(ch.bmin, ch.bmax, r, 0.0, tmax0);
                if (t0c[i] >= 0.0) hm |= 1u << i;
                c++;
            }
        }
    }
    seq = 0; n_near = 0;
    if (hm == 0) return 0;
    if (!ORDERED) {   // any-hit: visiting order is free
        uint32_t n = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) if ((hm >> i) & 1u) { seq |= (uint32_t)i << (4 * n); n++; }
        return n;
    }
    // ascending (entry distance, child index) — the reference's sorted leaf order (octree.cpp:297-300) — by RANKING:
    // entry distances are >= +0, so the raw 64-bit patterns compare like the values; a child that is not hit gets the sign bit and
    // ranks behind every hit.  rank(i) = number of children before child i: 28 integer comparisons, no double is moved.
    unsigned long long key[8];
#pragma unroll
    for (int i = 0; i < 8; i++) key[i] = (unsigned long long)__double_as_longlong(t0c[i]) | ((unsigned long long)((~hm >> i) & 1u) << 63);
    uint32_t rank = 0;   // 4 bits per child
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = i + 1; j < 8; j++) rank += key[j] < key[i] ? (1u << (4 * i)) : (1u << (4 * j));   // ties: the lower index comes first
#pragma unroll
    for (int i = 0; i < 8; i++) seq |= (uint32_t)i << (((rank >> (4 * i)) & 15u) * 4u);
    if (mark_t < CUDART_INF) {
        uint32_t far = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) far |= (t0c[i] >= mark_t ? 1u : 0u) << i;
        n_near = (uint32_t)__popc(hm & ~far);
    } else n_near = (uint32_t)__popc(hm);
    return (uint32_t)__popc(hm);
}
// one interior step: the hit children of nd in visiting order (4 bits each in seq), their number returned
template <bool IMPL, bool ORDERED>
__device__ __forceinline__ uint32_t interior_step(const DNode* __restrict__ nodes, const DNode& nd, const DRay& r, double tmax0, uint32_t& seq)
{
    uint32_t n_near;
    return children_general<IMPL, ORDERED>(nodes, nd, r, tmax0, seq, CUDART_INF, n_near);
}
template <bool IMPL>
__device__ __forceinline__ uint32_t interior_step_marked(const DNode* __restrict__ nodes, const DNode& nd, const DRay& r, uint32_t& seq, double mark_t, uint32_t& n_near)
{
    return children_general<IMPL, true>(nodes, nd, r, CUDART_INF, seq, mark_t, n_near);
}
__device__ __forceinline__ uint32_t child_node(const DNode& nd, uint32_t c) { return nd.child + __popc(nd.mask & ((1u << c) - 1u)); }

// ---- primitives (entities.h) ---------------------------------------------------------------------------------------------
// triangle::intersect, geometric part (entities.h:443-478): returns t > 0 or -1; u, v barycentrics
__device__ __forceinline__ double tri_hit(const double* g, const DRay& r, double& uo, double& vo)
{
    // a triangle's leaf record holds v0, v1 - v0, v2 - v0: the edges the reference recomputes per test (entities.h:447-448),
    // subtracted once at upload in the same fp64 arithmetic
    d3 v0 = mk3(g[0], g[1], g[2]);
    d3 edge1 = mk3(g[3], g[4], g[5]), edge2 = mk3(g[6], g[7], g[8]);
    d3 p = cross3(r.d, edge2);
    double det = dot3(edge1, p);
    if (det < GI_D_EPSILON && det > -GI_D_EPSILON) return -1.0;
    d3 tvec = r.o - v0;
    const double a = dot3(tvec, p);
    // u = a * fl(1 / det) (entities.h:455-459) is certainly negative when a and det differ in sign (|a| > 2^-1007 keeps the
    // product away from underflow) and certainly above 1 when |a| exceeds |det| by more than the two roundings can undo
    // ((1 + 2^-49)(1 - 2^-53) > 1 + 2^-50): those candidates are rejected before the fp64 division — it alone was 3-7 % of
    // the traversal kernels' instructions.  Everything else takes the reference's arithmetic.
    {
        const int ha = __double2hiint(a), hd = __double2hiint(det);
        if ((ha ^ hd) < 0 && (ha & 0x7ff00000) > 0x01000000) return -1.0;   // signs differ and |a| > 2^-1007
    }
    if (fabs(a) > fabs(det) * (1.0 + 0x1p-48)) return -1.0;
    double inv_det = 1.0 / det;
    double u = a * inv_det;
    if (u < 0 || u > 1) return -1.0;
    d3 q = cross3(tvec, edge1);
    double v = dot3(r.d, q) * inv_det;
    if (v < 0 || u + v > 1) return -1.0;
    double t = dot3(edge2, q) * inv_det;
    if (t <= 0) return -1.0;
    uo = u; vo = v;
    return t;
}
// sphere::intersect, geometric part (entities.h:60-84): hit point = o + d*t; returns 1 on hit.  pow(x,2) is x*x (what
// GCC emits for the reference at -O2)
__device__ __forceinline__ bool sphere_hit(const double* g, const DRay& r, double& t_out)
{
    d3 pos = mk3(g[0], g[1], g[2]); double rad = g[3];
    d3 oc = r.o - pos;
    double d = dot3(r.d, oc);
    double rr = (d * d - len2(oc) + rad * rad);
    if (rr < 0) return false;
    double sr = sqrt(rr);
    double t_1 = -1 * d - sr, t_2 = -1 * d + sr;
    if (t_1 < 0 && t_2 < 0) return false;
    t_out = ((t_1 < t_2 && t_1 > 0) || t_2 < 0) ? t_1 : t_2;
    return true;
}
// cone::intersect (entities.h:158-258); rot = cone::rot column-major; v*rot is glm's row-vector product
__device__ __forceinline__ d3 vec_mat(d3 v, const double* m)
{
    return mk3(m[0] * v.x + m[1] * v.y + m[2] * v.z, m[3] * v.x + m[4] * v.y + m[5] * v.z, m[6] * v.x + m[7] * v.y + m[8] * v.z);
}
// Out of line (cones are rare): everything goes in and out BY VALUE — a pointer to the caller's geometry registers or ray would force
// them into local memory for every primitive test of every kernel (the profile of the first version showed exactly that:
// ten 8-byte local stores per candidate).
struct ConeHit { double t; d3 n; bool ok; };
__device__ __noinline__ ConeHit cone_hit_v(d3 pos, double rad, double height, const double* __restrict__ rot, d3 ro, d3 rd)
{
    ConeHit H; H.ok = false; H.t = 0; H.n = mk3(0, 0, 0);
    d3 origin = vec_mat(ro - pos, rot), dir = vec_mat(rd, rot);
    double phiMax = 2 * GI_D_PI;
    double kq = (rad / height) * (rad / height);
    double A = dir.x * dir.x + dir.y * dir.y - kq * dir.z * dir.z;
    double B = 2 * (dir.x * origin.x + dir.y * origin.y - kq * dir.z * (origin.z - height));
    double C = origin.x * origin.x + origin.y * origin.y - kq * (origin.z - height) * (origin.z - height);
    double discrim = B * B - 4.0 * A * C;
    if (discrim < 0) return H;
    double rootDiscrim = sqrt(discrim);
    double q = (B < 0) ? -.5 * (B - rootDiscrim) : -.5 * (B + rootDiscrim);
    double t_1 = q / A, t_2 = C / q;
    if (t_1 < 0 && t_2 < 0) return H;
    if (t_1 > t_2) { double tmp = t_1; t_1 = t_2; t_2 = tmp; }
    double thit = t_1;
    if (t_1 < 0) thit = t_2; else if (t_2 < 0) thit = t_1;
    d3 phit = origin + dir * thit;
    double phi = atan2(phit.y, phit.x);
    if (phi < 0.) phi += 2.0 * GI_D_PI;
    if (phit.z < 0 || phit.z > height || phi > phiMax) {
        if (thit == t_2) return H;
        thit = t_2;
        phit = origin + dir * thit;
        phi = atan2(phit.y, phit.x);
        if (phi < 0.) phi += 2.0 * GI_D_PI;
        if (phit.z < 0 || phit.z > height || phi > phiMax) return H;
    }
    double vpar = phit.z / height;
    d3 dpdu = mk3(-phiMax * phit.y, phiMax * phit.x, 0);
    d3 dpdv = mk3(-phit.x / (1.0 - vpar), -phit.y / (1.0 - vpar), height);
    H.n = normalize3(cross3(dpdu, dpdv));
    H.t = thit; H.ok = true;
    return H;
}
__device__ __forceinline__ bool cone_hit(const double* g, const double* rot, const DRay& r, double& t_out, d3& n_out)
{
    const ConeHit H = cone_hit_v(mk3(g[0], g[1], g[2]), g[3], g[4], rot, r.o, r.d);
    t_out = H.t; n_out = H.n;
    return H.ok;
}

// ---- textures / materials (material.h) -----------------------------------------------------------------------------------
__device__ __forceinline__ const uint8_t* tex_pixel(const DScene& S, const gi_texture& t, double u, double v)
{
    int x = abs((int)(u * t.width * t.tile_u) % t.width);                              // material.h:65
    int y = t.height - abs((int)(v * t.height * t.tile_v) % t.height) - 1;
    return S.tex_pixels + t.pixel_offset + ((size_t)y * t.width + x) * 4;
}
__device__ __forceinline__ d3 tex_get(const DScene& S, uint32_t id, double u, double v)
{
    const gi_texture& t = S.tex[id];
    if (t.kind == GI_TEX_CONST) return ld3(t.a);
    if (t.kind == GI_TEX_CHECKER) {
        if ((((int)(u * t.tiles) % 2 == 0) ^ ((int)(v * t.tiles) % 2 == 0))) return ld3(t.a);
        return ld3(t.b);
    }
    const uint8_t* p = tex_pixel(S, t, u, v);
    double e = 1.0 / (1.0 / 2.2);                                                       // gamma(c, 1/GAMMA): material.h:67, util.h:94-97
    return mk3(pow(p[0] / 255.0, e), pow(p[1] / 255.0, e), pow(p[2] / 255.0, e));
}
__device__ __forceinline__ double tex_alpha(const DScene& S, uint32_t id, double u, double v)
{
    const gi_texture& t = S.tex[id];
    if (t.kind != GI_TEX_IMAGE || !t.has_alpha) return 1;
    return tex_pixel(S, t, u, v)[3] / 255.0;
}
// `drand() < material.getAlpha(uv) || IOR != 1` (raytracer.h:455,:297) with an occurrence-keyed counter draw
// `frac` (when given) is set when the candidate was cut out at a FRACTIONAL alpha (0 < a < 1): another occurrence of the same
// primitive in another leaf draws again and may pass — the closest-hit walk must not prune behind such a candidate (trace_walk).
__device__ __forceinline__ bool alpha_pass(const DScene& S, uint32_t prim, uint32_t node, double u, double v, uint64_t seed, uint64_t path, uint64_t depth, uint64_t site, bool* frac = nullptr)
{
    const gi_material& m = S.mats[S.prim_mat[prim]];
    if (m.ior != 1) return true;
    double a = m.opacity * tex_alpha(S, m.diffuse_tex, u, v);
    if (a >= 1.0) return true;
    const bool pass = gi_rand(seed, path, depth, SITE(site, ((uint64_t)node << 28) ^ prim)) < a;
    if (frac) *frac = !pass && a > 0.0;
    return pass;
}

// uv of a uv-writing primitive at barycentrics (u,v) (entities.h:482 for triangles, :93-96 for spheres)
__device__ __forceinline__ void prim_uv_at(const DScene& S, uint32_t prim, d3 hit, double u, double v, double& tu, double& tv)
{
    if (S.prim_type[prim] == GI_PRIM_TRIANGLE) {
        const double* t2 = S.prim_uv + 6 * (size_t)prim;
        double w = 1 - u - v;
        tu = (w * t2[0] + u * t2[2]) + v * t2[4];
        tv = (w * t2[1] + u * t2[3]) + v * t2[5];
    } else {  // sphere
        const double* g = S.prim_geom + 9 * (size_t)prim;
        double rad = g[3];
        d3 dd = mk3((g[0] - hit.x) / rad, (g[1] - hit.y) / rad, (g[2] - hit.z) / rad);
        tv = .5 + asin(dd.y) / GI_D_PI;
        tu = .5 + atan2(dd.z, dd.x) / (2 * GI_D_PI);
    }
}

// ---- closest hit: RayTracer::trace (raytracer.h:382-478) as an ordered stack traversal -----------------------------------------
// Leaves are visited in ascending entry distance, ties in child order (= the reference's sorted leaf list,
// octree.cpp:285-313 + SURVEY §A.3); inside a leaf every primitive is tested in stored order; a candidate is accepted if
// it is the first or STRICTLY closer in |hit-o|^2; the walk stops after the first leaf in which an accepted hit lies
// inside the leaf box (raytracer.h:446-472).  There is deliberately no pruning against the best distance: the reference
// has none, and keeping its exact visiting rule is what makes ids bit-identical.
//
// What the walk does NOT keep is the reference's walking on after the result is final.  Whenever the accepted hit was first met from a
// leaf that does not contain it (a triangle is stored in every leaf it overlaps), the reference never stops: the same hit met again in
// the leaf that does contain it is not STRICTLY closer, so `term` is never set (raytracer.h:457-466) and every remaining leaf along the
// ray is tested — 60-75 % of all node and primitive tests on the foliage / atrium stand-ins.  Nothing met in a node the ray enters at
// t0 >= t_hit can be accepted: |hit - o|^2 is monotone in the ray parameter, and a closer hit POINT lies in a leaf entered before t_hit,
// which is visited.  One exception: a candidate cut out at a FRACTIONAL alpha draws again in every other leaf that holds it
// (raytracer.h:455) and may be accepted there.  Rules (the test suite's CPU checker counts the same rules independently and compares
// the tallies; hits, ids, uvs are those of the unpruned walk bit for bit):
//   R1  a child entered at t0 >= t_hit (t_hit of the moment it is tested) is marked `beyond` (scenes without stochastic alpha: it is
//       simply not taken — the slab test runs against the segment [0, t_hit));
//   R2  a node popped unmarked while a hit exists is marked when its own t0 >= t_hit (entries pushed before the hit was found);
//   R3  a marked node is dropped when it is reached, provided no fractional-alpha rejection lies in front of the hit
//       (t_hit <= t_frac); otherwise it is walked like the reference does.
// Leaves are reached in ascending t0, so when R3 drops a node every leaf entered before t_hit has been tested already.
#define GI_NODE_BEYOND 0x80000000u
#ifdef GI_NO_PRUNE
#define GI_PRUNE_T(t) CUDART_INF
#else
#define GI_PRUNE_T(t) (t)
#endif
#define GI_NODE_INDEX 0x7fffffffu
struct DHit {
    uint32_t prim;     // GI_NO_HIT on miss
    double t, u, v;    // ray parameter (+inf on a miss: the walk's pruning bound) and barycentrics (triangles)
    double tu, tv;     // uv as RayTracer::trace returns it (FULL traversal only; otherwise derived from prim,u,v)
    d3 n;              // cone normal (cones only)
};

// The walk is resumable: its state is (stack, TraceState, out).  trace_begin tests the root, trace_walk pops leaves until the
// ray is done (stack empty, or the stop rule fired) — or, with BAIL, until fewer than `min_active` lanes of the warp are still
// walking (*warp_active, a shared-memory counter each lane decrements when its ray is done), so that a persistent kernel can
// hand the idle lanes new rays (k_bounce_p).
struct TraceState { int sp; bool term; double best_d2, cur_tu, cur_tv, frac_t; };   // frac_t: see R1-R3 above (+inf: none); t_hit is DHit::t (+inf while there is no hit)

// R2 + R3 for a node taken off the stack
template <bool FULL>
__device__ __forceinline__ bool node_dropped(const DNode& nd, uint32_t entry, const DRay& r, double hit_t, double frac_t)
{
#ifdef GI_NO_PRUNE   // A/B switch: the reference's full walk (profiles/r02/README.md)
    return false;
#endif
    if (!(hit_t < CUDART_INF)) return false;                         // no hit yet
    if (FULL) {
        if (hit_t > frac_t) return false;                             // a fractional-alpha rejection in front of the hit: the reference's walk
        if (entry & GI_NODE_BEYOND) return true;                      // R1 mark
    }
    return box_entry(nd.bmin, nd.bmax, r, 0.0, hit_t) < 0.0;          // R2: entered at or beyond the hit
}

__device__ __forceinline__ bool trace_begin(const DScene& S, const DRay& r, DHit& out, TraceState& st, TStack& stack, uint32_t& n_node)
{
    st.sp = 0; st.term = false; st.best_d2 = 0; st.frac_t = CUDART_INF;
    st.cur_tu = 0; st.cur_tv = 0;   // `uv` local of RayTracer::trace: survives across candidates (raytracer.h:385)
    out.prim = GI_NO_HIT; out.t = CUDART_INF; out.u = 0; out.v = 0; out.tu = 0; out.tv = 0; out.n = mk3(0, 0, 0);
    if (S.n_nodes == 0) return false;
    DNode root = load_node(S.nodes, 0);
    n_node++;
    if (box_entry(root.bmin, root.bmax, r, 0.0, CUDART_INF) < 0.0) return false;
    stack_push(stack, st.sp, 0u, S.err);
    return true;
}

// one leaf record: 5 x 16-byte loads through the read-only path
struct LeafRec { double g[9]; uint32_t prim, flags; };
__device__ __forceinline__ LeafRec load_leafref(const DLeafRef* refs, uint32_t k)
{
    const double2* rp = reinterpret_cast<const double2*>(refs + k);
    LeafRec L;
    const double2 a0 = __ldg(rp), a1 = __ldg(rp + 1), a2 = __ldg(rp + 2), a3 = __ldg(rp + 3);
    const uint4 tail = __ldg(reinterpret_cast<const uint4*>(rp + 4));
    L.g[0] = a0.x; L.g[1] = a0.y; L.g[2] = a1.x; L.g[3] = a1.y; L.g[4] = a2.x; L.g[5] = a2.y; L.g[6] = a3.x; L.g[7] = a3.y;
    L.g[8] = __hiloint2double((int)tail.y, (int)tail.x);
    L.prim = tail.z; L.flags = tail.w;
    return L;
}

template <bool FULL, bool IMPL, bool BAIL>
__device__ __forceinline__ void trace_walk(const DScene& S, const DRay& r, uint64_t seed, uint64_t path, uint64_t depth, DHit& out, TraceState& st, TStack& stack,
                                           uint32_t& n_node, uint32_t& n_prim, volatile int* warp_active = nullptr, int min_active = 0)
{
    int sp = st.sp;
    bool term = st.term;
    double frac_t = st.frac_t;
    // Two forms of the search for the next leaf, same rules, same tallies (measured on all five scenes, profiles/r02/ab_t17_walker_forms.txt):
    //  * FULL or BAIL: ONE pop site, ONE node load, ONE drop test for both ways of getting the next node (nearest child / popped entry);
    //  * otherwise the round-1 loop (pop + load in front of it, one load at its bottom) with the drop test folded in: a dropped node
    //    counts as an interior node without hit children.  8 % faster on the glass scene, 7 % on cornell, within 1 % on caustics.
    // What must not happen in either form is a copy of the load sequence in each branch of "nearest child or pop": the lanes of a warp
    // then wait for each other's loads (bounce kernels +45 %, profiles/r02/ab_t11_load_sites.txt).  Also measured and dropped: the inverse
    // direction in shared memory during the walk (to keep the direction in registers for the primitive tests: 30-50 % slower), the leaf
    // box re-read from memory when a hit is accepted instead of kept in registers (50-60 % slower), 72 / 80 registers (ab_t12, ab_t16).
    while (sp > 0 && !term) {
        bool have_leaf;
        uint32_t ent = 0, ni = 0;
        DNode nd;
        if (FULL || BAIL) {
            bool need_pop = true;
            have_leaf = false;
            for (;;) {
                if (need_pop) {
                    if (sp == 0) break;
                    ent = stack_pop(stack, sp);
                }
                ni = ent & GI_NODE_INDEX;
                nd = load_node(S.nodes, ni);
                if (need_pop && node_dropped<FULL>(nd, ent, r, out.t, frac_t)) continue;   // R2 / R3
                if (nd.mask == 0) { have_leaf = true; break; }
                // interior: the hit children in visiting order; the nearest is walked into at once, the others are pushed far-to-near
                uint32_t seq, n;
                n_node += __popc(nd.mask);
                if (FULL) {
                    uint32_t n_near;
                    n = interior_step_marked<IMPL>(S.nodes, nd, r, seq, GI_PRUNE_T(out.t), n_near);
                    if (n_near == 0 && out.t <= frac_t) n = 0;   // every child is beyond the hit and R3 holds: reached next, they would all be dropped
                    for (int j = (int)n - 1; j >= 1; j--) stack_push(stack, sp, child_node(nd, (seq >> (4 * j)) & 7u) | ((uint32_t)j >= n_near ? GI_NODE_BEYOND : 0u), S.err);
                } else {
                    n = interior_step<IMPL, true>(S.nodes, nd, r, GI_PRUNE_T(out.t), seq);   // R1 + R3 at once: the segment ends at the hit
                    for (int j = (int)n - 1; j >= 1; j--) stack_push(stack, sp, child_node(nd, (seq >> (4 * j)) & 7u), S.err);
                }
                need_pop = n == 0;
                if (!need_pop) ent = child_node(nd, seq & 7u);
            }
        } else {
            ent = stack_pop(stack, sp);
            ni = ent & GI_NODE_INDEX;
            nd = load_node(S.nodes, ni);
            bool popped = true;
            have_leaf = true;
            while (nd.mask != 0) {
                uint32_t seq = 0, n = 0;
                if (!(popped && node_dropped<FULL>(nd, ent, r, out.t, frac_t))) {   // R2 / R3: a dropped node has no hit children
                    n_node += __popc(nd.mask);
                    n = interior_step<IMPL, true>(S.nodes, nd, r, GI_PRUNE_T(out.t), seq);   // R1 + R3 at once: the segment ends at the hit
                    for (int j = (int)n - 1; j >= 1; j--) stack_push(stack, sp, child_node(nd, (seq >> (4 * j)) & 7u), S.err);
                }
                popped = n == 0;
                if (popped) {
                    if (sp == 0) { have_leaf = false; break; }
                    ent = stack_pop(stack, sp);
                } else ent = child_node(nd, seq & 7u);
                ni = ent & GI_NODE_INDEX;
                nd = load_node(S.nodes, ni);
            }
            if (have_leaf && popped && node_dropped<FULL>(nd, ent, r, out.t, frac_t)) have_leaf = false;   // a leaf taken off the stack
        }
        if (have_leaf) {
            const DLeafRef* refs = S.refs + nd.prim_off;
            n_prim += nd.prim_cnt;
            for (uint32_t k = 0; k < nd.prim_cnt; k++) {
                const LeafRec L = load_leafref(refs, k);
                const uint32_t prim = L.prim, flags = L.flags;
                double t, u = 0, v = 0; d3 cn = mk3(0, 0, 0);
                bool ok;
                uint32_t kind = LF_KIND(flags);
                if (kind == GI_PRIM_TRIANGLE) { t = tri_hit(L.g, r, u, v); ok = t > 0; }
                else if (kind == GI_PRIM_SPHERE) ok = sphere_hit(L.g, r, t);
                else ok = cone_hit(L.g, S.prim_nrm + 9 * (size_t)prim, r, t, cn);
                if (!ok) continue;
                d3 hit = r.o + r.d * t;
                if (FULL) {
                    if (flags & LF_WRITES_UV) prim_uv_at(S, prim, hit, u, v, st.cur_tu, st.cur_tv);
                    if (flags & LF_ALPHA) {
                        bool frac = false;
                        if (!alpha_pass(S, prim, ni, st.cur_tu, st.cur_tv, seed, path, depth, SITE_ALPHA_TRACE, &frac)) {
                            if (frac && t < frac_t) frac_t = t;
                            continue;
                        }
                    }
                }
                double d2 = len2(hit - r.o);
                if (out.prim == GI_NO_HIT || d2 < st.best_d2) {
                    out.prim = prim; out.t = t; out.u = u; out.v = v; out.n = cn; st.best_d2 = d2;
                    if (FULL) { out.tu = st.cur_tu; out.tv = st.cur_tv; }
                    if (box_contains(nd.bmin, nd.bmax, hit)) term = true;
                }
            }
        }
        if (BAIL) {
            if (sp == 0 || term) atomicSub((int*)warp_active, 1);   // this lane's ray is done
            else if (*warp_active < min_active) break;             // too few lanes still walking: go and fetch rays
        }
    }
    st.frac_t = frac_t;
    st.sp = sp; st.term = term;
}

template <bool FULL, bool IMPL>
__device__ __forceinline__ void trace_closest(const DScene& S, const DRay& r, uint64_t seed, uint64_t path, uint64_t depth, DHit& out, uint32_t& n_node, uint32_t& n_prim)
{
    GI_TSTACK_DECL(stack);
    TraceState st;
    if (trace_begin(S, r, out, st, stack, n_node)) trace_walk<FULL, IMPL, false>(S, r, seed, path, depth, out, st, stack, n_node, n_prim);
}

// ---- any hit: RayTracer::visible (raytracer.h:280-319) over Octree::Node::intersect (octree.cpp:256-282) -------------------------
// Returns true when nothing blocks the segment.  Visiting order is free: the alpha draw is keyed by the (leaf, primitive)
// occurrence, so the outcome equals the reference's first-blocker search for any order.  The children met by the segment come
// from the same interior step as the closest-hit walk (with the segment's tmax), nearest first.
template <bool FULL, bool IMPL>
__device__ __forceinline__ bool trace_visible(const DScene& S, const DRay& r, double mt, uint64_t seed, uint64_t path, uint64_t depth, uint64_t light, uint32_t& n_node, uint32_t& n_prim)
{
    if (S.n_nodes == 0) return true;
    const double tmax = sqrt(mt) - GI_D_SHADOW_BIAS;   // raytracer.h:283
    GI_TSTACK_DECL(stack);
    int sp = 0;
    {
        DNode root = load_node(S.nodes, 0);
        n_node++;
        if (box_entry(root.bmin, root.bmax, r, 0.0, tmax) < 0.0) return true;
        stack_push(stack, sp, 0u, S.err);
    }
    while (sp > 0) {
        uint32_t ni = stack_pop(stack, sp);
        DNode nd = load_node(S.nodes, ni);
        bool have_leaf = true;
        while (nd.mask != 0) {
            uint32_t seq;
            n_node += __popc(nd.mask);
            const uint32_t n = interior_step<IMPL, false>(S.nodes, nd, r, tmax, seq);
            for (int j = (int)n - 1; j >= 1; j--) stack_push(stack, sp, child_node(nd, (seq >> (4 * j)) & 7u), S.err);
            if (n == 0) {
                if (sp == 0) { have_leaf = false; break; }
                ni = stack_pop(stack, sp);
            } else ni = child_node(nd, seq & 7u);
            nd = load_node(S.nodes, ni);
        }
        if (!have_leaf) break;
        const DLeafRef* refs = S.refs + nd.prim_off;
        for (uint32_t k = 0; k < nd.prim_cnt; k++) {
            const LeafRec L = load_leafref(refs, k);
            const uint32_t prim = L.prim, flags = L.flags;
            double t, u = 0, v = 0; d3 cn;
            bool ok;
            n_prim++;
            uint32_t kind = LF_KIND(flags);
            if (kind == GI_PRIM_TRIANGLE) { t = tri_hit(L.g, r, u, v); ok = t > 0; }
            else if (kind == GI_PRIM_SPHERE) ok = sphere_hit(L.g, r, t);
            else ok = cone_hit(L.g, S.prim_nrm + 9 * (size_t)prim, r, t, cn);
            if (!ok) continue;
            d3 pos = r.o + r.d * t;
            if (FULL && (flags & LF_ALPHA)) {
                double tu = 0, tv = 0;   // `uv` is a fresh (0,0) per candidate in visible() (raytracer.h:295)
                if (flags & LF_WRITES_UV) prim_uv_at(S, prim, pos, u, v, tu, tv);
                if (!alpha_pass(S, prim, ni, tu, tv, seed, path, depth, SITE_ALPHA_SHADOW + (light << 8))) continue;
            }
            double t_shadow = len2(pos - r.o);
            if ((t_shadow < mt) && (t_shadow > 0)) return false;
        }
    }
    return true;
}

// ---- warp-cooperative traversals: one ray per warp --------------------------------------------------------------------------
// For incoherent rays and for the long tail of deep paths a thread-per-ray walk serialises 32 different control flows; here
// the 32 lanes of a warp work on ONE ray: lanes 0..7 test the eight child boxes of an interior node, all lanes test up to
// 32 primitives of a leaf at once (geometry only), and the reference's sequential acceptance rule is then replayed over
// the geometric hits in stored order (there are rarely more than one or two per leaf).  Same visiting order, same
// arithmetic, same results as trace_closest / trace_visible; `stack` is GI_STACK_MAX words of shared memory per warp.
template <bool FULL, bool IMPL>
__device__ __forceinline__ void trace_closest_warp(const DScene& S, const DRay& r, uint64_t seed, uint64_t path, uint64_t depth, DHit& out, uint32_t* stack, int lane,
                                                   uint32_t& n_node, uint32_t& n_prim)
{
    int sp = 0;
    out.prim = GI_NO_HIT; out.t = CUDART_INF; out.u = 0; out.v = 0; out.tu = 0; out.tv = 0; out.n = mk3(0, 0, 0);
    double best_d2 = 0, cur_tu = 0, cur_tv = 0;
    if (S.n_nodes == 0) return;
    {
        DNode root = load_node(S.nodes, 0);
        n_node++;
        if (box_entry(root.bmin, root.bmax, r, 0.0, CUDART_INF) < 0.0) return;
        if (lane == 0) stack[0] = 0;
        sp = 1;
        __syncwarp();
    }
    bool term = false;
    double frac_t = CUDART_INF;   // R1-R3 of trace_walk, uniform over the warp (t_hit is out.t)
    while (sp > 0 && !term) {
        const uint32_t ent = stack[sp - 1]; sp--;
        const uint32_t ni = ent & GI_NODE_INDEX;
        __syncwarp();
        DNode nd = load_node(S.nodes, ni);
        if (node_dropped<FULL>(nd, ent, r, out.t, frac_t)) continue;
        if (nd.mask == 0) {
            n_prim += nd.prim_cnt;
            for (uint32_t base = 0; base < nd.prim_cnt; base += 32) {
                bool have = base + lane < nd.prim_cnt;
                double t = 0, u = 0, v = 0; d3 cn = mk3(0, 0, 0);
                uint32_t prim = 0, flags = 0;
                bool ok = false;
                if (have) {
                    const double2* rp = reinterpret_cast<const double2*>(S.refs + nd.prim_off + base + lane);
                    double g[9];
                    double2 a0 = __ldg(rp), a1 = __ldg(rp + 1), a2 = __ldg(rp + 2), a3 = __ldg(rp + 3);
                    g[0] = a0.x; g[1] = a0.y; g[2] = a1.x; g[3] = a1.y; g[4] = a2.x; g[5] = a2.y; g[6] = a3.x; g[7] = a3.y;
                    uint4 tail = __ldg(reinterpret_cast<const uint4*>(rp + 4));
                    g[8] = __hiloint2double((int)tail.y, (int)tail.x);
                    prim = tail.z; flags = tail.w;
                    uint32_t kind = LF_KIND(flags);
                    if (kind == GI_PRIM_TRIANGLE) { t = tri_hit(g, r, u, v); ok = t > 0; }
                    else if (kind == GI_PRIM_SPHERE) ok = sphere_hit(g, r, t);
                    else ok = cone_hit(g, S.prim_nrm + 9 * (size_t)prim, r, t, cn);
                }
                uint32_t hm = __ballot_sync(0xffffffffu, ok);
                while (hm) {   // replay the sequential acceptance over the geometric hits, in stored order
                    int src = __ffs(hm) - 1; hm &= hm - 1;
                    double ct = __shfl_sync(0xffffffffu, t, src), cu = __shfl_sync(0xffffffffu, u, src), cv = __shfl_sync(0xffffffffu, v, src);
                    uint32_t cprim = __shfl_sync(0xffffffffu, prim, src), cflags = __shfl_sync(0xffffffffu, flags, src);
                    d3 ccn = mk3(0, 0, 0);
                    if (LF_KIND(cflags) == GI_PRIM_CONE) ccn = mk3(__shfl_sync(0xffffffffu, cn.x, src), __shfl_sync(0xffffffffu, cn.y, src), __shfl_sync(0xffffffffu, cn.z, src));
                    d3 hit = r.o + r.d * ct;
                    if (FULL) {
                        if (cflags & LF_WRITES_UV) prim_uv_at(S, cprim, hit, cu, cv, cur_tu, cur_tv);
                        if (cflags & LF_ALPHA) {
                            bool frac = false;
                            if (!alpha_pass(S, cprim, ni, cur_tu, cur_tv, seed, path, depth, SITE_ALPHA_TRACE, &frac)) {
                                if (frac && ct < frac_t) frac_t = ct;
                                continue;
                            }
                        }
                    }
                    double d2 = len2(hit - r.o);
                    if (out.prim == GI_NO_HIT || d2 < best_d2) {
                        out.prim = cprim; out.t = ct; out.u = cu; out.v = cv; out.n = ccn; best_d2 = d2;
                        if (FULL) { out.tu = cur_tu; out.tv = cur_tv; }
                        if (box_contains(nd.bmin, nd.bmax, hit)) term = true;
                    }
                }
            }
            continue;
        }
        n_node += __popc(nd.mask);
        // R1: without stochastic alpha the children are tested against the segment [0, t_hit); with it, against the whole ray and marked
        const double seg = FULL ? CUDART_INF : GI_PRUNE_T(out.t);
        double t0 = -1.0; uint32_t cidx = 0;
        if (lane < 8 && ((nd.mask >> lane) & 1u)) {
            cidx = nd.child + __popc(nd.mask & ((1u << lane) - 1u));
            if (IMPL) { PlaneT T; plane_params(nd, r, T); t0 = child_entry(T, lane, r, 0.0, seg); }
            else { DNode ch = load_node(S.nodes, cidx); t0 = box_entry(ch.bmin, ch.bmax, r, 0.0, seg); }
            if (FULL && t0 >= GI_PRUNE_T(out.t)) cidx |= GI_NODE_BEYOND;
        }
        uint32_t valid = __ballot_sync(0xffffffffu, t0 >= 0.0) & 0xffu;
        if (FULL && out.t <= frac_t && (__ballot_sync(0xffffffffu, t0 >= 0.0 && !(cidx & GI_NODE_BEYOND)) & 0xffu) == 0) valid = 0;   // all beyond, R3 holds: none is pushed
        // far-to-near on the stack: rank = children that come before me in descending (t0, child index) order
        int rank = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            double tj = __shfl_sync(0xffffffffu, t0, j);
            if (((valid >> j) & 1u) && j != lane && (tj > t0 || (tj == t0 && j > lane))) rank++;
        }
        if (((valid >> lane) & 1u) && lane < 8 && sp + rank < GI_STACK_MAX) stack[sp + rank] = cidx;
        sp += __popc(valid);
        if (sp > GI_STACK_MAX) { sp = GI_STACK_MAX; if (lane == 0) atomicOr(S.err, GI_DEV_ERR_STACK); }
        __syncwarp();
    }
}

template <bool FULL, bool IMPL>
__device__ __forceinline__ bool trace_visible_warp(const DScene& S, const DRay& r, double mt, uint64_t seed, uint64_t path, uint64_t depth, uint64_t light, uint32_t* stack,
                                                   int lane, uint32_t& n_node, uint32_t& n_prim)
{
    if (S.n_nodes == 0) return true;
    const double tmax = sqrt(mt) - GI_D_SHADOW_BIAS;   // raytracer.h:283
    int sp = 0;
    {
        DNode root = load_node(S.nodes, 0);
        n_node++;
        if (box_entry(root.bmin, root.bmax, r, 0.0, tmax) < 0.0) return true;
        if (lane == 0) stack[0] = 0;
        sp = 1;
        __syncwarp();
    }
    while (sp > 0) {
        uint32_t ni = stack[sp - 1]; sp--;
        __syncwarp();
        DNode nd = load_node(S.nodes, ni);
        if (nd.mask == 0) {
            for (uint32_t base = 0; base < nd.prim_cnt; base += 32) {
                bool have = base + lane < nd.prim_cnt;
                bool blocked = false;
                if (have) {
                    const double2* rp = reinterpret_cast<const double2*>(S.refs + nd.prim_off + base + lane);
                    double g[9];
                    double2 a0 = __ldg(rp), a1 = __ldg(rp + 1), a2 = __ldg(rp + 2), a3 = __ldg(rp + 3);
                    g[0] = a0.x; g[1] = a0.y; g[2] = a1.x; g[3] = a1.y; g[4] = a2.x; g[5] = a2.y; g[6] = a3.x; g[7] = a3.y;
                    uint4 tail = __ldg(reinterpret_cast<const uint4*>(rp + 4));
                    g[8] = __hiloint2double((int)tail.y, (int)tail.x);
                    uint32_t prim = tail.z, flags = tail.w;
                    double t, u = 0, v = 0; d3 cn;
                    bool ok;
                    uint32_t kind = LF_KIND(flags);
                    if (kind == GI_PRIM_TRIANGLE) { t = tri_hit(g, r, u, v); ok = t > 0; }
                    else if (kind == GI_PRIM_SPHERE) ok = sphere_hit(g, r, t);
                    else ok = cone_hit(g, S.prim_nrm + 9 * (size_t)prim, r, t, cn);
                    if (ok) {
                        d3 pos = r.o + r.d * t;
                        bool pass = true;
                        if (FULL && (flags & LF_ALPHA)) {
                            double tu = 0, tv = 0;
                            if (flags & LF_WRITES_UV) prim_uv_at(S, prim, pos, u, v, tu, tv);
                            pass = alpha_pass(S, prim, ni, tu, tv, seed, path, depth, SITE_ALPHA_SHADOW + (light << 8));
                        }
                        double t_shadow = len2(pos - r.o);
                        blocked = pass && (t_shadow < mt) && (t_shadow > 0);
                    }
                }
                uint32_t bm = __ballot_sync(0xffffffffu, blocked);
                n_prim += (nd.prim_cnt - base < 32u ? nd.prim_cnt - base : 32u);
                if (bm) return false;
            }
            continue;
        }
        n_node += __popc(nd.mask);
        bool in = false; uint32_t cidx = 0;
        if (lane < 8 && ((nd.mask >> lane) & 1u)) {
            cidx = nd.child + __popc(nd.mask & ((1u << lane) - 1u));
            if (IMPL) { PlaneT T; plane_params(nd, r, T); in = child_entry(T, lane, r, 0.0, tmax) >= 0.0; }
            else { DNode ch = load_node(S.nodes, cidx); in = box_entry(ch.bmin, ch.bmax, r, 0.0, tmax) >= 0.0; }
        }
        uint32_t valid = __ballot_sync(0xffffffffu, in) & 0xffu;
        if (in) { int pos = sp + __popc(valid & ((1u << lane) - 1u)); if (pos < GI_STACK_MAX) stack[pos] = cidx; }
        sp += __popc(valid);
        if (sp > GI_STACK_MAX) { sp = GI_STACK_MAX; if (lane == 0) atomicOr(S.err, GI_DEV_ERR_STACK); }
        __syncwarp();
    }
    return true;
}

// add a per-thread pair of work counters into two global u64 tallies: warp shuffle reduction, one atomic pair per warp.
// Must be called by all 32 lanes of the warp (inactive lanes pass zeros).
__device__ __forceinline__ void tally2(unsigned long long* dst, uint32_t a, uint32_t b)
{
    unsigned long long x = a, y = b;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { x += __shfl_down_sync(0xffffffffu, x, o); y += __shfl_down_sync(0xffffffffu, y, o); }
    if ((threadIdx.x & 31) == 0 && dst) { if (x) atomicAdd(dst, x); if (y) atomicAdd(dst + 1, y); }
}

// ---- hit reconstruction: what RayTracer::trace hands back (hit point, shading normal, uv) from (prim, t, u, v) --------------
__device__ __forceinline__ void hit_surface(const DScene& S, const DRay& r, const DHit& h, bool full, d3& p, d3& n, double& tu, double& tv)
{
    p = r.o + r.d * h.t;
    tu = 0; tv = 0;
    uint32_t kind = S.prim_type[h.prim];
    if (kind == GI_PRIM_TRIANGLE) {
        const double* nn = S.prim_nrm + 9 * (size_t)h.prim;
        d3 n0 = ld3(nn), n1 = ld3(nn + 3), n2 = ld3(nn + 6);
        if (len2(n0) > 0 && len2(n1) > 0 && len2(n2) > 0) {                    // entities.h:480-487
            double w = 1 - h.u - h.v;
            n = (n0 * w + n1 * h.u) + n2 * h.v;
            prim_uv_at(S, h.prim, p, h.u, h.v, tu, tv);
        } else n = ld3(S.prim_fnorm + 3 * (size_t)h.prim);
    } else if (kind == GI_PRIM_SPHERE) {
        const double* g = S.prim_geom + 9 * (size_t)h.prim;
        n = normalize3(p - mk3(g[0], g[1], g[2]));                            // entities.h:86-91
        prim_uv_at(S, h.prim, p, 0, 0, tu, tv);
    } else n = h.n;
    if (full) { tu = h.tu; tv = h.tv; }
}

// ---- samplers (util.h / util.cpp) -----------------------------------------------------------------------------------------------
// fastPrecisePow(double,double) (util.h:113-136)
__device__ __forceinline__ double fast_precise_pow(double a, double b)
{
    int e = (int)b;
    int hi = __double2hiint(a);
    hi = (int)((b - e) * (hi - 1072632447) + 1072632447);
    double frac = __hiloint2double(hi, 0);
    double rr = 1.0;
    while (e) { if (e & 1) rr *= a; a *= a; e >>= 1; }
    return rr * frac;
}
// the tangent frame of util.cpp:41-45 applied to v (glm column-major mat * vec)
__device__ __forceinline__ d3 frame_apply(d3 n, d3 v)
{
    double z = fabs(n.z);
    double k = (1.0 / (1 + z));
    double c0x = z + k * -n.y * -n.y, c0y = k * (n.x * -n.y), c0z = -n.x;
    double c1x = k * (n.x * -n.y), c1y = z + k * -n.x * -n.x, c1z = -n.y;
    double c2x = n.x, c2y = n.y, c2z = z;
    return mk3(c0x * v.x + c1x * v.y + c2x * v.z, c0y * v.x + c1y * v.y + c2y * v.z, c0z * v.x + c1z * v.y + c2z * v.z);
}
// phi / cosTheta / sinTheta are float in the reference; cos, sin, sqrt are the double versions of float values
__device__ __forceinline__ d3 lobe_dir(float v, float cosTheta)
{
    float phi = (float)((double)(v * 2.0f) * GI_D_PI);
    float sinTheta = (float)sqrt((double)(1.0f - cosTheta * cosTheta));
    return mk3(cos((double)phi) * (double)sinTheta, sin((double)phi) * (double)sinTheta, (double)cosTheta);
}
__device__ __forceinline__ d3 hemisphere_cos(d3 n, float u, float v, double power)   // util.cpp:38-58
{
    float cosTheta = (float)fast_precise_pow((double)(1.0f - u), (1.0f / power));
    d3 res = frame_apply(n, lobe_dir(v, cosTheta));
    if (n.z < 0) res.z *= -1.0;
    return res;
}
__device__ __forceinline__ d3 sphere_cap_cos(d3 n, float u, float v, double power, double frac)   // util.cpp:60-83
{
    float cosTheta = (float)(frac * fast_precise_pow((double)(1.0f - u), (1.0f / power)) + (1 - frac));
    d3 res = frame_apply(n, lobe_dir(v, cosTheta));
    if (n.z < 0) res.z *= -1.0;
    return res;
}
__device__ __forceinline__ d3 sample_phong(d3 outdir, double power, double sx, double sy)   // util.cpp:91-107
{
    float u = (float)sx, v = (float)sy;
    float cosTheta = (float)fast_precise_pow((double)(1.0f - u), (1.0f / power));
    d3 res = frame_apply(outdir, lobe_dir(v, cosTheta));
    if (outdir.z < 0) res.z *= -1.0;
    return res;
}
__device__ __forceinline__ d3 random_unit_vec(double x, double y)   // util.h:183-188
{
    double theta = acos(2 * y - 1);
    return mk3(sin(theta) * cos(2 * x * GI_D_PI), sin(theta) * sin(2 * x * GI_D_PI), cos(theta));
}
__device__ __forceinline__ d3 refr3(d3 I, d3 N, double eta)   // util.h:173-181
{
    double d = dot3(N, I);
    double k = 1.0 - eta * eta * (1.0 - d * d);
    if (k < GI_D_EPSILON) return reflect3(I, N);
    return I * eta - N * (eta * d + sqrt(k));
}
// pow(x, 2) / pow(x, 5) with small integer exponents and pow(d, 1/roughness): libm calls in the reference.  CUDA's pow is
// within 2 ulp of glibc's (<1 ulp); exponent 1 is returned exactly like glibc does.
__device__ __forceinline__ double pow_like_libm(double x, double y) { return y == 1.0 ? x : pow(x, y); }

// Light::getPoint (light.h:42-45) / getPointInRange (light.h:47-53)
__device__ __forceinline__ d3 light_point(const gi_light& l, double x, double y) { return ld3(l.pos) + random_unit_vec(x, y) * l.rad; }
__device__ __forceinline__ d3 light_point_in_range(const gi_light& l, double x, double y)
{
    if (l.angle < 1) return ld3(l.pos) + sphere_cap_cos(ld3(l.dir), (float)x, (float)y, 1, l.angle) * l.rad;
    return light_point(l, x, y);
}

// RayTracer::rayType (raytracer.h:481-506)
__device__ __forceinline__ int ray_type(const DScene& S, const gi_material& m, const DRay& r, d3 norm, double tu, double tv, uint64_t seed, uint64_t path, uint64_t depth)
{
    int type = 2;
    double IOR = m.ior;
    double opacity = tex_alpha(S, m.diffuse_tex, tu, tv) * m.opacity;
    double q0 = (1 - IOR) / (1 + IOR);
    double r0 = q0 * q0;
    double fs = r0 + (1 - r0) * pow(1 - dot3(reflect3(r.d, norm), norm), 5.0);
    if (m.roughness < .001) type = 0;
    if (gi_rand(seed, path, depth, SITE(SITE_TYPE_A, 0)) > opacity) {
        if (gi_rand(seed, path, depth, SITE(SITE_TYPE_B, 0)) < fs) type = 0; else type = 1;
    }
    return type;
}
// RayTracer::secondaryRay (raytracer.h:321-379): flips norm in place, returns refDir, updates f / contrib / offset
__device__ __forceinline__ d3 secondary_ray(const DScene& S, const gi_material& m, const DRay& r, d3& norm, double tu, double tv, double sx, double sy, d3 color, d3& f, d3& contrib,
                                            double& offset, uint64_t seed, uint64_t path, uint64_t depth)
{
    bool backface = false;
    if (dot3(norm, r.d) > 0) { norm = norm * -1.0; backface = true; }
    int type = ray_type(S, m, r, norm, tu, tv, seed, path, depth);
    d3 refDir;
    if (type == 1) {
        refDir = refr3(r.d, norm, backface ? m.ior : 1.0 / m.ior);
        offset *= -1;
        contrib = mk3(1, 1, 1);
        f = color * 1.0;
    } else if (type == 0) {
        refDir = reflect3(r.d, norm);
        contrib = mk3(1, 1, 1);
        f = color * 1.0;
    } else {
        refDir = hemisphere_cos(norm, (float)sx, (float)sy, 2);
        if (m.roughness < .9) {
            refDir = sample_phong(reflect3(r.d, norm), (1.0 / (m.roughness)) + 1, sx, sy);
            if (dot3(refDir, norm) < 0) refDir = reflect3(refDir, norm);
        }
        f = color * 1.0;
        contrib = contrib * color;
        contrib = contrib + (color - contrib) * 0.5;   // glm::mix(contrib, inf, 0.5)
    }
    return refDir;
}

// ---- atmosphere: HeightFog::density (atmosphere.h:50-81), Octree::atmosphereDensity / atmosphereBounds (octree.cpp:214-251),
// RayTracer::raymarch (raytracer.h:509-529) --------------------------------------------------------------------------------------
#define GI_D_RAYMARCH_STEPSIZE 0.04   // util.h:29
// fastPow (util.h:100-111): exponent bit-hack on the high word, low word cleared.  The (int) cast of an out-of-range double is
// what x86's cvttsd2si returns (INT_MIN), e.g. for a == 0.
__device__ __forceinline__ double fast_pow(double a, double b)
{
    const double v = b * (double)(__double2hiint(a) - 1072632447) + 1072632447.0;
    const int hi = (v >= 2147483648.0 || v < -2147483648.0 || v != v) ? (int)0x80000000 : (int)v;
    return __hiloint2double(hi, 0);
}
// one noise-grid read: the reference indexes a std::vector<double> with a double expression (truncated to size_t); an index
// past the end is undefined behaviour there and reads 0 here
__device__ __forceinline__ double fog_cell(const DScene& S, const gi_fog& g, double idx)
{
    const unsigned long long i = (unsigned long long)idx;
    return i < g.grid_count ? __ldg(S.fog_grid + g.grid_offset + i) : 0.0;
}
__device__ __forceinline__ double fog_density(const DScene& S, const gi_fog& g, d3 p)
{
    const double sx = g.size[0], sz = g.size[2];          // nscale == 1 after the constructor (atmosphere.h:47)
    const double ymax = g.pos[1] + .5 * g.size[1];
    const d3 rel = (p - ld3(g.bmin)) * 1.0;
    const int ix = (int)rel.x, iy = (int)rel.y, iz = (int)rel.z;
    const double dx = rel.x - ix, dy = rel.y - iy, dz = rel.z - iz;
    // the row stride is s.x for both outer terms (atmosphere.h:61-71), kept as written
    const double r0 = (ix * sx + iy) * sz + iz, r1 = ((ix + 1) * sx + iy) * sz + iz;
    const double r2 = (ix * sx + (iy + 1)) * sz + iz, r3 = ((ix + 1) * sx + (iy + 1)) * sz + iz;
    const double c00 = (1 - dx) * fog_cell(S, g, r0) + dx * fog_cell(S, g, r1);
    const double c01 = (1 - dx) * fog_cell(S, g, r0 + 1) + dx * fog_cell(S, g, r1 + 1);
    const double c10 = (1 - dx) * fog_cell(S, g, r2) + dx * fog_cell(S, g, r3);
    const double c11 = (1 - dx) * fog_cell(S, g, r2 + 1) + dx * fog_cell(S, g, r3 + 1);
    const double c0 = c00 * (1 - dy) + c10 * dy;
    const double c1 = c01 * (1 - dy) + c11 * dy;
    const double noise = fast_pow((1 - dz) * c0 + dz * c1, 7);
    return g.density * noise * fast_pow((ymax - p.y) / g.size[1], 2);
}
// Octree::atmosphereDensity (octree.cpp:214-226): sum over the volumes containing pos; col = the last one's colour
__device__ __forceinline__ double atmosphere_density(const DScene& S, d3 pos, d3& col)
{
    double d = 0;
    for (uint32_t k = 0; k < S.n_fog; k++) {
        const gi_fog& g = S.fogs[k];
        if (box_contains(g.bmin, g.bmax, pos)) { col = ld3(g.col); d += GI_D_RAYMARCH_STEPSIZE * fog_density(S, g, pos); }
    }
    return d;
}
// Octree::atmosphereBounds (octree.cpp:229-251): `min` starts at 0 and can only shrink, `max` starts at 0 and can only grow, so
// the march starts at the caller's mint and ends at the farthest exit, clipped to the caller's maxt
__device__ __forceinline__ bool atmosphere_bounds(const DScene& S, const DRay& r, double& mint, double& maxt)
{
    double mn = 0, mx = 0; bool hit = false;
    for (uint32_t k = 0; k < S.n_fog; k++) {
        const gi_fog& g = S.fogs[k];
        double tmin = mint, tmax = maxt; bool ok = true;
        {
            double t0 = (g.bmin[0] - r.o.x) * r.inv.x, t1 = (g.bmax[0] - r.o.x) * r.inv.x;
            if (r.inv.x < 0.0) { double tmp = t0; t0 = t1; t1 = tmp; }
            tmin = t0 > tmin ? t0 : tmin; tmax = t1 < tmax ? t1 : tmax;
            if (tmax <= tmin) ok = false;
        }
        if (ok) {
            double t0 = (g.bmin[1] - r.o.y) * r.inv.y, t1 = (g.bmax[1] - r.o.y) * r.inv.y;
            if (r.inv.y < 0.0) { double tmp = t0; t0 = t1; t1 = tmp; }
            tmin = t0 > tmin ? t0 : tmin; tmax = t1 < tmax ? t1 : tmax;
            if (tmax <= tmin) ok = false;
        }
        if (ok) {
            double t0 = (g.bmin[2] - r.o.z) * r.inv.z, t1 = (g.bmax[2] - r.o.z) * r.inv.z;
            if (r.inv.z < 0.0) { double tmp = t0; t0 = t1; t1 = tmp; }
            tmin = t0 > tmin ? t0 : tmin; tmax = t1 < tmax ? t1 : tmax;
            if (tmax <= tmin) ok = false;
        }
        if (ok) { mn = tmin < mn ? tmin : mn; mx = mx < tmax ? tmax : mx; hit = true; }
    }
    mint = mint < mn ? mn : mint;
    maxt = mx < maxt ? mx : maxt;
    return hit;
}
// RayTracer::raymarch (raytracer.h:509-529); one counter draw per step.  Steps outside every volume have density 0 and can
// never scatter (drand() < 0 is false), so only the position is advanced there — with the same sequential additions.
__device__ __noinline__ bool raymarch(const DScene& S, const DRay& r, d3& hit, d3& col, double mint, double maxt, uint64_t seed, uint64_t path, uint64_t depth, uint64_t site, uint64_t hi)
{
    double t = mint + GI_D_SHADOW_BIAS;
    d3 cur = r.o + r.d * mint;
    const d3 stepv = r.d * GI_D_RAYMARCH_STEPSIZE;
    for (uint64_t step = 0; t < maxt; step++) {
        const double dens = atmosphere_density(S, cur, col);
        if (dens > 0.0 && gi_rand(seed, path, depth, SITE(site, (hi << 32) | step)) < dens) { hit = cur; return true; }
        cur = cur + stepv;
        t += GI_D_RAYMARCH_STEPSIZE;
    }
    return false;
}
// the fog part of RayTracer::radiance (raytracer.h:209-228): true when the segment origin..surface hit scatters in a volume
__device__ __forceinline__ bool fog_scatter(const DScene& S, const DRay& r, d3 surf_hit, d3& fhit, d3& fcol, uint64_t seed, uint64_t path, uint64_t depth, uint64_t site)
{
    double tmin = 0, tmax = sqrt(len2(surf_hit - r.o));   // glm::length
    fcol = mk3(0, 0, 0);
    return atmosphere_bounds(S, r, tmin, tmax) && raymarch(S, r, fhit, fcol, tmin, tmax, seed, path, depth, site, 0);
}
// the fog part of RayTracer::visible (raytracer.h:308-316): tmax is the SQUARED distance, as written
__device__ __forceinline__ bool fog_blocks(const DScene& S, const DRay& r, double mt, uint64_t seed, uint64_t path, uint64_t depth, uint64_t light)
{
    double tmin = 0, tmax = mt; d3 h, c = mk3(0, 0, 0);
    return atmosphere_bounds(S, r, tmin, tmax) && raymarch(S, r, h, c, tmin, tmax, seed, path, depth, SITE_FOG_SHADOW, light);
}

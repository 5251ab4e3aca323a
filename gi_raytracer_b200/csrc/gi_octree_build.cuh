// gi_octree_build.cuh — the scene octree built on the device: a level-synchronous restatement of
// Octree::rebuild / Octree::Node::partition (octree.cpp:53-119, 316-384) with the reference's entity / cell tests
// (triangle::intersect(BoundingBox) entities.h:522-528 -> triBoxOverlap util.cpp:257-330 with its float temporaries;
// sphere::intersect(BoundingBox) entities.h:108-141; cones inherit Entity::intersect(BoundingBox) = false, entities.h:38-41).
// Output = the breadth-first arrays of gi_scene_desc (existing children contiguous, in child order), identical to what the
// host build + Octree::flatten produce (tests compare them bit for bit).
//
// One level = every node that is still being subdivided ("active"), each with its entity list stored contiguously:
//   k_ob_classify   one thread per (active node, entity) item: which of the 8 child cells take the entity -> 8 flag planes
//   scan            ONE exclusive scan over the 8 planes (plane-major): a set flag's scan value is the entity's position in the
//                   next level's list buffer, and a child's list is the contiguous run [scan(start of node), scan(end of node))
//   k_ob_slots      one thread per active node: the 8 child counts, the "did not improve" rule, which children exist / go on
//   scan x3         node index of every new child, active index of every child that goes on, leaf-reference offset of every final leaf
//   k_ob_emit       one thread per child slot: node record (box, counts), parent's mask / first child
//   k_ob_scatter    one thread per item: entity id + owner slot into the next list; ids of final leaves into leaf_prims
#pragma once
#include "gi_kernels.cuh"

#define GI_OB_MAX_LEAF 16u        // MAX_ENTITIES_PER_LEAF (util.h:14)
#define GI_OB_MIN_LEAF_SIZE .0015 // MIN_LEAF_SIZE (util.h:16)
#define GI_OB_SUBDIV_RATIO 0.75   // MAX_SUBDIV_RATIO (util.h:17)
#define GI_OB_NONE 0xFFFFFFFFu

// AXISTEST_* of util.cpp:218-255: projections in float, radius in float
__device__ __forceinline__ bool ob_separated(double a, double b, double va_i, double va_j, double vb_i, double vb_j, float fa, float fb, double hi, double hj, bool neg_first)
{
    const float p0 = neg_first ? (float)(-a * va_i + b * va_j) : (float)(a * va_i - b * va_j);
    const float p1 = neg_first ? (float)(-a * vb_i + b * vb_j) : (float)(a * vb_i - b * vb_j);
    float mn, mx;
    if (p0 < p1) { mn = p0; mx = p1; } else { mn = p1; mx = p0; }
    const float rad = (float)(fa * hi + fb * hj);
    return mn > rad || mx < -rad;
}
// planeBoxOverlap (util.cpp:195-216)
__device__ __forceinline__ bool ob_plane_box(d3 normal, float d, d3 maxbox)
{
    d3 vmin, vmax;
    if (normal.x > 0.0f) { vmin.x = -maxbox.x; vmax.x = maxbox.x; } else { vmin.x = maxbox.x; vmax.x = -maxbox.x; }
    if (normal.y > 0.0f) { vmin.y = -maxbox.y; vmax.y = maxbox.y; } else { vmin.y = maxbox.y; vmax.y = -maxbox.y; }
    if (normal.z > 0.0f) { vmin.z = -maxbox.z; vmax.z = maxbox.z; } else { vmin.z = maxbox.z; vmax.z = -maxbox.z; }
    if (dot3(normal, vmin) + d > 0.0f) return false;
    if (dot3(normal, vmax) + d >= 0.0f) return true;
    return false;
}
__device__ __forceinline__ bool ob_axis_out(double a0, double a1, double a2, double h)
{
    float mn, mx;
    mn = mx = (float)a0;
    if (a1 < mn) mn = (float)a1;
    if (a1 > mx) mx = (float)a1;
    if (a2 < mn) mn = (float)a2;
    if (a2 > mx) mx = (float)a2;
    return mn > h || mx < -h;
}
// triBoxOverlap (util.cpp:257-330)
__device__ __forceinline__ bool ob_tri_box(d3 c, d3 h, d3 t0, d3 t1, d3 t2)
{
    const d3 v0 = t0 - c, v1 = t1 - c, v2 = t2 - c;
    const d3 e0 = v1 - v0, e1 = v2 - v1, e2 = v0 - v2;
    float fex, fey, fez;
    fex = (float)fabs(e0.x); fey = (float)fabs(e0.y); fez = (float)fabs(e0.z);
    if (ob_separated(e0.z, e0.y, v0.y, v0.z, v2.y, v2.z, fez, fey, h.y, h.z, false)) return false;
    if (ob_separated(e0.z, e0.x, v0.x, v0.z, v2.x, v2.z, fez, fex, h.x, h.z, true)) return false;
    if (ob_separated(e0.y, e0.x, v1.x, v1.y, v2.x, v2.y, fey, fex, h.x, h.y, false)) return false;
    fex = (float)fabs(e1.x); fey = (float)fabs(e1.y); fez = (float)fabs(e1.z);
    if (ob_separated(e1.z, e1.y, v0.y, v0.z, v2.y, v2.z, fez, fey, h.y, h.z, false)) return false;
    if (ob_separated(e1.z, e1.x, v0.x, v0.z, v2.x, v2.z, fez, fex, h.x, h.z, true)) return false;
    if (ob_separated(e1.y, e1.x, v0.x, v0.y, v1.x, v1.y, fey, fex, h.x, h.y, false)) return false;
    fex = (float)fabs(e2.x); fey = (float)fabs(e2.y); fez = (float)fabs(e2.z);
    if (ob_separated(e2.z, e2.y, v0.y, v0.z, v1.y, v1.z, fez, fey, h.y, h.z, false)) return false;
    if (ob_separated(e2.z, e2.x, v0.x, v0.z, v1.x, v1.z, fez, fex, h.x, h.z, true)) return false;
    if (ob_separated(e2.y, e2.x, v1.x, v1.y, v2.x, v2.y, fey, fex, h.x, h.y, false)) return false;
    if (ob_axis_out(v0.x, v1.x, v2.x, h.x) || ob_axis_out(v0.y, v1.y, v2.y, h.y) || ob_axis_out(v0.z, v1.z, v2.z, h.z)) return false;
    const d3 normal = cross3(e0, e1);
    const float d = (float)(-dot3(normal, v0));
    return ob_plane_box(normal, d, h);
}
// Entity::intersect(BoundingBox) by primitive kind
__device__ __forceinline__ bool ob_entity_in_cell(uint32_t kind, const double* g, const double* cmin, const double* cmax)
{
    if (kind == GI_PRIM_TRIANGLE) {   // entities.h:522-528: the cell is grown by EPSILON, centre = min + .5*(max-min), half = d/2
        const d3 lo = mk3(cmin[0] - GI_D_EPSILON, cmin[1] - GI_D_EPSILON, cmin[2] - GI_D_EPSILON), hi = mk3(cmax[0] + GI_D_EPSILON, cmax[1] + GI_D_EPSILON, cmax[2] + GI_D_EPSILON);
        const d3 c = lo + (hi - lo) * 0.5;
        const d3 h = mk3((hi.x - lo.x) / 2, (hi.y - lo.y) / 2, (hi.z - lo.z) / 2);
        return ob_tri_box(c, h, mk3(g[0], g[1], g[2]), mk3(g[3], g[4], g[5]), mk3(g[6], g[7], g[8]));
    }
    if (kind == GI_PRIM_SPHERE) {     // entities.h:108-141
        double sq = 0.0;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            double out = 0;
            if (g[k] < cmin[k]) { const double val = cmin[k] - g[k]; out += val * val; }
            if (g[k] > cmax[k]) { const double val = g[k] - cmax[k]; out += val * val; }
            sq += out;
        }
        return sq <= g[3] * g[3];
    }
    return false;                     // cone: Entity::intersect(BoundingBox) (entities.h:38-41)
}

struct DOBLevel {
    uint32_t n_items, n_active;
    const uint32_t* list;        // [n_items] entity ids, the active nodes' lists (and, skipped, those of the previous level's final leaves)
    const uint32_t* owner;       // [n_items] slot (previous active index * 8 + child) that owns the item; nullptr at the root level
    const uint32_t* slot_active; // previous level: slot -> active index at this level, or NONE
    const double* a_box;         // [n_active][6]
    const uint32_t* a_node;      // [n_active] global node index
    const uint32_t* a_start;     // [n_active] first item
    const uint32_t* a_count;     // [n_active]
};

__global__ void k_ob_classify(DOBLevel L, const uint8_t* __restrict__ prim_type, const double* __restrict__ prim_geom, const double* __restrict__ prim_bbox, uint32_t* __restrict__ flags)
{
    const uint32_t it = blockIdx.x * blockDim.x + threadIdx.x;
    if (it >= L.n_items) return;
    uint32_t a = 0;
    if (L.owner) a = L.slot_active[L.owner[it]];
    uint32_t bits = 0;
    if (a != GI_OB_NONE) {
        const uint32_t e = L.list[it];
        const double* eb = prim_bbox + 6 * (size_t)e;
        const double* nb = L.a_box + 6 * (size_t)a;
        const double* g = prim_geom + 9 * (size_t)e;
        const uint32_t kind = prim_type[e];
        const bool wide = (eb[3] - eb[0]) > GI_D_EPSILON;   // `bbox.dx() > EPSILON` (octree.cpp:338)
        if (wide) {
#pragma unroll 1
            for (int i = 0; i < 8; i++) {
                double cmin[3], cmax[3];
                child_box(nb, nb + 3, i, cmin, cmax);
                const bool overlap = (cmin[0] <= eb[3] && cmax[0] >= eb[0]) && (cmin[1] <= eb[4] && cmax[1] >= eb[1]) && (cmin[2] <= eb[5] && cmax[2] >= eb[2]);   // bbox.h:33-38
                if (overlap && ob_entity_in_cell(kind, g, cmin, cmax)) bits |= 1u << i;
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 8; i++) flags[(size_t)i * L.n_items + it] = (bits >> i) & 1u;
    if (it == 0) flags[(size_t)8 * L.n_items] = 0;   // sentinel: its scan value is the total
}

// per active node: child counts from the scan, the rules of octree.cpp:346-383
struct DOBSlots { uint32_t* exists; uint32_t* cont; uint32_t* final_cnt; uint32_t* count; uint32_t* start; };
__global__ void k_ob_slots(DOBLevel L, const uint32_t* __restrict__ pos, DOBSlots S)
{
    const uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= L.n_active) return;
    const uint32_t s0 = L.a_start[a], s1 = s0 + L.a_count[a];
    uint32_t c[8], st[8];
    double avg = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        st[i] = pos[(size_t)i * L.n_items + s0];
        c[i] = pos[(size_t)i * L.n_items + s1] - st[i];   // s1 == n_items reads the next plane's first entry / the sentinel
        if (c[i]) avg += (double)c[i];
    }
    avg /= 8;
    const bool improved = !(avg > GI_OB_SUBDIV_RATIO * (double)L.a_count[a]);   // octree.cpp:363
    const double* nb = L.a_box + 6 * (size_t)a;
#pragma unroll 1
    for (int i = 0; i < 8; i++) {
        double cmin[3], cmax[3];
        child_box(nb, nb + 3, i, cmin, cmax);
        const bool ex = c[i] > 0;
        const bool go = ex && improved && c[i] > GI_OB_MAX_LEAF && (cmax[0] - cmin[0]) > GI_OB_MIN_LEAF_SIZE;   // octree.cpp:376
        const size_t s = (size_t)a * 8 + i;
        S.exists[s] = ex ? 1u : 0u; S.cont[s] = go ? 1u : 0u; S.final_cnt[s] = (ex && !go) ? c[i] : 0u; S.count[s] = c[i]; S.start[s] = st[i];
    }
}

struct DOBOut { double* node_box; uint32_t* node_child; uint8_t* node_mask; uint32_t* node_prim_off; uint32_t* node_prim_cnt; uint32_t* leaf_prims; };
struct DOBNext { double* a_box; uint32_t* a_node; uint32_t* a_start; uint32_t* a_count; uint32_t* slot_active; };
// per child slot: the new node's record; per active node (slot 0 of each): the parent's mask, first child and prim_off
__global__ void k_ob_emit(DOBLevel L, DOBSlots S, const uint32_t* __restrict__ node_rank, const uint32_t* __restrict__ active_rank, const uint32_t* __restrict__ leaf_rank, uint32_t node_base,
                          uint32_t leaf_base, DOBOut O, DOBNext N)
{
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= L.n_active * 8u) return;
    const uint32_t a = s >> 3; const int i = (int)(s & 7u);
    if (i == 0) {
        uint32_t mask = 0;
        for (int k = 0; k < 8; k++) if (S.exists[(size_t)a * 8 + k]) mask |= 1u << k;
        const uint32_t pn = L.a_node[a];
        O.node_mask[pn] = (uint8_t)mask;
        O.node_child[pn] = mask ? node_base + node_rank[s] : 0u;
        O.node_prim_cnt[pn] = 0;                      // Node::partition clears the list (octree.cpp:370-371); prim_off was set when the node was made
    }
    N.slot_active[s] = S.cont[s] ? active_rank[s] : GI_OB_NONE;
    if (!S.exists[s]) return;
    const uint32_t nn = node_base + node_rank[s];
    const double* nb = L.a_box + 6 * (size_t)a;
    double cmin[3], cmax[3];
    child_box(nb, nb + 3, i, cmin, cmax);
    double* ob = O.node_box + 6 * (size_t)nn;
    ob[0] = cmin[0]; ob[1] = cmin[1]; ob[2] = cmin[2]; ob[3] = cmax[0]; ob[4] = cmax[1]; ob[5] = cmax[2];
    // a child that goes on gets its mask / child / prim_off when it is split at the next level
    O.node_mask[nn] = 0; O.node_child[nn] = 0;
    O.node_prim_off[nn] = leaf_base + leaf_rank[s];
    O.node_prim_cnt[nn] = S.cont[s] ? 0u : S.count[s];
    if (S.cont[s]) {
        const uint32_t na = active_rank[s];
        double* ab = N.a_box + 6 * (size_t)na;
        ab[0] = cmin[0]; ab[1] = cmin[1]; ab[2] = cmin[2]; ab[3] = cmax[0]; ab[4] = cmax[1]; ab[5] = cmax[2];
        N.a_node[na] = nn; N.a_start[na] = S.start[s]; N.a_count[na] = S.count[s];
    }
}

__global__ void k_ob_scatter(DOBLevel L, const uint32_t* __restrict__ flags, const uint32_t* __restrict__ pos, DOBSlots S, const uint32_t* __restrict__ leaf_rank, uint32_t leaf_base,
                             uint32_t* __restrict__ next_list, uint32_t* __restrict__ next_owner, uint32_t* __restrict__ leaf_prims)
{
    const uint32_t it = blockIdx.x * blockDim.x + threadIdx.x;
    if (it >= L.n_items) return;
    uint32_t a = 0;
    if (L.owner) a = L.slot_active[L.owner[it]];
    if (a == GI_OB_NONE) return;
    const uint32_t e = L.list[it];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const size_t k = (size_t)i * L.n_items + it;
        if (!flags[k]) continue;
        const uint32_t p = pos[k];
        const uint32_t s = a * 8u + (uint32_t)i;
        if (S.cont[s]) { next_list[p] = e; next_owner[p] = s; }
        else { leaf_prims[leaf_base + leaf_rank[s] + (p - S.start[s])] = e; next_list[p] = e; next_owner[p] = s; }   // final leaf: stored order = parent's order
    }
}
__global__ void k_ob_iota(uint32_t n, uint32_t* out)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = i;
}

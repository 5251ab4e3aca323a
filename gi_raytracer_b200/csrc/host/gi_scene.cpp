// gi_scene.cpp — entities, the host octree build and its flattening (see gi_scene.hpp).
// Every routine that decides the SHAPE of the tree restates the reference's arithmetic exactly (file:line cited),
// because leaf membership and leaf order are what make GPU hit ids bit-identical to the reference's.
#include "gi_scene.hpp"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <deque>
#include <iostream>
#include <limits>
#include <unordered_map>

using gi::dvec2;
using gi::dvec3;
using gi::dmat3;

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

namespace gi {

dmat3 eulerAngleXYZ(double t1, double t2, double t3)
{
    // glm negates the angles (gtx/euler_angles.inl:142-147)
    double c1 = std::cos(-t1), c2 = std::cos(-t2), c3 = std::cos(-t3);
    double s1 = std::sin(-t1), s2 = std::sin(-t2), s3 = std::sin(-t3);
    dmat3 r;
    r.c[0][0] = c2 * c3;
    r.c[0][1] = -c1 * s3 + s1 * s2 * c3;
    r.c[0][2] = s1 * s3 + c1 * s2 * c3;
    r.c[1][0] = c2 * s3;
    r.c[1][1] = c1 * c3 + s1 * s2 * s3;
    r.c[1][2] = -s1 * c3 + c1 * s2 * s3;
    r.c[2][0] = -s2;
    r.c[2][1] = s1 * c2;
    r.c[2][2] = c1 * c2;
    return r;
}

dmat3 inverse(const dmat3& in)
{
    const double(*m)[3] = in.c;
    double ood = 1.0 / (+m[0][0] * (m[1][1] * m[2][2] - m[2][1] * m[1][2]) - m[1][0] * (m[0][1] * m[2][2] - m[2][1] * m[0][2]) +
                        m[2][0] * (m[0][1] * m[1][2] - m[1][1] * m[0][2]));
    dmat3 r;
    r.c[0][0] = +(m[1][1] * m[2][2] - m[2][1] * m[1][2]) * ood;
    r.c[1][0] = -(m[1][0] * m[2][2] - m[2][0] * m[1][2]) * ood;
    r.c[2][0] = +(m[1][0] * m[2][1] - m[2][0] * m[1][1]) * ood;
    r.c[0][1] = -(m[0][1] * m[2][2] - m[2][1] * m[0][2]) * ood;
    r.c[1][1] = +(m[0][0] * m[2][2] - m[2][0] * m[0][2]) * ood;
    r.c[2][1] = -(m[0][0] * m[2][1] - m[2][0] * m[0][1]) * ood;
    r.c[0][2] = +(m[0][1] * m[1][2] - m[1][1] * m[0][2]) * ood;
    r.c[1][2] = -(m[0][0] * m[1][2] - m[1][0] * m[0][2]) * ood;
    r.c[2][2] = +(m[0][0] * m[1][1] - m[1][0] * m[0][1]) * ood;
    return r;
}

// glm::inverse of the 4x4 matrix glm::eulerAngleXYZ returns (detail/func_matrix.inl:297-352), upper-left 3x3 of the
// result: what `cone::rot = glm::inverse(glm::eulerAngleXYZ(..))` stores (entities.h:153).  Not the 3x3 formula: the
// cofactor expansion differs in rounding and in the sign of zeros.
dmat3 inverse_euler4(const dmat3& in)
{
    double m[4][4] = { { in.c[0][0], in.c[0][1], in.c[0][2], 0 }, { in.c[1][0], in.c[1][1], in.c[1][2], 0 }, { in.c[2][0], in.c[2][1], in.c[2][2], 0 }, { 0, 0, 0, 1 } };
    double Coef00 = m[2][2] * m[3][3] - m[3][2] * m[2][3], Coef02 = m[1][2] * m[3][3] - m[3][2] * m[1][3], Coef03 = m[1][2] * m[2][3] - m[2][2] * m[1][3];
    double Coef04 = m[2][1] * m[3][3] - m[3][1] * m[2][3], Coef06 = m[1][1] * m[3][3] - m[3][1] * m[1][3], Coef07 = m[1][1] * m[2][3] - m[2][1] * m[1][3];
    double Coef08 = m[2][1] * m[3][2] - m[3][1] * m[2][2], Coef10 = m[1][1] * m[3][2] - m[3][1] * m[1][2], Coef11 = m[1][1] * m[2][2] - m[2][1] * m[1][2];
    double Coef12 = m[2][0] * m[3][3] - m[3][0] * m[2][3], Coef14 = m[1][0] * m[3][3] - m[3][0] * m[1][3], Coef15 = m[1][0] * m[2][3] - m[2][0] * m[1][3];
    double Coef16 = m[2][0] * m[3][2] - m[3][0] * m[2][2], Coef18 = m[1][0] * m[3][2] - m[3][0] * m[1][2], Coef19 = m[1][0] * m[2][2] - m[2][0] * m[1][2];
    double Coef20 = m[2][0] * m[3][1] - m[3][0] * m[2][1], Coef22 = m[1][0] * m[3][1] - m[3][0] * m[1][1], Coef23 = m[1][0] * m[2][1] - m[2][0] * m[1][1];
    const double Fac0[4] = { Coef00, Coef00, Coef02, Coef03 }, Fac1[4] = { Coef04, Coef04, Coef06, Coef07 }, Fac2[4] = { Coef08, Coef08, Coef10, Coef11 };
    const double Fac3[4] = { Coef12, Coef12, Coef14, Coef15 }, Fac4[4] = { Coef16, Coef16, Coef18, Coef19 }, Fac5[4] = { Coef20, Coef20, Coef22, Coef23 };
    const double Vec0[4] = { m[1][0], m[0][0], m[0][0], m[0][0] }, Vec1[4] = { m[1][1], m[0][1], m[0][1], m[0][1] };
    const double Vec2[4] = { m[1][2], m[0][2], m[0][2], m[0][2] }, Vec3[4] = { m[1][3], m[0][3], m[0][3], m[0][3] };
    const double SignA[4] = { +1, -1, +1, -1 }, SignB[4] = { -1, +1, -1, +1 };
    double Inv[4][4];
    for (int k = 0; k < 4; k++) {
        Inv[0][k] = ((Vec1[k] * Fac0[k] - Vec2[k] * Fac1[k]) + Vec3[k] * Fac2[k]) * SignA[k];
        Inv[1][k] = ((Vec0[k] * Fac0[k] - Vec2[k] * Fac3[k]) + Vec3[k] * Fac4[k]) * SignB[k];
        Inv[2][k] = ((Vec0[k] * Fac1[k] - Vec1[k] * Fac3[k]) + Vec3[k] * Fac5[k]) * SignA[k];
        Inv[3][k] = ((Vec0[k] * Fac2[k] - Vec1[k] * Fac4[k]) + Vec2[k] * Fac5[k]) * SignB[k];
    }
    const double Dot0[4] = { m[0][0] * Inv[0][0], m[0][1] * Inv[1][0], m[0][2] * Inv[2][0], m[0][3] * Inv[3][0] };
    const double ood = 1.0 / ((Dot0[0] + Dot0[1]) + (Dot0[2] + Dot0[3]));
    dmat3 r;
    for (int c = 0; c < 3; c++) for (int k = 0; k < 3; k++) r.c[c][k] = Inv[c][k] * ood;
    return r;
}

}  // namespace gi

// ---- triangle / box overlap: Akenine-Möller separating-axis test as the reference uses it (util.cpp:187-330).
// The reference keeps min/max/p/rad/d and the |edge| terms in `float` while vertices and edges are double; the
// narrowing decides borderline leaf membership, so it is reproduced here.
namespace {

struct SatAxis {
    // one "AXISTEST": project two vertices on the axis (a, b are edge components, i/j the vertex components used)
    static bool separated(double a, double b, double va_i, double va_j, double vb_i, double vb_j, float fa, float fb, double hi, double hj, bool neg_first)
    {
        float p0 = neg_first ? (float)(-a * va_i + b * va_j) : (float)(a * va_i - b * va_j);
        float p1 = neg_first ? (float)(-a * vb_i + b * vb_j) : (float)(a * vb_i - b * vb_j);
        float mn, mx;
        if (p0 < p1) { mn = p0; mx = p1; } else { mn = p1; mx = p0; }
        float rad = (float)(fa * hi + fb * hj);
        return mn > rad || mx < -rad;
    }
};

bool plane_box_overlap(const dvec3& normal, float d, const dvec3& maxbox)  // util.cpp:195-216
{
    dvec3 vmin, vmax;
    for (int q = 0; q <= 2; q++) {
        if (normal[q] > 0.0f) { vmin[q] = -maxbox[q]; vmax[q] = maxbox[q]; }
        else { vmin[q] = maxbox[q]; vmax[q] = -maxbox[q]; }
    }
    if (gi::dot(normal, vmin) + d > 0.0f) return false;
    if (gi::dot(normal, vmax) + d >= 0.0f) return true;
    return false;
}

bool tri_box_overlap(dvec3 boxcenter, dvec3 h, const dvec3 tv[3])  // util.cpp:257-330
{
    dvec3 v0 = tv[0] - boxcenter, v1 = tv[1] - boxcenter, v2 = tv[2] - boxcenter;
    dvec3 e0 = v1 - v0, e1 = v2 - v1, e2 = v0 - v2;
    float fex, fey, fez;
    // edge 0: X01 (v0,v2 on y/z), Y02 (v0,v2 on x/z, negated first term), Z12 (v1,v2 on x/y)
    fex = (float)std::abs(e0[0]); fey = (float)std::abs(e0[1]); fez = (float)std::abs(e0[2]);
    if (SatAxis::separated(e0[2], e0[1], v0[1], v0[2], v2[1], v2[2], fez, fey, h[1], h[2], false)) return false;
    if (SatAxis::separated(e0[2], e0[0], v0[0], v0[2], v2[0], v2[2], fez, fex, h[0], h[2], true)) return false;
    if (SatAxis::separated(e0[1], e0[0], v1[0], v1[1], v2[0], v2[1], fey, fex, h[0], h[1], false)) return false;
    // edge 1: X01, Y02, Z0 (v0,v1 on x/y)
    fex = (float)std::abs(e1[0]); fey = (float)std::abs(e1[1]); fez = (float)std::abs(e1[2]);
    if (SatAxis::separated(e1[2], e1[1], v0[1], v0[2], v2[1], v2[2], fez, fey, h[1], h[2], false)) return false;
    if (SatAxis::separated(e1[2], e1[0], v0[0], v0[2], v2[0], v2[2], fez, fex, h[0], h[2], true)) return false;
    if (SatAxis::separated(e1[1], e1[0], v0[0], v0[1], v1[0], v1[1], fey, fex, h[0], h[1], false)) return false;
    // edge 2: X2 (v0,v1 on y/z), Y1 (v0,v1 on x/z, negated first term), Z12
    fex = (float)std::abs(e2[0]); fey = (float)std::abs(e2[1]); fez = (float)std::abs(e2[2]);
    if (SatAxis::separated(e2[2], e2[1], v0[1], v0[2], v1[1], v1[2], fez, fey, h[1], h[2], false)) return false;
    if (SatAxis::separated(e2[2], e2[0], v0[0], v0[2], v1[0], v1[2], fez, fex, h[0], h[2], true)) return false;
    if (SatAxis::separated(e2[1], e2[0], v1[0], v1[1], v2[0], v2[1], fey, fex, h[0], h[1], false)) return false;
    // the three box axes: float min/max of the (double) vertex coordinates against the half size
    for (int ax = 0; ax < 3; ax++) {
        float mn, mx;
        mn = mx = (float)v0[ax];
        if (v1[ax] < mn) mn = (float)v1[ax];
        if (v1[ax] > mx) mx = (float)v1[ax];
        if (v2[ax] < mn) mn = (float)v2[ax];
        if (v2[ax] > mx) mx = (float)v2[ax];
        if (mn > h[ax] || mx < -h[ax]) return false;
    }
    // the triangle's plane
    dvec3 normal = gi::cross(e0, e1);
    float d = (float)(-gi::dot(normal, v0));
    return plane_box_overlap(normal, d, h);
}

}  // namespace

// ---- entities ------------------------------------------------------------------------------------------------
Entity::Entity() : material(Material(new texture(dvec3(1, 0, 0)), new texture(dvec3(0, 0, 0)), .75, 1)) {}  // entities.h:19

bool sphere::intersect(BoundingBox bbox)  // entities.h:108-141: squared distance from the centre to the box
{
    auto check = [](double v, double bmin, double bmax) {
        double out = 0;
        if (v < bmin) { double val = (bmin - v); out += val * val; }
        if (v > bmax) { double val = (v - bmax); out += val * val; }
        return out;
    };
    double sq = 0.0;
    sq += check(pos.x, bbox.min.x, bbox.max.x);
    sq += check(pos.y, bbox.min.y, bbox.max.y);
    sq += check(pos.z, bbox.min.z, bbox.max.z);
    return sq <= (rad * rad);
}

cone::cone(dvec3 position, dvec3 rotation, double radius, double h, const Material& m) : Entity(m), rad(radius), height(h)
{
    pos = position;
    rot = gi::inverse_euler4(gi::eulerAngleXYZ(rotation.x, rotation.y, rotation.z));  // entities.h:153 (a 4x4 inverse in the reference)
}

BoundingBox cone::boundingBox() const  // entities.h:260-299: AABB of the bounding pyramid
{
    const double inf = std::numeric_limits<double>::infinity();
    dvec3 mn(inf, inf, inf), mx(-inf, -inf, -inf);
    dvec3 verts[5] = { rad * dvec3(-1, -1, 0), rad * dvec3(-1, 1, 0), rad * dvec3(1, -1, 0), rad * dvec3(1, 1, 0), dvec3(0, 0, height) };
    dmat3 irot = gi::inverse(rot);
    for (int i = 0; i < 5; i++) {
        dvec3 v = verts[i] * irot + pos;
        for (int k = 0; k < 3; k++) {
            if (v[k] < mn[k]) mn[k] = v[k];
            if (v[k] > mx[k]) mx[k] = v[k];
        }
    }
    return BoundingBox(mn, mx);
}

triangle::triangle(vertex v1, vertex v2, vertex v3, const Material& m) : Entity(m)  // entities.h:335-342
{
    vertices = { { v1, v2, v3 } };
    norm = gi::normalize(gi::cross((v2.pos - v1.pos), (v3.pos - v1.pos)));
    inv_area = 1.0 / gi::length(gi::cross(vertices[0].pos - vertices[1].pos, vertices[0].pos - vertices[2].pos));
}

bool triangle::intersect(BoundingBox bbox)  // entities.h:522-528: the cell is grown by EPSILON before the SAT
{
    BoundingBox tmp(bbox.min - GI_EPSILON, bbox.max + GI_EPSILON);
    dvec3 verts[3] = { vertices[0].pos, vertices[1].pos, vertices[2].pos };
    return tri_box_overlap(tmp.center(), dvec3(tmp.dx() / 2, tmp.dy() / 2, tmp.dz() / 2), verts);
}

BoundingBox triangle::boundingBox() const  // entities.h:530-557: max is padded by EPSILON once per vertex iteration
{
    const double inf = std::numeric_limits<double>::infinity();
    dvec3 mn(inf, inf, inf), mx(-inf, -inf, -inf);
    for (int i = 0; i < 3; i++) {
        dvec3 v = vertices[i].pos;
        if (v.x < mn.x) mn.x = v.x;
        if (v.x > mx.x) mx.x = v.x;
        if (v.y < mn.y) mn.y = v.y;
        if (v.y > mx.y) mx.y = v.y;
        if (v.z < mn.z) mn.z = v.z;
        if (v.z > mx.z) mx.z = v.z;
        mx.x += GI_EPSILON;
        mx.y += GI_EPSILON;
        mx.z += GI_EPSILON;
    }
    return BoundingBox(mn, mx);
}

// ---- mesh generators (entities.h:562-785) ----------------------------------------------------------------------
static dvec3 mix_half(const dvec3& a, const dvec3& b) { return a + 0.5 * (b - a); }  // glm::mix(a,b,.5)

sphereMesh::sphereMesh(Octree* o, dvec3 position, double radius, int subdivs, const Material& m)  // entities.h:568-632
{
    typedef std::array<dvec3, 3> T3;
    std::vector<T3> tris, tmp;
    const dvec3 X(1, 0, 0), Y(0, 1, 0), Z(0, 0, 1), nX(-1, 0, 0), nY(0, -1, 0), nZ(0, 0, -1);
    tris = { T3{ nX, nY, nZ }, T3{ nY, X, nZ }, T3{ X, Y, nZ }, T3{ Y, nX, nZ }, T3{ nX, nY, Z }, T3{ nY, X, Z }, T3{ X, Y, Z }, T3{ Y, nX, Z } };
    for (int j = 0; j < subdivs; j++) {
        for (size_t i = 0; i < tris.size(); i++) {
            dvec3 v1 = gi::normalize(tris[i][0]), v2 = gi::normalize(tris[i][1]), v3 = gi::normalize(tris[i][2]);
            dvec3 a = gi::normalize(mix_half(v1, v2)), b = gi::normalize(mix_half(v2, v3)), c = gi::normalize(mix_half(v1, v3));
            tmp.push_back(T3{ v1, a, c });
            tmp.push_back(T3{ a, v2, b });
            tmp.push_back(T3{ a, b, c });
            tmp.push_back(T3{ c, b, v3 });
        }
        tris = std::move(tmp);
        tmp.clear();
    }
    auto uv = [](const dvec3& p) { return dvec2(.5 * std::acos(p.y) / (M_PI) + .5, .5 * std::atan(p.z / p.x) / (2 * M_PI) + .5); };
    for (size_t i = 0; i < tris.size(); i++) {
        o->push_back(new triangle(vertex(radius * tris[i][0] + position, tris[i][0], uv(tris[i][0])), vertex(radius * tris[i][1] + position, tris[i][1], uv(tris[i][1])),
                                  vertex(radius * tris[i][2] + position, tris[i][2], uv(tris[i][2])), m));
        count++;
    }
}

coneMesh::coneMesh(Octree* o, dvec3 position, dvec3 rotation, double radius, double height, int tris, const Material& m)  // entities.h:651-675
{
    dmat3 rot = gi::eulerAngleXYZ(rotation.x, rotation.y, rotation.z);
    dmat3 baseRot = gi::eulerAngleXYZ(0.0, 0.0, 2 * M_PI / tris);
    dvec3 last(radius, 0, 0);
    for (int i = 0; i < tris; i++) {
        dvec3 next = baseRot * last;
        o->push_back(new triangle(vertex(rot * last + position, rot * last), vertex(rot * next + position, rot * next), vertex(rot * dvec3(0, 0, height) + position, rot * last), m));
        o->push_back(new triangle(vertex(rot * last + position, rot * dvec3(0, 0, -1)), vertex(rot * next + position, rot * dvec3(0, 0, -1)),
                                  vertex(dvec3(0, 0, 0) + position, rot * dvec3(0, 0, -1)), m));
        last = next;
        count += 2;
    }
}

quadMesh::quadMesh(Octree* o, dvec3 v1, dvec3 v2, dvec3 v3, dvec3 v4, const Material& m)  // entities.h:723-727
{
    o->push_back(new triangle(vertex(v1), vertex(v2), vertex(v3), m));
    o->push_back(new triangle(vertex(v3), vertex(v2), vertex(v4), m));
}

boxMesh::boxMesh(Octree* o, dvec3 position, dvec3 size, dvec3 rotation, const Material& m)  // entities.h:744-774
{
    dmat3 rot = gi::eulerAngleXYZ(rotation.x, rotation.y, rotation.z);
    // 12 triangles over the corners (+-1,+-1,+-1); each corner is normalised before scaling (half extent = size/sqrt(3))
    static const int T[12][3][3] = {
        { { -1, -1, -1 }, { -1, 1, -1 }, { 1, -1, -1 } }, { { -1, 1, -1 }, { 1, 1, -1 }, { 1, -1, -1 } },
        { { -1, -1, -1 }, { -1, -1, 1 }, { -1, 1, -1 } }, { { -1, -1, 1 }, { -1, 1, 1 }, { -1, 1, -1 } },
        { { -1, -1, -1 }, { 1, -1, -1 }, { -1, -1, 1 } }, { { -1, -1, 1 }, { 1, -1, -1 }, { 1, -1, 1 } },
        { { -1, -1, 1 }, { 1, -1, 1 }, { -1, 1, 1 } },    { { 1, -1, 1 }, { 1, 1, 1 }, { -1, 1, 1 } },
        { { -1, 1, 1 }, { 1, 1, 1 }, { 1, 1, -1 } },      { { -1, 1, 1 }, { 1, 1, -1 }, { -1, 1, -1 } },
        { { 1, -1, -1 }, { 1, 1, -1 }, { 1, -1, 1 } },    { { 1, -1, 1 }, { 1, 1, -1 }, { 1, 1, 1 } },
    };
    for (int i = 0; i < 12; i++) {
        dvec3 c[3];
        for (int k = 0; k < 3; k++) c[k] = rot * (gi::normalize(dvec3(T[i][k][0], T[i][k][1], T[i][k][2])) * size) + position;
        o->push_back(new triangle(vertex(c[0]), vertex(c[1]), vertex(c[2]), m));
    }
}

// ---- octree ------------------------------------------------------------------------------------------------------
void Octree::push_back(Entity* object)  // octree.cpp:25-38
{
    if (_root._entities.empty()) {
        _root._bbox.max = object->boundingBox().max;
        _root._bbox.min = object->boundingBox().min;
    }
    _root._entities.push_back(object);
    object->_owner = this; object->_id = (uint32_t)_all.size();
    _all.push_back(object);
    _root._bbox.max = gi::vmax(_root._bbox.max, object->boundingBox().max);
    _root._bbox.min = gi::vmin(_root._bbox.min, object->boundingBox().min);
    valid = false;
}

void Octree::push_back(Light* light) { lights.push_back(light); }  // octree.cpp:41-46

void Octree::light_cones(const std::vector<Entity*>& list)  // octree.cpp:60-102
{
    // caustic emission cone per light, including the reference's habit of continuing to accumulate avgPos/count inside
    // the per-light loop (it only affects lights after the first)
    dvec3 avgPos(0, 0, 0);
    double count = 0;
    for (Entity* e : list)
        if (e->material.roughness < 0.1) { avgPos = avgPos + e->boundingBox().center(); count++; }
    if (count > 0) avgPos = avgPos / count;
    for (Light* l : lights) {
        double maxAngle = 0;
        l->dir = gi::normalize(avgPos - l->pos);
        for (Entity* e : list)
            if (e->material.roughness < 0.1) {
                BoundingBox bb = e->boundingBox();
                avgPos = avgPos + bb.center();
                double angle = 1.0 - std::acos(gi::dot(l->dir, gi::normalize(l->pos - bb.min))) / M_PI;
                maxAngle = std::max(maxAngle, angle);
                count++;
            }
        l->angle = maxAngle;
    }
}

void Octree::rebuild()  // octree.cpp:53-119
{
    light_cones(_root._entities);
    _device_built = false;
    if (_root._entities.size() > GI_MAX_ENTITIES_PER_LEAF) _root.partition();  // octree.cpp:106-112
    valid = true;
}

void Octree::entity_boxes(std::vector<double>& out6) const
{
    out6.resize(_all.size() * 6);
    for (size_t i = 0; i < _all.size(); i++) {
        BoundingBox b = _all[i]->boundingBox();
        double* o = &out6[i * 6];
        o[0] = b.min.x; o[1] = b.min.y; o[2] = b.min.z; o[3] = b.max.x; o[4] = b.max.y; o[5] = b.max.z;
    }
}

int Octree::rebuild(gi_ctx* ctx)  // octree.cpp:53-119 with Node::partition on the device
{
    // Built from the insertion-order list of ALL entities.  (Deviation, documented: the reference's second rebuild after a
    // partition sees only the entities pushed since — the root list was cleared, octree.cpp:370 — and leaves them unreachable
    // in the root; here a rebuild always covers every entity.)
    light_cones(_all);
    const size_t n = _all.size();
    std::vector<uint8_t> type(n);
    std::vector<double> geom(n * 9, 0.0), boxes;
    for (size_t i = 0; i < n; i++) {
        const Entity* e = _all[i];
        type[i] = (uint8_t)e->kind();
        double* g = &geom[i * 9];
        if (const triangle* t = dynamic_cast<const triangle*>(e)) { for (int k = 0; k < 3; k++) { g[3 * k] = t->vertices[k].pos.x; g[3 * k + 1] = t->vertices[k].pos.y; g[3 * k + 2] = t->vertices[k].pos.z; } }
        else if (const sphere* s = dynamic_cast<const sphere*>(e)) { g[0] = s->pos.x; g[1] = s->pos.y; g[2] = s->pos.z; g[3] = s->rad; }
        else if (const cone* c = dynamic_cast<const cone*>(e)) { g[0] = c->pos.x; g[1] = c->pos.y; g[2] = c->pos.z; g[3] = c->rad; g[4] = c->height; }
    }
    entity_boxes(boxes);
    const double root[6] = { _root._bbox.min.x, _root._bbox.min.y, _root._bbox.min.z, _root._bbox.max.x, _root._bbox.max.y, _root._bbox.max.z };
    uint32_t nn = 0, nr = 0;
    valid = false;
    int rc = gi_octree_build(ctx, (uint32_t)n, type.data(), geom.data(), boxes.data(), root, &nn, &nr, &last_build_ms);
    if (rc != GI_OK) return rc;
    _dev_box.resize((size_t)nn * 6); _dev_child.resize(nn); _dev_mask.resize(nn); _dev_off.resize(nn); _dev_cnt.resize(nn); _dev_leaf.resize(nr);
    rc = gi_octree_download(ctx, _dev_box.data(), _dev_child.data(), _dev_mask.data(), _dev_off.data(), _dev_cnt.data(), _dev_leaf.data());
    if (rc != GI_OK) return rc;
    _device_built = true;
    valid = true;
    return GI_OK;
}

// One level of Octree::Node::partition (octree.cpp:316-365): make the eight child boxes, hand every entity to each
// child it overlaps, drop empty children.  Returns false when the split "did not improve" (octree.cpp:363-368), in
// which case the children stay but are not subdivided further.
static bool split_once(Octree::Node* n)
{
    const BoundingBox& b = n->_bbox;
    dvec3 mid = b.min + 0.5 * (b.max - b.min);  // glm::mix(min, max, .5)
    const double hx = .5 * b.dx(), hy = .5 * b.dy(), hz = .5 * b.dz();
    // child i: bit0 = +x, bit1 = +z, bit2 = +y.  Low corners are min (+ half extent), high corners are mid (+ half
    // extent) — not the parent's max — exactly as written in the reference (octree.cpp:321-328); child 0 and 7 use
    // (min, mid) and (mid, max) directly.
    for (int i = 0; i < 8; i++) {
        dvec3 lo = b.min, hi = mid;
        if (i == 7) { lo = mid; hi = b.max; }
        else {
            if (i & 1) { lo.x = b.min.x + hx; hi.x = mid.x + hx; }
            if (i & 2) { lo.z = b.min.z + hz; hi.z = mid.z + hz; }
            if (i & 4) { lo.y = b.min.y + hy; hi.y = mid.y + hy; }
        }
        n->_children[i].reset(new Octree::Node(BoundingBox(lo, hi)));
    }
    for (Entity* e : n->_entities) {
        BoundingBox eb = e->boundingBox();
        for (int i = 0; i < 8; i++) {
            Octree::Node* c = n->_children[i].get();
            if (c->_bbox.intersect(eb) && e->intersect(c->_bbox) && eb.dx() > GI_EPSILON) c->_entities.push_back(e);
        }
    }
    double avg = 0;
    for (int i = 0; i < 8; i++) {
        if (n->_children[i]->_entities.empty()) n->_children[i].reset();
        else avg += (double)n->_children[i]->_entities.size();
    }
    avg /= 8;
    bool improved = !(avg > GI_MAX_SUBDIV_RATIO * n->_entities.size());
    n->_entities.clear();
    n->_entities.shrink_to_fit();
    return improved;
}

void Octree::Node::partition()  // octree.cpp:316-384 (the reference recurses; a work list builds the same tree)
{
    std::vector<Node*> work = { this };
    while (!work.empty()) {
        Node* n = work.back();
        work.pop_back();
        if (!split_once(n)) continue;
        for (int i = 0; i < 8; i++) {
            Node* c = n->_children[i].get();
            if (c && c->_entities.size() > GI_MAX_ENTITIES_PER_LEAF && c->_bbox.dx() > GI_MIN_LEAF_SIZE) work.push_back(c);
        }
    }
}

bool Octree::Node::is_leaf() const  // octree.cpp:386-393
{
    for (int i = 0; i < 8; ++i)
        if (_children[i]) return false;
    return true;
}

// ---- flatten: breadth-first SoA image of the tree + primitive / material / texture / light tables ------------------
gi_scene_desc FlatScene::desc() const
{
    gi_scene_desc d;
    std::memset(&d, 0, sizeof(d));
    d.n_nodes = (uint32_t)node_mask.size();
    d.node_box = node_box.data(); d.node_child = node_child.data(); d.node_mask = node_mask.data();
    d.node_prim_off = node_prim_off.data(); d.node_prim_cnt = node_prim_cnt.data();
    d.n_refs = (uint32_t)leaf_prims.size(); d.leaf_prims = leaf_prims.data();
    d.n_prims = (uint32_t)prim_type.size();
    d.prim_type = prim_type.data(); d.prim_geom = prim_geom.data(); d.prim_nrm = prim_nrm.data(); d.prim_uv = prim_uv.data();
    d.prim_fnorm = prim_fnorm.data(); d.prim_mat = prim_mat.data();
    d.n_mats = (uint32_t)mats.size(); d.mats = mats.data();
    d.n_tex = (uint32_t)tex.size(); d.tex = tex.data();
    d.tex_pixel_bytes = tex_pixels.size(); d.tex_pixels = tex_pixels.data();
    d.n_lights = (uint32_t)lights.size(); d.lights = lights.data();
    d.camera = camera;
    for (int i = 0; i < 3; i++) d.ambient[i] = ambient[i];
    d.n_fog = (uint32_t)fogs.size(); d.fogs = fogs.data();
    d.fog_grid_count = fog_grid.size(); d.fog_grid = fog_grid.data();
    return d;
}

// the counter generator of the device (gi_device.cuh: gi_mix64 / gi_rand), restated for the fog noise grid
static uint64_t mix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
HeightFog::HeightFog(dvec3 position, dvec3 size, dvec3 color, double density, double scatter, int noiseScale)
    : AtmosphereEntity(position, size, color, scatter), d(density), nscale(noiseScale), s(size)
{
    const long n = (long)((size.x + 1) * (size.y + 1) * (size.z + 1) * std::pow(noiseScale, 3));   // atmosphere.h:39-41
    uint64_t ent = 0;   // stream id: a hash of the volume's own parameters, so a fog definition always gets the same grid
    for (double v : { position.x, position.y, position.z, size.x, size.y, size.z, density, (double)noiseScale }) { uint64_t b; std::memcpy(&b, &v, 8); ent = mix64(ent ^ b); }
    noiseGrid.reserve(n > 0 ? (size_t)n : 0);
    for (long i = 0; i < n; i++) {
        uint64_t h = mix64(0x5EEDF06ull ^ mix64(ent ^ mix64((uint64_t)i)));
        noiseGrid.push_back((double)(h >> 11) * (1.0 / 9007199254740992.0));
    }
    nscale = 1;   // atmosphere.h:47
}

static void put3(double* p, const dvec3& v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }

void Octree::flatten(const Camera& cam, const dvec3& amb, FlatScene& out) const
{
    out = FlatScene();
    for (const AtmosphereEntity* a : at) {
        const HeightFog* hf = dynamic_cast<const HeightFog*>(a);
        if (!hf) continue;   // the base class has zero density (atmosphere.h:19-22)
        gi_fog g;
        std::memset(&g, 0, sizeof(g));
        put3(g.pos, hf->pos); put3(g.size, hf->s); put3(g.col, hf->col); put3(g.bmin, hf->bbox.min); put3(g.bmax, hf->bbox.max);
        g.density = hf->d; g.scatter = hf->sc;
        g.grid_offset = out.fog_grid.size(); g.grid_count = hf->noiseGrid.size();
        out.fog_grid.insert(out.fog_grid.end(), hf->noiseGrid.begin(), hf->noiseGrid.end());
        out.fogs.push_back(g);
    }
    std::unordered_map<const Entity*, uint32_t> eid;
    for (uint32_t i = 0; i < _all.size(); i++) eid[_all[i]] = i;
    // nodes, breadth-first; existing children contiguous in child order
    std::vector<const Node*> order;
    if (_device_built) {
        out.node_box = _dev_box; out.node_child = _dev_child; out.node_mask = _dev_mask; out.node_prim_off = _dev_off; out.node_prim_cnt = _dev_cnt; out.leaf_prims = _dev_leaf;
    } else order.push_back(&_root);
    for (size_t q = 0; q < order.size(); q++) {
        const Node* n = order[q];
        uint8_t mask = 0;
        uint32_t first = 0;
        for (int i = 0; i < 8; i++)
            if (n->_children[i]) {
                if (!mask) first = (uint32_t)order.size();
                mask |= (uint8_t)(1u << i);
                order.push_back(n->_children[i].get());
            }
        out.node_mask.push_back(mask);
        out.node_child.push_back(first);
        double b[6];
        put3(b, n->_bbox.min); put3(b + 3, n->_bbox.max);
        out.node_box.insert(out.node_box.end(), b, b + 6);
        out.node_prim_off.push_back((uint32_t)out.leaf_prims.size());
        out.node_prim_cnt.push_back((uint32_t)n->_entities.size());
        for (Entity* e : n->_entities) out.leaf_prims.push_back(eid.at(e));
    }
    // primitives; materials are de-duplicated by value, textures by identity
    std::unordered_map<const texture*, uint32_t> tid;
    auto tex_id = [&](const texture* t) -> uint32_t {
        auto it = tid.find(t);
        if (it != tid.end()) return it->second;
        gi_texture g;
        std::memset(&g, 0, sizeof(g));
        g.kind = t->kind();
        put3(g.a, t->color);
        if (const checkerboard* c = dynamic_cast<const checkerboard*>(t)) { put3(g.a, c->a); put3(g.b, c->b); g.tiles = c->tiles; }
        if (const imageTexture* im = dynamic_cast<const imageTexture*>(t)) {
            g.tile_u = im->tile.x; g.tile_v = im->tile.y; g.width = im->width; g.height = im->height; g.has_alpha = im->has_alpha ? 1 : 0;
            g.pixel_offset = out.tex_pixels.size();
            out.tex_pixels.insert(out.tex_pixels.end(), im->rgba.begin(), im->rgba.end());
        }
        uint32_t id = (uint32_t)out.tex.size();
        out.tex.push_back(g);
        tid[t] = id;
        return id;
    };
    auto mat_id = [&](const Material& m) -> uint32_t {
        gi_material g;
        std::memset(&g, 0, sizeof(g));
        g.diffuse_tex = tex_id(m.diffuse); g.emissive_tex = tex_id(m.emissive);
        g.roughness = m.roughness; g.opacity = m.opacity; g.ior = m.IOR;
        for (uint32_t i = 0; i < out.mats.size(); i++)
            if (std::memcmp(&out.mats[i], &g, sizeof(g)) == 0) return i;
        out.mats.push_back(g);
        return (uint32_t)out.mats.size() - 1;
    };
    size_t np = _all.size();
    out.prim_type.resize(np); out.prim_mat.resize(np);
    out.prim_geom.assign(np * 9, 0.0); out.prim_nrm.assign(np * 9, 0.0); out.prim_uv.assign(np * 6, 0.0); out.prim_fnorm.assign(np * 3, 0.0);
    for (size_t i = 0; i < np; i++) {
        const Entity* e = _all[i];
        out.prim_type[i] = (uint8_t)e->kind();
        out.prim_mat[i] = mat_id(e->material);
        double* g = &out.prim_geom[i * 9];
        double* nn = &out.prim_nrm[i * 9];
        if (const triangle* t = dynamic_cast<const triangle*>(e)) {
            for (int k = 0; k < 3; k++) {
                put3(g + 3 * k, t->vertices[k].pos); put3(nn + 3 * k, t->vertices[k].norm);
                out.prim_uv[i * 6 + 2 * k] = t->vertices[k].texCoord.x; out.prim_uv[i * 6 + 2 * k + 1] = t->vertices[k].texCoord.y;
            }
            put3(&out.prim_fnorm[i * 3], t->norm);
        } else if (const sphere* s = dynamic_cast<const sphere*>(e)) {
            put3(g, s->pos); g[3] = s->rad;
        } else if (const cone* c = dynamic_cast<const cone*>(e)) {
            put3(g, c->pos); g[3] = c->rad; g[4] = c->height;
            for (int col = 0; col < 3; col++) for (int row = 0; row < 3; row++) nn[col * 3 + row] = c->rot.c[col][row];
        }
    }
    for (const Light* l : lights) {
        gi_light g;
        put3(g.pos, l->pos); put3(g.col, l->col); g.rad = l->rad; put3(g.dir, l->dir); g.angle = l->angle;
        out.lights.push_back(g);
    }
    put3(out.camera.pos, cam.pos); put3(out.camera.forward, cam.forward); put3(out.camera.up, cam.up); put3(out.camera.right, cam.right);
    out.camera.sensor_diag = cam.sensorDiag; out.camera.focal_dist = cam.focalDist;
    put3(out.ambient, amb);
}

bool Image::writePPM(const char* path) const
{
    FILE* f = std::fopen(path, "wb");
    if (!f) return false;
    std::fprintf(f, "P6\n%d %d\n255\n", _w, _h);
    std::fwrite(rgb.data(), 1, rgb.size(), f);
    std::fclose(f);
    return true;
}

// ---- the reference's query members as batch-of-one device calls (octree.h:54,56; entities.h:24) -----------------------------------------------
void Octree::attach(gi_ctx* ctx, const FlatScene& flat)
{
    _query_ctx = ctx;
    _query_nodes.clear();
    const size_t nn = flat.node_mask.size();
    _query_nodes.reserve(nn);
    for (size_t i = 0; i < nn; i++) {
        const double* b = &flat.node_box[6 * i];
        std::unique_ptr<Node> n(new Node(BoundingBox(dvec3(b[0], b[1], b[2]), dvec3(b[3], b[4], b[5]))));
        if (!flat.node_mask[i]) for (uint32_t k = 0; k < flat.node_prim_cnt[i]; k++) n->_entities.push_back(_all[flat.leaf_prims[flat.node_prim_off[i] + k]]);
        _query_nodes.push_back(std::move(n));
    }
}

std::vector<Entity*> Octree::intersect(const Ray& ray, double tmin, double tmax) const
{
    std::vector<Entity*> res;
    if (!_query_ctx) return res;
    const double o[3] = { ray.origin.x, ray.origin.y, ray.origin.z }, d[3] = { ray.dir.x, ray.dir.y, ray.dir.z };
    uint32_t cap = 256, count = 0;   // res.reserve(256) in the reference (octree.cpp:153)
    std::vector<uint32_t> ids(cap);
    if (gi_octree_intersect(_query_ctx, 1, o, d, &tmin, &tmax, cap, ids.data(), &count) != GI_OK) return res;
    if (count > cap) { cap = count; ids.resize(cap); if (gi_octree_intersect(_query_ctx, 1, o, d, &tmin, &tmax, cap, ids.data(), &count) != GI_OK) return res; }
    res.reserve(count);
    for (uint32_t k = 0; k < count; k++) res.push_back(_all[ids[k]]);
    return res;
}

std::vector<std::pair<const Octree::Node*, double>> Octree::intersectSorted(const Ray& ray, double tmin, double tmax) const
{
    std::vector<std::pair<const Node*, double>> res;
    if (!_query_ctx) return res;
    const double o[3] = { ray.origin.x, ray.origin.y, ray.origin.z }, d[3] = { ray.dir.x, ray.dir.y, ray.dir.z };
    uint32_t cap = 64, count = 0;
    std::vector<uint32_t> nodes(cap); std::vector<double> t0(cap);
    if (gi_octree_intersect_sorted(_query_ctx, 1, o, d, &tmin, &tmax, cap, nodes.data(), t0.data(), &count) != GI_OK) return res;
    if (count > cap) { cap = count; nodes.resize(cap); t0.resize(cap); if (gi_octree_intersect_sorted(_query_ctx, 1, o, d, &tmin, &tmax, cap, nodes.data(), t0.data(), &count) != GI_OK) return res; }
    for (uint32_t k = 0; k < count; k++) if (nodes[k] < _query_nodes.size()) res.emplace_back(_query_nodes[nodes[k]].get(), t0[k]);
    return res;
}

bool Entity::intersect(const Ray& ray, dvec3& hit, dvec3& norm, gi::dvec2& uv) const
{
    gi_ctx* ctx = _owner ? _owner->attached() : nullptr;
    if (!ctx) return false;
    const double o[3] = { ray.origin.x, ray.origin.y, ray.origin.z }, d[3] = { ray.dir.x, ray.dir.y, ray.dir.z };
    uint8_t ok = 0, wrote = 0;
    double h[3], n[3], t[2];
    if (gi_prim_intersect(ctx, 1, &_id, o, d, &ok, h, n, t, &wrote) != GI_OK || !ok) return false;
    hit = dvec3(h[0], h[1], h[2]); norm = dvec3(n[0], n[1], n[2]);
    if (wrote) uv = gi::dvec2(t[0], t[1]);
    return true;
}

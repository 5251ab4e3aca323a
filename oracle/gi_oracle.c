/* gi_oracle.c — TEST INFRASTRUCTURE ONLY: plain-C CPU restatement of the reference hot path.  See gi_oracle.h for
 * the role and the parity status (PINNED against oracle/_ref, the reference itself compiled from /root/reference).
 * Build: oracle/Makefile (gcc -O2 -ffp-contract=off, no -march: fp64 operation order must match the reference's
 * -O2 x86-64 build).  Every function cites the reference file:line it follows.  Vector helpers restate the few glm
 * 0.9.8.2 functions the path uses with glm's evaluation order. */
#include "gi_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define GO_EPSILON 0.00001      /* util.h:18 */
#define GO_SHADOW_BIAS 0.0001   /* util.h:20 */
#define GO_MAX_PHOTONS_PER_LEAF 16 /* util.h:15 */
#define GO_PI 3.14159265358979323846 /* M_PI from <math.h> (util.h:11 only defines it if missing) */

typedef struct { double x, y, z; } v3;

static inline v3 V(double x, double y, double z) { v3 r = { x, y, z }; return r; }
static inline v3 ld3(const double* p) { return V(p[0], p[1], p[2]); }
static inline void st3(double* p, v3 a) { p[0] = a.x; p[1] = a.y; p[2] = a.z; }
static inline v3 add(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 sub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 mul(v3 a, v3 b) { return V(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline v3 scale(v3 a, double s) { return V(a.x * s, a.y * s, a.z * s); }
/* glm::dot: x*x' + y*y' + z*z' summed left to right (glm/detail/func_geometric.inl:54-60) */
static inline double dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
/* glm::cross (func_geometric.inl:74-85) */
static inline v3 cross(v3 x, v3 y) { return V(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y); }
/* glm::normalize = v * inversesqrt(dot(v,v)), inversesqrt = 1/sqrt (func_geometric.inl:88-95, func_exponential.inl:128-133) */
static inline v3 normalize(v3 a) { return scale(a, 1.0 / sqrt(dot(a, a))); }
static inline double length3(v3 a) { return sqrt(dot(a, a)); }
/* glm::reflect = I - N*dot(N,I)*2 (func_geometric.inl:110-115) */
static inline v3 reflect3(v3 I, v3 N) { double d = dot(N, I); return sub(I, scale(scale(N, d), 2.0)); }
/* vecLengthSquared (util.h:35-38) */
static inline double len2(v3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }

/* ------------------------------------------------------------------------------------------------------------
 * counter-based PRNG (specification shared with the CUDA path; replaces util.h:52-80 drand)
 * ---------------------------------------------------------------------------------------------------------- */
static inline uint64_t mix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
double go_rand(uint64_t seed, uint64_t path, uint64_t depth, uint64_t site)
{
    uint64_t h = mix64(seed ^ mix64(path ^ mix64(depth ^ mix64(site))));
    return (double)(h >> 11) * (1.0 / 9007199254740992.0);
}
/* Reference-PRNG replay (checker of the checker): the reference's drand() is a thread_local xorshift64* seeded with
 * time(0) (util.h:52-80).  gi_ref runs with time() interposed and OMP_NUM_THREADS=1, so its stream is known; when g_xs is
 * set, the alpha cut-out draws of trace()/visible() below are taken from THAT stream in the reference's own order —
 * one draw at the debug guard of trace() (raytracer.h:438) and one per geometric hit (`drand() < getAlpha(uv) || IOR != 1`,
 * raytracer.h:455 and :297: drand() is evaluated first, always) — which makes the ids of an alpha-textured scene
 * bit-comparable with the reference.  Only go_trace_closest_replay / go_trace_any_replay set it (sequential loops). */
static _Thread_local uint64_t* g_xs = NULL;
static inline double xs_next(uint64_t* st)
{
    uint64_t x = *st;
    x ^= x >> 12; x ^= x << 25; x ^= x >> 27;
    *st = x;
    return (double)(x * 2685821657736338717ull) / 18446744073709551615.0;   /* / (double)ULLONG_MAX == 2^64 */
}
/* sites */
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
#define SITE_FOG_RAD 11ull    /* raymarch draws: counter = step (radiance), light<<32|step (visible), step (photons) */
#define SITE_FOG_SHADOW 12ull
#define SITE_FOG_PHOTON 13ull
#define SITE_PH_FOG_U 14ull
#define SITE_PH_FOG_V 15ull
#define SITE(s, c) (((uint64_t)(s) << 56) | (uint64_t)(c))
#define PHOTON_PATH_BIT (1ull << 63)

/* ------------------------------------------------------------------------------------------------------------
 * Halton sampler (halton_sampler.h).  The generated header hard-codes, for each of the first 256 primes B, a
 * digit-block size k (largest k with B^k <= 500: 5,3,3,2,2,2,2 for 3..19, then 1), a block count n (largest n
 * with B^(k n) < 2^32) and the scale float(0.9999998807907104 / B^(k n)); this restatement derives the same
 * numbers from that rule (checked against all 256 generated functions through the known-answer table).
 * ---------------------------------------------------------------------------------------------------------- */
#define GO_NDIMS 256
static uint32_t g_primes[GO_NDIMS];
static go_halton_dim g_dims[GO_NDIMS];
static uint16_t* g_tables = NULL;
static size_t g_table_entries = 0;
static int g_halton_ready = 0;

/* Halton_sampler::invert (halton_sampler.h:890-900) */
static uint16_t invert_block(uint32_t base, uint32_t digits, uint32_t index, const uint16_t* perm)
{
    uint32_t result = 0;
    for (uint32_t i = 0; i < digits; ++i) {
        result = result * base + perm[index % base];
        index /= base;
    }
    return (uint16_t)result;
}

void go_halton_init(void)
{
#pragma omp critical(go_halton)
    if (!g_halton_ready) {
        /* first 256 primes: 2 .. 1619 */
        uint32_t np = 0;
        for (uint32_t c = 2; np < GO_NDIMS; c++) {
            int prime = 1;
            for (uint32_t d = 2; d * d <= c; d++) if (c % d == 0) { prime = 0; break; }
            if (prime) g_primes[np++] = c;
        }
        /* Faure permutations, built recursively from base 2 upward (halton_sampler.h:573-603) */
        const uint32_t max_base = 1619;
        uint16_t** perms = (uint16_t**)calloc(max_base + 1, sizeof(uint16_t*));
        for (uint32_t k = 1; k <= 3; k++) {
            perms[k] = (uint16_t*)malloc(k * sizeof(uint16_t));
            for (uint32_t i = 0; i < k; i++) perms[k][i] = (uint16_t)i;
        }
        for (uint32_t base = 4; base <= max_base; base++) {
            perms[base] = (uint16_t*)malloc(base * sizeof(uint16_t));
            uint32_t b = base / 2;
            if (base & 1) {
                for (uint32_t i = 0; i < base - 1; i++) perms[base][i + (i >= b)] = (uint16_t)(perms[base - 1][i] + (perms[base - 1][i] >= b));
                perms[base][b] = (uint16_t)b;
            } else {
                for (uint32_t i = 0; i < b; i++) {
                    perms[base][i] = (uint16_t)(2 * perms[b][i]);
                    perms[base][b + i] = (uint16_t)(2 * perms[b][i] + 1);
                }
            }
        }
        /* block tables (halton_sampler.h:902-1414) */
        size_t total = 0;
        for (int d = 0; d < GO_NDIMS; d++) {
            uint32_t B = g_primes[d];
            uint32_t k = 1; uint64_t bk = B;
            while (bk * B <= 500) { bk *= B; k++; }
            uint32_t n = 1; uint64_t pw = bk;
            while (pw * bk <= 0xFFFFFFFFull) { pw *= bk; n++; }
            g_dims[d].base = B; g_dims[d].block = (uint32_t)bk; g_dims[d].nblocks = n; g_dims[d].table_off = (uint32_t)total;
            g_dims[d].scale = (float)(0.9999998807907104 / (double)pw);
            total += (d == 0) ? 0 : bk;
            (void)k;
        }
        g_tables = (uint16_t*)malloc(total * sizeof(uint16_t));
        g_table_entries = total;
        for (int d = 1; d < GO_NDIMS; d++) {
            uint32_t B = g_dims[d].base, bk = g_dims[d].block;
            uint32_t k = 0; for (uint32_t t = bk; t > 1; t /= B) k++;
            for (uint32_t i = 0; i < bk; i++) g_tables[g_dims[d].table_off + i] = invert_block(B, k, i, perms[B]);
        }
        for (uint32_t b = 1; b <= max_base; b++) free(perms[b]);
        free(perms);
        g_halton_ready = 1;
    }
}

const uint16_t* go_halton_tables(size_t* n_entries) { go_halton_init(); if (n_entries) *n_entries = g_table_entries; return g_tables; }
const go_halton_dim* go_halton_dims(void) { go_halton_init(); return g_dims; }

/* halton2 (halton_sampler.h:1417-1431): bit reversal written into the mantissa */
static float halton2(uint32_t index)
{
    index = (index << 16) | (index >> 16);
    index = ((index & 0x00ff00ffu) << 8) | ((index & 0xff00ff00u) >> 8);
    index = ((index & 0x0f0f0f0fu) << 4) | ((index & 0xf0f0f0f0u) >> 4);
    index = ((index & 0x33333333u) << 2) | ((index & 0xccccccccu) >> 2);
    index = ((index & 0x55555555u) << 1) | ((index & 0xaaaaaaaau) >> 1);
    union { uint32_t u; float f; } r;
    r.u = 0x3f800000u | (index >> 9);
    return r.f - 1.f;
}

float go_halton_sample(uint32_t dim, uint32_t index)
{
    if (!g_halton_ready) go_halton_init();
    if (dim == 0) return halton2(index);
    if (dim >= GO_NDIMS) return 0.f; /* the reference falls back to rand() (halton_sampler.h:887); never reached with MAX_DEPTH 64 */
    const go_halton_dim* D = &g_dims[dim];
    const uint16_t* T = g_tables + D->table_off;
    /* e.g. halton3 (halton_sampler.h:1433-1439): sum_j perm[(index / block^j) % block] * block^(n-1-j), u32 arithmetic */
    uint32_t sum = 0, idx = index;
    uint32_t mult[8]; mult[D->nblocks - 1] = 1;
    for (int j = (int)D->nblocks - 2; j >= 0; j--) mult[j] = mult[j + 1] * D->block;
    for (uint32_t j = 0; j < D->nblocks; j++) {
        sum += (uint32_t)T[idx % D->block] * mult[j];
        idx /= D->block;
    }
    return (float)sum * D->scale;
}

/* ------------------------------------------------------------------------------------------------------------
 * Halton_enum (halton_enum.h:69-155)
 * ---------------------------------------------------------------------------------------------------------- */
static void ext_euclid(int a, int b, int* s, int* t) /* halton_enum.h:126-134 */
{
    if (!b) { *s = 1; *t = 0; return; }
    int q = a / b, r = a % b, s1, t1;
    ext_euclid(b, r, &s1, &t1);
    *s = t1; *t = s1 - q * t1;
}
void go_henum_init(go_henum* he, uint32_t width, uint32_t height)
{
    he->w = width; he->h = height;
    he->p2 = 0; uint32_t w = 1; while (w < width) { ++he->p2; w *= 2; }
    he->scale_x = (float)w;
    he->p3 = 0; uint32_t h = 1; while (h < height) { ++he->p3; h *= 3; }
    he->scale_y = (float)h;
    he->inc = w * h;
    int i1, i2; ext_euclid((int)h, (int)w, &i1, &i2);
    uint32_t inv2 = (i1 < 0) ? (uint32_t)(i1 + (int)w) : (uint32_t)(i1 % (int)w);
    uint32_t inv3 = (i2 < 0) ? (uint32_t)(i2 + (int)h) : (uint32_t)(i2 % (int)h);
    he->mx = h * inv2; he->my = w * inv3;
}
static uint32_t halton2_inverse(uint32_t index, uint32_t digits) /* halton_enum.h:136-144 */
{
    index = (index << 16) | (index >> 16);
    index = ((index & 0x00ff00ffu) << 8) | ((index & 0xff00ff00u) >> 8);
    index = ((index & 0x0f0f0f0fu) << 4) | ((index & 0xf0f0f0f0u) >> 4);
    index = ((index & 0x33333333u) << 2) | ((index & 0xccccccccu) >> 2);
    index = ((index & 0x55555555u) << 1) | ((index & 0xaaaaaaaau) >> 1);
    return digits ? index >> (32 - digits) : 0; /* digits==0 only for width 1 (x is then 0) */
}
static uint32_t halton3_inverse(uint32_t index, uint32_t digits) /* halton_enum.h:146-155 */
{
    uint32_t result = 0;
    for (uint32_t d = 0; d < digits; ++d) { result = result * 3 + index % 3; index /= 3; }
    return result;
}
uint32_t go_henum_index(const go_henum* he, uint32_t s, uint32_t x, uint32_t y) /* halton_enum.h:106-114 */
{
    uint64_t hx = halton2_inverse(x, he->p2), hy = halton3_inverse(y, he->p3);
    uint32_t offset = (uint32_t)((hx * he->mx + hy * he->my) % he->inc);
    return offset + s * he->inc; /* u32 wrap-around is the reference's behaviour */
}

/* ------------------------------------------------------------------------------------------------------------
 * camera rays (raytracer.h:74-78, 112-129; ray.h:7-17)
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct { v3 o, d, inv; } ray_t;
static inline ray_t make_ray(v3 o, v3 d) /* Ray::Ray -> setDir: dir = normalize(d); invDir = 1/dir */
{
    ray_t r; r.o = o; r.d = normalize(d);
    r.inv = V(1.0 / r.d.x, 1.0 / r.d.y, 1.0 / r.d.z);
    return r;
}
static inline ray_t ray_as_stored(v3 o, v3 dir) /* a Ray whose `dir` member is given as stored (no second normalisation) */
{
    ray_t r; r.o = o; r.d = dir;
    r.inv = V(1.0 / r.d.x, 1.0 / r.d.y, 1.0 / r.d.z);
    return r;
}
void go_camera_ray(const gi_camera* cam, const go_henum* he, int w, int h, int x, int y, int s, double org[3], double dir[3], uint32_t* index)
{
    double halfW = (cam->sensor_diag * w) / (sqrt((double)w * w + h * h));
    double halfH = halfW * ((double)h / w);
    v3 pos = ld3(cam->pos), fwd = ld3(cam->forward), up = ld3(cam->up);
    v3 center = add(pos, scale(fwd, cam->focal_dist));
    v3 right = normalize(cross(fwd, up));
    int idx = (int)go_henum_index(he, (uint32_t)s, (uint32_t)x, (uint32_t)y);
    double xr = go_halton_sample(0, (uint32_t)idx);
    double yr = go_halton_sample(1, (uint32_t)idx);
    double dx = (float)((float)xr * he->scale_x); /* Halton_enum::scale_x takes and returns float */
    double dy = (float)((float)yr * he->scale_y);
    v3 pixelPos = sub(add(center, scale(right, halfW * (dx / w - .5))), scale(up, halfH * (dy / h - .5)));
    /* FOCAL_BLUR = 0: eyePos = pos + 0*(xr-.5)*right + 0*(yr-.5)*up, kept as written so that -0/NaN behave alike */
    v3 eye = add(add(pos, scale(right, 0 * (xr - .5))), scale(up, 0 * (yr - .5)));
    ray_t r = make_ray(eye, normalize(sub(pixelPos, eye)));
    st3(org, r.o); st3(dir, r.d);
    if (index) *index = (uint32_t)idx;
}
void go_camera_rays(const gi_camera* cam, int w, int h, int x0, int y0, int x1, int y1, int s0, int s1, double* org, double* dir, uint32_t* index)
{
    go_henum he; go_henum_init(&he, (uint32_t)w, (uint32_t)h); go_halton_init();
    long tw = x1 - x0, th = y1 - y0;
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)(s1 - s0) * tw * th; i++) {
        int s = s0 + (int)(i / (tw * th)); long p = i % (tw * th);
        int y = y0 + (int)(p / tw), x = x0 + (int)(p % tw);
        go_camera_ray(cam, &he, w, h, x, y, s, org + 3 * i, dir + 3 * i, index ? index + i : NULL);
    }
}

/* ------------------------------------------------------------------------------------------------------------
 * boxes (bbox.h)
 * ---------------------------------------------------------------------------------------------------------- */
/* BoundingBox::intersect(ray, tmin, tmax, t0, t1) (bbox.h:47-73) */
static inline int box_hit(const double* b, const ray_t* r, double tmin, double tmax, double* tout)
{
    const double* o = &r->o.x; const double* inv = &r->inv.x;
    for (int i = 0; i < 3; i++) {
        double t0 = (b[i] - o[i]) * inv[i];
        double t1 = (b[3 + i] - o[i]) * inv[i];
        if (inv[i] < 0.0) { double tmp = t0; t0 = t1; t1 = tmp; }
        tmin = t0 > tmin ? t0 : tmin;
        tmax = t1 < tmax ? t1 : tmax;
        if (tmax <= tmin) return 0;
    }
    if (tout) *tout = tmin;
    return 1;
}
/* BoundingBox::contains (bbox.h:41-44), half-open */
static inline int box_contains(const double* b, v3 p)
{
    return p.x >= b[0] && p.y >= b[1] && p.z >= b[2] && p.x < b[3] && p.y < b[4] && p.z < b[5];
}
/* BoundingBox::intersect(other) (bbox.h:33-38), closed */
static inline int box_overlap(const double* a, const double* o)
{
    return (a[0] <= o[3] && a[3] >= o[0]) && (a[1] <= o[4] && a[4] >= o[1]) && (a[2] <= o[5] && a[5] >= o[2]);
}

/* ------------------------------------------------------------------------------------------------------------
 * primitives (entities.h)
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct { v3 p, n; double u, v; } hit_t; /* u,v only written when the primitive writes uv */
/* ray parameter of the last successful primitive test of this thread (what the device keeps as DHit::t); only the pruned
 * canonical traversal below reads it */
static __thread double g_last_t;

/* triangle::intersect(ray, hit, normal, uv) (entities.h:443-490) */
static inline int tri_hit(const gi_scene_desc* sc, uint32_t id, const ray_t* r, v3* hp, v3* hn, double* huv)
{
    const double* g = sc->prim_geom + 9 * (size_t)id;
    v3 v0 = ld3(g), v1 = ld3(g + 3), v2 = ld3(g + 6);
    v3 edge1 = sub(v1, v0), edge2 = sub(v2, v0);
    v3 p = cross(r->d, edge2);
    double det = dot(edge1, p);
    if (det < GO_EPSILON && det > -GO_EPSILON) return 0;
    double inv_det = 1.0 / det;
    v3 tvec = sub(r->o, v0);
    double u = dot(tvec, p) * inv_det;
    if (u < 0 || u > 1) return 0;
    v3 q = cross(tvec, edge1);
    double v = dot(r->d, q) * inv_det;
    if (v < 0 || u + v > 1) return 0;
    double t = dot(edge2, q) * inv_det;
    if (t <= 0) return 0;
    g_last_t = t;
    *hp = add(r->o, scale(r->d, t));
    const double* nn = sc->prim_nrm + 9 * (size_t)id;
    v3 n0 = ld3(nn), n1 = ld3(nn + 3), n2 = ld3(nn + 6);
    if (len2(n0) > 0 && len2(n1) > 0 && len2(n2) > 0) {
        double w = 1 - u - v;
        *hn = add(add(scale(n0, w), scale(n1, u)), scale(n2, v));
        const double* t2 = sc->prim_uv + 6 * (size_t)id;
        huv[0] = (w * t2[0] + u * t2[2]) + v * t2[4];
        huv[1] = (w * t2[1] + u * t2[3]) + v * t2[5];
    } else {
        *hn = ld3(sc->prim_fnorm + 3 * (size_t)id);
    }
    return 1;
}

/* sphere::intersect (entities.h:60-101) */
static inline int sphere_hit(const gi_scene_desc* sc, uint32_t id, const ray_t* r, v3* hp, v3* hn, double* huv)
{
    const double* g = sc->prim_geom + 9 * (size_t)id;
    v3 pos = ld3(g); double rad = g[3];
    v3 oc = sub(r->o, pos);
    double d = dot(r->d, oc);
    double rr = (pow(d, 2) - len2(oc) + pow(rad, 2));
    if (rr < 0) return 0;
    double sr = sqrt(rr);
    double t_1 = -1 * d - sr;
    double t_2 = -1 * d + sr;
    if (t_1 < 0 && t_2 < 0) return 0;
    v3 ip;
    if ((t_1 < t_2 && t_1 > 0) || t_2 < 0) { ip = add(r->o, scale(r->d, t_1)); g_last_t = t_1; }
    else { ip = add(r->o, scale(r->d, t_2)); g_last_t = t_2; }
    *hp = ip;
    *hn = normalize(sub(ip, pos));
    v3 dd = V((pos.x - ip.x) / rad, (pos.y - ip.y) / rad, (pos.z - ip.z) / rad);
    double vv = .5 + asin(dd.y) / GO_PI;
    double uu = .5 + atan2(dd.z, dd.x) / (2 * GO_PI);
    huv[0] = uu; huv[1] = vv;
    return 1;
}

/* cone::intersect (entities.h:158-258).  prim_nrm holds cone::rot (the inverse Euler rotation), column-major;
 * `v*rot` is glm's row-vector product (type_mat3x3.inl:437-443). */
static inline v3 vec_mat(v3 v, const double* m)
{
    return V(m[0] * v.x + m[1] * v.y + m[2] * v.z, m[3] * v.x + m[4] * v.y + m[5] * v.z, m[6] * v.x + m[7] * v.y + m[8] * v.z);
}
static inline int cone_hit(const gi_scene_desc* sc, uint32_t id, const ray_t* r, v3* hp, v3* hn)
{
    const double* g = sc->prim_geom + 9 * (size_t)id;
    const double* rot = sc->prim_nrm + 9 * (size_t)id;
    v3 pos = ld3(g); double rad = g[3], height = g[4];
    v3 origin = vec_mat(sub(r->o, pos), rot);
    v3 dir = vec_mat(r->d, rot);
    double t_1, t_2, thit, phi;
    double phiMax = 2 * GO_PI;
    v3 phit;
    double k = pow(rad / height, 2);
    double A = dir.x * dir.x + dir.y * dir.y - k * dir.z * dir.z;
    double B = 2 * (dir.x * origin.x + dir.y * origin.y - k * dir.z * (origin.z - height));
    double C = origin.x * origin.x + origin.y * origin.y - k * (origin.z - height) * (origin.z - height);
    double discrim = B * B - 4.f * A * C;
    if (discrim < 0) return 0;
    double rootDiscrim = sqrt(discrim);
    double q;
    if (B < 0) q = -.5f * (B - rootDiscrim);
    else q = -.5f * (B + rootDiscrim);
    t_1 = q / A;
    t_2 = C / q;
    if (t_1 < 0 && t_2 < 0) return 0;
    if (t_1 > t_2) { double tmp = t_1; t_1 = t_2; t_2 = tmp; }
    thit = t_1;
    if (t_1 < 0) thit = t_2;
    else if (t_2 < 0) thit = t_1;
    phit = add(origin, scale(dir, thit));
    phi = atan2(phit.y, phit.x);
    if (phi < 0.) phi += 2.f * GO_PI;
    if (phit.z < 0 || phit.z > height || phi > phiMax) {
        if (thit == t_2) return 0;
        thit = t_2;
        phit = add(origin, scale(dir, thit));
        phi = atan2(phit.y, phit.x);
        if (phi < 0.) phi += 2.f * GO_PI;
        if (phit.z < 0 || phit.z > height || phi > phiMax) return 0;
    }
    g_last_t = thit;
    *hp = add(r->o, scale(r->d, thit));
    double vpar = phit.z / height;
    v3 dpdu = V(-phiMax * phit.y, phiMax * phit.x, 0);
    v3 dpdv = V(-phit.x / (1.f - vpar), -phit.y / (1.f - vpar), height);
    *hn = normalize(cross(dpdu, dpdv));
    return 1;
}

static inline int prim_hit(const gi_scene_desc* sc, uint32_t id, const ray_t* r, v3* hp, v3* hn, double* huv)
{
    switch (sc->prim_type[id]) {
    case GI_PRIM_TRIANGLE: return tri_hit(sc, id, r, hp, hn, huv);
    case GI_PRIM_SPHERE: return sphere_hit(sc, id, r, hp, hn, huv);
    case GI_PRIM_CONE: return cone_hit(sc, id, r, hp, hn);
    default: return 0;
    }
}

/* ------------------------------------------------------------------------------------------------------------
 * textures / materials (material.h)
 * ---------------------------------------------------------------------------------------------------------- */
static inline void tex_pixel(const gi_scene_desc* sc, const gi_texture* t, const double* uv, const uint8_t** px)
{
    /* imageTexture::get / getAlpha (material.h:63-81): nearest neighbour, integer modulo, vertical flip */
    int x = abs((int)(uv[0] * t->width * t->tile_u) % t->width);
    int y = t->height - abs((int)(uv[1] * t->height * t->tile_v) % t->height) - 1;
    *px = sc->tex_pixels + t->pixel_offset + ((size_t)y * t->width + x) * 4;
}
static v3 tex_get(const gi_scene_desc* sc, uint32_t id, const double* uv)
{
    const gi_texture* t = &sc->tex[id];
    if (t->kind == GI_TEX_CONST) return ld3(t->a);                                  /* material.h:18-21 */
    if (t->kind == GI_TEX_CHECKER) {                                                /* material.h:39-45 */
        if ((((int)(uv[0] * t->tiles) % 2 == 0) ^ ((int)(uv[1] * t->tiles) % 2 == 0))) return ld3(t->a);
        return ld3(t->b);
    }
    const uint8_t* p; tex_pixel(sc, t, uv, &p);
    /* gamma(c/255, 1/GAMMA) = pow(c, 1/(1/2.2)) (material.h:67, util.h:94-97) */
    double g = 1.0 / 2.2;
    return V(pow(p[0] / 255.0, 1.0 / g), pow(p[1] / 255.0, 1.0 / g), pow(p[2] / 255.0, 1.0 / g));
}
static double tex_alpha(const gi_scene_desc* sc, uint32_t id, const double* uv)
{
    const gi_texture* t = &sc->tex[id];
    if (t->kind != GI_TEX_IMAGE || !t->has_alpha) return 1;                          /* material.h:23-26,70-73 */
    const uint8_t* p; tex_pixel(sc, t, uv, &p);
    return p[3] / 255.0;
}
/* Material::getAlpha (material.h:90-93) */
static inline double mat_alpha(const gi_scene_desc* sc, const gi_material* m, const double* uv) { return m->opacity * tex_alpha(sc, m->diffuse_tex, uv); }

void go_material_eval(const gi_scene_desc* sc, size_t n, const uint32_t* prim, const double* uv, double* diffuse, double* emissive, double* alpha)
{
    for (size_t i = 0; i < n; i++) {
        const gi_material* m = &sc->mats[sc->prim_mat[prim[i]]];
        st3(diffuse + 3 * i, tex_get(sc, m->diffuse_tex, uv + 2 * i));
        st3(emissive + 3 * i, tex_get(sc, m->emissive_tex, uv + 2 * i));
        alpha[i] = mat_alpha(sc, m, uv + 2 * i);
    }
}

/* alpha cut-out decision `drand() < getAlpha(uv) || IOR != 1` (raytracer.h:455, :297).  The draw is keyed by the
 * (leaf node, primitive) occurrence, so duplicates of a primitive in several leaves draw independently like the
 * reference, while the outcome does not depend on visiting order. */
static __thread int g_alpha_frac;   /* the last alpha_pass rejected a candidate whose alpha is fractional (0 < a < 1): another occurrence of it may pass */
static inline int alpha_pass(const gi_scene_desc* sc, uint32_t prim, uint32_t node, const double* uv, uint64_t seed, uint64_t path, uint64_t depth, uint64_t site)
{
    const gi_material* m = &sc->mats[sc->prim_mat[prim]];
    g_alpha_frac = 0;
    if (g_xs) { double r = xs_next(g_xs); return r < mat_alpha(sc, m, uv) || m->ior != 1; }   /* replay: the reference's stream and order */
    if (m->ior != 1) return 1;
    double a = mat_alpha(sc, m, uv);
    if (a >= 1.0) return 1; /* drand() < 1 holds for every draw of this generator ([0,1)) */
    const int pass = go_rand(seed, path, depth, SITE(site, ((uint64_t)node << 28) ^ prim)) < a;
    g_alpha_frac = !pass && a > 0.0;
    return pass;
}

/* ------------------------------------------------------------------------------------------------------------
 * closest hit: RayTracer::trace (raytracer.h:382-478) over Octree::intersectSorted (octree.cpp:188-211,285-313)
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct { uint32_t node; double t0; } leafref_t;
typedef struct { leafref_t* v; size_t n, cap; } leaflist_t;

static void leaf_insert(leaflist_t* L, uint32_t node, double t0)
{
    if (L->n == L->cap) { L->cap = L->cap ? 2 * L->cap : 64; L->v = (leafref_t*)realloc(L->v, L->cap * sizeof(leafref_t)); }
    /* std::partition_point with predicate t0 >= n.second: insert after every entry whose key is <= t0 (octree.cpp:297-300) */
    size_t lo = 0, hi = L->n;
    while (lo < hi) { size_t mid = (lo + hi) / 2; if (t0 >= L->v[mid].t0) lo = mid + 1; else hi = mid; }
    memmove(L->v + lo + 1, L->v + lo, (L->n - lo) * sizeof(leafref_t));
    L->v[lo].node = node; L->v[lo].t0 = t0; L->n++;
}
/* Octree::Node::intersectSorted (octree.cpp:285-313) */
static void enum_sorted(const gi_scene_desc* sc, uint32_t node, const ray_t* r, double tmin, double tmax, leaflist_t* L)
{
    double t0;
    if (!box_hit(sc->node_box + 6 * (size_t)node, r, tmin, tmax, &t0)) return;
    uint8_t mask = sc->node_mask[node];
    if (!mask) { if (sc->node_prim_cnt[node] > 0) leaf_insert(L, node, t0); return; }
    uint32_t c = sc->node_child[node];
    for (int i = 0; i < 8; i++) if (mask & (1u << i)) enum_sorted(sc, c++, r, tmin, tmax, L);
}

typedef struct { int hit; uint32_t prim; v3 p, n; double uv[2]; } closest_t;

static void trace_one(const gi_scene_desc* sc, const ray_t* r, uint64_t seed, uint64_t path, uint64_t depth, leaflist_t* L, closest_t* out)
{
    L->n = 0;
    if (sc->n_nodes) enum_sorted(sc, 0, r, 0, INFINITY, L);
    if (g_xs) (void)xs_next(g_xs);   /* replay: `drand() < .5 && nodes.size() > 500` (raytracer.h:438) always draws */
    v3 hit = V(0, 0, 0), norm = V(0, 0, 0); double uv[2] = { 0, 0 };
    out->hit = 0; out->prim = GI_NO_HIT; out->p = V(0, 0, 0); out->n = V(0, 0, 0); out->uv[0] = out->uv[1] = 0;
    int term = 0;
    for (size_t li = 0; li < L->n && !term; li++) {
        uint32_t node = L->v[li].node;
        const uint32_t* ids = sc->leaf_prims + sc->node_prim_off[node];
        for (uint32_t k = 0; k < sc->node_prim_cnt[node]; k++) {
            uint32_t id = ids[k];
            if (prim_hit(sc, id, r, &hit, &norm, uv) && alpha_pass(sc, id, node, uv, seed, path, depth, SITE_ALPHA_TRACE)) {
                if (!out->hit || len2(sub(hit, r->o)) < len2(sub(out->p, r->o))) {
                    out->prim = id; out->p = hit; out->n = norm; out->uv[0] = uv[0]; out->uv[1] = uv[1]; out->hit = 1;
                    if (box_contains(sc->node_box + 6 * (size_t)node, hit)) term = 1;
                }
            }
        }
    }
}

void go_trace_closest(const gi_scene_desc* sc, size_t n, const double* org, const double* dir, uint64_t alpha_seed, uint32_t* prim, double* hit, double* normal, double* uv)
{
#pragma omp parallel
    {
        leaflist_t L = { 0, 0, 0 };
#pragma omp for schedule(dynamic, 256)
        for (long i = 0; i < (long)n; i++) {
            ray_t r = ray_as_stored(ld3(org + 3 * i), ld3(dir + 3 * i));
            closest_t c; trace_one(sc, &r, alpha_seed, (uint64_t)i, 0, &L, &c);
            if (prim) prim[i] = c.prim;
            if (hit) st3(hit + 3 * i, c.p);
            if (normal) st3(normal + 3 * i, c.n);
            if (uv) { uv[2 * i] = c.uv[0]; uv[2 * i + 1] = c.uv[1]; }
        }
        free(L.v);
    }
}

/* Canonical ordered traversal: pop node; test each existing child box; descend front-to-back (entry t0,
 * ties in child order); at a non-empty leaf test every primitive in stored order; stop after the first leaf in which
 * an accepted hit lies inside the leaf box.  Same acceptance rules as trace_one; counts box and primitive tests. */
typedef struct { uint32_t node; double t0; int beyond; } stk_t;
/* prune != 0: the traversal the device executes (same hits, fewer tests; this variant counts them).
 * The reference keeps walking to the end of the ray whenever the accepted hit was first met from a leaf that does not contain it: the
 * same hit met again is not STRICTLY closer, so `term` never fires (raytracer.h:457-466).  But nothing met in a node the ray enters at
 * t0 >= t_hit can be accepted — |hit - o|^2 is monotone in the ray parameter, and a closer hit point lies in a leaf entered before
 * t_hit, which is visited — with one exception: a candidate cut out at a FRACTIONAL alpha draws again in every other leaf that holds
 * it (raytracer.h:455) and may be accepted there.  Rules (gi_device.cuh trace_walk follows them to the letter):
 *   R1  a child entered at t0 >= t_hit (t_hit of the moment it is tested) is marked `beyond`;
 *   R2  a node popped unmarked while a hit exists is marked when its own t0 >= t_hit (entries pushed before the hit was found);
 *   R3  a marked node is dropped when it is reached, provided no fractional-alpha rejection lies in front of the hit
 *       (t_hit <= t_frac); otherwise it is walked like the reference does.
 * Leaves are reached in ascending t0, so when R3 drops a node every leaf entered before t_hit has been tested already and t_frac
 * is final for the region in front of the hit. */
static void trace_one_cot(const gi_scene_desc* sc, const ray_t* r, uint64_t seed, uint64_t path, uint64_t depth, closest_t* out, uint32_t* nn, uint32_t* np, int prune)
{
    double hit_t = INFINITY, frac_t = INFINITY;   /* ray parameter of the accepted hit; least ray parameter of a fractional-alpha rejection */
    stk_t stack[8 * 64]; int sp = 0;
    uint32_t n_node = 0, n_prim = 0;
    v3 hit = V(0, 0, 0), norm = V(0, 0, 0); double uv[2] = { 0, 0 };
    out->hit = 0; out->prim = GI_NO_HIT; out->p = V(0, 0, 0); out->n = V(0, 0, 0); out->uv[0] = out->uv[1] = 0;
    double t0;
    if (sc->n_nodes) { n_node++; if (box_hit(sc->node_box, r, 0, INFINITY, &t0)) { stack[sp].node = 0; stack[sp].t0 = t0; stack[sp].beyond = 0; sp++; } }
    int term = 0;
    while (sp > 0 && !term) {
        stk_t cur = stack[--sp];
        uint8_t mask = sc->node_mask[cur.node];
        if (prune) {
            int beyond = cur.beyond;
            if (!beyond && out->hit && cur.t0 >= hit_t) beyond = 1;            /* R2 */
            if (beyond && hit_t <= frac_t) continue;                            /* R3 */
        }
        if (!mask) {
            const uint32_t* ids = sc->leaf_prims + sc->node_prim_off[cur.node];
            for (uint32_t k = 0; k < sc->node_prim_cnt[cur.node]; k++) {
                uint32_t id = ids[k]; n_prim++;
                if (!prim_hit(sc, id, r, &hit, &norm, uv)) continue;
                if (!alpha_pass(sc, id, cur.node, uv, seed, path, depth, SITE_ALPHA_TRACE)) {
                    if (g_alpha_frac && g_last_t < frac_t) frac_t = g_last_t;
                    continue;
                }
                if (!out->hit || len2(sub(hit, r->o)) < len2(sub(out->p, r->o))) {
                    out->prim = id; out->p = hit; out->n = norm; out->uv[0] = uv[0]; out->uv[1] = uv[1]; out->hit = 1;
                    hit_t = g_last_t;
                    if (box_contains(sc->node_box + 6 * (size_t)cur.node, hit)) term = 1;
                }
            }
            continue;
        }
        /* children that the ray enters, sorted by entry t0 (stable in child order), pushed far-to-near */
        stk_t ch[8]; int nc = 0; uint32_t c = sc->node_child[cur.node];
        for (int i = 0; i < 8; i++) if (mask & (1u << i)) {
            n_node++;
            if (box_hit(sc->node_box + 6 * (size_t)c, r, 0, INFINITY, &t0)) {
                int j = nc++;
                while (j > 0 && ch[j - 1].t0 > t0) { ch[j] = ch[j - 1]; j--; }
                ch[j].node = c; ch[j].t0 = t0; ch[j].beyond = out->hit && t0 >= hit_t;   /* R1 */
            }
            c++;
        }
        for (int j = nc - 1; j >= 0; j--) stack[sp++] = ch[j];
    }
    if (nn) *nn = n_node;
    if (np) *np = n_prim;
}
static void trace_closest_cot_n(const gi_scene_desc* sc, size_t n, const double* org, const double* dir, uint64_t alpha_seed, uint32_t* prim, double* hit, uint32_t* n_node_tests, uint32_t* n_prim_tests, int prune)
{
#pragma omp parallel for schedule(dynamic, 256)
    for (long i = 0; i < (long)n; i++) {
        ray_t r = ray_as_stored(ld3(org + 3 * i), ld3(dir + 3 * i));
        closest_t c; uint32_t a, b;
        trace_one_cot(sc, &r, alpha_seed, (uint64_t)i, 0, &c, &a, &b, prune);
        if (prim) prim[i] = c.prim;
        if (hit) st3(hit + 3 * i, c.p);
        if (n_node_tests) n_node_tests[i] = a;
        if (n_prim_tests) n_prim_tests[i] = b;
    }
}

void go_trace_closest_cot(const gi_scene_desc* sc, size_t n, const double* org, const double* dir, uint64_t alpha_seed, uint32_t* prim, double* hit, uint32_t* n_node_tests, uint32_t* n_prim_tests)
{
    trace_closest_cot_n(sc, n, org, dir, alpha_seed, prim, hit, n_node_tests, n_prim_tests, 0);
}
void go_trace_closest_cot_pruned(const gi_scene_desc* sc, size_t n, const double* org, const double* dir, uint64_t alpha_seed, uint32_t* prim, double* hit, uint32_t* n_node_tests, uint32_t* n_prim_tests)
{
    trace_closest_cot_n(sc, n, org, dir, alpha_seed, prim, hit, n_node_tests, n_prim_tests, 1);
}

/* ------------------------------------------------------------------------------------------------------------
 * any hit: RayTracer::visible (raytracer.h:280-319) over Octree::intersect / Node::intersect (octree.cpp:150-185,
 * 256-282) with BoundingBox::intersectSimple (bbox.h:117-138)
 * ---------------------------------------------------------------------------------------------------------- */
static int visible_rec(const gi_scene_desc* sc, uint32_t node, const ray_t* r, double tmax, double mt, uint64_t seed, uint64_t path, uint64_t depth, uint64_t light, uint32_t* nn, uint32_t* np)
{
    /* returns 1 when a blocker was found in this subtree.  The reference first collects every candidate and then
     * tests them in the same DFS order, stopping at the first blocker; testing while walking is equivalent. */
    if (nn) (*nn)++;
    if (!box_hit(sc->node_box + 6 * (size_t)node, r, 0, tmax, NULL)) return 0;
    uint8_t mask = sc->node_mask[node];
    if (!mask) {
        const uint32_t* ids = sc->leaf_prims + sc->node_prim_off[node];
        for (uint32_t k = 0; k < sc->node_prim_cnt[node]; k++) {
            uint32_t id = ids[k]; v3 pos, norm; double uv[2] = { 0, 0 };
            if (np) (*np)++;
            if (prim_hit(sc, id, r, &pos, &norm, uv) && alpha_pass(sc, id, node, uv, seed, path, depth, SITE_ALPHA_SHADOW + (light << 8))) {
                double t_shadow = len2(sub(pos, r->o));
                if ((t_shadow < mt) && (t_shadow > 0)) return 1;
            }
        }
        return 0;
    }
    uint32_t c = sc->node_child[node];
    for (int i = 0; i < 8; i++) if (mask & (1u << i)) { if (visible_rec(sc, c++, r, tmax, mt, seed, path, depth, light, nn, np)) return 1; }
    return 0;
}
static int visible_one(const gi_scene_desc* sc, const ray_t* r, double mt, uint64_t seed, uint64_t path, uint64_t depth, uint64_t light, uint32_t* nn, uint32_t* np)
{
    if (!sc->n_nodes) return 1;
    return !visible_rec(sc, 0, r, sqrt(mt) - GO_SHADOW_BIAS, mt, seed, path, depth, light, nn, np);
}
void go_trace_any(const gi_scene_desc* sc, size_t n, const double* org, const double* dir, const double* maxt2, uint64_t alpha_seed, uint8_t* vis)
{
#pragma omp parallel for schedule(dynamic, 256)
    for (long i = 0; i < (long)n; i++) {
        ray_t r = ray_as_stored(ld3(org + 3 * i), ld3(dir + 3 * i));
        vis[i] = (uint8_t)visible_one(sc, &r, maxt2[i], alpha_seed, (uint64_t)i, 0, 0, NULL, NULL);
    }
}
/* Octree::intersect (octree.cpp:150-185, 256-282): the entities of every non-empty leaf met by [tmin, tmax], in the recursion's order */
static void octree_intersect_rec(const gi_scene_desc* sc, uint32_t node, const ray_t* r, double tmin, double tmax, uint32_t cap, uint32_t* ids, uint32_t* cnt)
{
    if (!box_hit(sc->node_box + 6 * (size_t)node, r, tmin, tmax, NULL)) return;   /* intersectSimple: same folds, same early rejects */
    uint8_t mask = sc->node_mask[node];
    if (!mask) {
        const uint32_t* l = sc->leaf_prims + sc->node_prim_off[node];
        for (uint32_t k = 0; k < sc->node_prim_cnt[node]; k++, (*cnt)++) if (*cnt < cap) ids[*cnt] = l[k];
        return;
    }
    uint32_t c = sc->node_child[node];
    for (int i = 0; i < 8; i++) if (mask & (1u << i)) octree_intersect_rec(sc, c++, r, tmin, tmax, cap, ids, cnt);
}
void go_octree_intersect(const gi_scene_desc* sc, size_t n, const double* org, const double* dir, const double* tmin, const double* tmax, uint32_t cap, uint32_t* ids, uint32_t* counts)
{
#pragma omp parallel for schedule(dynamic, 64)
    for (long i = 0; i < (long)n; i++) {
        ray_t r = ray_as_stored(ld3(org + 3 * i), ld3(dir + 3 * i));
        uint32_t c = 0;
        if (sc->n_nodes) octree_intersect_rec(sc, 0, &r, tmin[i], tmax[i], cap, ids + (size_t)i * cap, &c);
        counts[i] = c;
    }
}
/* Octree::intersectSorted (octree.cpp:188-211, 285-313): flattened node index and entry distance of every non-empty leaf, sorted */
void go_octree_intersect_sorted(const gi_scene_desc* sc, size_t n, const double* org, const double* dir, const double* tmin, const double* tmax, uint32_t cap, uint32_t* nodes, double* t0,
                                uint32_t* counts)
{
#pragma omp parallel
    {
        leaflist_t L = { 0, 0, 0 };
#pragma omp for schedule(dynamic, 64)
        for (long i = 0; i < (long)n; i++) {
            ray_t r = ray_as_stored(ld3(org + 3 * i), ld3(dir + 3 * i));
            L.n = 0;
            if (sc->n_nodes) enum_sorted(sc, 0, &r, tmin[i], tmax[i], &L);
            for (size_t k = 0; k < L.n && k < cap; k++) { nodes[(size_t)i * cap + k] = L.v[k].node; t0[(size_t)i * cap + k] = L.v[k].t0; }
            counts[i] = (uint32_t)L.n;
        }
        free(L.v);
    }
}

/* sequential forms on the reference's own PRNG stream (see g_xs): *state = the interposed time() value, updated */
void go_trace_closest_replay(const gi_scene_desc* sc, size_t n, const double* org, const double* dir, uint64_t* state, uint32_t* prim, double* hit, double* normal, double* uv)
{
    leaflist_t L = { 0, 0, 0 };
    g_xs = state;
    for (size_t i = 0; i < n; i++) {
        ray_t r = ray_as_stored(ld3(org + 3 * i), ld3(dir + 3 * i));
        closest_t c; trace_one(sc, &r, 0, (uint64_t)i, 0, &L, &c);
        prim[i] = c.prim;
        if (hit) st3(hit + 3 * i, c.p);
        if (normal) st3(normal + 3 * i, c.n);
        if (uv) { uv[2 * i] = c.uv[0]; uv[2 * i + 1] = c.uv[1]; }
    }
    g_xs = NULL;
    free(L.v);
}
void go_trace_any_replay(const gi_scene_desc* sc, size_t n, const double* org, const double* dir, const double* maxt2, uint64_t* state, uint8_t* vis)
{
    g_xs = state;
    for (size_t i = 0; i < n; i++) {
        ray_t r = ray_as_stored(ld3(org + 3 * i), ld3(dir + 3 * i));
        vis[i] = (uint8_t)visible_one(sc, &r, maxt2[i], 0, (uint64_t)i, 0, 0, NULL, NULL);
    }
    g_xs = NULL;
}
void go_trace_any_cot(const gi_scene_desc* sc, size_t n, const double* org, const double* dir, const double* maxt2, uint64_t alpha_seed, uint8_t* vis, uint32_t* n_node_tests, uint32_t* n_prim_tests)
{
#pragma omp parallel for schedule(dynamic, 256)
    for (long i = 0; i < (long)n; i++) {
        ray_t r = ray_as_stored(ld3(org + 3 * i), ld3(dir + 3 * i));
        uint32_t a = 0, b = 0;
        vis[i] = (uint8_t)visible_one(sc, &r, maxt2[i], alpha_seed, (uint64_t)i, 0, 0, &a, &b);
        if (n_node_tests) n_node_tests[i] = a;
        if (n_prim_tests) n_prim_tests[i] = b;
    }
}

/* ------------------------------------------------------------------------------------------------------------
 * photon map (photonMap.cpp)
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct { double box[6]; uint32_t child; /* index of child 0, 0 = leaf */ uint32_t off, cnt; uint32_t depth; } pnode_t;
struct go_pmap {
    size_t n_photons; double* ph; /* copy of the 9-double photon records */
    pnode_t* nodes; size_t n_nodes, cap_nodes;
    uint32_t* ids; /* photon ids grouped by leaf */ size_t n_ids, cap_ids;
    uint32_t max_depth;
};
static uint32_t pm_new_nodes(go_pmap* m, size_t k)
{
    if (m->n_nodes + k > m->cap_nodes) { m->cap_nodes = (m->n_nodes + k) * 2; m->nodes = (pnode_t*)realloc(m->nodes, m->cap_nodes * sizeof(pnode_t)); }
    uint32_t r = (uint32_t)m->n_nodes; m->n_nodes += k;
    memset(m->nodes + r, 0, k * sizeof(pnode_t));
    return r;
}
static void set_box(double* b, double x0, double y0, double z0, double x1, double y1, double z1) { b[0] = x0; b[1] = y0; b[2] = z0; b[3] = x1; b[4] = y1; b[5] = z1; }
/* the eight child boxes exactly as written in PhotonMap::Node::partition (photonMap.cpp:139-149) and
 * Octree::Node::partition (octree.cpp:318-328): mid = mix(min,max,.5) = min + .5*(max-min); dx = max.x-min.x */
void go_child_boxes(const double* b, double out[8][6])
{
    double mx = b[0] + .5 * (b[3] - b[0]), my = b[1] + .5 * (b[4] - b[1]), mz = b[2] + .5 * (b[5] - b[2]);
    double dx = b[3] - b[0], dy = b[4] - b[1], dz = b[5] - b[2];
    set_box(out[0], b[0], b[1], b[2], mx, my, mz);
    set_box(out[1], b[0] + .5 * dx, b[1], b[2], mx + .5 * dx, my, mz);
    set_box(out[2], b[0], b[1], b[2] + .5 * dz, mx, my, mz + .5 * dz);
    set_box(out[3], b[0] + .5 * dx, b[1], b[2] + .5 * dz, mx + .5 * dx, my, mz + .5 * dz);
    set_box(out[4], b[0], b[1] + .5 * dy, b[2], mx, my + .5 * dy, mz);
    set_box(out[5], b[0] + .5 * dx, b[1] + .5 * dy, b[2], mx + .5 * dx, my + .5 * dy, mz);
    set_box(out[6], b[0], b[1] + .5 * dy, b[2] + .5 * dz, mx, my + .5 * dy, mz + .5 * dz);
    set_box(out[7], mx, my, mz, b[3], b[4], b[5]);
}
/* PhotonMap::Node::partition (photonMap.cpp:137-192); `list` holds the node's photon ids in insertion order */
static void pm_partition(go_pmap* m, uint32_t node, uint32_t* list, uint32_t n)
{
    double cb[8][6]; go_child_boxes(m->nodes[node].box, cb);
    uint32_t depth = m->nodes[node].depth;
    uint32_t c0 = pm_new_nodes(m, 8);
    m->nodes[node].child = c0; m->nodes[node].cnt = 0;
    uint32_t* bucket[8]; uint32_t bn[8];
    for (int i = 0; i < 8; i++) { bucket[i] = (uint32_t*)malloc((n ? n : 1) * sizeof(uint32_t)); bn[i] = 0; memcpy(m->nodes[c0 + i].box, cb[i], sizeof(cb[i])); m->nodes[c0 + i].depth = depth + 1; }
    if (depth + 1 > m->max_depth) m->max_depth = depth + 1;
    for (uint32_t k = 0; k < n; k++) {
        v3 p = ld3(m->ph + 9 * (size_t)list[k]);
        for (int i = 0; i < 8; i++) if (box_contains(cb[i], p)) bucket[i][bn[i]++] = list[k];
    }
    /* the "no improvement" early-out (photonMap.cpp:174-178) needs avg > .75*n, impossible without duplication, but keep it */
    double avg = 0; for (int i = 0; i < 8; i++) avg += (double)bn[i];
    avg /= 8;
    int stop = avg > 0.75 * n;
    for (int i = 0; i < 8; i++) {
        if (!stop && bn[i] > GO_MAX_PHOTONS_PER_LEAF) pm_partition(m, c0 + i, bucket[i], bn[i]);
        else {
            if (m->n_ids + bn[i] > m->cap_ids) { m->cap_ids = (m->n_ids + bn[i]) * 2 + 64; m->ids = (uint32_t*)realloc(m->ids, m->cap_ids * sizeof(uint32_t)); }
            m->nodes[c0 + i].off = (uint32_t)m->n_ids; m->nodes[c0 + i].cnt = bn[i];
            memcpy(m->ids + m->n_ids, bucket[i], bn[i] * sizeof(uint32_t)); m->n_ids += bn[i];
        }
        free(bucket[i]);
    }
}
go_pmap* go_pmap_build(size_t n, const double* photons9, const double box6[6])
{
    go_pmap* m = (go_pmap*)calloc(1, sizeof(go_pmap));
    m->n_photons = n; m->ph = (double*)malloc((n ? n : 1) * 9 * sizeof(double));
    memcpy(m->ph, photons9, n * 9 * sizeof(double));
    uint32_t root = pm_new_nodes(m, 1);
    memcpy(m->nodes[root].box, box6, 6 * sizeof(double));
    uint32_t* all = (uint32_t*)malloc((n ? n : 1) * sizeof(uint32_t));
    for (size_t i = 0; i < n; i++) all[i] = (uint32_t)i;
    if (n > GO_MAX_PHOTONS_PER_LEAF) pm_partition(m, root, all, (uint32_t)n);   /* PhotonMap::rebuild (photonMap.cpp:33-47) */
    else {
        m->ids = (uint32_t*)malloc((n ? n : 1) * sizeof(uint32_t)); m->cap_ids = n ? n : 1;
        memcpy(m->ids, all, n * sizeof(uint32_t)); m->n_ids = n; m->nodes[root].off = 0; m->nodes[root].cnt = (uint32_t)n;
    }
    free(all);
    return m;
}
void go_pmap_free(go_pmap* m) { if (!m) return; free(m->ph); free(m->nodes); free(m->ids); free(m); }
void go_pmap_info(const go_pmap* m, uint32_t* n_nodes, uint32_t* n_leaves, uint32_t* n_kept, uint32_t* max_depth)
{
    uint32_t leaves = 0; for (size_t i = 0; i < m->n_nodes; i++) if (!m->nodes[i].child) leaves++;
    if (n_nodes) *n_nodes = (uint32_t)m->n_nodes;
    if (n_leaves) *n_leaves = leaves;
    if (n_kept) *n_kept = (uint32_t)m->n_ids;
    if (max_depth) *max_depth = m->max_depth;
}
static void pm_dump_rec(const go_pmap* m, uint32_t node, size_t* ni, size_t* pi, double* box, uint8_t* leaf, uint32_t* cnt, uint32_t* ids)
{
    const pnode_t* nd = &m->nodes[node];
    memcpy(box + 6 * *ni, nd->box, 6 * sizeof(double)); leaf[*ni] = nd->child ? 0 : 1; cnt[*ni] = nd->cnt; (*ni)++;
    if (!nd->child) { memcpy(ids + *pi, m->ids + nd->off, nd->cnt * sizeof(uint32_t)); *pi += nd->cnt; return; }
    for (int i = 0; i < 8; i++) pm_dump_rec(m, nd->child + i, ni, pi, box, leaf, cnt, ids);
}
void go_pmap_dump(const go_pmap* m, double* node_box6, uint8_t* node_is_leaf, uint32_t* node_count, uint32_t* photon_ids)
{
    size_t ni = 0, pi = 0; pm_dump_rec(m, 0, &ni, &pi, node_box6, node_is_leaf, node_count, photon_ids);
}

/* PhotonMap::Node::getBounds (photonMap.cpp:115-134): leaf containing pos, grown by EPSILON; -inf box if pos is in no child */
static int pm_bounds(const go_pmap* m, v3 pos, double out[6], uint32_t* depth)
{
    uint32_t node = 0;
    for (;;) {
        const pnode_t* nd = &m->nodes[node];
        if (!nd->child) {
            for (int i = 0; i < 3; i++) { out[i] = nd->box[i] - GO_EPSILON; out[3 + i] = nd->box[3 + i] + GO_EPSILON; }
            if (depth) *depth = nd->depth;
            return 1;
        }
        int i = 0;
        while (i < 8 && !box_contains(m->nodes[nd->child + i].box, pos)) i++;
        if (i == 8) { if (depth) *depth = nd->depth; return 0; }
        node = nd->child + i;
    }
}
/* PhotonMap::Node::get (photonMap.cpp:71-92) */
static void pm_get(const go_pmap* m, uint32_t node, const double* q, uint32_t* out, size_t cap, size_t* n)
{
    const pnode_t* nd = &m->nodes[node];
    if (q[3] - q[0] <= 0) return;
    if (!nd->child) {
        for (uint32_t k = 0; k < nd->cnt; k++) { if (*n < cap) out[*n] = m->ids[nd->off + k]; (*n)++; }
        return;
    }
    for (int i = 0; i < 8; i++) if (box_overlap(m->nodes[nd->child + i].box, q)) pm_get(m, nd->child + i, q, out, cap, n);
}
size_t go_pmap_candidates(const go_pmap* m, const double pos[3], uint32_t* out, size_t cap)
{
    double q[6]; size_t n = 0;
    if (!pm_bounds(m, ld3(pos), q, NULL)) return 0; /* box(-inf,-inf): overlaps nothing (photonMap.cpp:132) */
    pm_get(m, 0, q, out, cap, &n);
    return n;
}

typedef struct { double d2; uint32_t id; } cand_t;
static int cand_cmp(const void* a, const void* b)
{
    const cand_t* x = (const cand_t*)a; const cand_t* y = (const cand_t*)b;
    if (x->d2 < y->d2) return -1;
    if (x->d2 > y->d2) return 1;
    return (x->id > y->id) - (x->id < y->id);
}
/* RayTracer::samplePhotons (raytracer.h:532-579) */
void go_gather(const go_pmap* m, size_t n, const double* pos, const double* dir, int k, double* rgb, uint32_t* knn, uint32_t* n_cand, uint32_t* depth_leaf)
{
#pragma omp parallel
    {
        size_t cap = 4096; uint32_t* ids = (uint32_t*)malloc(cap * sizeof(uint32_t)); cand_t* cd = (cand_t*)malloc(cap * sizeof(cand_t));
#pragma omp for schedule(dynamic, 256)
        for (long i = 0; i < (long)n; i++) {
            v3 p = ld3(pos + 3 * i), d = ld3(dir + 3 * i);
            double q[6]; size_t nc = 0; uint32_t dl = 0;
            if (pm_bounds(m, p, q, &dl)) {
                pm_get(m, 0, q, ids, cap, &nc);
                if (nc > cap) { cap = nc * 2; ids = (uint32_t*)realloc(ids, cap * sizeof(uint32_t)); cd = (cand_t*)realloc(cd, cap * sizeof(cand_t)); nc = 0; pm_get(m, 0, q, ids, cap, &nc); }
            }
            for (size_t c = 0; c < nc; c++) { cd[c].id = ids[c]; cd[c].d2 = len2(sub(ld3(m->ph + 9 * (size_t)ids[c]), p)); }
            /* std::partial_sort by squared distance (raytracer.h:547); exact ties are unordered in the reference, here: id */
            qsort(cd, nc, sizeof(cand_t), cand_cmp);
            int count = (int)nc < k ? (int)nc : k;
            v3 res = V(0, 0, 0);
            for (int c = 0; c < count; c++) {
                const double* ph = m->ph + 9 * (size_t)cd[c].id;
                res = add(res, scale(ld3(ph + 6), dot(ld3(ph + 3), d)));           /* raytracer.h:569 */
            }
            if (nc > 0) { double md = cd[count - 1].d2; double den = GO_PI * md; res = V(res.x / den, res.y / den, res.z / den); } /* :572-576 */
            if (rgb) st3(rgb + 3 * i, res);
            if (knn) for (int c = 0; c < k; c++) knn[(size_t)i * k + c] = c < count ? cd[c].id : GI_NO_HIT;
            if (n_cand) n_cand[i] = (uint32_t)nc;
            if (depth_leaf) depth_leaf[i] = dl;
        }
        free(ids); free(cd);
    }
}

/* ------------------------------------------------------------------------------------------------------------
 * samplers (util.h / util.cpp)
 * ---------------------------------------------------------------------------------------------------------- */
/* fastPrecisePow(double a, double b) (util.h:113-136): exponent bit-hack for the fractional part of b */
double go_fast_precise_pow(double a, double b)
{
    int e = (int)b;
    union { double d; int x[2]; } u; u.d = a;
    u.x[1] = (int)((b - e) * (u.x[1] - 1072632447) + 1072632447);
    u.x[0] = 0;
    double r = 1.0;
    while (e) { if (e & 1) r *= a; a *= a; e >>= 1; }
    return r * u.d;
}
/* tangent frame used by hemisphereSample_cos / sphereCapSample_cos / sample_phong (util.cpp:41-45): glm::dmat3x3 built
 * from nine scalars = three COLUMNS; res = rot*v (type_mat3x3.inl:428-434) */
static inline v3 frame_apply(v3 n, v3 v)
{
    double z = fabs(n.z);
    double c0x = z + (1.0 / (1 + z)) * -n.y * -n.y, c0y = (1.0 / (1 + z)) * (n.x * -n.y), c0z = -n.x;
    double c1x = (1.0 / (1 + z)) * (n.x * -n.y), c1y = z + (1.0 / (1 + z)) * -n.x * -n.x, c1z = -n.y;
    double c2x = n.x, c2y = n.y, c2z = z;
    return V(c0x * v.x + c1x * v.y + c2x * v.z, c0y * v.x + c1y * v.y + c2y * v.z, c0z * v.x + c1z * v.y + c2z * v.z);
}
/* the float-typed body shared by the three samplers: phi, cosTheta, sinTheta are `float`; cos/sin/sqrt are the
 * double overloads applied to float values (util.cpp:31-36,48-51,70-73) */
static inline v3 lobe_dir(float u, float v, float cosTheta)
{
    float phi = (float)(v * 2.0f * GO_PI);
    float sinTheta = (float)sqrt((double)(1.0f - cosTheta * cosTheta));
    return V(cos((double)phi) * sinTheta, sin((double)phi) * sinTheta, cosTheta);
}
void go_hemisphere_cos(const double n[3], float u, float v, double power, double out[3]) /* util.cpp:38-58 */
{
    v3 nn = ld3(n);
    float cosTheta = (float)go_fast_precise_pow(1.0f - u, (1.0f / power));
    v3 res = frame_apply(nn, lobe_dir(u, v, cosTheta));
    if (nn.z < 0) res.z *= -1.0;
    st3(out, res);
}
void go_sphere_cap_cos(const double n[3], float u, float v, double power, double frac, double out[3]) /* util.cpp:60-83 */
{
    v3 nn = ld3(n);
    float cosTheta = (float)(frac * go_fast_precise_pow(1.0f - u, (1.0f / power)) + (1 - frac));
    v3 res = frame_apply(nn, lobe_dir(u, v, cosTheta));
    if (nn.z < 0) res.z *= -1.0;
    st3(out, res);
}
void go_sample_phong(const double outdir[3], const double n[3], double power, double sx, double sy, double out[3]) /* util.cpp:91-107 */
{
    (void)n;
    v3 od = ld3(outdir);
    float u = (float)sx, v = (float)sy;
    float cosTheta = (float)go_fast_precise_pow(1.0f - u, (1.0f / power));   /* hemisphereSample_cos(u,v,power) util.cpp:29-36 */
    v3 res = frame_apply(od, lobe_dir(u, v, cosTheta));
    if (od.z < 0) res.z *= -1.0;
    st3(out, res);
}
void go_random_unit_vec(double x, double y, double out[3]) /* util.h:183-188 */
{
    double theta = acos(2 * y - 1);
    out[0] = sin(theta) * cos(2 * x * GO_PI); out[1] = sin(theta) * sin(2 * x * GO_PI); out[2] = cos(theta);
}
void go_refr(const double inc[3], const double n[3], double eta, double out[3]) /* util.h:173-181 */
{
    v3 I = ld3(inc), N = ld3(n);
    double d = dot(N, I);
    double k = 1.0 - eta * eta * (1.0 - d * d);
    if (k < GO_EPSILON) st3(out, reflect3(I, N));
    else st3(out, sub(scale(I, eta), scale(N, eta * d + sqrt(k))));
}

/* ------------------------------------------------------------------------------------------------------------
 * shading (raytracer.h:167-276, 321-379, 481-506), lights (light.h)
 * ---------------------------------------------------------------------------------------------------------- */
/* RayTracer::rayType (raytracer.h:481-506) */
static int ray_type(const gi_scene_desc* sc, const gi_material* m, const ray_t* r, v3 norm, const double* uv, uint64_t seed, uint64_t path, uint64_t depth)
{
    int type = 2;
    double IOR = m->ior;
    double opacity = tex_alpha(sc, m->diffuse_tex, uv) * m->opacity;
    double r0 = pow((1 - IOR) / (1 + IOR), 2);
    double fs = r0 + (1 - r0) * pow(1 - dot(reflect3(r->d, norm), norm), 5);
    if (m->roughness < .001) type = 0;
    if (go_rand(seed, path, depth, SITE(SITE_TYPE_A, 0)) > opacity) {
        if (go_rand(seed, path, depth, SITE(SITE_TYPE_B, 0)) < fs) type = 0;
        else type = 1;
    }
    return type;
}
/* RayTracer::secondaryRay (raytracer.h:321-379).  norm is flipped in place; returns refDir; updates f, contrib, offset */
static v3 secondary_ray(const gi_scene_desc* sc, const gi_material* m, const ray_t* r, v3* norm, const double* uv, double sx, double sy, v3* f, v3* contrib, double* offset, uint64_t seed, uint64_t path, uint64_t depth)
{
    int backface = 0;
    if (dot(*norm, r->d) > 0) { *norm = scale(*norm, -1.0); backface = 1; }
    v3 color = tex_get(sc, m->diffuse_tex, uv);
    int type = ray_type(sc, m, r, *norm, uv, seed, path, depth);
    v3 refDir;
    double d3[3], i3[3], n3[3];
    if (type == 1) {
        st3(i3, r->d); st3(n3, *norm);
        go_refr(i3, n3, backface ? m->ior : 1.0 / m->ior, d3);
        refDir = ld3(d3);
        *offset *= -1;
        *contrib = V(1, 1, 1);
        *f = scale(color, 1.0);
    } else if (type == 0) {
        refDir = reflect3(r->d, *norm);
        *contrib = V(1, 1, 1);
        *f = scale(color, 1.0);
    } else {
        st3(n3, *norm);
        go_hemisphere_cos(n3, (float)sx, (float)sy, 2, d3);
        refDir = ld3(d3);
        if (m->roughness < .9) {
            st3(i3, reflect3(r->d, *norm));
            go_sample_phong(i3, n3, (1.0 / (m->roughness)) + 1, sx, sy, d3);
            refDir = ld3(d3);
            if (dot(refDir, *norm) < 0) refDir = reflect3(refDir, *norm);
        }
        *f = scale(color, 1.0);
        v3 inf = color;
        *contrib = mul(*contrib, inf);
        /* glm::mix(contrib, inf, 0.5) = contrib + 0.5*(inf - contrib) (func_common.inl:143-150) */
        *contrib = add(*contrib, scale(sub(inf, *contrib), 0.5));
    }
    return refDir;
}
/* Light::getPoint(x,y) (light.h:42-45) */
static v3 light_point(const gi_light* l, double x, double y) { double u[3]; go_random_unit_vec(x, y, u); return add(ld3(l->pos), scale(ld3(u), l->rad)); }
/* Light::getPointInRange (light.h:47-53) */
static v3 light_point_in_range(const gi_light* l, double x, double y)
{
    if (l->angle < 1) { double u[3]; go_sphere_cap_cos(l->dir, (float)x, (float)y, 1, l->angle, u); return add(ld3(l->pos), scale(ld3(u), l->rad)); }
    return light_point(l, x, y);
}

/* ------------------------------------------------------------------------------------------------------------
 * atmosphere: HeightFog::density (atmosphere.h:50-81), Octree::atmosphereDensity / atmosphereBounds
 * (octree.cpp:214-251), RayTracer::raymarch (raytracer.h:509-529)
 * ---------------------------------------------------------------------------------------------------------- */
#define GO_RAYMARCH_STEPSIZE 0.04 /* util.h:29 */
/* fastPow (util.h:100-111): the exponent bit-hack on the high word, low word cleared */
static inline double fast_pow(double a, double b)
{
    union { double d; int32_t x[2]; } u;
    u.d = a;
    u.x[1] = (int32_t)(b * (u.x[1] - 1072632447) + 1072632447);
    u.x[0] = 0;
    return u.d;
}
/* one noise-grid read.  The reference indexes std::vector<double> with a double expression (converted to size_t by
 * truncation); an index past the end is undefined behaviour there and reads 0 here. */
static inline double fog_cell(const gi_scene_desc* sc, const gi_fog* g, double idx)
{
    uint64_t i = (uint64_t)idx;
    return i < g->grid_count ? sc->fog_grid[g->grid_offset + i] : 0.0;
}
static double fog_density(const gi_scene_desc* sc, const gi_fog* g, v3 p)
{
    const int nscale = 1;                                             /* atmosphere.h:47: the ctor resets nscale to 1 */
    const double sx = g->size[0], sy = g->size[1], sz = g->size[2];
    double ymax = g->pos[1] + .5 * sy;
    v3 rel = scale(sub(p, ld3(g->bmin)), (double)nscale);
    int ix = (int)rel.x, iy = (int)rel.y, iz = (int)rel.z;
    double dx = nscale * (rel.x - ix), dy = nscale * (rel.y - iy), dz = nscale * (rel.z - iz);
    (void)sy;
    /* the row stride is s.x for BOTH outer terms (atmosphere.h:61-71), kept as written */
    double c00 = (1 - dx) * fog_cell(sc, g, (ix * nscale * sx + iy) * nscale * sz + iz) + dx * fog_cell(sc, g, ((ix + 1) * nscale * sx + iy) * nscale * sz + iz);
    double c01 = (1 - dx) * fog_cell(sc, g, (ix * nscale * sx + iy) * nscale * sz + iz + 1) + dx * fog_cell(sc, g, ((ix + 1) * nscale * sx + iy) * nscale * sz + iz + 1);
    double c10 = (1 - dx) * fog_cell(sc, g, (ix * nscale * sx + (iy + 1)) * nscale * sz + iz) + dx * fog_cell(sc, g, ((ix + 1) * nscale * sx + (iy + 1)) * nscale * sz + iz);
    double c11 = (1 - dx) * fog_cell(sc, g, (ix * nscale * sx + (iy + 1)) * nscale * sz + iz + 1) + dx * fog_cell(sc, g, ((ix + 1) * nscale * sx + (iy + 1)) * nscale * sz + iz + 1);
    double c0 = c00 * (1 - dy) + c10 * dy;
    double c1 = c01 * (1 - dy) + c11 * dy;
    double noise = fast_pow((1 - dz) * c0 + dz * c1, 7);
    return g->density * noise * fast_pow((ymax - p.y) / g->size[1], 2);
}
/* Octree::atmosphereDensity (octree.cpp:214-226): sum of STEPSIZE*density over the volumes containing pos; col = the last one's */
static double atmosphere_density(const gi_scene_desc* sc, v3 pos, v3* col)
{
    double d = 0;
    for (uint32_t k = 0; k < sc->n_fog; k++) {
        const gi_fog* g = &sc->fogs[k];
        if (box_contains(g->bmin, pos)) { *col = ld3(g->col); d += GO_RAYMARCH_STEPSIZE * fog_density(sc, g, pos); }
    }
    return d;
}
/* BoundingBox::intersect with both outputs (bbox.h:47-73); bmin/bmax are adjacent in gi_fog, i.e. one 6-double box */
static inline int box_hit2(const double* b, const ray_t* r, double tmin, double tmax, double* t0o, double* t1o)
{
    const double* o = &r->o.x; const double* inv = &r->inv.x;
    for (int i = 0; i < 3; i++) {
        double t0 = (b[i] - o[i]) * inv[i];
        double t1 = (b[3 + i] - o[i]) * inv[i];
        if (inv[i] < 0.0) { double tmp = t0; t0 = t1; t1 = tmp; }
        tmin = t0 > tmin ? t0 : tmin;
        tmax = t1 < tmax ? t1 : tmax;
        if (tmax <= tmin) return 0;
    }
    *t0o = tmin; *t1o = tmax;
    return 1;
}
/* Octree::atmosphereBounds (octree.cpp:229-251).  `min` starts at 0 and only shrinks, `max` starts at 0 and only grows:
 * the march therefore always starts at the caller's mint and ends at the farthest exit (clipped to the caller's maxt). */
static int atmosphere_bounds(const gi_scene_desc* sc, const ray_t* r, double* mint, double* maxt)
{
    double mn = 0, mx = 0; int hit = 0;
    for (uint32_t k = 0; k < sc->n_fog; k++) {
        double a = 0, b = 0;
        if (box_hit2(sc->fogs[k].bmin, r, *mint, *maxt, &a, &b)) { mn = a < mn ? a : mn; mx = mx < b ? b : mx; hit = 1; }   /* std::min / std::max */
    }
    *mint = *mint < mn ? mn : *mint;
    *maxt = mx < *maxt ? mx : *maxt;
    return hit;
}
/* RayTracer::raymarch (raytracer.h:509-529); one counter draw per step: site(step) */
static int raymarch(const gi_scene_desc* sc, const ray_t* r, v3* hit, v3* col, double mint, double maxt, uint64_t seed, uint64_t path, uint64_t depth, uint64_t site, uint64_t hi)
{
    double t = mint + GO_SHADOW_BIAS;
    v3 cur = add(r->o, scale(r->d, mint));
    v3 stepv = scale(r->d, GO_RAYMARCH_STEPSIZE);
    for (uint64_t step = 0; t < maxt; step++) {
        if (go_rand(seed, path, depth, SITE(site, (hi << 32) | step)) < atmosphere_density(sc, cur, col)) { *hit = cur; return 1; }
        cur = add(cur, stepv);
        t += GO_RAYMARCH_STEPSIZE;
    }
    return 0;
}
/* the tail of RayTracer::visible (raytracer.h:308-316): tmax is the SQUARED distance, as written */
static int fog_blocks(const gi_scene_desc* sc, const ray_t* r, double mt, uint64_t seed, uint64_t path, uint64_t depth, uint64_t site, uint64_t light)
{
    double tmin = 0, tmax = mt; v3 h, c = V(0, 0, 0);
    if (!sc->n_fog) return 0;
    return atmosphere_bounds(sc, r, &tmin, &tmax) && raymarch(sc, r, &h, &c, tmin, tmax, seed, path, depth, site, light);
}
void go_fog_density(const gi_scene_desc* sc, size_t n, const double* pos, double* dens, double* col)
{
    for (size_t i = 0; i < n; i++) { v3 c = V(0, 0, 0); dens[i] = atmosphere_density(sc, ld3(pos + 3 * i), &c); if (col) st3(col + 3 * i, c); }
}
void go_atmosphere_bounds(const gi_scene_desc* sc, size_t n, const double* org, const double* dir, const double* tmax_in, uint8_t* hit, double* mint, double* maxt)
{
    for (size_t i = 0; i < n; i++) {
        ray_t r = ray_as_stored(ld3(org + 3 * i), ld3(dir + 3 * i));
        double a = 0, b = tmax_in[i];
        hit[i] = (uint8_t)atmosphere_bounds(sc, &r, &a, &b); mint[i] = a; maxt[i] = b;
    }
}
void go_raymarch(const gi_scene_desc* sc, size_t n, const double* org, const double* dir, const double* tmax_in, uint64_t seed, uint8_t* hit, double* pos, double* col)
{
#pragma omp parallel for schedule(dynamic, 64)
    for (long i = 0; i < (long)n; i++) {
        ray_t r = ray_as_stored(ld3(org + 3 * i), ld3(dir + 3 * i));
        double a = 0, b = tmax_in[i]; v3 h = V(0, 0, 0), c = V(0, 0, 0);
        hit[i] = (uint8_t)(atmosphere_bounds(sc, &r, &a, &b) && raymarch(sc, &r, &h, &c, a, b, seed, (uint64_t)i, 0, SITE_FOG_RAD, 0));
        if (!hit[i]) { h = V(0, 0, 0); c = V(0, 0, 0); }
        st3(pos + 3 * i, h); st3(col + 3 * i, c);
    }
}

typedef struct { uint64_t closest, shadow, gathers; } tally_t;

/* samplePhotons for one query (same arithmetic as go_gather) */
static v3 gather_one(const go_pmap* m, v3 p, v3 d, int k, uint32_t** ids, cand_t** cd, size_t* cap)
{
    double q[6]; size_t nc = 0;
    if (!m) return V(0, 0, 0);
    if (pm_bounds(m, p, q, NULL)) {
        pm_get(m, 0, q, *ids, *cap, &nc);
        if (nc > *cap) { *cap = nc * 2; *ids = (uint32_t*)realloc(*ids, *cap * sizeof(uint32_t)); *cd = (cand_t*)realloc(*cd, *cap * sizeof(cand_t)); nc = 0; pm_get(m, 0, q, *ids, *cap, &nc); }
    }
    for (size_t c = 0; c < nc; c++) { (*cd)[c].id = (*ids)[c]; (*cd)[c].d2 = len2(sub(ld3(m->ph + 9 * (size_t)(*ids)[c]), p)); }
    qsort(*cd, nc, sizeof(cand_t), cand_cmp);
    int count = (int)nc < k ? (int)nc : k;
    v3 res = V(0, 0, 0);
    for (int c = 0; c < count; c++) { const double* ph = m->ph + 9 * (size_t)(*cd)[c].id; res = add(res, scale(ld3(ph + 6), dot(ld3(ph + 3), d))); }
    if (nc > 0) { double den = GO_PI * (*cd)[count - 1].d2; res = V(res.x / den, res.y / den, res.z / den); }
    return res;
}

/* RayTracer::radiance (raytracer.h:167-276) unrolled into a loop: L = sum_k T_k*(color_k*i_k + cont_k*(emissive_k + color_k*caustic_k)),
 * T_{k+1} = T_k*f_k.  The reference evaluates the same sum recursively (innermost first), so results agree to
 * rounding, not bit-for-bit; the CUDA path uses this same loop form. */
static v3 radiance_path(const gi_scene_desc* sc, const go_pmap* pm, const gi_render_params* P, ray_t ray, uint32_t sample, uint64_t path,
                        leaflist_t* L, uint32_t** gid, cand_t** gcd, size_t* gcap, tally_t* tl)
{
    v3 Lsum = V(0, 0, 0), T = V(1, 1, 1), contrib = V(1, 1, 1);
    for (int depth = 0;; depth++) {
        if (depth > P->max_depth) break;                                                       /* :169 */
        float sx = go_halton_sample((uint32_t)(2 + 2 * depth), sample);                        /* :172-173 */
        float sy = go_halton_sample((uint32_t)(3 + 2 * depth), sample);
        double offset = GO_SHADOW_BIAS;
        closest_t c; trace_one(sc, &ray, P->seed, path, (uint64_t)depth, L, &c); tl->closest++; /* :190 */
        if (!c.hit) { Lsum = add(Lsum, mul(T, ld3(sc->ambient))); break; }                      /* :275 */
        const gi_material* m = &sc->mats[sc->prim_mat[c.prim]];
        v3 i = V(0, 0, 0);
        v3 color = tex_get(sc, m->diffuse_tex, c.uv);                                           /* :200 */
        double roughness = m->roughness;
        v3 f = V(1, 1, 1);
        v3 norm = c.n;
        v3 refDir = secondary_ray(sc, m, &ray, &norm, c.uv, sx, sy, &f, &contrib, &offset, P->seed, path, (uint64_t)depth); /* :207 */
        if (sc->n_fog) {                                                                        /* :209-228 */
            double tmin = 0, tmax = length3(sub(c.p, ray.o));
            if (atmosphere_bounds(sc, &ray, &tmin, &tmax)) {
                v3 fh, fcol = V(0, 0, 0);
                if (raymarch(sc, &ray, &fh, &fcol, tmin, tmax, P->seed, path, (uint64_t)depth, SITE_FOG_RAD, 0)) {
                    double u3[3]; go_random_unit_vec(sx, sy, u3);
                    c.p = fh; refDir = ld3(u3); f = fcol; color = fcol; contrib = fcol; roughness = 1;   /* normal, uv, offset stay the surface's */
                }
            }
        }
        for (uint32_t li = 0; li < sc->n_lights; li++) {                                        /* :230-256 */
            const gi_light* light = &sc->lights[li];
            v3 sp = add(c.p, scale(norm, GO_SHADOW_BIAS));
            v3 lightDir = sub(light_point(light, go_rand(P->seed, path, (uint64_t)depth, SITE(SITE_LIGHT_U, li)), go_rand(P->seed, path, (uint64_t)depth, SITE(SITE_LIGHT_V, li))), sp);
            double maxt = len2(lightDir);
            double hfrac = 1 / (GO_PI * len2(sub(ld3(light->pos), c.p)));
            ray_t sr = make_ray(sp, lightDir);
            int vis = visible_one(sc, &sr, maxt, P->seed, path, (uint64_t)depth, li, NULL, NULL); tl->shadow++;
            if (vis && fog_blocks(sc, &sr, maxt, P->seed, path, (uint64_t)depth, SITE_FOG_SHADOW, li)) vis = 0;   /* :308-316 */
            if (vis) {
                double d = dot(norm, normalize(sub(ld3(light->pos), c.p)));
                if (d < 0) d = 0;
                double l = pow(d, (1.0 / roughness));
                i = scale(scale(ld3(light->col), l), hfrac);
            }
        }
        v3 caustic = V(0, 0, 0);
        if (depth <= P->caustic_max_depth) { caustic = gather_one(pm, c.p, refDir, P->k_photons, gid, gcd, gcap); tl->gathers++; } /* :258 */
        double q = contrib.x < contrib.y ? contrib.y : contrib.x; q = q < contrib.z ? contrib.z : q;    /* compMax = std::max chain, :263, util.h:47-50 */
        int cont = depth <= P->min_depth || go_rand(P->seed, path, (uint64_t)depth, SITE(SITE_RR, 0)) < q; /* :265 */
        Lsum = add(Lsum, mul(T, mul(color, i)));
        if (!cont) break;                                                                       /* :272 */
        f = scale(f, depth <= P->min_depth ? 1.0 : (1.0 / q));                                  /* :267 */
        v3 em = tex_get(sc, m->emissive_tex, c.uv);
        Lsum = add(Lsum, mul(T, add(em, mul(color, caustic))));                                 /* :269 */
        T = mul(T, f);
        ray = make_ray(add(c.p, scale(norm, offset)), refDir);
    }
    return Lsum;
}

void go_render(const gi_scene_desc* sc, const go_pmap* pm, const gi_render_params* P, int x0, int y0, int x1, int y1, int s0, int s1, double* accum, gi_stats* stats)
{
    go_halton_init();
    go_henum he; go_henum_init(&he, (uint32_t)P->width, (uint32_t)P->height);
    long tw = x1 - x0, th = y1 - y0;
    uint64_t n_closest = 0, n_shadow = 0, n_gather = 0;
#pragma omp parallel reduction(+ : n_closest, n_shadow, n_gather)
    {
        leaflist_t L = { 0, 0, 0 };
        size_t gcap = 4096; uint32_t* gid = (uint32_t*)malloc(gcap * sizeof(uint32_t)); cand_t* gcd = (cand_t*)malloc(gcap * sizeof(cand_t));
        tally_t tl = { 0, 0, 0 };
#pragma omp for schedule(dynamic, 16)
        for (long p = 0; p < tw * th; p++) {
            int y = y0 + (int)(p / tw), x = x0 + (int)(p % tw);
            v3 sum = V(0, 0, 0);
            for (int s = s0; s < s1; s++) {
                double o[3], d[3]; uint32_t idx;
                go_camera_ray(&sc->camera, &he, P->width, P->height, x, y, s, o, d, &idx);
                ray_t ray = ray_as_stored(ld3(o), ld3(d));
                uint64_t path = ((uint64_t)((uint64_t)y * (uint64_t)P->width + (uint64_t)x) << 24) | (uint64_t)s;
                sum = add(sum, radiance_path(sc, pm, P, ray, idx, path, &L, &gid, &gcd, &gcap, &tl));
            }
            st3(accum + 3 * p, sum);
        }
        n_closest += tl.closest; n_shadow += tl.shadow; n_gather += tl.gathers;
        free(L.v); free(gid); free(gcd);
    }
    if (stats) { memset(stats, 0, sizeof(*stats)); stats->closest_rays = n_closest; stats->shadow_rays = n_shadow; stats->gathers = n_gather; }
}

/* The adaptive per-pixel loop of RayTracer::run (raytracer.h:100-148): running mean, smoothed change `var`, `samps` bookkeeping.
 * color[n][3] = final running mean, samples[n] = samples taken. */
void go_render_adaptive(const gi_scene_desc* sc, const go_pmap* pm, const gi_render_params* P, int min_samples, int max_samples, double noise_thresh, int x0, int y0, int x1, int y1,
                        double* color_out, uint32_t* samples_out)
{
    go_halton_init();
    go_henum he; go_henum_init(&he, (uint32_t)P->width, (uint32_t)P->height);
    long tw = x1 - x0, th = y1 - y0;
#pragma omp parallel
    {
        leaflist_t L = { 0, 0, 0 };
        size_t gcap = 4096; uint32_t* gid = (uint32_t*)malloc(gcap * sizeof(uint32_t)); cand_t* gcd = (cand_t*)malloc(gcap * sizeof(cand_t));
        tally_t tl = { 0, 0, 0 };
#pragma omp for schedule(dynamic, 16)
        for (long p = 0; p < tw * th; p++) {
            int y = y0 + (int)(p / tw), x = x0 + (int)(p % tw);
            v3 color = V(0.5, 0.5, 0.5), lastCol = V(0, 0, 0);          /* :102-103 */
            double var = 0;
            int samps = 0, s = 0;
            while (s < max_samples && samps < min_samples) {            /* :108 */
                lastCol = color;
                double o[3], d[3]; uint32_t idx;
                go_camera_ray(&sc->camera, &he, P->width, P->height, x, y, s, o, d, &idx);
                ray_t ray = ray_as_stored(ld3(o), ld3(d));
                uint64_t path = ((uint64_t)((uint64_t)y * (uint64_t)P->width + (uint64_t)x) << 24) | (uint64_t)s;
                v3 rad = radiance_path(sc, pm, P, ray, idx, path, &L, &gid, &gcd, &gcap, &tl);
                if (s == 0) color = rad;
                else color = scale(add(scale(color, 1.0 * s), rad), 1.0 / (s + 1));   /* :131-134 */
                if (s > 0) {
                    v3 dc = sub(color, lastCol);
                    var = (1.0 * 5 * var + sqrt(dot(dc, dc))) * (1.0 / (5 + 1));     /* :138 */
                }
                if (s > 0 && var > noise_thresh) samps -= 2;            /* :143-144 */
                s++; samps++;
            }
            st3(color_out + 3 * p, color);
            if (samples_out) samples_out[p] = (uint32_t)s;
        }
        free(L.v); free(gid); free(gcd);
    }
}

/* gamma + clamp + 8-bit (raytracer.h:150-156, util.h:94-97, image.h:14-16); accum holds the sum of spp samples.
 * The reference keeps a running mean (raytracer.h:131-134) which equals sum/spp up to rounding. */
void go_resolve(size_t n_pixels, const double* accum, int spp, uint8_t* rgb8)
{
    for (size_t i = 0; i < n_pixels * 3; i++) {
        double c = accum[i] * (1.0 / spp);
        c = pow(c, 1.0 / 2.2);
        c = c < 0.0 ? 0.0 : (c > 1.0 ? 1.0 : c);   /* glm::clamp = min(max(x,lo),hi); NaN -> 0 like (int)(255*NaN) is UB: keep 0 */
        if (!(c == c)) c = 0.0;
        rgb8[i] = (uint8_t)(int)(255 * c);
    }
}

/* ------------------------------------------------------------------------------------------------------------
 * photon tracing: RayTracer::tracePhotons (raytracer.h:582-715)
 * ---------------------------------------------------------------------------------------------------------- */
size_t go_trace_photons(const gi_scene_desc* sc, int count, int max_depth, uint64_t seed, double* photons9, uint64_t* tries_out, uint64_t* traces_out)
{
    go_halton_init();
    size_t nl = sc->n_lights;
    uint8_t* stored_flag = (uint8_t*)calloc((size_t)count * (nl ? nl : 1), 1);
    uint64_t tries_total = 0, traces_total = 0;
#pragma omp parallel reduction(+ : tries_total, traces_total)
    {
        leaflist_t L = { 0, 0, 0 };
#pragma omp for schedule(dynamic, 64)
        for (long i = 0; i < count; i++) {
            for (uint32_t li = 0; li < nl; li++) {
                const gi_light* l = &sc->lights[li];
                int tries = 0, stored = 0;
                while (!stored && tries < 500) {                                                     /* :602 */
                    uint64_t path = PHOTON_PATH_BIT | ((uint64_t)li << 48) | (uint64_t)((uint64_t)i * 500u + (uint64_t)tries);
                    float sx = go_halton_sample(0, (uint32_t)((int)i * 500 + tries));                /* :604-605 */
                    float sy = go_halton_sample(1, (uint32_t)((int)i * 500 + tries));
                    v3 pos = light_point_in_range(l, sx, sy);                                        /* :612 */
                    double nrm3[3], dir3[3];
                    st3(nrm3, normalize(sub(pos, ld3(l->pos))));
                    /* fmod(drand() + 5*i, 1), fmod(drand() + 13*i, 1) narrowed to float by the callee's signature (:613) */
                    float du = (float)fmod(go_rand(seed, path, 0, SITE(SITE_PH_DIR_U, 0)) + 5 * (int)i, 1);
                    float dv = (float)fmod(go_rand(seed, path, 0, SITE(SITE_PH_DIR_V, 0)) + 13 * (int)i, 1);
                    go_sphere_cap_cos(nrm3, du, dv, 2, l->angle, dir3);
                    ray_t r = make_ray(pos, ld3(dir3));
                    v3 col = scale(ld3(l->col), (1.0 / count) * .5 * l->angle);                      /* :618 */
                    int depth = 0, term = 0, isCaustic = 0;
                    closest_t c; trace_one(sc, &r, seed, path, 0, &L, &c); traces_total++;
                    if (!c.hit) { tries++; continue; }                                               /* :626-630 */
                    v3 hit = c.p, norm = c.n; double uv[2] = { c.uv[0], c.uv[1] }; uint32_t cur = c.prim;
                    while (depth < max_depth && !term) {                                             /* :633 */
                        double roughness = sc->mats[sc->prim_mat[cur]].roughness;
                        if (roughness < 0.1) {
                            trace_one(sc, &r, seed, path, (uint64_t)(depth + 1), &L, &c); traces_total++;  /* :640 */
                            if (!c.hit) { term = 1; continue; }
                            hit = c.p; norm = c.n; uv[0] = c.uv[0]; uv[1] = c.uv[1]; cur = c.prim;
                            const gi_material* m = &sc->mats[sc->prim_mat[cur]];
                            roughness = m->roughness;
                            v3 f = V(0, 0, 0), contrib = V(0, 0, 0); double offset = GO_SHADOW_BIAS;
                            double su = fmod(go_rand(seed, path, (uint64_t)(depth + 1), SITE(SITE_PH_SEC_U, 0)) + 5 * (int)i, 1);
                            double sv = fmod(go_rand(seed, path, (uint64_t)(depth + 1), SITE(SITE_PH_SEC_V, 0)) + 13 * (int)i, 1);
                            v3 refDir = secondary_ray(sc, m, &r, &norm, uv, su, sv, &f, &contrib, &offset, seed, path, (uint64_t)(depth + 1)); /* :656 */
                            if (sc->n_fog) {                                                         /* :658-675 */
                                double tmin = 0, tmax = length3(sub(hit, r.o));
                                if (atmosphere_bounds(sc, &r, &tmin, &tmax)) {
                                    v3 ah, acol = V(0, 0, 0);
                                    if (raymarch(sc, &r, &ah, &acol, tmin, tmax, seed, path, (uint64_t)(depth + 1), SITE_FOG_PHOTON, 0)) {
                                        double u3[3];
                                        go_random_unit_vec(fmod(go_rand(seed, path, (uint64_t)(depth + 1), SITE(SITE_PH_FOG_U, 0)) + 13 * (int)i, 1),
                                                           fmod(go_rand(seed, path, (uint64_t)(depth + 1), SITE(SITE_PH_FOG_V, 0)) + 7 * (int)i, 1), u3);
                                        hit = ah; refDir = ld3(u3); f = acol; roughness = 1;
                                    }
                                }
                            }
                            col = mul(col, f);                                                       /* :677 */
                            r = make_ray(add(hit, scale(norm, offset)), refDir);                     /* :679-680 */
                            isCaustic = 1;
                        }
                        if (depth > 0 && isCaustic && roughness >= 0.1) {                            /* :685-692 */
                            double* out = photons9 + 9 * ((size_t)i * nl + li);
                            st3(out, hit); st3(out + 3, r.d); st3(out + 6, col);
                            stored_flag[(size_t)i * nl + li] = 1;
                            term = 1; stored = 1;
                        }
                        depth++;
                    }
                    tries++;
                }
                tries_total += (uint64_t)tries;
            }
        }
        free(L.v);
    }
    /* compact in (i, light) order — the canonical photon order (the reference's order depends on thread timing, :702-711) */
    size_t ns = 0;
    for (size_t k = 0; k < (size_t)count * nl; k++) if (stored_flag[k]) { if (ns != k) memmove(photons9 + 9 * ns, photons9 + 9 * k, 9 * sizeof(double)); ns++; }
    free(stored_flag);
    if (tries_out) *tries_out = tries_total;
    if (traces_out) *traces_out = traces_total;
    return ns;
}

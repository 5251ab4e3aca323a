// gi_scene.hpp — host-side C++ mirror of GI_Raytracer's scene API (the drop-in boundary).
//
// Same class names, members and call signatures as the reference headers (entities.h, material.h, light.h,
// camera.h, bbox.h, ray.h, photon.h, octree.h, photonMap.h, raytracer.h, sceneLoader.h, meshLoader.h), written
// from scratch.  What differs by design:
//   * the classes only HOLD and BUILD the scene; every ray / photon / gather computation goes to the GPU through the
//     C ABI in include/gi_api.h (RayTracer::run -> gi_scene_upload / gi_photon_trace / gi_photon_map_build /
//     gi_render_tile / gi_resolve).  There is no CPU intersection code here and no CPU fallback;
//   * glm is replaced by the small gi::dvec3/dmat3 below (same operation order as glm 0.9.8.2);
//     inside the reference tree a maintainer can alias them to glm's types;
//   * Qt is gone: Image is a plain RGB8 buffer, textures are read from raw RGBA sidecars.
// The octree build (Octree::rebuild / Node::partition) stays on the host, as the north star asks, and reproduces the
// reference's tree exactly (child boxes, float-precision triangle/box SAT, padding quirks) so that leaf order, and
// with it every hit id, is identical; Octree::flatten() emits the SoA arrays of gi_scene_desc.
#pragma once
#include <array>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <memory>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "../../../include/gi_api.h"

#define GI_MAX_ENTITIES_PER_LEAF 16   // util.h:14
#define GI_MIN_LEAF_SIZE .0015        // util.h:16
#define GI_MAX_SUBDIV_RATIO 0.75      // util.h:17
#define GI_EPSILON 0.00001            // util.h:18
#define GI_PHOTONS 75000              // util.h:27
#define GI_PHOTON_DEPTH 5             // util.h:28
#define GI_MIN_SAMPLES 8              // util.h:25
#define GI_SAMPLES 32                 // util.h:26
#define GI_NOISE_THRESH 0.0015        // util.h:24
#define GI_FOCAL_DIST 240             // camera.h:4

namespace gi {

struct dvec2 { double x = 0, y = 0; dvec2() {} dvec2(double a, double b) : x(a), y(b) {} };
struct dvec3 {
    double x = 0, y = 0, z = 0;
    dvec3() {}
    dvec3(double a, double b, double c) : x(a), y(b), z(c) {}
    explicit dvec3(double s) : x(s), y(s), z(s) {}
    // any other 3-vector type with x, y, z members converts both ways — glm::dvec3 in a translation unit of the reference tree
    // (main.cpp, the scene loader's callers) is accepted wherever this API takes a gi::dvec3, and a result can be assigned back
    template <class V, class = decltype(double(std::declval<const V&>().x) + double(std::declval<const V&>().y) + double(std::declval<const V&>().z)),
              class = typename std::enable_if<!std::is_same<typename std::decay<V>::type, dvec3>::value>::type>
    dvec3(const V& v) : x(v.x), y(v.y), z(v.z) {}
    template <class V, class = decltype(V(0.0, 0.0, 0.0)), class = decltype(std::declval<V&>().x),
              class = typename std::enable_if<!std::is_same<typename std::decay<V>::type, dvec3>::value && !std::is_arithmetic<V>::value>::type>
    explicit operator V() const { return V(x, y, z); }
    double& operator[](int i) { return (&x)[i]; }
    const double& operator[](int i) const { return (&x)[i]; }
};
inline dvec3 operator+(const dvec3& a, const dvec3& b) { return { a.x + b.x, a.y + b.y, a.z + b.z }; }
inline dvec3 operator-(const dvec3& a, const dvec3& b) { return { a.x - b.x, a.y - b.y, a.z - b.z }; }
inline dvec3 operator*(const dvec3& a, const dvec3& b) { return { a.x * b.x, a.y * b.y, a.z * b.z }; }
inline dvec3 operator*(const dvec3& a, double s) { return { a.x * s, a.y * s, a.z * s }; }
inline dvec3 operator*(double s, const dvec3& a) { return { s * a.x, s * a.y, s * a.z }; }
inline dvec3 operator/(const dvec3& a, double s) { return { a.x / s, a.y / s, a.z / s }; }
inline dvec3 operator+(const dvec3& a, double s) { return { a.x + s, a.y + s, a.z + s }; }
inline dvec3 operator-(const dvec3& a, double s) { return { a.x - s, a.y - s, a.z - s }; }
inline double dot(const dvec3& a, const dvec3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline dvec3 cross(const dvec3& x, const dvec3& y) { return { x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y }; }
inline dvec3 normalize(const dvec3& v) { return v * (1.0 / std::sqrt(dot(v, v))); }
inline double length(const dvec3& v) { return std::sqrt(dot(v, v)); }
inline dvec3 vmax(const dvec3& a, const dvec3& b) { return { a.x < b.x ? b.x : a.x, a.y < b.y ? b.y : a.y, a.z < b.z ? b.z : a.z }; }
inline dvec3 vmin(const dvec3& a, const dvec3& b) { return { b.x < a.x ? b.x : a.x, b.y < a.y ? b.y : a.y, b.z < a.z ? b.z : a.z }; }

// column-major 3x3 like glm::dmat3x3: c[col][row]
struct dmat3 {
    double c[3][3] = { { 1, 0, 0 }, { 0, 1, 0 }, { 0, 0, 1 } };
};
inline dvec3 operator*(const dmat3& m, const dvec3& v)
{
    return { m.c[0][0] * v.x + m.c[1][0] * v.y + m.c[2][0] * v.z, m.c[0][1] * v.x + m.c[1][1] * v.y + m.c[2][1] * v.z,
             m.c[0][2] * v.x + m.c[1][2] * v.y + m.c[2][2] * v.z };
}
inline dvec3 operator*(const dvec3& v, const dmat3& m)
{
    return { m.c[0][0] * v.x + m.c[0][1] * v.y + m.c[0][2] * v.z, m.c[1][0] * v.x + m.c[1][1] * v.y + m.c[1][2] * v.z,
             m.c[2][0] * v.x + m.c[2][1] * v.y + m.c[2][2] * v.z };
}
dmat3 eulerAngleXYZ(double t1, double t2, double t3);   // upper-left 3x3 of glm::eulerAngleXYZ (gtx/euler_angles.inl:135-167)
dmat3 inverse(const dmat3& m);                          // glm::inverse (detail/func_matrix.inl:272-294)
dmat3 inverse_euler4(const dmat3& m);                   // glm::inverse of the 4x4 euler matrix, upper-left 3x3 (detail/func_matrix.inl:297-352)

}  // namespace gi

// ---- the reference's global-namespace scene classes ---------------------------------------------------------
struct Ray {  // ray.h:5-32
    Ray(gi::dvec3 o, gi::dvec3 d) : origin(o), dir(d) { setDir(d); }
    void setDir(const gi::dvec3& d)
    {
        dir = gi::normalize(d);
        invDir = gi::dvec3(1.0 / dir.x, 1.0 / dir.y, 1.0 / dir.z);
    }
    gi::dvec3 origin, dir, invDir;
};

struct BoundingBox {  // bbox.h:11-44 (the ray/box slab tests live on the device)
    BoundingBox(gi::dvec3 mn, gi::dvec3 mx) : min(mn), max(mx) {}
    double dx() const { return max.x - min.x; }
    double dy() const { return max.y - min.y; }
    double dz() const { return max.z - min.z; }
    gi::dvec3 center() const { return min + 0.5 * (max - min); }
    gi::dvec3 size() const { return max - min; }
    bool intersect(const BoundingBox& o) const
    {
        return (min.x <= o.max.x && max.x >= o.min.x) && (min.y <= o.max.y && max.y >= o.min.y) && (min.z <= o.max.z && max.z >= o.min.z);
    }
    bool contains(gi::dvec3 p) const { return p.x >= min.x && p.y >= min.y && p.z >= min.z && p.x < max.x && p.y < max.y && p.z < max.z; }
    gi::dvec3 min, max;
};

struct texture {  // material.h:11-29
    texture(gi::dvec3 col) : color(col) {}
    virtual ~texture() {}
    virtual int kind() const { return GI_TEX_CONST; }
    gi::dvec3 color;
};
struct checkerboard : texture {  // material.h:32-48
    checkerboard(int t, gi::dvec3 col0, gi::dvec3 col1) : texture(gi::dvec3(0, 0, 0)), a(col0), b(col1), tiles(t) {}
    int kind() const override { return GI_TEX_CHECKER; }
    gi::dvec3 a, b;
    int tiles;
};
struct imageTexture : texture {  // material.h:51-81; pixels come from the raw RGBA sidecar "<name>.rgba"
    imageTexture(const char* name, gi::dvec2 t);
    int kind() const override { return GI_TEX_IMAGE; }
    std::string fname;
    gi::dvec2 tile;
    int width = 0, height = 0;
    bool has_alpha = false;
    std::vector<uint8_t> rgba;
};

struct Material {  // material.h:84-100
    Material(texture* dif, texture* em, double r, double o, double i = 1) : diffuse(dif), emissive(em), roughness(r), opacity(o), IOR(i) {}
    texture* diffuse;
    texture* emissive;
    double roughness, opacity, IOR;
};

class Octree;
struct gi_ctx;

struct Entity {  // entities.h:17-49
    Entity();
    Entity(const Material& m) : material(m) {}
    virtual ~Entity() {}
    virtual int kind() const = 0;                       // GI_PRIM_* (replaces the virtual intersect(ray) dispatch)
    virtual bool intersect(BoundingBox) { return false; }  // entities.h:38-41 (cone inherits this: never assigned to a child)
    // entities.h:24: the ray test of this entity.  A batch-of-one gi_prim_intersect on the context that holds the entity's scene
    // (Octree::attach, done by RayTracer::run): false — like a miss — while the scene is not on a device.  uv is left alone where
    // the reference leaves it alone (cones, triangles without vertex normals).
    bool intersect(const Ray& ray, gi::dvec3& hit, gi::dvec3& norm, gi::dvec2& uv) const;
    virtual BoundingBox boundingBox() const = 0;
    gi::dvec3 pos, rot;
    Material material;
    Octree* _owner = nullptr;      // set by Octree::push_back
    uint32_t _id = 0;              // primitive id = insertion index
};

struct sphere : Entity {  // entities.h:51-142
    double rad;
    sphere(gi::dvec3 position, double radius, const Material& m) : Entity(m), rad(radius) { pos = position; }
    int kind() const override { return GI_PRIM_SPHERE; }
    BoundingBox boundingBox() const override { return BoundingBox(pos + rad * gi::dvec3(-1, -1, -1), pos + rad * gi::dvec3(1, 1, 1)); }
    bool intersect(BoundingBox bbox) override;
};

struct cone : Entity {  // entities.h:144-300
    double rad, height;
    gi::dmat3 rot;
    cone(gi::dvec3 position, gi::dvec3 rotation, double radius, double height, const Material& m);
    int kind() const override { return GI_PRIM_CONE; }
    BoundingBox boundingBox() const override;
};

struct vertex {  // entities.h:302-327
    gi::dvec3 pos, norm;
    gi::dvec2 texCoord;
    vertex() {}
    vertex(gi::dvec3 p, gi::dvec3 n, gi::dvec2 uv) : pos(p), norm(gi::normalize(n)), texCoord(uv) {}
    vertex(gi::dvec3 p, gi::dvec3 n) : pos(p), norm(gi::normalize(n)) {}
    vertex(gi::dvec3 p) : pos(p) {}
};

struct triangle : Entity {  // entities.h:329-558
    std::array<vertex, 3> vertices;
    gi::dvec3 norm;
    double inv_area;
    triangle(vertex v1, vertex v2, vertex v3, const Material& m);
    int kind() const override { return GI_PRIM_TRIANGLE; }
    bool intersect(BoundingBox bbox) override;
    BoundingBox boundingBox() const override;
};

// mesh generators: they only push triangles into the octree (entities.h:562-785)
struct sphereMesh { sphereMesh(Octree* o, gi::dvec3 position, double radius, int subdivs, const Material& m); int count = 0; };
struct coneMesh { coneMesh(Octree* o, gi::dvec3 position, gi::dvec3 rotation, double radius, double height, int tris, const Material& m); int count = 0; };
struct quadMesh { quadMesh(Octree* o, gi::dvec3 v1, gi::dvec3 v2, gi::dvec3 v3, gi::dvec3 v4, const Material& m); };
struct boxMesh { boxMesh(Octree* o, gi::dvec3 position, gi::dvec3 size, gi::dvec3 rotation, const Material& m); };

struct Light {  // light.h:10-58 (point sampling happens on the device)
    Light(gi::dvec3 position, gi::dvec3 color, double radius) : pos(position), col(color), rad(radius) {}
    Light(gi::dvec3 position, gi::dvec3 target, gi::dvec3 color, double radius) : pos(position), col(color), rad(radius) { dir = gi::normalize(target - position); }
    gi::dvec3 dir;
    double angle = .125;
    gi::dvec3 pos, col;
    double rad = 0;
};

struct Camera {  // camera.h:7-31
    explicit Camera(gi::dvec3 p) : Camera(p, gi::dvec3(0, 0, 0)) {}
    Camera(gi::dvec3 p, gi::dvec3 lookAt) : pos(p) { setDir(lookAt - p); }
    void setDir(gi::dvec3 d)
    {
        forward = gi::normalize(d);
        up = gi::dvec3(0, 1.0, 0);
        right = gi::normalize(gi::cross(up, forward));
        up = gi::cross(forward, right);
    }
    gi::dvec3 pos, up, forward, right;
    const double sensorDiag = 0.035 * GI_FOCAL_DIST * 2;
    const double focalDist = 0.04 * GI_FOCAL_DIST;
    Camera& operator=(const Camera& o) { pos = o.pos; up = o.up; forward = o.forward; right = o.right; return *this; }
    Camera(const Camera&) = default;
};

// atmosphere.h:11-83.  The density field itself is evaluated on the device (fog_density, gi_device.cuh); the host keeps the
// parameters and the noise grid.  The reference fills the grid from its time-seeded drand(); here the values come from the
// counter generator (stream 0x5EEDF06, entity, cell), so a scene renders the same fog every run.
struct AtmosphereEntity {
    AtmosphereEntity(gi::dvec3 position, gi::dvec3 size, gi::dvec3 color, double scatter) : pos(position), col(color), bbox(position - .5 * size, position + .5 * size), sc(scatter) {}
    virtual ~AtmosphereEntity() {}
    gi::dvec3 pos, col;
    BoundingBox bbox;
    double sc = 0;
};
struct HeightFog : AtmosphereEntity {
    HeightFog(gi::dvec3 position, gi::dvec3 size, gi::dvec3 color, double density, double scatter, int noiseScale);   // atmosphere.h:37-48
    double d;
    int nscale;
    std::vector<double> noiseGrid;
    gi::dvec3 s;
};

struct Photon {  // photon.h:5-15
    Photon(gi::dvec3 o, gi::dvec3 d, gi::dvec3 c) : origin(o), dir(d), col(c) {}
    gi::dvec3 origin, dir, col;
};

// A flattened scene: owns the arrays a gi_scene_desc points into.
struct FlatScene {
    std::vector<double> node_box;
    std::vector<uint32_t> node_child, node_prim_off, node_prim_cnt, leaf_prims;
    std::vector<uint8_t> node_mask;
    std::vector<uint8_t> prim_type;
    std::vector<double> prim_geom, prim_nrm, prim_uv, prim_fnorm;
    std::vector<uint32_t> prim_mat;
    std::vector<gi_material> mats;
    std::vector<gi_texture> tex;
    std::vector<uint8_t> tex_pixels;
    std::vector<gi_light> lights;
    gi_camera camera;
    double ambient[3] = { 0, 0, 0 };
    std::vector<gi_fog> fogs;
    std::vector<double> fog_grid;
    gi_scene_desc desc() const;
};

class Octree {  // octree.h:17-65
  public:
    struct Node {
        explicit Node(const BoundingBox& b) : _bbox(b) {}
        void partition();       // octree.cpp:316-384
        bool is_leaf() const;   // octree.cpp:386-393
        BoundingBox _bbox;
        std::vector<Entity*> _entities;
        std::array<std::unique_ptr<Node>, 8> _children;
    };
    Octree(gi::dvec3 mn = gi::dvec3(0, 0, 0), gi::dvec3 mx = gi::dvec3(0, 0, 0)) : _root(BoundingBox(mn, mx)) {}
    std::vector<Light*> lights;
    std::vector<AtmosphereEntity*> at;   // octree.h:60
    void push_back(AtmosphereEntity* a) { at.push_back(a); }   // octree.cpp:48-51
    void push_back(Entity* object);   // octree.cpp:25-38
    void push_back(Light* light);     // octree.cpp:41-46
    void rebuild();                   // octree.cpp:53-119
    // New: the same rebuild with Node::partition run on the device (gi_octree_build); flatten() then hands out the device-built
    // arrays.  Returns a GI_* code; on failure the tree is left invalid.
    int rebuild(gi_ctx* ctx);
    void entity_boxes(std::vector<double>& out6) const;   // Entity::boundingBox() of every entity, insertion order
    double last_build_ms = 0;         // device time of the last rebuild(ctx)
    // New: SoA image of the rebuilt tree for gi_scene_upload.  Primitive id = insertion order of push_back(Entity*).
    void flatten(const Camera& cam, const gi::dvec3& ambient, FlatScene& out) const;
    const std::vector<Entity*>& entities() const { return _all; }
    // octree.h:54,56 — the reference's two ray queries, answered by the device that holds the flattened tree (batch-of-one
    // gi_octree_intersect / gi_octree_intersect_sorted).  attach() names that context (RayTracer::run does it after the upload);
    // without one the queries return nothing.  intersectSorted's Node pointers are images of the flattened nodes (box and
    // entities of each leaf), valid until the next rebuild.
    std::vector<Entity*> intersect(const Ray& ray, double tmin, double tmax) const;
    std::vector<std::pair<const Node*, double>> intersectSorted(const Ray& ray, double tmin, double tmax) const;
    void attach(gi_ctx* ctx, const FlatScene& flat);
    gi_ctx* attached() const { return _query_ctx; }
    bool valid = false;
    Node _root;
    int nodes = 0, skipped_subdiv = 0;
  private:
    void light_cones(const std::vector<Entity*>& list);   // octree.cpp:60-102
    bool _device_built = false; // _dev_* hold the tree instead of _root's children
    std::vector<double> _dev_box; std::vector<uint32_t> _dev_child, _dev_off, _dev_cnt, _dev_leaf; std::vector<uint8_t> _dev_mask;
    std::vector<Entity*> _all;  // insertion order (the root list itself is cleared by partition, octree.cpp:370-371)
    gi_ctx* _query_ctx = nullptr;
    std::vector<std::unique_ptr<Node>> _query_nodes;   // one image per flattened node (attach)
};

class PhotonMap {  // photonMap.h:13-49 — a handle on the device-resident map
  public:
    PhotonMap(gi::dvec3 mn, gi::dvec3 mx) : min(mn), max(mx) {}
    void reserve(int) {}
    void push_back(Photon* p) { staged.push_back(*p); }   // photons supplied by the caller (uploaded by rebuild)
    void rebuild(gi_ctx* ctx);                             // photonMap.cpp:33-47 -> gi_photon_upload (if staged) + gi_photon_map_build
    // photonMap.h:45 — the candidate photons of a query point, Node::get's order (batch-of-one gi_photon_in_range on the context the
    // map was built in).  The Photon objects are a host copy of the device's photon array, fetched on first use.
    std::vector<Photon*> getInRange(gi::dvec3& pos, double& scale, double dist) const;
    bool valid = false;
    gi::dvec3 min, max;
    std::vector<Photon> staged;
  private:
    gi_ctx* _ctx = nullptr;
    mutable std::vector<Photon> _host;   // device photons, original order
};

struct Image {  // image.h:7-29 without Qt: RGB888 rows, top row first
    Image(int w, int h) : _w(w), _h(h), rgb((size_t)w * h * 3, 0) {}
    int width() const { return _w; }
    int height() const { return _h; }
    gi::dvec3 getPixel(int x, int y) const { const uint8_t* p = &rgb[((size_t)y * _w + x) * 3]; return { p[0] / 255., p[1] / 255., p[2] / 255. }; }
    void clear() { std::fill(rgb.begin(), rgb.end(), 0); }
    bool writePPM(const char* path) const;
    bool writePNG(const char* path) const;   // gi_png.cpp (the reference saves through QImage::save, gui.h:39-45)
    int _w, _h;
    std::vector<uint8_t> rgb;
};

class RayTracer {  // raytracer.h:23-735
  public:
    RayTracer() = delete;
    RayTracer(const Camera& camera) : _camera(camera), _image(std::make_shared<Image>(0, 0)) {}
    // The reference copies the tracer by value into Gui / Viewer (gui.h:19, viewer.h:16); copies share the scene, the photon
    // map and the image.  A copy gets its own (lazily created) device context.
    RayTracer(const RayTracer& o);
    RayTracer& operator=(const RayTracer&) = delete;
    ~RayTracer();
    void setScene(Octree* scene);       // raytracer.h:35-39
    // raytracer.h:41-165: octree rebuild if needed, photon phase once, then the frame — all device work through gi_*.
    // Returns 0 or a GI_ERR_* code (the reference returns void and prints).
    int run(int w, int h);
    bool running() const { return _running; }
    // raytracer.h:723-725.  stop() may be called from another thread while run() is in flight (viewer.h:29-34): bands that
    // have not started are skipped like the reference's rows, and the device call in flight is cancelled (gi_cancel).
    void stop();
    void start();
    std::shared_ptr<Image> getImage() const { return _image; }
    // Progressive display: > 0 renders the frame in bands of this many rows, top to bottom, and publishes each band's pixels
    // to the Image as soon as it is done (the reference writes pixels row by row while a 32 ms timer repaints, viewer.h:17-21).
    // 0 (default, headless) = the whole frame in one device call.  Bands are tiles: pixels are identical either way.
    int progressive_rows = 0;
    int rows_done() const { return _rows_done.load(); }   // rows published so far by the run() in flight / last run()

    int photons = GI_PHOTONS;
    int photon_depth = GI_PHOTON_DEPTH;
    int min_samples = GI_MIN_SAMPLES;
    int max_samples = GI_SAMPLES;
    double noise_thresh = GI_NOISE_THRESH;
    gi::dvec3 ambient = gi::dvec3(0, 0, 0);
    Camera _camera;

    // run-time forms of the reference's compile-time knobs (util.h:22-23) and the PRNG seed
    int max_depth = 64, min_depth = 2;
    uint64_t seed = 1;
    int device = 0;
    // New (the reference is one process on one machine's CPU cores): gpus > 1 splits a fixed-sample frame over the devices device ..
    // device + gpus - 1 — interleaved 16-row blocks per GPU (gi_render_rows_image), the photon map built once on the first and broadcast,
    // the 8-bit rows gathered there (NCCL from the C ABI).  The pixels are those of the one-GPU frame, bit for bit.
    int gpus = 1;
    gi_stats last_frame_stats{}, last_photon_stats{};
    double last_photon_ms = 0, last_frame_ms = 0;
    gi_ctx* context();                  // lazily created gi_ctx on `device`
    int run_multi(int w, int h);        // run() when gpus > 1 (fixed sample counts; adaptive sampling stays on one GPU)

  private:
    std::atomic<bool> _running{ false };
    std::atomic<int> _rows_done{ 0 };
    Octree* _scene = nullptr;
    std::shared_ptr<PhotonMap> _photon_map;   // shared by copies of the tracer (the reference shares a raw pointer it never frees)
    std::shared_ptr<Image> _image;
    std::atomic<gi_ctx*> _ctx{ nullptr };      // read by stop() / start() from other threads
    bool _uploaded = false;                     // this tracer's context holds the scene
    bool _map_in_ctx = false;                   // ... and the photon map (validity is per context: a copy has its own)
    std::vector<gi_ctx*> _peers;                // contexts on the other GPUs of a multi-GPU run (owned; communicator rank = index + 1)
    bool _peers_have_scene = false;
};

// PNG without Qt (gi_png.cpp): 8-bit RGBA rows top first, has_alpha as QImage::hasAlphaChannel reports it
bool gi_png_decode(const char* path, int& width, int& height, bool& has_alpha, std::vector<uint8_t>& rgba);
bool gi_png_encode(const char* path, int width, int height, const uint8_t* rgb);
// JPEG without Qt (gi_jpg.cpp): 8-bit RGBA rows top first (alpha 255: QImage::hasAlphaChannel is false for a JPEG), libjpeg's pixels
bool gi_jpg_decode(const char* path, int& width, int& height, std::vector<uint8_t>& rgba);

void loadScene(Octree* o, RayTracer& r, const char* fname);                                              // sceneLoader.h:5
void loadOBJ(Octree* o, const char* fname, gi::dvec3 pos, gi::dvec3 rot, const Material& material);     // meshLoader.h:4

// A translation unit written the way a user of the REFERENCE writes one (main.cpp:19-46 + direct use of the scene API), compiled
// against the mirror in gi_raytracer_b200/csrc/host.  Vectors are a foreign glm-style type (glm::dvec3 itself when GI_TEST_WITH_GLM is
// defined and glm is on the include path): the mirror's API accepts them through gi::dvec3's converting constructor.
// usage: boundary_main <scene.scn> <out.ppm>   (needs a B200: run() renders on the device; exit code 0 = every check held)
#include <cmath>
#include <cstdio>
#include <iostream>
#ifdef GI_TEST_WITH_GLM
#include <glm/glm.hpp>
typedef glm::dvec3 vec3;
typedef glm::dvec2 vec2;
#else
struct vec3 { double x, y, z; vec3() : x(0), y(0), z(0) {} vec3(double a, double b, double c) : x(a), y(b), z(c) {} };
struct vec2 { double x, y; vec2() : x(0), y(0) {} vec2(double a, double b) : x(a), y(b) {} };
#endif
#include "gi_scene.hpp"

#define CHECK(c) do { if (!(c)) { std::fprintf(stderr, "boundary_main: check failed at line %d: %s\n", __LINE__, #c); return 1; } } while (0)

int main(int argc, char** argv)
{
    if (argc < 3) return 2;
    Camera camera(vec3(10, 5, 0), vec3(0, 0, 0));          // main.cpp:30
    RayTracer raytracer(camera);                            // main.cpp:32
    Octree* scene = new Octree();                           // main.cpp:36
    loadScene(scene, raytracer, argv[1]);                   // main.cpp:38
    // entities through the C++ API, positions given as foreign vectors (sceneLoader.cpp builds them the same way)
    Material grey(new texture(vec3(0.7, 0.7, 0.7)), new texture(vec3(0, 0, 0)), 1, 1, 1);
    sphere* ball = new sphere(vec3(0.3, 1.1, -0.2), 0.45, grey);
    scene->push_back(ball);
    raytracer.setScene(scene);                              // main.cpp:41
    raytracer.photons = 20000;
    raytracer.min_samples = raytracer.max_samples = 2;
    raytracer.max_depth = 6;
    raytracer.start();                                      // viewer.h:48-54
    CHECK(raytracer.run(96, 96) == GI_OK);
    CHECK(raytracer.getImage()->width() == 96 && raytracer.getImage()->writePPM(argv[2]));
    double lum = 0;
    for (int y = 0; y < 96; y++) for (int x = 0; x < 96; x++) lum += raytracer.getImage()->getPixel(x, y).x;
    CHECK(lum > 0);

    // the reference's query members (octree.h:54,56; entities.h:24; photonMap.h:45), answered by the device that holds the scene
    const vec3 eye(10, 5, 0), at(0.3, 1.1, -0.2);
    Ray ray(eye, vec3(at.x - eye.x, at.y - eye.y, at.z - eye.z));
    std::vector<std::pair<const Octree::Node*, double>> leaves = scene->intersectSorted(ray, 0, INFINITY);   // raytracer.h:389
    CHECK(!leaves.empty());
    for (size_t i = 1; i < leaves.size(); i++) CHECK(leaves[i - 1].second <= leaves[i].second);
    bool ball_seen = false;
    double best = INFINITY, ball_d2 = -1; const Entity* first_hit = nullptr;
    for (auto& l : leaves) {
        CHECK(l.first->is_leaf() && !l.first->_entities.empty());
        for (Entity* e : l.first->_entities) {                                    // raytracer.h:446-472, without the early exit
            gi::dvec3 hit, norm; gi::dvec2 uv;
            if (e == ball) ball_seen = true;
            if (e->intersect(ray, hit, norm, uv)) {
                const gi::dvec3 d = hit - ray.origin;
                const double d2 = gi::dot(d, d);
                if (d2 < best) { best = d2; first_hit = e; }
                if (e == ball) ball_d2 = d2;
            }
        }
    }
    CHECK(ball_seen && first_hit != nullptr && ball_d2 > 0);                      // the ray was aimed at the ball's centre: it enters it at |centre - eye| - r
    CHECK(std::fabs(std::sqrt(ball_d2) - (std::sqrt(9.7 * 9.7 + 3.9 * 3.9 + 0.2 * 0.2) - 0.45)) < 1e-9);
    std::vector<Entity*> cands = scene->intersect(ray, 0, std::sqrt(ball_d2) + 1e-3);   // raytracer.h:283
    bool in_cands = false;
    for (Entity* e : cands) in_cands |= e == ball;
    CHECK(in_cands);
    // a ray that leaves the scene upwards meets leaves but hits nothing
    Ray up(vec3(0, 30, 0), vec3(0, 1, 0));
    CHECK(scene->intersectSorted(up, 0, INFINITY).empty());
    std::printf("boundary_main ok: %zu sorted leaves, %zu shadow candidates, nearest hit at %.9f, luminance %.1f\n", leaves.size(), cands.size(), std::sqrt(best), lum);
    return 0;
}

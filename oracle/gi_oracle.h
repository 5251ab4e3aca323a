/* gi_oracle.h — TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C, CPU restatement of the reference's rendering hot path (moepforfreedom/GI_Raytracer), used as the
 * checker for the CUDA path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it; the product (libgi_b200.so, gi_raytracer_b200/) never links or calls it.
 *
 * Parity status: PINNED.  Every PRNG-free function is checked bit-for-bit against the reference itself compiled
 * from /root/reference (oracle/_ref/gi_ref, see oracle/Makefile) — Halton values and pixel indices, camera rays,
 * closest-hit primitive ids / hit points / normals / uvs, shadow visibility, photon-map cells and candidate sets,
 * k-nearest index sets and radiance estimates, sampler known-answer tables — and the committed fixtures under
 * tests/golden/ were produced by that same binary (tests/golden/make_golden.py).  The reference's own repository
 * holds no tests or golden vectors.  PRNG-dependent stages (radiance, photon tracing) use a
 * counter-based generator instead of the reference's time-seeded thread-local xorshift64* (util.h:52-80) and are
 * compared with the reference statistically.
 */
#ifndef GI_ORACLE_H
#define GI_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#include "../include/gi_api.h"

#ifdef __cplusplus
extern "C" {
#endif

/* counter-based PRNG shared (as a specification) with the CUDA path: u01 = mix(seed, path, depth, site|counter) */
double go_rand(uint64_t seed, uint64_t path, uint64_t depth, uint64_t site);

/* Halton_sampler::sample after init_faure (halton_sampler.h:573-603, 626-888, 890-1414, 1417-3286) */
void go_halton_init(void);
float go_halton_sample(uint32_t dim, uint32_t index);
/* the device-side table image: u16 permutation tables for all 256 dims, plus per-dim descriptors */
typedef struct go_halton_dim { uint32_t base, block, nblocks, table_off; float scale; } go_halton_dim;
const uint16_t* go_halton_tables(size_t* n_entries);
const go_halton_dim* go_halton_dims(void);

/* Halton_enum (halton_enum.h:69-155) */
typedef struct go_henum { uint32_t p2, p3, mx, my, inc, w, h; float scale_x, scale_y; } go_henum;
void go_henum_init(go_henum* he, uint32_t width, uint32_t height);
uint32_t go_henum_index(const go_henum* he, uint32_t s, uint32_t x, uint32_t y);

/* camera ray of RayTracer::run (raytracer.h:74-78,112-129); dir = Ray::dir (normalised twice) */
void go_camera_ray(const gi_camera* cam, const go_henum* he, int w, int h, int x, int y, int s, double org[3], double dir[3],
                   uint32_t* index);
void go_camera_rays(const gi_camera* cam, int w, int h, int x0, int y0, int x1, int y1, int s0, int s1, double* org, double* dir,
                    uint32_t* index);

/* closest hit / any hit, batch, OpenMP.  counters (optional, per ray): nodes[i] / prims[i] = box tests / primitive
 * tests of the canonical ordered traversal */
void go_trace_closest(const gi_scene_desc* sc, size_t n, const double* org, const double* dir, uint64_t alpha_seed, uint32_t* prim,
                      double* hit, double* normal, double* uv);
void go_trace_any(const gi_scene_desc* sc, size_t n, const double* org, const double* dir, const double* maxt2, uint64_t alpha_seed,
                  uint8_t* vis);
/* canonical ordered traversal with early termination: same answers as go_trace_closest plus work counters */
void go_trace_closest_cot(const gi_scene_desc* sc, size_t n, const double* org, const double* dir, uint64_t alpha_seed,
                          uint32_t* prim, double* hit, uint32_t* n_node_tests, uint32_t* n_prim_tests);
/* the same hits with the device's pruning (children tested against [0, t_best) once a hit is accepted): counts what the kernels execute */
void go_trace_closest_cot_pruned(const gi_scene_desc* sc, size_t n, const double* org, const double* dir, uint64_t alpha_seed,
                          uint32_t* prim, double* hit, uint32_t* n_node_tests, uint32_t* n_prim_tests);
/* Octree::intersect / Octree::intersectSorted as callable queries (octree.cpp:150-211, 256-313) */
void go_octree_intersect(const gi_scene_desc* sc, size_t n, const double* org, const double* dir, const double* tmin, const double* tmax, uint32_t cap, uint32_t* ids,
                         uint32_t* counts);
void go_octree_intersect_sorted(const gi_scene_desc* sc, size_t n, const double* org, const double* dir, const double* tmin, const double* tmax, uint32_t cap, uint32_t* nodes,
                                double* t0, uint32_t* counts);
/* the same two functions on the reference's own PRNG stream (thread_local xorshift64*, util.h:52-80), sequential:
 * *state = the value gi_ref's interposed time() returned; makes alpha-textured scenes bit-comparable with gi_ref run with
 * OMP_NUM_THREADS=1 (one draw per trace() call + one per geometric hit, in the reference's order) */
void go_trace_closest_replay(const gi_scene_desc* sc, size_t n, const double* org, const double* dir, uint64_t* state, uint32_t* prim,
                             double* hit, double* normal, double* uv);
void go_trace_any_replay(const gi_scene_desc* sc, size_t n, const double* org, const double* dir, const double* maxt2, uint64_t* state, uint8_t* vis);
/* Material::diffuse->get(uv) / emissive->get(uv) / Material::getAlpha(uv) (material.h:18-26, 39-45, 63-81, 90-93) */
void go_material_eval(const gi_scene_desc* sc, size_t n, const uint32_t* prim, const double* uv, double* diffuse, double* emissive, double* alpha);
void go_trace_any_cot(const gi_scene_desc* sc, size_t n, const double* org, const double* dir, const double* maxt2,
                      uint64_t alpha_seed, uint8_t* vis, uint32_t* n_node_tests, uint32_t* n_prim_tests);

/* the eight child boxes of Octree::Node::partition / PhotonMap::Node::partition (octree.cpp:318-328, photonMap.cpp:139-149) */
void go_child_boxes(const double* box6, double out[8][6]);

/* photon map (photonMap.cpp) */
typedef struct go_pmap go_pmap;
go_pmap* go_pmap_build(size_t n, const double* photons9, const double box6[6]);
void go_pmap_free(go_pmap* m);
void go_pmap_info(const go_pmap* m, uint32_t* n_nodes, uint32_t* n_leaves, uint32_t* n_kept, uint32_t* max_depth);
/* DFS pre-order dump (children 0..7), same layout as gi_photon_map_download */
void go_pmap_dump(const go_pmap* m, double* node_box6, uint8_t* node_is_leaf, uint32_t* node_count, uint32_t* photon_ids);
/* getInRange (photonMap.cpp:50-66): candidate ids in the reference's order; returns count (cap = capacity of out) */
size_t go_pmap_candidates(const go_pmap* m, const double pos[3], uint32_t* out, size_t cap);
/* samplePhotons (raytracer.h:532-579), batch; knn [n][k] ascending distance (ties: lower photon id first),
 * n_cand [n]; depth_leaf (optional) [n] = depth of the leaf containing pos (for the bytes model) */
void go_gather(const go_pmap* m, size_t n, const double* pos, const double* dir, int k, double* rgb, uint32_t* knn, uint32_t* n_cand,
               uint32_t* depth_leaf);

/* samplers / shading helpers exposed for known-answer tests against the reference */
void go_hemisphere_cos(const double n[3], float u, float v, double power, double out[3]);          /* util.cpp:38-58 */
void go_sphere_cap_cos(const double n[3], float u, float v, double power, double frac, double out[3]); /* util.cpp:60-83 */
void go_sample_phong(const double outdir[3], const double n[3], double power, double sx, double sy, double out[3]); /* util.cpp:91-107 */
void go_random_unit_vec(double x, double y, double out[3]);                                        /* util.h:183-188 */
void go_refr(const double inc[3], const double n[3], double eta, double out[3]);                   /* util.h:173-181 */
double go_fast_precise_pow(double a, double b);                                                    /* util.h:113-136 */

/* atmosphere (atmosphere.h:50-81, octree.cpp:214-251, raytracer.h:509-529), batch forms for known-answer tests:
 * Octree::atmosphereDensity at points (includes the STEPSIZE factor; col = colour of the last containing volume),
 * Octree::atmosphereBounds on rays with mint = 0 and maxt = tmax_in, and atmosphereBounds + raymarch with the counter
 * PRNG (path = ray index, depth 0, one draw per step) */
void go_fog_density(const gi_scene_desc* sc, size_t n, const double* pos, double* dens, double* col);
void go_atmosphere_bounds(const gi_scene_desc* sc, size_t n, const double* org, const double* dir, const double* tmax_in, uint8_t* hit, double* mint, double* maxt);
void go_raymarch(const gi_scene_desc* sc, size_t n, const double* org, const double* dir, const double* tmax_in, uint64_t seed, uint8_t* hit, double* pos, double* col);

/* photon tracing (raytracer.h:582-715) with the counter PRNG; photons9 capacity = count * n_lights; returns stored */
size_t go_trace_photons(const gi_scene_desc* sc, int count, int max_depth, uint64_t seed, double* photons9, uint64_t* tries,
                        uint64_t* traces);

/* frame (raytracer.h:93-160,167-276): SUM over s in [s0,s1) of radiance() per pixel of the rectangle */
void go_render_adaptive(const gi_scene_desc* sc, const go_pmap* pm, const gi_render_params* p, int min_samples, int max_samples, double noise_thresh, int x0, int y0, int x1,
                        int y1, double* color, uint32_t* samples);
void go_render(const gi_scene_desc* sc, const go_pmap* pm, const gi_render_params* p, int x0, int y0, int x1, int y1, int s0, int s1,
               double* accum, gi_stats* stats);
void go_resolve(size_t n_pixels, const double* accum, int spp, uint8_t* rgb8);

#ifdef __cplusplus
}
#endif
#endif

/* gi_api.h — the C ABI of libgi_b200.so: the B200-native rendering hot path behind GI_Raytracer's
 * C++ scene API.
 *
 * The reference (moepforfreedom/GI_Raytracer) has no FFI layer; its hot path is reached through C++ member
 * functions.  Each entry point below replaces one of them (reference file:line given per function) and is what
 * a maintainer binds from the preserved C++ classes (see INTEGRATION.md for the stubs).
 *
 * Conventions
 *   - plain C: pointers and sizes only; every function returns 0 (GI_OK) or a negative GI_ERR_* code and
 *     gi_last_error(ctx) returns the text of the last failure;
 *   - the caller owns every host buffer, the library owns every device buffer;
 *   - "host" entry points take host pointers and copy (pinned staging) to/from the device; the "_dev" twins take
 *     device pointers (e.g. torch tensors' data_ptr) and run on the context's stream without copies;
 *   - one gi_ctx per device; a ctx is not thread-safe, different ctxs are independent;
 *   - there is NO CPU fallback: without a usable CUDA device gi_create fails with GI_ERR_NO_DEVICE.
 *
 * Geometry, rays, hits and photons are fp64, exactly like the reference (glm::dvec3); Halton samples are fp32,
 * like Halton_sampler::sample.  Device code on the parity path is compiled with -fmad=false and follows the
 * reference's operation order, so primitive ids, hit points and photon index sets are bit-exact.
 */
#ifndef GI_API_H
#define GI_API_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GI_OK 0
#define GI_ERR_INVALID (-1)     /* bad argument / bad scene description            */
#define GI_ERR_NO_DEVICE (-2)   /* no CUDA device, or the device is not sm_100     */
#define GI_ERR_CUDA (-3)        /* a CUDA runtime call or kernel failed            */
#define GI_ERR_NO_SCENE (-4)    /* call needs gi_scene_upload first                */
#define GI_ERR_NO_PHOTONS (-5)  /* call needs a built photon map                   */
#define GI_ERR_OOM (-6)
#define GI_ERR_CANCELLED (-7)   /* gi_cancel(ctx, 1) was raised: the call stopped at the next launch boundary */

#define GI_NO_HIT 0xFFFFFFFFu

/* primitive kinds (entities.h:51 sphere, :144 cone, :329 triangle; box/quad/sphere/cone *meshes* are triangles) */
#define GI_PRIM_TRIANGLE 0
#define GI_PRIM_SPHERE 1
#define GI_PRIM_CONE 2

/* texture kinds (material.h:11 texture, :32 checkerboard, :51 imageTexture) */
#define GI_TEX_CONST 0
#define GI_TEX_CHECKER 1
#define GI_TEX_IMAGE 2

typedef struct gi_ctx gi_ctx;

typedef struct gi_texture {
    int32_t kind;          /* GI_TEX_*                                                              */
    int32_t tiles;         /* checkerboard tile count (material.h:35)                               */
    double a[3];           /* const colour, or checker colour a                                     */
    double b[3];           /* checker colour b                                                      */
    double tile_u, tile_v; /* imageTexture::tile (material.h:55)                                    */
    int32_t width, height; /* image size                                                            */
    int32_t has_alpha;     /* QImage::hasAlphaChannel()                                             */
    int32_t _pad;
    uint64_t pixel_offset; /* byte offset of this image's RGBA8 rows (top row first) in tex_pixels  */
} gi_texture;

typedef struct gi_material { /* material.h:84-100 */
    uint32_t diffuse_tex, emissive_tex;
    double roughness, opacity, ior;
} gi_material;

typedef struct gi_light { /* light.h:10-58; dir/angle are the caustic cone written by Octree::rebuild (octree.cpp:79-102) */
    double pos[3], col[3], rad, dir[3], angle;
} gi_light;

/* HeightFog (atmosphere.h:30-83): an axis-aligned volume whose density is a noise grid (trilinear, ^7) times a height falloff.
 * grid values live in gi_scene_desc.fog_grid[grid_offset .. grid_offset + grid_count). */
typedef struct gi_fog {
    double pos[3], size[3], col[3];
    double density, scatter;        /* `d` and `sc` (scatter is stored by the reference, never read)              */
    double bmin[3], bmax[3];        /* pos -+ size/2 (atmosphere.h:13)                                            */
    uint64_t grid_offset, grid_count;
} gi_fog;

typedef struct gi_camera { /* camera.h:7-31 */
    double pos[3], forward[3], up[3], right[3], sensor_diag, focal_dist;
} gi_camera;

/* The scene octree flattened on the host (build stays on the host, only its output is uploaded).
 * Nodes are stored breadth-first; the existing children of a node are contiguous, in the reference's child
 * order 0..7 (octree.cpp:321-328), starting at node_child[n].  A node with child mask 0 is a leaf
 * (Octree::Node::is_leaf, octree.cpp:386-393) and owns leaf_prims[node_prim_off[n] .. +node_prim_cnt[n]) in the
 * reference's per-leaf entity order.  Primitive id = insertion index into Octree::_root._entities. */
typedef struct gi_scene_desc {
    uint32_t n_nodes;
    const double* node_box;         /* [n_nodes][6]  min.xyz, max.xyz                                  */
    const uint32_t* node_child;     /* [n_nodes]     index of the first existing child                 */
    const uint8_t* node_mask;       /* [n_nodes]     bit i set = child i exists                        */
    const uint32_t* node_prim_off;  /* [n_nodes]                                                        */
    const uint32_t* node_prim_cnt;  /* [n_nodes]                                                        */
    uint32_t n_refs;
    const uint32_t* leaf_prims;     /* [n_refs]                                                         */

    uint32_t n_prims;
    const uint8_t* prim_type;       /* [n_prims] GI_PRIM_*                                              */
    const double* prim_geom;        /* [n_prims][9]  tri: v0,v1,v2 | sphere: c.xyz,r | cone: pos.xyz,rad,height */
    const double* prim_nrm;         /* [n_prims][9]  tri: vertex normals | cone: inverse rotation, column-major   */
    const double* prim_uv;          /* [n_prims][6]  tri: vertex uvs                                    */
    const double* prim_fnorm;       /* [n_prims][3]  tri: face normal (entities.h:339)                  */
    const uint32_t* prim_mat;       /* [n_prims]                                                        */

    uint32_t n_mats;
    const gi_material* mats;
    uint32_t n_tex;
    const gi_texture* tex;
    uint64_t tex_pixel_bytes;
    const uint8_t* tex_pixels;

    uint32_t n_lights;
    const gi_light* lights;
    gi_camera camera;
    double ambient[3];              /* RayTracer::ambient (raytracer.h:726)                             */

    uint32_t n_fog;                 /* Octree::at (octree.h:60): atmosphere entities, in push_back order */
    const gi_fog* fogs;
    uint64_t fog_grid_count;
    const double* fog_grid;
} gi_scene_desc;

/* Run-time replacements for the reference's compile-time knobs (util.h:14-31) and RayTracer members. */
typedef struct gi_render_params {
    int32_t width, height;          /* frame size given to RayTracer::run(w,h)                          */
    int32_t max_depth;              /* MAX_DEPTH 64  (util.h:23)                                        */
    int32_t min_depth;              /* MIN_DEPTH 2   (util.h:22)                                        */
    int32_t spp;                    /* fixed samples per pixel: `samples N N t` (raytracer.h:108)       */
    int32_t k_photons;              /* 32 (raytracer.h:258)                                             */
    int32_t caustic_max_depth;      /* 10 (raytracer.h:258)                                             */
    int32_t _pad;
    uint64_t seed;                  /* counter-based PRNG seed (replaces the time-seeded xorshift)      */
} gi_render_params;

/* per-phase counters filled by gi_render_tile / gi_photon_trace (device-side tallies) */
typedef struct gi_stats {
    uint64_t closest_rays;          /* trace() calls        (raytracer.h:389)                           */
    uint64_t shadow_rays;           /* visible() calls      (raytracer.h:283)                           */
    uint64_t gathers;               /* samplePhotons() calls(raytracer.h:538)                           */
    uint64_t photon_tries;          /* emission tries       (raytracer.h:602)                           */
    uint64_t photons_stored;
    uint64_t kernel_launches;       /* kernels launched by the call                                     */
    /* work of the canonical ordered traversal, tallied on the device: box tests / primitive tests    */
    uint64_t closest_node_tests, closest_prim_tests;
    uint64_t shadow_node_tests, shadow_prim_tests;
    /* gather: sum of containing-leaf depths, of candidate counts C, and of min(k, C)                              */
    uint64_t gather_leaf_depth, gather_candidates, gather_selected;
    double trace_ms, shadow_ms, gather_ms, shade_ms, total_ms; /* CUDA-event times: bounce, direct, gather, tail (in shade_ms), whole call */
    /* the part of the totals above that was done inside the tail kernel (one warp per path, time in shade_ms); the
       bounce / direct / gather kernels did the rest in trace_ms / shadow_ms / gather_ms                              */
    uint64_t tail_closest_rays, tail_shadow_rays, tail_gathers;
    uint64_t tail_closest_node_tests, tail_closest_prim_tests, tail_shadow_node_tests, tail_shadow_prim_tests;
    uint64_t tail_gather_leaf_depth, tail_gather_candidates, tail_gather_selected;
    double bin_ms;                  /* ray binning between bounces (k_bin_*)                             */
} gi_stats;

/* ---- lifetime ------------------------------------------------------------------------------------------ */
int gi_create(int device, gi_ctx** out);
void gi_destroy(gi_ctx* ctx);
const char* gi_last_error(const gi_ctx* ctx);
const char* gi_version(void);
/* the CUDA stream the work of this ctx is enqueued on (a cudaStream_t), for callers that time with events or enqueue their own
 * work behind a *_dev call.  A frame also uses side streams the context owns (shadow rays, gather runs, the gather's long lists:
 * "sched_mode" below); every call joins them back into this stream before it returns, so work enqueued here afterwards sees
 * complete results. */
void* gi_stream(gi_ctx* ctx);
int gi_synchronize(gi_ctx* ctx);

/* ---- scene: replaces Octree::push_back/rebuild as the producer of traversal data (octree.cpp:25-119,316-384):
 *      the host builds the reference octree and hands it over flattened ----------------------------------- */
int gi_scene_upload(gi_ctx* ctx, const gi_scene_desc* scene);

/* traversal variant chosen for the uploaded scene: out = {full, implicit_boxes, n_nodes, n_leaf_refs}.  full = 1 when a material
 * can fail the alpha test or a primitive does not write uv (the general kernels are used); implicit_boxes = 1 when every child
 * box equals the partition formula of its parent, so the traversal derives child boxes instead of loading them. */
int gi_scene_info(gi_ctx* ctx, uint32_t out[4]);

/* ---- Halton (halton_sampler.h:626-888 sample, halton_enum.h:106-114 get_index) -------------------------- */
int gi_halton_sample(gi_ctx* ctx, size_t n, const uint32_t* dim, const uint32_t* index, float* out);
int gi_halton_index(gi_ctx* ctx, int width, int height, size_t n, const uint32_t* s, const uint32_t* x, const uint32_t* y,
                    uint32_t* out_index);

/* ---- camera rays: RayTracer::run lines 74-78 + 112-129.  One ray per (pixel, sample): pixels of the rectangle
 *      [x0,x1) x [y0,y1), row-major, for each sample s in [s0,s1) (sample-major).  org/dir are [n][3] fp64 (dir as
 *      stored in Ray::dir), index is the Halton index of the sample. ---------------------------------------- */
int gi_camera_rays(gi_ctx* ctx, int width, int height, int x0, int y0, int x1, int y1, int s0, int s1,
                   double* org, double* dir, uint32_t* index);

/* ---- closest hit: RayTracer::trace (raytracer.h:382-478) over Octree::intersectSorted (octree.cpp:188-211,
 *      285-313).  org/dir are Ray::origin / Ray::dir as the reference's Ray holds them (ray.h:7-17: dir already
 *      normalised by the constructor; invDir = 1/dir is derived on the device).
 *      Outputs (any may be NULL): prim id or GI_NO_HIT, hit point, un-normalised shading normal, uv.
 *      `alpha_seed`: stream seed for the stochastic alpha cut-out (raytracer.h:455); irrelevant for opaque scenes. */
int gi_trace_closest(gi_ctx* ctx, size_t n, const double* org, const double* dir, uint64_t alpha_seed,
                     uint32_t* prim, double* hit, double* normal, double* uv);
int gi_trace_closest_dev(gi_ctx* ctx, size_t n, const double* org, const double* dir, uint64_t alpha_seed,
                         uint32_t* prim, double* hit, double* normal, double* uv);

/* ---- any hit: RayTracer::visible (raytracer.h:280-319) over Octree::intersect (octree.cpp:150-185,256-282).
 *      maxt2 = squared distance bound `mt`; vis[i] = 1 if nothing blocks the segment. ----------------------- */
int gi_trace_any(gi_ctx* ctx, size_t n, const double* org, const double* dir, const double* maxt2, uint64_t alpha_seed,
                 uint8_t* vis);
int gi_trace_any_dev(gi_ctx* ctx, size_t n, const double* org, const double* dir, const double* maxt2, uint64_t alpha_seed,
                     uint8_t* vis);

/* ---- the scene-API queries the reference exposes beside its renderer, as batches (host pointers); the preserved C++ members
 *      Octree::intersect / intersectSorted, PhotonMap::getInRange and Entity::intersect (csrc/host) are batch-of-one calls of these.
 *      gi_octree_intersect: Octree::intersect (octree.cpp:150-185, 256-282): the entities of every non-empty leaf the segment
 *        [tmin, tmax] meets, in the recursion's order (a primitive once per leaf it sits in): prim_ids [n][cap], counts [n] = the
 *        full count (may exceed cap);
 *      gi_octree_intersect_sorted: Octree::intersectSorted (octree.cpp:188-211, 285-313): the non-empty leaves (flattened node
 *        index of gi_scene_desc) with their entry distance, ascending, equal keys in discovery order: node_ids / t0 [n][cap];
 *      gi_photon_in_range: PhotonMap::getInRange (photonMap.cpp:50-92, 115-134): the candidate photons of a query point (original
 *        photon indices, Node::get's order): photon_ids [n][cap], counts = the full count;
 *      gi_prim_intersect: Entity::intersect(ray, hit, norm, uv) of primitive prim[i] (entities.h:60-101, 158-258, 443-490); uv is
 *        written only where the reference writes it (wrote_uv). --------------------------------------------------------------------- */
int gi_octree_intersect(gi_ctx* ctx, size_t n, const double* org, const double* dir, const double* tmin, const double* tmax, uint32_t cap, uint32_t* prim_ids,
                        uint32_t* counts);
int gi_octree_intersect_sorted(gi_ctx* ctx, size_t n, const double* org, const double* dir, const double* tmin, const double* tmax, uint32_t cap, uint32_t* node_ids,
                               double* t0, uint32_t* counts);
int gi_photon_in_range(gi_ctx* ctx, size_t n, const double* pos, uint32_t cap, uint32_t* photon_ids, uint32_t* counts);
int gi_prim_intersect(gi_ctx* ctx, size_t n, const uint32_t* prim, const double* org, const double* dir, uint8_t* ok, double* hit, double* normal, double* uv,
                      uint8_t* wrote_uv);

/* ---- materials: texture::get / checkerboard::get / imageTexture::get + getAlpha and Material::getAlpha (material.h:18-26,
 *      39-45, 63-81, 90-93) for the material of primitive prim[i] at uv[i] — the values RayTracer::radiance reads at
 *      raytracer.h:200 / :269 and the alpha test at :455 / :297.  diffuse / emissive [n][3], alpha [n]; host pointers. ------ */
int gi_material_eval(gi_ctx* ctx, size_t n, const uint32_t* prim, const double* uv, double* diffuse, double* emissive, double* alpha);

/* ---- cancellation: RayTracer::stop / start and the `_running` poll of the row loop (raytracer.h:98, 723-725; viewer.h:29-34).
 *      gi_cancel(ctx, 1) may be called from ANY thread while another thread is inside gi_render_* / gi_photon_trace on the
 *      same context; the call in flight returns GI_ERR_CANCELLED at its next launch boundary (between bounce depths, path
 *      chunks, adaptive passes, photon rounds), its outputs are then incomplete.  The flag stays raised — every later
 *      render / photon call returns GI_ERR_CANCELLED at once — until gi_cancel(ctx, 0). ------------------------------ */
int gi_cancel(gi_ctx* ctx, int raise);

/* ---- scheduling knobs (new; no counterpart in the reference).  Results never depend on them (tested bit for bit).
 *      "overlap_threshold": 0 = every kernel of a frame on ONE stream (what a profiler or a per-kernel timing wants); otherwise
 *                           (default 2^20) the frame runs on the context's three streams as "sched_mode" says, and under
 *                           sched_mode 0 this is the hit count below which a depth counts as short;
 *      "sched_mode":        0 = a depth's shadow rays beside its gather run, both behind the next depth's bounce kernel only when the
 *                           depth is short, everything drained before the tail; 1 = "deferred": bounce kernels / binning / tail kernel
 *                           on the main stream at the highest stream priority, every depth's shadow rays on side stream 0, every
 *                           gather run on side stream 1, hit lists in a ring of "ring" (default 8, 2..8); 2 (default) = 1 unless
 *                           the last large frame of the scene had neither a tail nor a short depth, then 0;
 *      "tail_threshold" (32768), "bin_threshold" (65536), "bounce_mode" (0 auto / 1 thread per ray / 2 persistent),
 *      "trace_mode" (0 / 1 = warp per ray in the batch kernels), "tail_mode" (0 queued / 1 inline gathers). ------------- */
int gi_configure(gi_ctx* ctx, const char* key, long long value);

/* ---- atmosphere.  The fog itself acts inside gi_render_* / gi_photon_trace (RayTracer::radiance
 *      raytracer.h:209-228, ::visible :308-316, ::tracePhotons :658-675) whenever the uploaded scene has n_fog > 0; the
 *      batch forms below expose its pieces for comparison with the reference.
 *      gi_fog_density: Octree::atmosphereDensity (octree.cpp:214-226) = sum of RAYMARCH_STEPSIZE * HeightFog::density
 *        (atmosphere.h:50-81) over the volumes containing pos; col (optional) = colour of the last containing volume.
 *      gi_raymarch: Octree::atmosphereBounds (octree.cpp:229-251) with mint = 0, maxt = tmax[i] -> hit/t0/t1; with
 *        march != 0 followed by RayTracer::raymarch (raytracer.h:509-529, counter PRNG: path = ray index, depth 0, one
 *        draw per step) -> hit = scattered, pos = scatter point, col = volume colour (zeros when not scattered).
 *      gi_trace_any stays the geometric part of RayTracer::visible. ------------------------------------------- */
int gi_fog_density(gi_ctx* ctx, size_t n, const double* pos, double* dens, double* col);
int gi_raymarch(gi_ctx* ctx, size_t n, const double* org, const double* dir, const double* tmax, uint64_t seed, int march,
                uint8_t* hit, double* t0, double* t1, double* pos, double* col);

/* ---- scene octree build on the device: Octree::rebuild / Octree::Node::partition (octree.cpp:53-119,
 *      316-384) with triBoxOverlap's float temporaries (util.cpp:257-330), level-synchronous.  Inputs (host pointers): the
 *      primitives in insertion order (prim_type / prim_geom as in gi_scene_desc), prim_bbox [n][6] = Entity::boundingBox()
 *      (entities.h:103-106, 260-299, 530-557) and the root box Octree::push_back accumulated (octree.cpp:25-38).  The
 *      result stays on the device; gi_octree_download copies it out in gi_scene_desc's node layout (any pointer may be
 *      NULL) — the same arrays the host build + flatten produce, bit for bit. ------------------------------------- */
int gi_octree_build(gi_ctx* ctx, uint32_t n_prims, const uint8_t* prim_type, const double* prim_geom, const double* prim_bbox,
                    const double* root_box6, uint32_t* n_nodes, uint32_t* n_refs, double* build_ms);
int gi_octree_download(gi_ctx* ctx, double* node_box, uint32_t* node_child, uint8_t* node_mask, uint32_t* node_prim_off,
                       uint32_t* node_prim_cnt, uint32_t* leaf_prims);

/* ---- photons: RayTracer::tracePhotons (raytracer.h:582-715); photons are {origin, dir, col} = 9 fp64 (photon.h) */
int gi_photon_trace(gi_ctx* ctx, int count, int max_depth, uint64_t seed, uint64_t* n_stored, gi_stats* stats);
int gi_photon_upload(gi_ctx* ctx, size_t n, const double* photons9);
int gi_photon_count(gi_ctx* ctx, size_t* n);
int gi_photon_download(gi_ctx* ctx, size_t n, double* photons9);

/* ---- photon map: PhotonMap::rebuild / Node::partition (photonMap.cpp:33-47,137-192), built on the device over
 *      the photons currently held by the ctx; root box = scene root box (raytracer.h:38) unless box6 != NULL. - */
int gi_photon_map_build(gi_ctx* ctx, const double* box6);
/* structure readback for parity tests: nodes in DFS pre-order like the reference's recursion */
int gi_photon_map_info(gi_ctx* ctx, uint32_t* n_nodes, uint32_t* n_leaves, uint32_t* n_kept_photons, uint32_t* max_depth);
int gi_photon_map_download(gi_ctx* ctx, double* node_box6, uint8_t* node_is_leaf, uint32_t* node_count,
                           uint32_t* photon_ids);
/* serialise / adopt a built map as one slab (what gets broadcast to the other GPUs) */
int gi_photon_map_slab_size(gi_ctx* ctx, size_t* bytes);
int gi_photon_map_slab_ptr(gi_ctx* ctx, void** dev_ptr);
int gi_photon_map_adopt_slab(gi_ctx* ctx, size_t bytes);   /* after the slab bytes were written to slab_ptr */
int gi_photon_map_reserve_slab(gi_ctx* ctx, size_t bytes, void** dev_ptr);

/* ---- gather: RayTracer::samplePhotons (raytracer.h:532-579) over PhotonMap::getInRange (photonMap.cpp:50-92,
 *      115-134).  rgb [n][3]; knn (optional) [n][k] photon ids in ascending distance, GI_NO_HIT padded;
 *      n_cand (optional) [n] candidate count returned by getInRange. ------------------------------------------ */
int gi_photon_gather(gi_ctx* ctx, size_t n, const double* pos, const double* dir, int k, double* rgb, uint32_t* knn,
                     uint32_t* n_cand);
int gi_photon_gather_dev(gi_ctx* ctx, size_t n, const double* pos, const double* dir, int k, double* rgb, uint32_t* knn,
                         uint32_t* n_cand);

/* ---- frame: the row loop of RayTracer::run (raytracer.h:93-160) for the pixel rectangle and sample range.
 *      accum [(y1-y0)*(x1-x0)][3] fp64 receives the SUM over s in [s0,s1) of radiance(); host pointer. ---------- */
int gi_render_tile(gi_ctx* ctx, const gi_render_params* p, int x0, int y0, int x1, int y1, int s0, int s1, double* accum,
                   gi_stats* stats);
int gi_render_tile_dev(gi_ctx* ctx, const gi_render_params* p, int x0, int y0, int x1, int y1, int s0, int s1, double* accum,
                       gi_stats* stats);
/* RayTracer::run's frame as it reaches the Image (raytracer.h:93-160 + image.h:14-16): samples [s0,s1) rendered and resolved
 * (running mean over s1 - s0 samples, gamma 2.2, clamp, (int)(255c)) on the device; rgb8 [(y1-y0)*(x1-x0)][3] comes back,
 * accum (optional, may be NULL) = the fp64 sums gi_render_tile would give. */
int gi_render_image(gi_ctx* ctx, const gi_render_params* p, int x0, int y0, int x1, int y1, int s0, int s1, uint8_t* rgb8,
                    double* accum, gi_stats* stats);
/* ---- resolve: mean -> gamma 2.2 -> clamp -> (int)(255*c)  (raytracer.h:150-156, util.h:94-97, image.h:14-16) */
/* ---- tile split of the frame (the loop being split is the row loop of RayTracer::run, raytracer.h:93-160, whose rows the
 *      reference hands to its threads `schedule(dynamic, 10)`): part `part` of `nparts` owns the row blocks part, part + nparts, ..
 *      of block_rows rows each.  gi_rows_of_part = how many rows that is (or GI_ERR_INVALID); gi_render_rows renders them in ONE
 *      wavefront into accum [local rows][width][3] (compact, block after block) — the same pixels, bit for bit, as the one-GPU frame. */
int gi_rows_of_part(int height, int block_rows, int nparts, int part);
int gi_render_rows(gi_ctx* ctx, const gi_render_params* p, int block_rows, int nparts, int part, int s0, int s1, double* accum, gi_stats* stats);
int gi_render_rows_dev(gi_ctx* ctx, const gi_render_params* p, int block_rows, int nparts, int part, int s0, int s1, double* accum, gi_stats* stats);

/* Adaptive sampling: the per-pixel loop of RayTracer::run (raytracer.h:100-148) with `samples min max thresh`.  Every pixel
 * of the tile takes samples s = 0, 1, .. while s < max_samples && samps < min_samples (samps +1 per sample, -2 when the
 * smoothed change of the running mean stays above noise_thresh).  color[n_pixels][3] receives the final running-mean colour
 * (resolve it with spp = 1), samples[n_pixels] (may be NULL) the samples taken.  p->spp is ignored. */
int gi_render_adaptive(gi_ctx* ctx, const gi_render_params* p, int min_samples, int max_samples, double noise_thresh, int x0, int y0, int x1, int y1,
                       double* color, uint32_t* samples, gi_stats* stats);
int gi_render_adaptive_dev(gi_ctx* ctx, const gi_render_params* p, int min_samples, int max_samples, double noise_thresh, int x0, int y0, int x1, int y1,
                           double* color, uint32_t* samples, gi_stats* stats);
int gi_resolve(gi_ctx* ctx, size_t n_pixels, const double* accum, int spp, uint8_t* rgb8);
int gi_resolve_dev(gi_ctx* ctx, size_t n_pixels, const double* accum, int spp, uint8_t* rgb8);

/* ---- multi-GPU (new; the reference is a single process): one gi_ctx per GPU, NCCL over NVLink bound at run time (dlopen), used only
 *      for the two exchanges the path has — the photon map "built once and broadcast" (the reference builds it once per scene,
 *      raytracer.h:61-71) and the framebuffer at the end of a frame.  Everything runs on gi_stream(ctx).
 *      gi_comm_unique_id: 128 bytes made on one rank (ncclGetUniqueId); the caller hands them to the others (file, socket, MPI,
 *        torch.distributed, a shared variable between threads);
 *      gi_comm_init: collective over the nranks contexts (ncclCommInitRank); one communicator per ctx;
 *      gi_photon_map_bcast: the root's built map (one slab) to every rank, adopted after its header is validated;
 *      gi_framebuffer_reduce: sample split — fp64 partial sums (count doubles, device pointer) added onto the root, in place;
 *      gi_framebuffer_gather: tile split — each rank's compact rows (gi_render_rows' plan; row_bytes per image row: 3 * width for the
 *        resolved 8-bit image, 24 * width for the fp64 sums; device pointers) into the root's frame [height][row_bytes]. ------------ */
#define GI_COMM_ID_BYTES 128
int gi_comm_unique_id(void* id, size_t bytes);
int gi_comm_init(gi_ctx* ctx, const void* id, size_t bytes, int rank, int nranks);
int gi_comm_destroy(gi_ctx* ctx);
int gi_comm_info(gi_ctx* ctx, int* rank, int* nranks, int* nccl_version);
int gi_photon_map_bcast(gi_ctx* ctx, int root);
int gi_framebuffer_reduce(gi_ctx* ctx, double* accum_dev, size_t count, int root);
int gi_framebuffer_gather(gi_ctx* ctx, const void* local_dev, size_t row_bytes, int height, int block_rows, void* frame_dev, int root);
/* the tile split as one call per rank (RayTracer::run with several GPUs): render this context's rows (part = its rank in its communicator;
 * no communicator = the whole frame), resolve them on the device, gather the 8-bit rows on `root`, copy the frame to the root's HOST
 * buffer rgb8 [height][width][3] (ignored elsewhere).  Collective over the communicator. */
int gi_render_rows_image(gi_ctx* ctx, const gi_render_params* p, int block_rows, int s0, int s1, uint8_t* rgb8, int root, gi_stats* stats);

/* ---- kernel-level hooks for bench.py's roofline (device buffers owned by the ctx) -------------------------- */
/* average duration (ms) of the last launches of the named kernel family, measured with CUDA events on gi_stream */
int gi_last_kernel_ms(gi_ctx* ctx, const char* family, double* ms, uint64_t* launches);
/* device-side work tallies of the last call of a family: "trace_closest"/"trace_any": out = {rays, box tests, primitive
 * tests, 0}; "gather": out = {queries, sum leaf depth, sum candidates, sum selected} */
int gi_last_work(gi_ctx* ctx, const char* family, uint64_t out[4]);

#ifdef __cplusplus
}
#endif
#endif /* GI_API_H */

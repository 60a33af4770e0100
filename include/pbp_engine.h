/*
 * pbp_engine.h -- C ABI of the B200 batched collision-validation engine (libpbp_engine.so).
 *
 * This is the drop-in boundary for pyroboplan's collision hot path.  The reference has no
 * native FFI of its own for this path: it reaches Pinocchio/coal through boost.python from
 * the free functions in src/pyroboplan/core/utils.py and src/pyroboplan/planning/utils.py.
 * Each entry point below cites the reference call it replaces (paths relative to
 * /root/reference/).  The Python host side (pyroboplan_b200/engine.py) binds exactly these
 * symbols with ctypes; INTEGRATION.md shows the stub a pyroboplan maintainer would add.
 *
 * Conventions
 *  - plain pointers and sizes only; no torch / C++ types in signatures;
 *  - every function returns 0 on success, a negative pbp_status otherwise, and
 *    pbp_last_error() returns a thread-local message for the last failure;
 *  - "_dev" arguments are CUDA device pointers on the engine's device, "_host" arguments
 *    are host pointers; `stream` is a cudaStream_t passed as void* (NULL = default stream);
 *  - device-pointer entry points are asynchronous on `stream`; host-pointer entry points
 *    (suffix _host) copy in, run, copy out and synchronise before returning;
 *  - one engine owns one set of device workspaces (item bins, control block): its entry points serialise
 *    internally -- a mutex around every call, and an event that makes a call on another stream wait for the
 *    previous call's kernels -- so an engine may be shared by host threads and CUDA streams (results are ready
 *    in stream order); calls on ONE engine never overlap on the device, use one engine per stream for that.
 *    Every entry point leaves the caller's current CUDA device as it found it;
 *  - configurations are row-major float32 [N, nq]; SE(3) placements are 12 floats:
 *    rotation row-major (9) then translation (3);
 *  - there is NO CPU fallback: without a CUDA device creation fails with PBP_ERR_CUDA.
 */
#ifndef PBP_ENGINE_H
#define PBP_ENGINE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PBP_ABI_VERSION 6

typedef enum {
    PBP_OK = 0,
    PBP_ERR_ARG = -1,      /* bad argument (null pointer, negative size, unsupported value) */
    PBP_ERR_CUDA = -2,     /* CUDA runtime error, including "no device" */
    PBP_ERR_CAPACITY = -3, /* model exceeds a compiled-in table limit */
    PBP_ERR_ALLOC = -4     /* device or pinned-host allocation failed */
} pbp_status;

typedef enum { PBP_SHAPE_SPHERE = 0, PBP_SHAPE_BOX = 1, PBP_SHAPE_CYLINDER = 2, PBP_SHAPE_CAPSULE = 3, PBP_SHAPE_CONVEX = 4 } pbp_shape_type;
typedef enum { PBP_JOINT_FIXED = 0, PBP_JOINT_REVOLUTE = 1, PBP_JOINT_PRISMATIC = 2 } pbp_joint_type;

/* Compile-time table limits of the kernels. */
#define PBP_MAX_JOINTS 64
#define PBP_MAX_GEOMS 128
#define PBP_MAX_PAIRS 4096
#define PBP_MAX_NQ 64

/*
 * Flat model description (host pointers, copied at creation).  Replaces the pinocchio.Model /
 * pinocchio.GeometryModel pair every reference hot-path function takes
 * (src/pyroboplan/core/utils.py:8-15).  Joint 0 is the universe; joints are topologically
 * ordered (parent < child).  Geometry g is rigidly attached to joint geom_parent[g] with
 * placement geom_placement[g] (for geom_parent == 0 that is its world placement).
 */
typedef struct {
    int32_t nq, njoints, ngeoms, npairs, nverts, nframes;
    const int32_t *joint_parent;   /* [njoints] */
    const int32_t *joint_type;     /* [njoints] pbp_joint_type */
    const int32_t *joint_idx_q;    /* [njoints] index into q, -1 for the universe */
    const float *joint_placement;  /* [njoints][12] placement in the parent joint frame */
    const float *joint_axis;       /* [njoints][3] unit axis */
    const float *q_lower;          /* [nq] */
    const float *q_upper;          /* [nq] */
    const int32_t *geom_parent;    /* [ngeoms] parent joint */
    const int32_t *geom_shape;     /* [ngeoms] pbp_shape_type */
    const float *geom_placement;   /* [ngeoms][12] */
    const float *geom_params;      /* [ngeoms][4] sphere: r | box: hx hy hz | cylinder, capsule: r, half length */
    const int32_t *geom_vert_off;  /* [ngeoms] first hull vertex (convex shapes) */
    const int32_t *geom_vert_cnt;  /* [ngeoms] hull vertex count */
    const float *geom_outer;       /* [ngeoms][4] enclosing sphere (cx cy cz r), geometry frame */
    const float *geom_inner;       /* [ngeoms][4] inscribed sphere, SAME centre as geom_outer (cx cy cz r) */
    const float *verts;            /* [nverts][3] hull vertices, geometry frame */
    const int32_t *pairs;          /* [npairs][2] active collision pairs (first, second) */
    const int32_t *frame_parent;   /* [nframes] parent joint of each frame (may be NULL if nframes == 0) */
    const float *frame_placement;  /* [nframes][12] */
    /* float64 copies of the kinematic tables for the FP64 differential-IK kernel (each may be
     * NULL: the float32 table above is widened instead) */
    const double *joint_placement_f64; /* [njoints][12] */
    const double *joint_axis_f64;      /* [njoints][3] */
    const double *frame_placement_f64; /* [nframes][12] */
    const double *q_lower_f64;         /* [nq] */
    const double *q_upper_f64;         /* [nq] */
    /* optional (may be NULL): an enclosing capsule per geometry, geometry frame -- p0 (3), p1 (3),
     * radius, flag (!= 0: use it).  A second "certainly free" proxy of the broadphase for elongated
     * shapes; it must contain the whole shape. */
    const float *geom_capsule;         /* [ngeoms][8] */
    /* optional (all NULL / 0: every geometry is a solid): SURFACE semantics of triangle meshes.
     * Pinocchio loads a URDF <mesh> as a coal BVHModel -- the surface triangles, not the solid
     * (pinocchio.buildModelsFromUrdf, src/pyroboplan/models/panda.py:27): two meshes collide when two of
     * their triangles come within the security margin, a primitive (a solid) collides with a mesh when it
     * comes within the margin of a triangle, and a shape that lies wholly inside a mesh without reaching
     * its surface does NOT collide with it (1 in 1000 uniform Panda states: a finger inside link 0 / 1 /
     * 2 / 5 while the hand - link pair is disabled by the SRDF).  The engine brackets a mesh M between
     * two convex polytopes -- the hull H of its vertices (`verts`) and an inner core K inside the mesh
     * (`core_verts`: a subset of the vertices of the intersection of the inner half-spaces of all
     * triangles) -- runs its GJK kernels on those, and evaluates the triangles themselves (exact
     * triangle - triangle / triangle - primitive distances) for the items the two bounds leave open.
     *   planes            faces of the exact core {x : n.x <= off}, unit normals, geometry frame: the hull's
     *                     face planes first (geom_plane_hull_cnt of them, largest face first), then the
     *                     planes of mesh triangles that lie below the hull
     *   geom_surface_tol  two-sided distance between the hull's boundary and the mesh surface (metres)
     *   pair_enclose      bit 0: `second` could lie wholly inside `first`, and `first` is a surface;
     *                     bit 1: `first` could lie wholly inside `second`, and `second` is a surface
     *   core_verts        vertices of K; geom_core_cnt == 0: the geometry has no core (K = H is used when
     *                     the geometry is a solid; a surface without a core never certifies a hit by GJK)
     *   geom_core_eps     every point of H is within this distance of K (metres)
     *   mesh_verts, tris  the triangles: tris[t] = three vertex indices relative to geom_mvert_off */
    const float *planes;               /* [nplanes][4] nx ny nz off */
    const int32_t *geom_plane_off;     /* [ngeoms] first plane of the geometry */
    const int32_t *geom_plane_cnt;     /* [ngeoms] number of planes (0: none given) */
    const float *geom_surface_tol;     /* [ngeoms] */
    const uint8_t *pair_enclose;       /* [npairs] */
    int32_t nplanes;
    int32_t ncore, nmverts, ntris;
    const int32_t *geom_plane_hull_cnt; /* [ngeoms] */
    const float *core_verts;           /* [ncore][3] */
    const int32_t *geom_core_off;      /* [ngeoms] */
    const int32_t *geom_core_cnt;      /* [ngeoms] */
    const float *geom_core_eps;        /* [ngeoms] */
    const float *mesh_verts;           /* [nmverts][3] */
    const int32_t *geom_mvert_off;     /* [ngeoms] */
    const int32_t *tris;               /* [ntris][3] */
    const int32_t *geom_tri_off;       /* [ngeoms] */
    const int32_t *geom_tri_cnt;       /* [ngeoms] 0: the geometry is a solid */
    /* optional (may be NULL): up to four balls per geometry (cx cy cz r, geometry frame; r = 0: unused) that lie
     * INSIDE the solid -- inside the inner core K for a surface geometry.  Two geometries whose balls overlap
     * certainly collide: the item setup settles such pairs before any GJK iteration is spent on them.  Boxes
     * need none (the test against a box is exact). */
    const float *geom_inner_balls;     /* [ngeoms][4][4] */
} pbp_model_desc;

typedef struct pbp_engine pbp_engine;

/* Work counters since creation / last reset (for bench.py: gpu_launches, items per state). */
typedef struct {
    int64_t kernel_launches;   /* kernels launched by this engine */
    int64_t states;            /* configurations pushed through FK + broadphase */
    int64_t narrowphase_items; /* (state, pair) items that reached GJK (only counted by *_host / sync calls) */
    int64_t chunks;            /* broadphase / narrowphase chunk iterations */
} pbp_stats;

const char *pbp_last_error(void);
int pbp_abi_version(void);

/* Upload a model to `device`, allocate workspaces.  max_chunk_states <= 0 picks a default.
 * Replaces model.createData() / collision_model.createData() (core/utils.py:39-42). */
int pbp_engine_create(const pbp_model_desc *desc, int device, int64_t max_chunk_states, pbp_engine **out);
void pbp_engine_destroy(pbp_engine *e);
int pbp_engine_device(const pbp_engine *e);
int pbp_get_stats(pbp_engine *e, pbp_stats *out, int reset);

/* Optional per-kernel timing (CUDA events recorded on the caller's stream around the two hot
 * kernels of every round) and narrowphase item counting.  Off by default; bench.py turns it
 * on to report the dominant kernel's average launch duration for the roofline figure. */
typedef struct {
    double broadphase_ms;      /* sum of k_broadphase launch durations */
    double item_setup_ms;      /* sum of k_item_setup launch durations */
    double narrowphase_ms;     /* sum of k_narrowphase launch durations */
    int64_t rounds;            /* broadphase + narrowphase rounds timed */
    int64_t narrowphase_items; /* (state, pair) items that reached GJK */
    double surface_refine_ms;  /* sum of k_surface_refine launch durations (0 without surface semantics) */
} pbp_profile;
int pbp_set_profiling(pbp_engine *e, int enabled);
/* Synchronises the recorded events, accumulates, returns totals since the last reset. */
int pbp_get_profile(pbp_engine *e, pbp_profile *out, int reset);

/* Measured FP32 FMA throughput of the device (dependent-chain-free FFMA loop on every SM),
 * in TFLOP/s; used as the roofline denominator of this FP32-pipe-bound path. */
int pbp_measure_fp32_peak(int device, double *tflops_out);

/* Forward kinematics + placement of joints / geometries / frames for N configurations.
 * Replaces pinocchio.framesForwardKinematics + updateGeometryPlacements
 * (core/utils.py:253,297,345,374; implicit in computeCollisions at core/utils.py:48).
 * Any of the three outputs may be NULL.  Layouts: oMi [N][njoints][12], oMg [N][ngeoms][12],
 * oMf [N][nframes][12]. */
int pbp_fk(pbp_engine *e, const float *q_dev, int64_t n, float *oMi_dev, float *oMg_dev, float *oMf_dev, void *stream);

/* Boolean collision (+ optional min distance / first colliding pair) for N configurations.
 * Replaces check_collisions_at_state (core/utils.py:8-51) and, with min_dist,
 * get_minimum_distance_at_state (core/utils.py:54-87).
 *   hit[i]       = 1 iff some active pair has distance <= padding (padding >= 0)
 *   min_dist[i]  = min over pairs of max(distance, 0); +inf when the model has no pairs  (nullable)
 *   first_pair[i]= smallest pair index in collision, -1 if none                          (nullable) */
int pbp_check_states(pbp_engine *e, const float *q_dev, int64_t n, float padding, uint8_t *hit_dev, float *min_dist_dev,
                     int32_t *first_pair_dev, void *stream);
int pbp_check_states_host(pbp_engine *e, const float *q_host, int64_t n, float padding, uint8_t *hit_host,
                          float *min_dist_host, int32_t *first_pair_host);
/* The same from FLOAT64 host states -- what the reference's callers hold (numpy float64, pageable memory;
 * core/utils.py:8-15 takes q as such).  A few host threads narrow slice after slice into rotating pinned
 * float32 buffers while earlier slices cross PCIe and are checked: conversion, copies and kernels overlap. */
int pbp_check_states_host_f64(pbp_engine *e, const double *q_host, int64_t n, float padding, uint8_t *hit_host,
                              float *min_dist_host, int32_t *first_pair_host);

/* Edge validation: edge e = straight joint-space segment q1[e] -> q2[e], discretised on device
 * with ceil(||dq||_2 / step) + 1 samples (both ends included) and collision-checked.
 * Replaces has_collision_free_path (planning/utils.py:51-111) =
 * discretize_joint_space_path (planning/utils.py:114-140) + check_collisions_along_path
 * (core/utils.py:90-132).  hit[e] = 1 iff any sample collides (edge NOT free).
 *   first_bad[e] = smallest colliding sample index, -1 if none (nullable; disables early-out)
 * Samples are tested coarse to fine (every 32nd sample, then every 4th, then the rest; edges already
 * known to collide are dropped between levels); batches of <= 16384 edges use two levels, <= 512 one, because
 * every level costs a device -> host round trip.  This entry point therefore SYNCHRONISES `stream`
 * internally; the _host variant needs no intermediate round trip for <= 512 edges. */
int pbp_check_edges(pbp_engine *e, const float *q1_dev, const float *q2_dev, int64_t n_edges, double step, float padding,
                    uint8_t *hit_dev, int32_t *first_bad_dev, void *stream);
int pbp_check_edges_host(pbp_engine *e, const float *q1_host, const float *q2_host, int64_t n_edges, double step,
                         float padding, uint8_t *hit_host, int32_t *first_bad_host);

/* Path discretisation on device (planning/utils.py:114-140), two calls:
 *   counts[e]  = ceil(||q2[e]-q1[e]|| / step) + 1
 *   q_out[offsets[e] + i] = q1[e] + (i / (counts[e]-1)) * (q2[e] - q1[e]),  last sample s = 1
 * offsets = exclusive prefix sum of counts (caller computes it, e.g. torch.cumsum). */
int pbp_discretize_counts(pbp_engine *e, const float *q1_dev, const float *q2_dev, int64_t n_edges, double step,
                          int32_t *counts_dev, void *stream);
int pbp_discretize_fill(pbp_engine *e, const float *q1_dev, const float *q2_dev, int64_t n_edges,
                        const int32_t *counts_dev, const int64_t *offsets_dev, float *q_out_dev, void *stream);

/* Pre-discretised path (one group of samples): replaces check_collisions_along_path
 * (core/utils.py:90-132) for a batch of paths given as a ragged array.
 *   q [total][nq], path_offsets [n_paths + 1] (int64, device), hit[p] = 1 iff any sample collides. */
int pbp_check_paths(pbp_engine *e, const float *q_dev, const int64_t *path_offsets_dev, int64_t n_paths, int64_t total,
                    float padding, uint8_t *hit_dev, void *stream);

/* Rejection sampling of collision-free configurations on device: n uniform draws in
 * [lower + joint_padding, upper - joint_padding] per round (counter-based Philox stream
 * keyed by `seed`), collision-checked; free ones are compacted in draw order into q_out until
 * n_wanted are found or max_rounds rounds ran.  *n_found_host receives the count.
 * Batched sibling of get_random_collision_free_state (core/utils.py:195-229). */
int pbp_sample_free_states(pbp_engine *e, int64_t n_wanted, uint64_t seed, float joint_padding, float padding,
                           int max_rounds, float *q_out_dev, int64_t *n_found_host, int64_t *n_drawn_host, void *stream);

/* Distance, witness points and normal of EVERY active pair for N configurations (FP32 GJK with
 * witness tracking, thread per (pair, state)).  Replaces pinocchio.computeDistances as the
 * reference reads its results -- distanceResults[k].min_distance, getNearestPoint1/2(), normal
 * (src/pyroboplan/core/utils.py:505-516; src/pyroboplan/ik/nullspace_components.py:124-131).
 * Outputs are PAIR-MAJOR in the caller's pair order: dist_dev[k*N + i] (signed); p1/p2/normal_dev
 * [(k*N + i)*3 ..] in the world frame (each of the three may be NULL).
 *   cutoff   pairs certainly farther apart than this return a lower bound of their distance
 *            (> cutoff) and zero witness data; INFINITY = exact distances everywhere
 * Overlapping pairs report coal's convention: a NEGATIVE distance (the penetration depth, FP32
 * Expanding Polytope Algorithm), the deepest points and the normal; p2 - p1 = normal * distance. */
int pbp_compute_distances(pbp_engine *e, const float *q_dev, int64_t n, float cutoff, float *dist_dev, float *p1_dev,
                          float *p2_dev, float *normal_dev, void *stream);

/* Collision-avoidance null-space term for N configurations:
 *   component[i] = gain * sum_{pairs k with dist <= dist_padding} -(normal_k dist_k - dist_padding) @ (Jcoll2 - Jcoll1)
 * Replaces collision_avoidance_nullspace_component (src/pyroboplan/ik/nullspace_components.py:82-142)
 * including calculate_collision_vector_and_jacobians (src/pyroboplan/core/utils.py:478-538).
 *   component_dev  [N][nq] float32 */
int pbp_collision_nullspace(pbp_engine *e, const float *q_dev, int64_t n, float dist_padding, float gain,
                            float *component_dev, void *stream);

/* k nearest neighbours in joint space of every node among the other nodes, for batched PRM
 * construction.  Replaces one Graph.get_nearest_neighbors call per node
 * (src/pyroboplan/planning/graph.py:271-301, called from PRMPlanner.connect_node,
 * src/pyroboplan/planning/prm.py:186-190): neighbours with distance <= radius, sorted by distance
 * (ties keep index order), first k; the node itself is never returned.
 *   nodes_dev          [n][nq] float32
 *   predecessors_only  != 0: node i only sees nodes j < i (the reference's incremental roadmap)
 *   idx_dev            [n][k] int32, -1 where fewer than k neighbours qualify
 *   dist_dev           [n][k] float32 distances (inf where idx is -1)
 * Distances accumulate in FP64.  k <= 64. */
int pbp_knn(pbp_engine *e, const float *nodes_dev, int64_t n, int32_t k, double radius, int32_t predecessors_only,
            int32_t *idx_dev, float *dist_dev, void *stream);

/* The same for every block_stride-th block of 16 consecutive query nodes, starting with block block_first:
 * query block b = block_first + t * block_stride (t = 0, 1, ...) writes LOCAL rows [16 t, 16 t + 16) of idx_dev /
 * dist_dev ([ceil((nblocks - block_first) / block_stride) * 16][k]).  The ranks of a multi-GPU roadmap build
 * interleave the query blocks (rank r: block_first = r, block_stride = world) -- the work of a predecessor query
 * grows with its index, interleaving balances it -- and all-gather the rows. */
#define PBP_KNN_QUERY_BLOCK 16
int pbp_knn_blocks(pbp_engine *e, const float *nodes_dev, int64_t n, int32_t k, double radius, int32_t predecessors_only,
                   int64_t block_first, int64_t block_stride, int32_t *idx_dev, float *dist_dev, void *stream);

/* One try of damped-least-squares differential IK for N targets (thread per target, FP64 end to
 * end -- unlike the collision entry points the states here are DOUBLES, because an iterate that
 * crosses a singularity amplifies input rounding by up to 1/damping^2).
 * Restates the inner loop of DifferentialIk.solve (src/pyroboplan/ik/differential_ik.py:208-298):
 * up to max_iters iterations of  FK -> error = -log6(target^-1 * frame pose) -> converged? ->
 * LOCAL frame Jacobian -> (J W J^T + damping^2 I) y = error -> q += alpha W J^T y.
 *   frame_id      index into the model's frames (pbp_model_desc.frame_*)
 *   targets_dev   [N][12] desired frame placements (float64)
 *   q_dev         [N][nq] float64, in: initial states, out: final states (every coordinate wrapped
 *                 to [-pi, pi) when converged, differential_ik.py:221)
 *   active_mask   bit k set = joint coordinate k may move (ignore_joint_indices cleared)
 *   w_diag_host   [nq] diagonal of W = inverse joint weights (NULL = identity)
 *   nullspace_dev [N][nq] float64 null-space term, constant for the iterations of this call
 *                 (NULL = none): the step becomes alpha (W J^T solve(jjt, error - J ns) + ns) on the
 *                 active coordinates (:274-283).  Terms that depend on q (joint limits, collision
 *                 avoidance) are refreshed by calling with max_iters = 1 once per iteration.
 *   e0_dev        [N] in/out initial error norm of each solve (0 = not captured yet; it is kept
 *                 across retries, as the reference does, :200,264-265)
 *   status_dev    [N] bit 0: pose error within tolerance, bit 1: within joint limits (after wrapping)
 *   iters_dev     [N] iterations used (nullable)
 * The collision gate and the random restarts stay with the caller (pbp_check_states,
 * pbp_sample_free_states), exactly where the reference has them (:223-230, :306-311). */
typedef struct {
    int32_t max_iters;
    double max_translation_error, max_rotation_error, damping, min_step_size, max_step_size;
} pbp_ik_options;
int pbp_ik_iterate(pbp_engine *e, int32_t frame_id, const double *targets_dev, double *q_dev, int64_t n, const pbp_ik_options *opt,
                   uint64_t active_mask, const double *w_diag_host, const double *nullspace_dev, double *e0_dev,
                   int32_t *status_dev, int32_t *iters_dev, void *stream);

/* Introspection of the model-specialised broadphase (pyroboplan_b200/csrc/pbp_jit.h): boolean
 * queries run a kernel whose CUDA source is generated from the model (joint chain as straight-line
 * code, pair tests unrolled) and built with NVRTC when the engine is created.  These two entry
 * points need no CUDA device; they exist for tests and tooling (tools/jit_dump.py).
 *   pbp_jit_broadphase_source  writes the generated source (NUL terminated) when cap >= *needed;
 *                              *needed = 1 (empty string) when the model does not qualify
 *   pbp_jit_compile_check      NVRTC build for sm_<major><minor>a; *cubin_bytes = size of the result */
int pbp_jit_broadphase_source(const pbp_model_desc *desc, char *buf, size_t cap, size_t *needed);
int pbp_jit_compile_check(const pbp_model_desc *desc, int cc_major, int cc_minor, size_t *cubin_bytes, char *log_buf, size_t log_cap);

#ifdef __cplusplus
}
#endif
#endif /* PBP_ENGINE_H */

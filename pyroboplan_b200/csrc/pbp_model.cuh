// Device-side model tables and small SE(3) helpers (sm_100a, scalar FP32).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pbp_engine.h"

namespace pbp {

// SE(3) as 12 floats: rotation row-major r[0..8], translation r[9..11].

struct DevJoint {       // 96 B
    float P[12];        // placement in parent joint frame
    float axis[3];
    int32_t parent;
    int32_t type;
    int32_t idx_q;
    int32_t axis_code;  // 0 generic axis, 1/2/3 = +-x/y/z aligned (sign in axis[])
    int32_t parent_slot;  // broadphase: -1 parent is the previous joint (running product), -2 universe, >= 0 saved slot
    int32_t save_slot;    // broadphase: >= 0 -> keep this joint's transform in that slot (it has a non-adjacent child)
    int32_t geom_begin, geom_end;  // range in DevModel::joint_geoms of the moving geometries attached to this joint
    int32_t _pad;
};

constexpr int N_INNER_BALLS = 4;

struct __align__(16) DevGeom {        // 256 B -- narrowphase / item-setup view of a geometry
    float P[12];        // placement in parent joint frame (world placement if parent == 0)
    float par[4];       // shape parameters
    float centre[3];    // proxy-sphere centre in the geometry frame (initial GJK direction)
    float core_radius;  // sphere / capsule inflation applied outside GJK
    int32_t parent;
    int32_t shape;
    int32_t vert_off;   // offset into the padded float4 vertex table
    int32_t vert_cnt;   // padded to a multiple of 4 (last vertex repeated)
    int32_t map_off;    // convex: offset of this hull's support-map cell table (uint16 units), -1 if none
    int32_t list_off;   // convex: offset of this hull's candidate lists (uint16 units)
    int32_t has_axis;   // != 0: axis0 / axis1 is the segment of the shape's enclosing capsule (geometry frame)
    int32_t _pad;
    float axis0[3];     // used for the GJK start direction of elongated shapes
    float r_out;        // enclosing radius around `centre` (culls of the triangle stage)
    float axis1[3];
    float core_eps;     // surface geometry: every point of the hull is within this of the inner core (0: solid)
    // Surface semantics (pbp_surface.cuh).  A mesh geometry is bracketed by its hull H (vert_* above) and its
    // inner core K (kvert_* below, a convex polytope inside the mesh); solids have K == H.
    int32_t kvert_off;  // offset into the padded float4 core vertex table
    int32_t kvert_cnt;
    int32_t kmap_off;   // support map of the core (as map_off / list_off)
    int32_t klist_off;
    int32_t tri_off;    // first triangle (DevModel::tris), indices relative to mvert_off
    int32_t tri_cnt;    // 0: the geometry is a solid
    int32_t mvert_off;  // first mesh vertex (DevModel::mverts)
    int32_t plane_off;  // planes of the core: the hull's face planes first (largest face first) ...
    int32_t plane_cnt;
    int32_t plane_hull_cnt;   // ... this many; the rest are planes of mesh triangles below the hull
    float surface_tol;  // two-sided distance between the hull's boundary and the mesh surface
    int32_t n_balls;    // balls[0 .. n_balls) lie inside the solid / inside the core of a surface geometry
    float balls[N_INNER_BALLS][4];   // cx cy cz r, geometry frame
};

struct BpGeom {         // 64 B -- broadphase view: one proxy centre, two radii (+ optional enclosing capsule)
    float c[3];         // centre: parent-JOINT frame for moving geometries, world frame for static ones
    float r_out;        // enclosing radius
    float r_in;         // inscribed radius (same centre)
    int32_t slot;       // per-thread centre slot, -1 if static
    int32_t box;        // index into the static-box table, -1 if this is not a static box
    int32_t has_cap;    // != 0: cap0 / cap1 / cap_r describe a capsule that contains the shape
    float cap0[3];      // capsule segment end points, same frame as c
    float cap_r;
    float cap1[3];
    float _pad;
};

struct BpBox {          // 64 B -- static box: world placement + half sides
    float R[9];
    float t[3];
    float h[3];
    float _pad;
};

// Pairs are renumbered internally (DevModel::pair_orig maps back to the caller's index): first the
// pairs between two moving geometries, then the pairs against each static geometry as one
// contiguous group per static geometry, so the broadphase loads an obstacle once per group.
struct BpPair {         // 16 B
    int16_t slot_a, slot_b;  // per-thread centre slots of a / b, -1 if static
    uint8_t a, b;            // geometry ids (first, second)
    uint8_t kind;            // BP_*
    uint8_t _pad;
    float r_out;             // BP_SPHERES: r_out(a) + r_out(b) | BP_BOX_*: r_out of the non-box geometry
    float r_in;              // same for the inscribed radii
};

// Broadphase recipes (warp-uniform per pair).
enum : uint8_t {
    BP_SPHERES = 0,        // enclosing spheres for "free", inscribed spheres for "hit"
    BP_BOX_A = 1,          // a is a static box: exact point-to-box distance against b's spheres
    BP_BOX_B = 2,          // b is a static box
    BP_EXACT_SPHERES = 3,  // both shapes are spheres: the proxies are the shapes, decided exactly
};

// A run of consecutive (internal) pairs that share how their proxy distance is computed.
enum : int32_t {
    GRP_MOVING = 0,        // both geometries move: centre - centre
    GRP_STATIC_POINT = 1,  // one static sphere-proxy (centre loaded once) vs moving geometries
    GRP_STATIC_BOX = 2,    // one static box (placement loaded once) vs moving geometries
    GRP_GENERIC = 3,       // anything else (static vs static, ...)
};
struct BpGroup {        // 16 B
    int32_t cls;        // GRP_*
    int32_t sidx;       // GRP_STATIC_POINT: static geometry id | GRP_STATIC_BOX: static-box table index
    int32_t begin, end; // internal pair range
};

// Joint path of a pair for the item-setup FK: trunk (root .. lowest common ancestor), then the
// branch to a's joint, then the branch to b's joint.  Ops are joint indices.
struct PairPath {       // 8 B
    int32_t off;        // into DevModel::path_ops
    uint8_t n_trunk, n_a, n_b, _pad;
};

struct DevModel {
    int32_t nq, njoints, ngeoms, npairs, nverts_padded, nframes, nmoving, nsave, nboxes, ngroups;
    const DevJoint *joints;
    const int32_t *joint_geoms;  // moving geometry ids grouped by parent joint
    const DevGeom *geoms;
    const BpGeom *bpgeoms;
    const BpBox *bpboxes;
    const BpPair *bppairs;
    const BpGroup *groups;
    const int32_t *pair_orig;    // [npairs] internal pair index -> caller's pair index
    const PairPath *paths;
    const uint8_t *path_ops;
    const int32_t *bin_order;    // [npairs] pair indices sorted by decreasing narrowphase cost
    const float4 *verts;         // padded hull vertices
    const uint16_t *sup_cells;   // support maps: per hull SUPMAP_CELLS + 1 offsets
    const uint16_t *sup_lists;   // support maps: candidate vertex indices
    int32_t n_sup_cells, n_sup_lists;   // uint16 entries (padded to a multiple of 8)
    const float *q_lower, *q_upper;
    const int32_t *frame_parent;
    const float *frame_placement;
    // surface semantics of mesh geometries (pbp_model_desc.planes ...): hull face planes
    // (n.x <= w inside), per geometry (first plane, count), dent depth, per INTERNAL pair the
    // enclosure flags (bit 0: b could lie inside a's surface, bit 1: a inside b's); surface = any flag set
    const float4 *planes;
    const int2 *plane_idx;
    const float *surface_tol;
    const uint8_t *pair_enclose;
    int32_t surface, nplanes;
    // inner cores (boolean queries run GJK on them, distance queries on the hulls) and mesh triangles
    const float4 *kverts;        // padded core vertices (== verts when no geometry is a surface)
    const uint16_t *ksup_cells;
    const uint16_t *ksup_lists;
    int32_t nkverts_padded, n_ksup_cells, n_ksup_lists;
    int32_t has_tris;            // some geometry carries triangles: the narrowphase parks undecided items
    const float4 *mverts;        // mesh vertices (geometry frame)
    const int4 *tris;            // vertex indices (x, y, z) relative to the geometry's mvert_off
    const float4 *tverts;        // [ntris][3] the triangles' vertices, packed (what the triangle stage reads)
};

// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void se3_mul(const float *A, const float *B, float *C) {
    // C = A * B (C must not alias A or B)
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const float a0 = A[3 * i], a1 = A[3 * i + 1], a2 = A[3 * i + 2];
        C[3 * i + 0] = fmaf(a0, B[0], fmaf(a1, B[3], a2 * B[6]));
        C[3 * i + 1] = fmaf(a0, B[1], fmaf(a1, B[4], a2 * B[7]));
        C[3 * i + 2] = fmaf(a0, B[2], fmaf(a1, B[5], a2 * B[8]));
        C[9 + i] = fmaf(a0, B[9], fmaf(a1, B[10], fmaf(a2, B[11], A[9 + i])));
    }
}

// C = A^-1 * B
__device__ __forceinline__ void se3_inv_mul(const float *A, const float *B, float *C) {
    const float dx = B[9] - A[9], dy = B[10] - A[10], dz = B[11] - A[11];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const float a0 = A[i], a1 = A[3 + i], a2 = A[6 + i];  // column i of R_A = row i of R_A^T
        C[3 * i + 0] = fmaf(a0, B[0], fmaf(a1, B[3], a2 * B[6]));
        C[3 * i + 1] = fmaf(a0, B[1], fmaf(a1, B[4], a2 * B[7]));
        C[3 * i + 2] = fmaf(a0, B[2], fmaf(a1, B[5], a2 * B[8]));
        C[9 + i] = fmaf(a0, dx, fmaf(a1, dy, a2 * dz));
    }
}

__device__ __forceinline__ float3 se3_act(const float *T, float x, float y, float z) {
    return make_float3(fmaf(T[0], x, fmaf(T[1], y, fmaf(T[2], z, T[9]))),
                       fmaf(T[3], x, fmaf(T[4], y, fmaf(T[5], z, T[10]))),
                       fmaf(T[6], x, fmaf(T[7], y, fmaf(T[8], z, T[11]))));
}

// R^T (p - t)
__device__ __forceinline__ float3 se3_act_inv(const float *T, float3 p) {
    const float dx = p.x - T[9], dy = p.y - T[10], dz = p.z - T[11];
    return make_float3(fmaf(T[0], dx, fmaf(T[3], dy, T[6] * dz)), fmaf(T[1], dx, fmaf(T[4], dy, T[7] * dz)),
                       fmaf(T[2], dx, fmaf(T[5], dy, T[8] * dz)));
}

__device__ __forceinline__ void se3_identity(float *T) {
#pragma unroll
    for (int i = 0; i < 12; i++) T[i] = (i == 0 || i == 4 || i == 8) ? 1.f : 0.f;
}

// out = A * R where R rotates columns U and W (Rx: 1,2  Ry: 0,2 with the opposite sign  Rz: 0,1).
template <int U, int W, bool NEG>
__device__ __forceinline__ void rotate_columns(const float *A, float s, float c, float *out) {
    const float sg = NEG ? -s : s;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const float au = A[3 * i + U], aw = A[3 * i + W];
        out[3 * i + U] = fmaf(au, c, aw * sg);
        out[3 * i + W] = fmaf(aw, c, -au * sg);
        out[3 * i + (3 - U - W)] = A[3 * i + (3 - U - W)];
        out[9 + i] = A[9 + i];
    }
}

// oMi = oMparent * P * motion(q); `out` must not alias `parent`.
__device__ __forceinline__ void joint_transform(const DevJoint &J, const float *parent, float q, float *out) {
    float A[12];
    se3_mul(parent, J.P, A);
    if (J.type == PBP_JOINT_REVOLUTE && J.axis_code != 0) {
        // axis-aligned revolute: rotate two columns of A (rotation about -axis by q == about +axis by -q)
        float s, c;
        sincosf(q, &s, &c);
        if (J.axis_code == 1) rotate_columns<1, 2, false>(A, J.axis[0] * s, c, out);
        else if (J.axis_code == 2) rotate_columns<0, 2, true>(A, J.axis[1] * s, c, out);
        else rotate_columns<0, 1, false>(A, J.axis[2] * s, c, out);
    } else if (J.type == PBP_JOINT_REVOLUTE) {
        float s, c;
        sincosf(q, &s, &c);
        const float x = J.axis[0], y = J.axis[1], z = J.axis[2], C = 1.0f - c;
        float M[12];
        M[0] = fmaf(x * x, C, c);      M[1] = fmaf(x * y, C, -z * s); M[2] = fmaf(x * z, C, y * s);
        M[3] = fmaf(y * x, C, z * s);  M[4] = fmaf(y * y, C, c);      M[5] = fmaf(y * z, C, -x * s);
        M[6] = fmaf(z * x, C, -y * s); M[7] = fmaf(z * y, C, x * s);  M[8] = fmaf(z * z, C, c);
        M[9] = M[10] = M[11] = 0.0f;
        se3_mul(A, M, out);
    } else if (J.type == PBP_JOINT_PRISMATIC) {
        const float tx = J.axis[0] * q, ty = J.axis[1] * q, tz = J.axis[2] * q;
#pragma unroll
        for (int i = 0; i < 9; i++) out[i] = A[i];
#pragma unroll
        for (int i = 0; i < 3; i++) out[9 + i] = fmaf(A[3 * i], tx, fmaf(A[3 * i + 1], ty, fmaf(A[3 * i + 2], tz, A[9 + i])));
    } else {
#pragma unroll
        for (int i = 0; i < 12; i++) out[i] = A[i];
    }
}

}  // namespace pbp

// Device-side model tables and small SE(3) helpers (sm_100a, scalar FP32).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pbp_engine.h"

namespace pbp {

// SE(3) as 12 floats: rotation row-major r[0..8], translation r[9..11].
struct Se3 {
    float r[12];
};

struct DevJoint {    // 80 B
    float P[12];     // placement in parent joint frame
    float axis[3];
    int32_t parent;
    int32_t type;
    int32_t idx_q;
    int32_t chain;   // 1 if parent == index - 1 (previous transform can stay in registers)
    int32_t axis_code;  // 0 generic axis, 1/2/3 = +-x/y/z aligned (sign folded into axis_sign)
};

struct DevGeom {     // 128 B
    float P[12];     // placement in parent joint frame (world placement if parent == 0)
    float par[4];    // shape parameters
    float outer[4];  // enclosing sphere (geometry frame)
    float inner[4];  // inscribed sphere (geometry frame)
    int32_t parent;
    int32_t shape;
    int32_t vert_off;   // offset into the padded float4 vertex table
    int32_t vert_cnt;   // padded to a multiple of 4 (last vertex repeated)
    int32_t slot;       // index into the per-thread moving-geometry table, -1 if static
    int32_t identity;   // placement is the identity
    float core_radius;  // sphere / capsule inflation applied outside GJK
    int32_t _pad;
};

struct DevPair {     // 8 B
    int16_t a, b;    // geometry ids (first, second)
    int16_t kind;    // broadphase recipe, see BP_* below
    int16_t _pad;
};

// Broadphase recipes (warp-uniform per pair).
enum : int16_t {
    BP_SPHERES = 0,  // enclosing spheres for "free", inscribed spheres for "hit"
    BP_BOX_A = 1,    // a is a box: point-to-box distance in a's frame against b's spheres
    BP_BOX_B = 2,    // b is a box
    BP_EXACT_SPHERES = 3,  // both shapes are spheres: broadphase decides exactly
};

struct DevModel {
    int32_t nq, njoints, ngeoms, npairs, nverts_padded, nframes, nmoving, _pad;
    const DevJoint *joints;
    const DevGeom *geoms;
    const DevPair *pairs;
    const int32_t *bin_order;   // [npairs] pair indices sorted by decreasing narrowphase cost
    const float4 *verts;        // padded hull vertices
    const float *q_lower, *q_upper;
    const int32_t *frame_parent;
    const float *frame_placement;
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

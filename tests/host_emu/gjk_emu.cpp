// Host build of the engine's FP32 GJK (pyroboplan_b200/csrc/pbp_gjk.cuh) for CPU-side tests.
//
// TEST TOOLING ONLY: compiled by tests/ with g++ (no nvcc, no GPU), never loaded by the
// product.  It lets `pytest -m "not gpu"` exercise the exact float32 arithmetic of the device
// narrowphase (fmaf / IEEE sqrt / division are identical on host and device; only rsqrtf is
// emulated) against the float64 oracle on many random pairs.
#include <cmath>
#include <cstdint>
#include <cuda_runtime.h>

static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
static inline int __popc(int x) { return __builtin_popcount((unsigned)x); }

#include "../../pyroboplan_b200/csrc/pbp_gjk.cuh"

extern "C" {

// shape*: pbp_shape_type; par*: 3 floats; verts*: padded float4 table (n multiple of 4) or NULL.
// T: placement of B in A's frame; v0: initial direction.  Returns hit; writes dist and iters.
int emu_gjk(int shapeA, const float *parA, const float *vertsA, int nA, int shapeB, const float *parB, const float *vertsB,
            int nB, const float *T, const float *v0, float margin, float bound, int want_dist, float *dist, int *iters) {
    pbp::ShapeRef A, B;
    A.shape = shapeA; A.p0 = parA[0]; A.p1 = parA[1]; A.p2 = parA[2];
    A.verts = reinterpret_cast<const float4 *>(vertsA); A.nverts = nA;
    B.shape = shapeB; B.p0 = parB[0]; B.p1 = parB[1]; B.p2 = parB[2];
    B.verts = reinterpret_cast<const float4 *>(vertsB); B.nverts = nB;
    const float3 v = make_float3(v0[0], v0[1], v0[2]);
    pbp::GjkOut r = want_dist ? pbp::gjk_run<true>(A, B, T, v, margin, bound) : pbp::gjk_run<false>(A, B, T, v, margin, bound);
    *dist = r.dist;
    *iters = r.iters;
    return r.hit;
}

// batch over n items sharing the shapes; T [n][12], v0 [n][3]
void emu_gjk_batch(int shapeA, const float *parA, const float *vertsA, int nA, int shapeB, const float *parB,
                   const float *vertsB, int nB, const float *T, const float *v0, int64_t n, float margin, float bound,
                   int want_dist, int32_t *hit, float *dist, int32_t *iters) {
    for (int64_t i = 0; i < n; i++) {
        int it;
        hit[i] = emu_gjk(shapeA, parA, vertsA, nA, shapeB, parB, vertsB, nB, T + 12 * i, v0 + 3 * i, margin, bound, want_dist,
                         dist + i, &it);
        iters[i] = it;
    }
}
}

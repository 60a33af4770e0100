// Kernels of the collision-validation engine (sm_100a).
//
//   k_broadphase   one thread per state: FK of the joint tree with SE(3) products in registers,
//                  world centres of each geometry's enclosing / inscribed sphere, then a
//                  conservative test per active pair: certainly free / certainly colliding /
//                  undecided.  Undecided (state, pair) items are appended -- with the relative
//                  placement of the two geometries -- to that pair's bin (warp-aggregated
//                  atomics, coalesced SoA writes).
//   k_narrowphase  persistent blocks; hull vertices + shape tables staged once per block in
//                  shared memory; each warp pulls 32-item tiles of ONE pair and runs FP32 GJK.
//   k_edge_*       path discretisation on device (counts, explicit fill, sample descriptors).
//   k_fk_output    FK only, writes joint / geometry / frame placements.
#pragma once
#include <limits.h>

#include "pbp_gjk.cuh"

namespace pbp {

constexpr int BP_THREADS = 256;
constexpr int NP_THREADS = 256;
constexpr float BP_SLACK = 2e-5f;  // broadphase decisions keep this margin (metres) for FP32 rounding

// Where the states of a launch come from.
struct StateSource {
    const float *q;          // mode 0: explicit states [S][nq]
    const float *q1, *q2;    // mode 1: edge endpoints [E][nq]
    const int32_t *group;    // per-state group id (edge / path index); NULL -> group = state index
    const float *frac;       // mode 1: interpolation parameter of each state
    const int32_t *sample;   // optional: sample index of each state inside its group (for first_bad)
    int mode;
};

// Per-pair bins of undecided items (SoA; bin p occupies [p*cap, (p+1)*cap)).
struct Bins {
    int32_t *group;      // [npairs*cap]
    int32_t *sample;     // [npairs*cap] (only written when the source carries sample ids)
    float *T;            // [12][npairs*cap]
    uint32_t *count;     // [npairs]
    int64_t cap;
};

struct Outputs {
    uint8_t *hit;          // [groups]
    float *min_dist;       // [states] or NULL
    int32_t *first_pair;   // [states] or NULL   (scratch semantics: INT_MAX = none)
    int32_t *first_bad;    // [groups] or NULL   (scratch semantics: INT_MAX = none)
};

enum { MODE_BOOL = 0, MODE_BOOL_ALL = 1, MODE_DIST = 2 };
// MODE_BOOL      verdict only; a state stops at its first certain hit, dead groups are skipped
// MODE_BOOL_ALL  verdict + first_pair / first_bad: every pair of every state is resolved
// MODE_DIST      verdict + min distance: pairs are resolved unless certainly farther than the bound

__device__ __forceinline__ float point_box_distance(float3 p, float hx, float hy, float hz) {
    const float dx = fmaxf(fabsf(p.x) - hx, 0.f), dy = fmaxf(fabsf(p.y) - hy, 0.f), dz = fmaxf(fabsf(p.z) - hz, 0.f);
    return sqrtf(fmaf(dx, dx, fmaf(dy, dy, dz * dz)));
}

__device__ __forceinline__ void stage_words(void *dst, const void *src, int bytes) {
    const int n = bytes >> 2;
    uint32_t *d = reinterpret_cast<uint32_t *>(dst);
    const uint32_t *s = reinterpret_cast<const uint32_t *>(src);
    for (int i = threadIdx.x; i < n; i += blockDim.x) d[i] = s[i];
}

// Load this block's states into shared memory: sq[local_state * nq + k].
__device__ __forceinline__ void load_states(const StateSource &src, int nq, int64_t block_first, int count, float *sq) {
    if (src.mode == 0) {
        const float *base = src.q + block_first * nq;
        const int total = count * nq;
        if ((reinterpret_cast<uintptr_t>(base) & 15) == 0) {
            const int n4 = total >> 2;
            const float4 *b4 = reinterpret_cast<const float4 *>(base);
            float4 *s4 = reinterpret_cast<float4 *>(sq);
            for (int i = threadIdx.x; i < n4; i += blockDim.x) s4[i] = __ldg(b4 + i);
            for (int i = (n4 << 2) + threadIdx.x; i < total; i += blockDim.x) sq[i] = __ldg(base + i);
        } else {
            for (int i = threadIdx.x; i < total; i += blockDim.x) sq[i] = __ldg(base + i);
        }
    } else {
        if ((int)threadIdx.x < count) {
            const int64_t s = block_first + threadIdx.x;
            const int64_t g = src.group[s];
            const float f = src.frac[s];
            const float *a = src.q1 + g * nq, *b = src.q2 + g * nq;
            for (int k = 0; k < nq; k++) {
                const float qa = __ldg(a + k), qb = __ldg(b + k);
                sq[threadIdx.x * nq + k] = fmaf(f, qb - qa, qa);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// FK + broadphase
// ------------------------------------------------------------------------------------------
template <int MAXJ, int MAXM, int MODE>
__global__ void __launch_bounds__(BP_THREADS)
k_broadphase(DevModel M, StateSource src, int64_t n_states, float padding, Outputs out, Bins bins) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    DevJoint *sj = reinterpret_cast<DevJoint *>(smem_raw);
    DevGeom *sg = reinterpret_cast<DevGeom *>(sj + M.njoints);
    DevPair *sp = reinterpret_cast<DevPair *>(sg + M.ngeoms);
    float *sq = reinterpret_cast<float *>(sp + ((M.npairs + 1) & ~1));

    const int64_t block_first = (int64_t)blockIdx.x * BP_THREADS;
    const int count = (int)min((int64_t)BP_THREADS, n_states - block_first);
    stage_words(sj, M.joints, M.njoints * (int)sizeof(DevJoint));
    stage_words(sg, M.geoms, M.ngeoms * (int)sizeof(DevGeom));
    stage_words(sp, M.pairs, M.npairs * (int)sizeof(DevPair));
    load_states(src, M.nq, block_first, count, sq);
    __syncthreads();

    const int tid = threadIdx.x;
    const int64_t state = block_first + tid;
    const bool valid = tid < count;
    const int32_t group = valid ? (src.group ? src.group[state] : (int32_t)state) : 0;
    const int32_t sample = (valid && src.sample) ? src.sample[state] : 0;
    bool alive = valid;
    if (MODE == MODE_BOOL && src.group && valid && out.hit[group]) alive = false;  // group already decided

    // ---- forward kinematics: oMi kept in local memory for random parent access, the running
    //      transform of a serial chain stays in registers
    float oMi[MAXJ][12];
    float cw[MAXM][3], iw[MAXM][3];
    float cur[12];
#pragma unroll
    for (int i = 0; i < 12; i++) { cur[i] = (i == 0 || i == 4 || i == 8) ? 1.f : 0.f; oMi[0][i] = cur[i]; }
    const float *myq = sq + tid * M.nq;
    if (__any_sync(0xffffffffu, alive)) {
        for (int j = 1; j < M.njoints; j++) {
            const DevJoint &J = sj[j];
            float par[12];
            if (J.chain) {
#pragma unroll
                for (int i = 0; i < 12; i++) par[i] = cur[i];
            } else {
#pragma unroll
                for (int i = 0; i < 12; i++) par[i] = oMi[J.parent][i];
            }
            const float qj = valid ? myq[J.idx_q] : 0.f;
            joint_transform(J, par, qj, cur);
#pragma unroll
            for (int i = 0; i < 12; i++) oMi[j][i] = cur[i];
        }
        // world centres of the moving geometries' spheres
        for (int g = 0; g < M.ngeoms; g++) {
            const DevGeom &G = sg[g];
            if (G.slot < 0) continue;
            float T[12];
            if (G.identity) {
#pragma unroll
                for (int i = 0; i < 12; i++) T[i] = oMi[G.parent][i];
            } else {
                float P[12];
#pragma unroll
                for (int i = 0; i < 12; i++) P[i] = oMi[G.parent][i];
                se3_mul(P, G.P, T);
            }
            const float3 c = se3_act(T, G.outer[0], G.outer[1], G.outer[2]);
            const float3 ic = se3_act(T, G.inner[0], G.inner[1], G.inner[2]);
            cw[G.slot][0] = c.x; cw[G.slot][1] = c.y; cw[G.slot][2] = c.z;
            iw[G.slot][0] = ic.x; iw[G.slot][1] = ic.y; iw[G.slot][2] = ic.z;
        }
    }

    auto outer_centre = [&](const DevGeom &G) {
        return G.slot < 0 ? make_float3(G.outer[0], G.outer[1], G.outer[2]) : make_float3(cw[G.slot][0], cw[G.slot][1], cw[G.slot][2]);
    };
    auto inner_centre = [&](const DevGeom &G) {
        return G.slot < 0 ? make_float3(G.inner[0], G.inner[1], G.inner[2]) : make_float3(iw[G.slot][0], iw[G.slot][1], iw[G.slot][2]);
    };
    auto world_placement = [&](const DevGeom &G, float *T) {
        if (G.slot < 0) {
#pragma unroll
            for (int i = 0; i < 12; i++) T[i] = G.P[i];
        } else if (G.identity) {
#pragma unroll
            for (int i = 0; i < 12; i++) T[i] = oMi[G.parent][i];
        } else {
            float P[12];
#pragma unroll
            for (int i = 0; i < 12; i++) P[i] = oMi[G.parent][i];
            se3_mul(P, G.P, T);
        }
    };
    // lower / upper bound of the distance of one pair from the sphere (and box) proxies
    auto pair_bounds = [&](const DevPair &pr, const DevGeom &GA, const DevGeom &GB, float &lb, float &ub) {
        if (pr.kind == BP_BOX_A || pr.kind == BP_BOX_B) {
            const DevGeom &BX = pr.kind == BP_BOX_A ? GA : GB;
            const DevGeom &OT = pr.kind == BP_BOX_A ? GB : GA;
            float T[12];
            world_placement(BX, T);
            const float3 pc = se3_act_inv(T, outer_centre(OT));
            const float3 pi = se3_act_inv(T, inner_centre(OT));
            lb = point_box_distance(pc, BX.par[0], BX.par[1], BX.par[2]) - OT.outer[3];
            ub = point_box_distance(pi, BX.par[0], BX.par[1], BX.par[2]) - OT.inner[3];
        } else {
            const float3 d = sub3(outer_centre(GA), outer_centre(GB));
            const float3 e = sub3(inner_centre(GA), inner_centre(GB));
            lb = sqrtf(dot3(d, d)) - GA.outer[3] - GB.outer[3];
            ub = sqrtf(dot3(e, e)) - GA.inner[3] - GB.inner[3];
        }
    };

    float bound = INFINITY;     // MODE_DIST: certain upper bound of this state's min distance
    int first_pair = INT_MAX;   // MODE_BOOL_ALL
    bool state_hit = false;

    if (MODE == MODE_DIST) {
        // pass 1: the smallest certain upper bound over all pairs
        for (int p = 0; p < M.npairs; p++) {
            const DevPair pr = sp[p];
            float lb, ub;
            pair_bounds(pr, sg[pr.a], sg[pr.b], lb, ub);
            const float slack = pr.kind == BP_EXACT_SPHERES ? 0.f : BP_SLACK;
            bound = fminf(bound, fmaxf(ub + slack, 0.f));
        }
    }

    const unsigned lane = tid & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    for (int p = 0; p < M.npairs; p++) {
        if (MODE == MODE_BOOL && !__any_sync(0xffffffffu, alive)) break;
        const DevPair pr = sp[p];
        const DevGeom &GA = sg[pr.a];
        const DevGeom &GB = sg[pr.b];
        bool undecided = false;
        if (alive) {
            float lb, ub;
            pair_bounds(pr, GA, GB, lb, ub);
            const bool exact = pr.kind == BP_EXACT_SPHERES;
            const float slack = exact ? 0.f : BP_SLACK;
            const bool sure = ub <= padding - slack;
            if (sure) {
                state_hit = true;
                if (MODE == MODE_BOOL) alive = false;
                if (MODE == MODE_BOOL_ALL) first_pair = min(first_pair, p);
            }
            if (MODE == MODE_DIST) {
                if (exact) bound = fminf(bound, fmaxf(lb, 0.f));  // lb == ub == exact distance
                undecided = !exact && (lb - slack <= bound);
            } else {
                const bool free_ = exact ? !sure : (lb > padding + slack);
                undecided = !sure && !free_;
            }
        }
        const unsigned um = __ballot_sync(0xffffffffu, undecided);
        if (um) {
            uint32_t base = 0;
            if (lane == (unsigned)(__ffs(um) - 1)) base = atomicAdd(bins.count + p, (uint32_t)__popc(um));
            base = __shfl_sync(0xffffffffu, base, __ffs(um) - 1);
            if (undecided) {
                float TA[12], TB[12], R[12];
                world_placement(GA, TA);
                world_placement(GB, TB);
                se3_inv_mul(TA, TB, R);
                const int64_t slot = (int64_t)p * bins.cap + base + __popc(um & lt_mask);
                const int64_t stride = (int64_t)M.npairs * bins.cap;
                bins.group[slot] = group;
                if (src.sample) bins.sample[slot] = sample;
#pragma unroll
                for (int k = 0; k < 12; k++) bins.T[k * stride + slot] = R[k];
            }
        }
    }

    if (valid) {
        if (MODE == MODE_BOOL) {
            if (state_hit) out.hit[group] = 1;
        } else if (MODE == MODE_BOOL_ALL) {
            if (state_hit) {
                out.hit[group] = 1;
                if (out.first_pair) atomicMin(out.first_pair + state, first_pair);
                if (out.first_bad) atomicMin(out.first_bad + group, sample);
            }
        } else {
            if (state_hit) out.hit[group] = 1;
            // bound is a certain upper bound (+slack) of the true minimum; items refine it downwards
            out.min_dist[state] = bound;
        }
    }
}

// ------------------------------------------------------------------------------------------
// Narrowphase (GJK), persistent, pair-major tiles
// ------------------------------------------------------------------------------------------
template <int MODE, bool VERTS_IN_SMEM>
__global__ void __launch_bounds__(NP_THREADS)
k_narrowphase(DevModel M, float padding, Outputs out, Bins bins, uint32_t *tile_counter, bool has_sample) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4 *sv = reinterpret_cast<float4 *>(smem_raw);
    DevGeom *sg = reinterpret_cast<DevGeom *>(sv + (VERTS_IN_SMEM ? M.nverts_padded : 0));
    DevPair *sp = reinterpret_cast<DevPair *>(sg + M.ngeoms);
    uint32_t *prefix = reinterpret_cast<uint32_t *>(sp + ((M.npairs + 1) & ~1));  // [npairs + 1] tiles before bin k
    int32_t *order = reinterpret_cast<int32_t *>(prefix + M.npairs + 1);

    if (VERTS_IN_SMEM) stage_words(sv, M.verts, M.nverts_padded * (int)sizeof(float4));
    stage_words(sg, M.geoms, M.ngeoms * (int)sizeof(DevGeom));
    stage_words(sp, M.pairs, M.npairs * (int)sizeof(DevPair));
    stage_words(order, M.bin_order, M.npairs * (int)sizeof(int32_t));
    __syncthreads();
    if (threadIdx.x < 32) {
        // warp scan of tiles per bin, in bin order
        uint32_t run = 0;
        for (int k0 = 0; k0 < M.npairs; k0 += 32) {
            const int k = k0 + threadIdx.x;
            uint32_t t = k < M.npairs ? (bins.count[order[k]] + 31u) >> 5 : 0u;
            uint32_t inc = t;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, inc, o);
                if ((int)threadIdx.x >= o) inc += y;
            }
            if (k < M.npairs) prefix[k] = run + inc - t;
            run += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (threadIdx.x == 0) prefix[M.npairs] = run;
    }
    __syncthreads();
    const uint32_t total_tiles = prefix[M.npairs];
    const unsigned lane = threadIdx.x & 31;
    const float4 *verts = VERTS_IN_SMEM ? sv : M.verts;
    const int64_t stride = (int64_t)M.npairs * bins.cap;

    for (;;) {
        uint32_t tile = 0;
        if (lane == 0) tile = atomicAdd(tile_counter, 1u);
        tile = __shfl_sync(0xffffffffu, tile, 0);
        if (tile >= total_tiles) break;
        // bin = last k with prefix[k] <= tile
        int lo = 0, hi = M.npairs;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (prefix[mid] <= tile) lo = mid; else hi = mid;
        }
        const int p = order[lo];
        const uint32_t idx = ((tile - prefix[lo]) << 5) + lane;
        if (idx >= bins.count[p]) continue;
        const int64_t slot = (int64_t)p * bins.cap + idx;
        const int32_t group = bins.group[slot];
        if (MODE == MODE_BOOL && out.hit[group]) continue;  // another item / the broadphase already decided it
        float T[12];
#pragma unroll
        for (int k = 0; k < 12; k++) T[k] = bins.T[k * stride + slot];

        const DevPair pr = sp[p];
        const DevGeom &GA = sg[pr.a];
        const DevGeom &GB = sg[pr.b];
        ShapeRef A, B;
        A.shape = GA.shape; A.p0 = GA.par[0]; A.p1 = GA.par[1]; A.p2 = GA.par[2];
        A.verts = verts + GA.vert_off; A.nverts = GA.vert_cnt;
        B.shape = GB.shape; B.p0 = GB.par[0]; B.p1 = GB.par[1]; B.p2 = GB.par[2];
        B.verts = verts + GB.vert_off; B.nverts = GB.vert_cnt;
        // initial direction: centre(A) - centre(B) in A's frame (geometry-frame sphere centres)
        const DevGeom *ga_glob = M.geoms + pr.a, *gb_glob = M.geoms + pr.b;
        (void)ga_glob; (void)gb_glob;
        const float radii = GA.core_radius + GB.core_radius;
        float3 ca, cb;
        {
            // moving geometries keep geometry-frame centres in `outer`; static ones hold world
            // centres there, so recover the geometry-frame centre through the placement
            if (GA.slot < 0) ca = se3_act_inv(GA.P, make_float3(GA.outer[0], GA.outer[1], GA.outer[2]));
            else ca = make_float3(GA.outer[0], GA.outer[1], GA.outer[2]);
            float3 cbl;
            if (GB.slot < 0) cbl = se3_act_inv(GB.P, make_float3(GB.outer[0], GB.outer[1], GB.outer[2]));
            else cbl = make_float3(GB.outer[0], GB.outer[1], GB.outer[2]);
            cb = se3_act(T, cbl.x, cbl.y, cbl.z);
        }
        const float3 v0 = sub3(ca, cb);

        if (MODE == MODE_DIST) {
            const float cur = out.min_dist[group];  // group == state in this mode
            const GjkOut r = gjk_run<true>(A, B, T, v0, padding + radii, cur + radii);
            const float d = fmaxf(r.dist - radii, 0.f);
            if (r.hit) out.hit[group] = 1;
            atomicMin(reinterpret_cast<unsigned int *>(out.min_dist + group), __float_as_uint(d));
        } else {
            const GjkOut r = gjk_run<false>(A, B, T, v0, padding + radii, 0.f);
            if (r.hit) {
                if (MODE == MODE_BOOL) {
                    out.hit[group] = 1;
                } else {
                    // MODE_BOOL_ALL: `group` holds the state index unless the source has groups
                    if (out.first_pair) { atomicMin(out.first_pair + group, p); out.hit[group] = 1; }
                    if (out.first_bad) { atomicMin(out.first_bad + group, has_sample ? bins.sample[slot] : 0); out.hit[group] = 1; }
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// Path discretisation (reference planning/utils.py:114-140)
// ------------------------------------------------------------------------------------------
// counts[e] = ceil(||q2 - q1||_2 / step) + 1, evaluated in FP64 so the sample count matches the
// reference's float64 arithmetic on the same (float32-representable) endpoints.
__global__ void k_edge_counts(const float *q1, const float *q2, int64_t n_edges, int nq, double step, int32_t *counts) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_edges) return;
    double s = 0.0;
    for (int k = 0; k < nq; k++) {
        const double d = (double)q2[e * nq + k] - (double)q1[e * nq + k];
        s += d * d;
    }
    counts[e] = (int32_t)ceil(sqrt(s) / step) + 1;
}

// q_out[offsets[e] + i] = q1 + linspace(0,1,n)[i] * (q2 - q1); one warp per edge, lanes stride
// over (sample, coordinate) so the stores are contiguous.
__global__ void k_edge_fill(const float *q1, const float *q2, int64_t n_edges, int nq, const int32_t *counts,
                            const int64_t *offsets, float *q_out) {
    const int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (e >= n_edges) return;
    const int lane = threadIdx.x & 31;
    const int n = counts[e];
    const float inv = n > 1 ? 1.0f / (float)(n - 1) : 0.f;
    const int64_t total = (int64_t)n * nq;
    float *dst = q_out + offsets[e] * nq;
    for (int64_t t = lane; t < total; t += 32) {
        const int i = (int)(t / nq), k = (int)(t - (int64_t)i * nq);
        const float f = (i == n - 1 && n > 1) ? 1.0f : (float)i * inv;
        const float a = q1[e * nq + k], b = q2[e * nq + k];
        dst[t] = fmaf(f, b - a, a);
    }
}

// Sample descriptors for one refinement level of every live edge.
//
// The reference walks an edge's samples in order and stops at the first colliding one
// (core/utils.py:125-130).  On device the samples of all edges are tested level by level,
// coarse to fine -- level 0 = samples {0, S, 2S, ...} plus the end point, level l = the
// multiples of S/2^l not yet visited -- and edges already known to collide are dropped before
// the next level is emitted.  The verdict (any sample collides) is order independent.
struct LevelSpec {
    int stride;     // sample indices 0, stride, 2*stride, ... < n
    int skip_mod;   // skip indices divisible by this (visited at a coarser level); 0 = none
    int level0;     // 1: also emit the end point n-1; 0: never emit it (level 0 did)
};

template <typename F>
__device__ __forceinline__ int for_level_samples(int n, LevelSpec spec, F emit) {
    int c = 0;
    for (int i = 0; i < n; i += spec.stride) {
        if (spec.skip_mod && (i % spec.skip_mod) == 0) continue;
        if (!spec.level0 && i == n - 1) continue;
        emit(i);
        c++;
    }
    if (spec.level0 && ((n - 1) % spec.stride) != 0) { emit(n - 1); c++; }
    return c;
}

// Pass 1: total number of descriptors of this level (warp-reduced atomic).
__global__ void k_level_count(const int32_t *counts, const uint8_t *dead, int64_t n_edges, LevelSpec spec,
                              unsigned long long *total) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int c = 0;
    if (e < n_edges && !(dead && dead[e])) c = for_level_samples(counts[e], spec, [](int) {});
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(total, (unsigned long long)c);
}

// Pass 2: write (edge, fraction, sample index) descriptors; slots from a warp-aggregated cursor.
__global__ void k_level_emit(const int32_t *counts, const uint8_t *dead, int64_t n_edges, LevelSpec spec,
                             unsigned long long *cursor, int32_t *group, float *frac, int32_t *sample) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = e < n_edges && !(dead && dead[e]);
    const int n = live ? counts[e] : 0;
    const int c = live ? for_level_samples(n, spec, [](int) {}) : 0;
    // exclusive warp scan of c
    int inc = c;
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += y;
    }
    const int warp_total = __shfl_sync(0xffffffffu, inc, 31);
    unsigned long long base = 0;
    if (lane == 0 && warp_total) base = atomicAdd(cursor, (unsigned long long)warp_total);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (!c) return;
    int64_t o = (int64_t)base + inc - c;
    const float inv = n > 1 ? 1.0f / (float)(n - 1) : 0.f;
    for_level_samples(n, spec, [&](int i) {
        group[o] = (int32_t)e;
        frac[o] = (i == n - 1 && n > 1) ? 1.0f : (float)i * inv;
        if (sample) sample[o] = i;
        o++;
    });
}

// group ids of a ragged batch of pre-discretised paths: group[s] = p for offsets[p] <= s < offsets[p+1]
__global__ void k_path_groups(const int64_t *offsets, int64_t n_paths, int64_t total, int32_t *group, int32_t *sample) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= total) return;
    int64_t lo = 0, hi = n_paths;
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (offsets[mid] <= s) lo = mid; else hi = mid;
    }
    group[s] = (int32_t)lo;
    if (sample) sample[s] = (int32_t)(s - offsets[lo]);
}

// ------------------------------------------------------------------------------------------
// FK output
// ------------------------------------------------------------------------------------------
template <int MAXJ>
__global__ void __launch_bounds__(BP_THREADS)
k_fk_output(DevModel M, const float *q, int64_t n_states, float *oMi_out, float *oMg_out, float *oMf_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    DevJoint *sj = reinterpret_cast<DevJoint *>(smem_raw);
    DevGeom *sg = reinterpret_cast<DevGeom *>(sj + M.njoints);
    stage_words(sj, M.joints, M.njoints * (int)sizeof(DevJoint));
    stage_words(sg, M.geoms, M.ngeoms * (int)sizeof(DevGeom));
    __syncthreads();
    const int64_t s = (int64_t)blockIdx.x * BP_THREADS + threadIdx.x;
    if (s >= n_states) return;
    float oMi[MAXJ][12];
#pragma unroll
    for (int i = 0; i < 12; i++) oMi[0][i] = (i == 0 || i == 4 || i == 8) ? 1.f : 0.f;
    for (int j = 1; j < M.njoints; j++) {
        const DevJoint &J = sj[j];
        float par[12], cur[12];
#pragma unroll
        for (int i = 0; i < 12; i++) par[i] = oMi[J.parent][i];
        joint_transform(J, par, q[s * M.nq + J.idx_q], cur);
#pragma unroll
        for (int i = 0; i < 12; i++) oMi[j][i] = cur[i];
    }
    if (oMi_out)
        for (int j = 0; j < M.njoints; j++)
            for (int i = 0; i < 12; i++) oMi_out[(s * M.njoints + j) * 12 + i] = oMi[j][i];
    if (oMg_out)
        for (int g = 0; g < M.ngeoms; g++) {
            const DevGeom &G = sg[g];
            float P[12], T[12];
#pragma unroll
            for (int i = 0; i < 12; i++) P[i] = oMi[G.parent][i];
            se3_mul(P, G.P, T);
            for (int i = 0; i < 12; i++) oMg_out[(s * M.ngeoms + g) * 12 + i] = T[i];
        }
    if (oMf_out)
        for (int f = 0; f < M.nframes; f++) {
            float P[12], F[12], T[12];
            const int pj = M.frame_parent[f];
#pragma unroll
            for (int i = 0; i < 12; i++) { P[i] = oMi[pj][i]; F[i] = M.frame_placement[f * 12 + i]; }
            se3_mul(P, F, T);
            for (int i = 0; i < 12; i++) oMf_out[(s * M.nframes + f) * 12 + i] = T[i];
        }
}

// ------------------------------------------------------------------------------------------
// small utility kernels
// ------------------------------------------------------------------------------------------
__global__ void k_fill_i32(int32_t *p, int64_t n, int32_t v) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
__global__ void k_fill_f32(float *p, int64_t n, float v) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
// total += sum(count[0..n))   (profiling: items that reached the narrowphase)
__global__ void k_sum_counts(const uint32_t *count, int n, unsigned long long *total) {
    unsigned long long s = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += count[i];
    if (s) atomicAdd(total, s);
}
// INT_MAX (none) -> -1
__global__ void k_finalize_index(int32_t *p, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && p[i] == INT_MAX) p[i] = -1;
}

}  // namespace pbp

// Kernels of the collision-validation engine (sm_100a, scalar FP32).
//
//   k_broadphase   one thread per state.  FK of the joint tree with the running SE(3) product in
//                  registers; only the world centre of each moving geometry's proxy sphere (3
//                  floats) is kept, thread-interleaved in shared memory.  Every active pair is then
//                  classified from squared distances (no sqrt): certainly free / certainly
//                  colliding / undecided.  Undecided (state, pair) items are appended to the
//                  pair's bin (one warp-aggregated atomic per pair, coalesced stores).
//   k_item_setup   one thread per item, pair-major tiles: recomputes the FK chain of just the two
//                  geometries of the item's pair (uniform control flow inside a warp, everything in
//                  registers) and stores the relative placement T = oMg[a]^-1 oMg[b].
//   k_narrowphase  persistent blocks, hull vertices staged once per block in shared memory; each
//                  warp streams the items of ONE pair through FP32 GJK, one iteration per trip for
//                  all active lanes, refilling finished lanes from the pair's bin.
//   k_edge_* / k_level_*   path discretisation on device (counts, explicit fill, level descriptors).
//   k_fk_output    FK only, writes joint / geometry / frame placements.
#pragma once
#include <limits.h>

#include "pbp_gjk.cuh"

namespace pbp {

#ifndef PBP_IS_MIN_BLOCKS
#define PBP_IS_MIN_BLOCKS 3   // minimum resident blocks the item setup is compiled for (80 registers; measured 121 -> 102 us)
#endif
#ifndef PBP_NP_REFILL_MIN
#define PBP_NP_REFILL_MIN 8   // measured on config 2: 173.7 us (1) / 172.9 (4) / 170.0 (8) / 170.7 (12) / 171.0 (16)
#endif
#ifndef PBP_NP_LOCKSTEP
#define PBP_NP_LOCKSTEP 1
#endif
#ifndef PBP_NP_MIN_BLOCKS
#define PBP_NP_MIN_BLOCKS 2
#endif
constexpr int BP_THREADS = 128;
constexpr int NP_THREADS = 256;
constexpr int IS_THREADS = 256;
#ifndef PBP_TILE_ITEMS
#define PBP_TILE_ITEMS 64    // measured: 128 -> 99.7 us, 64 -> 98.5 us (1 Mi states), 41.4 -> 38.6 us (256 Ki); 32 and 256 are worse
#endif
constexpr int TILE_ITEMS = PBP_TILE_ITEMS;    // items of one pair handed to a warp per tile fetch
constexpr int BP_MAX_LOCAL_CENTRES = PBP_MAX_GEOMS;
constexpr int BP_MAX_SAVES = 16;
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
    int32_t *state;      // [npairs*cap] written by the broadphase: state index inside the round
    int32_t *group;      // [npairs*cap] written by the item setup: output group, -1 = dropped
    int32_t *sample;     // [npairs*cap] (only when the source carries sample ids)
    float *T;            // [12][npairs*cap] relative placement of b in a's frame
    uint32_t *count;     // [npairs] items appended by the broadphase
    uint32_t *count_np;  // [npairs] items the item setup kept for the narrowphase (compacted to the bin's front)
    int64_t cap;
};

struct Outputs {
    uint8_t *hit;          // [groups]
    float *min_dist;       // [states] or NULL
    int32_t *first_pair;   // [states] or NULL   (scratch semantics: INT_MAX = none)
    int32_t *first_bad;    // [groups] or NULL   (scratch semantics: INT_MAX = none)
};

enum { MODE_BOOL = 0, MODE_BOOL_ALL = 1, MODE_DIST = 2 };
// bit set in Bins::group by the narrowphase on a hull hit that k_surface_refine has to decide
constexpr int32_t SR_PARKED = 0x40000000;
// MODE_BOOL      verdict only; a state stops at its first certain hit, decided groups are skipped
// MODE_BOOL_ALL  verdict + first_pair / first_bad: every pair of every state is resolved
// MODE_DIST      verdict + min distance: pairs are resolved unless certainly farther than the bound

// Copy a model table into shared memory.  Both pointers are 16-byte aligned (cudaMalloc bases, smem
// carved with align16): 16-byte loads, four in flight per thread -- a word-by-word loop is a chain
// of dependent L2 latencies (60 of them for the 120 KB k_surface_refine stages, 40 % of its time).
__device__ __forceinline__ void stage_words(void *dst, const void *src, int bytes) {
    const int n16 = bytes >> 4;
    uint4 *d4 = reinterpret_cast<uint4 *>(dst);
    const uint4 *s4 = reinterpret_cast<const uint4 *>(src);
#pragma unroll 4
    for (int i = threadIdx.x; i < n16; i += blockDim.x) d4[i] = __ldg(s4 + i);
    const int n = bytes >> 2;
    uint32_t *d = reinterpret_cast<uint32_t *>(dst);
    const uint32_t *s = reinterpret_cast<const uint32_t *>(src);
    for (int i = (n16 << 2) + threadIdx.x; i < n; i += blockDim.x) d[i] = s[i];
}

__host__ __device__ __forceinline__ size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

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

// Shared-memory footprint of the broadphase tables (everything but the per-thread store).
__host__ __device__ inline size_t bp_table_bytes(int njoints, int nmoving, int ngeoms, int nboxes, int npairs, int ngroups, int nq) {
    const size_t np1 = npairs > 0 ? npairs : 1;
    return align16((size_t)njoints * sizeof(DevJoint)) + align16((size_t)(nmoving > 0 ? nmoving : 1) * 4) +
           align16((size_t)(ngeoms > 0 ? ngeoms : 1) * sizeof(BpGeom)) + align16((size_t)(nboxes > 0 ? nboxes : 1) * sizeof(BpBox)) +
           align16(np1 * sizeof(BpPair)) + align16((size_t)(ngroups > 0 ? ngroups : 1) * sizeof(BpGroup)) + align16(np1 * sizeof(float2)) +
           align16((size_t)BP_THREADS * nq * 4) + align16((size_t)((npairs + 31) >> 5) * BP_THREADS * 4 + 16) +
           align16((size_t)(BP_THREADS / 32) * np1 * 4);
}

// ------------------------------------------------------------------------------------------
// FK + broadphase
// ------------------------------------------------------------------------------------------
// SMEM_STORE: per-thread proxy centres / saved branch transforms live thread-interleaved in shared
// memory (no local memory at all); otherwise (very large models) in local arrays.
template <int MODE, bool SMEM_STORE>
__global__ void __launch_bounds__(BP_THREADS)
k_broadphase(DevModel M, StateSource src, int64_t n_states, float padding, Outputs out, Bins bins) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned char *ptr = smem_raw;
    DevJoint *sj = reinterpret_cast<DevJoint *>(ptr);   ptr += align16((size_t)M.njoints * sizeof(DevJoint));
    int32_t *sjg = reinterpret_cast<int32_t *>(ptr);     ptr += align16((size_t)max(M.nmoving, 1) * 4);
    BpGeom *sg = reinterpret_cast<BpGeom *>(ptr);        ptr += align16((size_t)max(M.ngeoms, 1) * sizeof(BpGeom));
    BpBox *sb = reinterpret_cast<BpBox *>(ptr);          ptr += align16((size_t)max(M.nboxes, 1) * sizeof(BpBox));
    BpPair *sp = reinterpret_cast<BpPair *>(ptr);        ptr += align16((size_t)max(M.npairs, 1) * sizeof(BpPair));
    BpGroup *sgrp = reinterpret_cast<BpGroup *>(ptr);    ptr += align16((size_t)max(M.ngroups, 1) * sizeof(BpGroup));
    float2 *sthr = reinterpret_cast<float2 *>(ptr);      ptr += align16((size_t)max(M.npairs, 1) * sizeof(float2));
    float *sq = reinterpret_cast<float *>(ptr);          ptr += align16((size_t)BP_THREADS * M.nq * 4);
    uint32_t *sbits = reinterpret_cast<uint32_t *>(ptr); ptr += align16((size_t)((M.npairs + 31) >> 5) * BP_THREADS * 4 + 16);
    uint32_t *swcount = reinterpret_cast<uint32_t *>(ptr); ptr += align16((size_t)(BP_THREADS / 32) * max(M.npairs, 1) * 4);
    float *sstore = reinterpret_cast<float *>(ptr);      // [(nmoving*3 + nsave*12)][BP_THREADS]

    const int64_t block_first = (int64_t)blockIdx.x * BP_THREADS;
    const int count = (int)min((int64_t)BP_THREADS, n_states - block_first);
    stage_words(sj, M.joints, M.njoints * (int)sizeof(DevJoint));
    stage_words(sjg, M.joint_geoms, M.nmoving * 4);
    stage_words(sg, M.bpgeoms, M.ngeoms * (int)sizeof(BpGeom));
    stage_words(sb, M.bpboxes, M.nboxes * (int)sizeof(BpBox));
    stage_words(sp, M.bppairs, M.npairs * (int)sizeof(BpPair));
    stage_words(sgrp, M.groups, M.ngroups * (int)sizeof(BpGroup));
    load_states(src, M.nq, block_first, count, sq);
    for (int i = threadIdx.x; i < (BP_THREADS / 32) * M.npairs; i += BP_THREADS) swcount[i] = 0;
    // squared decision thresholds of every pair for this call's padding:
    //   d2 > x -> certainly free,  d2 <= y -> certainly colliding  (y < 0: never)
    for (int p = threadIdx.x; p < M.npairs; p += BP_THREADS) {
        const BpPair pr = M.bppairs[p];
        const bool exact = pr.kind == BP_EXACT_SPHERES;
        const float slack = exact ? 0.f : BP_SLACK;
        const float rf = pr.r_out + padding + slack, rs = pr.r_in + padding - slack;
        const float y = rs >= 0.f ? rs * rs : -1.f;
        sthr[p] = make_float2(exact ? y : rf * rf, y);
    }
    __syncthreads();

    const int tid = threadIdx.x;
    const int64_t state = block_first + tid;
    const bool valid = tid < count;
    const int32_t group = valid ? (src.group ? src.group[state] : (int32_t)state) : 0;
    bool alive = valid;
    if (MODE == MODE_BOOL && src.group && valid && out.hit[group]) alive = false;  // group already decided

    float lstore[SMEM_STORE ? 1 : (BP_MAX_LOCAL_CENTRES * 3 + BP_MAX_SAVES * 12)];
    auto store = [&](int idx) -> float & {
        if (SMEM_STORE) return sstore[idx * BP_THREADS + tid];
        return lstore[SMEM_STORE ? 0 : idx];
    };
    const int save_base = M.nmoving * 3;

    // ---- forward kinematics: running product in registers
    if (__any_sync(0xffffffffu, alive)) {
        float cur[12];
        se3_identity(cur);
        const float *myq = sq + tid * M.nq;
        for (int j = 1; j < M.njoints; j++) {
            const DevJoint &J = sj[j];
            float par[12];
            if (J.parent_slot == -1) {
#pragma unroll
                for (int i = 0; i < 12; i++) par[i] = cur[i];
            } else if (J.parent_slot == -2) {
                se3_identity(par);
            } else {
#pragma unroll
                for (int i = 0; i < 12; i++) par[i] = store(save_base + J.parent_slot * 12 + i);
            }
            const float qj = valid ? myq[J.idx_q] : 0.f;
            joint_transform(J, par, qj, cur);
            if (J.save_slot >= 0) {
#pragma unroll
                for (int i = 0; i < 12; i++) store(save_base + J.save_slot * 12 + i) = cur[i];
            }
            for (int gi = J.geom_begin; gi < J.geom_end; gi++) {
                const BpGeom &G = sg[sjg[gi]];
                const float3 c = se3_act(cur, G.c[0], G.c[1], G.c[2]);
                store(G.slot * 3 + 0) = c.x;
                store(G.slot * 3 + 1) = c.y;
                store(G.slot * 3 + 2) = c.z;
            }
        }
    }

    auto moving_centre = [&](int slot) { return make_float3(store(slot * 3 + 0), store(slot * 3 + 1), store(slot * 3 + 2)); };
    auto any_centre = [&](int g, int slot) {
        return slot >= 0 ? moving_centre(slot) : make_float3(sg[g].c[0], sg[g].c[1], sg[g].c[2]);
    };
    auto box_dist2 = [&](const BpBox &bx, float3 c) {
        const float dx = c.x - bx.t[0], dy = c.y - bx.t[1], dz = c.z - bx.t[2];
        const float lx = fmaf(bx.R[0], dx, fmaf(bx.R[3], dy, bx.R[6] * dz));
        const float ly = fmaf(bx.R[1], dx, fmaf(bx.R[4], dy, bx.R[7] * dz));
        const float lz = fmaf(bx.R[2], dx, fmaf(bx.R[5], dy, bx.R[8] * dz));
        const float ex = fmaxf(fabsf(lx) - bx.h[0], 0.f), ey = fmaxf(fabsf(ly) - bx.h[1], 0.f), ez = fmaxf(fabsf(lz) - bx.h[2], 0.f);
        return fmaf(ex, ex, fmaf(ey, ey, ez * ez));
    };
    // proxy distance of any pair, without the per-group hoisting (MODE_DIST pass 1, generic groups)
    auto generic_dist2 = [&](const BpPair &pr) {
        if (pr.kind == BP_BOX_A) return box_dist2(sb[sg[pr.a].box], any_centre(pr.b, pr.slot_b));
        if (pr.kind == BP_BOX_B) return box_dist2(sb[sg[pr.b].box], any_centre(pr.a, pr.slot_a));
        const float3 d = sub3(any_centre(pr.a, pr.slot_a), any_centre(pr.b, pr.slot_b));
        return dot3(d, d);
    };

    float bound = INFINITY;     // MODE_DIST: certain upper bound of this state's min distance
    int first_pair = INT_MAX;   // MODE_BOOL_ALL (caller's pair numbering)
    bool state_hit = false;

    if (MODE == MODE_DIST && alive) {
        // pass 0: the smallest certain upper bound over all pairs
        for (int p = 0; p < M.npairs; p++) {
            const BpPair pr = sp[p];
            const float slack = pr.kind == BP_EXACT_SPHERES ? 0.f : BP_SLACK;
            bound = fminf(bound, fmaxf(sqrtf(generic_dist2(pr)) - pr.r_in + slack, 0.f));
        }
    }

    // ---- pass 1: classify every pair, group by group; undecided pairs become bits of per-thread
    //      words kept in shared memory (no atomics on the critical path)
    const int nwords = (M.npairs + 31) >> 5;
    uint32_t bits = 0;
    auto classify = [&](int p, float d2) {
        bool undecided = false;
        if (alive) {
            const float2 thr = sthr[p];
            const bool sure = d2 <= thr.y;
            if (sure) {
                state_hit = true;
                if (MODE == MODE_BOOL) alive = false;
                if (MODE == MODE_BOOL_ALL) first_pair = min(first_pair, M.pair_orig[p]);
            }
            if (MODE == MODE_DIST) {
                const BpPair pr = sp[p];
                const bool exact = pr.kind == BP_EXACT_SPHERES;
                const float lb = sqrtf(d2) - pr.r_out;
                if (exact) bound = fminf(bound, fmaxf(lb, 0.f));  // the proxies are the shapes
                undecided = !exact && (lb - BP_SLACK <= bound);
            } else {
                undecided = !sure && !(d2 > thr.x);
            }
        }
        bits |= (undecided ? 1u : 0u) << (p & 31);
        if ((p & 31) == 31 || p == M.npairs - 1) {   // word complete (uniform branch)
            sbits[(p >> 5) * BP_THREADS + tid] = bits;
            bits = 0;
        }
    };
    for (int gi = 0; gi < M.ngroups; gi++) {
        const BpGroup grp = sgrp[gi];
        if (MODE == MODE_BOOL && !__any_sync(0xffffffffu, alive)) {
            // nobody needs the remaining pairs: their words are zero
            for (int p = grp.begin; p < grp.end; p++) classify(p, 0.f);
            continue;
        }
        if (grp.cls == GRP_MOVING) {
            for (int p = grp.begin; p < grp.end; p++) {
                const BpPair pr = sp[p];
                const float3 d = sub3(moving_centre(pr.slot_a), moving_centre(pr.slot_b));
                classify(p, dot3(d, d));
            }
        } else if (grp.cls == GRP_STATIC_POINT) {
            const float3 c0 = make_float3(sg[grp.sidx].c[0], sg[grp.sidx].c[1], sg[grp.sidx].c[2]);
            for (int p = grp.begin; p < grp.end; p++) {
                const BpPair pr = sp[p];
                const float3 d = sub3(moving_centre(pr.slot_a >= 0 ? pr.slot_a : pr.slot_b), c0);
                classify(p, dot3(d, d));
            }
        } else if (grp.cls == GRP_STATIC_BOX) {
            const BpBox &bx = sb[grp.sidx];
            for (int p = grp.begin; p < grp.end; p++) {
                const BpPair pr = sp[p];
                classify(p, box_dist2(bx, moving_centre(pr.slot_a >= 0 ? pr.slot_a : pr.slot_b)));
            }
        } else {
            for (int p = grp.begin; p < grp.end; p++) classify(p, generic_dist2(sp[p]));
        }
    }
    // ---- append the undecided items to their pairs' bins: every thread walks its OWN bits twice
    //      with shared-memory atomics -- count per pair, one global reservation per (block, pair),
    //      then rank + store -- so a thread pays only for its ~1.3 items
    uint32_t *scnt = swcount, *sbase = swcount + M.npairs;
    if (MODE == MODE_BOOL && state_hit) {
        // a state decided "colliding" needs none of its items
        for (int w = 0; w < nwords; w++) sbits[w * BP_THREADS + tid] = 0;
    }
    for (int w = 0; w < nwords; w++) {
        uint32_t b = sbits[w * BP_THREADS + tid];
        while (b) {
            const int k = __ffs(b) - 1;
            b &= b - 1;
            atomicAdd(&scnt[(w << 5) + k], 1u);
        }
    }
    __syncthreads();
    for (int p = tid; p < M.npairs; p += BP_THREADS) {
        const uint32_t c = scnt[p];
        sbase[p] = c ? atomicAdd(bins.count + p, c) : 0u;
        scnt[p] = 0u;
    }
    __syncthreads();
    for (int w = 0; w < nwords; w++) {
        uint32_t b = sbits[w * BP_THREADS + tid];
        while (b) {
            const int k = __ffs(b) - 1;
            b &= b - 1;
            const int p = (w << 5) + k;
            const uint32_t r = atomicAdd(&scnt[p], 1u);
            bins.state[(int64_t)p * bins.cap + sbase[p] + r] = (int32_t)state;
        }
    }

    if (valid) {
        if (MODE == MODE_BOOL) {
            if (state_hit) out.hit[group] = 1;
        } else if (MODE == MODE_BOOL_ALL) {
            if (state_hit) {
                out.hit[group] = 1;
                if (out.first_pair) atomicMin(out.first_pair + state, first_pair);
                if (out.first_bad) atomicMin(out.first_bad + group, src.sample ? src.sample[state] : 0);
            }
        } else {
            if (state_hit) out.hit[group] = 1;
            // bound is a certain upper bound (+slack) of the true minimum; items refine it downwards
            out.min_dist[state] = bound;
        }
    }
}

// ------------------------------------------------------------------------------------------
// Tile scheduling shared by the item setup and the narrowphase: tiles of TILE_ITEMS items of one
// pair, bins visited in `order` (most expensive pairs first).  prefix[k] = tiles before bin k.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void compute_tile_prefix(const uint32_t *count, const int32_t *order, int npairs, uint32_t *prefix) {
    if (threadIdx.x < 32) {
        uint32_t run = 0;
        for (int k0 = 0; k0 < npairs; k0 += 32) {
            const int k = k0 + threadIdx.x;
            const uint32_t t = k < npairs ? (count[order[k]] + TILE_ITEMS - 1) / TILE_ITEMS : 0u;
            uint32_t inc = t;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, inc, o);
                if ((int)threadIdx.x >= o) inc += y;
            }
            if (k < npairs) prefix[k] = run + inc - t;
            run += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (threadIdx.x == 0) prefix[npairs] = run;
    }
}

// Fetch the next tile for this warp.  Returns false when the work is exhausted.
__device__ __forceinline__ bool next_tile(uint32_t *tile_counter, const uint32_t *prefix, const int32_t *order, int npairs,
                                          const uint32_t *count, int &pair, uint32_t &begin, uint32_t &end) {
    uint32_t tile = 0;
    if ((threadIdx.x & 31) == 0) tile = atomicAdd(tile_counter, 1u);
    tile = __shfl_sync(0xffffffffu, tile, 0);
    if (tile >= prefix[npairs]) return false;
    int lo = 0, hi = npairs;  // last k with prefix[k] <= tile
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (prefix[mid] <= tile) lo = mid; else hi = mid;
    }
    pair = order[lo];
    begin = (tile - prefix[lo]) * TILE_ITEMS;
    end = min(begin + TILE_ITEMS, count[pair]);
    return true;
}

// ------------------------------------------------------------------------------------------
// Surface semantics: can this hull hit be an enclosure at all?
// ------------------------------------------------------------------------------------------
// A shape wholly inside a mesh SURFACE does not collide with it (k_surface_refine below).  Most hull
// hits of a flagged pair are ordinary partial overlaps, and one direction d in which the would-be
// inner shape reaches at least as far as the would-be outer one, h_inner(d) + padding >= h_outer(d),
// proves that.  Tried, per flagged role: the line between the proxy centres (the inner shape pokes
// out on the side its centre is displaced to) and both directions of the inner shape's long axis
// (a finger half inside a link pokes out with one end).  Two support evaluations per direction, on
// the lane that found the hit; only the items no direction settles are parked for the plane test.
#ifndef PBP_SR_AXIS_DIRS
#define PBP_SR_AXIS_DIRS 0   // measured on config 2: the centre line settles 73 % of the flagged hull hits, the axis directions 1 % more
#endif
__device__ __forceinline__ float support_value(const ShapeRef &S, float3 d, float radius, float len) {
    const float3 sl = support_local(S, d);
    return dot3(d, sl) + radius * len;
}
__device__ __forceinline__ float support_value_placed(const ShapeRef &S, const float *T, float3 d, float radius, float len) {
    const float3 dl = make_float3(fmaf(T[0], d.x, fmaf(T[3], d.y, T[6] * d.z)), fmaf(T[1], d.x, fmaf(T[4], d.y, T[7] * d.z)),
                                  fmaf(T[2], d.x, fmaf(T[5], d.y, T[8] * d.z)));
    const float3 sl = support_local(S, dl);
    return dot3(d, se3_act(T, sl.x, sl.y, sl.z)) + radius * len;
}
__device__ __noinline__ bool may_be_enclosed(int fl, const DevGeom &GA, const DevGeom &GB, const ShapeRef &A, const ShapeRef &B, const float *T,
                                             float padding) {
    const float3 ca = make_float3(GA.centre[0], GA.centre[1], GA.centre[2]);
    const float3 cb = se3_act(T, GB.centre[0], GB.centre[1], GB.centre[2]);
    float3 d = sub3(cb, ca);
    if (!(dot3(d, d) > 1e-12f)) d = make_float3(1.f, 0.f, 0.f);
    if (fl & 1) {   // b inside a's surface?
        const float3 p0 = se3_act(T, GB.axis0[0], GB.axis0[1], GB.axis0[2]), p1 = se3_act(T, GB.axis1[0], GB.axis1[1], GB.axis1[2]);
        const float3 u = sub3(p1, p0);
        bool out = false;
#pragma unroll 1
        for (int k = 0; k < (PBP_SR_AXIS_DIRS && GB.has_axis ? 3 : 1) && !out; k++) {
            const float3 dir = k == 0 ? d : (k == 1 ? u : make_float3(-u.x, -u.y, -u.z));
            const float len = sqrtf(dot3(dir, dir));
            out = support_value_placed(B, T, dir, GB.core_radius, len) + padding * len >= support_value(A, dir, GA.core_radius, len);
        }
        if (!out) return true;
    }
    if (fl & 2) {   // a inside b's surface?
        const float3 u = make_float3(GA.axis1[0] - GA.axis0[0], GA.axis1[1] - GA.axis0[1], GA.axis1[2] - GA.axis0[2]);
        bool out = false;
#pragma unroll 1
        for (int k = 0; k < (PBP_SR_AXIS_DIRS && GA.has_axis ? 3 : 1) && !out; k++) {
            const float3 dir = k == 0 ? make_float3(-d.x, -d.y, -d.z) : (k == 1 ? u : make_float3(-u.x, -u.y, -u.z));
            const float len = sqrtf(dot3(dir, dir));
            out = support_value(A, dir, GA.core_radius, len) + padding * len >= support_value_placed(B, T, dir, GB.core_radius, len);
        }
        if (!out) return true;
    }
    return false;
}

// ------------------------------------------------------------------------------------------
// Item setup: T = oMg[a]^-1 * oMg[b] for every undecided (state, pair) item
// ------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(IS_THREADS, PBP_IS_MIN_BLOCKS)
k_item_setup(DevModel M, StateSource src, float padding, Outputs out, Bins bins, uint32_t *tile_counter) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned char *ptr = smem_raw;
    DevJoint *sj = reinterpret_cast<DevJoint *>(ptr);    ptr += align16((size_t)M.njoints * sizeof(DevJoint));
    DevGeom *sg = reinterpret_cast<DevGeom *>(ptr);      ptr += align16((size_t)max(M.ngeoms, 1) * sizeof(DevGeom));
    BpPair *sp = reinterpret_cast<BpPair *>(ptr);        ptr += align16((size_t)max(M.npairs, 1) * sizeof(BpPair));
    PairPath *spath = reinterpret_cast<PairPath *>(ptr); ptr += align16((size_t)max(M.npairs, 1) * sizeof(PairPath));
    uint32_t *prefix = reinterpret_cast<uint32_t *>(ptr); ptr += align16(((size_t)M.npairs + 1) * 4);
    int32_t *order = reinterpret_cast<int32_t *>(ptr);   ptr += align16((size_t)max(M.npairs, 1) * 4);

    stage_words(sj, M.joints, M.njoints * (int)sizeof(DevJoint));
    stage_words(sg, M.geoms, M.ngeoms * (int)sizeof(DevGeom));
    stage_words(sp, M.bppairs, M.npairs * (int)sizeof(BpPair));
    stage_words(spath, M.paths, M.npairs * (int)sizeof(PairPath));
    stage_words(order, M.bin_order, M.npairs * 4);
    __syncthreads();
    compute_tile_prefix(bins.count, order, M.npairs, prefix);
    __syncthreads();

    const unsigned lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int64_t stride = (int64_t)M.npairs * bins.cap;
    int p;
    uint32_t begin, end;
    while (next_tile(tile_counter, prefix, order, M.npairs, bins.count, p, begin, end)) {
        const PairPath path = spath[p];
        const uint8_t *ops = M.path_ops + path.off;   // tiny, L1/L2 resident, warp-uniform addresses
        const DevGeom &GA = sg[sp[p].a];
        const DevGeom &GB = sg[sp[p].b];
        // shapes of the pair (hull tables read through L1: warp-uniform base addresses)
        ShapeRef A, B;
        A.shape = GA.shape; A.p0 = GA.par[0]; A.p1 = GA.par[1]; A.p2 = GA.par[2];
        A.verts = M.verts + GA.vert_off; A.nverts = GA.vert_cnt;
        A.cell_off = GA.map_off >= 0 ? M.sup_cells + GA.map_off : nullptr; A.cell_list = M.sup_lists + GA.list_off;
        B.shape = GB.shape; B.p0 = GB.par[0]; B.p1 = GB.par[1]; B.p2 = GB.par[2];
        B.verts = M.verts + GB.vert_off; B.nverts = GB.vert_cnt;
        B.cell_off = GB.map_off >= 0 ? M.sup_cells + GB.map_off : nullptr; B.cell_list = M.sup_lists + GB.list_off;
        const float radii = GA.core_radius + GB.core_radius;
        for (uint32_t base = begin; base < end; base += 32) {
            const uint32_t idx = base + lane;
            bool keep = idx < end;
            int32_t state = 0, group = 0;
            float R[12];
            if (keep) {
                state = bins.state[(int64_t)p * bins.cap + idx];
                group = src.group ? src.group[state] : state;
                if (MODE == MODE_BOOL && out.hit[group]) keep = false;   // the broadphase already decided this group
            }
            if (keep) {
                // joint value of this state
                auto qval = [&](int k) {
                    if (src.mode == 0) return __ldg(src.q + (int64_t)state * M.nq + k);
                    const float f = src.frac[state];
                    const float qa = __ldg(src.q1 + (int64_t)group * M.nq + k), qb = __ldg(src.q2 + (int64_t)group * M.nq + k);
                    return fmaf(f, qb - qa, qa);
                };
                float cur[12], lca[12], TA[12], nxt[12];
                se3_identity(cur);
                int o = 0;
                for (int i = 0; i < path.n_trunk; i++, o++) {
                    const DevJoint &J = sj[ops[o]];
                    joint_transform(J, cur, qval(J.idx_q), nxt);
#pragma unroll
                    for (int k = 0; k < 12; k++) cur[k] = nxt[k];
                }
#pragma unroll
                for (int k = 0; k < 12; k++) lca[k] = cur[k];
                for (int i = 0; i < path.n_a; i++, o++) {
                    const DevJoint &J = sj[ops[o]];
                    joint_transform(J, cur, qval(J.idx_q), nxt);
#pragma unroll
                    for (int k = 0; k < 12; k++) cur[k] = nxt[k];
                }
                se3_mul(cur, GA.P, TA);   // oMg[a]
#pragma unroll
                for (int k = 0; k < 12; k++) cur[k] = lca[k];
                for (int i = 0; i < path.n_b; i++, o++) {
                    const DevJoint &J = sj[ops[o]];
                    joint_transform(J, cur, qval(J.idx_q), nxt);
#pragma unroll
                    for (int k = 0; k < 12; k++) cur[k] = nxt[k];
                }
                float TB[12];
                se3_mul(cur, GB.P, TB);   // oMg[b]
                se3_inv_mul(TA, TB, R);
                // First GJK iteration, here where every lane of the warp is busy with the same pair:
                // about two thirds of the items are already separated along the line between the
                // proxy centres and never reach the (divergent) narrowphase.  Same arithmetic as the
                // narrowphase's own first trip, so the decision is the one it would have taken.
                GjkState gs;
                GjkOut go;
                gjk_init(gs, gjk_start_direction(GA, GB, R));
                const float margin = padding + radii;
                const float bound = MODE == MODE_DIST ? out.min_dist[group] + radii : 0.f;
                const bool finished = (MODE == MODE_DIST) ? gjk_support_phase<true>(A, B, R, gs, margin, bound, go)
                                                          : gjk_support_phase<false>(A, B, R, gs, margin, 0.f, go);
                if (finished) keep = false;   // at n == 0 the only exit is "certainly farther than the cut"
            }
            // compact the survivors to the front of the bin (warp-aggregated slot reservation)
            const unsigned km = __ballot_sync(0xffffffffu, keep);
            if (km == 0u) continue;
            uint32_t first = 0;
            if (lane == 0) first = atomicAdd(bins.count_np + p, (uint32_t)__popc(km));
            first = __shfl_sync(0xffffffffu, first, 0);
            if (keep) {
                const int64_t slot = (int64_t)p * bins.cap + first + __popc(km & lt_mask);
                bins.group[slot] = group;
                if (src.sample) bins.sample[slot] = src.sample[state];
#pragma unroll
                for (int k = 0; k < 12; k++) bins.T[k * stride + slot] = R[k];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// Narrowphase (GJK): persistent warps, one item per lane, lanes refilled as they finish
// ------------------------------------------------------------------------------------------
// The items of all bins, concatenated in bin order, form one global sequence; warps claim chunks
// of NP_CHUNK consecutive items from an atomic cursor and keep every lane busy: each trip of the
// loop runs ONE GJK iteration for every active lane, and lanes whose item finished take the next
// item of the chunk.  A lane carries its own pair (neighbouring items almost always share it, so
// shape tables are read at the same shared-memory addresses), which means a warp never has to
// drain when it crosses from one bin into the next.
#ifndef PBP_NP_CHUNK
#define PBP_NP_CHUNK 128
#endif
constexpr int NP_CHUNK = PBP_NP_CHUNK;
#ifndef PBP_NP_QUORUM
#define PBP_NP_QUORUM 12
#endif
constexpr int NP_SIMPLEX_QUORUM = PBP_NP_QUORUM;   // lanes waiting for a simplex phase that make it worth a trip

template <int MODE, bool TABLES_IN_SMEM>
__global__ void __launch_bounds__(NP_THREADS, PBP_NP_MIN_BLOCKS)
k_narrowphase(DevModel M, float padding, Outputs out, Bins bins, uint32_t *chunk_cursor, bool has_sample) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned char *ptr = smem_raw;
    float4 *sv = reinterpret_cast<float4 *>(ptr);        ptr += TABLES_IN_SMEM ? align16((size_t)M.nverts_padded * sizeof(float4)) : 0;
    uint16_t *scell = reinterpret_cast<uint16_t *>(ptr); ptr += TABLES_IN_SMEM ? align16((size_t)M.n_sup_cells * 2) : 0;
    uint16_t *slist = reinterpret_cast<uint16_t *>(ptr); ptr += TABLES_IN_SMEM ? align16((size_t)M.n_sup_lists * 2) : 0;
    DevGeom *sg = reinterpret_cast<DevGeom *>(ptr);      ptr += align16((size_t)max(M.ngeoms, 1) * sizeof(DevGeom));
    BpPair *sp = reinterpret_cast<BpPair *>(ptr);        ptr += align16((size_t)max(M.npairs, 1) * sizeof(BpPair));
    uint32_t *prefix = reinterpret_cast<uint32_t *>(ptr); ptr += align16(((size_t)M.npairs + 1) * 4);   // items before bin k
    int32_t *order = reinterpret_cast<int32_t *>(ptr);   ptr += align16((size_t)max(M.npairs, 1) * 4);

    if (TABLES_IN_SMEM) {
        stage_words(sv, M.verts, M.nverts_padded * (int)sizeof(float4));
        stage_words(scell, M.sup_cells, M.n_sup_cells * 2);
        stage_words(slist, M.sup_lists, M.n_sup_lists * 2);
    }
    stage_words(sg, M.geoms, M.ngeoms * (int)sizeof(DevGeom));
    stage_words(sp, M.bppairs, M.npairs * (int)sizeof(BpPair));
    stage_words(order, M.bin_order, M.npairs * 4);
    __syncthreads();
    if (threadIdx.x < 32) {   // exclusive scan of the bin counts, in bin order
        uint32_t run = 0;
        for (int k0 = 0; k0 < M.npairs; k0 += 32) {
            const int k = k0 + threadIdx.x;
            const uint32_t t = k < M.npairs ? bins.count_np[order[k]] : 0u;
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

    const unsigned lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const float4 *verts = TABLES_IN_SMEM ? sv : M.verts;
    const uint16_t *cells = TABLES_IN_SMEM ? scell : M.sup_cells;
    const uint16_t *lists = TABLES_IN_SMEM ? slist : M.sup_lists;
    const int64_t stride = (int64_t)M.npairs * bins.cap;
    const uint32_t total = prefix[M.npairs];

    // lane-resident item; phase: 0 = empty, 1 = needs a support phase, 2 = needs a simplex phase
    int phase = 0;
    GjkState s;
    float T[12];
    int32_t group = 0;
    int p = 0;               // this lane's pair
    int64_t slot = 0;        // this lane's bin slot
    float bound = 0.f;
    // chunk claimed by the warp
    uint32_t cursor = 0, chunk_end = 0;
    bool more = true;
    for (;;) {
        unsigned need = __ballot_sync(0xffffffffu, phase == 0);
        // batch the refills: the refill path (bin search, 12 strided loads, GJK init) runs with only
        // the free lanes active, so it is taken once enough lanes are free (or nothing else is left)
        if (need != 0xffffffffu && __popc(need) < PBP_NP_REFILL_MIN) need = 0u;
        if (need && cursor >= chunk_end && more) {
            // guided self-scheduling: NP_CHUNK items while plenty are left, down to one warp-load
            // near the end, so the last chunk of a launch (and every chunk of a small launch) is
            // short -- a 128-item chunk is ~100 us of serial GJK trips for one warp
            uint32_t c = 0, sz = 0;
            if (lane == 0) {
                const uint32_t seen = *reinterpret_cast<volatile uint32_t *>(chunk_cursor);
                const uint32_t left = seen < total ? total - seen : 0u;
                sz = min((uint32_t)NP_CHUNK, max(32u, left / (2u * gridDim.x * (NP_THREADS / 32))));
                c = atomicAdd(chunk_cursor, sz);
            }
            c = __shfl_sync(0xffffffffu, c, 0);
            sz = __shfl_sync(0xffffffffu, sz, 0);
            if (c >= total) more = false;
            else { cursor = c; chunk_end = min(c + sz, total); }
        }
        // ---- refill finished lanes with the next items of the chunk
        if (need && cursor < chunk_end) {
            const uint32_t gi = cursor + __popc(need & lt_mask);
            if (phase == 0 && gi < chunk_end) {
                int lo = 0, hi = M.npairs;  // last bin k with prefix[k] <= gi
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if (prefix[mid] <= gi) lo = mid; else hi = mid;
                }
                p = order[lo];
                slot = (int64_t)p * bins.cap + (gi - prefix[lo]);
                const int32_t g = bins.group[slot];
                if (g >= 0 && !(MODE == MODE_BOOL && out.hit[g])) {
#pragma unroll
                    for (int k = 0; k < 12; k++) T[k] = bins.T[k * stride + slot];
                    const DevGeom &GA = sg[sp[p].a];
                    const DevGeom &GB = sg[sp[p].b];
                    gjk_init(s, gjk_start_direction(GA, GB, T));
                    group = g;
                    if (MODE == MODE_DIST) bound = out.min_dist[g] + GA.core_radius + GB.core_radius;  // group == state here
                    phase = 1;
                }
            }
            cursor += __popc(need);
        }
        // ---- one GJK iteration for every live lane: the support phase, then -- for the lanes it did
        //      not finish -- the simplex phase.  Since the item setup removed the items that end
        //      after one support call, the items here all run several iterations, and keeping the
        //      lanes in lock-step (every live lane is at "support" at the top of a trip) wastes only
        //      the simplex slot of a lane's last trip.  (PBP_NP_LOCKSTEP=0 restores the older
        //      policy: per trip, run only the phase most lanes are waiting for.)
        auto publish = [&](const GjkOut &r, float radii) {
            // (distance mode: only OVERLAPPING hulls can be an enclosure; a pair merely within the
            // padding keeps its measured distance)
            if (r.hit && M.surface && M.pair_enclose[p] && (MODE != MODE_DIST || r.dist - radii <= 0.f)) {
                // surface semantics: one shape of this pair could be enclosed by the other's surface,
                // which is the free placement -- k_surface_refine decides the parked item (all its
                // lanes busy with such items, instead of one lane of this warp)
                bins.group[slot] = group | SR_PARKED;
                return;
            }
            if (MODE == MODE_DIST) {
                if (r.hit) out.hit[group] = 1;
                atomicMin(reinterpret_cast<unsigned int *>(out.min_dist + group), __float_as_uint(fmaxf(r.dist - radii, 0.f)));
            } else if (r.hit) {
                out.hit[group] = 1;
                if (MODE == MODE_BOOL_ALL) {
                    if (out.first_pair) atomicMin(out.first_pair + group, M.pair_orig[p]);
                    if (out.first_bad) atomicMin(out.first_bad + group, has_sample ? bins.sample[slot] : 0);
                }
            }
        };
        auto support = [&](GjkOut &r, float &radii) {
            const DevGeom &GA = sg[sp[p].a];
            const DevGeom &GB = sg[sp[p].b];
            ShapeRef A, B;
            A.shape = GA.shape; A.p0 = GA.par[0]; A.p1 = GA.par[1]; A.p2 = GA.par[2];
            A.verts = verts + GA.vert_off; A.nverts = GA.vert_cnt;
            A.cell_off = GA.map_off >= 0 ? cells + GA.map_off : nullptr; A.cell_list = lists + GA.list_off;
            B.shape = GB.shape; B.p0 = GB.par[0]; B.p1 = GB.par[1]; B.p2 = GB.par[2];
            B.verts = verts + GB.vert_off; B.nverts = GB.vert_cnt;
            B.cell_off = GB.map_off >= 0 ? cells + GB.map_off : nullptr; B.cell_list = lists + GB.list_off;
            radii = GA.core_radius + GB.core_radius;
            const float margin = padding + radii;
            return (MODE == MODE_DIST) ? gjk_support_phase<true>(A, B, T, s, margin, bound, r)
                                       : gjk_support_phase<false>(A, B, T, s, margin, 0.f, r);
        };
        auto simplex = [&](GjkOut &r, float &radii) {
            const float margin = padding + sg[sp[p].a].core_radius + sg[sp[p].b].core_radius;
            radii = margin - padding;
            return (MODE == MODE_DIST) ? gjk_simplex_phase<true>(s, margin, r) : gjk_simplex_phase<false>(s, margin, r);
        };
        const int n_sup = __popc(__ballot_sync(0xffffffffu, phase == 1));
        const int n_smp = __popc(__ballot_sync(0xffffffffu, phase == 2));
        if (n_sup == 0 && n_smp == 0) {
            if (!more && cursor >= chunk_end) break;
            continue;
        }
        GjkOut r;
        float radii = 0.f;
#if PBP_NP_LOCKSTEP
        if (phase == 1) {
            const bool finished = support(r, radii);
            phase = finished ? 0 : 2;
            if (finished) publish(r, radii);
        }
        if (phase == 2) {
            const bool finished = simplex(r, radii);
            phase = finished ? 0 : 1;
            if (finished) publish(r, radii);
        }
#else
        const bool refillable = more || cursor < chunk_end;
        const bool run_simplex = n_smp > 0 && (n_sup == 0 || n_smp >= NP_SIMPLEX_QUORUM || (!refillable && n_smp >= n_sup));
        bool finished = false;
        if (run_simplex) {
            if (phase == 2) {
                finished = simplex(r, radii);
                phase = finished ? 0 : 1;
            }
        } else if (phase == 1) {
            finished = support(r, radii);
            phase = finished ? 0 : 2;
        }
        if (finished) publish(r, radii);
#endif
    }
}

// ------------------------------------------------------------------------------------------
// Narrowphase, block-cooperative variant: a pool of NP_THREADS items in shared memory, re-sorted
// by phase every round
// ------------------------------------------------------------------------------------------
// In k_narrowphase a lane owns its item from start to finish, so a warp always holds a mix of
// lanes that need a support phase, lanes that need a simplex phase and lanes being refilled: on
// average 13 of 32 lanes are active (ncu smsp__thread_inst_executed_per_inst_executed).  Here the
// per-item GJK state lives in shared memory instead of registers, and every round the block
// partitions its pool into "needs support" and "needs simplex" lists and hands list entries to
// threads in order -- warps are (nearly) full and run ONE phase type.  The price is a
// load/store of the 17-word state plus the 12-word placement around each phase and three
// barriers per round.
constexpr int POOL = NP_THREADS;          // items resident per block
enum { PF_V = 0, PF_VV = 3, PF_W0 = 4, PF_W1 = 7, PF_W2 = 10, PF_W = 13, PF_T = 16, PF_BOUND = 28, PF_WORDS = 29 };

template <int MODE>
__global__ void __launch_bounds__(NP_THREADS, 2)
k_narrowphase_pool(DevModel M, float padding, Outputs out, Bins bins, uint32_t *chunk_cursor, bool has_sample) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned char *ptr = smem_raw;
    float4 *sv = reinterpret_cast<float4 *>(ptr);        ptr += align16((size_t)M.nverts_padded * sizeof(float4));
    uint16_t *scell = reinterpret_cast<uint16_t *>(ptr); ptr += align16((size_t)M.n_sup_cells * 2);
    uint16_t *slist = reinterpret_cast<uint16_t *>(ptr); ptr += align16((size_t)M.n_sup_lists * 2);
    DevGeom *sg = reinterpret_cast<DevGeom *>(ptr);      ptr += align16((size_t)max(M.ngeoms, 1) * sizeof(DevGeom));
    BpPair *sp = reinterpret_cast<BpPair *>(ptr);        ptr += align16((size_t)max(M.npairs, 1) * sizeof(BpPair));
    uint32_t *prefix = reinterpret_cast<uint32_t *>(ptr); ptr += align16(((size_t)M.npairs + 1) * 4);   // items before bin k
    int32_t *order = reinterpret_cast<int32_t *>(ptr);   ptr += align16((size_t)max(M.npairs, 1) * 4);
    float *pf = reinterpret_cast<float *>(ptr);          ptr += (size_t)PF_WORDS * POOL * 4;     // [PF_WORDS][POOL]
    int32_t *pgroup = reinterpret_cast<int32_t *>(ptr);  ptr += (size_t)POOL * 4;
    int32_t *pslot = reinterpret_cast<int32_t *>(ptr);   ptr += (size_t)POOL * 4;                // bin slot (low 32 bits suffice per pair: idx in bin)
    int32_t *pmeta = reinterpret_cast<int32_t *>(ptr);   ptr += (size_t)POOL * 4;                // pair | n << 16 | stalls << 20 | it << 24
    uint8_t *pphase = reinterpret_cast<uint8_t *>(ptr);  ptr += (size_t)POOL;                    // 0 free, 1 support, 2 simplex
    uint16_t *lsup = reinterpret_cast<uint16_t *>(ptr);  ptr += (size_t)POOL * 2;
    uint16_t *lsmp = reinterpret_cast<uint16_t *>(ptr);  ptr += (size_t)POOL * 2;
    uint32_t *wcnt = reinterpret_cast<uint32_t *>(ptr);  ptr += 3 * (NP_THREADS / 32) * 4 + 16;  // per-warp counts: free, support, simplex
    __shared__ uint32_t s_base, s_take, s_nsup, s_nsmp;
    __shared__ int s_more;

    stage_words(sv, M.verts, M.nverts_padded * (int)sizeof(float4));
    stage_words(scell, M.sup_cells, M.n_sup_cells * 2);
    stage_words(slist, M.sup_lists, M.n_sup_lists * 2);
    stage_words(sg, M.geoms, M.ngeoms * (int)sizeof(DevGeom));
    stage_words(sp, M.bppairs, M.npairs * (int)sizeof(BpPair));
    stage_words(order, M.bin_order, M.npairs * 4);
    pphase[threadIdx.x] = 0;
    if (threadIdx.x == 0) s_more = 1;
    __syncthreads();
    if (threadIdx.x < 32) {   // exclusive scan of the bin counts, in bin order
        uint32_t run = 0;
        for (int k0 = 0; k0 < M.npairs; k0 += 32) {
            const int k = k0 + threadIdx.x;
            const uint32_t t = k < M.npairs ? bins.count_np[order[k]] : 0u;
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

    const int tid = threadIdx.x;
    const unsigned lane = tid & 31, warp = tid >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int64_t stride = (int64_t)M.npairs * bins.cap;
    const uint32_t total = prefix[M.npairs];
    constexpr int NW = NP_THREADS / 32;

    for (;;) {
        // ---- (a) count free slots, claim that many items from the global cursor
        const bool is_free = pphase[tid] == 0;
        const unsigned fm = __ballot_sync(0xffffffffu, is_free);
        if (lane == 0) wcnt[warp] = __popc(fm);
        __syncthreads();
        if (tid == 0) {
            uint32_t nfree = 0;
            for (int w = 0; w < NW; w++) nfree += wcnt[w];
            uint32_t base = 0, take = 0;
            if (s_more && nfree >= (uint32_t)(POOL / 4)) {    // refill in chunks: one atomic per >= POOL/4 items
                // guided: near the end of the launch blocks take smaller bites so that they finish together
                const uint32_t seen = *reinterpret_cast<volatile uint32_t *>(chunk_cursor);
                const uint32_t left = seen < total ? total - seen : 0u;
                const uint32_t want = min(nfree, max(32u, left / (2u * gridDim.x)));
                base = atomicAdd(chunk_cursor, want);
                if (base >= total) s_more = 0;
                else take = min(want, total - base);
                if (base + want >= total) s_more = 0;
            }
            s_base = base;
            s_take = take;
        }
        __syncthreads();
        // ---- (b) free slots pick up their item: rank among the free slots of the block
        if (is_free && s_take > 0) {
            uint32_t rank = __popc(fm & lt_mask);
            for (unsigned w = 0; w < warp; w++) rank += wcnt[w];
            if (rank < s_take) {
                const uint32_t gi = s_base + rank;
                int lo = 0, hi = M.npairs;  // last bin k with prefix[k] <= gi
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if (prefix[mid] <= gi) lo = mid; else hi = mid;
                }
                const int p = order[lo];
                const uint32_t idx = gi - prefix[lo];
                const int64_t slot = (int64_t)p * bins.cap + idx;
                const int32_t g = bins.group[slot];
                if (g >= 0 && !(MODE == MODE_BOOL && out.hit[g])) {
                    float T[12];
#pragma unroll
                    for (int k = 0; k < 12; k++) T[k] = bins.T[k * stride + slot];
                    const DevGeom &GA = sg[sp[p].a];
                    const DevGeom &GB = sg[sp[p].b];
                    GjkState s;
                    gjk_init(s, gjk_start_direction(GA, GB, T));
                    pf[(PF_V + 0) * POOL + tid] = s.v.x; pf[(PF_V + 1) * POOL + tid] = s.v.y; pf[(PF_V + 2) * POOL + tid] = s.v.z;
                    pf[PF_VV * POOL + tid] = s.vv;
#pragma unroll
                    for (int k = 0; k < 12; k++) pf[(PF_T + k) * POOL + tid] = T[k];
                    if (MODE == MODE_DIST) pf[PF_BOUND * POOL + tid] = out.min_dist[g] + GA.core_radius + GB.core_radius;
                    pgroup[tid] = g;
                    pslot[tid] = (int32_t)idx;
                    pmeta[tid] = p;            // n = 0, stalls = 0, it = 0
                    pphase[tid] = 1;
                }
            }
        }
        // ---- (c) phase lists
        const int ph = pphase[tid];
        const unsigned ms = __ballot_sync(0xffffffffu, ph == 1), mq = __ballot_sync(0xffffffffu, ph == 2);
        if (lane == 0) { wcnt[NW + warp] = __popc(ms); wcnt[2 * NW + warp] = __popc(mq); }
        __syncthreads();
        {
            uint32_t bs = 0, bq = 0, ts = 0, tq = 0;
            for (int w = 0; w < NW; w++) {
                const uint32_t a = wcnt[NW + w], b = wcnt[2 * NW + w];
                if (w < (int)warp) { bs += a; bq += b; }
                ts += a; tq += b;
            }
            if (ph == 1) lsup[bs + __popc(ms & lt_mask)] = (uint16_t)tid;
            if (ph == 2) lsmp[bq + __popc(mq & lt_mask)] = (uint16_t)tid;
            if (tid == 0) { s_nsup = ts; s_nsmp = tq; }
        }
        __syncthreads();
        const uint32_t nsup = s_nsup, nsmp = s_nsmp;
        if (nsup == 0 && nsmp == 0) {
            if (!s_more) break;
            continue;    // pool empty but the cursor has more (only when a claimed chunk was all dropped)
        }
        // ---- (d) work: support items first, simplex items from the next warp boundary
        const uint32_t sup_pad = (nsup + 31u) & ~31u;
        for (uint32_t wi = tid; wi < sup_pad + nsmp; wi += NP_THREADS) {
            const bool do_sup = wi < nsup;
            const bool do_smp = wi >= sup_pad;
            if (!do_sup && !do_smp) continue;
            const int it = do_sup ? lsup[wi] : lsmp[wi - sup_pad];
            const int meta = pmeta[it];
            const int p = meta & 0xffff;
            GjkState s;
            s.v = make_float3(pf[(PF_V + 0) * POOL + it], pf[(PF_V + 1) * POOL + it], pf[(PF_V + 2) * POOL + it]);
            s.vv = pf[PF_VV * POOL + it];
            s.W0 = make_float3(pf[(PF_W0 + 0) * POOL + it], pf[(PF_W0 + 1) * POOL + it], pf[(PF_W0 + 2) * POOL + it]);
            s.W1 = make_float3(pf[(PF_W1 + 0) * POOL + it], pf[(PF_W1 + 1) * POOL + it], pf[(PF_W1 + 2) * POOL + it]);
            s.W2 = make_float3(pf[(PF_W2 + 0) * POOL + it], pf[(PF_W2 + 1) * POOL + it], pf[(PF_W2 + 2) * POOL + it]);
            s.n = (meta >> 16) & 0xf; s.stalls = (meta >> 20) & 0xf; s.it = (meta >> 24) & 0xff;
            const DevGeom &GA = sg[sp[p].a];
            const DevGeom &GB = sg[sp[p].b];
            const float radii = GA.core_radius + GB.core_radius;
            const float margin = padding + radii;
            bool finished;
            GjkOut r;
            if (do_sup) {
                float T[12];
#pragma unroll
                for (int k = 0; k < 12; k++) T[k] = pf[(PF_T + k) * POOL + it];
                ShapeRef A, B;
                A.shape = GA.shape; A.p0 = GA.par[0]; A.p1 = GA.par[1]; A.p2 = GA.par[2];
                A.verts = sv + GA.vert_off; A.nverts = GA.vert_cnt;
                A.cell_off = GA.map_off >= 0 ? scell + GA.map_off : nullptr; A.cell_list = slist + GA.list_off;
                B.shape = GB.shape; B.p0 = GB.par[0]; B.p1 = GB.par[1]; B.p2 = GB.par[2];
                B.verts = sv + GB.vert_off; B.nverts = GB.vert_cnt;
                B.cell_off = GB.map_off >= 0 ? scell + GB.map_off : nullptr; B.cell_list = slist + GB.list_off;
                const float bound = MODE == MODE_DIST ? pf[PF_BOUND * POOL + it] : 0.f;
                finished = (MODE == MODE_DIST) ? gjk_support_phase<true>(A, B, T, s, margin, bound, r)
                                               : gjk_support_phase<false>(A, B, T, s, margin, 0.f, r);
                if (!finished) {
                    pf[(PF_W + 0) * POOL + it] = s.w.x; pf[(PF_W + 1) * POOL + it] = s.w.y; pf[(PF_W + 2) * POOL + it] = s.w.z;
                    pphase[it] = 2;
                }
            } else {
                s.w = make_float3(pf[(PF_W + 0) * POOL + it], pf[(PF_W + 1) * POOL + it], pf[(PF_W + 2) * POOL + it]);
                finished = (MODE == MODE_DIST) ? gjk_simplex_phase<true>(s, margin, r) : gjk_simplex_phase<false>(s, margin, r);
                if (!finished) {
                    pf[(PF_V + 0) * POOL + it] = s.v.x; pf[(PF_V + 1) * POOL + it] = s.v.y; pf[(PF_V + 2) * POOL + it] = s.v.z;
                    pf[PF_VV * POOL + it] = s.vv;
                    pf[(PF_W0 + 0) * POOL + it] = s.W0.x; pf[(PF_W0 + 1) * POOL + it] = s.W0.y; pf[(PF_W0 + 2) * POOL + it] = s.W0.z;
                    pf[(PF_W1 + 0) * POOL + it] = s.W1.x; pf[(PF_W1 + 1) * POOL + it] = s.W1.y; pf[(PF_W1 + 2) * POOL + it] = s.W1.z;
                    pf[(PF_W2 + 0) * POOL + it] = s.W2.x; pf[(PF_W2 + 1) * POOL + it] = s.W2.y; pf[(PF_W2 + 2) * POOL + it] = s.W2.z;
                    pphase[it] = 1;
                }
            }
            if (!finished) {
                pmeta[it] = p | (s.n << 16) | (min(s.stalls, 15) << 20) | (min(s.it, 255) << 24);
            } else {
                pphase[it] = 0;
                const int32_t group = pgroup[it];
                if (MODE == MODE_DIST) {
                    if (r.hit) out.hit[group] = 1;
                    atomicMin(reinterpret_cast<unsigned int *>(out.min_dist + group), __float_as_uint(fmaxf(r.dist - radii, 0.f)));
                } else if (r.hit) {
                    out.hit[group] = 1;
                    if (MODE == MODE_BOOL_ALL) {
                        if (out.first_pair) atomicMin(out.first_pair + group, M.pair_orig[p]);
                        if (out.first_bad) atomicMin(out.first_bad + group, has_sample ? bins.sample[(int64_t)p * bins.cap + pslot[it]] : 0);
                    }
                }
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// Surface semantics of mesh geometries: enclosure test for parked hull hits
// ------------------------------------------------------------------------------------------
// The reference's meshes are SURFACES (coal BVHModel): a shape that lies wholly inside a mesh without
// reaching its surface does not collide with it.  For a convex hull A = {x : n_f.x <= w_f} and a
// convex shape B, the gap of B to A's boundary is  min_f (w_f - h_B(n_f))  (h_B: support function);
// B is enclosed with clearance beyond `padding` when that gap exceeds padding.  (The mesh surface lies
// up to DevModel::surface_tol below its hull where a quad is triangulated along its valley; an
// enclosed shape closer than that to a dented face is the residual difference to the reference,
// a few states in a million -- tools/soup_study.py, tools/gpu_surface_check.py.)  The faces are split
// over `nlanes` cooperating lanes; the caller reduces with a warp maximum when nlanes > 1.
//   T: placement of `inner` in the frame of the geometry `g_outer`; returns max_f (h_inner(n_f) - w_f).
__device__ __forceinline__ float enclosure_excess(const DevModel &M, const float4 *planes, int g_outer, const ShapeRef &inner, float inner_radius,
                                                  const float *T, float padding, int lane, int nlanes, int max_faces = INT_MAX) {
    const int2 pi = M.plane_idx[g_outer];
    const int nfaces = min(pi.y, max_faces);
    float worst = -FLT_MAX;
    for (int f0 = 0; f0 < nfaces; f0 += nlanes) {
        const int f = f0 + lane;
        if (f < nfaces) {
            const float4 pl = planes[pi.x + f];
            const float3 dl = make_float3(fmaf(T[0], pl.x, fmaf(T[3], pl.y, T[6] * pl.z)), fmaf(T[1], pl.x, fmaf(T[4], pl.y, T[7] * pl.z)),
                                          fmaf(T[2], pl.x, fmaf(T[5], pl.y, T[8] * pl.z)));
            const float3 sl = support_local(inner, dl);
            const float3 sw = se3_act(T, sl.x, sl.y, sl.z);
            worst = fmaxf(worst, fmaf(pl.x, sw.x, fmaf(pl.y, sw.y, pl.z * sw.z)) + inner_radius - pl.w);
        }
        // one face the inner shape reaches (to within padding) settles it: not enclosed.  Nearly all
        // parked items end here after the first batch of faces.
        const bool out = worst >= -padding;
        if (nlanes > 1 ? __any_sync(0xffffffffu, out) : out) return FLT_MAX;
    }
    return worst;
}

// flags bit 0: b could lie inside a's surface; bit 1: a inside b's.  T: placement of b in a's frame.
// gap_out (optional): the clearance of the enclosed shape to the enclosing hull.
// max_faces: only the first (largest) faces of the enclosing hull -- a "false" is then a proof that the
// inner shape pokes out, a "true" only means that none of those faces was reached.
__device__ __forceinline__ bool surface_enclosed(const DevModel &M, const float4 *planes, int flags, int ga, int gb, const ShapeRef &A,
                                                 const ShapeRef &B, float ra, float rb, const float *T, float padding, int lane, int nlanes,
                                                 float *gap_out = nullptr, int max_faces = INT_MAX) {
    if (flags & 1) {
        float ex = enclosure_excess(M, planes, ga, B, rb, T, padding, lane, nlanes, max_faces);
        if (nlanes > 1) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) ex = fmaxf(ex, __shfl_xor_sync(0xffffffffu, ex, o));
        }
        if (-ex > padding) { if (gap_out) *gap_out = -ex; return true; }
    }
    if (flags & 2) {
        float Ti[12];
#pragma unroll
        for (int i = 0; i < 3; i++) {
            Ti[3 * i + 0] = T[i]; Ti[3 * i + 1] = T[3 + i]; Ti[3 * i + 2] = T[6 + i];
            Ti[9 + i] = -(T[i] * T[9] + T[3 + i] * T[10] + T[6 + i] * T[11]);
        }
        float ex = enclosure_excess(M, planes, gb, A, ra, Ti, padding, lane, nlanes, max_faces);
        if (nlanes > 1) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) ex = fmaxf(ex, __shfl_xor_sync(0xffffffffu, ex, o));
        }
        if (-ex > padding) { if (gap_out) *gap_out = -ex; return true; }
    }
    return false;
}

// Parked items (the narrowphase marked bins.group with SR_PARKED instead of publishing the hull hit)
// are about 0.1 per Panda state; 1 in 100 of them is a real enclosure.  The bins of the flagged pairs
// are cut into chunks of 32 items, the chunks dealt round-robin to the warps.  Pass 1, one item per
// lane: may_be_enclosed() -- a direction in which the would-be inner shape pokes out publishes the
// hit.  Pass 2, one item per warp: the face planes of the enclosing hull split over the 32 lanes.
// The hull tables and planes are staged in shared memory when they fit (every support evaluation is a
// chain of dependent table reads).
constexpr int SR_THREADS = 512;        // (768 threads at 80 registers: 72 us, 1024 at 64: 74 us, against 62)
constexpr int SR_DIR = 256;       // flagged pairs the chunk directory holds
constexpr int SR_LANE_FACES = 8;  // faces of the enclosing hull each lane tries by itself in pass 1

template <int MODE, bool TABLES_IN_SMEM>
__global__ void __launch_bounds__(SR_THREADS)
k_surface_refine(DevModel M, float padding, Outputs out, Bins bins, uint32_t *chunk_cursor, bool has_sample) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned char *ptr = smem_raw;
    float4 *sv = reinterpret_cast<float4 *>(ptr);        ptr += TABLES_IN_SMEM ? align16((size_t)M.nverts_padded * sizeof(float4)) : 0;
    float4 *spl = reinterpret_cast<float4 *>(ptr);       ptr += TABLES_IN_SMEM ? align16((size_t)M.nplanes * sizeof(float4)) : 0;
    uint16_t *scell = reinterpret_cast<uint16_t *>(ptr); ptr += TABLES_IN_SMEM ? align16((size_t)M.n_sup_cells * 2) : 0;
    uint16_t *slist = reinterpret_cast<uint16_t *>(ptr); ptr += TABLES_IN_SMEM ? align16((size_t)M.n_sup_lists * 2) : 0;
    if (TABLES_IN_SMEM) {
        stage_words(sv, M.verts, M.nverts_padded * (int)sizeof(float4));
        stage_words(spl, M.planes, M.nplanes * (int)sizeof(float4));
        stage_words(scell, M.sup_cells, M.n_sup_cells * 2);
        stage_words(slist, M.sup_lists, M.n_sup_lists * 2);
        __syncthreads();
    }
    const float4 *verts = TABLES_IN_SMEM ? sv : M.verts;
    const float4 *planes = TABLES_IN_SMEM ? spl : M.planes;
    const uint16_t *cells = TABLES_IN_SMEM ? scell : M.sup_cells;
    const uint16_t *lists = TABLES_IN_SMEM ? slist : M.sup_lists;
    const unsigned lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * SR_THREADS + threadIdx.x) >> 5, nwarps = gridDim.x * (SR_THREADS / 32);
    const int64_t stride = (int64_t)M.npairs * bins.cap;
    const float pad_test = MODE == MODE_DIST ? 0.f : padding;
    auto commit = [&](int p, int32_t group, int64_t slot, bool enclosed, float gap) {
        if (MODE == MODE_DIST) {
            // the distance of an enclosed shape is its clearance, whatever the padding (pad_test = 0)
            if (!enclosed || gap <= padding) out.hit[group] = 1;
            atomicMin(reinterpret_cast<unsigned int *>(out.min_dist + group), __float_as_uint(enclosed ? gap : 0.f));
        } else if (!enclosed) {
            out.hit[group] = 1;
            if (MODE == MODE_BOOL_ALL) {
                if (out.first_pair) atomicMin(out.first_pair + group, M.pair_orig[p]);
                if (out.first_bad) atomicMin(out.first_bad + group, has_sample ? bins.sample[slot] : 0);
            }
        }
    };
    // chunk directory: the flagged pairs and the running number of 32-item chunks before each
    __shared__ int s_pair[SR_DIR];
    __shared__ uint32_t s_first[SR_DIR + 1];
    __shared__ int s_n;
    if (threadIdx.x == 0) {
        int nf = 0;
        uint32_t acc = 0;
        for (int k = 0; k < M.npairs && nf >= 0; k++) {
            if (!M.pair_enclose[k]) continue;
            if (nf == SR_DIR) { nf = -1; break; }   // directory full: every warp scans the pair list itself
            s_pair[nf] = k;
            s_first[nf] = acc;
            acc += (bins.count_np[k] + 31u) >> 5;
            nf++;
        }
        if (nf >= 0) s_first[nf] = acc;
        s_n = nf;
    }
    __syncthreads();
    (void)warp; (void)nwarps;
    for (;;) {
        // next chunk of the concatenated flagged bins, claimed from a global cursor: the chunks differ
        // widely in cost (parked lanes, true enclosures), a fixed round-robin leaves warps idle
        uint32_t c = 0;
        if (lane == 0) c = atomicAdd(chunk_cursor, 1u);
        c = __shfl_sync(0xffffffffu, c, 0);
        int p = -1;
        uint32_t local = 0;
        if (s_n >= 0) {
            if (c >= s_first[s_n]) break;
            int lo = 0, hi = s_n;   // last entry with s_first[lo] <= c
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (s_first[mid] <= c) lo = mid; else hi = mid;
            }
            // (entries with an empty bin share their s_first with the next one: take the last of the run)
            p = s_pair[lo];
            local = c - s_first[lo];
        } else {
            uint32_t acc = 0;
            for (int k = 0; k < M.npairs; k++) {
                if (!M.pair_enclose[k]) continue;
                const uint32_t nch = (bins.count_np[k] + 31u) >> 5;
                if (c < acc + nch) { p = k; local = c - acc; break; }
                acc += nch;
            }
        }
        if (p < 0) break;
        const int fl = M.pair_enclose[p];
        const uint32_t cnt = bins.count_np[p];
        const BpPair pr = M.bppairs[p];
        const DevGeom &GA = M.geoms[pr.a];
        const DevGeom &GB = M.geoms[pr.b];
        ShapeRef A, B;
        A.shape = GA.shape; A.p0 = GA.par[0]; A.p1 = GA.par[1]; A.p2 = GA.par[2];
        A.verts = verts + GA.vert_off; A.nverts = GA.vert_cnt;
        A.cell_off = GA.map_off >= 0 ? cells + GA.map_off : nullptr; A.cell_list = lists + GA.list_off;
        B.shape = GB.shape; B.p0 = GB.par[0]; B.p1 = GB.par[1]; B.p2 = GB.par[2];
        B.verts = verts + GB.vert_off; B.nverts = GB.vert_cnt;
        B.cell_off = GB.map_off >= 0 ? cells + GB.map_off : nullptr; B.cell_list = lists + GB.list_off;
        const uint32_t idx = local * 32u + lane;
        const int64_t slot = (int64_t)p * bins.cap + idx;
        const int32_t g = idx < cnt ? bins.group[slot] : 0;
        bool parked = idx < cnt && g >= 0 && (g & SR_PARKED);
        const int32_t group = g & ~SR_PARKED;
        if (parked && MODE == MODE_BOOL && out.hit[group]) parked = false;   // decided meanwhile
        float T[12];
#pragma unroll
        for (int k = 0; k < 12; k++) T[k] = parked ? bins.T[k * stride + slot] : 0.f;
        bool maybe = false;
        if (parked) {
            // the centre line settles 73 % of the parked items, the enclosing hull's 8 largest faces
            // 76 % of the rest (measured on config 2): 1.4 x the true enclosures reach pass 2
            maybe = may_be_enclosed(fl, GA, GB, A, B, T, pad_test) &&
                    surface_enclosed(M, planes, fl, pr.a, pr.b, A, B, GA.core_radius, GB.core_radius, T, pad_test, 0, 1, nullptr, SR_LANE_FACES);
            if (!maybe) commit(p, group, slot, false, 0.f);
        }
        unsigned todo = __ballot_sync(0xffffffffu, maybe);
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            float Ts[12];
#pragma unroll
            for (int k = 0; k < 12; k++) Ts[k] = __shfl_sync(0xffffffffu, T[k], src);
            const int32_t gs = __shfl_sync(0xffffffffu, group, src);
            if (MODE == MODE_BOOL && out.hit[gs]) continue;   // warp-uniform
            float gap = 0.f;
            const bool enclosed = surface_enclosed(M, planes, fl, pr.a, pr.b, A, B, GA.core_radius, GB.core_radius, Ts, pad_test, (int)lane, 32, &gap);
            if (lane == 0) commit(p, gs, (int64_t)p * bins.cap + local * 32u + src, enclosed, gap);
        }
    }
}

// ------------------------------------------------------------------------------------------
// Small batches: one fused kernel, one thread per (pair, state)
// ------------------------------------------------------------------------------------------
// The three-kernel pipeline is built for throughput; a single state or a single edge (the calls of a
// sequential RRT loop) pays three latency-bound launches plus their memsets.  For at most a few
// ten thousand (pair, state) items this kernel does everything in one launch: FK of the pair's two
// geometries along its joint path, the broadphase's proxy classification (same thresholds), and --
// only where that is undecided -- the whole GJK.  Boolean verdicts only.
constexpr int SMALL_THREADS = 128;

// One (pair, state) item.  Returns 0: nothing to publish, 1: the pair collides, 2: the hulls overlap
// but one shape may be enclosed by the other's surface -- R (placement of b in a's frame) is then
// handed to the warp-cooperative enclosure test of the kernel.
enum { SMALL_NONE = 0, SMALL_HIT = 1, SMALL_PARKED = 2 };
__device__ __forceinline__ int small_item(const DevModel &M, const StateSource &src, int64_t n_states, float padding, const Outputs &out,
                                          int64_t t, float *R, int32_t &group, int &p) {
    p = (int)(t / n_states);                            // pair-major: the lanes of a warp share the pair
    const int64_t state = t - (int64_t)p * n_states;
    group = src.group ? src.group[state] : (int32_t)state;
    if (out.hit[group]) return SMALL_NONE;              // already decided by another item of this launch
    const BpPair pr = M.bppairs[p];
    const DevGeom &GA = M.geoms[pr.a];
    const DevGeom &GB = M.geoms[pr.b];
    auto qval = [&](int k) {
        if (src.mode == 0) return __ldg(src.q + state * M.nq + k);
        const float f = src.frac[state];
        const float qa = __ldg(src.q1 + (int64_t)group * M.nq + k), qb = __ldg(src.q2 + (int64_t)group * M.nq + k);
        return fmaf(f, qb - qa, qa);
    };
    // ---- relative placement of b in a's frame (as k_item_setup)
    const PairPath path = M.paths[p];
    const uint8_t *ops = M.path_ops + path.off;
    float cur[12], lca[12], TA[12], TB[12], nxt[12];
    se3_identity(cur);
    int o = 0;
    for (int i = 0; i < path.n_trunk; i++, o++) {
        const DevJoint &J = M.joints[ops[o]];
        joint_transform(J, cur, qval(J.idx_q), nxt);
#pragma unroll
        for (int k = 0; k < 12; k++) cur[k] = nxt[k];
    }
#pragma unroll
    for (int k = 0; k < 12; k++) lca[k] = cur[k];
    for (int i = 0; i < path.n_a; i++, o++) {
        const DevJoint &J = M.joints[ops[o]];
        joint_transform(J, cur, qval(J.idx_q), nxt);
#pragma unroll
        for (int k = 0; k < 12; k++) cur[k] = nxt[k];
    }
    se3_mul(cur, GA.P, TA);
#pragma unroll
    for (int k = 0; k < 12; k++) cur[k] = lca[k];
    for (int i = 0; i < path.n_b; i++, o++) {
        const DevJoint &J = M.joints[ops[o]];
        joint_transform(J, cur, qval(J.idx_q), nxt);
#pragma unroll
        for (int k = 0; k < 12; k++) cur[k] = nxt[k];
    }
    se3_mul(cur, GB.P, TB);
    se3_inv_mul(TA, TB, R);
    // ---- proxy classification with the broadphase's thresholds, in a's frame
    const float3 ca = make_float3(GA.centre[0], GA.centre[1], GA.centre[2]);
    const float3 cb = se3_act(R, GB.centre[0], GB.centre[1], GB.centre[2]);
    float d2;
    if (pr.kind == BP_BOX_A) {
        const float ex = fmaxf(fabsf(cb.x) - GA.par[0], 0.f), ey = fmaxf(fabsf(cb.y) - GA.par[1], 0.f), ez = fmaxf(fabsf(cb.z) - GA.par[2], 0.f);
        d2 = fmaf(ex, ex, fmaf(ey, ey, ez * ez));
    } else if (pr.kind == BP_BOX_B) {
        const float3 l = se3_act_inv(R, ca);
        const float ex = fmaxf(fabsf(l.x) - GB.par[0], 0.f), ey = fmaxf(fabsf(l.y) - GB.par[1], 0.f), ez = fmaxf(fabsf(l.z) - GB.par[2], 0.f);
        d2 = fmaf(ex, ex, fmaf(ey, ey, ez * ez));
    } else {
        const float3 d = sub3(ca, cb);
        d2 = dot3(d, d);
    }
    const bool exact = pr.kind == BP_EXACT_SPHERES;
    const float slack = exact ? 0.f : BP_SLACK;
    const float rf = pr.r_out + padding + slack, rs = pr.r_in + padding - slack;
    if (rs >= 0.f && d2 <= rs * rs) return SMALL_HIT;                    // certainly colliding
    if (exact || d2 > rf * rf) return SMALL_NONE;                        // certainly free (or decided exactly)
    // ---- undecided: the whole GJK
    ShapeRef A, B;
    A.shape = GA.shape; A.p0 = GA.par[0]; A.p1 = GA.par[1]; A.p2 = GA.par[2];
    A.verts = M.verts + GA.vert_off; A.nverts = GA.vert_cnt;
    A.cell_off = GA.map_off >= 0 ? M.sup_cells + GA.map_off : nullptr; A.cell_list = M.sup_lists + GA.list_off;
    B.shape = GB.shape; B.p0 = GB.par[0]; B.p1 = GB.par[1]; B.p2 = GB.par[2];
    B.verts = M.verts + GB.vert_off; B.nverts = GB.vert_cnt;
    B.cell_off = GB.map_off >= 0 ? M.sup_cells + GB.map_off : nullptr; B.cell_list = M.sup_lists + GB.list_off;
    const float margin = padding + GA.core_radius + GB.core_radius;
    const GjkOut r = gjk_run<false>(A, B, R, gjk_start_direction(GA, GB, R), margin, 0.f);
    if (!r.hit) return SMALL_NONE;
    if (M.surface && M.pair_enclose[p] && may_be_enclosed(M.pair_enclose[p], GA, GB, A, B, R, padding)) return SMALL_PARKED;
    return SMALL_HIT;
}

__global__ void __launch_bounds__(SMALL_THREADS)
k_small_check(DevModel M, StateSource src, int64_t n_states, float padding, Outputs out) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float R[12];
#pragma unroll
    for (int k = 0; k < 12; k++) R[k] = 0.f;
    int32_t group = 0;
    int p = 0;
    const int status = t < n_states * M.npairs ? small_item(M, src, n_states, padding, out, t, R, group, p) : SMALL_NONE;
    if (status == SMALL_HIT) out.hit[group] = 1;
    if (!M.surface) return;
    // surface semantics: the (rare) possibly-enclosed items of this warp, one after the other, the
    // enclosing hull's faces split over the 32 lanes (a single lane would walk ~300 faces alone)
    const unsigned lane = threadIdx.x & 31;
    unsigned todo = __ballot_sync(0xffffffffu, status == SMALL_PARKED);
    while (todo) {
        const int srcl = __ffs(todo) - 1;
        todo &= todo - 1;
        float Rs[12];
#pragma unroll
        for (int k = 0; k < 12; k++) Rs[k] = __shfl_sync(0xffffffffu, R[k], srcl);
        const int ps = __shfl_sync(0xffffffffu, p, srcl);
        const int32_t gs = __shfl_sync(0xffffffffu, group, srcl);
        if (out.hit[gs]) continue;   // warp-uniform
        const BpPair pr = M.bppairs[ps];
        const DevGeom &GA = M.geoms[pr.a];
        const DevGeom &GB = M.geoms[pr.b];
        ShapeRef A, B;
        A.shape = GA.shape; A.p0 = GA.par[0]; A.p1 = GA.par[1]; A.p2 = GA.par[2];
        A.verts = M.verts + GA.vert_off; A.nverts = GA.vert_cnt;
        A.cell_off = GA.map_off >= 0 ? M.sup_cells + GA.map_off : nullptr; A.cell_list = M.sup_lists + GA.list_off;
        B.shape = GB.shape; B.p0 = GB.par[0]; B.p1 = GB.par[1]; B.p2 = GB.par[2];
        B.verts = M.verts + GB.vert_off; B.nverts = GB.vert_cnt;
        B.cell_off = GB.map_off >= 0 ? M.sup_cells + GB.map_off : nullptr; B.cell_list = M.sup_lists + GB.list_off;
        const bool enclosed = surface_enclosed(M, M.planes, M.pair_enclose[ps], pr.a, pr.b, A, B, GA.core_radius, GB.core_radius, Rs, padding, (int)lane, 32);
        if (!enclosed && lane == 0) out.hit[gs] = 1;
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
__global__ void __launch_bounds__(128)
k_fk_output(DevModel M, const float *q, int64_t n_states, float *oMi_out, float *oMg_out, float *oMf_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    DevJoint *sj = reinterpret_cast<DevJoint *>(smem_raw);
    DevGeom *sg = reinterpret_cast<DevGeom *>(smem_raw + align16((size_t)M.njoints * sizeof(DevJoint)));
    stage_words(sj, M.joints, M.njoints * (int)sizeof(DevJoint));
    stage_words(sg, M.geoms, M.ngeoms * (int)sizeof(DevGeom));
    __syncthreads();
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_states) return;
    float oMi[PBP_MAX_JOINTS][12];
    se3_identity(oMi[0]);
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

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
#include "pbp_surface.cuh"

namespace pbp {

#ifndef PBP_IS_MIN_BLOCKS
#define PBP_IS_MIN_BLOCKS 3   // minimum resident blocks the item setup is compiled for (80 registers; measured 121 -> 102 us)
#endif
#ifndef PBP_NP_REFILL_MIN
#define PBP_NP_REFILL_MIN 8   // measured on config 2: 173.7 us (1) / 172.9 (4) / 170.0 (8) / 170.7 (12) / 171.0 (16)
#endif
#ifndef PBP_NP_LOCKSTEP
// 0: per trip, run only the phase (support / simplex) most live lanes are waiting for; 1: every trip runs the support
// phase and then the simplex phase for every live lane.  Round 1 (0.28 items per state, most of them ending after
// two or three iterations): lock-step 170 against 194 us.  Round 2 (the inner balls take the easy hits away: 0.15
// items per state, 4.7 iterations each, a lane load or two per warp): 125 us in lock-step, 100 us without.
#define PBP_NP_LOCKSTEP 0
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
    // boolean queries: the (few) items the narrowphase leaves UNDECIDED between core and hull, compacted so that
    // the refine kernel runs 32 of them per warp: (pair, index in the pair's bin).  Entries beyond the
    // capacity are marked in `group` (SR_DOUBT) instead and found by a scan of all bins.
    int2 *doubt;
    float4 *doubt_v;         // ... and the direction the narrowphase's core run converged on (a warm start for the refine kernel)
    uint32_t *doubt_count;
    uint32_t doubt_cap;
    // items the refine kernel leaves to a whole warp (walk over the enclosing core's planes, triangle stage):
    // (pair | status << 24, slot) and (U, n) -- settled by k_surface_settle, one warp per item
    int2 *witem;
    float4 *witem_v;
    uint32_t *witem_count;
    uint32_t witem_cap;
};

struct Outputs {
    uint8_t *hit;          // [groups]
    float *min_dist;       // [states] or NULL
    int32_t *first_pair;   // [states] or NULL   (scratch semantics: INT_MAX = none)
    int32_t *first_bad;    // [groups] or NULL   (scratch semantics: INT_MAX = none)
};


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
// Item setup: T = oMg[a]^-1 * oMg[b] for every undecided (state, pair) item
// ------------------------------------------------------------------------------------------
#ifndef PBP_IS_NOENC
#define PBP_IS_NOENC 1        // developer switch: 0 leaves every core hit of a pair with an enclosure flag to the refine kernel
#endif
#ifndef PBP_INNER_BALLS
#define PBP_INNER_BALLS 1     // developer switch: 0 disables the inner-ball hit certificates of the item setup
#endif
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
        // shapes of the pair (tables read through L1: warp-uniform base addresses): boolean queries run
        // on the inner cores of surface geometries, distance queries on the hulls (pbp_surface.cuh)
        SurfTables tab{};
        tab.hverts = M.verts; tab.hcells = M.sup_cells; tab.hlists = M.sup_lists;
        tab.kverts = M.kverts; tab.kcells = M.ksup_cells; tab.klists = M.ksup_lists;
        const ShapeRef A = MODE == MODE_DIST ? hull_shape(tab, GA) : core_shape(tab, GA);
        const ShapeRef B = MODE == MODE_DIST ? hull_shape(tab, GB) : core_shape(tab, GB);
        const float radii = GA.core_radius + GB.core_radius;
        const float eps_pair = pair_eps(GA, GB);
        const int fl_enc_any = M.has_tris ? (M.pair_enclose[p] & PAIR_ENCLOSE) : 0;
        const int fl_enc = MODE != MODE_DIST ? fl_enc_any : 0;
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
                // (the trunk from the base to the lowest common ancestor cancels in oMg[a]^-1 * oMg[b]: both
                //  branches start from the identity at the ancestor)
                float cur[12], TA[12], nxt[12];
                se3_identity(cur);
                int o = path.n_trunk;
                for (int i = 0; i < path.n_a; i++, o++) {
                    const DevJoint &J = sj[ops[o]];
                    joint_transform(J, cur, qval(J.idx_q), nxt);
#pragma unroll
                    for (int k = 0; k < 12; k++) cur[k] = nxt[k];
                }
                se3_mul(cur, GA.P, TA);   // placement of a relative to the ancestor
                se3_identity(cur);
                for (int i = 0; i < path.n_b; i++, o++) {
                    const DevJoint &J = sj[ops[o]];
                    joint_transform(J, cur, qval(J.idx_q), nxt);
#pragma unroll
                    for (int k = 0; k < 12; k++) cur[k] = nxt[k];
                }
                float TB[12];
                se3_mul(cur, GB.P, TB);   // placement of b relative to the ancestor
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
                                                          : gjk_support_phase<false>(A, B, R, gs, margin, margin + eps_pair, go);
                if (finished) keep = false;   // at n == 0 the only exit is "certainly farther than the cut"
                if (keep) {
                    // What the first iteration could not separate is mostly colliding.  Overlapping inner balls
                    // (pbp_surface.cuh) settle most of it here, before the narrowphase spends GJK iterations on it.
                    // Pairs where one shape could sit inside the other's surface: most placements rule that out
                    // along the centre line (the would-be inner shape reaches at least as far as the outer hull);
                    // decided here, where the whole warp works on the same pair, a hit by the balls stands and a
                    // later core hit needs no enclosure test.
                    const float pad = MODE == MODE_DIST ? 0.f : padding;
                    const bool balls = PBP_INNER_BALLS && inner_balls_hit(GA, GB, R, pad);
                    bool pokes = false;
                    if (fl_enc_any && (balls || (PBP_IS_NOENC && fl_enc))) pokes = enclosure_quick_pokes(fl_enc_any, GA, GB, hull_shape(tab, GA), hull_shape(tab, GB), R, pad);
                    if (balls && (!fl_enc_any || pokes)) {
                        out.hit[group] = 1;
                        if (MODE == MODE_BOOL_ALL) {
                            if (out.first_pair) atomicMin(out.first_pair + group, M.pair_orig[p]);
                            if (out.first_bad) atomicMin(out.first_bad + group, src.sample ? src.sample[state] : 0);
                        }
                        if (MODE == MODE_DIST) atomicMin(reinterpret_cast<unsigned int *>(out.min_dist + group), __float_as_uint(0.f));
                        keep = false;
                    } else if (fl_enc && pokes) {
                        group |= SR_NOENC;
                    }
                }
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
    // boolean queries run on the inner cores of surface geometries, distance queries on the hulls
    constexpr bool CORES = MODE != MODE_DIST;
    const float4 *gverts = CORES ? M.kverts : M.verts;
    const uint16_t *gcells = CORES ? M.ksup_cells : M.sup_cells, *glists = CORES ? M.ksup_lists : M.sup_lists;
    const int nv = CORES ? M.nkverts_padded : M.nverts_padded, ncell = CORES ? M.n_ksup_cells : M.n_sup_cells,
              nlist = CORES ? M.n_ksup_lists : M.n_sup_lists;
    float4 *sv = reinterpret_cast<float4 *>(ptr);        ptr += TABLES_IN_SMEM ? align16((size_t)nv * sizeof(float4)) : 0;
    uint16_t *scell = reinterpret_cast<uint16_t *>(ptr); ptr += TABLES_IN_SMEM ? align16((size_t)ncell * 2) : 0;
    uint16_t *slist = reinterpret_cast<uint16_t *>(ptr); ptr += TABLES_IN_SMEM ? align16((size_t)nlist * 2) : 0;
    DevGeom *sg = reinterpret_cast<DevGeom *>(ptr);      ptr += align16((size_t)max(M.ngeoms, 1) * sizeof(DevGeom));
    BpPair *sp = reinterpret_cast<BpPair *>(ptr);        ptr += align16((size_t)max(M.npairs, 1) * sizeof(BpPair));
    uint32_t *prefix = reinterpret_cast<uint32_t *>(ptr); ptr += align16(((size_t)M.npairs + 1) * 4);   // items before bin k
    int32_t *order = reinterpret_cast<int32_t *>(ptr);   ptr += align16((size_t)max(M.npairs, 1) * 4);

    if (TABLES_IN_SMEM) {
        stage_words(sv, gverts, nv * (int)sizeof(float4));
        stage_words(scell, gcells, ncell * 2);
        stage_words(slist, glists, nlist * 2);
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
    const float4 *verts = TABLES_IN_SMEM ? sv : gverts;
    const uint16_t *cells = TABLES_IN_SMEM ? scell : gcells;
    const uint16_t *lists = TABLES_IN_SMEM ? slist : glists;
    const int64_t stride = (int64_t)M.npairs * bins.cap;
    const uint32_t total = prefix[M.npairs];

    // lane-resident item; phase: 0 = empty, 1 = needs a support phase, 2 = needs a simplex phase
    int phase = 0;
    GjkState s;
    float T[12];
    int32_t group = 0;
    bool noenc = false;      // the item setup ruled an enclosure out (pairs with an enclosure flag)
    int p = 0;               // this lane's pair
    int64_t slot = 0;        // this lane's bin slot
    float bound = 0.f;
    // chunk claimed by the warp
    uint32_t cursor = 0, chunk_end = 0;
    bool more = true;
    const long long tn0 = DBG_CLOCK();
    if (blockIdx.x == 0 && threadIdx.x == 0) DBG_ADD(30, total);
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
                const int32_t graw = bins.group[slot];
                const int32_t g = graw & ~SR_NOENC;
                if (graw >= 0 && !(MODE == MODE_BOOL && out.hit[g])) {
                    noenc = (graw & SR_NOENC) != 0;
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
        //      the simplex slot of a lane's last trip.  (PBP_NP_LOCKSTEP=0, the default since round 2:
        //      per trip, run only the phase most lanes are waiting for -- see the macro.)
        auto publish = [&](const GjkOut &r, float radii) {
            const int fl = M.has_tris ? M.pair_enclose[p] : 0;
            if (MODE == MODE_DIST) {
                if (fl & PAIR_SURFACE) {
                    // the hull distance is a lower bound of the surface distance; hull distance + eps an upper
                    // bound (through the cores) unless one shape may be enclosed by the other.  The refine
                    // kernel settles the items that can define the state's minimum.
                    if (r.far) return;
                    const float dh = fmaxf(r.dist - radii, 0.f);
                    const float eps = pair_eps(sg[sp[p].a], sg[sp[p].b]);
                    if (!((fl & PAIR_ENCLOSE) && r.dist <= 0.f) && !(fl & PAIR_NOCORE))
                        atomicMin(reinterpret_cast<unsigned int *>(out.min_dist + group), __float_as_uint(dh + eps));
                    bins.group[slot] = group | SR_DOUBT;
                    return;
                }
                if (r.hit) out.hit[group] = 1;
                atomicMin(reinterpret_cast<unsigned int *>(out.min_dist + group), __float_as_uint(fmaxf(r.dist - radii, 0.f)));
                return;
            }
            if (fl & PAIR_SURFACE) {
                if (r.hit && (fl & PAIR_ENCLOSE) && !noenc) DBG_ADD(35, 1);
                if (!r.hit && !r.far) DBG_ADD(36, 1);
                if (r.hit && (fl & PAIR_NOCORE)) { bins.group[slot] = group | SR_PARKED; return; }   // a hull hit proves nothing: the triangles decide
                if (r.hit && (fl & PAIR_ENCLOSE) && !noenc) {
                    // one shape of this pair could be enclosed by the other's surface, which is the free
                    // placement -- the refine kernel decides (all its lanes busy with such items)
                    bins.group[slot] = group | SR_PARKED;
                    return;
                }
                if (!r.hit && !r.far) {   // between the bounds: the triangles decide
                    const uint32_t di = atomicAdd(bins.doubt_count, 1u);
                    if (di < bins.doubt_cap) {
                        bins.doubt[di] = make_int2(p, (int)(slot - (int64_t)p * bins.cap));
                        bins.doubt_v[di] = make_float4(s.v.x, s.v.y, s.v.z, 0.f);
                    }
                    else bins.group[slot] = group | SR_DOUBT;
                    return;
                }
            }
            if (r.hit) {
                DBG_ADD(37, 1);
                out.hit[group] = 1;
                if (MODE == MODE_BOOL_ALL) {
                    if (out.first_pair) atomicMin(out.first_pair + group, M.pair_orig[p]);
                    if (out.first_bad) atomicMin(out.first_bad + group, has_sample ? bins.sample[slot] : 0);
                }
            } else DBG_ADD(38, 1);
        };
        auto support = [&](GjkOut &r, float &radii) {
            const DevGeom &GA = sg[sp[p].a];
            const DevGeom &GB = sg[sp[p].b];
            ShapeRef A, B;
            A.shape = GA.shape; A.p0 = GA.par[0]; A.p1 = GA.par[1]; A.p2 = GA.par[2];
            A.verts = verts + (CORES ? GA.kvert_off : GA.vert_off); A.nverts = CORES ? GA.kvert_cnt : GA.vert_cnt;
            const int moa = CORES ? GA.kmap_off : GA.map_off, mob = CORES ? GB.kmap_off : GB.map_off;
            A.cell_off = moa >= 0 ? cells + moa : nullptr; A.cell_list = lists + (CORES ? GA.klist_off : GA.list_off);
            B.shape = GB.shape; B.p0 = GB.par[0]; B.p1 = GB.par[1]; B.p2 = GB.par[2];
            B.verts = verts + (CORES ? GB.kvert_off : GB.vert_off); B.nverts = CORES ? GB.kvert_cnt : GB.vert_cnt;
            B.cell_off = mob >= 0 ? cells + mob : nullptr; B.cell_list = lists + (CORES ? GB.klist_off : GB.list_off);
            radii = GA.core_radius + GB.core_radius;
            const float margin = padding + radii;
            return (MODE == MODE_DIST) ? gjk_support_phase<true>(A, B, T, s, margin, bound, r)
                                       : gjk_support_phase<false, false, true>(A, B, T, s, margin, margin + pair_eps(GA, GB), r);
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
        if (lane == 0) { DBG_ADD(31, 1); DBG_ADD(32, n_sup); }
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
    if (lane == 0) { const long long dt = DBG_CLOCK() - tn0; DBG_ADD(33, dt); DBG_MAX(34, dt); }
}

// ------------------------------------------------------------------------------------------
// Surface semantics: settle the items the GJK kernels parked (pbp_surface.cuh)
// ------------------------------------------------------------------------------------------
// Parked items are a few in a thousand Panda states (undecided between hull and core) plus 0.1 per state
// (core hits of pairs with an enclosure flag, 1 in 100 of them a real enclosure).  The bins of the pairs
// with a surface are cut into chunks of 32 items, claimed from a global cursor.  One item per lane first
// (decide_boolean / decide_distance), then the warp settles what is left (settle_warp_items).  The hull
// tables and the planes are staged in shared memory when they fit (every support evaluation is a chain of
// dependent table reads); cores and triangles, read by a few items only, stay in global memory.
#ifndef PBP_SR_THREADS
#define PBP_SR_THREADS 512
#endif
constexpr int SR_THREADS = PBP_SR_THREADS;
constexpr int SR_DIR = 256;       // pairs the chunk directory holds
#ifndef PBP_SR_LIST_CHUNK
#define PBP_SR_LIST_CHUNK 8
#endif
constexpr uint32_t SR_LIST_CHUNK = PBP_SR_LIST_CHUNK;   // undecided items a warp takes at a time (<= 32)

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
    }
    SurfTables tab;
    tab.hverts = TABLES_IN_SMEM ? sv : M.verts;
    tab.hcells = TABLES_IN_SMEM ? scell : M.sup_cells;
    tab.hlists = TABLES_IN_SMEM ? slist : M.sup_lists;
    tab.planes = TABLES_IN_SMEM ? spl : M.planes;
    tab.kverts = M.kverts; tab.kcells = M.ksup_cells; tab.klists = M.ksup_lists;
    tab.mverts = M.mverts; tab.tris = M.tris; tab.tverts = M.tverts;
    __shared__ uint16_t s_tri[SR_THREADS / 32][2][TRI_LIST_CAP];
    uint16_t *la = s_tri[threadIdx.x >> 5][0], *lb = s_tri[threadIdx.x >> 5][1];
    const unsigned lane = threadIdx.x & 31;
    const int64_t stride = (int64_t)M.npairs * bins.cap;
    auto on_hit = [&](int p, int32_t group, int64_t slot) {
        out.hit[group] = 1;
        if (MODE == MODE_BOOL_ALL) {
            if (out.first_pair) atomicMin(out.first_pair + group, M.pair_orig[p]);
            if (out.first_bad) atomicMin(out.first_bad + group, has_sample ? bins.sample[slot] : 0);
        }
    };
    auto on_value = [&](int, int32_t group, int64_t, float d) {
        if (d <= padding) out.hit[group] = 1;
        atomicMin(reinterpret_cast<unsigned int *>(out.min_dist + group), __float_as_uint(d));
    };
    auto decided = [&](int32_t group) { return MODE == MODE_BOOL && out.hit[group] != 0; };
    // chunk directory: the pairs whose bins hold parked items and the running number of 32-item chunks before
    // each.  Boolean queries: the pairs with an enclosure flag (core hits parked by the narrowphase) -- the
    // undecided items come through the compact list; distance queries, and boolean ones whose list
    // overflowed: every pair with a surface.
    __shared__ int s_pair[SR_DIR];
    __shared__ uint32_t s_first[SR_DIR + 1];
    __shared__ int s_n;
    __shared__ uint32_t s_list_n;
    const uint32_t listed = MODE == MODE_DIST ? 0u : *bins.doubt_count;
    const bool scan_all = MODE == MODE_DIST || listed > bins.doubt_cap;
    const int want = scan_all ? PAIR_SURFACE : (PAIR_ENCLOSE | PAIR_NOCORE);
    if (threadIdx.x == 0) {
        int nf = 0;
        uint32_t acc = 0;
        for (int k = 0; k < M.npairs && nf >= 0; k++) {
            if (!(M.pair_enclose[k] & want)) continue;
            if (nf == SR_DIR) { nf = -1; break; }   // directory full: every warp scans the pair list itself
            s_pair[nf] = k;
            s_first[nf] = acc;
            acc += (bins.count_np[k] + 31u) >> 5;
            nf++;
        }
        if (nf >= 0) s_first[nf] = acc;
        s_n = nf;
        s_list_n = min(listed, bins.doubt_cap);
    }
    __syncthreads();
    const uint32_t list_n = s_list_n;
    uint32_t bin_chunks = 0;
    if (s_n >= 0) {
        bin_chunks = s_first[s_n];
    } else {
        for (int k = 0; k < M.npairs; k++)
            if (M.pair_enclose[k] & want) bin_chunks += (bins.count_np[k] + 31u) >> 5;
    }
    // the undecided items of the list (two converged GJK runs each, then maybe triangles) cost many times a
    // parked item: they go FIRST, in short chunks, so the launch does not end on a few warps still chewing them
    const uint32_t list_chunks = (list_n + SR_LIST_CHUNK - 1) / SR_LIST_CHUNK;
    const uint32_t total_chunks = bin_chunks + list_chunks;
    const long long tk0 = DBG_CLOCK();
    if (blockIdx.x == 0 && threadIdx.x == 0) { DBG_ADD(22, bin_chunks); DBG_ADD(23, list_n); }
    for (;;) {
        // next chunk of the concatenated bins / of the list, claimed from a global cursor: the chunks differ
        // widely in cost (parked lanes, true enclosures, triangle stages), a fixed round-robin leaves warps idle
        uint32_t c = 0;
        if (lane == 0) c = atomicAdd(chunk_cursor, 1u);
        c = __shfl_sync(0xffffffffu, c, 0);
        if (c >= total_chunks) break;
        const long long tl0 = DBG_CLOCK();
        int p = -1;
        int64_t slot = 0;
        bool active = false;
        int32_t g = 0;
        float3 v0 = make_float3(0.f, 0.f, 0.f);   // warm start of an undecided item (zero: none)
        const bool is_list_chunk = c < list_chunks;
        if (c < list_chunks) {
            // undecided items of the compact list, one per lane, each with its own pair
            const uint32_t i = c * SR_LIST_CHUNK + lane;
            if (lane < SR_LIST_CHUNK && i < list_n) {
                const int2 it = bins.doubt[i];
                p = it.x;
                slot = (int64_t)p * bins.cap + it.y;
                g = bins.group[slot] | SR_DOUBT;
                const float4 w4 = bins.doubt_v[i];
                v0 = make_float3(w4.x, w4.y, w4.z);
                active = true;
            }
        } else {
            uint32_t local = 0;
            c -= list_chunks;
            if (s_n >= 0) {
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
                    if (!(M.pair_enclose[k] & want)) continue;
                    const uint32_t nch = (bins.count_np[k] + 31u) >> 5;
                    if (c < acc + nch) { p = k; local = c - acc; break; }
                    acc += nch;
                }
            }
            const uint32_t cnt = bins.count_np[p];
            const uint32_t idx = local * 32u + lane;
            slot = (int64_t)p * bins.cap + idx;
            g = idx < cnt ? bins.group[slot] : 0;
            active = idx < cnt && g >= 0 && (g & SR_BITS);
        }
        if (p < 0) p = 0;
        const int fl = M.pair_enclose[p];
        const BpPair pr = M.bppairs[p];
        const DevGeom &GA = M.geoms[pr.a];
        const DevGeom &GB = M.geoms[pr.b];
        const int32_t group = g & ~(SR_BITS | SR_NOENC);
        if (active && decided(group)) active = false;
        float T[12];
#pragma unroll
        for (int k = 0; k < 12; k++) T[k] = active ? bins.T[k * stride + slot] : 0.f;
        LaneVerdict v;
        v.status = ST_NONE; v.dist = 0.f; v.U = 0.f; v.n = make_float3(0.f, 0.f, 0.f); v.have_n = 0;
        if (active) {
            if (MODE == MODE_DIST) v = decide_distance(tab, fl, GA, GB, T, out.min_dist[group]);
            else v = decide_boolean(tab, fl, (g & SR_PARKED) != 0 && !(fl & PAIR_NOCORE), GA, GB, T, padding, v0);
            if (v.status == ST_HIT) on_hit(p, group, slot);
            if (v.status == ST_VALUE) on_value(p, group, slot, fmaxf(v.dist - GA.core_radius - GB.core_radius, 0.f));
            DBG_ADD(0, 1); DBG_ADD(1 + v.status, 1);
            if (g & SR_PARKED) DBG_ADD(24, 1);
        }
        if (lane == 0) { const long long dtl = DBG_CLOCK() - tl0; DBG_ADD(11, dtl); DBG_MAX(is_list_chunk ? 43 : 44, dtl); }
        // what needs a whole warp goes to a list that k_surface_settle works off one item per warp: these items
        // cluster in a few bins (a finger inside link 5 ...), settled by the warp that found them they were the
        // tail of the launch
        const bool wants_warp = v.status == ST_WALK || v.status == ST_TRIANGLES;
        const unsigned wm = __ballot_sync(0xffffffffu, wants_warp);
        if (wm) {
            uint32_t first = 0;
            if (lane == 0) first = atomicAdd(bins.witem_count, (uint32_t)__popc(wm));
            first = __shfl_sync(0xffffffffu, first, 0);
            const uint32_t wi = first + __popc(wm & ((1u << lane) - 1u));
            if (wants_warp && wi < bins.witem_cap) {
                bins.witem[wi] = make_int2(p | (v.status << 24) | (v.have_n << 28), (int)(slot - (int64_t)p * bins.cap));
                bins.witem_v[wi] = make_float4(v.U, v.n.x, v.n.y, v.n.z);
                v.status = ST_NONE;
            }
            if (first + (uint32_t)__popc(wm) > bins.witem_cap)   // (list full: settled here, as k_small_check does)
                settle_warp_items<MODE>(M, tab, v, p, group, slot, T, padding, la, lb, (int)lane, on_hit, on_value, decided);
        }
    }
    if (lane == 0) { const long long dt = DBG_CLOCK() - tk0; DBG_ADD(14, dt); DBG_MAX(13, dt); }
}

// One warp per listed item: the walk over the enclosing core's planes, the triangle stage (pbp_surface.cuh).
// Tables are read through L1 / L2: a few thousand items per million states do not pay for staging them.
constexpr int SS_THREADS = 128;

template <int MODE>
__global__ void __launch_bounds__(SS_THREADS)
k_surface_settle(DevModel M, float padding, Outputs out, Bins bins, bool has_sample) {
    SurfTables tab;
    tab.hverts = M.verts; tab.hcells = M.sup_cells; tab.hlists = M.sup_lists; tab.planes = M.planes;
    tab.kverts = M.kverts; tab.kcells = M.ksup_cells; tab.klists = M.ksup_lists;
    tab.mverts = M.mverts; tab.tris = M.tris; tab.tverts = M.tverts;
    __shared__ uint16_t s_tri[SS_THREADS / 32][2][TRI_LIST_CAP];
    uint16_t *la = s_tri[threadIdx.x >> 5][0], *lb = s_tri[threadIdx.x >> 5][1];
    const unsigned lane = threadIdx.x & 31;
    const int64_t stride = (int64_t)M.npairs * bins.cap;
    auto on_hit = [&](int p, int32_t group, int64_t slot) {
        out.hit[group] = 1;
        if (MODE == MODE_BOOL_ALL) {
            if (out.first_pair) atomicMin(out.first_pair + group, M.pair_orig[p]);
            if (out.first_bad) atomicMin(out.first_bad + group, has_sample ? bins.sample[slot] : 0);
        }
    };
    auto on_value = [&](int, int32_t group, int64_t, float d) {
        if (d <= padding) out.hit[group] = 1;
        atomicMin(reinterpret_cast<unsigned int *>(out.min_dist + group), __float_as_uint(d));
    };
    auto decided = [&](int32_t group) { return MODE == MODE_BOOL && out.hit[group] != 0; };
    const uint32_t n = min(*bins.witem_count, bins.witem_cap);
    const uint32_t nwarps = gridDim.x * (SS_THREADS / 32);
    for (uint32_t i = blockIdx.x * (SS_THREADS / 32) + (threadIdx.x >> 5); i < n; i += nwarps) {
        const int2 it = bins.witem[i];
        const float4 w4 = bins.witem_v[i];
        const int p = it.x & 0xffffff;
        const int64_t slot = (int64_t)p * bins.cap + it.y;
        const int32_t group = bins.group[slot] & ~(SR_BITS | SR_NOENC);
        LaneVerdict v;
        v.status = lane == 0 ? ((it.x >> 24) & 0xf) : ST_NONE;     // lane 0 owns the item, the warp settles it
        v.dist = 0.f; v.U = w4.x; v.n = make_float3(w4.y, w4.z, w4.w); v.have_n = (it.x >> 28) & 1;
        float T[12];
#pragma unroll
        for (int k = 0; k < 12; k++) T[k] = bins.T[k * stride + slot];
        settle_warp_items<MODE>(M, tab, v, p, group, slot, T, padding, la, lb, (int)lane, on_hit, on_value, decided);
    }
}

// ------------------------------------------------------------------------------------------
// Small batches: one fused kernel, one thread per (pair, state)
// ------------------------------------------------------------------------------------------
// The pipeline is built for throughput; a single state or a single edge (the calls of a sequential RRT
// loop) pays four latency-bound launches plus their memsets.  For at most a few ten thousand
// (pair, state) items this kernel does everything in one launch: FK of the pair's two geometries along
// its joint path, the broadphase's proxy classification (same thresholds), and -- only where that is
// undecided -- the whole GJK on the cores; what that leaves open is settled as in k_surface_refine.
// Boolean verdicts only.
constexpr int SMALL_THREADS = 128;

// One (pair, state) item.  Returns ST_NONE: nothing to publish, ST_HIT: the pair collides, ST_WALK /
// ST_TRIANGLES ...: see LaneVerdict; R (placement of b in a's frame) is handed on to the warp.
__device__ __forceinline__ LaneVerdict small_item(const DevModel &M, const SurfTables &tab, const StateSource &src, int64_t n_states, float padding,
                                                  const Outputs &out, int64_t t, float *R, int32_t &group, int &p) {
    LaneVerdict none;
    none.status = ST_NONE; none.dist = 0.f; none.U = 0.f; none.n = make_float3(0.f, 0.f, 0.f); none.have_n = 0;
    LaneVerdict hit = none;
    hit.status = ST_HIT;
    p = (int)(t / n_states);                            // pair-major: the lanes of a warp share the pair
    const int64_t state = t - (int64_t)p * n_states;
    group = src.group ? src.group[state] : (int32_t)state;
    if (out.hit[group]) return none;                    // already decided by another item of this launch
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
    float cur[12], TA[12], TB[12], nxt[12];
    se3_identity(cur);
    int o = path.n_trunk;      // (the trunk cancels in the relative placement)
    for (int i = 0; i < path.n_a; i++, o++) {
        const DevJoint &J = M.joints[ops[o]];
        joint_transform(J, cur, qval(J.idx_q), nxt);
#pragma unroll
        for (int k = 0; k < 12; k++) cur[k] = nxt[k];
    }
    se3_mul(cur, GA.P, TA);
    se3_identity(cur);
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
    if (rs >= 0.f && d2 <= rs * rs) return hit;                          // certainly colliding
    if (exact || d2 > rf * rf) return none;                              // certainly free (or decided exactly)
    // ---- undecided: the whole GJK, on the cores.  (No inner-ball test here: with one pair per lane its cold
    //      table reads cost a single-state call 11 us and save a 1024-state call nothing.)
    const int fl = M.has_tris ? M.pair_enclose[p] : 0;
    const ShapeRef A = core_shape(tab, GA), B = core_shape(tab, GB);
    const float margin = padding + GA.core_radius + GB.core_radius;
    GjkState gs;
    GjkOut r;
    gjk_init(gs, gjk_start_direction(GA, GB, R));
    while (!gjk_step<false>(A, B, R, gs, margin, margin + pair_eps(GA, GB), r)) {
    }
    if (!(fl & PAIR_SURFACE)) return r.hit ? hit : none;
    if (r.far) return none;
    if (r.hit && !(fl & (PAIR_ENCLOSE | PAIR_NOCORE))) return hit;
    // (an item that converged between the bounds hands its direction on: one hull run, no second core run)
    const bool undecided = !r.hit && !(fl & PAIR_NOCORE);
    return decide_boolean(tab, fl, r.hit && !(fl & PAIR_NOCORE), GA, GB, R, padding, undecided ? gs.v : make_float3(0.f, 0.f, 0.f));
}

__global__ void __launch_bounds__(SMALL_THREADS)
k_small_check(DevModel M, StateSource src, int64_t n_states, float padding, Outputs out) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    SurfTables tab;
    tab.hverts = M.verts; tab.hcells = M.sup_cells; tab.hlists = M.sup_lists;
    tab.kverts = M.kverts; tab.kcells = M.ksup_cells; tab.klists = M.ksup_lists;
    tab.planes = M.planes; tab.mverts = M.mverts; tab.tris = M.tris; tab.tverts = M.tverts;
    float R[12];
#pragma unroll
    for (int k = 0; k < 12; k++) R[k] = 0.f;
    int32_t group = 0;
    int p = 0;
    LaneVerdict v;
    v.status = ST_NONE; v.dist = 0.f; v.U = 0.f; v.n = make_float3(0.f, 0.f, 0.f); v.have_n = 0;
    if (t < n_states * M.npairs) v = small_item(M, tab, src, n_states, padding, out, t, R, group, p);
    if (v.status == ST_HIT) out.hit[group] = 1;
    if (!M.has_tris) return;
    // surface semantics: the (rare) items that need the warp, one after the other
    __shared__ uint16_t s_tri[SMALL_THREADS / 32][2][TRI_LIST_CAP];
    auto on_hit = [&](int, int32_t g, int64_t) { out.hit[g] = 1; };
    auto on_value = [&](int, int32_t, int64_t, float) {};
    auto decided = [&](int32_t g) { return out.hit[g] != 0; };
    settle_warp_items<MODE_BOOL>(M, tab, v, p, group, (int64_t)0, R, padding, s_tri[threadIdx.x >> 5][0], s_tri[threadIdx.x >> 5][1],
                                 (int)(threadIdx.x & 31), on_hit, on_value, decided);
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

// libpbp_engine.so -- host side of the C ABI declared in include/pbp_engine.h.
//
// Owns the device copy of the compiled model, the per-pair item bins and the scratch used by
// the chunked broadphase -> narrowphase pipeline.  All launches go to the caller's stream.
#include <cub/device/device_select.cuh>

#include <algorithm>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "pbp_kernels.cuh"
#include "pbp_ik.cuh"
#include "pbp_knn.cuh"
#include "pbp_witness.cuh"
#include "pbp_jit.h"
#include "pbp_supmap.h"

static_assert(pbp::SUPMAP_KD == pbp::SUPMAP_K, "support map resolution mismatch");

using namespace pbp;

namespace {

thread_local std::string g_last_error;

int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

#define CUDA_TRY(expr)                                                                                  \
    do {                                                                                                \
        cudaError_t _e = (expr);                                                                        \
        if (_e != cudaSuccess)                                                                          \
            return fail(PBP_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

struct DeviceBuffer {
    void *p = nullptr;
    size_t bytes = 0;
    int ensure(size_t need) {
        if (need <= bytes) return 0;
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
        size_t want = std::max(need, (size_t)256);
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) return fail(PBP_ERR_ALLOC, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
        bytes = want;
        return 0;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
    }
    template <typename T> T *as() { return reinterpret_cast<T *>(p); }
};

}  // namespace

struct PinnedBuffer {
    void *p = nullptr;
    size_t bytes = 0;
    int ensure(size_t need) {
        if (need <= bytes) return 0;
        if (p) cudaFreeHost(p);
        p = nullptr;
        bytes = 0;
        cudaError_t e = cudaMallocHost(&p, need);
        if (e != cudaSuccess) return fail(PBP_ERR_ALLOC, "cudaMallocHost(%zu) failed: %s", need, cudaGetErrorString(e));
        bytes = need;
        return 0;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; bytes = 0; }
};

// true when `ptr` is page-locked host memory known to CUDA (async copies really are async)
static bool is_pinned_host(const void *ptr) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, ptr) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

struct pbp_engine {
    int device = 0;
    int num_sms = 0;
    int nq = 0, njoints = 0, ngeoms = 0, npairs = 0, nframes = 0, nmoving = 0, nverts_padded = 0;
    bool bp_smem_store = true;  // per-thread broadphase store fits in shared memory
    bool verts_in_smem = true;
    DevModel dm{};
    DeviceBuffer d_joints, d_jgeoms, d_geoms, d_bpgeoms, d_bpboxes, d_bppairs, d_groups, d_pairorig, d_paths, d_pathops, d_order, d_verts, d_supcells, d_suplists, d_qlim, d_frames, d_planes, d_plane_idx, d_surf_tol, d_pair_enclose, d_kverts, d_ksupcells, d_ksuplists, d_mverts, d_tris, d_tverts;
    bool surface = false;   // some pair carries an enclosure flag (surface semantics of mesh geometries)
    bool sr_tables_in_smem = false;
    size_t smem_sr = 0;
    int sr_blocks = 1;
    // pipeline workspaces
    int64_t chunk = 0;  // states per broadphase/narrowphase round = bin capacity
    DeviceBuffer d_bin_state, d_bin_group, d_bin_sample, d_bin_T, d_ctrl, d_doubt, d_doubt_v, d_witem, d_witem_v;
    DeviceBuffer d_counts, d_desc_group, d_desc_frac, d_desc_sample, d_scratch_i32, d_scratch_u8, d_cub, d_sel;
    DeviceBuffer d_io_q, d_io_q2, d_io_hit, d_io_f32, d_io_i32;   // staging for the *_host entry points
    unsigned long long *h_pinned = nullptr;  // small pinned mailbox for device -> host counters
    cudaStream_t s_copy = nullptr, s_comp = nullptr;   // *_host entry points: copy / compute pipeline
    std::vector<cudaEvent_t> io_events;
    PinnedBuffer h_io_hit, h_io_f32, h_io_i32;   // staging for results bound for pageable caller memory
    PinnedBuffer h_stage;                        // rotating float32 slices of a float64 host batch being converted
    int64_t host_slice = 1 << 18;                // states per slice of the *_host copy/compute pipeline
    // Two workspaces (halves of every bin + one control block each) let the *_host pipeline run
    // consecutive slices on two streams: the latency-bound tail of one slice's narrowphase
    // overlaps the next slice's broadphase.  ws selects the workspace run_round uses.
    // model-specialised broadphase (pbp_jit.h): module + function through the driver API
    CUmodule spec_module = nullptr;
    CUfunction spec_bp = nullptr;
    pbp::jit::SpecArgs spec_args{};
    float spec_thr_padding = -1.f;
    std::vector<float> spec_r_out, spec_r_in;      // per internal pair, for the per-call thresholds
    std::vector<float> spec_r_cap;                 // capsule-based "free" radius sum of the pair (< 0: none)
    std::vector<uint8_t> spec_exact;
    int64_t small_items = 400000;   // (pair, state) items up to which boolean queries use the fused k_small_check (measured crossover with the pipeline: ~8 k Panda states = 600 k items)
    PinnedBuffer h_desc;            // host-built sample descriptors of small edge batches
    int ws = 0;
    bool ws_split = false;
    cudaStream_t s_comp2 = nullptr;
    // float64 host copies of the kinematic tables for the IK kernel + the last uploaded chain
    std::vector<int32_t> h_joint_parent, h_joint_type, h_joint_idx_q, h_frame_parent;
    std::vector<double> h_joint_placement, h_joint_axis, h_frame_placement, h_q_lower, h_q_upper;
    DeviceBuffer d_ik_tables;
    DeviceBuffer d_ws_oMi, d_ws_oMg, d_ws_dist, d_ws_p1, d_ws_p2, d_ws_nrm, d_pairs_orig;   // distance / witness scratch
    pbp::IkTables ik_cached{}, ik_build{};
    bool ik_cached_valid = false;
    size_t smem_bp = 0, smem_is = 0, smem_np = 0, smem_np_dist = 0, smem_fk = 0;
    int np_blocks = 0, is_blocks = 0;
    bool verts_in_smem_dist = true;
    pbp_stats stats{};
    // One engine = one set of bins, counters and staging buffers: calls are serialised.  Host side by this
    // mutex (held for the whole entry point); device side by an event recorded at the end of every call
    // that the next call waits for when it arrives on a different stream.
    std::recursive_mutex mu;
    int guard_depth = 0;
    cudaEvent_t ev_last = nullptr;
    cudaStream_t last_stream = nullptr;
    bool last_valid = false;
    // CUDA graphs of whole boolean rounds (memsets + four kernels), keyed by the call's arguments: a replay
    // is ONE launch for the host thread.  Captured on s_cap (the caller's stream may be the legacy default
    // stream, which cannot be captured), launched into the caller's stream.
    struct RoundGraph {
        const float *q; uint8_t *hit; int64_t n; float padding;
        cudaGraphExec_t exec; int launches; uint64_t last_use;
    };
    std::vector<RoundGraph> graphs;
    uint64_t graph_clock = 0;
    bool graphs_enabled = true;
    cudaStream_t s_cap = nullptr;
    // optional per-kernel timing
    bool profiling = false;
    std::vector<cudaEvent_t> prof_events;   // five per round: before broadphase / item setup / narrowphase / surface refine, after
    std::vector<cudaEvent_t> prof_pool;     // recycled events
    pbp_profile prof{};
};

namespace {

// Every extern "C" entry point that touches an engine runs under one of these: serialises the host side,
// orders the device side against the previous call when the stream changed, switches to the engine's
// device and restores the caller's current device on the way out (torch reads it from the runtime).
struct CallGuard {
    pbp_engine *e;
    cudaStream_t st;
    int prev_dev = -1;
    CallGuard(pbp_engine *eng, cudaStream_t stream) : e(eng), st(stream) {
        e->mu.lock();
        if (e->guard_depth++ == 0) {
            cudaGetDevice(&prev_dev);
            if (prev_dev != e->device) cudaSetDevice(e->device);
            if (e->last_valid && e->last_stream != st && e->ev_last) cudaStreamWaitEvent(st, e->ev_last, 0);
        }
    }
    ~CallGuard() {
        if (--e->guard_depth == 0) {
            if (e->ev_last && cudaEventRecord(e->ev_last, st) == cudaSuccess) { e->last_stream = st; e->last_valid = true; }
            if (prev_dev >= 0 && prev_dev != e->device) cudaSetDevice(prev_dev);
        }
        e->mu.unlock();
    }
    CallGuard(const CallGuard &) = delete;
    CallGuard &operator=(const CallGuard &) = delete;
};

// control block layout (uint32 words): [0, npairs) bin counts, [npairs] item-setup tile counter,
// [npairs + 1] narrowphase chunk cursor, [npairs + 2, 2 npairs + 2) narrowphase bin counts (items the
// item setup kept), [2 npairs + 2] surface-refine chunk cursor, then four 64-bit counters at an 8-byte aligned offset
// (level total, emit cursor, select count, profiled item total, IK work cursor).
inline size_t ctrl_words(const pbp_engine *e) { return 2 * (size_t)e->npairs + 5; }   // + the surface-refine chunk cursor, the undecided-item count, the warp-item count
// There are two such blocks (workspaces 0 / 1, see pbp_engine::ws); the 64-bit counters live in block 0.
inline size_t ctrl_bytes(const pbp_engine *e) { return ((ctrl_words(e) * 4 + 7) & ~(size_t)7) + 64; }   // 8 x 64-bit counters
inline uint32_t *ctrl_base(pbp_engine *e) {
    return reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(e->d_ctrl.p) + (size_t)e->ws * ctrl_bytes(e));
}
inline uint32_t *ctrl_counts(pbp_engine *e) { return ctrl_base(e); }
inline uint32_t *ctrl_tile_setup(pbp_engine *e) { return ctrl_base(e) + e->npairs; }
inline uint32_t *ctrl_chunk_np(pbp_engine *e) { return ctrl_base(e) + e->npairs + 1; }
inline uint32_t *ctrl_counts_np(pbp_engine *e) { return ctrl_base(e) + e->npairs + 2; }
inline uint32_t *ctrl_chunk_sr(pbp_engine *e) { return ctrl_base(e) + 2 * e->npairs + 2; }
inline uint32_t *ctrl_doubt(pbp_engine *e) { return ctrl_base(e) + 2 * e->npairs + 3; }
inline uint32_t *ctrl_witem(pbp_engine *e) { return ctrl_base(e) + 2 * e->npairs + 4; }
inline unsigned long long *ctrl_u64(pbp_engine *e, int i) {
    const size_t off = (ctrl_words(e) * 4 + 7) & ~(size_t)7;
    return reinterpret_cast<unsigned long long *>(reinterpret_cast<char *>(e->d_ctrl.p) + off) + i;
}
// states one broadphase -> narrowphase round may take (half the bins each when two workspaces alternate)
inline int64_t round_cap(const pbp_engine *e) { return e->ws_split ? e->chunk / 2 : e->chunk; }

inline unsigned grid_for(int64_t n, int threads) { return (unsigned)((n + threads - 1) / threads); }

// Driver API entry points obtained through the runtime (no link-time dependency on libcuda, so the
// library still loads on a machine without a driver).
struct DriverApi {
    CUresult (*moduleLoadData)(CUmodule *, const void *) = nullptr;
    CUresult (*moduleUnload)(CUmodule) = nullptr;
    CUresult (*moduleGetFunction)(CUfunction *, CUmodule, const char *) = nullptr;
    CUresult (*launchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void **, void **) = nullptr;
    bool ok = false;
};
inline const DriverApi &driver_api() {
    static DriverApi api;
    static bool tried = false;
    if (!tried) {
        tried = true;
        cudaDriverEntryPointQueryResult qr;
        auto get = [&](const char *name, void **fn) {
            return cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess && *fn;
        };
        api.ok = get("cuModuleLoadData", (void **)&api.moduleLoadData) && get("cuModuleUnload", (void **)&api.moduleUnload) &&
                 get("cuModuleGetFunction", (void **)&api.moduleGetFunction) && get("cuLaunchKernel", (void **)&api.launchKernel);
        cudaGetLastError();
    }
    return api;
}

#define LAUNCH_CHECK(what)                                                                                  \
    do {                                                                                                    \
        cudaError_t _le = cudaGetLastError();                                                               \
        if (_le != cudaSuccess) return fail(PBP_ERR_CUDA, "%s launch failed: %s", what, cudaGetErrorString(_le)); \
        e->stats.kernel_launches++;                                                                         \
    } while (0)

template <int MODE>
int launch_round_kernels(pbp_engine *e, const StateSource &src, int64_t n, float padding, const Outputs &out, const Bins &bins,
                         cudaEvent_t *ev, cudaStream_t st) {
    const unsigned grid = grid_for(n, BP_THREADS);
    const bool has_sample = src.sample != nullptr;
    if (MODE == MODE_BOOL && e->spec_bp) {
        // model-specialised broadphase: the thresholds of this call's padding ride in the kernel parameter
        static_assert(pbp::jit::SPEC_THREADS == BP_THREADS, "block size mismatch");
        pbp::jit::SpecArgs &a = e->spec_args;
        a.q = src.q; a.q1 = src.q1; a.q2 = src.q2; a.group = src.group; a.frac = src.frac; a.mode = src.mode; a._pad = 0;
        a.n_states = n; a.hit = out.hit; a.bin_state = bins.state; a.bin_count = bins.count; a.cap = bins.cap;
        if (padding != e->spec_thr_padding) {
            for (int p = 0; p < e->npairs; p++) {   // same float arithmetic as the generic kernel's prologue
                const bool exact = e->spec_exact[p] != 0;
                const float slack = exact ? 0.f : BP_SLACK;
                const float rf = e->spec_r_out[p] + padding + slack, rs = e->spec_r_in[p] + padding - slack;
                const float y = rs >= 0.f ? rs * rs : -1.f;
                a.thr[p][0] = exact ? y : rf * rf;
                a.thr[p][1] = y;
                const float rc = e->spec_r_cap[p] + padding + BP_SLACK;
                a.thr[p][2] = e->spec_r_cap[p] >= 0.f ? rc * rc : INFINITY;
            }
            e->spec_thr_padding = padding;
        }
        void *params[] = {&a};
        const CUresult cr = driver_api().launchKernel(e->spec_bp, grid, 1, 1, BP_THREADS, 1, 1, 0, (CUstream)st, params, nullptr);
        if (cr != CUDA_SUCCESS) return fail(PBP_ERR_CUDA, "specialised broadphase launch failed (CUresult %d)", (int)cr);
        e->stats.kernel_launches++;
    } else if (e->bp_smem_store)
        k_broadphase<MODE, true><<<grid, BP_THREADS, e->smem_bp, st>>>(e->dm, src, n, padding, out, bins);
    else
        k_broadphase<MODE, false><<<grid, BP_THREADS, e->smem_bp, st>>>(e->dm, src, n, padding, out, bins);
    LAUNCH_CHECK("k_broadphase");
    if (ev) CUDA_TRY(cudaEventRecord(ev[1], st));
    // persistent kernels: never more blocks than the round can possibly feed (at most one item per
    // state and pair) -- a single-state call then stages the tables in a handful of blocks, not 444
    const int64_t max_items = n * (int64_t)e->npairs;
    const unsigned is_grid = (unsigned)std::min<int64_t>(e->is_blocks, std::max<int64_t>(1, (max_items + IS_THREADS - 1) / IS_THREADS));
    const unsigned np_grid = (unsigned)std::min<int64_t>(e->np_blocks, std::max<int64_t>(1, (max_items + NP_THREADS - 1) / NP_THREADS));
    k_item_setup<MODE><<<is_grid, IS_THREADS, e->smem_is, st>>>(e->dm, src, padding, out, bins, ctrl_tile_setup(e));
    LAUNCH_CHECK("k_item_setup");
    if (ev) CUDA_TRY(cudaEventRecord(ev[2], st));
    {
        const bool in_smem = MODE == MODE_DIST ? e->verts_in_smem_dist : e->verts_in_smem;
        const size_t smem = MODE == MODE_DIST ? e->smem_np_dist : e->smem_np;
        if (in_smem)
            k_narrowphase<MODE, true><<<np_grid, NP_THREADS, smem, st>>>(e->dm, padding, out, bins, ctrl_chunk_np(e), has_sample);
        else
            k_narrowphase<MODE, false><<<np_grid, NP_THREADS, smem, st>>>(e->dm, padding, out, bins, ctrl_chunk_np(e), has_sample);
    }
    LAUNCH_CHECK("k_narrowphase");
    if (ev) CUDA_TRY(cudaEventRecord(ev[3], st));
    if (e->surface) {
        // hull hits of pairs with an enclosure flag were parked by the narrowphase: decide them
        const unsigned sr_grid = (unsigned)std::min<int64_t>(e->sr_blocks, std::max<int64_t>(1, (max_items + 31) / 32 / (SR_THREADS / 32)));
        if (e->sr_tables_in_smem && n >= 65536)   // staging the tables only pays when the blocks have many items
            k_surface_refine<MODE, true><<<sr_grid, SR_THREADS, e->smem_sr, st>>>(e->dm, padding, out, bins, ctrl_chunk_sr(e), has_sample);
        else
            k_surface_refine<MODE, false><<<sr_grid, SR_THREADS, 0, st>>>(e->dm, padding, out, bins, ctrl_chunk_sr(e), has_sample);
        LAUNCH_CHECK("k_surface_refine");
        k_surface_settle<MODE><<<(unsigned)std::min<int64_t>(e->num_sms * 4, std::max<int64_t>(1, max_items / 4)), SS_THREADS, 0, st>>>(e->dm, padding, out, bins, has_sample);
        LAUNCH_CHECK("k_surface_settle");
    }
    return 0;
}

// One broadphase -> item setup -> narrowphase round over n <= chunk states.
int run_round(pbp_engine *e, int mode, const StateSource &src, int64_t n, float padding, const Outputs &out, cudaStream_t st) {
    if (n <= 0) return 0;
    if (e->npairs == 0) return 0;  // nothing can collide; outputs were initialised by the caller
    Bins bins;
    if (n > round_cap(e)) return fail(PBP_ERR_ARG, "internal: round of %lld states exceeds the bin capacity", (long long)n);
    const int64_t ws_off = e->ws_split ? (int64_t)e->ws * (e->chunk / 2) : 0;   // this workspace's half of every bin
    bins.state = e->d_bin_state.as<int32_t>() + ws_off;
    bins.group = e->d_bin_group.as<int32_t>() + ws_off;
    bins.sample = e->d_bin_sample.as<int32_t>() + ws_off;
    bins.T = e->d_bin_T.as<float>() + ws_off;
    bins.count = ctrl_counts(e);
    bins.count_np = ctrl_counts_np(e);
    bins.cap = e->chunk;
    bins.doubt = e->d_doubt.as<int2>() + ws_off;
    bins.doubt_v = e->d_doubt_v.as<float4>() + ws_off;
    bins.doubt_count = ctrl_doubt(e);
    bins.doubt_cap = (uint32_t)round_cap(e);
    bins.witem = e->d_witem.as<int2>() + ws_off;
    bins.witem_v = e->d_witem_v.as<float4>() + ws_off;
    bins.witem_count = ctrl_witem(e);
    bins.witem_cap = (uint32_t)round_cap(e);
    CUDA_TRY(cudaMemsetAsync(ctrl_base(e), 0, ctrl_words(e) * 4, st));
    int rc;
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    if (e->profiling) {
        for (int i = 0; i < 5; i++) {
            if (!e->prof_pool.empty()) { ev[i] = e->prof_pool.back(); e->prof_pool.pop_back(); }
            else CUDA_TRY(cudaEventCreate(&ev[i]));
        }
        CUDA_TRY(cudaEventRecord(ev[0], st));
    }
    cudaEvent_t *evp = e->profiling ? ev : nullptr;
    if (mode == MODE_BOOL) rc = launch_round_kernels<MODE_BOOL>(e, src, n, padding, out, bins, evp, st);
    else if (mode == MODE_BOOL_ALL) rc = launch_round_kernels<MODE_BOOL_ALL>(e, src, n, padding, out, bins, evp, st);
    else rc = launch_round_kernels<MODE_DIST>(e, src, n, padding, out, bins, evp, st);
    if (rc) return rc;
    if (e->profiling) {
        CUDA_TRY(cudaEventRecord(ev[4], st));
        for (int i = 0; i < 5; i++) e->prof_events.push_back(ev[i]);
        k_sum_counts<<<1, 256, 0, st>>>(bins.count_np, e->npairs, ctrl_u64(e, 3));
        e->stats.kernel_launches++;
    }
    e->stats.states += n;
    e->stats.chunks++;
    return 0;
}

int check_args(pbp_engine *e, const void *a, int64_t n) {
    if (!e) return fail(PBP_ERR_ARG, "engine is NULL");
    if (n < 0) return fail(PBP_ERR_ARG, "negative count");
    if (n > 0 && !a) return fail(PBP_ERR_ARG, "NULL buffer");
    // (group ids travel through the bins with three flag bits, SR_PARKED / SR_DOUBT / SR_NOENC, on top)
    if (n >= (int64_t)SR_NOENC) return fail(PBP_ERR_ARG, "more than 2^28-1 states/edges per call are not supported");
    return 0;
}

}  // namespace

#define PBP_HOST_FAIL fail
#include "pbp_host_tables.h"
static_assert(pbp::BP_MAX_SAVES_HOST == pbp::BP_MAX_SAVES, "broadphase save slots mismatch");

// CUDA source of the model-specialised broadphase (empty string when the model does not qualify).
static std::string specialised_broadphase_source(const pbp_model_desc *d, const HostTables &H) {
    if (d->npairs < 1 || d->npairs > pbp::jit::SPEC_MAX_PAIRS || d->njoints > 48 || d->ngeoms > 96) return std::string();
    std::vector<int32_t> gparent(std::max(d->ngeoms, 1), 0);
    for (int g = 0; g < d->ngeoms; g++) gparent[g] = H.geoms[g].parent;
    pbp::jit::Input in;
    in.nq = d->nq; in.njoints = d->njoints; in.ngeoms = d->ngeoms; in.npairs = d->npairs;
    in.joints = H.joints.data(); in.bpgeoms = H.bpgeoms.data(); in.geom_parent = gparent.data();
    in.bpboxes = H.bpboxes.data(); in.bppairs = H.bppairs.data(); in.groups = H.groups.data(); in.ngroups = (int)H.groups.size();
    return pbp::jit::generate_broadphase(in);
}

// ==========================================================================================
extern "C" {

const char *pbp_last_error(void) { return g_last_error.c_str(); }
int pbp_abi_version(void) { return PBP_ABI_VERSION; }
int pbp_engine_device(const pbp_engine *e) { return e ? e->device : -1; }

int pbp_get_stats(pbp_engine *e, pbp_stats *out, int reset) {
    if (!e || !out) return fail(PBP_ERR_ARG, "NULL argument");
    *out = e->stats;
    if (reset) e->stats = pbp_stats{};
    return 0;
}

#ifdef PBP_DEBUG_STATS
// developer build only (not in include/pbp_engine.h): the kernels' debug counters of the current device
int pbp_debug_stats(unsigned long long *out48, int reset) {
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaMemcpyFromSymbol(out48, pbp::g_dbg, sizeof(unsigned long long) * 48));
    if (reset) {
        unsigned long long z[48] = {0};
        CUDA_TRY(cudaMemcpyToSymbol(pbp::g_dbg, z, sizeof(z)));
    }
    return 0;
}
#endif

int pbp_set_profiling(pbp_engine *e, int enabled) {
    if (!e) return fail(PBP_ERR_ARG, "engine is NULL");
    e->profiling = enabled != 0;
    return 0;
}

int pbp_get_profile(pbp_engine *e, pbp_profile *out, int reset) {
    if (!e || !out) return fail(PBP_ERR_ARG, "NULL argument");
    CallGuard guard(e, nullptr);
    for (size_t i = 0; i + 4 < e->prof_events.size(); i += 5) {
        float a = 0.f, b = 0.f, c = 0.f, r = 0.f;
        CUDA_TRY(cudaEventSynchronize(e->prof_events[i + 4]));
        CUDA_TRY(cudaEventElapsedTime(&a, e->prof_events[i], e->prof_events[i + 1]));
        CUDA_TRY(cudaEventElapsedTime(&b, e->prof_events[i + 1], e->prof_events[i + 2]));
        CUDA_TRY(cudaEventElapsedTime(&c, e->prof_events[i + 2], e->prof_events[i + 3]));
        CUDA_TRY(cudaEventElapsedTime(&r, e->prof_events[i + 3], e->prof_events[i + 4]));
        e->prof.broadphase_ms += a;
        e->prof.item_setup_ms += b;
        e->prof.narrowphase_ms += c;
        e->prof.surface_refine_ms += r;
        e->prof.rounds++;
        for (int k = 0; k < 5; k++) e->prof_pool.push_back(e->prof_events[i + k]);
    }
    e->prof_events.clear();
    unsigned long long items = 0;
    CUDA_TRY(cudaMemcpy(&items, ctrl_u64(e, 3), 8, cudaMemcpyDeviceToHost));
    e->prof.narrowphase_items = (int64_t)items;
    *out = e->prof;
    if (reset) {
        e->prof = pbp_profile{};
        CUDA_TRY(cudaMemset(ctrl_u64(e, 3), 0, 8));
    }
    return 0;
}

void pbp_engine_destroy(pbp_engine *e) {
    if (!e) return;
    cudaSetDevice(e->device);
    for (cudaEvent_t ev : e->prof_events) cudaEventDestroy(ev);
    for (cudaEvent_t ev : e->prof_pool) cudaEventDestroy(ev);
    DeviceBuffer *bufs[] = {&e->d_joints, &e->d_jgeoms, &e->d_geoms, &e->d_bpgeoms, &e->d_bpboxes, &e->d_bppairs, &e->d_groups, &e->d_pairorig, &e->d_paths,
                            &e->d_pathops, &e->d_order, &e->d_verts, &e->d_supcells, &e->d_suplists, &e->d_qlim, &e->d_frames, &e->d_bin_state,
                            &e->d_bin_group, &e->d_bin_sample, &e->d_bin_T, &e->d_ctrl, &e->d_doubt, &e->d_doubt_v, &e->d_witem, &e->d_witem_v, &e->d_counts, &e->d_desc_group,
                            &e->d_desc_frac, &e->d_desc_sample, &e->d_scratch_i32, &e->d_scratch_u8, &e->d_cub, &e->d_sel,
                            &e->d_io_q, &e->d_io_q2, &e->d_io_hit, &e->d_io_f32, &e->d_io_i32, &e->d_ik_tables, &e->d_ws_oMi, &e->d_ws_oMg, &e->d_ws_dist,
                            &e->d_ws_p1, &e->d_ws_p2, &e->d_ws_nrm, &e->d_pairs_orig, &e->d_planes, &e->d_plane_idx, &e->d_surf_tol, &e->d_pair_enclose, &e->d_kverts, &e->d_ksupcells,
                            &e->d_ksuplists, &e->d_mverts, &e->d_tris, &e->d_tverts};
    for (DeviceBuffer *b : bufs) b->release();
    if (e->h_pinned) cudaFreeHost(e->h_pinned);
    e->h_io_hit.release(); e->h_io_f32.release(); e->h_io_i32.release(); e->h_stage.release();
    if (e->spec_module && driver_api().ok) driver_api().moduleUnload(e->spec_module);
    e->h_desc.release();
    for (cudaEvent_t ev : e->io_events) cudaEventDestroy(ev);
    if (e->s_copy) cudaStreamDestroy(e->s_copy);
    if (e->s_comp) cudaStreamDestroy(e->s_comp);
    if (e->s_comp2) cudaStreamDestroy(e->s_comp2);
    for (auto &g : e->graphs) cudaGraphExecDestroy(g.exec);
    if (e->s_cap) cudaStreamDestroy(e->s_cap);
    if (e->ev_last) cudaEventDestroy(e->ev_last);
    delete e;
}

int pbp_engine_create(const pbp_model_desc *d, int device, int64_t max_chunk_states, pbp_engine **out) {
    if (!d || !out) return fail(PBP_ERR_ARG, "NULL argument");
    *out = nullptr;
    if (d->njoints < 1 || d->njoints > PBP_MAX_JOINTS) return fail(PBP_ERR_CAPACITY, "njoints %d outside [1, %d]", d->njoints, PBP_MAX_JOINTS);
    if (d->ngeoms < 0 || d->ngeoms > PBP_MAX_GEOMS) return fail(PBP_ERR_CAPACITY, "ngeoms %d outside [0, %d]", d->ngeoms, PBP_MAX_GEOMS);
    if (d->npairs < 0 || d->npairs > PBP_MAX_PAIRS) return fail(PBP_ERR_CAPACITY, "npairs %d outside [0, %d]", d->npairs, PBP_MAX_PAIRS);
    if (d->nq < 1 || d->nq > PBP_MAX_NQ) return fail(PBP_ERR_CAPACITY, "nq %d outside [1, %d]", d->nq, PBP_MAX_NQ);
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0)
        return fail(PBP_ERR_CUDA, "no CUDA device available (%s); this engine has no CPU fallback", cudaGetErrorString(ce));
    if (device < 0 || device >= ndev) return fail(PBP_ERR_ARG, "device %d out of range [0, %d)", device, ndev);
    struct DeviceRestore {   // the caller's current device survives this call (torch reads it from the runtime)
        int prev = -1;
        DeviceRestore() { cudaGetDevice(&prev); }
        ~DeviceRestore() { if (prev >= 0) cudaSetDevice(prev); }
    } restore;
    CUDA_TRY(cudaSetDevice(device));

    pbp_engine *e = new pbp_engine();
    e->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete e; return fail(PBP_ERR_CUDA, "cudaGetDeviceProperties failed"); }
    e->num_sms = prop.multiProcessorCount;
    e->nq = d->nq; e->njoints = d->njoints; e->ngeoms = d->ngeoms; e->npairs = d->npairs; e->nframes = d->nframes;
    {
        auto widen = [](std::vector<double> &dst, const double *f64, const float *f32, size_t cnt) {
            dst.resize(cnt);
            for (size_t k = 0; k < cnt; k++) dst[k] = f64 ? f64[k] : (double)f32[k];
        };
        e->h_joint_parent.assign(d->joint_parent, d->joint_parent + d->njoints);
        e->h_joint_type.assign(d->joint_type, d->joint_type + d->njoints);
        e->h_joint_idx_q.assign(d->joint_idx_q, d->joint_idx_q + d->njoints);
        widen(e->h_joint_placement, d->joint_placement_f64, d->joint_placement, 12 * (size_t)d->njoints);
        widen(e->h_joint_axis, d->joint_axis_f64, d->joint_axis, 3 * (size_t)d->njoints);
        widen(e->h_q_lower, d->q_lower_f64, d->q_lower, (size_t)d->nq);
        widen(e->h_q_upper, d->q_upper_f64, d->q_upper, (size_t)d->nq);
        if (d->nframes > 0) {
            e->h_frame_parent.assign(d->frame_parent, d->frame_parent + d->nframes);
            widen(e->h_frame_placement, d->frame_placement_f64, d->frame_placement, 12 * (size_t)d->nframes);
        }
    }

    auto bail = [&](int code) { pbp_engine_destroy(e); return code; };
#define CUDA_TRY_BAIL(expr)                                                                             \
    do {                                                                                                \
        cudaError_t _e = (expr);                                                                        \
        if (_e != cudaSuccess)                                                                          \
            return bail(fail(PBP_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__)); \
    } while (0)

    HostTables H;
    {
        const int hrc = build_host_tables(d, H);
        if (hrc) return bail(hrc);
    }
    auto &joints = H.joints; auto &geoms = H.geoms; auto &bpgeoms = H.bpgeoms; auto &bpboxes = H.bpboxes; auto &verts = H.verts;
    auto &sup_cells = H.sup_cells; auto &sup_lists = H.sup_lists; auto &jgeoms = H.jgeoms; auto &bppairs = H.bppairs;
    auto &groups = H.groups; auto &pair_orig = H.pair_orig; auto &paths = H.paths; auto &pathops = H.pathops; auto &order = H.order;
    const int nmoving = H.nmoving, nsave = H.nsave;
    const int np1 = std::max(d->npairs, 1);
    e->nmoving = nmoving;
    e->nverts_padded = (int)verts.size();

    auto upload = [&](DeviceBuffer &buf, const void *src, size_t bytes) -> int {
        int rc = buf.ensure(bytes);
        if (rc) return rc;
        cudaError_t er = cudaMemcpy(buf.p, src, bytes, cudaMemcpyHostToDevice);
        if (er != cudaSuccess) return fail(PBP_ERR_CUDA, "model upload failed: %s", cudaGetErrorString(er));
        return 0;
    };
    std::vector<float> qlim(2 * d->nq);
    for (int i = 0; i < d->nq; i++) { qlim[i] = d->q_lower ? d->q_lower[i] : 0.f; qlim[d->nq + i] = d->q_upper ? d->q_upper[i] : 0.f; }
    std::vector<int32_t> fparent(std::max(d->nframes, 1), 0);
    for (int f = 0; f < d->nframes; f++) {
        fparent[f] = d->frame_parent[f];
        if (fparent[f] < 0 || fparent[f] >= d->njoints) return bail(fail(PBP_ERR_ARG, "frame %d: bad parent joint", f));
    }
    int rc = 0;
    if ((rc = upload(e->d_joints, joints.data(), joints.size() * sizeof(DevJoint))) ||
        (rc = upload(e->d_jgeoms, jgeoms.data(), jgeoms.size() * 4)) ||
        (rc = upload(e->d_geoms, geoms.data(), geoms.size() * sizeof(DevGeom))) ||
        (rc = upload(e->d_bpgeoms, bpgeoms.data(), bpgeoms.size() * sizeof(BpGeom))) ||
        (rc = upload(e->d_bpboxes, bpboxes.data(), bpboxes.size() * sizeof(BpBox))) ||
        (rc = upload(e->d_bppairs, bppairs.data(), bppairs.size() * sizeof(BpPair))) ||
        (rc = upload(e->d_groups, groups.data(), groups.size() * sizeof(BpGroup))) ||
        (rc = upload(e->d_pairorig, pair_orig.data(), pair_orig.size() * 4)) ||
        (rc = upload(e->d_paths, paths.data(), paths.size() * sizeof(PairPath))) ||
        (rc = upload(e->d_pathops, pathops.data(), pathops.size())) ||
        (rc = upload(e->d_order, order.data(), order.size() * sizeof(int32_t))) ||
        (rc = upload(e->d_verts, verts.data(), verts.size() * sizeof(float4))) ||
        (rc = upload(e->d_supcells, sup_cells.data(), sup_cells.size() * 2)) ||
        (rc = upload(e->d_suplists, sup_lists.data(), sup_lists.size() * 2)) ||
        (rc = upload(e->d_qlim, qlim.data(), qlim.size() * sizeof(float))) ||
        (rc = upload(e->d_pairs_orig, d->pairs, std::max<size_t>((size_t)d->npairs, 1) * 8)))
        return bail(rc);
    {
        // frames: [nframes] int32 parents followed (16-byte aligned) by [nframes][12] placements
        const size_t poff = (((size_t)std::max(d->nframes, 1) * 4) + 15) & ~(size_t)15;
        std::vector<unsigned char> blob(poff + (size_t)std::max(d->nframes, 1) * 48, 0);
        memcpy(blob.data(), fparent.data(), (size_t)std::max(d->nframes, 1) * 4);
        if (d->nframes > 0) memcpy(blob.data() + poff, d->frame_placement, (size_t)d->nframes * 48);
        if ((rc = upload(e->d_frames, blob.data(), blob.size()))) return bail(rc);
        e->dm.frame_parent = e->d_frames.as<int32_t>();
        e->dm.frame_placement = reinterpret_cast<const float *>(reinterpret_cast<char *>(e->d_frames.p) + poff);
    }
    e->dm.nq = d->nq; e->dm.njoints = d->njoints; e->dm.ngeoms = d->ngeoms; e->dm.npairs = d->npairs;
    e->dm.nverts_padded = e->nverts_padded; e->dm.nframes = d->nframes; e->dm.nmoving = nmoving; e->dm.nsave = nsave;
    e->dm.nboxes = (int)bpboxes.size();
    e->dm.joints = e->d_joints.as<DevJoint>();
    e->dm.joint_geoms = e->d_jgeoms.as<int32_t>();
    e->dm.geoms = e->d_geoms.as<DevGeom>();
    e->dm.bpgeoms = e->d_bpgeoms.as<BpGeom>();
    e->dm.bpboxes = e->d_bpboxes.as<BpBox>();
    e->dm.bppairs = e->d_bppairs.as<BpPair>();
    e->dm.groups = e->d_groups.as<BpGroup>();
    e->dm.pair_orig = e->d_pairorig.as<int32_t>();
    e->dm.ngroups = (int)groups.size();
    e->dm.paths = e->d_paths.as<PairPath>();
    e->dm.path_ops = e->d_pathops.as<uint8_t>();
    e->dm.bin_order = e->d_order.as<int32_t>();
    e->dm.verts = e->d_verts.as<float4>();
    e->dm.sup_cells = e->d_supcells.as<uint16_t>();
    e->dm.sup_lists = e->d_suplists.as<uint16_t>();
    e->dm.n_sup_cells = (int)sup_cells.size();
    e->dm.n_sup_lists = (int)sup_lists.size();
    {
        // surface semantics tables (optional): planes as float4, per-geometry (offset, count), tolerance,
        // pair flags in INTERNAL pair order, inner cores, mesh vertices and triangles
        std::vector<float> pl(4, 0.f), tol(std::max(d->ngeoms, 1), 0.f);
        std::vector<int32_t> pidx(2 * (size_t)std::max(d->ngeoms, 1), 0);
        const bool have = d->planes && d->geom_plane_off && d->geom_plane_cnt && d->nplanes > 0;
        if (have) {
            pl.assign(d->planes, d->planes + 4 * (size_t)d->nplanes);
            for (int g = 0; g < d->ngeoms; g++) {
                pidx[2 * g] = d->geom_plane_off[g];
                pidx[2 * g + 1] = d->geom_plane_cnt[g];
                if (pidx[2 * g] < 0 || pidx[2 * g + 1] < 0 || pidx[2 * g] + pidx[2 * g + 1] > d->nplanes)
                    return bail(fail(PBP_ERR_ARG, "geometry %d: plane range outside [0, %d)", g, d->nplanes));
                if (d->geom_surface_tol) tol[g] = d->geom_surface_tol[g];
            }
        }
        for (int p = 0; p < d->npairs; p++)
            if (H.pair_flags[p]) e->surface = true;
        if ((rc = upload(e->d_planes, pl.data(), pl.size() * 4)) || (rc = upload(e->d_plane_idx, pidx.data(), pidx.size() * 4)) ||
            (rc = upload(e->d_surf_tol, tol.data(), tol.size() * 4)) || (rc = upload(e->d_pair_enclose, H.pair_flags.data(), H.pair_flags.size())))
            return bail(rc);
        e->dm.planes = e->d_planes.as<float4>();
        e->dm.plane_idx = e->d_plane_idx.as<int2>();
        e->dm.surface_tol = e->d_surf_tol.as<float>();
        e->dm.pair_enclose = e->d_pair_enclose.as<uint8_t>();
        e->dm.surface = e->surface ? 1 : 0;
        e->dm.nplanes = (int)(pl.size() / 4);
        e->dm.has_tris = H.has_tris ? 1 : 0;
        if (H.has_tris) {
            if ((rc = upload(e->d_kverts, H.kverts.data(), H.kverts.size() * sizeof(float4))) ||
                (rc = upload(e->d_ksupcells, H.ksup_cells.data(), H.ksup_cells.size() * 2)) ||
                (rc = upload(e->d_ksuplists, H.ksup_lists.data(), H.ksup_lists.size() * 2)) ||
                (rc = upload(e->d_mverts, H.mverts.data(), H.mverts.size() * sizeof(float4))) ||
                (rc = upload(e->d_tris, H.tris.data(), H.tris.size() * sizeof(int4))) ||
                (rc = upload(e->d_tverts, H.tverts.data(), H.tverts.size() * sizeof(float4))))
                return bail(rc);
            e->dm.kverts = e->d_kverts.as<float4>();
            e->dm.ksup_cells = e->d_ksupcells.as<uint16_t>();
            e->dm.ksup_lists = e->d_ksuplists.as<uint16_t>();
            e->dm.nkverts_padded = (int)H.kverts.size();
            e->dm.n_ksup_cells = (int)H.ksup_cells.size();
            e->dm.n_ksup_lists = (int)H.ksup_lists.size();
            e->dm.mverts = e->d_mverts.as<float4>();
            e->dm.tris = e->d_tris.as<int4>();
            e->dm.tverts = e->d_tverts.as<float4>();
        } else {
            e->dm.kverts = e->dm.verts;
            e->dm.ksup_cells = e->dm.sup_cells;
            e->dm.ksup_lists = e->dm.sup_lists;
            e->dm.nkverts_padded = e->dm.nverts_padded;
            e->dm.n_ksup_cells = e->dm.n_sup_cells;
            e->dm.n_ksup_lists = e->dm.n_sup_lists;
            e->dm.mverts = e->dm.verts;
            e->dm.tris = nullptr;
            e->dm.tverts = nullptr;
        }
    }
    e->dm.q_lower = e->d_qlim.as<float>();
    e->dm.q_upper = e->d_qlim.as<float>() + d->nq;

    // ---- shared memory budgets and launch geometry
    const size_t max_optin = prop.sharedMemPerBlockOptin;
    const size_t bp_tables = bp_table_bytes(d->njoints, nmoving, d->ngeoms, e->dm.nboxes, d->npairs, (int)groups.size(), d->nq);
    const size_t bp_store = (size_t)(nmoving * 3 + nsave * 12) * BP_THREADS * 4;
    e->bp_smem_store = bp_tables + bp_store <= 100 * 1024 && nmoving <= BP_MAX_LOCAL_CENTRES;
    e->smem_bp = bp_tables + (e->bp_smem_store ? bp_store : 0) + 16;
    e->smem_fk = align16((size_t)d->njoints * sizeof(DevJoint)) + align16((size_t)np1 * 0 + (size_t)std::max(d->ngeoms, 1) * sizeof(DevGeom));
    const size_t tile_tables = align16(((size_t)d->npairs + 1) * 4) + align16((size_t)np1 * 4);
    e->smem_is = align16((size_t)d->njoints * sizeof(DevJoint)) + align16((size_t)std::max(d->ngeoms, 1) * sizeof(DevGeom)) +
                 align16((size_t)np1 * sizeof(BpPair)) + align16((size_t)np1 * sizeof(PairPath)) + tile_tables + 16;
    const size_t np_tables = align16((size_t)std::max(d->ngeoms, 1) * sizeof(DevGeom)) + align16((size_t)np1 * sizeof(BpPair)) + tile_tables + 16;
    const size_t vert_bytes = align16((size_t)e->nverts_padded * sizeof(float4)) + align16(sup_cells.size() * 2) + align16(sup_lists.size() * 2);
    // boolean queries stage the inner cores, distance queries the hulls (the same tables when no geometry is a surface)
    const size_t kvert_bytes = H.has_tris ? align16(H.kverts.size() * sizeof(float4)) + align16(H.ksup_cells.size() * 2) + align16(H.ksup_lists.size() * 2)
                                          : vert_bytes;
    e->verts_in_smem = np_tables + kvert_bytes <= 110 * 1024;   // two narrowphase blocks per SM must fit in 227 KB
    e->smem_np = np_tables + (e->verts_in_smem ? kvert_bytes : 0);
    e->verts_in_smem_dist = np_tables + vert_bytes <= 110 * 1024;
    e->smem_np_dist = np_tables + (e->verts_in_smem_dist ? vert_bytes : 0);
    if (e->smem_bp > max_optin || e->smem_np > max_optin || e->smem_np_dist > max_optin || e->smem_is > max_optin)
        return bail(fail(PBP_ERR_CAPACITY, "model tables exceed shared memory (%zu / %zu / %zu bytes)", e->smem_bp, e->smem_is, e->smem_np));
    // The dynamic shared-memory cap is a per-function attribute shared by every engine of the
    // process: raise it to the device's opt-in maximum once (a per-engine value would lower
    // the cap under another engine's feet).
#define SET_SMEM(kern) CUDA_TRY_BAIL(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_optin))
    SET_SMEM((k_broadphase<MODE_BOOL, true>));
    SET_SMEM((k_broadphase<MODE_BOOL_ALL, true>));
    SET_SMEM((k_broadphase<MODE_DIST, true>));
    SET_SMEM((k_broadphase<MODE_BOOL, false>));
    SET_SMEM((k_broadphase<MODE_BOOL_ALL, false>));
    SET_SMEM((k_broadphase<MODE_DIST, false>));
    SET_SMEM((k_item_setup<MODE_BOOL>));
    SET_SMEM((k_item_setup<MODE_BOOL_ALL>));
    SET_SMEM((k_item_setup<MODE_DIST>));
    SET_SMEM((k_narrowphase<MODE_BOOL, true>));
    SET_SMEM((k_narrowphase<MODE_BOOL_ALL, true>));
    SET_SMEM((k_narrowphase<MODE_DIST, true>));
    SET_SMEM((k_narrowphase<MODE_BOOL, false>));
    SET_SMEM((k_narrowphase<MODE_BOOL_ALL, false>));
    SET_SMEM((k_narrowphase<MODE_DIST, false>));
    SET_SMEM(k_fk_output);
#undef SET_SMEM
    // (static __shared__ chunk directory and per-warp triangle lists: leave room for them under the opt-in cap)
    const int sr_static = 32 * 1024;
#define SET_SMEM_SR(kern) CUDA_TRY_BAIL(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_optin - sr_static))
    SET_SMEM_SR((k_surface_refine<MODE_BOOL, true>));
    SET_SMEM_SR((k_surface_refine<MODE_BOOL_ALL, true>));
    SET_SMEM_SR((k_surface_refine<MODE_DIST, true>));
#undef SET_SMEM_SR
    e->smem_sr = vert_bytes + align16((size_t)e->dm.nplanes * sizeof(float4));
    e->sr_tables_in_smem = e->smem_sr + sr_static <= max_optin;   // one 16-warp block per SM is enough for this kernel
    if (const char *v = getenv("PBP_SR_SMEM")) e->sr_tables_in_smem = e->sr_tables_in_smem && atoi(v) != 0;   // developer switch
    {
        int occ_sr = 1;
        if (e->sr_tables_in_smem) CUDA_TRY_BAIL(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_sr, k_surface_refine<MODE_BOOL, true>, SR_THREADS, e->smem_sr));
        else CUDA_TRY_BAIL(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_sr, k_surface_refine<MODE_BOOL, false>, SR_THREADS, 0));
        e->sr_blocks = e->num_sms * std::max(occ_sr, 1);
    }
    int occ = 1;
    if (e->verts_in_smem)
        CUDA_TRY_BAIL(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_narrowphase<MODE_BOOL, true>, NP_THREADS, e->smem_np));
    else
        CUDA_TRY_BAIL(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_narrowphase<MODE_BOOL, false>, NP_THREADS, e->smem_np));
    e->np_blocks = e->num_sms * std::max(occ, 1);
    CUDA_TRY_BAIL(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_item_setup<MODE_BOOL>, IS_THREADS, e->smem_is));
    e->is_blocks = e->num_sms * std::max(occ, 1);

    // ---- model-specialised broadphase (optional: any failure leaves the generic kernel in charge)
    if (!getenv("PBP_NO_JIT") && driver_api().ok) {
        const std::string src_text = specialised_broadphase_source(d, H);
        if (!src_text.empty()) {
            std::vector<char> cubin;
            std::string log;
            if (pbp::jit::compile(src_text, prop.major, prop.minor, cubin, log) == 0 &&
                driver_api().moduleLoadData(&e->spec_module, cubin.data()) == CUDA_SUCCESS &&
                driver_api().moduleGetFunction(&e->spec_bp, e->spec_module, "k_bp_spec") == CUDA_SUCCESS) {
                e->spec_r_out.resize(np1); e->spec_r_in.resize(np1); e->spec_exact.resize(np1); e->spec_r_cap.assign(np1, -1.f);
                for (int p = 0; p < d->npairs; p++) {
                    e->spec_r_out[p] = bppairs[p].r_out;
                    e->spec_r_in[p] = bppairs[p].r_in;
                    e->spec_exact[p] = bppairs[p].kind == BP_EXACT_SPHERES;
                    if (bppairs[p].kind == BP_SPHERES) {   // must mirror the generator's choice (pbp_jit.h)
                        const BpGeom &ga = bpgeoms[bppairs[p].a], &gb = bpgeoms[bppairs[p].b];
                        auto usable = [](const BpGeom &g) {
                            const float du[3] = {g.cap1[0] - g.cap0[0], g.cap1[1] - g.cap0[1], g.cap1[2] - g.cap0[2]};
                            return g.has_cap && (du[0] * du[0] + du[1] * du[1] + du[2] * du[2]) > 0.f;
                        };
                        const bool ca = usable(ga), cb = usable(gb);
                        if (ca && cb) e->spec_r_cap[p] = ga.cap_r + gb.cap_r;
                        else if (ca) e->spec_r_cap[p] = ga.cap_r + gb.r_out;
                        else if (cb) e->spec_r_cap[p] = gb.cap_r + ga.r_out;
                    }
                }
            } else {
                e->spec_bp = nullptr;
                if (getenv("PBP_JIT_VERBOSE")) fprintf(stderr, "pbp: specialised broadphase unavailable: %s\n", log.c_str());
            }
        }
    }

    // ---- bins: one slot per (state of a chunk, pair)
    const size_t item_bytes = 4 + 4 + 4 + 48;
    int64_t chunk = max_chunk_states > 0 ? max_chunk_states : (int64_t)1 << 20;
    const size_t budget = (size_t)6 << 30;
    if (d->npairs > 0) {
        const int64_t fit = (int64_t)(budget / (item_bytes * (size_t)d->npairs));
        chunk = std::max<int64_t>(std::min(chunk, fit), 1024);
    }
    chunk = (chunk + 255) & ~(int64_t)255;
    e->chunk = chunk;
    const size_t slots = (size_t)np1 * (size_t)chunk;
    if ((rc = e->d_bin_state.ensure(slots * 4)) || (rc = e->d_bin_group.ensure(slots * 4)) || (rc = e->d_bin_sample.ensure(slots * 4)) ||
        (rc = e->d_bin_T.ensure(slots * 48)) || (rc = e->d_ctrl.ensure(2 * ctrl_bytes(e))) || (rc = e->d_doubt.ensure((size_t)chunk * sizeof(int2))) ||
        (rc = e->d_doubt_v.ensure((size_t)chunk * sizeof(float4))) || (rc = e->d_witem.ensure((size_t)chunk * sizeof(int2))) ||
        (rc = e->d_witem_v.ensure((size_t)chunk * sizeof(float4))))
        return bail(rc);
    CUDA_TRY_BAIL(cudaMemset(e->d_ctrl.p, 0, 2 * ctrl_bytes(e)));
    CUDA_TRY_BAIL(cudaMallocHost(reinterpret_cast<void **>(&e->h_pinned), 64));
    if (const char *sm = getenv("PBP_SMALL_ITEMS")) e->small_items = atoll(sm);
    if (const char *hs = getenv("PBP_HOST_SLICE")) {
        const long long v = atoll(hs);
        if (v >= 1024) e->host_slice = v;
    }
    CUDA_TRY_BAIL(cudaStreamCreateWithFlags(&e->s_copy, cudaStreamNonBlocking));
    CUDA_TRY_BAIL(cudaStreamCreateWithFlags(&e->s_comp, cudaStreamNonBlocking));
    CUDA_TRY_BAIL(cudaStreamCreateWithFlags(&e->s_comp2, cudaStreamNonBlocking));
    CUDA_TRY_BAIL(cudaStreamCreateWithFlags(&e->s_cap, cudaStreamNonBlocking));
    CUDA_TRY_BAIL(cudaEventCreateWithFlags(&e->ev_last, cudaEventDisableTiming));
    if (const char *g = getenv("PBP_GRAPHS")) e->graphs_enabled = atoi(g) != 0;
    *out = e;
    return 0;
}

#undef CUDA_TRY_BAIL
// ------------------------------------------------------------------------------------------
int pbp_fk(pbp_engine *e, const float *q_dev, int64_t n, float *oMi_dev, float *oMg_dev, float *oMf_dev, void *stream) {
    int rc = check_args(e, q_dev, n);
    if (rc) return rc;
    if (n == 0) return 0;
    CallGuard guard(e, (cudaStream_t)stream);
    cudaStream_t st = (cudaStream_t)stream;
    k_fk_output<<<grid_for(n, 128), 128, e->smem_fk, st>>>(e->dm, q_dev, n, oMi_dev, oMg_dev, oMf_dev);
    CUDA_TRY(cudaGetLastError());
    e->stats.kernel_launches++;
    return 0;
}

int pbp_check_states(pbp_engine *e, const float *q_dev, int64_t n, float padding, uint8_t *hit_dev, float *min_dist_dev,
                     int32_t *first_pair_dev, void *stream) {
    int rc = check_args(e, q_dev, n);
    if (rc) return rc;
    if (n > 0 && !hit_dev) return fail(PBP_ERR_ARG, "hit buffer is NULL");
    if (!(padding >= 0.0f)) return fail(PBP_ERR_ARG, "distance padding must be >= 0 (got %g)", (double)padding);
    if (n == 0) return 0;
    CallGuard guard(e, (cudaStream_t)stream);
    cudaStream_t st = (cudaStream_t)stream;
    // A whole boolean round (two memsets + four kernels) as ONE launch: the graph captured for these
    // arguments is replayed; a planner or a benchmark that calls with the same buffers again and again
    // (and eight ranks sharing the host's cores) then spends one launch per call on the host thread.
    const bool graphable = !min_dist_dev && !first_pair_dev && e->graphs_enabled && !e->profiling && !e->ws_split && e->ws == 0 &&
                           e->npairs > 0 && n <= round_cap(e) && n * (int64_t)e->npairs > e->small_items;
    if (graphable) {
        for (auto &g : e->graphs) {
            if (g.q == q_dev && g.hit == hit_dev && g.n == n && g.padding == padding) {
                CUDA_TRY(cudaGraphLaunch(g.exec, st));
                g.last_use = ++e->graph_clock;
                e->stats.kernel_launches += g.launches;
                e->stats.states += n;
                e->stats.chunks++;
                return 0;
            }
        }
        if (cudaStreamBeginCapture(e->s_cap, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
            const int64_t launches0 = e->stats.kernel_launches;
            StateSource src{};
            src.q = q_dev;
            src.mode = 0;
            Outputs out{};
            out.hit = hit_dev;
            cudaError_t ce = cudaMemsetAsync(hit_dev, 0, (size_t)n, e->s_cap);
            rc = ce == cudaSuccess ? run_round(e, MODE_BOOL, src, n, padding, out, e->s_cap) : fail(PBP_ERR_CUDA, "memset in capture: %s", cudaGetErrorString(ce));
            cudaGraph_t graph = nullptr;
            ce = cudaStreamEndCapture(e->s_cap, &graph);
            cudaGraphExec_t exec = nullptr;
            if (rc == 0 && ce == cudaSuccess && graph && cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess) {
                cudaGraphDestroy(graph);
                if (e->graphs.size() >= 16) {   // evict the least recently used
                    size_t lru = 0;
                    for (size_t i = 1; i < e->graphs.size(); i++)
                        if (e->graphs[i].last_use < e->graphs[lru].last_use) lru = i;
                    cudaGraphExecDestroy(e->graphs[lru].exec);
                    e->graphs.erase(e->graphs.begin() + (long)lru);
                }
                e->graphs.push_back(pbp_engine::RoundGraph{q_dev, hit_dev, n, padding, exec, (int)(e->stats.kernel_launches - launches0), ++e->graph_clock});
                CUDA_TRY(cudaGraphLaunch(exec, st));
                return 0;
            }
            // capture or instantiation failed: fall back to plain launches from now on
            if (graph) cudaGraphDestroy(graph);
            cudaGetLastError();
            e->graphs_enabled = false;
            e->stats.kernel_launches = launches0;
            e->stats.states -= n;
            e->stats.chunks--;
            if (rc) return rc;
        } else {
            cudaGetLastError();
            e->graphs_enabled = false;
        }
    }
    CUDA_TRY(cudaMemsetAsync(hit_dev, 0, (size_t)n, st));
    if (min_dist_dev) {
        k_fill_f32<<<grid_for(n, 256), 256, 0, st>>>(min_dist_dev, n, INFINITY);
        e->stats.kernel_launches++;
    }
    if (first_pair_dev) {
        k_fill_i32<<<grid_for(n, 256), 256, 0, st>>>(first_pair_dev, n, INT_MAX);
        e->stats.kernel_launches++;
    }
    // min distance wins over first_pair for the mode; both requested -> two passes
    const int mode = min_dist_dev ? MODE_DIST : (first_pair_dev ? MODE_BOOL_ALL : MODE_BOOL);
    if (mode == MODE_BOOL && e->npairs > 0 && n * (int64_t)e->npairs <= e->small_items) {
        // small batch: FK + proxies + GJK fused in one launch
        StateSource src{};
        src.q = q_dev;
        src.mode = 0;
        Outputs out{};
        out.hit = hit_dev;
        k_small_check<<<grid_for(n * e->npairs, SMALL_THREADS), SMALL_THREADS, 0, st>>>(e->dm, src, n, padding, out);
        LAUNCH_CHECK("k_small_check");
        e->stats.states += n;
        return 0;
    }
    for (int64_t off = 0; off < n; off += round_cap(e)) {
        const int64_t cnt = std::min(round_cap(e), n - off);
        StateSource src{};
        src.q = q_dev + off * e->nq;
        src.mode = 0;
        Outputs out{};
        out.hit = hit_dev + off;
        out.min_dist = min_dist_dev ? min_dist_dev + off : nullptr;
        out.first_pair = (mode == MODE_BOOL_ALL) ? first_pair_dev + off : nullptr;
        if ((rc = run_round(e, mode, src, cnt, padding, out, st))) return rc;
        if (min_dist_dev && first_pair_dev) {
            // second pass for the pair index (scratch hit buffer: verdicts are already final)
            if ((rc = e->d_scratch_u8.ensure((size_t)e->chunk))) return rc;
            Outputs o2{};
            o2.hit = e->d_scratch_u8.as<uint8_t>();
            o2.first_pair = first_pair_dev + off;
            if ((rc = run_round(e, MODE_BOOL_ALL, src, cnt, padding, o2, st))) return rc;
        }
    }
    if (first_pair_dev) {
        k_finalize_index<<<grid_for(n, 256), 256, 0, st>>>(first_pair_dev, n);
        e->stats.kernel_launches++;
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int pbp_check_paths(pbp_engine *e, const float *q_dev, const int64_t *path_offsets_dev, int64_t n_paths, int64_t total,
                    float padding, uint8_t *hit_dev, void *stream) {
    int rc = check_args(e, hit_dev, n_paths);
    if (rc) return rc;
    if ((rc = check_args(e, q_dev, total))) return rc;
    if (!(padding >= 0.0f)) return fail(PBP_ERR_ARG, "distance padding must be >= 0");
    if (n_paths == 0) return 0;
    if (!path_offsets_dev) return fail(PBP_ERR_ARG, "path_offsets is NULL");
    CallGuard guard(e, (cudaStream_t)stream);
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(cudaMemsetAsync(hit_dev, 0, (size_t)n_paths, st));
    if (total == 0) return 0;
    if ((rc = e->d_desc_group.ensure((size_t)total * 4))) return rc;
    k_path_groups<<<grid_for(total, 256), 256, 0, st>>>(path_offsets_dev, n_paths, total, e->d_desc_group.as<int32_t>(), nullptr);
    e->stats.kernel_launches++;
    if (e->npairs > 0 && total * (int64_t)e->npairs <= e->small_items) {   // few samples: the fused latency kernel
        StateSource src{};
        src.q = q_dev;
        src.group = e->d_desc_group.as<int32_t>();
        src.mode = 0;
        Outputs out{};
        out.hit = hit_dev;
        k_small_check<<<grid_for(total * e->npairs, SMALL_THREADS), SMALL_THREADS, 0, st>>>(e->dm, src, total, padding, out);
        LAUNCH_CHECK("k_small_check");
        e->stats.states += total;
        return 0;
    }
    for (int64_t off = 0; off < total; off += e->chunk) {
        const int64_t cnt = std::min(e->chunk, total - off);
        StateSource src{};
        src.q = q_dev + off * e->nq;
        src.group = e->d_desc_group.as<int32_t>() + off;
        src.mode = 0;
        Outputs out{};
        out.hit = hit_dev;
        if ((rc = run_round(e, MODE_BOOL, src, cnt, padding, out, st))) return rc;
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int pbp_discretize_counts(pbp_engine *e, const float *q1_dev, const float *q2_dev, int64_t n_edges, double step,
                          int32_t *counts_dev, void *stream) {
    int rc = check_args(e, q1_dev, n_edges);
    if (rc) return rc;
    if (n_edges > 0 && (!q2_dev || !counts_dev)) return fail(PBP_ERR_ARG, "NULL buffer");
    if (!(step > 0.0)) return fail(PBP_ERR_ARG, "step must be > 0");
    if (n_edges == 0) return 0;
    CallGuard guard(e, (cudaStream_t)stream);
    k_edge_counts<<<grid_for(n_edges, 256), 256, 0, (cudaStream_t)stream>>>(q1_dev, q2_dev, n_edges, e->nq, step, counts_dev);
    CUDA_TRY(cudaGetLastError());
    e->stats.kernel_launches++;
    return 0;
}

int pbp_discretize_fill(pbp_engine *e, const float *q1_dev, const float *q2_dev, int64_t n_edges, const int32_t *counts_dev,
                        const int64_t *offsets_dev, float *q_out_dev, void *stream) {
    int rc = check_args(e, q1_dev, n_edges);
    if (rc) return rc;
    if (n_edges > 0 && (!q2_dev || !counts_dev || !offsets_dev || !q_out_dev)) return fail(PBP_ERR_ARG, "NULL buffer");
    if (n_edges == 0) return 0;
    CallGuard guard(e, (cudaStream_t)stream);
    k_edge_fill<<<grid_for(n_edges * 32, 256), 256, 0, (cudaStream_t)stream>>>(q1_dev, q2_dev, n_edges, e->nq, counts_dev, offsets_dev, q_out_dev);
    CUDA_TRY(cudaGetLastError());
    e->stats.kernel_launches++;
    return 0;
}

// known_total >= 0: the caller already knows the number of samples of all edges together (host
// entry point with few edges): the single-level path then needs no device -> host round trip.
static int check_edges_core(pbp_engine *e, const float *q1_dev, const float *q2_dev, int64_t n_edges, double step, float padding,
                            uint8_t *hit_dev, int32_t *first_bad_dev, void *stream, int64_t known_total,
                            const int32_t *host_counts = nullptr) {
    int rc = check_args(e, q1_dev, n_edges);
    if (host_counts && known_total >= 0 && !first_bad_dev && e->npairs > 0 && n_edges > 0 && q2_dev && hit_dev && step > 0.0 &&
        padding >= 0.0f && known_total * (int64_t)e->npairs <= e->small_items) {
        // few samples in total: descriptors (edge, fraction) built here, one fused launch
        // (no count / emit kernels, no refinement levels, no device -> host round trip)
        if (rc) return rc;
        CallGuard guard(e, (cudaStream_t)stream);
        cudaStream_t st = (cudaStream_t)stream;
        CUDA_TRY(cudaMemsetAsync(hit_dev, 0, (size_t)n_edges, st));
        if (known_total == 0) return 0;
        const size_t tot = (size_t)known_total;
        if ((rc = e->h_desc.ensure(tot * 8)) || (rc = e->d_desc_group.ensure(tot * 4)) || (rc = e->d_desc_frac.ensure(tot * 4))) return rc;
        int32_t *hg = (int32_t *)e->h_desc.p;
        float *hf = (float *)((char *)e->h_desc.p + tot * 4);
        size_t o = 0;
        for (int64_t ed = 0; ed < n_edges; ed++) {
            const int nn = host_counts[ed];
            const float inv = nn > 1 ? 1.0f / (float)(nn - 1) : 0.f;       // the same expression as k_level_emit
            for (int i = 0; i < nn; i++, o++) {
                hg[o] = (int32_t)ed;
                hf[o] = (i == nn - 1 && nn > 1) ? 1.0f : (float)i * inv;
            }
        }
        CUDA_TRY(cudaMemcpyAsync(e->d_desc_group.p, hg, tot * 4, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemcpyAsync(e->d_desc_frac.p, hf, tot * 4, cudaMemcpyHostToDevice, st));
        StateSource src{};
        src.q1 = q1_dev; src.q2 = q2_dev;
        src.group = e->d_desc_group.as<int32_t>(); src.frac = e->d_desc_frac.as<float>();
        src.mode = 1;
        Outputs out{};
        out.hit = hit_dev;
        k_small_check<<<grid_for((int64_t)tot * e->npairs, SMALL_THREADS), SMALL_THREADS, 0, st>>>(e->dm, src, (int64_t)tot, padding, out);
        LAUNCH_CHECK("k_small_check");
        e->stats.states += (int64_t)tot;
        return 0;
    }
    if (rc) return rc;
    if (n_edges > 0 && (!q2_dev || !hit_dev)) return fail(PBP_ERR_ARG, "NULL buffer");
    if (!(step > 0.0)) return fail(PBP_ERR_ARG, "step must be > 0");
    if (!(padding >= 0.0f)) return fail(PBP_ERR_ARG, "distance padding must be >= 0");
    if (n_edges == 0) return 0;
    CallGuard guard(e, (cudaStream_t)stream);
    cudaStream_t st = (cudaStream_t)stream;
    if ((rc = e->d_counts.ensure((size_t)n_edges * 4))) return rc;
    int32_t *counts = e->d_counts.as<int32_t>();
    k_edge_counts<<<grid_for(n_edges, 256), 256, 0, st>>>(q1_dev, q2_dev, n_edges, e->nq, step, counts);
    CUDA_TRY(cudaGetLastError());
    e->stats.kernel_launches++;
    CUDA_TRY(cudaMemsetAsync(hit_dev, 0, (size_t)n_edges, st));
    if (first_bad_dev) {
        k_fill_i32<<<grid_for(n_edges, 256), 256, 0, st>>>(first_bad_dev, n_edges, INT_MAX);
        e->stats.kernel_launches++;
    }
    if (e->npairs == 0) {
        if (first_bad_dev) {
            k_finalize_index<<<grid_for(n_edges, 256), 256, 0, st>>>(first_bad_dev, n_edges);
            e->stats.kernel_launches++;
        }
        return 0;
    }

    // Refinement levels.  With first_bad every sample is tested (one level, stride 1).  Every level
    // costs a device -> host round trip (the sample total sizes the next launch), so small batches --
    // the single-edge calls of a sequential planner -- use one level, medium ones two.
    std::vector<LevelSpec> levels;
    if (first_bad_dev || n_edges <= 512) {
        levels.push_back(LevelSpec{1, 0, 1});
    } else if (n_edges <= 16384) {
        levels.push_back(LevelSpec{16, 0, 1});
        levels.push_back(LevelSpec{1, 16, 0});
    } else if (const char *lv = getenv("PBP_EDGE_LEVELS")) {
        // developer knob: comma-separated strides, coarse to fine, each dividing the previous, ending in 1
        int prev = 0;
        for (const char *c = lv; *c;) {
            const int s = atoi(c);
            if (s < 1) break;
            levels.push_back(LevelSpec{s, prev, prev == 0 ? 1 : 0});
            prev = s;
            while (*c && *c != ',') c++;
            if (*c == ',') c++;
        }
        if (levels.empty() || levels.back().stride != 1) return fail(PBP_ERR_ARG, "PBP_EDGE_LEVELS must end with stride 1");
    } else {
        // large batches: every 32nd sample (+ the end point), then every 4th, then the rest.  Measured
        // on 100 k random Panda edges: the full 128, 64, ..., 1 hierarchy evaluates the fewest samples
        // (1.35 M) but pays eight latency-bound rounds: 1.09 ms; {32, 4, 1} evaluates 1.63 M in 0.61 ms.
        levels.push_back(LevelSpec{32, 0, 1});
        levels.push_back(LevelSpec{4, 32, 0});
        levels.push_back(LevelSpec{1, 4, 0});
    }
    const int mode = first_bad_dev ? MODE_BOOL_ALL : MODE_BOOL;
    unsigned long long *d_total = ctrl_u64(e, 0), *d_cursor = ctrl_u64(e, 1);
    for (const LevelSpec &lv : levels) {
        CUDA_TRY(cudaMemsetAsync(d_total, 0, 16, st));
        const uint8_t *dead = (first_bad_dev || levels.size() == 1) ? nullptr : hit_dev;
        int64_t total;
        if (known_total >= 0 && levels.size() == 1) {
            total = known_total;
        } else {
            k_level_count<<<grid_for(n_edges, 256), 256, 0, st>>>(counts, dead, n_edges, lv, d_total);
            CUDA_TRY(cudaGetLastError());
            e->stats.kernel_launches++;
            CUDA_TRY(cudaMemcpyAsync(e->h_pinned, d_total, 8, cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaStreamSynchronize(st));
            total = (int64_t)e->h_pinned[0];
        }
        if (total == 0) continue;
        const size_t slack = known_total >= 0 ? (size_t)n_edges : 0;   // host-computed total: belt and braces
        if ((rc = e->d_desc_group.ensure((total + slack) * 4)) || (rc = e->d_desc_frac.ensure((total + slack) * 4))) return rc;
        if (first_bad_dev && (rc = e->d_desc_sample.ensure((total + slack) * 4))) return rc;
        int32_t *dg = e->d_desc_group.as<int32_t>();
        float *df = e->d_desc_frac.as<float>();
        int32_t *ds = first_bad_dev ? e->d_desc_sample.as<int32_t>() : nullptr;
        if (known_total >= 0) {
            // every descriptor slot holds a valid sample (edge 0, fraction 0) even if the device's
            // counts disagreed with the host's; the caller compares the emit cursor afterwards
            CUDA_TRY(cudaMemsetAsync(dg, 0, (total + slack) * 4, st));
            CUDA_TRY(cudaMemsetAsync(df, 0, (total + slack) * 4, st));
            if (ds) CUDA_TRY(cudaMemsetAsync(ds, 0, (total + slack) * 4, st));
        }
        k_level_emit<<<grid_for(n_edges, 256), 256, 0, st>>>(counts, dead, n_edges, lv, d_cursor, dg, df, ds);
        CUDA_TRY(cudaGetLastError());
        e->stats.kernel_launches++;
        for (int64_t off = 0; off < total; off += e->chunk) {
            const int64_t cnt = std::min(e->chunk, total - off);
            StateSource src{};
            src.q1 = q1_dev; src.q2 = q2_dev;
            src.group = dg + off; src.frac = df + off; src.sample = ds ? ds + off : nullptr;
            src.mode = 1;
            Outputs out{};
            out.hit = hit_dev;
            out.first_bad = first_bad_dev;
            if ((rc = run_round(e, mode, src, cnt, padding, out, st))) return rc;
        }
    }
    if (first_bad_dev) {
        k_finalize_index<<<grid_for(n_edges, 256), 256, 0, st>>>(first_bad_dev, n_edges);
        e->stats.kernel_launches++;
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int pbp_check_edges(pbp_engine *e, const float *q1_dev, const float *q2_dev, int64_t n_edges, double step, float padding,
                    uint8_t *hit_dev, int32_t *first_bad_dev, void *stream) {
    return check_edges_core(e, q1_dev, q2_dev, n_edges, step, padding, hit_dev, first_bad_dev, stream, -1);
}

// ------------------------------------------------------------------------------------------
// host-buffer entry points: H2D, run, D2H, synchronise
// ------------------------------------------------------------------------------------------
// Host buffers in, host buffers out.  The batch is cut into slices; every slice's H2D copy is
// queued on a copy stream up front and its FK/broadphase/narrowphase rounds + result D2H follow on
// a compute stream as soon as that slice has landed, so PCIe transfers overlap the kernels
// (with pinned host memory; pageable memory still works, just without overlap).
int pbp_check_states_host(pbp_engine *e, const float *q_host, int64_t n, float padding, uint8_t *hit_host,
                          float *min_dist_host, int32_t *first_pair_host) {
    int rc = check_args(e, q_host, n);
    if (rc) return rc;
    if (n > 0 && !hit_host) return fail(PBP_ERR_ARG, "hit buffer is NULL");
    if (n == 0) return 0;
    CallGuard guard(e, e->s_comp);
    if ((rc = e->d_io_q.ensure((size_t)n * e->nq * 4)) || (rc = e->d_io_hit.ensure((size_t)n))) return rc;
    if (min_dist_host && (rc = e->d_io_f32.ensure((size_t)n * 4))) return rc;
    if (first_pair_host && (rc = e->d_io_i32.ensure((size_t)n * 4))) return rc;
    // Results bound for pageable memory are staged in the engine's pinned buffers: a device ->
    // pageable cudaMemcpyAsync blocks the host thread and would stall the launch pipeline.
    uint8_t *hh = hit_host;
    float *hd = min_dist_host;
    int32_t *hp = first_pair_host;
    if (!is_pinned_host(hit_host)) { if ((rc = e->h_io_hit.ensure((size_t)n))) return rc; hh = (uint8_t *)e->h_io_hit.p; }
    if (min_dist_host && !is_pinned_host(min_dist_host)) { if ((rc = e->h_io_f32.ensure((size_t)n * 4))) return rc; hd = (float *)e->h_io_f32.p; }
    if (first_pair_host && !is_pinned_host(first_pair_host)) { if ((rc = e->h_io_i32.ensure((size_t)n * 4))) return rc; hp = (int32_t *)e->h_io_i32.p; }
    // Slices small enough that the compute of the last one (all that is left once the final H2D
    // copy lands) is short, large enough to fill the 148 SMs for a few waves.
    const int64_t slice = e->host_slice;
    const int64_t n_slices = (n + slice - 1) / slice;
    while ((int64_t)e->io_events.size() < n_slices) {
        cudaEvent_t ev;
        CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        e->io_events.push_back(ev);
    }
    float *dq = e->d_io_q.as<float>();
    for (int64_t k = 0; k < n_slices; k++) {
        const int64_t off = k * slice, cnt = std::min(slice, n - off);
        CUDA_TRY(cudaMemcpyAsync(dq + off * e->nq, q_host + off * e->nq, (size_t)cnt * e->nq * 4, cudaMemcpyHostToDevice, e->s_copy));
        CUDA_TRY(cudaEventRecord(e->io_events[k], e->s_copy));
    }
    // consecutive slices alternate between the two workspaces / compute streams (not when both
    // the distance and the pair index are wanted: that second pass shares one scratch buffer)
    struct SplitGuard {
        pbp_engine *e;
        ~SplitGuard() { e->ws_split = false; e->ws = 0; }
    } split_guard{e};
    e->ws_split = n_slices > 1 && slice <= e->chunk / 2 && !(min_dist_host && first_pair_host);
    for (int64_t k = 0; k < n_slices; k++) {
        const int64_t off = k * slice, cnt = std::min(slice, n - off);
        e->ws = e->ws_split ? (int)(k & 1) : 0;
        cudaStream_t sc = e->ws ? e->s_comp2 : e->s_comp;
        CUDA_TRY(cudaStreamWaitEvent(sc, e->io_events[k], 0));
        uint8_t *dh = e->d_io_hit.as<uint8_t>() + off;
        float *dd = min_dist_host ? e->d_io_f32.as<float>() + off : nullptr;
        int32_t *dp = first_pair_host ? e->d_io_i32.as<int32_t>() + off : nullptr;
        if ((rc = pbp_check_states(e, dq + off * e->nq, cnt, padding, dh, dd, dp, sc))) return rc;
        CUDA_TRY(cudaMemcpyAsync(hh + off, dh, (size_t)cnt, cudaMemcpyDeviceToHost, sc));
        if (dd) CUDA_TRY(cudaMemcpyAsync(hd + off, dd, (size_t)cnt * 4, cudaMemcpyDeviceToHost, sc));
        if (dp) CUDA_TRY(cudaMemcpyAsync(hp + off, dp, (size_t)cnt * 4, cudaMemcpyDeviceToHost, sc));
    }
    CUDA_TRY(cudaStreamSynchronize(e->s_comp));
    if (e->ws_split) CUDA_TRY(cudaStreamSynchronize(e->s_comp2));
    if (hh != hit_host) memcpy(hit_host, hh, (size_t)n);
    if (min_dist_host && hd != min_dist_host) memcpy(min_dist_host, hd, (size_t)n * 4);
    if (first_pair_host && hp != first_pair_host) memcpy(first_pair_host, hp, (size_t)n * 4);
    return 0;
}

// dst[i] = (float)src[i].  The destination is a pinned staging buffer that only the DMA engine reads next:
// streaming stores keep it out of the caches and spare the read-for-ownership (a quarter of the memory traffic
// of this loop, which is bound by the host's memory bandwidth, not by the conversion).
static void narrow_f64_to_f32(float *dst, const double *src, int64_t m) {
    int64_t i = 0;
#if defined(__SSE2__)
    if ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
        for (; i + 4 <= m; i += 4) {
            const __m128 lo = _mm_cvtpd_ps(_mm_loadu_pd(src + i)), hi = _mm_cvtpd_ps(_mm_loadu_pd(src + i + 2));
            _mm_stream_ps(dst + i, _mm_movelh_ps(lo, hi));
        }
        _mm_sfence();
    }
#endif
    for (; i < m; i++) dst[i] = (float)src[i];
}

// float64 host states (what the reference's callers hold: numpy float64, pageable).  The batch is cut into
// slices; a few host threads convert slice after slice into rotating pinned float32 buffers while earlier
// slices are already crossing PCIe and being checked, so conversion, copies and kernels overlap.
int pbp_check_states_host_f64(pbp_engine *e, const double *q_host, int64_t n, float padding, uint8_t *hit_host,
                              float *min_dist_host, int32_t *first_pair_host) {
    int rc = check_args(e, q_host, n);
    if (rc) return rc;
    if (n > 0 && !hit_host) return fail(PBP_ERR_ARG, "hit buffer is NULL");
    if (n == 0) return 0;
    CallGuard guard(e, e->s_comp);
    const int nq = e->nq;
    if ((rc = e->d_io_q.ensure((size_t)n * nq * 4)) || (rc = e->d_io_hit.ensure((size_t)n))) return rc;
    if (min_dist_host && (rc = e->d_io_f32.ensure((size_t)n * 4))) return rc;
    if (first_pair_host && (rc = e->d_io_i32.ensure((size_t)n * 4))) return rc;
    uint8_t *hh = hit_host;
    float *hd = min_dist_host;
    int32_t *hp = first_pair_host;
    if (!is_pinned_host(hit_host)) { if ((rc = e->h_io_hit.ensure((size_t)n))) return rc; hh = (uint8_t *)e->h_io_hit.p; }
    if (min_dist_host && !is_pinned_host(min_dist_host)) { if ((rc = e->h_io_f32.ensure((size_t)n * 4))) return rc; hd = (float *)e->h_io_f32.p; }
    if (first_pair_host && !is_pinned_host(first_pair_host)) { if ((rc = e->h_io_i32.ensure((size_t)n * 4))) return rc; hp = (int32_t *)e->h_io_i32.p; }
    // Compute slices as in the float32 path; every slice is narrowed in SUB-CHUNKS claimed by the host threads
    // (one thread narrows ~5 GB/s of doubles: a whole slice per thread would leave the pipeline waiting for the
    // slowest of a few threads), into a ring of NBUF pinned slice buffers.
    const int64_t slice = e->host_slice;
    const int64_t n_slices = (n + slice - 1) / slice;
    constexpr int NBUF = 4;
    constexpr int64_t SUB = 1 << 15;                       // states per conversion chunk (2.4 MB of doubles)
    const int64_t subs_per_slice = (slice + SUB - 1) / SUB;
    if ((rc = e->h_stage.ensure((size_t)NBUF * slice * nq * 4))) return rc;
    while ((int64_t)e->io_events.size() < n_slices) {
        cudaEvent_t ev;
        CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        e->io_events.push_back(ev);
    }
    const int64_t n_subs = n_slices * subs_per_slice;
    int nthreads = (int)std::min<int64_t>(std::min<int64_t>(n_subs, 12), std::max(1u, std::thread::hardware_concurrency()));
    if (const char *ht = getenv("PBP_HOST_THREADS")) nthreads = std::max(1, std::min(atoi(ht), 64));
    std::vector<std::atomic<int>> ready((size_t)n_slices), issued((size_t)n_slices);   // ready: sub-chunks of the slice narrowed so far
    for (int64_t k = 0; k < n_slices; k++) { ready[(size_t)k].store(0); issued[(size_t)k].store(0); }
    std::atomic<int64_t> next{0};
    std::atomic<int> abort_flag{0};
    float *stage = (float *)e->h_stage.p;
    const int device = e->device;
    auto worker = [&]() {
        bool device_set = false;   // (binding a fresh thread to the CUDA context costs a fraction of a millisecond:
                                   //  only threads that have to wait for a copy pay it)
        for (;;) {
            const int64_t j = next.fetch_add(1);
            if (j >= n_subs) return;
            const int64_t k = j / subs_per_slice, sub = j - k * subs_per_slice;
            if (k >= NBUF) {   // the buffer's previous slice must have left the host
                while (!issued[(size_t)(k - NBUF)].load(std::memory_order_acquire))
                    if (abort_flag.load()) return; else std::this_thread::yield();
                if (!device_set) { cudaSetDevice(device); device_set = true; }
                cudaEventSynchronize(e->io_events[(size_t)(k - NBUF)]);
            }
            const int64_t off = k * slice, cnt = std::min(slice, n - off);
            const int64_t s0 = std::min(sub * SUB, cnt), s1 = std::min(s0 + SUB, cnt);
            const double *src = q_host + (off + s0) * nq;
            float *dst = stage + (k % NBUF) * slice * nq + s0 * nq;
            const int64_t m = (s1 - s0) * nq;
            narrow_f64_to_f32(dst, src, m);
            ready[(size_t)k].fetch_add(1, std::memory_order_release);
        }
    };
    std::vector<std::thread> pool;
    for (int t = 0; t < nthreads; t++) pool.emplace_back(worker);
    struct Join {
        std::vector<std::thread> &pool; std::atomic<int> &abort_flag; pbp_engine *e;
        ~Join() { abort_flag.store(1); for (auto &t : pool) if (t.joinable()) t.join(); e->ws_split = false; e->ws = 0; }
    } join{pool, abort_flag, e};
    float *dq = e->d_io_q.as<float>();
    e->ws_split = n_slices > 1 && slice <= e->chunk / 2 && !(min_dist_host && first_pair_host);
    auto step = [&](int64_t k) -> int {
        const int64_t off = k * slice, cnt = std::min(slice, n - off);
        while (ready[(size_t)k].load(std::memory_order_acquire) < subs_per_slice) std::this_thread::yield();
        CUDA_TRY(cudaMemcpyAsync(dq + off * nq, stage + (k % NBUF) * slice * nq, (size_t)cnt * nq * 4, cudaMemcpyHostToDevice, e->s_copy));
        CUDA_TRY(cudaEventRecord(e->io_events[(size_t)k], e->s_copy));
        issued[(size_t)k].store(1, std::memory_order_release);
        e->ws = e->ws_split ? (int)(k & 1) : 0;
        cudaStream_t sc = e->ws ? e->s_comp2 : e->s_comp;
        CUDA_TRY(cudaStreamWaitEvent(sc, e->io_events[(size_t)k], 0));
        uint8_t *dh = e->d_io_hit.as<uint8_t>() + off;
        float *dd = min_dist_host ? e->d_io_f32.as<float>() + off : nullptr;
        int32_t *dp = first_pair_host ? e->d_io_i32.as<int32_t>() + off : nullptr;
        int r = pbp_check_states(e, dq + off * nq, cnt, padding, dh, dd, dp, sc);
        if (r) return r;
        CUDA_TRY(cudaMemcpyAsync(hh + off, dh, (size_t)cnt, cudaMemcpyDeviceToHost, sc));
        if (dd) CUDA_TRY(cudaMemcpyAsync(hd + off, dd, (size_t)cnt * 4, cudaMemcpyDeviceToHost, sc));
        if (dp) CUDA_TRY(cudaMemcpyAsync(hp + off, dp, (size_t)cnt * 4, cudaMemcpyDeviceToHost, sc));
        return 0;
    };
    for (int64_t k = 0; k < n_slices; k++)
        if ((rc = step(k))) return rc;
    CUDA_TRY(cudaStreamSynchronize(e->s_comp));
    CUDA_TRY(cudaStreamSynchronize(e->s_comp2));
    if (hh != hit_host) memcpy(hit_host, hh, (size_t)n);
    if (min_dist_host && hd != min_dist_host) memcpy(min_dist_host, hd, (size_t)n * 4);
    if (first_pair_host && hp != first_pair_host) memcpy(first_pair_host, hp, (size_t)n * 4);
    return 0;
}

int pbp_check_edges_host(pbp_engine *e, const float *q1_host, const float *q2_host, int64_t n_edges, double step,
                         float padding, uint8_t *hit_host, int32_t *first_bad_host) {
    int rc = check_args(e, q1_host, n_edges);
    if (rc) return rc;
    if (n_edges > 0 && (!q2_host || !hit_host)) return fail(PBP_ERR_ARG, "NULL buffer");
    if (n_edges == 0) return 0;
    CallGuard guard(e, e->s_comp);
    const size_t qb = (size_t)n_edges * e->nq * 4;
    if ((rc = e->d_io_q.ensure(qb)) || (rc = e->d_io_q2.ensure(qb)) || (rc = e->d_io_hit.ensure((size_t)n_edges))) return rc;
    if (first_bad_host && (rc = e->d_io_i32.ensure((size_t)n_edges * 4))) return rc;
    cudaStream_t st = e->s_comp;
    CUDA_TRY(cudaMemcpyAsync(e->d_io_q.p, q1_host, qb, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(e->d_io_q2.p, q2_host, qb, cudaMemcpyHostToDevice, st));
    // few edges: the sample counts are cheap to compute here (the same FP64 expression as
    // k_edge_counts), which spares the single-level path its only device -> host round trip
    int64_t known_total = -1;
    std::vector<int32_t> hcounts;
    if (n_edges <= 512 && step > 0.0) {
        known_total = 0;
        hcounts.resize((size_t)n_edges);
        for (int64_t i = 0; i < n_edges; i++) {
            double acc = 0.0;
            for (int k = 0; k < e->nq; k++) {
                const double dlt = (double)q2_host[i * e->nq + k] - (double)q1_host[i * e->nq + k];
                acc += dlt * dlt;
            }
            hcounts[(size_t)i] = (int32_t)ceil(sqrt(acc) / step) + 1;
            known_total += hcounts[(size_t)i];
        }
    }
    const bool fused = known_total >= 0 && !first_bad_host && e->npairs > 0 && known_total * (int64_t)e->npairs <= e->small_items;
    rc = check_edges_core(e, e->d_io_q.as<float>(), e->d_io_q2.as<float>(), n_edges, step, padding, e->d_io_hit.as<uint8_t>(),
                          first_bad_host ? e->d_io_i32.as<int32_t>() : nullptr, st, known_total, hcounts.empty() ? nullptr : hcounts.data());
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(hit_host, e->d_io_hit.p, (size_t)n_edges, cudaMemcpyDeviceToHost, st));
    if (first_bad_host) CUDA_TRY(cudaMemcpyAsync(first_bad_host, e->d_io_i32.p, (size_t)n_edges * 4, cudaMemcpyDeviceToHost, st));
    if (known_total > 0 && e->npairs > 0 && !fused) CUDA_TRY(cudaMemcpyAsync(e->h_pinned, ctrl_u64(e, 1), 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (known_total > 0 && e->npairs > 0 && !fused && (int64_t)e->h_pinned[0] != known_total) {
        // host and device sample counts disagree (never observed: both are the same IEEE FP64
        // expression): redo the batch on the path that asks the device
        rc = check_edges_core(e, e->d_io_q.as<float>(), e->d_io_q2.as<float>(), n_edges, step, padding, e->d_io_hit.as<uint8_t>(),
                              first_bad_host ? e->d_io_i32.as<int32_t>() : nullptr, st, -1);
        if (rc) return rc;
        CUDA_TRY(cudaMemcpyAsync(hit_host, e->d_io_hit.p, (size_t)n_edges, cudaMemcpyDeviceToHost, st));
        if (first_bad_host) CUDA_TRY(cudaMemcpyAsync(first_bad_host, e->d_io_i32.p, (size_t)n_edges * 4, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
    }
    return 0;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------
// rejection sampling on device
// ------------------------------------------------------------------------------------------
namespace {

// Philox4x32-10 (Salmon et al., SC'11), counter = (draw index, coordinate block), key = seed.
__device__ __forceinline__ void philox_round(uint32_t &c0, uint32_t &c1, uint32_t &c2, uint32_t &c3, uint32_t k0, uint32_t k1) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
}
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
    uint32_t c0 = ctr.x, c1 = ctr.y, c2 = ctr.z, c3 = ctr.w, k0 = key.x, k1 = key.y;
#pragma unroll
    for (int r = 0; r < 10; r++) {
        philox_round(c0, c1, c2, c3, k0, k1);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}

// q[i][k] = lo_k + u * (hi_k - lo_k), u in [0,1) with 24 random bits; draw index = first_draw + i
__global__ void k_sample_uniform(const float *lower, const float *upper, int nq, float pad, uint64_t seed, uint64_t first_draw,
                                 int64_t n, float *q) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t draw = first_draw + (uint64_t)i;
    for (int k0 = 0; k0 < nq; k0 += 4) {
        const uint4 r = philox4x32_10(make_uint4((uint32_t)draw, (uint32_t)(draw >> 32), (uint32_t)(k0 >> 2), 0u),
                                      make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
        const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
        for (int k = k0; k < min(k0 + 4, nq); k++) {
            const float u = (float)(rr[k - k0] >> 8) * (1.0f / 16777216.0f);
            const float lo = lower[k] + pad, hi = upper[k] - pad;
            q[i * nq + k] = fmaf(u, hi - lo, lo);
        }
    }
}

__global__ void k_invert_flags(const uint8_t *hit, int64_t n, uint8_t *free_flag) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) free_flag[i] = hit[i] ? 0 : 1;
}

__global__ void k_gather_rows(const float *q, const int32_t *rows, int64_t n_rows, int nq, int64_t out_first, int64_t out_cap,
                              float *q_out) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t r = t / nq;
    const int k = (int)(t - r * nq);
    if (r >= n_rows || out_first + r >= out_cap) return;
    q_out[(out_first + r) * nq + k] = q[(int64_t)rows[r] * nq + k];
}

__global__ void k_iota_i32(int32_t *p, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = (int32_t)i;
}

}  // namespace

extern "C" int pbp_sample_free_states(pbp_engine *e, int64_t n_wanted, uint64_t seed, float joint_padding, float padding,
                                      int max_rounds, float *q_out_dev, int64_t *n_found_host, int64_t *n_drawn_host,
                                      void *stream) {
    int rc = check_args(e, q_out_dev, n_wanted);
    if (rc) return rc;
    if (!n_found_host) return fail(PBP_ERR_ARG, "n_found_host is NULL");
    *n_found_host = 0;
    if (n_drawn_host) *n_drawn_host = 0;
    if (n_wanted == 0) return 0;
    if (max_rounds <= 0) max_rounds = 64;
    CallGuard guard(e, (cudaStream_t)stream);
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t batch = std::min<int64_t>(std::max<int64_t>(n_wanted, 4096), e->chunk);
    if ((rc = e->d_io_q.ensure((size_t)batch * e->nq * 4)) || (rc = e->d_io_hit.ensure((size_t)batch)) ||
        (rc = e->d_scratch_u8.ensure((size_t)std::max(batch, e->chunk))) || (rc = e->d_scratch_i32.ensure((size_t)batch * 4)) ||
        (rc = e->d_sel.ensure((size_t)batch * 4 + 16)))
        return rc;
    float *q = e->d_io_q.as<float>();
    uint8_t *hit = e->d_io_hit.as<uint8_t>();
    uint8_t *flag = e->d_scratch_u8.as<uint8_t>();
    int32_t *iota = e->d_scratch_i32.as<int32_t>();
    int32_t *sel = e->d_sel.as<int32_t>();
    int *d_nsel = reinterpret_cast<int *>(ctrl_u64(e, 2));
    k_iota_i32<<<grid_for(batch, 256), 256, 0, st>>>(iota, batch);
    size_t cub_bytes = 0;
    cub::DeviceSelect::Flagged(nullptr, cub_bytes, iota, flag, sel, d_nsel, (int)batch, st);
    if ((rc = e->d_cub.ensure(cub_bytes))) return rc;
    int64_t found = 0, drawn = 0;
    for (int round = 0; round < max_rounds && found < n_wanted; round++) {
        k_sample_uniform<<<grid_for(batch, 256), 256, 0, st>>>(e->dm.q_lower, e->dm.q_upper, e->nq, joint_padding, seed,
                                                              (uint64_t)drawn, batch, q);
        e->stats.kernel_launches++;
        if ((rc = pbp_check_states(e, q, batch, padding, hit, nullptr, nullptr, st))) return rc;
        k_invert_flags<<<grid_for(batch, 256), 256, 0, st>>>(hit, batch, flag);
        cub::DeviceSelect::Flagged(e->d_cub.p, cub_bytes, iota, flag, sel, d_nsel, (int)batch, st);
        e->stats.kernel_launches += 2;
        CUDA_TRY(cudaMemcpyAsync(e->h_pinned, d_nsel, 4, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        const int64_t nsel = (int64_t)(*reinterpret_cast<int *>(e->h_pinned));
        const int64_t take = std::min(nsel, n_wanted - found);
        if (take > 0) {
            k_gather_rows<<<grid_for(take * e->nq, 256), 256, 0, st>>>(q, sel, take, e->nq, found, n_wanted, q_out_dev);
            e->stats.kernel_launches++;
        }
        found += take;
        drawn += batch;
    }
    CUDA_TRY(cudaStreamSynchronize(st));
    CUDA_TRY(cudaGetLastError());
    *n_found_host = found;
    if (n_drawn_host) *n_drawn_host = drawn;
    return 0;
}

// ------------------------------------------------------------------------------------------
// FP32 FMA peak microbenchmark (roofline denominator of this FP32-pipe-bound path)
// ------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256) k_fma_peak(float *out, int iters, float a, float b) {
    float x0 = threadIdx.x * 1e-3f, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f, x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f,
          x7 = x0 + 7.f;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
            x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
}  // namespace

extern "C" int pbp_measure_fp32_peak(int device, double *tflops_out) {
    if (!tflops_out) return fail(PBP_ERR_ARG, "NULL argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return fail(PBP_ERR_CUDA, "no such CUDA device");
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 4096;
    float *buf = nullptr;
    CUDA_TRY(cudaMalloc(&buf, (size_t)blocks * threads * 4));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 6; rep++) {
        CUDA_TRY(cudaEventRecord(e0, 0));
        k_fma_peak<<<blocks, threads>>>(buf, iters, 0.999f, 1e-3f);
        CUDA_TRY(cudaEventRecord(e1, 0));
        CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        const double flops = 2.0 * 64.0 * (double)iters * (double)blocks * (double)threads;
        if (rep > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(buf);
    CUDA_TRY(cudaGetLastError());
    *tflops_out = best;
    return 0;
}

// ------------------------------------------------------------------------------------------
// differential IK (one try for N targets)
// ------------------------------------------------------------------------------------------
extern "C" int pbp_ik_iterate(pbp_engine *e, int32_t frame_id, const double *targets_dev, double *q_dev, int64_t n,
                              const pbp_ik_options *opt, uint64_t active_mask, const double *w_diag_host,
                              const double *nullspace_dev, double *e0_dev, int32_t *status_dev, int32_t *iters_dev, void *stream) {
    int rc = check_args(e, q_dev, n);
    if (rc) return rc;
    if (n > 0 && (!targets_dev || !e0_dev || !status_dev || !opt)) return fail(PBP_ERR_ARG, "NULL buffer");
    if (frame_id < 0 || frame_id >= e->nframes) return fail(PBP_ERR_ARG, "frame %d out of range [0, %d)", frame_id, e->nframes);
    if (opt->max_iters < 0) return fail(PBP_ERR_ARG, "max_iters must be >= 0");
    if (n == 0) return 0;
    CallGuard guard(e, (cudaStream_t)stream);
    cudaStream_t st = (cudaStream_t)stream;
    // chain root -> frame, as float64 tables
    std::vector<int> path;
    for (int j = e->h_frame_parent[frame_id]; j > 0; j = e->h_joint_parent[j])
        if (e->h_joint_type[j] != PBP_JOINT_FIXED) path.push_back(j);
    std::reverse(path.begin(), path.end());
    if ((int)path.size() > IK_MAX_PATH) return fail(PBP_ERR_CAPACITY, "frame chain longer than %d joints", IK_MAX_PATH);
    IkTables &tab = e->ik_build;
    memset(&tab, 0, sizeof(tab));
    // fixed joints between two moving ones fold into the next moving joint's placement
    {
        std::vector<int> full;
        for (int j = e->h_frame_parent[frame_id]; j > 0; j = e->h_joint_parent[j]) full.push_back(j);
        std::reverse(full.begin(), full.end());
        double acc[12] = {1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0};
        int k = 0;
        auto mul = [](const double *A, const double *B, double *C) {
            for (int i = 0; i < 3; i++) {
                for (int c = 0; c < 3; c++) C[3 * i + c] = A[3 * i] * B[c] + A[3 * i + 1] * B[3 + c] + A[3 * i + 2] * B[6 + c];
                C[9 + i] = A[3 * i] * B[9] + A[3 * i + 1] * B[10] + A[3 * i + 2] * B[11] + A[9 + i];
            }
        };
        for (int j : full) {
            double tmp[12];
            mul(acc, e->h_joint_placement.data() + 12 * (size_t)j, tmp);
            if (e->h_joint_type[j] == PBP_JOINT_FIXED) {
                memcpy(acc, tmp, sizeof(acc));
                continue;
            }
            IkJoint &J = tab.joint[k++];
            memcpy(J.P, tmp, sizeof(tmp));
            for (int a = 0; a < 3; a++) J.axis[a] = e->h_joint_axis[3 * (size_t)j + a];
            J.type = e->h_joint_type[j];
            J.idx_q = e->h_joint_idx_q[j];
            const bool active = J.idx_q < 64 && ((active_mask >> J.idx_q) & 1ull);
            J.w = active ? (w_diag_host ? w_diag_host[J.idx_q] : 1.0) : 0.0;
            if (J.idx_q < 64) tab.path_mask |= 1ull << J.idx_q;
            const double ident[12] = {1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0};
            memcpy(acc, ident, sizeof(acc));
        }
        double fr[12];
        mul(acc, e->h_frame_placement.data() + 12 * (size_t)frame_id, fr);
        memcpy(tab.frame, fr, sizeof(fr));
        tab.npath = k;
    }
    tab.nq = e->nq;
    tab.active_mask = active_mask;
    for (int k = 0; k < e->nq; k++) { tab.q_lower[k] = e->h_q_lower[k]; tab.q_upper[k] = e->h_q_upper[k]; }
    if ((rc = e->d_ik_tables.ensure(sizeof(IkTables)))) return rc;
    if (!e->ik_cached_valid || memcmp(&tab, &e->ik_cached, sizeof(IkTables)) != 0) {
        // rare (new frame / weights): a synchronous upload, ordered after earlier work on `st`
        CUDA_TRY(cudaStreamSynchronize(st));
        CUDA_TRY(cudaMemcpy(e->d_ik_tables.p, &tab, sizeof(IkTables), cudaMemcpyHostToDevice));
        memcpy(&e->ik_cached, &tab, sizeof(IkTables));
        e->ik_cached_valid = true;
    }
    IkOptions o;
    o.max_iters = opt->max_iters;
    o.tol_translation = opt->max_translation_error;
    o.tol_rotation = opt->max_rotation_error;
    o.damping = opt->damping;
    o.min_step = opt->min_step_size;
    o.max_step = opt->max_step_size;
    const IkTables *dt = e->d_ik_tables.as<IkTables>();
    unsigned long long *cursor = ctrl_u64(e, 4);
    CUDA_TRY(cudaMemsetAsync(cursor, 0, 8, st));
    // persistent lanes (targets are pulled from the cursor): exactly the blocks that are resident at once
    auto launch = [&](auto kern) -> int {
        int occ = 1;
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 128, 0));
        const unsigned grid = (unsigned)std::min<int64_t>((n + 127) / 128, (int64_t)e->num_sms * std::max(occ, 1));
        kern<<<grid, 128, 0, st>>>(dt, targets_dev, q_dev, n, o, nullspace_dev, e0_dev, status_dev, iters_dev, cursor);
        return 0;
    };
    if (tab.npath <= 8) rc = launch(k_ik_iterate<8>);
    else if (tab.npath <= 16) rc = launch(k_ik_iterate<16>);
    else rc = launch(k_ik_iterate<IK_MAX_PATH>);
    if (rc) return rc;
    LAUNCH_CHECK("k_ik_iterate");
    return 0;
}

// ------------------------------------------------------------------------------------------
// k nearest neighbours (batched PRM construction)
// ------------------------------------------------------------------------------------------
extern "C" int pbp_knn_blocks(pbp_engine *e, const float *nodes_dev, int64_t n, int32_t k, double radius, int32_t predecessors_only,
                              int64_t block_first, int64_t block_stride, int32_t *idx_dev, float *dist_dev, void *stream) {
    int rc = check_args(e, nodes_dev, n);
    if (rc) return rc;
    if (n > 0 && (!idx_dev || !dist_dev)) return fail(PBP_ERR_ARG, "NULL buffer");
    if (k < 1 || k > KNN_MAX_K) return fail(PBP_ERR_ARG, "k must be in [1, %d]", KNN_MAX_K);
    if (e->nq > KNN_MAX_NQ) return fail(PBP_ERR_CAPACITY, "nq > %d", KNN_MAX_NQ);
    if (!(radius >= 0.0)) return fail(PBP_ERR_ARG, "radius must be >= 0");
    if (block_first < 0 || block_stride < 1) return fail(PBP_ERR_ARG, "bad query block range");
    const int64_t nblocks = (n + KNN_QUERIES - 1) / KNN_QUERIES;
    if (n == 0 || block_first >= nblocks) return 0;
    CallGuard guard(e, (cudaStream_t)stream);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = knn_smem_bytes(e->nq, k);
    const double r2 = std::isinf(radius) ? INFINITY : radius * radius;
    const unsigned grid = (unsigned)((nblocks - block_first + block_stride - 1) / block_stride);
#define KNN_LAUNCH(NQ)                                                                                                   \
    do {                                                                                                                 \
        CUDA_TRY(cudaFuncSetAttribute(k_knn<NQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));               \
        k_knn<NQ><<<grid, KNN_THREADS, smem, st>>>(nodes_dev, n, e->nq, k, r2, predecessors_only, block_first, block_stride, idx_dev, dist_dev); \
    } while (0)
    switch (e->nq) {
    case 2: KNN_LAUNCH(2); break;
    case 6: KNN_LAUNCH(6); break;
    case 7: KNN_LAUNCH(7); break;
    case 9: KNN_LAUNCH(9); break;
    default: KNN_LAUNCH(0); break;
    }
#undef KNN_LAUNCH
    LAUNCH_CHECK("k_knn");
    return 0;
}

extern "C" int pbp_knn(pbp_engine *e, const float *nodes_dev, int64_t n, int32_t k, double radius, int32_t predecessors_only,
                       int32_t *idx_dev, float *dist_dev, void *stream) {
    return pbp_knn_blocks(e, nodes_dev, n, k, radius, predecessors_only, 0, 1, idx_dev, dist_dev, stream);
}

// ------------------------------------------------------------------------------------------
// per-pair distances with witness points; collision-avoidance null-space term
// ------------------------------------------------------------------------------------------
namespace {
int launch_distances(pbp_engine *e, const float *q_dev, int64_t n, float cutoff, float *oMi, float *oMg, float *dist, float *p1,
                     float *p2, float *nrm, cudaStream_t st) {
    k_fk_output<<<grid_for(n, 128), 128, e->smem_fk, st>>>(e->dm, q_dev, n, oMi, oMg, nullptr);
    LAUNCH_CHECK("k_fk_output");
    const int64_t items = n * e->npairs;
    k_pair_distances<<<grid_for(items, 128), 128, 0, st>>>(e->dm, oMg, n, cutoff, dist, p1, p2, nrm);
    LAUNCH_CHECK("k_pair_distances");
    return 0;
}
}  // namespace

extern "C" int pbp_compute_distances(pbp_engine *e, const float *q_dev, int64_t n, float cutoff, float *dist_dev, float *p1_dev,
                                     float *p2_dev, float *normal_dev, void *stream) {
    int rc = check_args(e, q_dev, n);
    if (rc) return rc;
    if (n > 0 && !dist_dev) return fail(PBP_ERR_ARG, "dist buffer is NULL");
    if (!(cutoff >= 0.0f)) return fail(PBP_ERR_ARG, "cutoff must be >= 0");
    if (n == 0 || e->npairs == 0) return 0;
    if (n * (int64_t)e->npairs > (int64_t)INT32_MAX * 64) return fail(PBP_ERR_ARG, "too many (state, pair) items");
    CallGuard guard(e, (cudaStream_t)stream);
    if ((rc = e->d_ws_oMg.ensure((size_t)n * e->ngeoms * 48))) return rc;
    return launch_distances(e, q_dev, n, cutoff, nullptr, e->d_ws_oMg.as<float>(), dist_dev, p1_dev, p2_dev, normal_dev,
                            (cudaStream_t)stream);
}

extern "C" int pbp_collision_nullspace(pbp_engine *e, const float *q_dev, int64_t n, float dist_padding, float gain,
                                       float *component_dev, void *stream) {
    int rc = check_args(e, q_dev, n);
    if (rc) return rc;
    if (n > 0 && !component_dev) return fail(PBP_ERR_ARG, "component buffer is NULL");
    if (!(dist_padding >= 0.0f)) return fail(PBP_ERR_ARG, "dist_padding must be >= 0");
    if (n == 0) return 0;
    CallGuard guard(e, (cudaStream_t)stream);
    cudaStream_t st = (cudaStream_t)stream;
    if (e->npairs == 0) {
        CUDA_TRY(cudaMemsetAsync(component_dev, 0, (size_t)n * e->nq * 4, st));
        return 0;
    }
    const size_t items = (size_t)n * e->npairs;
    if ((rc = e->d_ws_oMi.ensure((size_t)n * e->njoints * 48)) || (rc = e->d_ws_oMg.ensure((size_t)n * e->ngeoms * 48)) ||
        (rc = e->d_ws_dist.ensure(items * 4)) || (rc = e->d_ws_p1.ensure(items * 12)) || (rc = e->d_ws_p2.ensure(items * 12)) ||
        (rc = e->d_ws_nrm.ensure(items * 12)))
        return rc;
    if ((rc = launch_distances(e, q_dev, n, dist_padding, e->d_ws_oMi.as<float>(), e->d_ws_oMg.as<float>(), e->d_ws_dist.as<float>(),
                               e->d_ws_p1.as<float>(), e->d_ws_p2.as<float>(), e->d_ws_nrm.as<float>(), st)))
        return rc;
    k_collision_nullspace<<<grid_for(n, 128), 128, 0, st>>>(e->dm, e->d_pairs_orig.as<int32_t>(), e->d_ws_oMi.as<float>(),
                                                           e->d_ws_dist.as<float>(), e->d_ws_p1.as<float>(), e->d_ws_p2.as<float>(),
                                                           e->d_ws_nrm.as<float>(), n, dist_padding, gain, component_dev);
    LAUNCH_CHECK("k_collision_nullspace");
    return 0;
}

// ------------------------------------------------------------------------------------------
// JIT introspection (tests, tooling): generated source / offline NVRTC build, no device needed
// ------------------------------------------------------------------------------------------
extern "C" int pbp_jit_broadphase_source(const pbp_model_desc *d, char *buf, size_t cap, size_t *needed) {
    if (!d || !needed) return fail(PBP_ERR_ARG, "NULL argument");
    HostTables H;
    int rc = build_host_tables(d, H);
    if (rc) return rc;
    const std::string src_text = specialised_broadphase_source(d, H);
    *needed = src_text.size() + 1;
    if (buf && cap >= *needed) memcpy(buf, src_text.c_str(), *needed);
    return 0;
}

extern "C" int pbp_jit_compile_check(const pbp_model_desc *d, int cc_major, int cc_minor, size_t *cubin_bytes, char *log_buf, size_t log_cap) {
    if (!d || !cubin_bytes) return fail(PBP_ERR_ARG, "NULL argument");
    HostTables H;
    int rc = build_host_tables(d, H);
    if (rc) return rc;
    const std::string src_text = specialised_broadphase_source(d, H);
    if (src_text.empty()) return fail(PBP_ERR_CAPACITY, "model does not qualify for the specialised broadphase");
    std::vector<char> cubin;
    std::string log;
    const int crc = pbp::jit::compile(src_text, cc_major, cc_minor, cubin, log);
    if (log_buf && log_cap) { strncpy(log_buf, log.c_str(), log_cap - 1); log_buf[log_cap - 1] = 0; }
    if (crc) return fail(PBP_ERR_CUDA, "NVRTC failed: %s", log.c_str());
    *cubin_bytes = cubin.size();
    if (const char *dump = getenv("PBP_JIT_DUMP_CUBIN")) {
        FILE *f = fopen(dump, "wb");
        if (f) { fwrite(cubin.data(), 1, cubin.size(), f); fclose(f); }
    }
    return 0;
}

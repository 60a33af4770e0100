#!/usr/bin/env python
"""Headline benchmark: collision-checked Panda configurations per second (BASELINE.json config 2).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A *step* is one pass of the hot path (FK + placement + boolean collision over the 74 active
pairs) over one batch of 1 M uniformly sampled Panda configurations per GPU.  Under torchrun
(N > 1) every rank owns its own batches (weak scaling, no data-path collective); the timed
region is bracketed by a barrier + synchronize and the max over ranks is reported.

Output: ONE JSON line on rank 0 (see DESIGN.md "Measurement" for every field).
`--impl reference` times the CPU restatement of the reference path (oracle/, float64, all host
threads) on a bounded sample of the same workload; Pinocchio/coal cannot be installed here.
"""

import argparse
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "collision-checked configs/sec (Panda 7-DOF, 1/2/4/8 B200) vs reference CPU"
UNIT = "configs/s"
WORKLOAD = "config2: Panda nq=9, 16 geoms, 74 active pairs (20 self + 32 box + 22 sphere), uniform states in joint limits, padding 0, mesh surfaces as in the reference (BVHModel semantics)"
N_STATES = 1 << 20          # states per GPU per step (BASELINE.json config 2: 1 M random states)
N_BATCHES = 4               # distinct input batches rotated between steps (4 x 37.7 MB > 126 MB L2)

# Algorithmic FP32 work per configuration, frozen from the float64 oracle's counters on the
# config-2 distribution (seed 0, 100 k states, cull-aware counting mode; DESIGN.md "F_alg"):
#   FK 1.5 kflop + broadphase 74 pairs x 30 flop + 6 flop x 1598 hull vertices scanned
#   + 150 flop x 7.57 GJK iterations  = 14.4 kflop / configuration.
F_BP_FLOP_PER_CONFIG = 1.5e3 + 74 * 30.0          # k_broadphase: FK + 74 proxy tests
F_NP_FLOP_PER_CONFIG = 6.0 * 1598 + 150.0 * 7.57  # k_narrowphase: support scans + simplex updates
F_ALG_FLOP_PER_CONFIG = F_BP_FLOP_PER_CONFIG + F_NP_FLOP_PER_CONFIG  # = 14.4 kflop
# The GJK work is split over two kernels: k_item_setup settles the items that end after one iteration
# or by the inner-ball certificate (about 45 % of the oracle's iterations), k_narrowphase the rest (the
# split is the round-1 measurement, kept so the fractions stay comparable across rounds); the item
# setup's FK recomputation is redundant work, not algorithmic flops.
F_KERNEL = {"k_broadphase": F_BP_FLOP_PER_CONFIG, "k_item_setup": 0.45 * F_NP_FLOP_PER_CONFIG, "k_narrowphase": 0.55 * F_NP_FLOP_PER_CONFIG}
HBM_BYTES_PER_CONFIG = 9 * 4 + 1   # 9 float32 in, 1 verdict byte out
# DRAM traffic per launch (dram__bytes_read.sum + dram__bytes_write.sum) and pipe utilisation of
# the five kernels of a 1 Mi-state round, from the committed `ncu --set full` capture
# profiles/r02_v6_ncu_summary.txt (cold-cache, serialised launches of this same command).
# k_surface_refine / k_surface_settle (what the cores and hulls leave open: enclosures, triangle stage;
# DESIGN.md 4) carry no algorithmic flops in F_alg -- the oracle's counters were frozen on the hull
# arithmetic -- so their time only lowers the step's achieved figure.
NCU = {
    "source": "profiles/r02_v6_ncu_summary.txt",
    "k_broadphase": {"traffic_bytes": 37.81e6 + 0.002e6, "fma_pipe_pct": 34.2, "alu_pipe_pct": 61.8, "issue_active_pct": 76.3,
                     "threads_per_inst": 28.2, "duration_us": 104.8},
    "k_item_setup": {"traffic_bytes": 40.90e6 + 4.20e6, "fma_pipe_pct": 24.6, "alu_pipe_pct": 36.8, "issue_active_pct": 58.2,
                     "threads_per_inst": 20.5, "duration_us": 102.5},
    "k_narrowphase": {"traffic_bytes": 9.32e6 + 0.0003e6, "fma_pipe_pct": 16.0, "alu_pipe_pct": 30.0, "issue_active_pct": 44.5,
                      "threads_per_inst": 12.1, "duration_us": 100.0},
    "k_surface_refine": {"traffic_bytes": 9.14e6 + 0.18e6, "fma_pipe_pct": 4.5, "alu_pipe_pct": 10.8, "issue_active_pct": 16.9,
                         "threads_per_inst": 15.2, "duration_us": 39.2},
    "k_surface_settle": {"traffic_bytes": 1.82e6 + 0.0, "fma_pipe_pct": 6.7, "alu_pipe_pct": 13.7, "issue_active_pct": 19.8,
                         "threads_per_inst": 28.0, "duration_us": 34.4},
}


def panda_models():
    from pyroboplan_b200.models import panda

    model, cmodel, vmodel = panda.load_models()
    panda.add_self_collisions(model, cmodel)
    panda.add_object_collisions(model, cmodel, vmodel)
    return model, cmodel


def make_batch(model, n, seed):
    rng = np.random.default_rng(seed)
    return rng.uniform(model.lowerPositionLimit, model.upperPositionLimit, size=(n, model.nq)).astype(np.float32)


class ClockSampler:
    """Samples SM clock, power and throttle reasons through NVML (a poll every ~2 ms from a thread;
    `nvidia-smi -lms` cannot start, let alone sample, inside a timed region of a few milliseconds).
    Samples carry a timestamp; `stop(t0, t1)` reports those taken inside [t0, t1]."""

    REASONS = {  # nvmlClocksEventReasons bits (nvml.h)
        0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
    }

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.samples = []   # (t, sm_mhz, power_w, reasons bitmask)
        self.running = False
        self.handle = None
        self.sm_max = None
        try:
            import pynvml

            self.nv = pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES when it lists indices
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            idx = gpu_index
            if vis and all(v.strip().isdigit() for v in vis.split(",")):
                idx = int(vis.split(",")[gpu_index])
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(idx)
            elif vis and vis.split(",")[gpu_index].strip().startswith("GPU-"):
                self.handle = pynvml.nvmlDeviceGetHandleByUUID(vis.split(",")[gpu_index].strip())
            else:
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception as exc:  # no NVML: report that, never fail the bench
            self.error = repr(exc)
            self.handle = None

    def _poll(self):
        nv = self.nv
        while self.running:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)
                try:
                    reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                except Exception:
                    reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                power = nv.nvmlDeviceGetPowerUsage(self.handle) / 1000.0
                self.samples.append((time.perf_counter(), float(sm), power, int(reasons)))
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.handle is None:
            return
        self.running = True
        self.thread = threading.Thread(target=self._poll, daemon=True)
        self.thread.start()

    def stop(self, t0, t1):
        if self.handle is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [f"NVML unavailable: {getattr(self, 'error', '')}"]}
        self.running = False
        self.thread.join(timeout=1.0)
        inside = [x for x in self.samples if t0 <= x[0] <= t1]
        window = "timed region"
        if not inside:  # a timed region shorter than one poll: fall back to the whole loaded window
            inside, window = self.samples, "warm-up + timed region (timed region shorter than one NVML poll)"
        bits = 0
        for x in inside:
            bits |= x[3]
        return {
            "sm_mhz": statistics.median(x[1] for x in inside) if inside else None,
            "sm_min_mhz": min(x[1] for x in inside) if inside else None,
            "sm_max_mhz": self.sm_max,
            "power_w_max": max(x[2] for x in inside) if inside else None,
            "samples": len(inside),
            "window": window,
            "reasons": sorted(name for bit, name in self.REASONS.items() if bits & bit),
        }


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


def run_reference(args, rank):
    """CPU arm: the float64 oracle (port of the reference path), all host threads, bounded sample."""
    if rank != 0:
        return
    from oracle import pinocchio_ref
    from oracle.oracle import Oracle

    model, cmodel = panda_models()
    threads = len(os.sched_getaffinity(0))
    sample = int(os.environ.get("PBP_REF_SAMPLE", args.states))   # the GPU arm's step: 1 Mi states
    if pinocchio_ref.available():
        # Tier A: the reference's own arithmetic (pinocchio.computeCollisions through its check_collisions_at_state
        # loop), one model copy per worker process, on a bounded prefix of the same batches
        sub = min(sample, int(os.environ.get("PBP_PIN_SAMPLE", 200_000)))
        rates = []
        for k in range(max(args.steps, 1)):
            rate, _ = pinocchio_ref.timed_rate(make_batch(model, sub, 7919 * 0 + (k % 2)).astype(np.float64), threads)
            rates.append(rate)
        value = float(np.mean(rates))
        line = {
            "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": sample / value * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "states_per_gpu_per_step": sample,
                       "note": f"CPU arm: Pinocchio + coal (the reference's own check_collisions_at_state loop), {threads} worker processes, "
                               f"{sub}-state prefix of each step's batch"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "reference",
                             "sample": f"{sub} states/step x {args.steps} steps of the config-2 distribution through pinocchio.computeCollisions"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "pinocchio": "importable: Tier-A reference timing",
        }
        print(json.dumps(line), flush=True)
        return
    orc = Oracle(model, cmodel)
    batches = [make_batch(model, sample, 7919 * 0 + b).astype(np.float64) for b in range(2)]   # rank 0's first two batches of the GPU arm
    for w in range(max(args.warmup, 1)):
        orc.check_states(batches[w % 2][: sample // 4], use_cull=1, nthreads=threads)
    t0 = time.perf_counter()
    for k in range(args.steps):
        orc.check_states(batches[k % 2], use_cull=1, nthreads=threads)
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    line = {
        "impl": "reference",
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "states_per_gpu_per_step": sample,
                   "l2": f"inputs rotate over {N_BATCHES} batches ({N_BATCHES * sample * 36 / 1e6:.0f} MB > 126 MB L2)",
                   "parallelism": f"{args.gpus} rank(s), states sharded, no data-path collective",
                   "note": "CPU arm: float64 restatement of the reference path (oracle/pbp_oracle.c, kind 'port') on all host threads of rank 0 "
                           "-- NOT upstream Pinocchio/coal timing: the probe for `import pinocchio` failed on this box (see `pinocchio`)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{sample} states/step x {args.steps} steps of the config-2 distribution, bounding-sphere cull + first-hit exit, exact triangle tests where hulls touch (surface semantics)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "pinocchio": "not importable: oracle port timed instead",
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--states", type=int, default=N_STATES, help="states per GPU per step")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist

    from pyroboplan_b200.engine import CollisionEngine, measure_fp32_peak

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # keep stdout to the single JSON line: NCCL prints its version banner there at VERSION and WARN level
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
            os.environ.pop("NCCL_DEBUG")
        dist.init_process_group("nccl", device_id=dev)

    if world > 1:
        # one process per GPU on a shared host: give every rank its own slice of the cores, so that a rank's
        # launch thread is not descheduled behind another rank's (the step is ~0.4 ms of GPU time)
        try:
            cores = sorted(os.sched_getaffinity(0))
            mine = cores[local_rank::world] if len(cores) >= 2 * world else cores
            os.sched_setaffinity(0, mine)
        except (AttributeError, OSError):
            pass
    model, cmodel = panda_models()
    eng = CollisionEngine(model, cmodel, device=local_rank)
    n = args.states
    seed_rank = int(os.environ.get("PBP_BENCH_SEED_RANK", rank))   # (developer knob: another rank's inputs on this rank)
    host_batches = [torch.from_numpy(make_batch(model, n, 7919 * seed_rank + b)).pin_memory() for b in range(N_BATCHES)]
    dev_batches = [b.to(dev, non_blocking=True) for b in host_batches]
    hits = [torch.empty(n, dtype=torch.bool, device=dev) for _ in range(N_BATCHES)]   # fixed buffers: every round replays its CUDA graph
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()

    # ---------------- device-resident throughput
    # W untimed warm-up steps (at least one per input batch, so no batch is touched for the
    # first time inside the timed region)
    sampler = ClockSampler(local_rank)
    sampler.start()   # every rank samples its own GPU (a poll every 2 ms on one of the rank's own cores; the timed region
                      # replays CUDA graphs, one launch per step, so the poller cannot starve the launch thread)
    for w in range(max(args.warmup, N_BATCHES + 1)):
        eng.check_states(dev_batches[w % N_BATCHES], out=hits[w % N_BATCHES])
    torch.cuda.synchronize()
    fp32_peak = measure_fp32_peak(local_rank)
    eng.stats(reset=True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    torch.cuda.synchronize()
    t_region0 = time.perf_counter()
    ev0.record()
    for k in range(args.steps):
        eng.check_states(dev_batches[k % N_BATCHES], out=hits[k % N_BATCHES])
    ev1.record()
    torch.cuda.synchronize()
    t_region1 = time.perf_counter()
    barrier()
    clocks = sampler.stop(t_region0, t_region1)
    clocks_by_rank = None
    if world > 1:
        allc = [None] * world
        dist.all_gather_object(allc, {"sm_mhz": clocks.get("sm_mhz"), "power_w_max": clocks.get("power_w_max"), "reasons": clocks.get("reasons")})
        clocks_by_rank = allc
    ms = ev0.elapsed_time(ev1)
    stats = eng.stats(reset=True)
    # per-kernel launch durations: the same K steps once more with CUDA events around every kernel (plain
    # launches: the headline region above replays one CUDA graph per step and carries no events)
    eng.set_profiling(True)
    eng.profile(reset=True)
    for k in range(args.steps):
        eng.check_states(dev_batches[k % N_BATCHES], out=hits[k % N_BATCHES])
    torch.cuda.synchronize()
    prof = eng.profile(reset=True)
    eng.set_profiling(False)
    eng.stats(reset=True)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    rank_ms = None
    if world > 1:
        gathered = [torch.zeros(1, dtype=torch.float64, device=dev) for _ in range(world)]
        dist.all_gather(gathered, t)
        rank_ms = [float(x.item()) / args.steps for x in gathered]
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * n * args.steps / (ms_max * 1e-3)

    # ---------------- end to end through the host-buffer C ABI (pinned inputs, H2D + D2H timed)
    host_np = [b.numpy() for b in host_batches]
    for w in range(2):
        eng.check_states(host_np[w % N_BATCHES])
    barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(args.steps):
        hit_host = eng.check_states(host_np[k % N_BATCHES])
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * n * args.steps / float(t.item())
    # the same from float64 PAGEABLE numpy arrays -- what the reference's callers hold and what the drop-in
    # functions receive: the library narrows to float32 on host threads, overlapped with copies and kernels
    host_f64 = [b.numpy().astype(np.float64) for b in host_batches[:2]]
    for w in range(2):
        eng.check_states(host_f64[w % 2])
    barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(args.steps):
        hit_f64 = eng.check_states(host_f64[k % 2])
    torch.cuda.synchronize()
    f64_s = time.perf_counter() - t0
    t = torch.tensor([f64_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_f64_value = world * n * args.steps / float(t.item())
    # pinned H2D bandwidth of this box, for context (the e2e path moves 36 B per state over PCIe)
    h2d0, h2d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h2d0.record()
    for b in range(N_BATCHES):
        dev_batches[b].copy_(host_batches[b], non_blocking=True)
    h2d1.record()
    torch.cuda.synchronize()
    h2d_gbs = N_BATCHES * n * 36 / (h2d0.elapsed_time(h2d1) * 1e-3) / 1e9

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- roofline of the dominant kernel (per-launch CUDA events inside the timed region)
    peaks, peak_src = measured_peaks()
    rounds = max(prof["rounds"], 1)
    bp_ms, np_ms = prof["broadphase_ms"] / rounds, prof["narrowphase_ms"] / rounds
    dominant = "k_narrowphase" if np_ms >= bp_ms else "k_broadphase"
    states_per_launch = n * args.steps / rounds
    kernel_ms = max(np_ms, bp_ms)
    f_dom = F_KERNEL[dominant]
    achieved_tflops = f_dom * states_per_launch / (kernel_ms * 1e-3) / 1e12
    sr_ms = prof.get("surface_refine_ms", 0.0) / rounds
    step_kernel_ms = bp_ms + np_ms + prof["item_setup_ms"] / rounds + sr_ms
    step_tflops = F_ALG_FLOP_PER_CONFIG * states_per_launch / (step_kernel_ms * 1e-3) / 1e12
    hbm_gbs = HBM_BYTES_PER_CONFIG * states_per_launch / (step_kernel_ms * 1e-3) / 1e9
    roofline = {
        "bound": "fp32", "achieved": achieved_tflops, "peak": fp32_peak, "unit": "TFLOP/s",
        "frac": achieved_tflops / fp32_peak if fp32_peak else None,
        "traffic": NCU[dominant]["traffic_bytes"] * (states_per_launch / float(1 << 20)),
        "traffic_source": f"{NCU['source']} (per 1 Mi-state launch, scaled to this launch size)",
        "kernel": dominant,
        "kernel_avg_ms": kernel_ms,
        "kernel_share_of_step": kernel_ms / step_kernel_ms if step_kernel_ms else None,
        "peak_source": "FFMA microbenchmark run in this process (pbp_measure_fp32_peak); MEASURED_PEAKS.json has no FP32 entry",
        "note": "scalar FP32 geometry: neither HBM- nor tensor-bound (SURVEY.md 8d); bound = FP32 pipe. `achieved` counts the "
                "ALGORITHMIC flops frozen from the oracle's counters (full hull scans, DESIGN.md 5); the kernels execute fewer "
                "(exact support maps scan ~3 of ~150 vertices), so the executed-instruction view is in `ncu`",
        "ncu": NCU,
        "kernel_ms": {"k_broadphase": bp_ms, "k_item_setup": prof["item_setup_ms"] / rounds, "k_narrowphase": np_ms,
                      "k_surface_refine+k_surface_settle": sr_ms},
        "by_kernel": {name: {"ms": k_ms, "achieved": F_KERNEL[name] * states_per_launch / (k_ms * 1e-3) / 1e12 if k_ms else None,
                             "frac": F_KERNEL[name] * states_per_launch / (k_ms * 1e-3) / 1e12 / fp32_peak if (k_ms and fp32_peak) else None}
                      for name, k_ms in (("k_broadphase", bp_ms), ("k_item_setup", prof["item_setup_ms"] / rounds), ("k_narrowphase", np_ms))},
        "states_per_launch": states_per_launch,
        "narrowphase_items_per_state": prof["narrowphase_items"] / max(n * args.steps, 1),
        "flop_per_config": F_KERNEL,
        "step": {"achieved": step_tflops, "frac": step_tflops / fp32_peak if fp32_peak else None, "unit": "TFLOP/s"},
        "hbm": {"bound": "hbm", "achieved": hbm_gbs, "peak": peaks.get("hbm_gbs"), "unit": "GB/s",
                "frac": hbm_gbs / peaks.get("hbm_gbs", 1.0), "peak_source": f"MEASURED_PEAKS.json ({peak_src})"},
    }

    # ---------------- CPU baseline + parity on a bounded sample (rank 0, N = 1 only)
    cpu_baseline = None
    parity = None
    if world == 1:
        from oracle.oracle import Oracle

        orc = Oracle(model, cmodel)
        threads = len(os.sched_getaffinity(0))
        sample_n = min(n, int(os.environ.get("PBP_CPU_SAMPLE", 500_000)))
        qs = host_np[(args.steps - 1) % N_BATCHES][:sample_n].astype(np.float64)
        orc.check_states(qs[:20000], use_cull=1, nthreads=threads)
        t0 = time.perf_counter()
        hit_cpu, _, _, _ = orc.check_states(qs, use_cull=1, nthreads=threads)
        dt = time.perf_counter() - t0
        cpu_baseline = {"value": sample_n / dt, "unit": UNIT, "cores": threads, "kind": "port",
                        "sample": f"first {sample_n} states of the last timed batch, float64 oracle, bounding-sphere cull + first-hit exit, exact triangle tests where hulls touch (surface semantics)"}
        bad = np.nonzero(hit_cpu.astype(bool) != hit_host[:sample_n])[0]
        beyond = 0
        if len(bad):
            # classify: verdicts may differ only where the reference min distance is within 1e-6 m
            _, md_bad, _, _ = orc.check_states(qs[bad], want_all=True, use_cull=0, nthreads=threads)
            beyond = int((md_bad > 1e-6).sum())
        parity = {"checked": sample_n, "mismatches": int(len(bad)), "mismatches_beyond_1e-6m": beyond, "checker": "oracle (float64 port)"}
        from oracle import pinocchio_ref

        if pinocchio_ref.available():   # Tier A: the reference's own arithmetic is on this box -- time it and check against it
            sub = min(sample_n, int(os.environ.get("PBP_PIN_SAMPLE", 200_000)))
            rate, hit_pin = pinocchio_ref.timed_rate(qs[:sub], threads)
            cpu_baseline = {"value": rate, "unit": UNIT, "cores": threads, "kind": "reference",
                            "sample": f"first {sub} states of the last timed batch through pinocchio.computeCollisions (the reference's check_collisions_at_state loop), one process per core",
                            "port_value": cpu_baseline["value"]}
            badp = np.nonzero(hit_pin != hit_host[:sub])[0]
            beyondp = 0
            if len(badp):
                _, md_bad, _, _ = orc.check_states(qs[badp], want_all=True, use_cull=0, nthreads=threads)
                beyondp = int((md_bad > 1e-6).sum())
            parity["pinocchio"] = {"checked": sub, "mismatches": int(len(badp)), "mismatches_beyond_1e-6m": beyondp}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "ms_per_step_by_rank": rank_ms,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "states_per_gpu_per_step": n,
                   "l2": f"inputs rotate over {N_BATCHES} batches ({N_BATCHES * n * 36 / 1e6:.0f} MB > 126 MB L2)",
                   "parallelism": f"{world} rank(s), states sharded, no data-path collective"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n * 9 * 4, "d2h_bytes_per_step": n,
                "api": "CollisionEngine.check_states(numpy) -> pbp_check_states_host (pinned host buffers, copies overlapped with kernels)",
                "pinned_h2d_gbs_measured": h2d_gbs, "pcie_bound_value": h2d_gbs * 1e9 / 36.0 * world,
                "pageable_f64": {"value": e2e_f64_value, "unit": UNIT, "h2d_bytes_per_step": n * 9 * 4, "d2h_bytes_per_step": n,
                                 "api": "CollisionEngine.check_states(float64 pageable numpy) -> pbp_check_states_host_f64 (threaded narrowing into "
                                        "rotating pinned slices, overlapped with H2D and kernels): the call the drop-in functions make"}},
        "gpu_launches": stats["kernel_launches"],
        "clocks": clocks,
        "clocks_by_rank": clocks_by_rank,
        "roofline": roofline,
        "cpu_baseline": cpu_baseline,
        "parity": parity,
        "p_collision": float(hits[(args.steps - 1) % N_BATCHES].float().mean().item()),
        "timing": "headline region: K steps, one CUDA-graph launch each (memsets + five kernels), no per-kernel events; per-kernel durations from a second pass of K steps",
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

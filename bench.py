#!/usr/bin/env python
"""Headline benchmark: agent-steps/sec of a 65,536-cyclist TwoDBicycle open-plane crowd
(BASELINE.json metric), one process per GPU, agent-range sharding + payload all-gather.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W     # CPU oracle port, all host threads

Prints ONE JSON line on rank 0.  See DESIGN.md "Measurement" for how every field is obtained.
Environment knobs (tuning / diagnostics, never needed for the reported numbers): CSF_BENCH_N (crowd size),
CSF_BENCH_GRAPH=0 (kernel-by-kernel launches), CSF_BENCH_EXCHANGE=nccl|peer, CSF_BENCH_BALANCE=0,
CSF_BENCH_EMULATE_WORLD=k (time one shard of a k-way split on one GPU), CSF_BENCH_DEBUG=1 (per-step times).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_AGENTS = int(os.environ.get("CSF_BENCH_N", 65536))
FLOP_PER_PAIR = 76.0          # 68 FP32 flops + 8 special-function results, dense convention
BYTES_PER_AGENT_STEP = 124.0  # TwoDBicycle per-agent kernel, fp32 SoA
SEED = 1
METRIC = f"agent-steps/sec, N={N_AGENTS:,} TwoDBicycle"
WORKLOAD = f"{N_AGENTS}-cyclist TwoDBicycle open-plane crowd (SURVEY 8d recipe, seed {SEED}, 4 m spacing)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary BASELINE config-4 measurement")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md recipe) during the timed region
# ------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------
# CPU arm: the oracle port on host cores (bounded sample of the same workload)
# ------------------------------------------------------------------------------------------
def cpu_oracle_run(steps, warmup, sample_agents=None):
    """Times the CPU oracle on a bounded sample: `sample_agents` agents of the N-agent crowd are
    stepped per CPU step; each of them interacts with all N sources (full per-agent cost)."""
    from oracle import cpu_port
    return cpu_port.timed_sample(N_AGENTS, SEED, steps, warmup, sample_agents)


def main():
    args = parse()
    # stdout carries exactly ONE JSON line: libraries that chat on fd 1 (the NCCL version banner) are
    # sent to stderr for the duration of the run
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    try:
        return run(args, real_stdout)
    finally:
        sys.stdout.flush()
        real_stdout.flush()


def run(args, out):
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))

    if args.impl == "reference":
        if rank != 0:
            return 0
        # torchrun pins OMP_NUM_THREADS=1 for its children: the CPU arm uses every host thread
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
        r = cpu_oracle_run(args.steps, args.warmup)
        cfg = {"workload": WORKLOAD + f"; CPU arm: a bounded sample of it -- {r.get('sample_agents')} of the "
                           f"{N_AGENTS} agents are stepped per CPU step, each against all {N_AGENTS} sources (the "
                           "full per-agent cost); value = sampled agent-steps / wall time, ms_per_step = measured "
                           "time of one sample step",
               "n_agents": N_AGENTS, "sample_agents": r.get("sample_agents"),
               "ms_per_full_step_extrapolated": r.get("ms_per_full_step_extrapolated"),
               "implementation": "C/OpenMP restatement of the reference's TwoDBicycle stepping path (oracle/csf_oracle_c.c),"
                                 " every host thread; the reference's own Python is O(N^4) and cannot run N > ~200"}
        if os.environ.get("CSF_BENCH_NUMPY_ORACLE", "1") != "0":
            from oracle import cpu_port
            try:        # the restated numpy oracle on a FULL 4,096-cyclist crowd (config 3), one core
                cfg["restated_numpy_oracle_full_crowd"] = cpu_port.timed_numpy_oracle_full(4096, SEED, steps=2, warmup=1)
            except Exception as e:      # noqa: BLE001
                cfg["restated_numpy_oracle_full_crowd"] = {"error": str(e)[:200]}
        line = {
            "impl": "reference", "metric": METRIC, "value": r["value"], "unit": "agent-steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": r["value"], "unit": "agent-steps/s", "cores": r["cores"],
                             "kind": r["kind"], "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": "agent-steps/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0},
        }
        print(json.dumps(line), file=out, flush=True)
        return 0

    import torch
    import torch.distributed as dist
    from cyclistsocialforce_b200 import _lib, parameters as P
    from cyclistsocialforce_b200.distributed import (PayloadExchange, PeerExchange, balanced_bounds, neighbour_work,
                                                     shard_bounds)
    from cyclistsocialforce_b200.engine import AgentGroup, Engine
    from cyclistsocialforce_b200.synthetic import queues_with_start, spatial_order, synthetic_crowd

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    lib = _lib.load()

    # ---- the crowd, sharded by agent range -------------------------------------------------
    s0, q = synthetic_crowd(N_AGENTS, seed=SEED)
    # spatial domain decomposition: agents are numbered along a Hilbert curve through their initial
    # positions, so that the contiguous agent range of a rank is a compact region of the plane
    order = spatial_order(s0[:, 0], s0[:, 1])
    s0, q = s0[order], q[order]
    queues = queues_with_start(s0, q)
    # CSF_BENCH_EMULATE_WORLD=8: time ONE rank's shard of an 8-way split on a single GPU (tuning aid;
    # the payload of the other shards stays frozen, no exchange) -- never used for reported numbers
    emu = int(os.environ.get("CSF_BENCH_EMULATE_WORLD", 0))
    # ranks get contiguous ranges of equal estimated pair work (agents at the rim of the crowd have
    # fewer neighbours): every rank computes the same bounds from the same initial positions
    nparts = emu if emu else world
    balance = os.environ.get("CSF_BENCH_BALANCE", "1") != "0"
    bounds = (balanced_bounds(neighbour_work(s0[:, 0], s0[:, 1], 160.0), nparts) if (balance and nparts > 1)
              else shard_bounds(N_AGENTS, nparts))
    lo, hi = bounds[0] if emu else bounds[rank]
    # Q-format frame of the f32 payload, the same on every rank: centred on the crowd and its destinations
    origin, extent = P.payload_frame([s0[:, :2], q[..., :2]])
    group = AgentGroup("twod", s0[lo:hi], P.InvPendulumBicycleParameters(), destqueues=list(queues[lo:hi]),
                       dtype=torch.float32, device=dev)
    exchange_kind = os.environ.get("CSF_BENCH_EXCHANGE", "peer") if (world > 1 and not emu) else "none"
    xb = None if emu else bounds
    exch = None
    if exchange_kind == "peer":
        try:
            exch = PeerExchange(N_AGENTS, rank, world, torch.float32, dev, bounds=xb)
        except RuntimeError as e:      # raised on every rank alike (see PeerExchange.__init__)
            print(f"[bench] peer-memory exchange unavailable ({e}); using the NCCL all-gather", file=sys.stderr)
            exchange_kind = "nccl"
    if exch is None:
        exch = PayloadExchange(N_AGENTS, rank, world, bounds=xb)
    if emu:
        full = AgentGroup("twod", s0, P.InvPendulumBicycleParameters(), destqueues=list(queues),
                          dtype=torch.float32, device=dev)
        e_full = Engine([full], dtype=torch.float32, device=dev, extent=extent, origin=origin, pair_mode="dense")
        frozen_payload = e_full.payload.clone()
        del e_full, full
    pair_mode = os.environ.get("CSF_PAIR_MODE", "tiled")
    use_graph = os.environ.get("CSF_BENCH_GRAPH", "1") != "0"
    eng = Engine([group], dtype=torch.float32, device=dev, extent=extent, origin=origin, n_global=N_AGENTS, global_offset=lo,
                 exchange=exch, pair_mode=pair_mode, count_pairs=True, graph=False)
    exch(eng.payload)
    if emu:
        eng.payload.copy_(frozen_payload)
    n_local = hi - lo

    def sync():
        torch.cuda.synchronize(dev)

    def barrier():
        if world > 1:
            dist.barrier()

    # ---- FP32 peak of this device (FFMA micro-benchmark; MEASURED_PEAKS.json has no fp32 figure) ----
    import ctypes as C
    sink = torch.zeros(1, dtype=torch.float32, device=dev)
    flops = C.c_double(0.0)
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    lib.csf_ffma_peak(2000, sink.data_ptr(), C.byref(flops), st)
    sync()
    best = 0.0
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        lib.csf_ffma_peak(20000, sink.data_ptr(), C.byref(flops), st)
        e1.record()
        sync()
        best = max(best, flops.value / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    fp32_peak_tflops = best

    # ---- L2 flush buffer (126 MB L2): written between timed iterations ---------------------
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    for _ in range(max(args.warmup, 3)):
        eng.step()
    sync()
    # pair evaluations actually executed per launch (the tiled kernel skips tiles outside the
    # target's field of view); counted once, outside the timed region
    executed_pairs = float(n_local) * float(N_AGENTS)
    if eng.tiled:
        eng.pair_stats.zero_()
        eng._pair_and_road()
        sync()
        executed_pairs = float(eng.pair_stats[0].item())
        eng._pair_calls -= 1
        if os.environ.get("CSF_BENCH_ROLES"):      # cycle accounting of a -DCSF_TILED_PROF build (tools/k1_roles.py)
            st_ = eng.pair_stats.cpu().numpy().astype(float)
            if st_[1] > 0:
                n_it = int(lib.csf_tiled_num_items(N_AGENTS, n_local, 4))
                tr = eng.pair_stats.cpu().numpy()[16:16 + 4 * n_it].reshape(-1, 4).astype(np.int64)
                np.save(os.path.join(ROOT, "gpurun_out", "k1_item_trace.npy"), tr)
                t0_ = tr[:, 0][tr[:, 0] > 0].min()
                dur = (tr[:, 2] - tr[:, 0]) / 1e3
                print("item trace: kernel span %.1f us; item (claim -> retire) us: mean %.1f median %.1f max %.1f; claim->publish "
                      "mean %.1f; last claim at %.1f us, last retire at %.1f us" % (
                          (tr[:, 2].max() - t0_) / 1e3, dur.mean(), np.median(dur), dur.max(),
                          ((tr[:, 1] - tr[:, 0]) / 1e3).mean(), (tr[:, 0].max() - t0_) / 1e3, (tr[:, 2].max() - t0_) / 1e3),
                      file=sys.stderr, flush=True)
                print("roles: evaluate waiting %.1f%%; filter waiting for a stage %.1f%%, for a free slot %.1f%%; producer "
                      "waiting %.1f%%; buffers %d stages %d units %d" % (100 * st_[2] / st_[1], 100 * st_[4] / st_[3],
                      100 * st_[5] / st_[3], 100 * st_[7] / st_[6], st_[8], st_[9], st_[10]), file=sys.stderr, flush=True)
    eng.pair_stats_ptr_backup = eng.pair_stats
    eng.pair_stats = None                      # no counting inside the timed region

    eng.use_graph = use_graph                  # step() = CUDA-graph replay of the step's kernels
    for _ in range(2):
        eng.step()                             # (captures the graph)
    sync()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    launches0 = eng.gpu_launches
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    sync()
    for k in range(args.steps):
        flush.fill_(k & 0xFF)                      # L2 flush, outside the per-step events
        a, c = ev[k]
        a.record()
        eng.step()                                 # K1 (+ tile build, partial-sum reduce), K2+K3 fused, all-gather
        c.record()
    sync()
    barrier()
    launches = eng.gpu_launches - launches0
    clocks = sampler.stop()
    step_ms = [a.elapsed_time(c) for a, c in ev]
    total_ms = sum(step_ms)
    if os.environ.get("CSF_BENCH_DEBUG"):
        print("per-step ms:", " ".join(f"{t:.3f}" for t in step_ms), file=sys.stderr, flush=True)
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    # per-kernel durations: events cannot be placed inside a graph replay, so the split of a step into
    # [tile build + block bounds (+ wait for the peers)] / K1 / [reduce + K2+K3 (+ exchange)] is taken in a second
    # pass of the same K steps launched kernel by kernel, with CUDA events between the launches
    eng.use_graph = False
    fused = eng._fused
    ev2 = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
    for k in range(args.steps):
        flush.fill_(k & 0xFF)
        e = ev2[k]
        if fused:
            eng._step_fused(mark=lambda i, e=e: e[i].record())
        else:
            e[0].record(); e[1].record()
            have_rep = eng._pair_and_road()
            e[2].record()
            eng._agent_step(have_rep)
            e[3].record()
    sync()
    barrier()
    prep_ms = [e[0].elapsed_time(e[1]) for e in ev2]
    pair_ms = [e[1].elapsed_time(e[2]) for e in ev2]
    agent_ms = [e[2].elapsed_time(e[3]) for e in ev2]
    ungraphed_ms = statistics.mean(e[0].elapsed_time(e[3]) for e in ev2)
    eng.check_status()
    if os.environ.get("CSF_BENCH_DEBUG"):
        print("second pass, pair kernel ms per step (first pair call of the pass: %d):" % (eng._pair_calls - args.steps),
              " ".join(f"{t:.3f}" for t in pair_ms), file=sys.stderr, flush=True)
        if eng.tiled and fused:
            eng.pair_stats = eng.pair_stats_ptr_backup
            eng.pair_stats.zero_()
            eng._pair_alone()
            sync()
            print("executed pairs per launch now: %.4g (before the timed steps: %.4g)" % (
                float(eng.pair_stats[0].item()), executed_pairs), file=sys.stderr, flush=True)
            eng.pair_stats = None
    # On a sharded crowd the intervals of that pass also hold the ranks' skew (waits for the peers, launches that
    # a host issues late): the pair kernel's own duration is taken from launches of [tile build, pair kernel]
    # alone on this rank's shard while every rank is quiescent (the pair kernel never talks to a peer)
    pair_timing = "second pass of the same K steps, launched kernel by kernel"
    if world > 1 and fused:
        barrier()
        sync()
        ev3 = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(min(args.steps, 30))]
        for k, e in enumerate(ev3):
            flush.fill_(k & 0xFF)
            eng._pair_alone(mark=lambda i, e=e: e[i].record())
        sync()
        barrier()
        pair_ms = [e[1].elapsed_time(e[2]) for e in ev3]
        pair_timing = ("%d launches of [tile build, pair kernel] alone on rank 0's shard after the timed steps, L2 flushed "
                       "before each; prepare / per-agent kernel (with their waits for the peers) from a second pass of "
                       "the same K steps launched kernel by kernel" % len(ev3))
    value = N_AGENTS * args.steps / (total_ms * 1e-3)
    ms_per_step = total_ms / args.steps

    # ---- end to end: host buffers in, host buffers out, every step --------------------------
    # the CSF state (x, y double; psi, v, delta float) travels as one pinned slab each way, the total force
    # as a second device->host copy
    host_in = group.state_slab.cpu().pin_memory()
    host_out = torch.empty_like(host_in).pin_memory()
    host_force = torch.empty((n_local, 2), dtype=torch.float32).pin_memory()
    h2d = host_in.numel()
    d2h = h2d + host_force.numel() * host_force.element_size()
    eng.use_graph = use_graph
    for _ in range(2):
        eng.step_host(host_in, host_out, host_force)
    barrier()
    sync()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        eng.step_host(host_in, host_out, host_force)
        host_in, host_out = host_out, host_in
    sync()
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = N_AGENTS * args.steps / e2e_s

    if rank == 0:
        pair_s = statistics.mean(pair_ms) * 1e-3
        pairs_per_launch = float(n_local) * float(N_AGENTS - 1)
        # achieved = flops the kernel actually executed (76 per evaluated pair); the dense-convention
        # figure (every ordered pair counted) is reported next to it
        achieved = executed_pairs * FLOP_PER_PAIR / pair_s / 1e12
        dense_equiv = pairs_per_launch * FLOP_PER_PAIR / pair_s / 1e12
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "pair_kernel_traffic.json")
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        agent_s = statistics.mean(agent_ms) * 1e-3
        line = {
            "metric": METRIC, "value": value, "unit": "agent-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "n_agents": N_AGENTS, "step": "CUDA-graph replay" if use_graph else "kernel-by-kernel launches", "partition": "contiguous agent ranges of a Hilbert order of the initial positions (spatial decomposition)" + (", ranges balanced by estimated neighbour count" if (balance and nparts > 1) else ""), "exchange": {"peer": "NVLink peer-memory stores + flags inside the step graph (csf_peer_*)", "nccl": "NCCL all_gather_into_tensor", "none": "none (1 GPU)"}[exchange_kind], "parallelism": f"agent-range x{world}" + (f" (EMULATED 1/{emu} shard, not a result)" if emu else ""), "pair_kernel": "tiled+culled" if eng.tiled else "dense",
                       "l2": "flushed between timed steps (256 MiB write)", "q_scale_m": eng.q_scale,
                       "pair_interactions_per_s": float(N_AGENTS) * (N_AGENTS - 1) * args.steps / (total_ms * 1e-3)},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "agent-steps/s", "h2d_bytes_per_step": h2d * world,
                    "d2h_bytes_per_step": d2h * world},
            "gpu_launches": launches,
            "roofline": {"bound": "fp32",
                         "kernel": ("pair_tiled_kernel<float>" if fused else "pair_tiled_kernel<float> (+tile build, partial reduce)") if eng.tiled
                         else "pair_kernel<float,2,512> (+partial reduce)",
                         "achieved": achieved, "dense_convention_tflops": dense_equiv,
                         "dense_convention_frac": dense_equiv / fp32_peak_tflops,
                         "executed_pairs_per_launch": executed_pairs,
                         "executed_pair_fraction": executed_pairs / (float(n_local) * float(N_AGENTS)),
                         "peak": fp32_peak_tflops, "unit": "TFLOP/s", "frac": achieved / fp32_peak_tflops,
                         "traffic": traffic, "peak_source": "csf_ffma_peak micro-benchmark in this run",
                         "flop_per_pair": FLOP_PER_PAIR, "pairs_per_launch": pairs_per_launch,
                         "kernel_ms": pair_s * 1e3, "share_of_step": pair_s * 1e3 / (ungraphed_ms if world == 1 else ms_per_step),
                         "prepare_kernel_ms": statistics.mean(prep_ms),
                         "timing": pair_timing + " (ms_per_step of that pass: %.4f)" % ungraphed_ms},
            "roofline_agent_kernel": {"bound": "hbm", "kernel": "agent_kernel<float,TWOD,STEP>" + (" (+ reduction of the pair kernel's partial sums" + (", payload exchange" if world > 1 else "") + ")" if fused else ""),
                                      "achieved": n_local * BYTES_PER_AGENT_STEP / agent_s / 1e9, "peak": hbm_peak,
                                      "unit": "GB/s", "frac": n_local * BYTES_PER_AGENT_STEP / agent_s / 1e9 / hbm_peak,
                                      "kernel_ms": agent_s * 1e3,
                                      "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650"},
        }
        if world == 1 and not emu and not args.no_extra and N_AGENTS == 65536:
            # secondary: BASELINE config 4 (65,536 independent 8-agent InvertedPendulumBicycle scenarios in one
            # batch, no communication) on this GPU -- the workload whose step is the per-agent kernel
            try:
                del eng
                torch.cuda.empty_cache()
                sys.path.insert(0, os.path.join(ROOT, "tools"))
                import bench_scenarios as bs
                c4 = bs.run_scenarios(65536, 8, steps=30, warmup=5, model="invpendulum", dev=dev, hbm_peak_gbs=hbm_peak)
                line["extra"] = {"config4": {
                    "metric": "agent-steps/sec, 65,536 independent 8-agent InvertedPendulumBicycle scenarios, 1 GPU",
                    "value": c4["n_agents"] * 30 / (c4["ms_total"] * 1e-3), "unit": "agent-steps/s",
                    "ms_per_step": c4["ms_per_step"], "steps": 30, "pair_kernel_ms": c4["pair_kernel_ms"],
                    "agent_kernel_ms": c4["agent_kernel_ms"], "roofline": c4.get("roofline"),
                    "config": {"workload": "BASELINE config 4 (one GPU's worth: all 65,536 scenarios)",
                               "step": "CUDA-graph replay", "dtype": "f32, dynamic state f64"}}}
            except Exception as ex:                      # the headline line must not depend on the extra
                line["extra"] = {"config4": {"error": repr(ex)[:200]}}
        if not args.no_cpu_baseline:
            r = cpu_oracle_run(steps=2, warmup=1)
            line["cpu_baseline"] = {"value": r["value"], "unit": "agent-steps/s", "cores": r["cores"],
                                    "kind": r["kind"], "sample": r["sample"]}
        print(json.dumps(line), file=out, flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

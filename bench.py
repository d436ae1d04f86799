#!/usr/bin/env python
"""bench.py -- SP+TM timesteps/sec with learning on (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

N=1 workload = BASELINE configs[1]: 2048 columns x 1024-bit input, 41 active
columns (2 %), 32 cells/column, learning on, example.py's input recipe (100
patterns of density 0.2, 5 % bit-flip noise per step).  N>1 = N independent
networks of that size, one per GPU (trivially partitioned, no collective).

One JSON line on stdout (rank 0).  `value` = device-timed throughput with the
inputs already in HBM (CUDA events around every step, L2 flushed before each
step so state comes from HBM); `e2e` = the same metric through
HierarchicalTemporalMemory.process with host inputs (H2D + D2H inside the timed
region); `roofline` = the dominant kernel against the measured HBM peak;
`cpu_baseline` = the NumPy oracle (port of the reference's path) on a host core.
`--impl reference` times that CPU port alone.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG2 = dict(input_dim=1024, column_dim=2048, cell_dim=32, active_columns=41,
            patterns=100, density=0.2, noise=0.05, seed=0)
METRIC = "SP+TM timesteps/sec (learn on)"
UNIT = "steps/s"


def make_inputs(cfg, steps, seed):
    g = np.random.default_rng(1000 + seed)
    base = g.random((cfg["patterns"], cfg["input_dim"])) < cfg["density"]
    flips = g.random((steps, cfg["input_dim"])) < cfg["noise"]
    return base[np.arange(steps) % cfg["patterns"]] ^ flips


def workload_name(cfg):
    return (f"cfg2: SP {cfg['column_dim']} columns x {cfg['input_dim']}-bit input, k={cfg['active_columns']} (2%), "
            f"TM {cfg['cell_dim']} cells/column, learning on")


# ------------------------------------------------------------------------------ CPU port
def cpu_port(cfg, steps, warmup, seed):
    """The oracle in its reference-literal mode (dense float64 compare per step,
    projections.py:18-21), single thread like the reference."""
    from oracle.htm_oracle import HTMOracle, OracleConfig

    xs = make_inputs(cfg, steps + warmup, seed)
    orc = HTMOracle(OracleConfig(cfg["input_dim"], cfg["column_dim"], cfg["cell_dim"], cfg["active_columns"]),
                    rng=np.random.RandomState(seed), overlap="dense")
    for t in range(warmup):
        orc.step(xs[t])
    t0 = time.perf_counter()
    for t in range(warmup, warmup + steps):
        orc.step(xs[t])
    dt = time.perf_counter() - t0
    return steps / dt, dt


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = CFG2
    steps, warmup = args.steps, max(args.warmup, 3)
    steps = min(steps, 3000)  # bounded sample: ~150 steps/s on one core
    value, dt = cpu_port(cfg, steps, min(warmup, 100), cfg["seed"])
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": min(warmup, 100), "ms_per_step": 1e3 / value, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64/f32/int", "data": "synthetic",
        "config": {"workload": workload_name(cfg), "sample": f"{steps} timesteps of the same input recipe"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port",
                         "sample": f"{steps} timesteps, NumPy oracle (reference-literal dense overlap), "
                                   f"{os.cpu_count()} host cores visible, path is single-threaded"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap,timestamp")

    def __init__(self, index):
        self.index = index
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i",
                 str(self.index)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def mark(self):
        """Start of the region whose samples count (the sampler itself is started early: nvidia-smi
        takes a few hundred ms to come up)."""
        self.t_mark = time.time()

    def stop(self):
        self.t_stop = time.time()
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        if os.path.getsize(self.path) == 0:  # the loop never got to print: one synchronous sample
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=20).stdout
                open(self.path, "w").write(out)
            except Exception:
                pass
        import datetime

        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = []
        for ln in open(self.path):
            f = [p.strip() for p in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                row = (float(f[1]), float(f[2]), f[5:9])
            except ValueError:
                continue
            ts = None
            if len(f) > 9:
                try:
                    ts = datetime.datetime.strptime(f[9], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                except ValueError:
                    ts = None
            rows.append((ts, row))
        t0 = getattr(self, "t_mark", None)
        inside = [r for ts, r in rows if ts is not None and t0 is not None and t0 - 0.05 <= ts <= self.t_stop + 0.05]
        for a, b, flags in (inside or [r for _, r in rows]):
            sm.append(a)
            smax.append(b)
            for name, v in zip(names, flags):
                if v.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.unlink(self.path)
        except OSError:
            pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------ roofline
sys.path.insert(0, os.path.join(ROOT, "tools"))
from cfg3_kernels import algorithmic_bytes, network_stats  # noqa: E402  (shared with tools/cfg3_kernels.py)


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed
    `ncu --set full` capture of this bench (profiles/r01_traffic.json), else None."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))[kernel]["dram_bytes_per_launch"]
    except Exception:
        return None


def nat_launches(eng):
    """Kernel launches one device-resident step issues (fused: 1; per-stage: 15 + ring fetch)."""
    from bithtm_b200 import _native as nat

    n = nat.lib.bh_step_launches(eng.ref, 1)
    return n if eng.ctx.fused_mode else n + 1


# ------------------------------------------------------------------------------ our arm
def ours(args):
    import torch
    import torch.distributed as dist

    import bithtm_b200 as bithtm

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg = CFG2
    K, W = args.steps, max(args.warmup, 3)
    seed = cfg["seed"] + rank  # independent networks on each GPU
    total = W + K

    def build(ring_len, rng_sync):
        np.random.seed(seed)
        return bithtm.HierarchicalTemporalMemory(cfg["input_dim"], cfg["column_dim"], cfg["cell_dim"],
                                                 cfg["active_columns"], rng_sync=rng_sync, ring_len=ring_len,
                                                 max_segments=1 << 17, fused=args.fused, fused_ctas=args.fused_ctas)

    xs = make_inputs(cfg, total, seed)
    sampler = ClockSampler(local)
    sampler.start()  # early: nvidia-smi takes a few hundred ms to deliver its first sample

    # ---------------- device-resident arm: inputs in an HBM ring, one CUDA graph per step
    htm = build(total, "lazy")
    eng = htm.engine
    htm.temporal_memory._rng.before(eng)  # upload np.random's MT19937 state once
    eng.load_ring(xs)
    graph1 = eng.graph(1, learning=True)
    launches = nat_launches(eng)
    exec_mode = {0: "one kernel per stage (15 launches/step)",
                 1: f"whole step in one kernel on a thread-block cluster of {eng.ctx.fused_ctas} CTAs",
                 2: f"whole step in one cooperative kernel, {eng.ctx.fused_ctas} CTAs"}[eng.ctx.fused_mode]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    sampler.mark()  # clocks are reported from the warm-up on (same load as the timed region)
    for _ in range(W):
        flush.fill_(1)
        eng.launch_graph(graph1, 1)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    for a, b in ev:
        flush.fill_(1)  # evict L2 so the step streams its state from HBM
        a.record()
        eng.launch_graph(graph1, 1)
        b.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop()
    step_ms = np.array([a.elapsed_time(b) for a, b in ev])
    dev_s = float(step_ms.sum()) / 1e3
    sc = eng.scalars()
    eng.check_status(sc[12])
    # warm (L2-resident) variant: graphs of 50 steps back to back, no flush
    per = 50
    graph50 = eng.graph(per, learning=True)
    reps = max(1, min(K, 2000) // per)
    # the ring wraps: inputs repeat, state keeps evolving
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        eng.launch_graph(graph50, per)
    e1.record()
    torch.cuda.synchronize()
    warm_ms_per_step = e0.elapsed_time(e1) / (reps * per)

    # ---------------- per-kernel timing of the same step (CUDA events after every launch)
    words = eng.buf["input_ring"][:eng.ctx.input_words]
    prof = {}
    n_prof = 40
    for _ in range(n_prof):
        flush.fill_(1)
        for name, ms in eng.profile_step(words, learning=True):
            prof[name] = prof.get(name, 0.0) + ms / n_prof
    stats = network_stats(eng)
    if eng.ctx.fused_mode:
        # the step IS one kernel: its launch duration is the per-step time of the timed region above
        # (CUDA events around each graph launch, L2 flushed before each), not the one-off profile launch
        prof = {next(iter(prof)): dev_s / K * 1e3}
    dominant = max(prof, key=prof.get)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    ab = algorithmic_bytes(cfg, dominant, stats)
    achieved = (ab / (prof[dominant] * 1e-3) / 1e9) if ab else None
    roofline = {"bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": (achieved / peak) if achieved else None, "traffic": ncu_traffic(dominant),
                "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback (B200_PROFILING.md)",
                "kernel_us": {k: round(v * 1e3, 2) for k, v in sorted(prof.items(), key=lambda kv: -kv[1])},
                "algorithmic_bytes_per_launch": ab, "state": stats,
                "note": ("cfg2's whole working set is ~3 MB: one step is ~2.6 MB of algorithmic traffic = 0.4 us of HBM "
                         "time, so the step is bound by its chain of dependent phases (9 barrier-separated phases), not "
                         "by bandwidth; the HBM-bound sizes of the same kernels are under roofline_hbm_kernels "
                         "(cfg3: whole step and per kernel)")}
    del htm, eng

    # ---------------- end-to-end arm: host inputs through the reference-facing API
    htm2 = build(0, "step")
    e2e_steps = min(K, 2000)
    for t in range(W):
        htm2.process(xs[t])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    bursts = 0
    for t in range(W, W + e2e_steps):
        sp_state, tm_state = htm2.process(xs[t])  # H2D input, step, D2H summary inside
        bursts += int(tm_state.active_column_bursting.sum())
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    h2d = htm2.engine.ctx.input_words * 4
    d2h = (4 + 4 * cfg["active_columns"] + 625) * 4

    # ---------------- max over ranks
    times = torch.tensor([dev_s, e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dev_s, e2e_s = float(times[0]), float(times[1])

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        v, dt = cpu_port(cfg, 1500, 100, cfg["seed"])
        cpu = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": f"1500 timesteps of the same workload, NumPy oracle in reference-literal mode "
                         f"({dt:.1f} s), single thread as the reference; {os.cpu_count()} host cores visible"}
    hbm = None
    if rank == 0 and world == 1 and not args.no_hbm:
        # the HBM-bound kernels of the path at cfg3 size (65536 x 16384): the cfg2 step is
        # latency-bound, so kernel quality against the HBM roofline is shown here
        try:
            import cfg3_kernels

            torch.cuda.empty_cache()
            hbm = cfg3_kernels.measure(65536, 16384, 250, 20)
        except Exception as e:  # never lose the headline line
            hbm = {"error": repr(e)}
    sharded = None
    if world > 1 and not args.no_hbm:
        # the multi-GPU path of ONE network: cfg3 sharded over all ranks, SP by column, TM by
        # segment id, one cooperative kernel per shard with in-kernel exchanges over NVLink peer memory
        try:
            import cfg3_sharded

            torch.cuda.empty_cache()
            sharded = cfg3_sharded.measure(300, 65536, 16384, "fused")
        except Exception as e:  # never lose the headline line
            sharded = {"error": repr(e)}
    batched = None
    if not args.no_hbm:
        # independent streams side by side on every GPU (BASELINE configs[3]: 1024 streams over 8 GPUs = 128 per
        # GPU, partitioned trivially): aggregate throughput, max-over-ranks time
        try:
            import stream_batch

            torch.cuda.empty_cache()
            if world > 1:
                dist.barrier()
            batched = stream_batch.measure(128, 4, 400, 200, seed0=1000 * rank, threads=512)
            if world > 1:
                ms = torch.tensor([batched["ms"]], dtype=torch.float64, device="cuda")
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
                batched["ms"] = float(ms)
                batched["streams"] *= world
                batched["aggregate_steps_per_s"] = batched["streams"] * batched["steps_per_stream"] / (batched["ms"] * 1e-3)
            batched["note"] = (f"{128 * world} independent cfg2 networks, 128 per GPU, one 4-CTA x 512-thread cluster kernel each (2 CTAs per SM), one "
                               "CUDA graph of 50 steps x 128 launches per GPU; L2-resident; no collective")
        except Exception as e:
            batched = {"error": repr(e)}
    shared_mask = None
    if rank == 0 and world == 1 and not args.no_hbm:
        # many inputs against ONE connected mask (inference over streams that share a spatial pooler): the
        # overlap as an int8 tensor-core contraction next to the popcount kernel, both exact
        try:
            import batched_overlap

            torch.cuda.empty_cache()
            shared_mask = {"cfg4_shape": batched_overlap.measure(1024, 2048, 1024, 100),
                           "cfg3_shape": batched_overlap.measure(256, 65536, 16384, 10)}
        except Exception as e:
            shared_mask = {"error": repr(e)}
    if rank == 0:
        launches_per_step = launches
        line = {
            "metric": METRIC, "value": world * K / dev_s, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": dev_s / K * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64 permanence / f32 / u32 bit-words", "data": "synthetic",
            "config": {"workload": workload_name(cfg), "parallelism": f"{world} independent network(s), 1 per GPU",
                       "execution": exec_mode,
                       "l2": "flushed before every timed step (256 MiB write); per-step CUDA events summed",
                       "inputs": "device-resident ring, one CUDA graph launch per step"},
            "l2_resident": {"value": world * 1e3 / warm_ms_per_step, "unit": UNIT, "ms_per_step": warm_ms_per_step,
                            "note": "no flush, graphs of 50 steps back to back (state stays in L2 as in real use)"},
            "e2e": {"value": world * e2e_steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                    "note": "HierarchicalTemporalMemory.process(host bool array), np.random kept in lock-step"},
            "gpu_launches": launches_per_step * K,
            "roofline": roofline, "roofline_hbm_kernels": hbm, "streams_batched": batched, "sharded_cfg3": sharded,
            "shared_mask_batched_overlap": shared_mask,
            "cpu_baseline": cpu,
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=200)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-hbm", action="store_true", help="skip the cfg3-size HBM-bound kernel measurements")
    ap.add_argument("--fused", default="auto", choices=["auto", "cluster", "grid", "off"],
                    help="execution mode of the step (default: one kernel on a thread-block cluster at this size)")
    ap.add_argument("--fused-ctas", type=int, default=None, help="CTAs of the fused kernel")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        ours(args)


if __name__ == "__main__":
    main()

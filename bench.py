#!/usr/bin/env python
"""bench.py -- SP+TM timesteps/sec with learning on (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload = BASELINE configs[2] (the configuration the metric "at 1/2/4/8 B200" is quoted on):
ONE network of 65536 columns x 16384-bit input, k = 1311 active columns (2 %), 32 cells per column,
learning on, example.py's input recipe (50 patterns of density 0.2, 5 % bit-flip noise per step).
N = 1: the whole step is one cooperative kernel.  N > 1 (torchrun, one rank per GPU): the SAME
network sharded over the N ranks -- spatial pooler by column, synapse rows by segment id, the
exchanges done inside the step kernel over NVLink peer memory -- i.e. strong scaling.

One JSON line on stdout (rank 0).  `value` = device-timed throughput of the K timed steps, inputs
already in HBM (device input ring), max over ranks; `e2e` = the same metric through
HierarchicalTemporalMemory.process(host array) with the H2D / D2H copies inside the timed region;
`roofline` = the step kernel against the measured HBM peak; `cpu_baseline` = the reference's own
NumPy path (oracle/_ref when present, else the NumPy port) on a host core.  `--impl reference`
times that CPU path alone.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

CFG3 = dict(input_dim=16384, column_dim=65536, cell_dim=32, active_columns=1311,
            patterns=50, density=0.2, noise=0.05, seed=0)
# BASELINE configs[1] (latency-bound; an extra of the N=1 line, and the shape of the independent streams)
CFG2 = dict(input_dim=1024, column_dim=2048, cell_dim=32, active_columns=41,
            patterns=100, density=0.2, noise=0.05, seed=0)
WORKLOADS = {"cfg3": CFG3, "cfg2": CFG2}  # cfg2: smoke runs of this script (tests/test_cpu.py)
METRIC = "SP+TM timesteps/sec (learn on)"
UNIT = "steps/s"
PERM_BLOCKS = 8  # the permanence matrix is drawn in 8 row blocks, so that every N in {1,2,4,8} builds the SAME network


def make_inputs(cfg, steps, seed):
    g = np.random.default_rng(1000 + seed)
    base = g.random((cfg["patterns"], cfg["input_dim"])) < cfg["density"]
    flips = g.random((steps, cfg["input_dim"])) < cfg["noise"]
    return base[np.arange(steps) % cfg["patterns"]] ^ flips


def workload_config(cfg):
    """The `config` object: identical in both arms (ours and --impl reference)."""
    name = "cfg3" if cfg is CFG3 else "cfg2"
    return {"workload": (f"{name}: ONE SP+TM network, {cfg['column_dim']} columns x {cfg['input_dim']}-bit input, "
                         f"k={cfg['active_columns']} (2%), {cfg['cell_dim']} cells/column, learning on"),
            "inputs": f"{cfg['patterns']} patterns of density {cfg['density']}, {int(cfg['noise'] * 100)}% bit-flip "
                      "noise per step (example.py's recipe), synthetic"}


# ------------------------------------------------------------------------------ CPU arm
def _reference_package():
    """The unmodified reference package if the build step copied it next to the oracle
    (oracle/Makefile `ref`; git-ignored, travels with the snapshot), else None."""
    path = os.path.join(ROOT, "oracle", "_ref")
    if os.path.isdir(os.path.join(path, "bithtm")):
        if path not in sys.path:
            sys.path.insert(0, path)
        try:
            import bithtm  # noqa: F401

            return sys.modules["bithtm"]
        except Exception:
            return None
    return None


def cpu_path(cfg, steps, warmup, seed):
    """Times `steps` timesteps of the workload on one host core (the reference's path is
    single-threaded NumPy): the reference's own classes when available, else the oracle in
    reference-literal mode.  Returns (steps/s, seconds, kind, setup seconds)."""
    xs = make_inputs(cfg, steps + warmup, seed)
    t_setup = time.perf_counter()
    np.random.seed(seed)
    ref = _reference_package()
    if ref is not None:
        htm = ref.HierarchicalTemporalMemory(cfg["input_dim"], cfg["column_dim"], cfg["cell_dim"],
                                             cfg["active_columns"])
        step = htm.process
        kind = "reference"
    else:
        from oracle.htm_oracle import HTMOracle, OracleConfig

        orc = HTMOracle(OracleConfig(cfg["input_dim"], cfg["column_dim"], cfg["cell_dim"], cfg["active_columns"]),
                        rng=np.random.RandomState(seed), overlap="dense")
        step = orc.step
        kind = "port"
    t_setup = time.perf_counter() - t_setup
    for t in range(warmup):
        step(xs[t])
    t0 = time.perf_counter()
    for t in range(warmup, warmup + steps):
        step(xs[t])
    dt = time.perf_counter() - t0
    return steps / dt, dt, kind, t_setup


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = WORKLOADS[args.workload]
    # bounded sample: one timestep of cfg3 is ~3 s of single-threaded NumPy (SP overlap over the 8 GiB
    # float64 permanence, projections.py:18-21) and the constructor ~45 s (projections.py:16)
    big = cfg["column_dim"] * cfg["input_dim"] > (1 << 28)
    steps, warmup = max(1, min(args.steps, 5 if big else 3000)), max(0, min(args.warmup, 1 if big else 100))
    value, dt, kind, t_setup = cpu_path(cfg, steps, warmup, cfg["seed"])
    sample = (f"{steps} timesteps (+{warmup} warm-up) of the same workload from a freshly constructed network "
              f"(construction {t_setup:.0f} s, untimed); {os.cpu_count()} host cores visible, the path is "
              "single-threaded NumPy")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": 1e3 / value, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64 permanence / f32 / int64", "data": "synthetic",
        "config": workload_config(cfg),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind, "sample": sample},
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
from cfg3_kernels import algorithmic_bytes, network_stats  # noqa: E402  (shared with tools/cfg3_kernels.py)


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed
    `ncu --set full` capture of this bench (profiles/r02_traffic.json, else round 1's), else None."""
    for name in ("r02_traffic.json", "r01_traffic.json"):
        try:
            return json.load(open(os.path.join(ROOT, "profiles", name)))[kernel]["dram_bytes_per_launch"]
        except Exception:
            continue
    return None


def hbm_peak():
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(peaks["hbm_gbs"]), "MEASURED_PEAKS.json (burst copy bandwidth)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------ the network
def build_network(cfg, rank, world, ring_len, rng_sync, fused=None):
    """cfg3 as ONE network; with world > 1 this rank's shard of it.  The float64 permanence is drawn on the
    device in PERM_BLOCKS row blocks with per-block seeds, so every world size builds the same matrix
    (performance run: np.random.randn on the host takes 45 s for this size; parity runs are tests/)."""
    import torch

    import bithtm_b200 as bithtm
    from bithtm_b200.projections import DenseProjection

    C, I, c, k = cfg["column_dim"], cfg["input_dim"], cfg["cell_dim"], cfg["active_columns"]
    rows = C // world
    blk = C // PERM_BLOCKS
    perm = torch.empty(rows, I, dtype=torch.float64, device="cuda")
    for b in range(PERM_BLOCKS):
        lo = b * blk - rank * rows
        if lo < 0 or lo >= rows:
            continue
        gen = torch.Generator(device="cuda")
        gen.manual_seed(4321 + b)
        perm[lo:lo + blk] = torch.randn(blk, I, dtype=torch.float64, device="cuda", generator=gen) * 0.1
    np.random.seed(cfg["seed"])
    sp = bithtm.SpatialPooler(I, C, k, proximal_projection=DenseProjection(I, C, permanence=perm))
    if fused is None:
        fused = "shard" if world > 1 else ("grid" if cfg is CFG3 else "auto")  # (small networks: the cluster kernel)
    htm = bithtm.HierarchicalTemporalMemory(I, C, c, k, spatial_pooler=sp, rng_sync=rng_sync, ring_len=ring_len,
                                            column_shard=True if world > 1 else None,
                                            max_segments=min(1 << 21, 8 * C * c), max_synapses_per_segment=64,
                                            fused=fused)
    del perm
    sp.proximal_projection._host_permanence = None
    torch.cuda.empty_cache()
    return htm


def run_ring(eng, n, per=50):
    """n steps from the device input ring, at most `per` per launch; returns the number of launches."""
    launches = 0
    while n > 0:
        m = min(n, per)
        eng.launch_graph(eng.graph(m, learning=True), m)
        n -= m
        launches += 1
    return launches


def state_check(htm):
    """Cheap size-independent check value of the learned state: equal for every world size (the sharded network
    IS the single-GPU network)."""
    eng = htm.engine
    sc = eng.scalars()
    cur = (int(sc[0]) - 1) & 1
    k = eng.k
    cols = eng.buf["active_cols"][cur * k:(cur + 1) * k].cpu().numpy()
    nseg = eng.buf["cell_nseg"].cpu().numpy()
    return {"steps": int(sc[0]), "n_segments": int(sc[2]), "matching": int(sc[4]),
            "active_columns_crc32": zlib.crc32(cols.tobytes()), "segments_per_cell_crc32": zlib.crc32(nseg.tobytes())}


def sharded_parity(world, rank, steps=300):
    """The `mid` golden trace (recorded from the unmodified reference) run sharded over all ranks with the fused
    shard kernel: every step's digest must equal the reference's.  Collective."""
    import torch
    import torch.distributed as dist

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import bithtm_b200 as bithtm
    from helpers import golden_inputs, gpu_record, load_golden, step_digest

    info = load_golden("mid")
    xs = golden_inputs(info, steps)
    np.random.seed(info["seed"])
    htm = bithtm.HierarchicalTemporalMemory(info["I"], info["C"], info["c"], info["k"], column_shard=True,
                                            fused="shard", fused_ctas=32)
    bad = -1
    for t in range(steps):
        sp_state, tm_state = htm.process(xs[t])
        if step_digest(**gpu_record(htm, sp_state, tm_state)) != int(info["g"]["digests"][t]) and bad < 0:
            bad = t
    flag = torch.tensor([bad], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MAX)
    del htm
    torch.cuda.empty_cache()
    worst = int(flag.item())
    if worst >= 0:
        return f"MISMATCH vs reference trace at step {worst} (mid, {world} ranks)"
    return f"bit-exact vs reference trace (mid: 512 columns x 256 inputs, {steps} steps, sharded over {world} ranks)"


# ------------------------------------------------------------------------------ our arm
def ours(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg = WORKLOADS[args.workload]
    K, W, P = args.steps, max(args.warmup, 3), args.pretrain
    ring_len = 2 * cfg["patterns"]
    xs = make_inputs(cfg, ring_len, cfg["seed"])
    sampler = ClockSampler(local)
    sampler.start()  # early: nvidia-smi takes a few hundred ms to deliver its first sample

    parity = None
    if world > 1 and not args.no_parity:
        parity = sharded_parity(world, rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_ring(htm, warm, steps):
        eng = htm.engine
        run_ring(eng, warm)
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        launches = run_ring(eng, steps)
        b.record()
        barrier()
        ms = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        eng.check_status()
        return float(ms) * 1e-3, launches

    # ---------------- the network as a fresh one: the first W + K steps of its life (all columns burst, every
    # winner cell grows a new segment, every learning row draws its priorities)
    htm = build_network(cfg, rank, world, ring_len, "step")
    eng = htm.engine
    tm = htm.temporal_memory
    tm._rng.before(eng)  # upload np.random's MT19937 state once
    eng.load_ring(xs)
    scratch_s, _ = timed_ring(htm, W, K)
    scratch_stats = network_stats(eng)
    # ---------------- steady state: P more steps of setup, then W warm-up and the K timed steps
    run_ring(eng, max(0, P - W - K))
    sampler.mark()
    dev_s, launches = timed_ring(htm, W, K)
    clocks = sampler.stop()
    stats = network_stats(eng)
    check = state_check(htm)
    eng_pipe, eng_mode = int(eng.ctx.pipe_ctas), int(eng.ctx.fused_mode)
    exec_mode = (f"whole step = one cooperative kernel per shard ({eng.ctx.fused_ctas} CTAs), both exchanges inside it "
                 f"over NVLink peer memory ({getattr(htm, 'exchange_transport', 'local')})" if world > 1 else
                 f"whole step = one cooperative kernel ({eng.ctx.fused_ctas} CTAs"
                 f"{', one thread-block cluster' if eng.ctx.fused_mode == 1 else ''})")
    if eng.ctx.pipe_ctas > 0:
        exec_mode += (f"; two-pipeline schedule in multi-step launches: the spatial pooler of step s+1 on "
                      f"{eng.ctx.fused_ctas - eng.ctx.pipe_ctas} CTAs beside the temporal memory of step s on {eng.ctx.pipe_ctas}")

    # phase split of the last step (globaltimer stamps of CTA 0, rank 0)
    phases = None
    try:
        import phase_names

        phases = phase_names.read(eng)
        if world > 1 and eng.ctx.xch_ll:
            mine = {"phases": phases, "exchange_phases": phase_names.read_ll(eng)}
            every = [None] * world
            dist.all_gather_object(every, mine)
            phases = {f"rank{r}": v for r, v in enumerate(every)}
    except Exception:
        pass

    # ---------------- roofline of the step kernel (this rank's share of the algorithmic bytes)
    peak, peak_src = hbm_peak()
    piped = eng_pipe > 0
    kernel = ("step_shard_pipe" if piped else "step_shard") if world > 1 else ("step_pipe" if piped else "step_fused_grid")
    if world == 1 and eng_mode == 1:
        kernel = "step_fused_cluster"
    ab_total = algorithmic_bytes(cfg, "step_fused_grid", stats)
    ab = ab_total / world
    step_s = dev_s / K
    achieved = ab / step_s / 1e9
    roofline = {"bound": "hbm", "kernel": {"step_fused_grid": "k_step_fused<2>", "step_fused_cluster": "k_step_fused<1>"}.get(kernel, "k_" + kernel),
                "achieved": achieved,
                "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic(kernel) if world == 1 else None,  # (ncu profiles one GPU)
                "peak_source": peak_src, "algorithmic_bytes_per_launch_step": ab,
                "algorithmic_bytes_whole_network": ab_total, "state": stats,
                "note": "per GPU: the step's algorithmic bytes (SURVEY 8d: SP + TM, from the live network state) / "
                        "number of GPUs / measured step time; one launch runs up to 50 steps, bytes and time are per step"}

    # ---------------- end to end: host inputs through the reference-facing API, same network
    tm.sync_rng()  # np.random continues from the device's stream position
    e2e_steps = max(1, min(K, 200))
    for t in range(3):
        htm.process(xs[t])
    barrier()
    t0 = time.perf_counter()
    bursts = 0
    for t in range(e2e_steps):
        sp_state, tm_state = htm.process(xs[(3 + t) % ring_len])  # H2D input, step, D2H summary inside
        bursts += int(tm_state.active_column_bursting.sum())
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    times = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    e2e_s = float(times[0])
    h2d = eng.ctx.input_words * 4
    d2h = (4 + 4 * cfg["active_columns"] + 625) * 4
    eng.check_status()
    del htm, eng, tm
    torch.cuda.empty_cache()

    extras = {}
    if rank == 0 and world == 1 and not args.no_cpu:
        v, dt, kind, t_setup = cpu_path(cfg, 3, 1, cfg["seed"])
        extras["cpu_baseline"] = {
            "value": v, "unit": UNIT, "cores": 1, "kind": kind,
            "sample": f"3 timesteps (+1 warm-up) of the same workload from a freshly constructed network ({dt:.1f} s; "
                      f"construction {t_setup:.0f} s, untimed); single thread as the reference; {os.cpu_count()} host "
                      "cores visible"}
    if not args.no_extras:
        extras.update(extra_measurements(args, rank, world))
    if rank == 0:
        line = {
            "metric": METRIC, "value": K / dev_s, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": dev_s / K * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64 permanence / f32 / u32 bit-words", "data": "synthetic",
            "config": workload_config(cfg),
            "execution": exec_mode,
            "regime": (f"steady state: the network ran {max(P, W + K)} timesteps of the same input stream before the "
                       f"{W} warm-up steps (setup); `from_scratch` is the same measurement on the first {W}+{K} steps "
                       "of a fresh network"),
            "l2": ("no flush needed: one step streams > 500 MB (connected mask 134 MB + 1311 random 128 KiB permanence "
                   "rows + the segment store), 4x the 126 MB L2") if cfg is CFG3 else
                  ("NOT flushed: this small network's working set (3 MB) stays in L2 between the steps of a launch; "
                   "the flushed measurement of this size is `latency_cfg2` of the default run"),
            "inputs": "device-resident ring of 100 packed inputs, up to 50 steps per kernel launch",
            "from_scratch": {"value": K / scratch_s, "unit": UNIT, "ms_per_step": scratch_s / K * 1e3,
                             "state": scratch_stats},
            "e2e": {"value": e2e_steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps,
                    "note": "HierarchicalTemporalMemory.process(host bool array) per step, np.random kept in lock-step; "
                            "every rank feeds the same input"},
            "gpu_launches": launches,
            "roofline": roofline, "phases_us_last_step": phases, "state_check": check,
            "parity": parity,
            "cpu_baseline": extras.pop("cpu_baseline", None),
            "clocks": clocks,
        }
        line.update(extras)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def extra_measurements(args, rank, world):
    """Measurements next to the headline: independent streams (BASELINE configs[3]), and on one GPU the
    latency-bound cfg2 network (BASELINE configs[1]), the per-kernel HBM fractions at cfg3 size and the
    shared-mask batched overlap.  Failures never cost the headline line."""
    import torch
    import torch.distributed as dist

    out = {}
    try:  # 128 independent cfg2 streams per GPU (cfg4: 1024 streams over 8 GPUs), no collective
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
        batched["note"] = (f"{128 * world} independent cfg2 networks, 128 per GPU, one 4-CTA x 512-thread cluster kernel "
                           "each (2 CTAs per SM), one CUDA graph of 50 steps x 128 launches per GPU; L2-resident; "
                           "no collective")
        out["streams_batched"] = batched
    except Exception as e:
        out["streams_batched"] = {"error": repr(e)}
    if rank != 0 or world != 1:
        return out
    try:
        import cfg2_latency

        torch.cuda.empty_cache()
        out["latency_cfg2"] = cfg2_latency.measure(CFG2, 400, 100)
    except Exception as e:
        out["latency_cfg2"] = {"error": repr(e)}
    try:
        import cfg3_kernels

        torch.cuda.empty_cache()
        out["roofline_hbm_kernels"] = cfg3_kernels.measure(65536, 16384, 300, 20)
    except Exception as e:
        out["roofline_hbm_kernels"] = {"error": repr(e)}
    try:
        import batched_overlap

        torch.cuda.empty_cache()
        out["shared_mask_batched_overlap"] = {"cfg4_shape": batched_overlap.measure(1024, 2048, 1024, 100),
                                              "cfg3_shape": batched_overlap.measure(256, 65536, 16384, 10)}
    except Exception as e:
        out["shared_mask_batched_overlap"] = {"error": repr(e)}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pretrain", type=int, default=300,
                    help="timesteps the network has run before the warm-up (setup; steady-state regime)")
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS),
                    help="cfg3 = the benchmark (BASELINE configs[2]); cfg2 = small network for smoke runs of this script")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extras", action="store_true", help="headline line only (no cfg2 / stream / per-kernel extras)")
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the sharded golden-trace parity run")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        ours(args)


if __name__ == "__main__":
    main()

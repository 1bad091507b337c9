#!/usr/bin/env python
"""bench.py — agent-steps/s of the MAPF hot path (step + observe) on B200, next to the CPU path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one joint step + one observation build over all worlds of the rank (mapf_step_observe, one launch):
the rollout loop's per-step env work (runner.py:64-100) with the policy excluded.
Workload (N=1) = BASELINE.json configs[2]: 65 536 lockstep 40x40 worlds, 32 agents, obstacle density U[0,0.3],
uniform random actions.  Weak scaling: every rank owns 65 536 worlds; worlds never communicate.

Printed JSON (rank 0): value = whole-job agent-steps/s with inputs resident in HBM; `e2e` = the same through the
host-buffer C-ABI call (actions from pinned host memory, per-agent results read back to the host every step);
`roofline` for the dominant kernel (the fused step_observe_kernel) from CUDA events inside the timed region; `cpu_baseline` = the C port
of the reference's algorithm (oracle/) on this box's host cores.  `--impl reference` times that CPU path alone.
BENCH_E2E_DEBUG=1 prints the wall time of every e2e call on stderr.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = "65536 batched 40x40 worlds, 32 agents, density 0-0.3, step+observe (BASELINE.json configs[2])"
H = WD = 40
N_AGENTS = 32
FOV, CH = 9, 6


def algorithmic_bytes(n_agents=N_AGENTS, h=H, wd=WD, c=CH, f=FOV):
    """SURVEY.md §8d, per agent-step: step ~ 46 + (H*Wd+8)/N, observe ~ 16 + 4*C*F^2 + (H*Wd+8)/N + 8, and for the
    fused step+observe launch B = 62 + 4*C*F^2 + (H*Wd+8)/N (the world's map and human are read once)."""
    shared = (h * wd + 8) / n_agents
    step = 46 + shared
    observe = 16 + 4 * c * f * f + shared + 8
    fused = 62 + 4 * c * f * f + shared
    return step, observe, fused


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` on this workload, from the committed
    `ncu --set full` capture (profiles/traffic.json, written by tools/ncu_traffic.py); None if not captured."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        d = json.load(open(p))
        return float(d[kernel]["dram_bytes_per_launch"])
    except Exception:
        return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region.  NVML in-process (nvidia_ml_py) every 2 ms from a
    background thread; falls back to an `nvidia-smi -lms` child when NVML is unavailable.  Start it before the warm-up
    (`start()`), bracket the timed region with `mark_begin()` / `mark_end()`; `summary()` uses the samples in between."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc, self.thread = index, [], None, None
        self.stop = threading.Event()
        self.t0 = self.t1 = None
        self.source = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            idx = self.index
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                ent = vis.split(",")[idx].strip()
                idx = int(ent) if ent.isdigit() else idx
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            mx = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            R = pynvml
            bits = (("hw_slowdown", getattr(R, "nvmlClocksEventReasonHwSlowdown", 0x8)),
                    ("hw_thermal_slowdown", getattr(R, "nvmlClocksEventReasonHwThermalSlowdown", 0x40)),
                    ("sw_thermal_slowdown", getattr(R, "nvmlClocksEventReasonSwThermalSlowdown", 0x20)),
                    ("sw_power_cap", getattr(R, "nvmlClocksEventReasonSwPowerCap", 0x4)))
            get_reasons = getattr(R, "nvmlDeviceGetCurrentClocksEventReasons", None) or R.nvmlDeviceGetCurrentClocksThrottleReasons

            def loop():
                while not self.stop.is_set():
                    try:
                        sm = float(R.nvmlDeviceGetClockInfo(h, R.NVML_CLOCK_SM))
                        rs = int(get_reasons(h))
                        self.rows.append((time.perf_counter(), sm, mx, [n for n, b in bits if rs & b]))
                    except Exception:
                        pass
                    time.sleep(0.002)
            self.thread = threading.Thread(target=loop, daemon=True)
            self.thread.start()
            self.source = "nvml"
        except Exception:
            try:
                self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                              "--format=csv,noheader,nounits", "-lms", "20"],
                                             stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
                self.thread = threading.Thread(target=self._read_smi, daemon=True)
                self.thread.start()
                self.source = "nvidia-smi"
            except Exception:
                self.proc = None
        return self

    def _read_smi(self):
        for line in self.proc.stdout:
            r = [x.strip() for x in line.split(",")]
            try:
                reasons = [n for n, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6])
                           if val.lower().startswith("active")]
                self.rows.append((time.perf_counter(), float(r[0]), float(r[1]), reasons))
            except Exception:
                continue

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def close(self):
        self.stop.set()
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass

    # context-manager form: the whole `with` body is the timed region
    def __enter__(self):
        if self.thread is None:
            self.start()
        self.mark_begin()
        return self

    def __exit__(self, *a):
        self.mark_end()
        time.sleep(0.01)
        self.close()

    def summary(self):
        rows = self.rows
        inside = [r for r in rows if self.t0 is not None and self.t0 <= r[0] <= (self.t1 or r[0])]
        if not inside and rows and self.t0 is not None:       # region shorter than the sampling period: nearest samples
            mid = 0.5 * (self.t0 + (self.t1 or self.t0))
            inside = sorted(rows, key=lambda r: abs(r[0] - mid))[:3]
        if not inside:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": self.source}
        reasons = sorted({n for r in inside for n in r[3]})
        return {"sm_mhz": float(np.median([r[1] for r in inside])), "sm_max_mhz": float(max(r[2] for r in inside)),
                "reasons": reasons, "samples": len(inside), "source": self.source}


def build_scenario(worlds, seed):
    from primal_ppo_b200 import random_scenario
    return random_scenario(worlds, H, WD, N_AGENTS, density=(0.0, 0.3), queue_len=16, seed=seed,
                           unique_maps=min(256, worlds), fov=FOV, num_channel=CH)


def cpu_port_throughput(budget_s, worlds, threads, seed=7):
    """agent-steps/s of the oracle (C port of the reference's algorithm) on the host cores: the reference's 5-call step
    + getAllObservations on the same kind of worlds and actions, for ~budget_s seconds."""
    from oracle import OracleMapfGym, build_oracle
    from primal_ppo_b200 import random_actions
    build_oracle()
    sc = build_scenario(worlds, seed)
    env = OracleMapfGym(sc, seed=1234, threads=threads, use_tape=False)
    acts = random_actions(16, worlds, N_AGENTS, seed=seed)
    bufs = env.step_observe(acts[0])                            # warm-up
    t0 = time.perf_counter()
    steps = 0
    while True:
        env.step_observe(acts[steps % 16], bufs)
        steps += 1
        if time.perf_counter() - t0 >= budget_s:
            break
    dt = time.perf_counter() - t0
    return worlds * N_AGENTS * steps / dt, steps, dt


def run_reference(args):
    """--impl reference: the CPU implementation of the path alone (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # all host threads this process may use (torchrun exports OMP_NUM_THREADS=1; the oracle takes the count explicitly)
    threads = len(os.sched_getaffinity(0))
    worlds = 32 * threads
    from oracle import OracleMapfGym, build_oracle
    from primal_ppo_b200 import random_actions
    build_oracle()
    sc = build_scenario(worlds, 7)
    env = OracleMapfGym(sc, seed=1234, threads=threads, use_tape=False)
    acts = random_actions(16, worlds, N_AGENTS, seed=7)
    bufs = env.step_observe(acts[0])
    # each bench "step" is a bounded sample: `inner` env steps over `worlds` worlds
    inner = 64
    for w in range(args.warmup):
        for k in range(inner):
            env.step_observe(acts[k % 16], bufs)
    t0 = time.perf_counter()
    for s in range(args.steps):
        for k in range(inner):
            env.step_observe(acts[k % 16], bufs)
    dt = time.perf_counter() - t0
    value = worlds * N_AGENTS * inner * args.steps / dt
    sample = f"{worlds} worlds 40x40x32 agents x {inner} env steps per bench step, {threads} OpenMP threads"
    line = {"impl": "reference", "metric": "agent-steps/sec step+observe (40x40, 32 agents)", "value": value,
            "unit": "agent-steps/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8/i16 state, f32 obs", "data": "synthetic",
            "config": {"workload": WORKLOAD, "note": "CPU path on a bounded sample of the same workload"},
            "cpu_baseline": {"value": value, "unit": "agent-steps/s", "cores": threads, "kind": "port", "sample": sample,
                             "note": "C restatement of mapf_gym.py (oracle/mapf_oracle.c); the Python reference itself "
                                     "measured 1.4e3 agent-steps/s per core at this shape (SURVEY.md §6) and cannot travel "
                                     "to the GPU box"},
            "e2e": {"value": value, "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--worlds", type=int, default=65536, help="worlds per GPU")
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of CPU-baseline sampling (rank 0, N=1)")
    ap.add_argument("--no-extras", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from primal_ppo_b200 import BatchedMapfGym, gae
    from primal_ppo_b200.build import build
    build()

    rank = int(os.environ.get("RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world_size > 1:
        dist.init_process_group("nccl", device_id=dev)
    Wn, N = args.worlds, N_AGENTS
    K, Wu = args.steps, max(args.warmup, 3)

    sc = build_scenario(Wn, seed=100 + rank)
    env = BatchedMapfGym(sc, device=dev, seed=1234 + rank, use_tape=False)
    obs = torch.empty((Wn, N, CH, FOV, FOV), dtype=torch.float32, device=dev)     # the policy's input tensors
    vec = torch.empty((Wn, N, 4), dtype=torch.float32, device=dev)
    gen = torch.Generator(device=dev); gen.manual_seed(1234 + rank)
    ring = [torch.randint(0, 5, (Wn, N), generator=gen, device=dev, dtype=torch.int8) for _ in range(8)]
    env.getAllObservations(out=(obs, vec))

    def barrier():
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    hb = env.make_host_buffers(with_obs=False, action_slots=8)     # one pinned slab: results + the runner's action ring
    for k, r in enumerate(ring):
        hb["action_ring"][k].copy_(r)
    host_ring = [hb["action_ring"][k] for k in range(8)]

    def run_e2e(n_calls):
        """n_calls of the host-buffer C-ABI call, wall clock, after 3 untimed calls.  Returns (seconds, h2d, d2h bytes)."""
        for i in range(3):
            env.step_observe_host(hb, obs, vec, actions=host_ring[i % 8])
        barrier()
        t0 = time.perf_counter()
        stamps = []
        for i in range(n_calls):
            h2d_, d2h_ = env.step_observe_host(hb, obs, vec, actions=host_ring[i % 8])   # the runner's pinned host action array
            _ = float(hb["reward"][0, 0])                            # the step's result is read on the host
            stamps.append(time.perf_counter())
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        if os.environ.get("BENCH_E2E_DEBUG"):
            d = [round((b - a) * 1e3, 3) for a, b in zip([t0] + stamps[:-1], stamps)]
            print("e2e per-call ms:", d, "mean", round(dt / n_calls * 1e3, 3), file=sys.stderr)
        return dt, h2d_, d2h_

    # ---------------- device-resident throughput (value): one fused step+observe launch per step ------------------
    clk = ClockSampler(local_rank).start()          # running before the warm-up so that it has samples in the timed region
    for i in range(Wu):
        env.step_observe(ring[i % 8], obs_out=(obs, vec))
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(K)]
    barrier()
    with clk:
        t_start = torch.cuda.Event(enable_timing=True); t_end = torch.cuda.Event(enable_timing=True)
        t_start.record()
        for i in range(K):
            ev[i][0].record()
            env.step_observe(ring[i % 8], obs_out=(obs, vec))
            ev[i][1].record()
        t_end.record()
        barrier()
    total_ms = t_start.elapsed_time(t_end)
    fused_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in ev]))
    tmax = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world_size > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    total_ms_max = float(tmax.item())
    value = Wn * N * K * world_size / (total_ms_max * 1e-3)
    err_frac = float((env.state()["err"] != 0).float().mean())

    # ---------------- the same work as two launches (mapf_step, mapf_observe): per-kernel CUDA-event timings ------
    Kk = max(3, min(K, 20))
    evk = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(Kk)]
    for i in range(3):
        env.step(ring[i % 8]); env.getAllObservations(out=(obs, vec))
    torch.cuda.synchronize(dev)
    for i in range(Kk):
        evk[i][0].record()
        env.step(ring[i % 8])
        evk[i][1].record()
        env.getAllObservations(out=(obs, vec))
        evk[i][2].record()
    torch.cuda.synchronize(dev)
    step_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in evk]))
    obs_ms = float(np.mean([e[1].elapsed_time(e[2]) for e in evk]))

    # ---------------- end to end through the host-buffer C-ABI call ---------------------------------------------
    Ke = max(3, min(K, 50))
    e2e_s, h2d, d2h = run_e2e(Ke)
    te = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world_size > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = Wn * N * Ke * world_size / float(te.item())

    line_extra = {}
    if rank == 0 and not args.no_extras:
        # observations copied to host as well (what the reference's getAllObservations returns): PCIe-bound
        try:
            hbo = env.make_host_buffers(with_obs=True)
            hbo["actions"].copy_(host_ring[0]); env.step_observe_host(hbo, obs, vec)
            t0 = time.perf_counter()
            for i in range(2):
                hbo["actions"].copy_(host_ring[i % 8]); _, d2h_full = env.step_observe_host(hbo, obs, vec)
            dt = time.perf_counter() - t0
            line_extra["e2e_obs_to_host"] = {"value": Wn * N * 2 / dt, "unit": "agent-steps/s (1 GPU)",
                                             "d2h_bytes_per_step": d2h_full, "note": "obs+vec also copied to pinned host memory"}
            del hbo
        except Exception as ex:   # host memory for 4 GB pinned buffers may be unavailable
            line_extra["e2e_obs_to_host"] = {"error": str(ex)[:200]}
        # BFS maps (all W*N maps of a reset) and GAE (T=256) with their own algorithmic-byte rooflines
        peak, _ = measured_peaks()
        # context for a write-only kernel: how fast a plain device fill of 4 GiB runs on this GPU
        buf = torch.empty(1 << 30, dtype=torch.float32, device=dev)
        buf.zero_(); torch.cuda.synchronize(dev)
        fa, fb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fa.record()
        for _ in range(5):
            buf.zero_()
        fb.record(); torch.cuda.synchronize(dev)
        line_extra["write_only_fill_gbs"] = 5 * buf.numel() * 4 / (fa.elapsed_time(fb) * 1e-3) / 1e9
        del buf
        nb = min(Wn, 8192)
        ids = torch.arange(nb * N, device=dev, dtype=torch.int32)
        bfs_out = torch.empty((nb * N, H, WD), dtype=torch.int16, device=dev)
        env.bfs_maps(agent_ids=ids, out=bfs_out); torch.cuda.synchronize(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            env.bfs_maps(agent_ids=ids, out=bfs_out)
        b.record(); torch.cuda.synchronize(dev)
        bfs_ms = a.elapsed_time(b) / 3
        bfs_bytes = nb * N * (2 * H * WD + H * WD / N)
        line_extra["bfs"] = {"maps": nb * N, "ms": bfs_ms, "maps_per_s": nb * N / (bfs_ms * 1e-3),
                             "achieved_gbs": bfs_bytes / (bfs_ms * 1e-3) / 1e9, "frac": bfs_bytes / (bfs_ms * 1e-3) / 1e9 / peak}
        del bfs_out
        # BFS "all W*N maps" (reset-time) and "arrivals only" (in-loop mapf_bfs_refresh after a step)
        try:
            maps = env.bfs_maps()                                   # [W, N, H, Wd] int16
            torch.cuda.synchronize(dev)
            a.record(); env.bfs_maps(out=maps); b.record(); torch.cuda.synchronize(dev)
            all_ms = a.elapsed_time(b)
            evr = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(10)]
            arrived = torch.zeros((), dtype=torch.int64, device=dev)
            for i in range(10):
                so = env.step(ring[i % 8])
                arrived += so.goals_reached.sum()
                evr[i][0].record(); env.refresh_bfs(maps); evr[i][1].record()
            torch.cuda.synchronize(dev)
            line_extra["bfs_all_and_refresh"] = {"all_maps": Wn * N, "all_maps_ms": all_ms,
                                                 "refresh_ms_per_step": float(np.mean([x.elapsed_time(y) for x, y in evr])),
                                                 "arrivals_per_step": float(arrived.item()) / 10}
            del maps
        except Exception as ex:
            line_extra["bfs_all_and_refresh"] = {"error": str(ex)[:200]}
        T, cols = 256, 8192 * N
        r = torch.randn((T, cols), device=dev); v = torch.randn((T, cols), device=dev); lv = torch.randn((cols,), device=dev)
        gae(r, v, lv); torch.cuda.synchronize(dev)
        a.record()
        for _ in range(3):
            gae(r, v, lv)
        b.record(); torch.cuda.synchronize(dev)
        gae_ms = a.elapsed_time(b) / 3
        gae_bytes = 12.0 * T * cols
        line_extra["gae"] = {"T": T, "cols": cols, "ms": gae_ms, "elements_per_s": T * cols / (gae_ms * 1e-3),
                             "achieved_gbs": gae_bytes / (gae_ms * 1e-3) / 1e9, "frac": gae_bytes / (gae_ms * 1e-3) / 1e9 / peak}
        del r, v, lv
        # optional bf16 observation format (same values, half the bytes; not the reference layout, not the headline)
        try:
            ob16 = torch.empty((Wn, N, CH, FOV, FOV), dtype=torch.bfloat16, device=dev)
            for i in range(3):
                env.step_observe(ring[i % 8], obs_out=(ob16, vec))
            torch.cuda.synchronize(dev)
            a.record()
            for i in range(10):
                env.step_observe(ring[i % 8], obs_out=(ob16, vec))
            b.record(); torch.cuda.synchronize(dev)
            ms16 = a.elapsed_time(b) / 10
            line_extra["bf16_observations"] = {"ms_per_step": ms16, "agent_steps_per_s": Wn * N / (ms16 * 1e-3),
                                               "note": "mapf_step_observe_bf16: optional output format, outside the reference layout"}
            del ob16
        except Exception as ex:
            line_extra["bf16_observations"] = {"error": str(ex)[:200]}
        # reset-time work (not on the step path): on-device generation of 65 536 worlds, mapf_reset
        try:
            from primal_ppo_b200 import generate_scenario_device
            nw = min(Wn, 65536)
            generate_scenario_device(256, H, WD, N, kind="density", density=(0.0, 0.3), queue_len=16, seed=1, device=dev)
            torch.cuda.synchronize(dev)
            a.record()
            dsc = generate_scenario_device(nw, H, WD, N, kind="density", density=(0.0, 0.3), queue_len=16, seed=2, device=dev)
            b.record(); torch.cuda.synchronize(dev)
            gen_ms = a.elapsed_time(b)
            a.record(); env.reset(); b.record(); torch.cuda.synchronize(dev)
            line_extra["reset"] = {"worlds": nw, "generate_scenario_ms": gen_ms, "mapf_reset_ms": a.elapsed_time(b),
                                   "generator_flagged_worlds": float(((dsc.gen_err & ~4) != 0).float().mean())}
            del dsc
        except Exception as ex:
            line_extra["reset"] = {"error": str(ex)[:200]}

    if rank == 0:
        peak, peak_src = measured_peaks()
        b_step, b_obs, b_fused = algorithmic_bytes()
        obs_bytes = b_obs * Wn * N
        step_bytes = b_step * Wn * N
        fused_bytes = b_fused * Wn * N
        achieved = fused_bytes / (fused_ms * 1e-3) / 1e9
        cpu = None
        if world_size == 1 and args.cpu_budget > 0:
            threads = len(os.sched_getaffinity(0))
            cw = 32 * threads
            cv, csteps, cdt = cpu_port_throughput(args.cpu_budget, cw, threads)
            cpu = {"value": cv, "unit": "agent-steps/s", "cores": threads, "kind": "port",
                   "sample": f"{cw} worlds 40x40x32 agents, {csteps} steps in {cdt:.1f} s, {threads} OpenMP threads (oracle/mapf_oracle.c)"}
        line = {"metric": "agent-steps/sec step+observe (40x40, 32 agents)", "value": value, "unit": "agent-steps/s",
                "n_gpus": world_size, "steps": K, "warmup": Wu, "ms_per_step": total_ms_max / K, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u8/i16 state, f32 obs", "data": "synthetic",
                "config": {"workload": WORKLOAD, "worlds_per_gpu": Wn, "agents": N, "grid": [H, WD], "fov": FOV,
                           "channels": CH, "actions": "uniform random, device-resident ring of 8",
                           "call": "mapf_step_observe (one fused launch per step)",
                           "l2": "working set per step (obs 4.08 GB/GPU written) far exceeds the 126 MB L2; no flush needed",
                           "worlds_with_error_flags": err_frac},
                "clocks": clk.summary(),
                "e2e": {"value": e2e_value, "unit": "agent-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "steps": Ke, "note": "mapf_step_observe_host: actions from pinned host memory; per-agent step results "
                                             "(status, reward, cost, goals, violations, shadow goals) copied back to pinned host "
                                             "memory every step; observations and trainValid stay in HBM as the policy's / learner's "
                                             "input tensors"},
                "gpu_launches": K,
                "roofline": {"bound": "hbm", "kernel": "step_observe_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak, "traffic": ncu_traffic("step_observe_kernel"), "peak_source": peak_src,
                             "algorithmic_bytes_per_launch": fused_bytes, "ms_per_launch": fused_ms,
                             "algorithmic_bytes_per_agent_step": b_fused},
                "kernels": {"step_observe_kernel": {"ms": fused_ms, "algorithmic_bytes": fused_bytes, "achieved_gbs": achieved,
                                                    "frac": achieved / peak, "traffic": ncu_traffic("step_observe_kernel")},
                            "step_kernel": {"ms": step_ms, "algorithmic_bytes": step_bytes,
                                            "achieved_gbs": step_bytes / (step_ms * 1e-3) / 1e9,
                                            "frac": step_bytes / (step_ms * 1e-3) / 1e9 / peak,
                                            "traffic": ncu_traffic("step_kernel")},
                            "observe_kernel": {"ms": obs_ms, "algorithmic_bytes": obs_bytes,
                                               "achieved_gbs": obs_bytes / (obs_ms * 1e-3) / 1e9,
                                               "frac": obs_bytes / (obs_ms * 1e-3) / 1e9 / peak,
                                               "traffic": ncu_traffic("observe_kernel")},
                            "note": "step_kernel / observe_kernel: the same work as two launches (mapf_step, mapf_observe), "
                                    "timed in a separate loop of %d steps" % Kk},
                "cpu_baseline": cpu}
        line.update(line_extra)
        print(json.dumps(line), flush=True)
    if world_size > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

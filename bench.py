#!/usr/bin/env python
"""bench.py — headline benchmark of the fused Cosserat-rod substep path.

Workload (BASELINE.json configs[1]): SoftPendulum-v0, 4096 envs per GPU, one rod of
n_elem=50 each, FP64, one env-step = 400 PositionVerlet substeps = ONE kernel launch.
A "step" is one env-step of every env on every rank.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Prints ONE JSON line on rank 0 (contract in the task statement).  The reference arm
(`--impl reference`) times the CPU restatement of the reference path (oracle/, C,
all host threads); PyElastica itself is not installable offline.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_ELEM = 50
STEP_SKIP = 400            # int(1 / (25 * 1e-4)), soft_pendulum.py:78
FLOP_PER_ELEM_SUBSTEP = 440.0
BYTES_PER_ROD_LAUNCH = 16.0 * (18 * N_ELEM + 6)   # state read once + written once, FP64


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=25)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--envs-per-gpu", type=int, default=4096)
    p.add_argument("--math", default="fast", choices=["fast", "faithful"])
    p.add_argument("--no-cpu-baseline", action="store_true")
    return p.parse_args()


# ----------------------------------------------------------------------------- CPU legs
def cpu_rollout(n_env, n_steps, n_threads, warmup=1):
    """Oracle (C port of the reference path) on host cores; returns env-steps/s."""
    import ctypes as C
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import rod_oracle as ro
    L = ro.lib()
    envs = [ro.OracleSoftPendulum() for _ in range(n_env)]
    for i, e in enumerate(envs):
        e.reset(seed=42 + i)
    handles = (C.c_void_p * n_env)(*[e.rod._h for e in envs])
    rng = np.random.default_rng(42)
    obs = np.empty((n_env, 4), np.float32)
    rew = np.empty(n_env, np.float64)
    term = np.empty(n_env, np.int32)
    trunc = np.empty(n_env, np.int32)
    times = []
    for s in range(warmup + n_steps):
        a = rng.uniform(-22, 22, n_env).astype(np.float32)
        t0 = time.perf_counter()
        L.ro_softpendulum_step_batch(handles, n_env, a.ctypes.data, STEP_SKIP, 5.0, obs.ctypes.data,
                                     rew.ctypes.data, term.ctypes.data, trunc.ctypes.data, n_threads)
        if s >= warmup:
            times.append(time.perf_counter() - t0)
    for e in envs:
        e.rod.close()
    total = sum(times)
    return n_env * n_steps / total, total / n_steps


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import rod_oracle as ro
    cores = ro.lib().ro_max_threads()
    n_env = 16 * cores  # bounded sample of the 4096-env workload: 16 envs per thread per step
    v, sec = cpu_rollout(n_env, args.steps, cores, warmup=max(1, args.warmup))
    line = {
        "impl": "reference",
        "metric": "env-steps/s", "value": v, "unit": "env-steps/s",
        "rod_element_substeps_per_s": v * STEP_SKIP * N_ELEM,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "SoftPendulum-v0 batched 4096 envs/GPU, single rod n_elem=50, "
                               "400 substeps per env-step, FP64 (BASELINE configs[1])",
                   "envs_per_gpu": 4096, "n_elem": N_ELEM, "substeps_per_step": STEP_SKIP,
                   "envs_per_step_sample": n_env,
                   "note": "CPU arm: each step advances a bounded sample of the workload's envs by one env-step"},
        "cpu_baseline": {"value": v, "unit": "env-steps/s", "cores": cores, "kind": "port",
                         "sample": f"{n_env} envs x {args.steps} env-steps (of the 4096-env workload), "
                                   f"C restatement of PyElastica path, {cores} pthreads"},
        "e2e": {"value": v, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons during the timed region (NVML)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._halt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
        }
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


# ----------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import gym_softrobot_b200 as gsb
    from gym_softrobot_b200 import _native as nat

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU: there is no CPU fallback for the product path"
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    n_env = args.envs_per_gpu
    math = nat.MATH_FAST if args.math == "fast" else nat.MATH_FAITHFUL
    env = gsb.make_vec("SoftPendulum-v0", n_env, device=local, math=math, autoreset=False,
                       env_offset=rank * n_env)
    env.reset(seed=42)
    h = env.handle
    K, W = args.steps, args.warmup
    gen = torch.Generator(device=dev).manual_seed(42 + rank)
    actions = (torch.rand((W + K, n_env, 1), generator=gen, device=dev) * 44 - 22).float()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2
    obs, rew, term = env.obs, env.reward, env.terminated

    fp64_peak = nat.measure_fp64_peak(local)
    fp64_peak_3reg = nat.measure_fp64_peak(local, three_register_operands=True)

    # ---- device-resident leg: inputs already in HBM, one launch per step ----------
    for s in range(W):
        h.step(actions[s], STEP_SKIP, obs, rew, term)
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    launches0 = h.launch_count
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    for s in range(K):
        flush.fill_(s & 0xFF)                       # L2 flush between timed iterations (untimed)
        ev[s][0].record()
        h.step(actions[W + s], STEP_SKIP, obs, rew, term)
        ev[s][1].record()
    barrier()
    launches = h.launch_count - launches0
    clocks = sampler.stop()
    per_step_ms = [a.elapsed_time(b) for a, b in ev]
    total_s = max_over_ranks(sum(per_step_ms) * 1e-3)
    assert int(term.sum().item()) == 0, "NaN termination during the benchmark"
    env_steps = n_env * world * K
    value = env_steps / total_s
    kernel_s = sum(per_step_ms) * 1e-3 / K          # this rank's average launch duration
    elem_substeps_per_launch = n_env * N_ELEM * STEP_SKIP
    achieved_tflops = FLOP_PER_ELEM_SUBSTEP * elem_substeps_per_launch / kernel_s / 1e12
    achieved_gbs = BYTES_PER_ROD_LAUNCH * n_env / kernel_s / 1e9

    # ---- end-to-end leg: host buffers through the C-ABI (H2D + launch + D2H per step) ----
    env.reset(seed=42)
    a_host = actions.cpu().numpy()
    o_host = np.empty((n_env, 4), np.float32)
    r_host = np.empty(n_env, np.float64)
    t_host = np.empty(n_env, np.uint8)
    for s in range(W):
        h.step_host(a_host[s], STEP_SKIP, o_host, r_host, t_host)
    barrier()
    t0 = time.perf_counter()
    for s in range(K):
        h.step_host(a_host[W + s], STEP_SKIP, o_host, r_host, t_host)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e_value = env_steps / e2e_s

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"

    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "latest_traffic.json"))).get("dram_bytes_per_launch")
    except Exception:
        pass

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v1, _ = cpu_rollout(16, 64, 1)   # 1024 env-steps on one core (~10 s)
        cpu_baseline = {"value": v1, "unit": "env-steps/s", "cores": 1, "kind": "port",
                        "sample": "16 envs x 64 env-steps (1024 of the 4096x25 env-steps), C restatement "
                                  "of the PyElastica path (oracle/rod_oracle.c), 1 thread"}

    if rank == 0:
        line = {
            "metric": "env-steps/s", "value": value, "unit": "env-steps/s",
            "rod_element_substeps_per_s": value * STEP_SKIP * N_ELEM,
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": total_s / K * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "SoftPendulum-v0 batched 4096 envs/GPU, single rod n_elem=50, "
                                   "400 substeps per env-step, FP64 (BASELINE configs[1])",
                       "envs_per_gpu": n_env, "n_elem": N_ELEM, "substeps_per_step": STEP_SKIP,
                       "math": args.math, "l2": "flushed between timed iterations (256 MiB fill)",
                       "parallelism": f"env-sharded x{world}, no hot-path collective"},
            "roofline": {"bound": "fp64", "achieved": achieved_tflops, "peak": fp64_peak,
                         "unit": "TFLOP/s", "frac": achieved_tflops / fp64_peak,
                         "peak_source": "DFMA-chain microbenchmark run live (sr_measure_fp64_peak)",
                         "flop_per_element_substep": FLOP_PER_ELEM_SUBSTEP, "traffic": traffic,
                         "peak_three_register_operands": fp64_peak_3reg,
                         "frac_of_three_register_peak": achieved_tflops / fp64_peak_3reg,
                         "hbm": {"achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s",
                                 "frac": achieved_gbs / hbm_peak, "peak_source": hbm_src}},
            "cpu_baseline": cpu_baseline,
            "e2e": {"value": e2e_value, "unit": "env-steps/s",
                    "h2d_bytes_per_step": n_env * 4 * world, "d2h_bytes_per_step": n_env * (16 + 8 + 1) * world},
            "gpu_launches": launches,
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    env.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)

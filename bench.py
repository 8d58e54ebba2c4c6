#!/usr/bin/env python
"""bench.py — benchmark of the fused Cosserat-rod substep path.

Default workload (BASELINE.json configs[1], the config the metric is quoted on): SoftPendulum-v0, 4096 envs per GPU,
one rod of n_elem=50 each, FP64, one env-step = 400 PositionVerlet substeps = ONE kernel launch.  A "step" is one
env-step of every env on every rank.  `--config 3|4|5` selects the other BASELINE configs:
  3  single rod n_elem=100, gravity + analytical damping, clamped base: 65 536 envs sharded over the GPUs (strong scaling)
  4  8-arm assembly (8 x n_elem=40 + rigid head + joints + plane friction), 16 384 envs sharded over the GPUs
  5  long slender rod n_elem=512 on the frictional plane, FP64 (the FP32 / FP64 speed ratio is added to the line)

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 2|3|4|5] [--scaling weak|strong]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Prints ONE JSON line on rank 0 (contract in the task statement).  The reference arm (`--impl reference`) times the
CPU restatement of the reference path (oracle/, C, all host threads) on the same workload; PyElastica itself is not
installable offline.
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_ROD = 440.0        # SURVEY.md §8(d) / Appendix A.7, per rod-element-substep
FLOP_CONTACT = 280.0    # plane contact + anisotropic friction add-on (A.7: 250-300)

# per config: workload text, default envs (total for strong scaling / per GPU for weak), elements per env, substeps per
# step, algorithmic FLOP per element-substep, default scaling mode
CONFIGS = {
    2: dict(workload="SoftPendulum-v0 batched 4096 envs/GPU, single rod n_elem=50, 400 substeps per env-step, FP64 "
                     "(BASELINE configs[1])", envs=4096, n_elem=50, elems_per_env=50, substeps=400, flop=FLOP_ROD, scaling="weak"),
    3: dict(workload="single Cosserat rod n_elem=100, gravity + analytical damping, clamped base, 400 substeps per step, "
                     "65536 envs sharded over the GPUs, FP64 (BASELINE configs[2])", envs=65536, n_elem=100, elems_per_env=100,
            substeps=400, flop=FLOP_ROD, scaling="strong"),
    4: dict(workload="8-arm radial assembly (8 rods x n_elem=40 + rigid head + FixedJoint2Rigid joints + plane friction, "
                     "build_octopus topology, dt=3e-5), 400 substeps per step, 16384 envs sharded over the GPUs, FP64 "
                     "(BASELINE configs[3])", envs=16384, n_elem=40, elems_per_env=320, substeps=400,
            flop=FLOP_ROD + FLOP_CONTACT, scaling="strong"),
    5: dict(workload="long slender rod n_elem=512 on the frictional plane, rest-curvature actuation, 50 substeps per step, "
                     "4096 envs/GPU, FP64 (BASELINE configs[4])", envs=4096, n_elem=512, elems_per_env=512, substeps=50,
            flop=FLOP_ROD + FLOP_CONTACT, scaling="weak"),
}


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=25)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS))
    p.add_argument("--scaling", default=None, choices=["weak", "strong"])
    p.add_argument("--envs-per-gpu", type=int, default=None, help="override the env count (per GPU)")
    p.add_argument("--math", default="fast", choices=["fast", "faithful"])
    p.add_argument("--no-cpu-baseline", action="store_true")
    return p.parse_args()


def envs_per_rank(args, world):
    cfg = CONFIGS[args.config]
    if args.envs_per_gpu:
        return args.envs_per_gpu
    scaling = args.scaling or cfg["scaling"]
    return cfg["envs"] // world if scaling == "strong" else cfg["envs"]


# ----------------------------------------------------------------------------- CPU legs (oracle: the checker, timed)
def _oracle():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import rod_oracle as ro
    return ro


def cpu_systems(config, n):
    """n independent CPU-oracle systems of the workload, and a function advancing them all by one step."""
    import ctypes as C
    import numpy as np
    ro = _oracle()
    cfg = CONFIGS[config]
    K = cfg["substeps"]
    if config == 2:
        L = ro.lib()
        envs = [ro.OracleSoftPendulum() for _ in range(n)]
        for i, e in enumerate(envs):
            e.reset(seed=42 + i)
        handles = (C.c_void_p * n)(*[e.rod._h for e in envs])
        rng = np.random.default_rng(42)
        obs, rew = np.empty((n, 4), np.float32), np.empty(n, np.float64)
        term, trunc = np.empty(n, np.int32), np.empty(n, np.int32)

        def step(n_threads):
            a = rng.uniform(-22, 22, n).astype(np.float32)
            L.ro_softpendulum_step_batch(handles, n, a.ctypes.data, K, 1e30, obs.ctypes.data, rew.ctypes.data,
                                         term.ctypes.data, trunc.ctypes.data, n_threads)
        return envs, step
    if config == 3:
        ang = np.deg2rad(np.random.default_rng(42).uniform(-5, 5, n))
        systems = [ro.OracleRod(100, np.zeros(3), (np.cos(a), np.sin(a), 0.0), (-np.sin(a), np.cos(a), 0.0), 1.0, 0.025, 1000.0,
                                1e6, 5e-5, gravity=(0.0, -9.80665, 0.0), damping_constant=2e-3, bc_kind=ro.BC_ONE_END_FIXED)
                   for a in ang]
    elif config == 4:
        systems = [ro.octopus_assembly(n_arm=8, n_elem=40, time_step=3e-5) for _ in range(n)]
        rk = np.random.default_rng(2).uniform(-5, 5, (n, 8))
        for s, r in zip(systems, rk):
            for a, rod in enumerate(s.arms):
                rod.rest_kappa[0, :] = r[a]
    else:
        from gym_softrobot_b200.envs.arm_single import arm_contact_params
        contact = {**arm_contact_params(), "plane_origin": [0.0, 0.0, -0.005]}
        systems = [ro.OracleRod(512, np.zeros(3), (1.0, 0.0, 0.0), (0.0, 0.0, 1.0), 1.0, 0.005, 1000.0, 1e6, 5e-6,
                                gravity=(0.0, 0.0, -9.81), damping_constant=1e-2, contact=contact) for _ in range(n)]
        for s, k in zip(systems, np.random.default_rng(1).uniform(-3, 3, n)):
            s.rest_kappa[0, :] = k

    def step(n_threads):
        ro.substeps_batch(systems, K, n_threads)
    return systems, step


def cpu_rate(config, n_env, n_steps, n_threads, warmup=1):
    """env-steps/s (and s/step) of the C port of the reference path on host cores."""
    systems, step = cpu_systems(config, n_env)
    times = []
    for s in range(warmup + n_steps):
        t0 = time.perf_counter()
        step(n_threads)
        if s >= warmup:
            times.append(time.perf_counter() - t0)
    for e in systems:
        (e.rod if hasattr(e, "rod") else e).close()
    total = sum(times)
    return n_env * n_steps / total, total / n_steps


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ro = _oracle()
    cfg = CONFIGS[args.config]
    cores = ro.lib().ro_max_threads()
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    n_workload = envs_per_rank(args, world)
    # config 2: every step advances the whole 4096-env workload of one GPU; the heavier configs step a bounded sample
    # (a multiple of the thread count) so that the run still ends within minutes
    n_env = n_workload if args.config == 2 else {3: 16, 4: 2, 5: 2}[args.config] * cores
    v, sec = cpu_rate(args.config, n_env, args.steps, cores, warmup=max(1, min(args.warmup, 2)))
    scaling = args.scaling or cfg["scaling"]
    line = {
        "impl": "reference",
        "metric": "env-steps/s", "value": v, "unit": "env-steps/s",
        "rod_element_substeps_per_s": v * cfg["substeps"] * cfg["elems_per_env"],
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, n_workload, world),
        "cpu_baseline": {"value": v, "unit": "env-steps/s", "cores": cores, "kind": "port",
                         "sample": f"{n_env} envs x {args.steps} steps of {cfg['substeps']} substeps per timed run "
                                   f"({'the whole per-GPU workload' if n_env == n_workload else 'a bounded sample of the workload'}), "
                                   f"C restatement of the PyElastica path (oracle/rod_oracle.c), {cores} pthreads"},
        "e2e": {"value": v, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, n_env, world):
    cfg = CONFIGS[args.config]
    return {"workload": cfg["workload"], "envs_per_gpu": n_env, "n_elem": cfg["n_elem"],
            "substeps_per_step": cfg["substeps"], "math": args.math,
            "l2": "flushed between timed iterations (256 MiB fill)",
            "parallelism": f"env-sharded x{world}, no hot-path collective"}


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
            time.sleep(0.02)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


# ----------------------------------------------------------------------------- GPU workloads
class Workload:
    """One rank's share of a config: `device_step(i)` launches one step with inputs already in HBM,
    `host_step(i)` is the same step through the public host-buffer API (H2D + launch + D2H)."""

    def __init__(self, args, n_env, local, rank):
        import numpy as np
        import torch
        import gym_softrobot_b200 as gsb
        from gym_softrobot_b200 import _native as nat
        self.torch, self.np, self.n_env, self.cfg = torch, np, n_env, CONFIGS[args.config]
        self.config = args.config
        dev = torch.device(f"cuda:{local}")
        K, W = args.steps, args.warmup
        S = self.cfg["substeps"]
        math = nat.MATH_FAST if args.math == "fast" else nat.MATH_FAITHFUL
        self.env = None
        if args.config == 2:
            env = gsb.make_vec("SoftPendulum-v0", n_env, device=local, math=math, autoreset=False, env_offset=rank * n_env)
            env.reset(seed=42)
            self.env, self.handle = env, env.handle
            gen = torch.Generator(device=dev).manual_seed(42 + rank)
            self.actions = (torch.rand((W + K, n_env, 1), generator=gen, device=dev) * 44 - 22).float()
            self.a_host = self.actions.cpu().numpy()
            self.obs, self.rew, self.term = env.obs, env.reward, env.terminated
            self.host_out = (np.empty((n_env, 4), np.float32), np.empty(n_env, np.float64), np.empty(n_env, np.uint8))
            self.device_step = lambda i: self.handle.step(self.actions[i], S, self.obs, self.rew, self.term)
            self.host_step = lambda i: self.handle.step_host(self.a_host[i], S, *self.host_out)
            self.reset = lambda: env.reset(seed=42)
            self.h2d, self.d2h = n_env * 4, n_env * (16 + 8 + 1)
        elif args.config in (3, 5):
            if args.config == 3:
                h = nat.Handle(model=nat.MODEL_ROD, n_env=n_env, n_elem=100, dt=5e-5, base_length=1.0, base_radius=0.025,
                               density=1000.0, youngs_modulus=1e6, gravity=(0, -9.80665, 0), damping_constant=2e-3,
                               bc_kind=nat.BC_ONE_END_FIXED, device=local, math=math)
                ang = np.deg2rad(np.random.default_rng(42 + rank).uniform(-5, 5, n_env))
                init = np.zeros((n_env, 9))
                init[:, 3], init[:, 4], init[:, 6], init[:, 7] = np.cos(ang), np.sin(ang), -np.sin(ang), np.cos(ang)
                rk = None
            else:
                from gym_softrobot_b200.envs.arm_single import arm_contact_params
                self._make5 = lambda dtype: nat.Handle(
                    model=nat.MODEL_ROD, n_env=n_env, n_elem=512, dt=5e-6, gravity=(0, 0, -9.81), damping_constant=1e-2,
                    bc_kind=nat.BC_FREE, contact={**arm_contact_params(), "plane_origin": [0.0, 0.0, -0.005]}, base_length=1.0,
                    base_radius=0.005, density=1000.0, youngs_modulus=1e6, dtype=dtype, device=local)
                h = self._make5(nat.DTYPE_F64)
                init = np.zeros((n_env, 9)); init[:, 3] = 1.0; init[:, 8] = 1.0
                rk = np.random.default_rng(1 + rank).uniform(-3, 3, (n_env, 1)) * np.ones((1, 511))
            self.handle, self._init, self._rk = h, init, rk
            self.obs = torch.empty((n_env, 6), dtype=torch.float32, device=dev)
            self.rew = torch.empty(n_env, dtype=torch.float64, device=dev)
            self.term = torch.empty(n_env, dtype=torch.uint8, device=dev)
            self.host_out = (np.empty((n_env, 6), np.float32), np.empty(n_env, np.float64), np.empty(n_env, np.uint8))
            self.device_step = lambda i: h.step(None, S, self.obs, self.rew, self.term)
            self.host_step = lambda i: h.step_host(None, S, *self.host_out)
            self.reset = self._reset_rod
            self.reset()
            self.h2d, self.d2h = 0, n_env * (24 + 8 + 1)
        else:
            from gym_softrobot_b200.envs.octo_flat import OctoFlatVectorEnv
            # recording_fps chosen so that the public env.step advances the same 400 substeps as the device leg
            env = OctoFlatVectorEnv(n_env, n_elems=40, time_step=3e-5, recording_fps=83, device=local, autoreset=False,
                                    env_offset=rank * n_env)
            assert abs(env.step_skip - S) <= 2, env.step_skip
            self.cfg = dict(self.cfg, substeps=env.step_skip)
            S = env.step_skip
            env.reset(seed=42)
            self.env, self.handle = env, env.handle
            gen = torch.Generator(device=dev).manual_seed(42 + rank)
            self.actions = (torch.rand((W + K, n_env, 24), generator=gen, device=dev) * 10 - 5).float()
            self.a_host = self.actions.cpu().pin_memory()
            o6, self.rew, self.term = env._scratch
            W_ = env._W
            rkt = self.handle.rest_kappa_tensor().unflatten(0, (n_env, 8))

            def device_step(i):     # what env.step launches, minus the host-side reward / observation reductions
                a3 = self.actions[i].double().reshape(n_env, 8, 3)
                rkt[:, :, 0, :] = a3 @ W_.T
                self.handle.step(None, S, o6, self.rew, self.term)

            def host_step(i):       # the public vector-env API: host actions in, host observations / rewards out
                obs, rew, term, trunc, _ = env.step(self.a_host[i].to(dev, non_blocking=True))
                return obs["individual"].cpu(), obs["shared"].cpu(), rew.cpu(), term.cpu()
            self.device_step, self.host_step = device_step, host_step
            self.reset = lambda: env.reset(seed=42)
            d_ind = env.obs_shapes["individual"][0] * env.obs_shapes["individual"][1]
            self.h2d, self.d2h = n_env * 24 * 4, n_env * ((d_ind + 13) * 4 + 8 + 1)

    def _reset_rod(self):
        self.handle.reset_host(self._init)
        if self._rk is not None:
            t = self.handle.rest_kappa_tensor()
            t[:, 0, :] = self.torch.as_tensor(self._rk, device=t.device).to(t.dtype)

    def check(self):
        assert int(self.term.sum().item()) == 0, "NaN termination during the benchmark"

    def close(self):
        (self.env or self.handle).close()


def run_ours(args):
    import torch
    import torch.distributed as dist
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

    n_env = envs_per_rank(args, world)
    wl = Workload(args, n_env, local, rank)
    cfg = wl.cfg
    K, W = args.steps, args.warmup
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2
    fp64_peak = nat.measure_fp64_peak(local)
    fp64_peak_3reg = nat.measure_fp64_peak(local, three_register_operands=True)

    # ---- device-resident leg: inputs already in HBM ------------------------------------------------------------
    for s in range(W):
        wl.device_step(s)
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    launches0 = wl.handle.launch_count
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    for s in range(K):
        flush.fill_(s & 0xFF)                       # L2 flush between timed iterations (untimed)
        ev[s][0].record()
        wl.device_step(W + s)
        ev[s][1].record()
    barrier()
    launches = wl.handle.launch_count - launches0
    clocks = sampler.stop()
    per_step_ms = [a.elapsed_time(b) for a, b in ev]
    total_s = max_over_ranks(sum(per_step_ms) * 1e-3)
    wl.check()
    env_steps = n_env * world * K
    value = env_steps / total_s
    kernel_s = sum(per_step_ms) * 1e-3 / K          # this rank's average step (= launch) duration
    elem_substeps_per_launch = n_env * cfg["elems_per_env"] * cfg["substeps"]
    achieved_tflops = cfg["flop"] * elem_substeps_per_launch / kernel_s / 1e12
    # state read once + written once per launch: 18 integrated doubles per element (+6 for the tip node) per rod
    bytes_per_launch = 16.0 * (18 * cfg["elems_per_env"] + 6 * (cfg["elems_per_env"] // cfg["n_elem"])) * n_env
    achieved_gbs = bytes_per_launch / kernel_s / 1e9

    # ---- end-to-end leg: host buffers through the public API (H2D + launch + D2H per step) ------------------------
    wl.reset()
    for s in range(W):
        wl.host_step(s)
    barrier()
    t0 = time.perf_counter()
    for s in range(K):
        wl.host_step(W + s)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e_value = env_steps / e2e_s

    extra = {}
    if args.config == 5 and rank == 0:      # BASELINE config 5 asks for the FP32 vs FP64 mode comparison
        wl.handle.close()
        h32 = wl._make5(nat.DTYPE_F32)
        wl.handle = h32
        wl._reset_rod()
        for s in range(W):
            h32.step(None, cfg["substeps"], wl.obs, wl.rew, wl.term)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in range(K):
            h32.step(None, cfg["substeps"], wl.obs, wl.rew, wl.term)
        e1.record()
        torch.cuda.synchronize()
        ms32 = e0.elapsed_time(e1) / K
        extra["fp32_mode"] = {"ms_per_step": ms32, "speedup_vs_fp64": (sum(per_step_ms) / K) / ms32}

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"

    traffic = None
    if args.config == 2:
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "latest_traffic.json"))).get("dram_bytes_per_launch")
        except Exception:
            pass

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n_s, k_s = {2: (16, 64), 3: (4, 24), 4: (1, 12), 5: (1, 24)}[args.config]   # ~10-20 s on one core
        v1, _ = cpu_rate(args.config, n_s, k_s, 1)
        cpu_baseline = {"value": v1, "unit": "env-steps/s", "cores": 1, "kind": "port",
                        "sample": f"{n_s} envs x {k_s} steps of {cfg['substeps']} substeps (a bounded sample of the workload), "
                                  "C restatement of the PyElastica path (oracle/rod_oracle.c), 1 thread"}

    if rank == 0:
        srt = sorted(per_step_ms)
        line = {
            "metric": "env-steps/s", "value": value, "unit": "env-steps/s",
            "rod_element_substeps_per_s": value * cfg["substeps"] * cfg["elems_per_env"],
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": total_s / K * 1e3,
            "step_ms": {"min": srt[0], "median": statistics.median(srt), "max": srt[-1]},
            "higher_is_better": True, "scaling": args.scaling or CONFIGS[args.config]["scaling"], "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, n_env, world),
            "roofline": {"bound": "fp64", "achieved": achieved_tflops, "peak": fp64_peak,
                         "unit": "TFLOP/s", "frac": achieved_tflops / fp64_peak,
                         "peak_source": "DFMA-chain microbenchmark run live (sr_measure_fp64_peak)",
                         "flop_per_element_substep": cfg["flop"], "traffic": traffic,
                         "peak_three_register_operands": fp64_peak_3reg,
                         "frac_of_three_register_peak": achieved_tflops / fp64_peak_3reg,
                         "hbm": {"achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s",
                                 "frac": achieved_gbs / hbm_peak, "peak_source": hbm_src}},
            "cpu_baseline": cpu_baseline,
            "e2e": {"value": e2e_value, "unit": "env-steps/s",
                    "h2d_bytes_per_step": wl.h2d * world, "d2h_bytes_per_step": wl.d2h * world},
            "gpu_launches": launches,
            "clocks": clocks,
        }
        line.update(extra)
        print(json.dumps(line), flush=True)
    wl.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)

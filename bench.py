#!/usr/bin/env python
"""Headline benchmark: vehicle-steps/s of the fused dynamics + INDI hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--envs E] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Workload (``config.workload``): BASELINE.json configs[4] "scaling sweep 1M-64M vehicles, fused 8
substeps/control step" at 4 Mi vehicles PER GPU (262144 envs x 16 drones; weak scaling), each env
being configs[3]'s heterogeneous swarm: 8 quads (robobee / tello) + 8 hexa_6DOF, ground effect +
drag + downwash, K = 8 physics substeps per INDI evaluation, hover targets, noise off, quaternion
integrator.  4 Mi vehicles = 503 MB of resident state per GPU, 4x the 126 MB L2, so every step
streams its state from HBM (no L2 flush needed between steps; said in ``config.l2``).

One "step" = one control step = ONE launch of the fused kernel over the whole shard
(K substeps + one INDI evaluation per vehicle) = N * K vehicle-steps.

* ``value``      device-resident: targets already in HBM, CUDA events around exactly ``steps``
                 launches, max over ranks.
* ``e2e``        the same metric through the reference-facing C-ABI call with HOST buffers
                 (``ds_step_host``): per step, pinned host targets [N][4] -> device, the fused
                 step, per-env done flags -> pinned host memory, stream synchronised.
* ``roofline``   algorithmic bytes (241 B per vehicle per control step) / launch
                 duration against the measured HBM copy bandwidth; ``fp32`` is the second
                 roofline the path is bounded by (FP32 instruction issue), measured live.
* ``cpu_baseline`` the CPU oracle (per-vehicle FP64 Python restatement of the reference path,
                 ``kind: port``) on all host cores, on a bounded sample of the same workload.

``--impl reference`` times that CPU path alone (the reference has no runnable implementation of
this path: its DYN code is dead and its controllers need PyBullet).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "vehicle-steps/sec (dynamics+INDI ctrl)"
UNIT = "vehicle-steps/s"
K_SUBSTEPS = 8
DRONES = 16
HBM_FALLBACK_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--envs", type=int, default=262144, help="environments PER GPU (x16 drones)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = min(steps, 50)")
    ap.add_argument("--cpu-steps", type=int, default=0, help="control steps of the CPU baseline sample; 0 = sized for ~12 s")
    ap.add_argument("--no-others", action="store_true", help="skip the secondary (single-type) workloads")
    ap.add_argument("--no-sweep", action="store_true", help="skip the 1 Mi / 16 Mi / 64 Mi points of the configs[4] sweep")
    ap.add_argument("--no-parity", action="store_true", help="skip the in-run check of the timed swarm against the FP64 oracle")
    ap.add_argument("--parity-envs", type=int, default=128, help="environments of the timed swarm checked against the oracle")
    return ap.parse_args()


def workload_config(envs_per_gpu: int, n_gpus: int) -> dict:
    return {
        "workload": "hetero16 swarm (BASELINE configs[3]/[4]): %d envs x 16 drones = %d vehicles per GPU, "
                    "8 quad (robobee/tello) + 8 hexa_6DOF per env, ground effect + drag + downwash, K=8 substeps "
                    "fused per INDI control step, hover targets, quaternion integrator; layout 4x4 grid at 1.0 m pitch, "
                    "z = 2.0 + 0.25 slot (SURVEY 8d sketches 0.5 m pitch: the reference's downwash term is singular for "
                    "laterally close vehicles crossing altitudes, dronesim_b200/workloads.py)" % (envs_per_gpu, envs_per_gpu * DRONES),
        "vehicles_per_gpu": envs_per_gpu * DRONES,
        "vehicles_total": envs_per_gpu * DRONES * n_gpus,
        "substeps_per_control_step": K_SUBSTEPS,
        "sim_freq_hz": 240,
        "parallelism": "envs sharded over %d GPU(s), no per-step collective" % n_gpus,
        "l2": "inputs larger than L2 (resident state %.0f MB per GPU vs 126 MB L2); no flush" % (envs_per_gpu * DRONES * 120 / 1e6),
    }


# --------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """nvidia-smi-equivalent clock / throttle-reason samples (NVML) taken DURING the timed regions."""

    def __init__(self, index: int, period: float = 0.025):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.power = [], set(), []
        self.sm_max = None
        self._halt = threading.Event()
        self.error = None

    def _open(self):
        """NVML handle and constants, on the caller's thread BEFORE the timed region starts (the import and nvmlInit take
        longer than a short timed region)."""
        import pynvml as nv

        nv.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = self.index
        if vis:
            try:
                idx = int(vis.split(",")[self.index])
            except Exception:
                idx = self.index
        self._nv, self._h = nv, nv.nvmlDeviceGetHandleByIndex(idx)
        self.sm_max = nv.nvmlDeviceGetMaxClockInfo(self._h, nv.NVML_CLOCK_SM)
        self._names = {
            "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
            "hw_power_brake": getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80),
        }

    def _sample(self, with_power: bool):
        nv, h = self._nv, self._h
        clk = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
        r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
        self.samples.append((clk, 0))
        if with_power:  # a slower driver call: only on every eighth sample
            try:
                self.power.append(nv.nvmlDeviceGetPowerUsage(h) / 1000.0)
            except Exception:
                pass
        for k, bit in self._names.items():
            if r & bit:
                self.reasons.add(k)

    def start(self):
        try:
            self._open()
        except Exception as e:  # NVML missing: report, do not fail the bench
            self.error = repr(e)
            return
        super().start()

    def run(self):
        try:
            n = 0
            while not self._halt.is_set():
                self._sample(n % 8 == 0)
                n += 1
                time.sleep(self.period)
        except Exception as e:
            self.error = repr(e)

    def stop(self) -> dict:
        self._halt.set()
        if self.is_alive():
            self.join(timeout=2.0)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons), "error": self.error}
        clk = sorted(c for c, _ in self.samples)
        return {"sm_mhz": float(clk[len(clk) // 2]), "sm_max_mhz": float(self.sm_max) if self.sm_max else None,
                "reasons": sorted(self.reasons), "samples": len(clk),
                "power_w_max": max(self.power) if self.power else None}


def bind_to_gpu_numa_node(index: int):
    """Pin this process to the CPUs closest to its GPU (NVML's ideal affinity) so that the pinned host buffers of the
    end-to-end path are allocated on that NUMA node: with one process per GPU the H2D copies of the ranks then do not
    all cross the same socket interconnect.  Best effort - returns a description or None."""
    try:
        import pynvml as nv

        nv.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = int(vis.split(",")[index]) if vis else index
        h = nv.nvmlDeviceGetHandleByIndex(idx)
        nv.nvmlDeviceSetCpuAffinity(h)
        cpus = sorted(os.sched_getaffinity(0))
        return "cpus %d-%d (%d)" % (cpus[0], cpus[-1], len(cpus))
    except Exception as e:  # no NVML / not permitted: keep the inherited affinity
        return "unbound (%s)" % type(e).__name__


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------------------------
def cpu_arms(kinds, cores, budget_s=12.0):
    """Time the CPU implementations of the same workload on the host cores (oracle/cpu_bench.py), each on a bounded sample
    sized from a 2-step calibration to about ``budget_s`` seconds.  Returns {kind: result}."""
    from oracle.cpu_bench import time_oracle

    out = {}
    for kind in kinds:
        epw = 64 if kind == "vectorised" else 1
        try:
            cal = time_oracle(steps=2, warmup=1, workers=cores, envs_per_worker=epw, kind=kind)
            n = int(min(600, max(3, round(budget_s / max(cal["seconds"] / 2.0, 1e-3)))))
            r = time_oracle(steps=n, warmup=1, workers=cores, envs_per_worker=epw, kind=kind)
            out[kind] = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": kind, "sample": r["sample"],
                         "seconds": r["seconds"], "finite": r["finite"]}
        except Exception as e:  # a CPU arm must never take the GPU numbers down with it
            out[kind] = {"value": None, "unit": UNIT, "cores": cores, "kind": kind, "error": repr(e)[:200]}
    return out


def run_reference(args):
    """The reference arm: the reference's own algorithm for this path on the host cores.

    The reference's implementation of the path cannot be stepped as shipped (its explicit-dynamics code is dead and the
    live path needs PyBullet).  Where the reference checkout exists (this container) its own
    INDIControl / INDIControl_6DOF classes run behind a three-function pybullet shim next to the restated substep
    (``kind: reference-executed``); on the GPU box, which has no checkout, the oracle port runs (``kind: port``).
    One process per core, each stepping whole hetero16 envs."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle.cpu_bench import reference_available, time_oracle

    cores = os.cpu_count() or 1
    kind = "reference-executed" if reference_available() else "port"
    # bounded sample: one env (16 vehicles) per core per step is ~0.2 s of Python, so the driver's --steps/--warmup
    # are honoured exactly up to 1000/100 steps (~4 min); beyond that they are clamped (and the line says so)
    steps = max(1, min(args.steps, 1000))
    warm = max(1, min(args.warmup, 100))
    r = time_oracle(steps=steps, warmup=warm, workers=cores, envs_per_worker=1, kind=kind)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.envs, args.gpus),
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "reference" if kind == "reference-executed" else "port",
                         "detail": kind, "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "CPU implementation of the reference path (%s), requested steps=%d warmup=%d clamped to %d/%d to bound the run"
                % (kind, args.steps, args.warmup, steps, warm),
    }
    _emit(line)
    return 0


# --------------------------------------------------------------------------------------------
def timed_steps(core, targets, steps, barrier, max_over_ranks, torch):
    """EXACTLY ``steps`` control steps between barriers, one CUDA event per step on the launching stream.
    -> (total ms: max over ranks, per-step ms list of this rank)"""
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    barrier()
    evs[0].record()
    for i in range(steps):
        core.step(targets, 1)
        evs[i + 1].record()
    barrier()
    total = max_over_ranks(evs[0].elapsed_time(evs[steps]))
    per = [evs[i].elapsed_time(evs[i + 1]) for i in range(steps)]
    return total, per


def spread(per):
    s = sorted(per)
    n = len(s)
    return {"median_ms": s[n // 2], "p95_ms": s[min(n - 1, int(0.95 * n))], "min_ms": s[0], "max_ms": s[-1]}


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from dronesim_b200 import _lib as L
    from dronesim_b200.core import SwarmCore
    from dronesim_b200.sharding import allreduce_stats
    from dronesim_b200.vehicles import load_vehicle
    from dronesim_b200.workloads import hetero16, hetero16_bytes_per_control_step

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa_node(local_rank)  # pinned staging buffers are first-touched on the GPU's own NUMA node
    if world > 1:
        # keep stdout for the ONE JSON line: NCCL's own log lines (e.g. its version banner) go to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    n_gpus = world

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    E = args.envs
    N = E * DRONES
    models, K, flags, pos0, act0, tgt = hetero16(E, seed=0, env_offset=rank * E)
    core = SwarmCore(models, E, integrator="quat", aggregate_phy_steps=K, stats=True, device=local_rank,
                     env_offset=rank * E, **flags)
    tgt32 = np.ascontiguousarray(tgt, dtype=np.float32)
    d_tgt = torch.from_numpy(tgt32).to(dev)
    d_vel = torch.zeros((N, 4), dtype=torch.float32, device=dev)
    d_acc = torch.zeros((N, 4), dtype=torch.float32, device=dev)
    targets = core.targets_per_vehicle(d_tgt, vel=d_vel, acc=d_acc)

    # ---- parity of what is about to be timed: the first envs of this very swarm, 1 s closed loop, against the FP64 oracle
    parity = None
    if not args.no_parity:
        core.reset(pos0, action0=act0)
        pe, ps = min(args.parity_envs, E), 30  # 30 control steps x 8 substeps = 1 s at 240 Hz
        core.step(targets, ps)
        v = core.views()
        g_pos = v["pos"][: pe * DRONES].cpu().numpy().astype(np.float64)
        g_quat = v["quat"][: pe * DRONES].cpu().numpy().astype(np.float64)
        if rank == 0:
            from oracle.batch import BatchOracle

            bo = BatchOracle([load_vehicle(m) for m in models], pe, gnd=flags["ground"], drag=flags["drag"], dw=flags["downwash"],
                             aggregate_phy_steps=K)
            bo.reset(pos0[:pe])
            act = act0[:pe].copy()
            tp = tgt[: pe * DRONES, :3].reshape(pe, DRONES, 3)
            for _ in range(ps):
                bo.physics_step(act)
                act = bo.control_step(tp)
            dq = np.abs(np.sum(g_quat * bo.quat.reshape(-1, 4), axis=1)) / np.linalg.norm(g_quat, axis=1)
            parity = {"max_dpos": float(np.abs(g_pos - bo.pos.reshape(-1, 3)).max()),
                      "max_datt": float((2.0 * np.arccos(np.clip(dq, -1.0, 1.0))).max()),
                      "n_envs": int(pe), "steps": int(ps), "substeps": int(ps * K),
                      "tolerance": {"pos_m": 1e-4, "att_rad": 1e-4},
                      "oracle": "oracle/batch.py (FP64, pinned to the per-vehicle oracle by tests/test_oracle_batch.py) on the "
                                "first %d envs of the timed swarm, same initial state and targets" % pe}
            parity["ok"] = bool(parity["max_dpos"] <= 1e-4 and parity["max_datt"] <= 1e-4)
    core.reset(pos0, action0=act0)
    core.stats_reset()

    sampler = ClockSampler(local_rank)
    # ---- device-resident timing -------------------------------------------------------------
    core.step(targets, args.warmup)
    barrier()
    sampler.start()
    l0 = core.launch_count()
    ms_total, per_step = timed_steps(core, targets, args.steps, barrier, max_over_ranks, torch)
    launches = core.launch_count() - l0
    ms_per_step = ms_total / args.steps
    value = N * n_gpus * K * args.steps / (ms_total * 1e-3)

    # ---- end to end through the C ABI with HOST buffers -------------------------------------
    e2e = None
    if not args.no_e2e:
        e2e_steps = args.e2e_steps or min(args.steps, 48)
        chunk = 16  # control steps per rollout call (1 GiB of pinned per-vehicle targets at 4 Mi vehicles)
        e2e_steps = max(chunk, (e2e_steps // chunk) * chunk)
        # (1) headline: per-vehicle set-points [N][4] f32 every step, H2D / D2H overlapped with the compute
        h_roll = torch.from_numpy(tgt32).unsqueeze(0).repeat(chunk, 1, 1).contiguous().pin_memory()
        h_roll_done = torch.zeros((chunk, E), dtype=torch.uint8).pin_memory()
        core.rollout_host(h_roll, h_roll_done)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps // chunk):
            core.rollout_host(h_roll, h_roll_done)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        barrier()
        e2e_s = max_over_ranks(t1 - t0)
        e2e = {"value": N * n_gpus * K * e2e_steps / e2e_s, "unit": UNIT,
               "h2d_bytes_per_step": int(N * 16 * n_gpus), "d2h_bytes_per_step": int(E * n_gpus),
               "steps": e2e_steps, "ms_per_step": 1e3 * e2e_s / e2e_steps, "host_affinity": numa,
               "h2d_gbs_per_gpu": N * 16 / (e2e_s / e2e_steps) / 1e9,
               "call": "ds_rollout_host (%d control steps per call): per step, pinned targets [N][4] f32 -> HBM on a copy "
                       "stream in 8 MiB pieces, fused step, per-env done u8 -> pinned host; copies overlap the previous / next "
                       "step's compute; synchronised at the end of each call" % chunk}
        del h_roll
        # (2) compact: the set-points the way the reference scripts pass them - a waypoint table resident on the device
        # (here one hover row + the resident per-vehicle offsets) and ONE int32 index per vehicle and step from the host
        tab = np.zeros((1, 10))
        off = np.concatenate([tgt32[:, :3], np.zeros((N, 1), np.float32)], axis=1)
        t_tab = core.targets_table(tab, offset=torch.from_numpy(off).to(dev))
        h_wp = torch.zeros((chunk, N), dtype=torch.int32).pin_memory()
        core.rollout_host_table(t_tab, h_wp, h_roll_done)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps // chunk):
            core.rollout_host_table(t_tab, h_wp, h_roll_done)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        barrier()
        c_s = max_over_ranks(t1 - t0)
        e2e["compact_targets"] = {"value": N * n_gpus * K * e2e_steps / c_s, "unit": UNIT, "steps": e2e_steps,
                                  "ms_per_step": 1e3 * c_s / e2e_steps, "h2d_bytes_per_step": int(N * 4 * n_gpus),
                                  "d2h_bytes_per_step": int(E * n_gpus),
                                  "call": "ds_rollout_host_table: device-resident waypoint table + per-vehicle offsets, per step "
                                          "one int32 waypoint index per vehicle from pinned host memory (fly_INDI.py:230-245 "
                                          "passes TARGET_POS[wp_counters[j]]), per-env done u8 back"}
        del h_wp, h_roll_done, t_tab
        # (3) strictly synchronous per-step call (copy, step, copy, sync) for comparison
        h_tgt = torch.from_numpy(tgt32).pin_memory()
        h_done = torch.zeros((E,), dtype=torch.uint8).pin_memory()
        for _ in range(3):
            core.step_host(h_tgt, None, h_done)
        barrier()
        t0 = time.perf_counter()
        n_sync = max(8, e2e_steps // 2)
        for _ in range(n_sync):
            core.step_host(h_tgt, None, h_done)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        barrier()
        sync_s = max_over_ranks(t1 - t0)
        e2e["per_step_sync"] = {"value": N * n_gpus * K * n_sync / sync_s, "unit": UNIT, "steps": n_sync,
                                "ms_per_step": 1e3 * sync_s / n_sync, "call": "ds_step_host"}
        # (4) the gym-style variant that also returns every vehicle's 22-float state vector to the host
        h_obs = torch.empty((N, L.DS_OBS_STRIDE), dtype=torch.float32).pin_memory()
        core.step_host(h_tgt, h_obs, h_done)
        barrier()
        t0 = time.perf_counter()
        n_obs = max(3, e2e_steps // 8)
        for _ in range(n_obs):
            core.step_host(h_tgt, h_obs, h_done)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        barrier()
        obs_s = max_over_ranks(t1 - t0)
        e2e["with_full_obs"] = {"value": N * n_gpus * K * n_obs / obs_s, "unit": UNIT,
                                "d2h_bytes_per_step": int((h_obs.numel() * 4 + h_done.numel()) * n_gpus), "steps": n_obs,
                                "call": "ds_step_host with the [N][22] observation copied back"}
        del h_obs, h_tgt, h_done
    clocks = sampler.stop()

    # ---- end-of-rollout statistics: the ONLY collective of the path --------------------------
    stats = allreduce_stats(core.stats(), device=dev)
    sane = (stats["non_finite"] == 0)
    core.close()
    del d_tgt, d_vel, d_acc, targets, pos0, act0

    # ---- secondary workloads ------------------------------------------------------------------
    peak, peak_src = hbm_peak()
    others = []
    if not args.no_others:
        from dronesim_b200.workloads import algorithmic_bytes_per_control_step, single_type

        # (a) the single-type configs of BASELINE.json (configs[1], configs[2]) and the plain K=8 path, at the same vehicle count
        for name in ("quad_k8", "traj_quad", "hexa_circle"):
            models_o, K_o, flags_o, p0, a0, tab, wp0 = single_type(name, N, seed=0, env_offset=rank * N)
            c = SwarmCore(models_o, N, integrator="quat", aggregate_phy_steps=K_o, stats=True, device=local_rank, **flags_o)
            c.reset(p0, action0=a0, wp0=wp0)
            del p0, a0
            tg = c.targets_table(tab)
            steps_o = max(20, args.steps // 2)
            c.step(tg, args.warmup)
            ms_o_total, per_o = timed_steps(c, tg, steps_o, barrier, max_over_ranks, torch)
            ms_o = ms_o_total / steps_o
            st_o = allreduce_stats(c.stats(), device=dev)
            n_u = 6 if "hexa" in models_o[0] else 4
            bytes_o = algorithmic_bytes_per_control_step(n_u, per_vehicle_targets=False)
            gbs = bytes_o * N / (ms_o * 1e-3) / 1e9
            others.append({"workload": name, "vehicles_per_gpu": N, "substeps_per_control_step": K_o, "flags": flags_o,
                           "targets": "shared waypoint table + per-vehicle counter" + (
                               " (the reference trajGenerator's 3-gate table, tests/golden/traj_3gates.npz)" if name == "traj_quad" else ""),
                           "steps": steps_o, "ms_per_step": ms_o, "per_step": spread(per_o),
                           "value": N * n_gpus * K_o / (ms_o * 1e-3), "unit": UNIT,
                           "control_steps_per_s": N * n_gpus / (ms_o * 1e-3),
                           "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                                        "algorithmic_bytes_per_vehicle_control_step": bytes_o},
                           "sane": st_o["non_finite"] == 0})
            sane = sane and st_o["non_finite"] == 0
            c.close()
            del tg
        # (b) BASELINE configs[4]: the 1M-64M sweep of the headline swarm on one GPU (single-GPU runs only: 8 GB of
        # state + 5 GB of host arrays per rank at 64 Mi vehicles)
        def sweep_point(envs_s):
            m_s, K_s, f_s, p0, a0, tg_np = hetero16(envs_s, seed=0, dtype=np.float32)
            c = SwarmCore(m_s, envs_s, integrator="quat", aggregate_phy_steps=K_s, stats=True, device=local_rank, **f_s)
            try:
                c.reset(p0, action0=a0)
                del p0, a0
                tg = c.targets_per_vehicle(torch.from_numpy(tg_np).to(dev))
                del tg_np
                steps_s = 40 if envs_s <= 1048576 else 12
                c.step(tg, 5)
                ms_s_total, per_s = timed_steps(c, tg, steps_s, barrier, max_over_ranks, torch)
                ms_s = ms_s_total / steps_s
                st_s = c.stats()
                n_s = envs_s * DRONES
                return {"workload": "hetero16 sweep point (BASELINE configs[4])", "vehicles_per_gpu": n_s,
                        "substeps_per_control_step": K_s, "steps": steps_s, "ms_per_step": ms_s, "per_step": spread(per_s),
                        "value": n_s * K_s / (ms_s * 1e-3), "unit": UNIT, "resident_state_gb": n_s * 120 / 1e9,
                        "sane": st_s["non_finite"] == 0}
            finally:
                c.close()

        if world == 1 and not args.no_sweep:
            for envs_s in (65536, 1048576, 4194304):
                if envs_s == E:
                    continue
                try:
                    o = sweep_point(envs_s)
                    sane = sane and o["sane"]
                except Exception as e:  # e.g. a host too small for the 64 Mi point: the headline line must still be printed
                    o = {"workload": "hetero16 sweep point (BASELINE configs[4])", "vehicles_per_gpu": envs_s * DRONES,
                         "error": repr(e)[:200]}
                    torch.cuda.empty_cache()
                others.append(o)

    # ---- rooflines ---------------------------------------------------------------------------
    # The kernel is NOT bound by HBM: its binding limits are FP32 execution and instruction issue.
    # `roofline` reports the FP32 one on EXECUTED FLOP (ncu, per SASS opcode incl. the packed FFMA2 / FMUL2 / FADD2) against
    # the FFMA rate measured live on this GPU; the HBM and issue fractions sit beside it.
    bytes_per_launch = hetero16_bytes_per_control_step() * N
    hbm_achieved = bytes_per_launch / (ms_per_step * 1e-3) / 1e9
    hbm = {"achieved": hbm_achieved, "peak": peak, "unit": "GB/s", "frac": hbm_achieved / peak, "peak_source": peak_src,
           "algorithmic_bytes_per_launch": bytes_per_launch,
           "algorithmic_bytes_per_vehicle_control_step": hetero16_bytes_per_control_step()}
    kernel_name = "ds_step_kernel<QUAT, DW=symmetric16, NU6, WARPSYNC, FUSED, FX=ground+drag, EXT=off, mixed types, CoM offsets>"
    roofline = {"bound": "hbm", "achieved": hbm_achieved, "peak": peak, "unit": "GB/s", "frac": hbm_achieved / peak,
                "traffic": None, "peak_source": peak_src, "kernel": kernel_name, "hbm": hbm}
    fp32_peak = None
    if rank == 0:
        import ctypes as C

        pk = C.c_double(0.0)
        if L.lib().ds_debug_fp32_peak(local_rank, C.byref(pk)) == 0 and pk.value > 0:
            fp32_peak = pk.value
    prof = os.path.join(ROOT, "profiles", "roofline_inputs.json")
    if os.path.isfile(prof):
        try:
            with open(prof) as f:
                pin = json.load(f)
            fresh = pin.get("source_hash") == L.source_hash()
            roofline["inputs"] = {"file": "profiles/roofline_inputs.json", "report": pin.get("report"),
                                  "source_hash": pin.get("source_hash"), "built_source_hash": L.source_hash(), "fresh": fresh}
            if fresh and pin.get("vehicles_per_launch"):
                scale = N / float(pin["vehicles_per_launch"])
                roofline["traffic"] = pin["dram_bytes_per_launch"] * scale
                flop = pin["fp32_flop_executed_per_launch"] * scale
                wi = pin["warp_inst_per_launch"] * scale
                sm = float(clocks.get("sm_mhz") or 0.0) * 1e6
                sms = torch.cuda.get_device_properties(local_rank).multi_processor_count
                if sm > 0:
                    roofline["issue"] = {"warp_inst_per_launch": wi, "achieved_ginst_s": wi / (ms_per_step * 1e-3) / 1e9,
                                         "peak_ginst_s": sms * 4 * sm / 1e9, "frac": wi / (ms_per_step * 1e-3) / (sms * 4 * sm),
                                         "packed_fp32_share_of_fp32_inst": pin["packed_fp32_warp_inst_per_launch"] / max(1.0, pin["fp32_warp_inst_per_launch"])}
                if fp32_peak:
                    tf = flop / (ms_per_step * 1e-3) / 1e12
                    roofline.update({"bound": "fp32", "achieved": tf, "peak": fp32_peak, "unit": "TFLOP/s", "frac": tf / fp32_peak,
                                     "peak_source": "FFMA micro-benchmark run live on this GPU (ds_debug_fp32_peak; MEASURED_PEAKS.json "
                                                    "carries no FP32 CUDA-core figure); theoretical 148 SM x 128 lanes x 2 x 1.965 GHz = 74.5",
                                     "flop_executed_per_launch": flop,
                                     "flop_source": "ncu per-opcode predicated-on thread instructions (tools/roofline_inputs.py)"})
        except Exception as e:
            roofline["inputs"] = {"error": repr(e)[:200]}

    # ---- CPU baseline (rank 0, N = 1 only) -----------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle.cpu_bench import reference_available

        cores = os.cpu_count() or 1
        try:  # the GPU-side NUMA binding must not shrink the CPU arm: its workers inherit this process's affinity
            os.sched_setaffinity(0, range(cores))
            cores = len(os.sched_getaffinity(0))
        except Exception:
            pass
        kinds = (["reference-executed"] if reference_available() else []) + ["port", "vectorised"]
        arms = cpu_arms(kinds, cores)
        main_kind = "reference-executed" if "reference-executed" in arms and arms["reference-executed"].get("value") else "port"
        cpu = dict(arms[main_kind])
        cpu["kind"] = "reference" if main_kind == "reference-executed" else "port"
        cpu["detail"] = main_kind
        cpu["vectorised"] = arms.get("vectorised")
        if main_kind != "port":
            cpu["port"] = arms.get("port")

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "per_step": spread(per_step), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(E, n_gpus),
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "parity_check": parity,
            "control_steps_per_s": N * n_gpus * args.steps / (ms_total * 1e-3),
            "rollout_stats": stats, "sane": bool(sane and (parity is None or parity["ok"])), "other_workloads": others,
        }
        _emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0 if (sane and (parity is None or rank != 0 or parity["ok"])) else 3


def _emit(line: dict):
    """The ONE JSON line, written to the process's original stdout."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1

if __name__ == "__main__":
    a = parse_args()
    # Native libraries print to fd 1 behind Python's back (NCCL's version banner under torchrun): keep the real stdout for
    # the JSON line only and send everything else to stderr.
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    sys.exit(run_reference(a) if a.impl == "reference" else run_ours(a))

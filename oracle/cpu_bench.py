"""Time the CPU oracle (the reference's per-vehicle Python path, restated) on the host cores.

TEST / BENCH INFRASTRUCTURE (see ``oracle/__init__.py``): used only by ``bench.py``'s
``cpu_baseline`` leg and its ``--impl reference`` arm.  The reference itself is single-threaded
Python (its only loop over vehicles is ``for i in range(self.NUM_DRONES)``, BaseAviary.py:522);
to use "all the host threads it can use" the swarm's envs are dealt to one process per core -
envs are independent, so this is exactly how a user would scale the reference on one box.

Each worker owns whole envs of the ``hetero16`` workload and runs the example loop
(``examples/fly_INDI.py:217-245``): K substeps with the held action, then one controller call per
vehicle.  Throughput = vehicle-steps of all workers / slowest worker's wall time.
"""
from __future__ import annotations

import multiprocessing as mp
import os
import time


def _worker(rank, envs_per_worker, warmup, steps, barrier, out_q, seed, kind="port"):
    import numpy as np

    from dronesim_b200.vehicles import load_vehicle
    from dronesim_b200.workloads import hetero16

    models, K, flags, pos0, act0, tgt = hetero16(envs_per_worker, seed=seed, env_offset=rank * envs_per_worker)
    vts = [load_vehicle(m) for m in models]
    if kind == "vectorised":
        from oracle.batch import BatchOracle

        orc = BatchOracle(vts, envs_per_worker, gnd=flags["ground"], drag=flags["drag"], dw=flags["downwash"],
                          aggregate_phy_steps=K)
    else:
        from oracle.sim import OracleSwarm

        orc = OracleSwarm(vts, envs_per_worker, integrator="quat", composite=True, gnd=flags["ground"], drag=flags["drag"],
                          dw=flags["downwash"], aggregate_phy_steps=K)
        if kind == "reference-executed":
            # the reference's OWN controller classes (dronesim/control/INDIControl.py, INDIControl_6DOF.py), unmodified,
            # behind the three-function pybullet shim; only the dead explicit-dynamics substep is the restatement
            import contextlib
            import io

            from oracle import ref_shims

            with contextlib.redirect_stdout(io.StringIO()):
                orc.ctrl = [[(ref_shims.hexa_controller(m) if vt.INDI_OUTPUT_NR == 6 else ref_shims.quad_controller(m))
                             for m, vt in zip(models, vts)] for _ in range(envs_per_worker)]
    orc.reset(pos0)
    tpos = tgt[:, :3].reshape(envs_per_worker, 16, 3)
    act = act0.copy()
    for _ in range(warmup):
        orc.physics_step(act)
        act = orc.control_step(tpos)
    barrier.wait()
    t0 = time.perf_counter()
    for _ in range(steps):
        orc.physics_step(act)
        act = orc.control_step(tpos)
    dt = time.perf_counter() - t0
    ok = bool(np.isfinite(orc.pos).all())
    out_q.put((rank, dt, ok))


KINDS = {
    "port": "FP64 per-vehicle Python oracle (oracle/sim.py: the reference's loop and formulas restated)",
    "reference-executed": "the reference's own INDIControl / INDIControl_6DOF classes executed behind oracle/ref_shims.py + the "
                          "restated explicit-dynamics substep (the reference's is dead code), per vehicle",
    "vectorised": "numpy-vectorised FP64 oracle (oracle/batch.py) - NOT the reference's code: a fairer CPU bound",
}


def reference_available() -> bool:
    from oracle import ref_shims

    return ref_shims.reference_available()


def time_oracle(steps: int, warmup: int = 1, workers: int = 0, envs_per_worker: int = 1, seed: int = 0,
                kind: str = "port") -> dict:
    """Run ``steps`` control steps (after ``warmup``) of ``workers * envs_per_worker`` hetero16 envs with the CPU
    implementation ``kind`` (see ``KINDS``).

    Returns {"value": vehicle-steps/s, "cores": workers, "seconds": slowest worker, "sample": text,
    "ms_per_step": ..., "kind": kind}."""
    workers = workers or (os.cpu_count() or 1)
    ctx = mp.get_context("spawn")  # the parent may hold a CUDA context: never fork it
    barrier = ctx.Barrier(workers)
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, envs_per_worker, warmup, steps, barrier, q, seed, kind), daemon=True)
             for r in range(workers)]
    for p in procs:
        p.start()
    res = []
    while len(res) < len(procs):  # a worker that died (import error, ...) must fail the call, not hang it
        try:
            res.append(q.get(timeout=5))
        except Exception:
            if any((not p.is_alive()) and p.exitcode not in (0, None) for p in procs):
                for p in procs:
                    p.kill()
                raise RuntimeError("oracle worker died (exit codes %s)" % [p.exitcode for p in procs])
    for p in procs:
        p.join()
    slowest = max(r[1] for r in res)
    n_vehicles = workers * envs_per_worker * 16
    K = 8
    return {
        "value": n_vehicles * K * steps / slowest,
        "cores": workers,
        "seconds": slowest,
        "ms_per_step": 1e3 * slowest / max(steps, 1),
        "finite": all(r[2] for r in res),
        "kind": kind,
        "sample": "%d envs x 16 drones (hetero16: 8 quad + 8 hexa, ground+drag+downwash, K=8) x %d control steps, "
                  "one process per core; %s" % (workers * envs_per_worker, steps, KINDS[kind]),
    }

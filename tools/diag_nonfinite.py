"""Diagnose which vehicles of the hetero16 swarm leave the finite range, and when (GPU)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from dronesim_b200.core import SwarmCore
from dronesim_b200.workloads import hetero16

E = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
models, K, flags, pos0, act0, tgt = hetero16(E)
core = SwarmCore(models, E, integrator="quat", aggregate_phy_steps=K, stats=True, **flags)
core.reset(pos0, action0=act0)
t = core.targets_per_vehicle(tgt.astype(np.float32))
bad_prev = torch.zeros(E * 16, dtype=torch.bool, device="cuda")
first = {}
for it in range(30):
    core.step(t, 10)
    v = core.views()
    p = v["pos"]
    bad = ~torch.isfinite(p).all(dim=1) | (p.abs().max(dim=1).values > 1e3)
    new = bad & ~bad_prev
    idx = torch.nonzero(new).flatten().cpu().numpy()
    if len(idx):
        slots = np.bincount(idx % 16, minlength=16)
        print("after %3d control steps: %d new bad vehicles, by slot %s" % ((it + 1) * 10, len(idx), slots.tolist()))
        for i in idx[:3]:
            e = i // 16
            print("   vehicle", i, "env", e, "slot", i % 16, "pos0 env z:", np.round(pos0[e, :, 2], 3).tolist())
            print("   env pos now z:", np.round(p.view(E, 16, 3)[e, :, 2].cpu().numpy(), 3).tolist())
    bad_prev |= bad
    z = p.view(E, 16, 3)[:, :, 2]
    # closest vertical approach between any pair within lateral distance < 1.2 m, over the swarm
    zs = torch.sort(z, dim=1).values
    gap = (zs[:, 1:] - zs[:, :-1]).min().item()
    print("step %3d: min altitude gap within an env = %.4f m, z range [%.2f, %.2f], stats %s" % (
        (it + 1) * 10, gap, z.min().item(), z.max().item(), {k: v2 for k, v2 in core.stats().items() if k in ("non_finite", "wls_slow_path", "saturated_cmds")}))

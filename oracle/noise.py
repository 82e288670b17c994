"""Counter-based rotor noise: the FP64 twin of ``ds_philox4x32`` / ``ds_box_muller`` / ``ds_normals12``
(dronesim_b200/csrc/ds_device.cuh).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  The reference draws its rotor noise from the unseeded global
``np.random.normal`` (BaseAviary.py:1429-1432, 1518-1525), so its noisy runs are not reproducible; the new core
replaces the SOURCE by Philox-4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11 - the
published algorithm, round constants 0xD2511F53 / 0xCD9E8D57, key increments 0x9E3779B9 / 0xBB67AE85) keyed by the
seed with counter (vehicle, substep, draw, 0), and keeps the reference's use of the numbers.  Uniforms are the
top 24 bits + 0.5 scaled by 2^-24 (exact in FP32 and FP64), normals by Box-Muller.
"""
import math

import numpy as np

M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
MASK = 0xFFFFFFFF


def philox4x32(c, k0, k1):
    """c: 4 ints -> 4 ints after 10 rounds."""
    c0, c1, c2, c3 = [int(x) & MASK for x in c]
    k0, k1 = int(k0) & MASK, int(k1) & MASK
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        hi0, lo0, hi1, lo1 = p0 >> 32, p0 & MASK, p1 >> 32, p1 & MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & MASK, lo1, (hi0 ^ c3 ^ k1) & MASK, lo0
        k0, k1 = (k0 + W0) & MASK, (k1 + W1) & MASK
    return c0, c1, c2, c3


def box_muller(a, b):
    u1 = ((a >> 8) + 0.5) / 16777216.0
    u2 = ((b >> 8) + 0.5) / 16777216.0
    r = math.sqrt(-2.0 * math.log(u1))
    return r * math.cos(2.0 * math.pi * u2), r * math.sin(2.0 * math.pi * u2)


def normals12(vehicle, substep, seed):
    """The 12 standard normals of (global vehicle id, substep index): draws 0..2."""
    k0, k1 = seed & MASK, (seed >> 32) & MASK
    out = []
    for j in range(3):
        x = philox4x32((vehicle, substep, j, 0), k0, k1)
        out.extend(box_muller(x[0], x[1]))
        out.extend(box_muller(x[2], x[3]))
    return np.array(out)

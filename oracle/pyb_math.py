"""Restatement of the three pure-math PyBullet functions the reference calls.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).

The reference imports ``pybullet`` (Bullet3, third-party, NOT under /root/reference, version
unpinned: ``setup.py:14``) and uses exactly three math helpers on the hot path:

* ``p.getEulerFromQuaternion``  - INDIControl.py:225,301 ; INDIControl_6DOF.py:334,419,548 ;
                                  BaseAviary.py:729
* ``p.getQuaternionFromEuler``  - INDIControl.py:388 ; INDIControl_6DOF.py:538 ;
                                  BaseAviary.py:688,1817
* ``p.getMatrixFromQuaternion`` - INDIControl.py:428 ; INDIControl_6DOF.py:567 ;
                                  BaseAviary.py:1719,1786

Published algorithm restated here (Bullet3 ``examples/pybullet/pybullet.c`` /
``btQuaternion::getEulerZYX`` / ``btMatrix3x3::setRotation``): quaternions are ``xyzw``; Euler
angles are roll-pitch-yaw about fixed X, Y, Z axes (= intrinsic Z-Y'-X''); the Euler
extraction switches to a gimbal branch when ``|sin(pitch)| >= 0.99999``.  Parity unpinned (no
Bullet in the container): cross-checked against ``scipy.spatial.transform.Rotation`` in
``tests/test_oracle_math.py``.
"""
import math

import numpy as np

_GIMBAL = 0.99999


def getEulerFromQuaternion(q):
    """xyzw quaternion -> (roll, pitch, yaw).  The quaternion is NOT normalised first."""
    x, y, z, w = float(q[0]), float(q[1]), float(q[2]), float(q[3])
    sqx, sqy, sqz, squ = x * x, y * y, z * z, w * w
    sarg = -2.0 * (x * z - w * y)
    if sarg <= -_GIMBAL:
        return (0.0, -0.5 * math.pi, 2.0 * math.atan2(x, -y))
    if sarg >= _GIMBAL:
        return (0.0, 0.5 * math.pi, 2.0 * math.atan2(-x, y))
    return (
        math.atan2(2.0 * (y * z + w * x), squ - sqx - sqy + sqz),
        math.asin(sarg),
        math.atan2(2.0 * (x * y + w * z), squ + sqx - sqy - sqz),
    )


def getQuaternionFromEuler(rpy):
    """(roll, pitch, yaw) -> normalised xyzw quaternion."""
    phi, the, psi = 0.5 * float(rpy[0]), 0.5 * float(rpy[1]), 0.5 * float(rpy[2])
    sph, cph = math.sin(phi), math.cos(phi)
    sth, cth = math.sin(the), math.cos(the)
    sps, cps = math.sin(psi), math.cos(psi)
    x = sph * cth * cps - cph * sth * sps
    y = cph * sth * cps + sph * cth * sps
    z = cph * cth * sps - sph * sth * cps
    w = cph * cth * cps + sph * sth * sps
    n = math.sqrt(x * x + y * y + z * z + w * w)
    return (x / n, y / n, z / n, w / n)


def getMatrixFromQuaternion(q):
    """xyzw quaternion -> row-major 9-tuple of the rotation matrix (body -> world)."""
    x, y, z, w = float(q[0]), float(q[1]), float(q[2]), float(q[3])
    d = x * x + y * y + z * z + w * w
    s = 2.0 / d
    xs, ys, zs = x * s, y * s, z * s
    wx, wy, wz = w * xs, w * ys, w * zs
    xx, xy, xz = x * xs, x * ys, x * zs
    yy, yz, zz = y * ys, y * zs, z * zs
    return (
        1.0 - (yy + zz), xy - wz, xz + wy,
        xy + wz, 1.0 - (xx + zz), yz - wx,
        xz - wy, yz + wx, 1.0 - (xx + yy),
    )


def rotmat(q):
    """Convenience: 3x3 ndarray of ``getMatrixFromQuaternion``."""
    return np.array(getMatrixFromQuaternion(q), dtype=np.float64).reshape(3, 3)

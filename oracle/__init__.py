"""CPU oracle for the dronesim dynamics + INDI hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the shipped product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and there only as the checker / the CPU arm being
timed.  The product path (``dronesim_b200``) never imports this package and fails loudly
when its CUDA library is missing.

Contents (every function cites the reference file:line it restates; paths are relative to
the reference checkout, enac-drones/dronesim):

* ``pyb_math``   - the three pure-math PyBullet functions the reference calls
                   (third-party Bullet3, un-vendored and unpinned in ``setup.py:14``).
* ``control``    - FP64 restatement of ``dronesim/control/INDIControl.py``,
                   ``INDIControl_6DOF.py``, ``wls_alloc.py`` and ``utils/math.py``.
* ``dynamics``   - FP64 restatement of ``BaseAviary._dynamics/_groundEffect/_drag/_downwash``
                   (dead code in the reference) with the documented repairs.
* ``sim``        - the example-script loop (physics K substeps, then control) over envs x drones.
* ``ref_shims``  - ``sys.modules`` shims that let the UNMODIFIED reference controller classes
                   import in a container without pybullet/gym.  Used only where
                   ``/root/reference`` exists (golden-vector generation, restatement checks).

Parity pinning: the control half is pinned against the reference's own code executed behind
``ref_shims`` (fixtures in ``tests/golden/``, generator ``tests/golden/make_golden.py``) and
against the one known-answer test the reference ships (``wls_alloc.py:381-408``); so are the
``VelocityAviary`` / ``RPYTAviary`` ``_preprocessAction`` restatements and the ``Logger`` data model.
The dynamics half restates formulas the reference cannot step (dead code + PyBullet absent), but
whose bodies run unbound on a stand-in ``self``: pinned that way for the 4-rotor airframes
(``dyn_*.npz``); the rotor-geometry generalisation for the hexa and the quaternion integrator are
restated only - **parity unpinned** for those two and for the Bullet math trio
(self-consistency and scipy cross-checks only).
"""

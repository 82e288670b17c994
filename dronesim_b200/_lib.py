"""ctypes binding of ``libdronesim_b200.so`` (the C ABI declared in ``include/dronesim_b200.h``).

There is no CPU or PyTorch fallback: if the shared library is missing, importing this module
raises, and every entry point returns an error status when no CUDA device is present.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
# DRONESIM_B200_LIB selects an alternative build of the same ABI (kernel experiments: tools/build_variant.py)
LIB_PATH = os.environ.get("DRONESIM_B200_LIB") or os.path.join(_HERE, "libdronesim_b200.so")
CSRC = os.path.join(_HERE, "csrc")
INCLUDE = os.path.join(_HERE, "..", "include")

DS_MAX_ROTORS = 6
DS_MAX_TYPES = 8
DS_MAX_DRONES_PER_ENV = 32
DS_OBS_STRIDE = 22
DS_NUM_STATS = 16

DS_OK, DS_ERR_INVALID, DS_ERR_CUDA, DS_ERR_STATE, DS_ERR_UNSUPPORTED = 0, 1, 2, 3, 4
DS_INTEG_QUAT, DS_INTEG_RPY = 0, 1
DS_FLAG_GROUND, DS_FLAG_DRAG, DS_FLAG_DOWNWASH, DS_FLAG_STATS, DS_FLAG_DW_ORDERED_PAIRS, DS_FLAG_TYPES_IN_SMEM, DS_FLAG_GROUND_PLANE = 1, 2, 4, 8, 16, 32, 64
DS_FLAG_DEBUG_REDZONES = 128
DS_LAW_QUAD, DS_LAW_6DOF = 0, 1
DS_DONE_GOAL, DS_DONE_FLOOR, DS_DONE_TIME = 1, 2, 4
DS_ORDER_PHYSICS_THEN_CONTROL, DS_ORDER_CONTROL_THEN_PHYSICS = 0, 1

_R = DS_MAX_ROTORS


class ds_config(C.Structure):
    _fields_ = [
        ("n_envs", C.c_int32), ("drones_per_env", C.c_int32), ("substeps", C.c_int32), ("integrator", C.c_int32),
        ("flags", C.c_uint32), ("device", C.c_int32), ("sim_freq", C.c_float), ("gravity", C.c_float),
        ("neighbourhood_radius", C.c_float), ("done_goal_enable", C.c_int32), ("goal", C.c_float * 3),
        ("goal_radius", C.c_float), ("done_floor_enable", C.c_int32), ("z_min", C.c_float),
        ("max_steps", C.c_int32), ("env_offset", C.c_int32),
        ("motor_tau", C.c_float), ("acc_filter_hz", C.c_float), ("reward_mode", C.c_int32),
        ("noise_force_sigma", C.c_float), ("noise_torque_sigma", C.c_float), ("noise_seed", C.c_uint64),
        ("ground_plane_z", C.c_float), ("reserved0", C.c_int32),
    ]


class ds_type_params(C.Structure):
    _fields_ = [
        ("n_u", C.c_int32), ("n_v", C.c_int32), ("law", C.c_int32), ("rotor_model", C.c_int32),
        ("mass", C.c_double), ("J", C.c_double * 9), ("r_com", C.c_double * 3), ("kf", C.c_double), ("km", C.c_double),
        ("rotor_pos", (C.c_double * 3) * _R), ("rotor_axis", (C.c_double * 3) * _R),
        ("torque_axis", (C.c_double * 3) * _R), ("rotor_spin", C.c_double * _R),
        ("pwm2rpm_scale", C.c_double * _R), ("pwm2rpm_const", C.c_double * _R),
        ("min_pwm", C.c_double * _R), ("max_pwm", C.c_double * _R),
        ("gnd_eff_coeff", C.c_double), ("prop_radius", C.c_double), ("gnd_eff_h_clip", C.c_double),
        ("drag_coeff", C.c_double * 3), ("dw_coeff", C.c_double * 3),
        ("kp_pos", C.c_double), ("kd_pos", C.c_double), ("att_gain", C.c_double * 3), ("rate_gain", C.c_double * 3),
        ("G1", (C.c_double * _R) * _R), ("alloc", (C.c_double * _R) * _R),
        ("wls_wv", C.c_double * _R), ("wls_gamma", C.c_double), ("init_cmd", C.c_double), ("init_thrust", C.c_double),
        ("max_speed_kmh", C.c_double), ("adv_coeff", C.c_double * 14), ("adv_radius", C.c_double),
    ]


class ds_targets(C.Structure):
    _fields_ = [
        ("mode", C.c_int32), ("num_wp", C.c_int32), ("advance_wp", C.c_int32), ("reserved", C.c_int32),
        ("pos_yaw", C.c_void_p), ("vel", C.c_void_p), ("acc", C.c_void_p), ("table", C.c_void_p), ("offset", C.c_void_p),
        ("wp", C.c_void_p),
    ]


class ds_state_views(C.Structure):
    _fields_ = [
        ("n", C.c_int64), ("n_pad", C.c_int64),
        ("pos_thrust", C.c_void_p), ("quat", C.c_void_p), ("vel_rpm", C.c_void_p), ("omega_wp", C.c_void_p),
        ("lastvel_done", C.c_void_p), ("lastrates_err", C.c_void_p), ("cmd0123", C.c_void_p), ("cmd45", C.c_void_p),
        ("slot_type", C.c_void_p), ("step_counter", C.c_int64),
        ("rpm0123", C.c_void_p), ("rpm45", C.c_void_p), ("ang_acc_filt", C.c_void_p),
    ]


# every symbol include/dronesim_b200.h declares: name -> (restype, argtypes)
_H = C.c_void_p
SYMBOLS = {
    "ds_create": (C.c_int, [C.POINTER(ds_config), C.POINTER(_H)]),
    "ds_destroy": (None, [_H]),
    "ds_set_types": (C.c_int, [_H, C.POINTER(ds_type_params), C.c_int32, C.POINTER(C.c_uint8)]),
    "ds_reset": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ds_reset_envs": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ds_set_step_counter": (C.c_int, [_H, C.c_int64]),
    "ds_step": (C.c_int, [_H, C.POINTER(ds_targets), C.c_int32, C.c_int32, C.c_void_p]),
    "ds_physics_step": (C.c_int, [_H, C.c_void_p, C.c_void_p]),
    "ds_control_step": (C.c_int, [_H, C.POINTER(ds_targets), C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ds_control_from_state": (C.c_int, [_H, C.c_void_p, C.POINTER(ds_targets), C.c_float, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p]),
    "ds_rate_control_step": (C.c_int, [_H, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p]),
    "ds_get_obs": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ds_set_env_outputs": (C.c_int, [_H, C.c_void_p, C.c_void_p]),
    "ds_views": (C.c_int, [_H, C.POINTER(ds_state_views)]),
    "ds_stats": (C.c_int, [_H, C.POINTER(C.c_double), C.c_int32, C.c_void_p]),
    "ds_stats_reset": (C.c_int, [_H, C.c_void_p]),
    "ds_log_attach": (C.c_int, [_H, C.c_void_p, C.c_int32, C.c_int32]),
    "ds_log_read": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.POINTER(C.c_int32), C.c_void_p]),
    "ds_step_host": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ds_rollout_host": (C.c_int, [_H, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "ds_rollout_host_table": (C.c_int, [_H, C.POINTER(ds_targets), C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "ds_debug_wls": (C.c_int, [_H, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                               C.c_int32, C.c_void_p]),
    "ds_debug_check_redzones": (C.c_int, [_H, C.POINTER(C.c_int64)]),
    "ds_debug_fp32_peak": (C.c_int, [C.c_int32, C.POINTER(C.c_double)]),
    "ds_strerror": (C.c_char_p, [C.c_int]),
    "ds_last_cuda_error": (C.c_int, [_H]),
    "ds_abi_version": (C.c_int, []),
    "ds_launch_count": (C.c_int64, [_H]),
}

# -prec-div / -prec-sqrt off, -ftz on: divisions and square roots are 2-ulp MUFU sequences without the denormal
# fix-up code around them (the kernel's own MUFU forms are .ftz already); measured -1.3 % / -0.7 % on the headline
# kernel, -2 % on the K = 2 workloads, all parity tests unchanged
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-prec-div=false",
              "-prec-sqrt=false", "-ftz=true", "-Xcompiler", "-fPIC"]
# translation units: the C ABI + small kernels, and the step kernel's instantiations split by (integrator, mode, rotors)
_UNITS = [("ds_api.cu", "api", [])] + [
    ("ds_step_inst.cu", "step_%s%d_%d" % ("qr"[i], m, n), ["-DDS_INST_INTEG=%d" % i, "-DDS_INST_MODE=%d" % m, "-DDS_INST_NU6=%d" % n])
    for i in (0, 1) for m in (0, 1, 2) for n in (0, 1)]


def build(verbose: bool = False, force: bool = False, extra_flags=(), out_path: str = None, only_units=None) -> str:
    """Compile the CUDA core for sm_100a, in-tree: ``nvcc -c`` of every translation unit in parallel
    (objects under ``csrc/build/``), then one ``nvcc -shared`` link."""
    from concurrent.futures import ThreadPoolExecutor

    out_path = out_path or LIB_PATH
    srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh"))]
    srcs.append(os.path.join(INCLUDE, "dronesim_b200.h"))
    if not force and os.path.isfile(out_path) and all(os.path.getmtime(out_path) >= os.path.getmtime(s) for s in srcs):
        return out_path
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    tag = os.path.basename(out_path).replace("libdronesim_b200", "").replace(".so", "").strip(".") or "main"
    objdir = os.path.join(CSRC, "build", tag)
    os.makedirs(objdir, exist_ok=True)

    def compile_unit(unit):
        src, name, defs = unit
        obj = os.path.join(objdir, name + ".o")
        cmd = ([nvcc] + NVCC_FLAGS + list(extra_flags) + defs + (["-Xptxas", "-v"] if verbose else [])
               + ["-c", "-o", obj, os.path.join(CSRC, src)])
        return obj, subprocess.run(cmd, capture_output=True, text=True)

    units = [u for u in _UNITS if only_units is None or u[1] in only_units]
    with ThreadPoolExecutor(max_workers=min(len(units), os.cpu_count() or 4)) as ex:
        results = list(ex.map(compile_unit, units))
    for obj, res in results:
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError("nvcc failed building %s" % obj)
        if verbose:
            sys.stderr.write(res.stderr)
    objs = [os.path.join(objdir, u[1] + ".o") for u in _UNITS]
    res = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out_path] + objs,
                         capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed linking libdronesim_b200.so")
    return out_path


def source_hash() -> str:
    """sha256 over everything the step kernel is a function of: the device-side sources it is compiled from and the nvcc
    flags (not the host API in ds_api.cu / the small kernels in ds_aux_kernels.cuh).  ``profiles/roofline_inputs.json`` is
    stamped with it; ``bench.py`` refuses ncu-derived constants that were measured on other kernel sources."""
    import hashlib

    h = hashlib.sha256()
    for f in ("ds_lanes.cuh", "ds_device.cuh", "ds_wls.cuh", "ds_control.cuh", "ds_physics.cuh", "ds_kernels.cuh",
              "ds_step_inst.cuh", "ds_step_inst.cu"):
        h.update(f.encode())
        h.update(open(os.path.join(CSRC, f), "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()[:16]


_lib = None


def lib() -> C.CDLL:
    """The loaded library.  Raises if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(
                "dronesim_b200: %s is missing - build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  There is no CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)  # AttributeError if the ABI lost a symbol
            fn.restype = res
            fn.argtypes = args
        if L.ds_abi_version() != 2:
            raise RuntimeError("dronesim_b200: ABI version mismatch")
        _lib = L
    return _lib


class DsError(RuntimeError):
    pass


def check(status: int, handle=None):
    if status != DS_OK:
        msg = lib().ds_strerror(status).decode()
        if status == DS_ERR_CUDA and handle is not None:
            msg += " [cudaError %d]" % lib().ds_last_cuda_error(handle)
        raise DsError("dronesim_b200: %s" % msg)

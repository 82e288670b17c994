"""CPU: the C-ABI library loads and exports every symbol include/dronesim_b200.h declares; argument
validation and the no-CPU-fallback contract (no compute calls are made - there is no GPU here)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from dronesim_b200 import _lib

    _lib.build()
    return _lib.lib()


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "dronesim_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ds_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree(lib):
    from dronesim_b200 import _lib

    declared = _declared_symbols()
    assert len(declared) >= 18
    assert sorted(_lib.SYMBOLS) == declared
    for name in declared:
        assert getattr(lib, name) is not None


def test_exports_via_nm():
    import subprocess

    from dronesim_b200 import _lib

    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(l.split()[-1] for l in out.splitlines() if " T " in l)
    for name in _declared_symbols():
        assert name in exported, name


def test_struct_layouts_match_header(lib):
    """sizeof of the ctypes mirrors == sizeof in the header, compiled with gcc."""
    import subprocess
    import tempfile

    from dronesim_b200 import _lib

    prog = ('#include "dronesim_b200.h"\n#include <stdio.h>\n#include <stddef.h>\nint main(){printf("%zu %zu %zu %zu %zu %zu\\n",'
            "sizeof(ds_config),sizeof(ds_type_params),sizeof(ds_targets),sizeof(ds_state_views),"
            "offsetof(ds_type_params,alloc),offsetof(ds_config,env_offset));return 0;}\n")
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(prog)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "t.c"), "-o", os.path.join(d, "t")])
        sizes = [int(x) for x in subprocess.check_output([os.path.join(d, "t")]).split()]
    assert sizes == [C.sizeof(_lib.ds_config), C.sizeof(_lib.ds_type_params), C.sizeof(_lib.ds_targets),
                     C.sizeof(_lib.ds_state_views), _lib.ds_type_params.alloc.offset, _lib.ds_config.env_offset.offset]


def test_misc_entry_points(lib):
    assert lib.ds_abi_version() == 2
    assert lib.ds_strerror(0) == b"ok"
    assert b"no CPU fallback" in lib.ds_strerror(2)
    assert lib.ds_launch_count(None) == 0
    lib.ds_destroy(None)  # no-op


def test_create_validates_and_refuses_without_gpu(lib):
    import torch

    from dronesim_b200 import _lib

    h = C.c_void_p()
    cfg = _lib.ds_config()
    assert lib.ds_create(C.byref(cfg), C.byref(h)) == _lib.DS_ERR_INVALID  # zero envs
    cfg.n_envs, cfg.drones_per_env, cfg.substeps, cfg.sim_freq = 4, 64, 1, 240.0
    assert lib.ds_create(C.byref(cfg), C.byref(h)) == _lib.DS_ERR_UNSUPPORTED  # > DS_MAX_DRONES_PER_ENV
    cfg.drones_per_env = 2
    cfg.integrator = 7
    assert lib.ds_create(C.byref(cfg), C.byref(h)) == _lib.DS_ERR_INVALID
    cfg.integrator = 0
    if not torch.cuda.is_available():
        # the product path fails loudly: there is no CPU implementation behind the ABI
        assert lib.ds_create(C.byref(cfg), C.byref(h)) == _lib.DS_ERR_CUDA
        assert not h.value
        from dronesim_b200.core import SwarmCore

        with pytest.raises(_lib.DsError):
            SwarmCore(["robobee"], 1)


def test_product_does_not_import_oracle():
    """The shipped package must never route through oracle/ (or any CPU fallback)."""
    pkg = os.path.join(ROOT, "dronesim_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


def test_graft_entry_build_runs_on_a_cpu_only_box():
    """The driver's build check: __graft_entry__.build() compiles (or finds up to date) the in-tree library, loads it and
    checks the ABI version - without a GPU."""
    import importlib

    g = importlib.import_module("__graft_entry__")
    g.build()

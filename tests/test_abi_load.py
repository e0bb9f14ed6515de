"""The C-ABI library loads on a machine without a GPU and exports every symbol include/quadsim.h declares; the ctypes
mirrors have the library's struct sizes; argument validation works without launching anything."""
import ctypes as C
import os
import re

from gym_pybullet_drones_b200 import _native as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "quadsim.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(qs_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    N.build()                                   # no-op when the in-tree .so is newer than its sources
    lib = N.lib()
    syms = declared_symbols()
    assert len(syms) >= 14 and set(N.EXPORTS) == set(syms)
    for s in syms:
        assert getattr(lib, s) is not None, s
    assert lib.qs_abi_version() == N.ABI_VERSION


def test_struct_mirrors_match_the_library():
    lib = N.lib()
    assert lib.qs_sizeof_params() == C.sizeof(N.QsParams) and lib.qs_sizeof_state() == C.sizeof(N.QsState)
    assert lib.qs_sizeof_step_io() == C.sizeof(N.QsStepIO) and lib.qs_sizeof_rollout_io() == C.sizeof(N.QsRolloutIO)
    assert lib.qs_sizeof_host_io() == C.sizeof(N.QsHostIO)


def test_argument_errors_without_a_gpu():
    """Every entry point validates its arguments before touching CUDA: negative codes + a message, no launch."""
    lib = N.lib()
    P, st, io = N.QsParams(), N.QsState(), N.QsStepIO()
    assert lib.qs_step(None, C.byref(st), C.byref(io), 0, 1, 4, 1, 8, 0, 0, None) == -1
    assert lib.qs_step(C.byref(P), C.byref(st), C.byref(io), 0, 1, 4, 1, 8, 0, 0, None) == -1          # NULL planes
    assert b"planes" in lib.qs_last_error()
    buf = (C.c_float * 64)()
    st.planes = C.addressof(buf) + 4
    st.step_counter = C.addressof(buf)
    assert lib.qs_step(C.byref(P), C.byref(st), C.byref(io), 0, 1, 4, 1, 8, 0, 0, None) == -2          # misaligned planes
    assert lib.qs_rollout_max_ticks(N.ACT_RPM, 15, 2) == 255 and lib.qs_rollout_max_ticks(N.ACT_RPM, 120, 2) == 0
    assert lib.qs_rollout_max_ticks(N.ACT_ONE_D_RPM, 15, 1) > 1000 and lib.qs_rollout_max_ticks(99, 15, 1) == 0
    assert lib.qs_pid_control(C.byref(P), None, 0.01, None, 3, None, 4, None, 3, None, None, None, None, 8, None, None, None, None) == -1
    assert lib.qs_downwash(C.byref(P), C.byref(st), 0, 4, None, None) == -1
    assert lib.qs_reset(C.byref(P), C.byref(st), None, 1, 1, 0, None, 12, 0, None) == -2
    try:
        N.check(-3, "x")
        raise AssertionError("check() must raise")
    except ValueError:
        pass


def test_product_package_does_not_import_the_oracle():
    """The oracle is test infrastructure: nothing under the product package may reference it."""
    pkg = os.path.join(ROOT, "gym_pybullet_drones_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(base, f)).read()
                assert "oracle" not in src.replace("oracle/", "").replace("the oracle", "") or "import oracle" not in src and "from oracle" not in src, f
                assert "import oracle" not in src and "from oracle" not in src, f

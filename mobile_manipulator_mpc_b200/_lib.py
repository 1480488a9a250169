"""Loader of the CUDA shared library (libmmpc_b200.so) behind include/mmpc.h.

The library is built in-tree by ``__graft_entry__.build()`` / ``build_library()``.  There is no
CPU fallback: if the library is missing or no B200-class device is present, every compute entry
point raises.
"""
import ctypes as C
import os
import subprocess

from . import _abi

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MMPC_LIB") or os.path.join(_PKG, "libmmpc_b200.so")  # MMPC_LIB: A/B builds of the same sources
SOURCES = [os.path.join(_PKG, "csrc", f) for f in ("mmpc_api.cu", "mmpc_resident.cu", "mmpc_resident_pose.cu", "mmpc_staged.cuh", "mmpc_team.cuh", "mmpc_parts.cuh", "mmpc_episode.cuh", "mmpc_ipm.cuh", "mmpc_model.cuh", "mmpc_warp.cuh")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]

_lib = None


class MmpcError(RuntimeError):
    pass


def build_library(force=False, verbose=False):
    """nvcc-compile csrc/ into libmmpc_b200.so for sm_100a (cross-compiles without a GPU)."""
    src_time = max(os.path.getmtime(s) for s in SOURCES + [os.path.join(_PKG, "..", "include", "mmpc.h")])
    if not force and os.path.exists(LIB_PATH) and os.path.getmtime(LIB_PATH) >= src_time:
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("MMPC_NVCC_EXTRA", "").split()   # A/B builds: -D switches
    flags = [f for f in NVCC_FLAGS if f != "-shared"] + (["-Xptxas", "-v"] if verbose else []) + extra
    # three translation units (staged, resident, resident with the pose-reference cost), compiled side by side, then linked
    objs = [LIB_PATH + "." + os.path.basename(s)[:-3] + ".o" for s in SOURCES[:3]]
    procs = [subprocess.Popen([nvcc] + flags + ["-c", "-o", o, s], cwd=os.path.join(_PKG, "csrc")) for o, s in zip(objs, SOURCES[:3])]
    for pr in procs:
        if pr.wait() != 0:
            raise subprocess.CalledProcessError(pr.returncode, pr.args)
    subprocess.check_call([nvcc] + NVCC_FLAGS[:2] + ["-shared", "-o", LIB_PATH] + objs, cwd=os.path.join(_PKG, "csrc"))
    return LIB_PATH


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MmpcError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(nvcc, sm_100a).  There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, dp = C.c_void_p, C.c_int32, C.c_int64, C.c_void_p
    L.mmpc_version.restype = C.c_int
    L.mmpc_error_string.restype = C.c_char_p
    L.mmpc_error_string.argtypes = [C.c_int]
    L.mmpc_default_config.argtypes = [C.POINTER(_abi.MmpcConfig)]
    L.mmpc_default_config.restype = None
    L.mmpc_struct_sizes.argtypes = [C.POINTER(i32)] * 3
    L.mmpc_create.argtypes = [C.POINTER(_abi.MmpcConfig), i32, i32, C.POINTER(vp)]
    L.mmpc_destroy.argtypes = [vp]
    L.mmpc_set_weights.argtypes = [vp, dp, dp, dp, dp, C.c_double]
    L.mmpc_set_kernel.argtypes = [vp, i32]
    L.mmpc_set_profile.argtypes = [vp, i32]
    L.mmpc_window.argtypes = [vp, i32, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp]
    L.mmpc_workspace_bytes.argtypes = [vp, C.POINTER(C.c_int64)]
    L.mmpc_phase_times.argtypes = [vp, dp, C.POINTER(C.c_int64), C.POINTER(i32)]
    L.mmpc_solve.argtypes = [vp, i32, C.POINTER(_abi.MmpcBatchIn), C.POINTER(_abi.MmpcBatchOut), vp]
    L.mmpc_solve_host.argtypes = [vp, i32, C.POINTER(_abi.MmpcBatchIn), C.POINTER(_abi.MmpcBatchOut)]
    L.mmpc_eval_model.argtypes = [vp, i32, dp, dp, dp, dp, dp, dp, dp, vp]
    L.mmpc_shift.argtypes = [vp, i32, dp, dp, vp]
    L.mmpc_plant_step.argtypes = [vp, i32, dp, dp, dp, vp]
    L.mmpc_ik.argtypes = [vp, i32, dp, dp, dp, vp, vp]
    L.mmpc_episode_update.argtypes = [vp, i32, i32, i32, C.POINTER(_abi.MmpcEpisodeIO), vp]
    L.mmpc_launch_count.argtypes = [vp]
    L.mmpc_launch_count.restype = i64
    L.mmpc_last_solver.argtypes = [vp]
    L.mmpc_occupancy.argtypes = [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
    L.mmpc_bench_fp64.argtypes = [i32, C.POINTER(C.c_double)]
    a, b, c = i32(), i32(), i32()
    L.mmpc_struct_sizes(C.byref(a), C.byref(b), C.byref(c))
    if (a.value, b.value, c.value) != (C.sizeof(_abi.MmpcConfig), C.sizeof(_abi.MmpcBatchIn), C.sizeof(_abi.MmpcBatchOut)):
        raise MmpcError("ABI mismatch between _abi.py and libmmpc_b200.so")
    _lib = L
    return L


def check(rc):
    if rc != _abi.OK:
        raise MmpcError(f"mmpc error {rc}: {lib().mmpc_error_string(rc).decode()}")


EXPORTS = ("mmpc_version", "mmpc_error_string", "mmpc_default_config", "mmpc_create", "mmpc_destroy",
           "mmpc_set_weights", "mmpc_set_kernel", "mmpc_set_profile", "mmpc_phase_times", "mmpc_workspace_bytes", "mmpc_solve", "mmpc_solve_host", "mmpc_eval_model", "mmpc_shift",
           "mmpc_plant_step", "mmpc_window", "mmpc_ik", "mmpc_episode_update", "mmpc_launch_count", "mmpc_last_solver", "mmpc_struct_sizes", "mmpc_occupancy", "mmpc_bench_fp64")

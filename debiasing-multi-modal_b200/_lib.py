"""ctypes binding of libdbmm.so (the C ABI declared in include/dbmm.h).

There is no CPU fallback: importing the package works anywhere (so CLI parsing, data conversion and
the CPU tests of host logic run without a GPU), but every compute entry point raises if the shared
library is missing or no CUDA device is present.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdbmm.so")
CSRC = os.path.join(_HERE, "csrc")
INCLUDE = os.path.join(os.path.dirname(_HERE), "include")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--expt-relaxed-constexpr", "--extended-lambda", "-shared", "-Xcompiler", "-fPIC", "-ldl"]
SOURCES = ["dbmm_api.cu"]

MAX_H, MAX_C, MAX_G = 128, 16, 16
PHASE_GEMM1, PHASE_ROWS, PHASE_WGRAD, PHASE_UPDATE, PHASE_ALL = 1, 2, 4, 8, 15
OP_EVAL, OP_TRAIN, OP_HEAD, OP_SUPCON = 1, 2, 3, 4


class DbmmError(RuntimeError):
    pass


class AdapterPtrs(C.Structure):
    """struct dbmm_adapter"""
    _fields_ = [("W1", C.c_void_p), ("b1", C.c_void_p), ("gamma", C.c_void_p), ("beta", C.c_void_p),
                ("running_mean", C.c_void_p), ("running_var", C.c_void_p), ("num_batches_tracked", C.c_void_p),
                ("W2", C.c_void_p), ("b2", C.c_void_p)]


class BatchStats(C.Structure):
    """struct dbmm_batch_stats"""
    _fields_ = [("loss_sum", C.c_void_p), ("counts", C.c_void_p)]


class Member(C.Structure):
    """struct dbmm_member (one member of a batched sweep epoch)"""
    _fields_ = [("order", C.c_void_p), ("old_ad", C.POINTER(AdapterPtrs)), ("ad", C.POINTER(AdapterPtrs)),
                ("grads", C.c_void_p), ("momentum_buf", C.c_void_p), ("stats", BatchStats),
                ("ws", C.c_void_p), ("ws_bytes", C.c_size_t)]


def sources_newer_than_lib() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(INCLUDE, "dbmm.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a into libdbmm.so next to this file (nvcc cross-compiles without a GPU)."""
    if not force and not sources_newer_than_lib():
        return LIB_PATH
    cmd = ["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + SOURCES
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if res.returncode != 0:
        raise DbmmError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB_PATH


_lib = None

_i32, _i64, _f32, _vp, _sz = C.c_int, C.c_int64, C.c_float, C.c_void_p, C.c_size_t
_AP = C.POINTER(AdapterPtrs)

SIGNATURES = {
    "dbmm_abi_version": (C.c_int, []),
    "dbmm_last_error": (C.c_char_p, []),
    "dbmm_build_info": (C.c_char_p, []),
    "dbmm_workspace_bytes": (_sz, [_i32, _i64, _i32, _i32, _i32, _i32]),
    "dbmm_train_accum_layout": (C.c_int, [_i32, _i32, C.POINTER(_sz), C.POINTER(_sz), C.POINTER(_sz), C.POINTER(_sz)]),
    "dbmm_normalize_text": (C.c_int, [_vp, _vp, _i32, _i32, _vp]),
    "dbmm_eval_fwd": (C.c_int, [_vp, _i64, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _AP, _AP, _f32, _vp, _f32,
                                _i64, BatchStats, _vp, _vp, _vp, _sz, _vp]),
    "dbmm_eval_f16_supported": (C.c_int, [_i32, _i32, _i32]),
    "dbmm_eval_f16_workspace_bytes": (_sz, [_i64, _i32, _i32, _i32, _i32]),
    "dbmm_eval_fwd_f16": (C.c_int, [_vp, _i64, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _AP, _AP, _f32, _vp, _f32,
                                    _i64, BatchStats, _vp, _vp, _vp, _sz, _vp]),
    "dbmm_contrastive_workspace_bytes": (_sz, [_i32, _i32, _i32]),
    "dbmm_contrastive_forward": (C.c_int, [_vp, _i64, _vp, _vp, _i32, _i32, _i32, _AP, _i32, _vp, _vp, _vp, _sz, _vp]),
    "dbmm_contrastive_backward": (C.c_int, [_vp, _i64, _vp, _i32, _i32, _i32, _AP, _i32, _vp, _vp, _f32, _vp, _vp, _sz, _vp]),
    "dbmm_contrastive_apply": (C.c_int, [_i32, _i32, _i32, _AP, _vp, _vp, _f32, _f32, _f32, _i32, _vp, _sz, _vp]),
    "dbmm_contrastive_step": (C.c_int, [_vp, _i64, _vp, _vp, _i32, _i32, _i32, _AP, _i32, _f32, _f32, _vp, _vp, _f32, _f32, _f32, _i32,
                                        _vp, _vp, _vp, _sz, _vp]),
    "dbmm_device_pci_bus_id": (C.c_int, [_i32, C.c_char_p, _i32]),
    "dbmm_train_step": (C.c_int, [_i32, _vp, _i64, _vp, _vp, _vp, _i32, _i64, _i32, _i32, _i32, _i32, _AP, _AP, _f32,
                                  _vp, _f32, _vp, _vp, _f32, _f32, _f32, _i32, BatchStats, _i64, _vp, _sz, _vp]),
    "dbmm_train_step_ex": (C.c_int, [_i32, _i32, _vp, _i64, _vp, _vp, _vp, _i32, _i64, _i32, _i32, _i32, _i32, _AP, _AP, _f32,
                                     _vp, _f32, _vp, _vp, _f32, _vp, _f32, _f32, _i32, BatchStats, _i64, _vp, _sz, _vp]),
    "dbmm_train_epoch": (C.c_int, [_vp, _i64, _vp, _i64, _i32, _vp, _vp, _i32, _i32, _i32, _i32, _AP, _AP, _f32,
                                   _vp, _f32, _vp, _vp, C.POINTER(C.c_float), _f32, _f32, _i32, BatchStats,
                                   _vp, _sz, _vp]),
    "dbmm_train_forward": (C.c_int, [_vp, _i64, _vp, _i32, _i32, _i32, _i32, _AP, _AP, _f32, _vp, _f32, _vp, _i32, _vp, _sz, _vp]),
    "dbmm_train_backward": (C.c_int, [_vp, _i64, _vp, _i32, _i32, _i32, _i32, _AP, _AP, _f32, _vp, _f32, _vp, _vp, _vp, _sz, _vp]),
    "dbmm_batched_workspace_bytes": (_sz, [_i32, _i64]),
    "dbmm_train_epoch_batched": (C.c_int, [_i32, C.POINTER(Member), _vp, _i64, _i64, _i32, _vp, _vp, _i32, _i32, _i32, _i32,
                                           _f32, _vp, _f32, C.POINTER(C.c_float), _f32, _f32, _i32, _vp, _sz, _vp]),
    "dbmm_comm_unique_id": (C.c_int, [_vp]),
    "dbmm_comm_init": (C.c_int, [_vp, _i32, _i32, C.POINTER(_vp)]),
    "dbmm_comm_destroy": (C.c_int, [_vp]),
    "dbmm_comm_has_p2p": (C.c_int, [_vp]),
    "dbmm_comm_check": (C.c_int, [_vp]),
    "dbmm_train_epoch_dp": (C.c_int, [_vp, _i32, _i32, _i32, _vp, _i64, _vp, _i64, _i32, _vp, _vp, _i32, _i32, _i32, _i32, _AP, _AP,
                                      _f32, _vp, _f32, _vp, _vp, C.POINTER(C.c_float), _f32, _f32, _i32, BatchStats, _i32,
                                      _vp, _sz, _vp]),
    "dbmm_train_epoch_profile": (C.c_int, [_vp, _i64, _vp, _i64, _i32, _vp, _vp, _i32, _i32, _i32, _i32, _AP, _AP, _f32,
                                           _vp, _f32, _vp, _vp, C.POINTER(C.c_float), _f32, _f32, BatchStats,
                                           _vp, _sz, _vp, C.POINTER(C.c_float)]),
    "dbmm_sgd_step": (C.c_int, [_vp, _vp, _vp, _i64, _f32, _f32, _f32, _i32, _vp]),
    "dbmm_export_embeddings": (C.c_int, [_vp, _i64, _vp, _i64, _i32, _i32, _AP, _AP, _f32, _i32, _vp, _i32, _vp, _i32, _f32,
                                         _vp, _i64, _vp, _vp, _vp, _sz, _vp]),
    "dbmm_train_tail_mode": (C.c_int, [_i32, _i32, _i32, _i32, _i32, _i32]),
    "dbmm_widen_f16": (C.c_int, [_vp, _i64, _vp, _i64, _i64, _i32, _vp]),
    "dbmm_head_workspace_bytes": (_sz, [_i64, _i32, _i32, _i32]),
    "dbmm_logits_ce": (C.c_int, [_vp, _i64, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _vp, _vp, _f32, _i32, _i64, BatchStats, _vp,
                                 _vp, _sz, _vp]),
    "dbmm_timeline_dump": (C.c_int, [_vp, _vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "dbmm_head_f16_workspace_bytes": (_sz, [_i64, _i32, _i32]),
    "dbmm_logits_ce_f16": (C.c_int, [_vp, _i64, _vp, _vp, _i64, _i32, _i32, _i32, _vp, _vp, _f32, _i32, _i64, BatchStats, _vp,
                                     _vp, _sz, _vp]),
    "dbmm_linear_train_epoch": (C.c_int, [_vp, _i64, _vp, _i64, _i32, _vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp,
                                          C.POINTER(C.c_float), _f32, _f32, _i32, BatchStats, _vp]),
    "dbmm_supcon_workspace_bytes": (_sz, [_i32, _i32, _i32]),
    "dbmm_supcon_fwd": (C.c_int, [_vp, _i32, _i32, _i64, _i32, _vp, _f32, _vp, _vp, _vp, _vp, _sz, _vp]),
    "dbmm_supcon_bwd": (C.c_int, [_vp, _i32, _i32, _i64, _i32, _f32, _vp, _vp, _vp, _i32, _vp, _sz, _vp]),
    "dbmm_group_counts": (C.c_int, [_vp, _vp, _vp, _i64, _i32, _i32, _i64, BatchStats, _vp, _vp]),
}


def load(require_gpu: bool = True):
    """Load libdbmm.so; fails loudly (no fallback) when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise DbmmError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(nvcc, sm_100a). There is no CPU fallback for the adapter kernels.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        if lib.dbmm_abi_version() != 1:
            raise DbmmError("libdbmm.so ABI version mismatch")
        _lib = lib
    if require_gpu:
        import torch
        if not torch.cuda.is_available():
            raise DbmmError("no CUDA device: the dbmm kernels are sm_100a-only and there is no CPU fallback")
    return _lib


def check(rc: int):
    if rc != 0:
        msg = load(require_gpu=False).dbmm_last_error().decode()
        raise DbmmError(f"libdbmm error {rc}: {msg}")

"""ctypes binding of ``liblfe.so`` (C ABI declared in ``include/lfe.h``).

The CUDA library is the product: there is no CPU or PyTorch fallback.  Importing
this module without a built ``leanfe_b200/lib/liblfe.so`` raises, and every
compute entry point raises if no CUDA device is present.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "liblfe.so")

# dtype / enum mirrors of include/lfe.h
LFE_I32, LFE_I64, LFE_U32, LFE_U64, LFE_F64, LFE_U8 = 0, 1, 2, 3, 4, 5
ORDER_SORTED, ORDER_FIRST, ORDER_GLOBAL = 0, 1, 0x100
VCOV_IID, VCOV_HC1, VCOV_CLUSTER = 0, 1, 2
VCOV_CLUSTER_LOCAL = 0x100
N_TIMERS = 16
TIMER_NAMES = ["factorize", "counts_singletons", "compact", "sort", "gather", "demean_setup", "demean_sweeps",
               "demean_check", "demean_apply", "gram", "solve", "resid_vcov", "compress", "collectives", "repartition", "_15"]
ERR_CAPACITY = -3
OPT_COLUMNS_HINT, OPT_L2_BLOCK_BYTES, OPT_DETERMINISTIC = 1, 2, 3

EXPORTS = [
    "lfe_version", "lfe_last_error", "lfe_create", "lfe_destroy", "lfe_set_profiling", "lfe_reset_timers",
    "lfe_get_timers", "lfe_get_counters", "lfe_set_option", "lfe_factorize", "lfe_factorize_counts", "lfe_group_counts", "lfe_singleton_mask", "lfe_compact",
    "lfe_sort_rows", "lfe_gather", "lfe_demean", "lfe_demean_plan_create", "lfe_demean_plan_run",
    "lfe_demean_plan_destroy", "lfe_demean_plan_fit", "lfe_gram", "lfe_solve_spd", "lfe_resid_vcov", "lfe_compress", "lfe_yoco_vcov",
    "lfe_compress_merge", "lfe_key_ranges", "lfe_comm_unique_id", "lfe_comm_init", "lfe_comm_destroy", "lfe_comm_set_active",
    "lfe_allreduce_f64", "lfe_shard_plan", "lfe_shard_exchange",
    "lfe_expand_dummies", "lfe_uploader_create", "lfe_uploader_destroy", "lfe_uploader_enqueue", "lfe_uploader_stream_wait", "lfe_uploader_drain",
]


class LfeError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"liblfe error {code}: {msg}")
        self.code = code


_lib = None


def load_library() -> C.CDLL:
    """dlopen liblfe.so and declare argument types.  Fails loudly when the extension is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: the CUDA extension is not built. Run `python -c 'import __graft_entry__ as g; "
            "g.build()'` or leanfe_b200/csrc/build.sh (nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, i64, i32, dbl = C.c_void_p, C.c_int64, C.c_int, C.c_double
    P = C.POINTER
    lib.lfe_version.restype = i32
    lib.lfe_last_error.restype = C.c_char_p
    lib.lfe_create.argtypes = [i32, P(vp)]
    lib.lfe_destroy.argtypes = [vp]
    lib.lfe_set_profiling.argtypes = [vp, i32]
    lib.lfe_set_option.argtypes = [vp, i32, i64]
    lib.lfe_reset_timers.argtypes = [vp]
    lib.lfe_get_timers.argtypes = [vp, P(dbl)]
    lib.lfe_get_counters.argtypes = [vp, P(i64)]
    lib.lfe_factorize.argtypes = [vp, vp, i32, i64, i32, vp, P(i64), vp, i64, vp]
    lib.lfe_factorize_counts.argtypes = [vp, vp, i64, vp, P(i32), vp]
    lib.lfe_group_counts.argtypes = [vp, vp, i64, i64, vp, vp]
    lib.lfe_singleton_mask.argtypes = [vp, P(vp), P(i64), i32, i64, i32, P(vp), vp, P(i64), P(i64), vp]
    lib.lfe_compact.argtypes = [vp, vp, i64, P(vp), P(vp), i32, P(vp), P(vp), i32, vp]
    lib.lfe_sort_rows.argtypes = [vp, vp, i64, i64, vp, P(i32), vp]
    lib.lfe_gather.argtypes = [vp, vp, i64, P(vp), P(vp), i32, P(vp), P(vp), i32, vp]
    lib.lfe_demean.argtypes = [vp, P(vp), i32, P(vp), P(i64), P(i32), i32, i64, vp, dbl, i32, i32, i32, i32,
                               P(i32), P(dbl), P(i32), vp]
    lib.lfe_demean_plan_create.argtypes = [vp, P(vp), P(i64), i32, i64, vp, i32, P(vp), P(vp), vp]
    lib.lfe_demean_plan_run.argtypes = [vp, vp, P(vp), P(vp), i32, P(i32), dbl, i32, i32, i32, P(i32), P(dbl), P(i32), vp]
    lib.lfe_demean_plan_destroy.argtypes = [vp, vp]
    lib.lfe_demean_plan_fit.argtypes = [vp, vp, P(vp), i32, P(i32), dbl, i32, i32, i32, i32, vp, i64, i32, i64, i64,
                                        vp, vp, vp, vp, P(i32), P(dbl), P(i32), P(i64), vp]
    lib.lfe_gram.argtypes = [vp, P(vp), i32, i64, vp, i32, vp, vp]
    lib.lfe_solve_spd.argtypes = [vp, vp, vp, i32, vp, vp, P(i32), vp]
    lib.lfe_resid_vcov.argtypes = [vp, P(vp), i32, i64, vp, vp, i32, P(vp), P(i64), i32, vp, i32, i64, i64,
                                   vp, vp, vp, vp, P(i64), vp]
    lib.lfe_compress.argtypes = [vp, P(vp), P(i32), P(i64), i32, vp, vp, i64, i64, P(vp), vp, vp, vp, vp, P(i64), vp]
    lib.lfe_key_ranges.argtypes = [vp, P(vp), P(i32), i32, i64, P(i64), P(i64), vp]
    lib.lfe_yoco_vcov.argtypes = [vp, P(vp), i32, i64, vp, vp, vp, vp, vp, i32, P(vp), P(i64), i32, i32, i64, i64,
                                  vp, vp, vp, P(i64), vp]
    lib.lfe_comm_unique_id.argtypes = [vp]
    lib.lfe_comm_init.argtypes = [vp, vp, i32, i32]
    lib.lfe_comm_destroy.argtypes = [vp]
    lib.lfe_comm_set_active.argtypes = [vp, i32]
    lib.lfe_compress_merge.argtypes = [vp, P(vp), P(i64), i32, vp, vp, vp, vp, i64, i64, P(vp), vp, vp, vp, vp, P(i64), vp]
    lib.lfe_allreduce_f64.argtypes = [vp, vp, i64, vp]
    lib.lfe_expand_dummies.argtypes = [vp, vp, i64, P(i32), i32, vp, i32, P(vp), vp]
    lib.lfe_uploader_create.argtypes = [vp, i64, i32, i32, P(vp)]
    lib.lfe_uploader_destroy.argtypes = [vp]
    lib.lfe_uploader_enqueue.argtypes = [vp, vp, i64, vp, vp, P(i64)]
    lib.lfe_uploader_stream_wait.argtypes = [vp, i64, vp]
    lib.lfe_uploader_drain.argtypes = [vp]
    lib.lfe_shard_plan.argtypes = [vp, vp, i64, i64, P(i64), P(i64), P(i64), vp, vp]
    lib.lfe_shard_exchange.argtypes = [vp, vp, P(i64), P(i64), P(vp), P(i32), P(i64), i32, i64, P(vp), vp]
    for name in EXPORTS:
        if name not in ("lfe_last_error",):
            getattr(lib, name).restype = i32
    _lib = lib
    return lib


def check(rc: int):
    if rc != 0:
        raise LfeError(rc, load_library().lfe_last_error().decode("utf-8", "replace"))


def ptr_array(ptrs):
    """HOST array of device pointers."""
    arr = (C.c_void_p * max(len(ptrs), 1))()
    for i, p in enumerate(ptrs):
        arr[i] = p
    return arr


def i64_array(vals):
    arr = (C.c_int64 * max(len(vals), 1))()
    for i, v in enumerate(vals):
        arr[i] = int(v)
    return arr


def i32_array(vals):
    arr = (C.c_int * max(len(vals), 1))()
    for i, v in enumerate(vals):
        arr[i] = int(v)
    return arr

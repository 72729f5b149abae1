"""ctypes binding of libmmsim.so (include/mmsim.h).  No CPU fallback: a missing library or device is an error."""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MMSIM_LIB") or os.path.join(_HERE, "libmmsim.so")   # MMSIM_LIB: experiment builds

METRICS = {"squaredeuclidean": 0, "euclidean": 1, "l1": 2}
LOSS_BATCH_HARD, LOSS_LIFTED = 0, 1
KNN_MAX_K = 112
KNN_MAX_D = 256            # widest embedding the tcgen05 sweep takes (4 K atoms); wider ones use the exact per-query path
EVAL_SMEM_MAX_N = 16385     # largest N whose leave-one-out ranking fits the one-CTA-per-query kernel (csrc/eval.cu)

# name -> (restype, argtypes); mirrors include/mmsim.h one to one (tests/test_abi.py checks the header against this)
SIGNATURES = {
    "mmsim_version": (c_int, []),
    "mmsim_last_error": (c_char_p, []),
    "mmsim_kernel_launches": (c_int64, [c_int]),
    "mmsim_sqdist_f32": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int, c_void_p, c_int64, c_void_p]),
    "mmsim_loss_workspace_bytes": (c_int, [c_int64, c_int64, POINTER(c_size_t)]),
    "mmsim_loss_f32": (c_int, [c_int, c_void_p, c_void_p, c_int64, c_int64, c_int, c_float, c_int,
                               c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                               c_void_p, c_void_p, c_size_t, c_void_p]),
    "mmsim_knn_workspace_bytes": (c_int, [c_int64, c_int64, c_int64, c_int, POINTER(c_size_t)]),
    "mmsim_knn_f32": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int, c_int, c_int64,
                              c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "mmsim_knn_finish_f32": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int, c_int, c_int64,
                                     c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "mmsim_knn_host_workspace_bytes": (c_int, [c_int64, c_int64, c_int64, c_int, POINTER(c_size_t)]),
    "mmsim_knn_shard_fallback_f32": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int, c_int, c_int64, c_void_p, c_int,
                                             c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, c_int]),
    "mmsim_knn_shard_host_f32": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int64, c_int, c_int, c_int, c_int64,
                                         c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, c_int, c_int64, c_int64]),
    "mmsim_knn_merge_patch": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                      c_void_p, c_void_p]),
    "mmsim_knn_f32_phases": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int, c_int, c_int64,
                                     c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, c_int]),
    "mmsim_knn_host_f32": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int, c_int, c_int64,
                                   c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "mmsim_knn_shard_f32": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int, c_int, c_int, c_int64,
                                    c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, c_int, c_int64, c_int64]),
    "mmsim_knn_pivot_region": (c_int, [c_int64, c_int64, c_int64, c_int, POINTER(c_size_t), POINTER(c_size_t)]),
    "mmsim_knn_plan": (c_int, [c_int64, c_int64, c_int64, c_int, c_int, POINTER(c_int64), c_int]),
    "mmsim_knn_merge_pivots": (c_int, [c_void_p, c_int, c_int64, c_int64, c_void_p, c_void_p]),
    "mmsim_knn_merge_certified": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int, c_int64, c_int, c_int, c_void_p,
                                          c_int64, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "mmsim_semihard_mask_f32": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_int64, c_float, c_void_p, c_void_p,
                                        c_void_p]),
    "mmsim_semihard_pick_f32": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_int64, c_float, c_void_p, c_void_p]),
    "mmsim_evaluate_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_void_p, c_int64, c_double, c_int,
                                   c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mmsim_evaluate_large_workspace_bytes": (c_int, [c_int64, c_int64, POINTER(c_size_t)]),
    "mmsim_evaluate_large_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_void_p, c_int64, c_double, c_int,
                                         c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "mmsim_evaluate_ws_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_void_p, c_int64, c_double, c_int,
                                      c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, c_int]),
    "mmsim_evaluate_confusion_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p,
                                             c_void_p, c_void_p]),
    "mmsim_project_normalize_f32": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_int64, c_int, c_float, c_void_p, c_void_p]),
    "mmsim_triplet_semihard_workspace_bytes": (c_int, [c_int64, POINTER(c_size_t)]),
    "mmsim_triplet_semihard_f32": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_float, c_void_p, c_void_p, c_void_p, c_size_t,
                                           c_void_p]),
    "mmsim_lifted_struct_workspace_bytes": (c_int, [c_int64, POINTER(c_size_t)]),
    "mmsim_lifted_struct_f32": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_float, c_void_p, c_void_p, c_void_p, c_size_t,
                                        c_void_p]),
    "mmsim_knn_merge": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int, c_int64, c_int, c_void_p, c_void_p, c_void_p]),
}


class MmsimError(RuntimeError):
    pass


_lib = None


def load() -> ctypes.CDLL:
    """Load libmmsim.so, building it first when it is missing.  A library older than its sources is loaded with a warning
    (set MMSIM_AUTO_REBUILD=1 to rebuild instead: never on by default, N ranks of one job would race on the same file)."""
    global _lib
    if _lib is not None:
        return _lib
    if os.path.exists(LIB_PATH) and not os.environ.get("MMSIM_LIB"):
        try:
            from .build import needs_build
            stale = needs_build()
        except Exception:  # noqa: BLE001 -- sources not shipped with the library: nothing to compare against
            stale = False
        if stale and os.environ.get("MMSIM_AUTO_REBUILD") == "1":
            from .build import build
            build()
        elif stale:
            import warnings
            warnings.warn("libmmsim.so is older than csrc/ or include/mmsim.h; run `python -m multimodal_similarity_b200.build` "
                          "(or set MMSIM_AUTO_REBUILD=1) -- loading the existing library")
    if not os.path.exists(LIB_PATH):
        try:
            from .build import build
            build()
        except Exception as e:  # noqa: BLE001
            raise MmsimError(
                f"libmmsim.so is missing and could not be built ({e}); run `python -m multimodal_similarity_b200.build`. "
                "There is no CPU fallback.") from e
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().mmsim_last_error()
        raise MmsimError(f"{what} failed with code {rc}: {msg.decode() if msg else ''}")


def ptr(t) -> int:
    """Device pointer of a torch tensor (None -> NULL)."""
    return 0 if t is None else t.data_ptr()


def require_cuda(torch):
    if not torch.cuda.is_available():
        raise MmsimError("multimodal_similarity_b200 needs a CUDA device (sm_100a); there is no CPU fallback")

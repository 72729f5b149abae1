"""The C-ABI library loads and exports exactly what include/mmsim.h declares (no compute: runs without a GPU)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def header_functions():
    src = open(os.path.join(ROOT, "include", "mmsim.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"MMSIM_API\s+([\w\s\*]+?)\s*\b(mmsim_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = [a.strip() for a in m.group(3).replace("\n", " ").split(",")]
        out[m.group(2)] = [] if args == ["void"] else args
    return out


def test_library_builds_and_exports_every_declared_symbol():
    from multimodal_similarity_b200 import _lib
    lib = _lib.load()
    decl = header_functions()
    assert len(decl) >= 8
    assert set(decl) == set(_lib.SIGNATURES), "include/mmsim.h and the ctypes table disagree"
    for name, args in decl.items():
        assert hasattr(lib, name), f"{name} not exported"
        assert len(args) == len(_lib.SIGNATURES[name][1]), f"{name}: arity differs between header and binding"
    assert lib.mmsim_version() >= 100


def test_argument_errors_are_reported_without_a_device():
    from multimodal_similarity_b200 import _lib
    lib = _lib.load()
    n = ctypes.c_size_t()
    assert lib.mmsim_loss_workspace_bytes(4096, 128, ctypes.byref(n)) == -4      # N too large
    assert b"unsupported" in lib.mmsim_last_error()
    assert lib.mmsim_knn_workspace_bytes(0, 10, 128, 5, ctypes.byref(n)) == -1
    assert lib.mmsim_knn_workspace_bytes(1000, 100000, 128, 100, ctypes.byref(n)) == 0 and n.value > 0
    assert lib.mmsim_sqdist_f32(None, 1, None, 1, 1, 0, None, 1, None) == -1
    assert lib.mmsim_knn_merge(None, None, 0, None, 1, 1, 1, None, None, None) == -1
    assert lib.mmsim_knn_host_f32(None, 8, None, 8, 4, 1, 0, 0, None, None, None, None, None, None, 0, None) == -1
    assert b"knn_host" in lib.mmsim_last_error()


def test_every_entry_point_rejects_null_arguments():
    """All-zero / NULL arguments: each compute entry point returns MMSIM_ERR_ARG with a message naming itself, before
    touching the device (this box has none) or any pointer."""
    from multimodal_similarity_b200 import _lib
    lib = _lib.load()
    for name, (_, argtypes) in _lib.SIGNATURES.items():
        if name in ("mmsim_version", "mmsim_last_error", "mmsim_kernel_launches"):
            continue
        args = [0.0 if t in (ctypes.c_float, ctypes.c_double) else 0 if t in (ctypes.c_int, ctypes.c_int32, ctypes.c_int64,
                                                                               ctypes.c_size_t) else None for t in argtypes]
        assert getattr(lib, name)(*args) == -1, name
        assert lib.mmsim_last_error(), name


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import numpy as np
    import multimodal_similarity_b200 as mm
    x = np.zeros((4, 8), np.float32)
    with pytest.raises(mm.MmsimError):
        mm.pairwise_distance(x, x)
    with pytest.raises(mm.MmsimError):
        mm.retrieve(x, x, 2)
    with pytest.raises(mm.MmsimError):
        mm.retrieve_host(x, x, 2)
    with pytest.raises(mm.MmsimError):
        mm.batch_hard(x, np.array([1, 1, 2, 2], np.float32))
    with pytest.raises(mm.MmsimError):
        mm.evaluate(x, np.array([1, 1, 2, 2], np.int32))
    with pytest.raises(mm.MmsimError):
        mm.select_triplets_facenet(np.array([1, 1, 2, 2]), np.zeros((4, 4), np.float32), 4)
    with pytest.raises(mm.MmsimError):
        mm.project_normalize(x, np.zeros((8, 3), np.float32))
    with pytest.raises(mm.MmsimError):
        mm.triplet_semihard_loss(np.array([1, 1, 2, 2]), x)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "multimodal_similarity_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, flags=re.M), f

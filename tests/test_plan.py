"""Launch-geometry invariants of the kNN pipeline (pure host logic of the C-ABI library; runs without a GPU)."""
import ctypes

import pytest

FIELDS = ["Dp", "katoms", "n_qblocks", "n_tiles", "n_splits", "tiles_per_split", "grid", "logcap", "use_pivots",
          "n_sample_tiles", "n_sample", "total_bytes"]


def plan(nq, ng, d, k=100, sms=148):
    from multimodal_similarity_b200 import _lib
    lib = _lib.load()
    out = (ctypes.c_int64 * 12)()
    assert lib.mmsim_knn_plan(nq, ng, d, k, sms, out, 12) == 0, lib.mmsim_last_error()
    return dict(zip(FIELDS, list(out)))


@pytest.mark.parametrize("nq,ng,d", [(1, 1, 1), (100, 7, 128), (129, 4000, 256), (5924, 5924, 128), (20000, 200000, 256),
                                     (100000, 1000000, 128), (100000, 125000, 128), (100000, 10000000, 256), (300, 2049, 160)])
def test_plan_invariants(nq, ng, d):
    p = plan(nq, ng, d)
    assert p["Dp"] % 64 == 0 and p["Dp"] >= d and p["katoms"] == p["Dp"] // 64 <= 4
    assert p["n_qblocks"] == -(-nq // 128) and p["n_tiles"] == -(-ng // 256)
    # the splits tile the gallery exactly
    assert 1 <= p["n_splits"] <= p["n_tiles"]
    assert p["n_splits"] * p["tiles_per_split"] >= p["n_tiles"] > (p["n_splits"] - 1) * p["tiles_per_split"]
    assert 1 <= p["grid"] <= 148 and p["grid"] <= p["n_qblocks"] * p["n_splits"]
    # small galleries are logged whole, larger ones get a pre-pass over a systematic sample of ~1/61 of the rows
    if ng <= p["logcap"]:
        assert not p["use_pivots"]
    else:
        assert p["use_pivots"] and 1 <= p["n_sample_tiles"] <= p["n_tiles"]
        assert ng / 200 <= p["n_sample"] <= max(ng / 20, 64), (p["n_sample"], ng)
        assert p["n_sample_tiles"] == -(-p["n_sample"] // 256)
    assert p["total_bytes"] < 40e9                      # fits beside a 10M x 256 gallery in 180 GB with room to spare


def test_workspace_query_is_device_independent_upper_bound():
    from multimodal_similarity_b200 import _lib
    lib = _lib.load()
    n = ctypes.c_size_t()
    assert lib.mmsim_knn_workspace_bytes(100000, 1000000, 128, 100, ctypes.byref(n)) == 0
    for sms in (1, 64, 132, 148, 160):
        assert plan(100000, 1000000, 128, sms=sms)["total_bytes"] <= n.value
    # the host-buffer call (more gallery splits, a staging block for the sample) has its own, larger size: a device-resident
    # call no longer reserves it
    h = ctypes.c_size_t()
    assert lib.mmsim_knn_host_workspace_bytes(100000, 1000000, 128, 100, ctypes.byref(h)) == 0
    assert n.value < 4e9 and n.value < h.value < 9e9


def test_sample_rows_are_spread_and_in_range():
    """The pivot sample takes every 61st row with a per-segment offset: rows are distinct, ascending and inside the gallery
    (pure arithmetic, restated here from csrc/knn.h sample_row)."""
    def sample_row(j, div, seg):
        h = ((j // seg) * 2654435761) & 0xffffffff
        return j * div + (h >> 7) % div
    for ng in (2049, 5924, 200000, 1000000):
        p = plan(100, ng, 128)
        n, div = p["n_sample"], 61
        seg = -(-n // 16)
        rows = [sample_row(j, div, seg) for j in range(n)]
        assert n == ng // div and rows == sorted(set(rows)) and rows[-1] < ng
        assert len({r % div for r in rows}) >= min(8, n // seg)      # offsets differ between segments

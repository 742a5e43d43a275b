"""The sharding arithmetic of the multi-device handle (include/b200_multi.h) on its own, no device: the ranges the calls cut queries and
target slices into, and the merge of per-device top lists -- the same functions b200_multi_ungapped_scan runs between its device calls.
(The device side is tests/test_multi_device.py; the multi-process layout is tests/test_sharding_gloo.py.)"""
import ctypes

import numpy as np

from mmseqs2_b200 import load_library
from mmseqs2_b200.api import HIT_DTYPE, _p


def _ranges(weights, parts):
    lib = load_library()
    w = np.ascontiguousarray(weights, np.uint64)
    b = np.zeros(parts + 1, np.uint64)
    lib.b200h_balanced_ranges(_p(w), ctypes.c_uint64(len(w)), int(parts), _p(b))
    return b.astype(np.int64)


def test_balanced_ranges_cover_everything_once_and_balance_the_weight():
    rng = np.random.default_rng(1)
    for trial in range(200):
        n = int(rng.integers(1, 400))
        parts = int(rng.integers(1, 9))
        w = rng.integers(1, 3000, n) if trial % 3 else np.full(n, 7)
        b = _ranges(w, parts)
        assert b[0] == 0 and b[-1] == n and np.all(np.diff(b) >= 0)
        shares = np.array([w[b[p]:b[p + 1]].sum() for p in range(parts)])
        assert shares.sum() == w.sum()
        # no range is further from the ideal share than one element's weight
        assert np.all(np.abs(shares - w.sum() / parts) <= w.max() + 1e-9), (trial, shares)
    b = _ranges([5], 4)                                        # fewer elements than parts: empty ranges, nothing lost
    assert b[0] == 0 and b[-1] == 1 and np.all(np.diff(b) >= 0)


def test_merge_of_per_device_lists_equals_one_global_sort():
    lib = load_library()
    lib.b200h_merge_top_hits.restype = ctypes.c_uint32
    rng = np.random.default_rng(2)
    for trial in range(300):
        nd = int(rng.integers(1, 9))
        k = int(rng.integers(1, 40))
        sizes = rng.integers(1, 500, nd)                     # targets per device slice
        first = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
        lists, counts, everything = [], np.zeros(nd, np.uint32), []
        for d in range(nd):
            # the slice's own top list: scores from a small range so that ties across devices are the rule, local ids ascending within a score
            scores = rng.integers(20, 26, int(sizes[d])).astype(np.int32)
            order = np.lexsort((np.arange(sizes[d]), -scores))[:k]
            arr = np.zeros(k, HIT_DTYPE)
            arr["id"][:len(order)] = order
            arr["score"][:len(order)] = scores[order]
            lists.append(arr); counts[d] = len(order)
            everything += [(int(first[d]) + int(i), int(scores[i])) for i in range(int(sizes[d]))]
        ptrs = (ctypes.c_void_p * nd)(*[a.ctypes.data for a in lists])
        out = np.zeros(k, HIT_DTYPE)
        n = lib.b200h_merge_top_hits(ptrs, _p(counts), _p(first), nd, ctypes.c_uint32(k), _p(out))
        exp = sorted(everything, key=lambda h: (-h[1], h[0]))[:k]          # what one device holding the whole DB would report
        assert n == len(exp) and [(int(h["id"]), int(h["score"])) for h in out[:n]] == exp, trial

"""The permutation argument behind the quickselect-faithful k-d build (kdtree.cu, kd_build_faithful), as a numpy model
on the CPU: one Lomuto pass of the reference (kd_tree.f90:112-146, pivot = middle element, strict '<') is reproduced
from a prefix sum over the '<' flags --

  * '<' elements go to the front in encounter order;
  * '>=' elements form a FIFO queue whose front moves to its back at every swap after the first '>=' element, so the
    item appended by loop iteration e is the element found there or, if that one was '<', the item popped by that swap:
    item number F[e] - F[i0];
  * the queue finally holds items T..P-1 in positions store..hi-1, then the pivot is swapped in at `store`.

The CUDA kernel evaluates exactly these formulas (one thread per position); the GPU tests compare its tree with the
oracle's node by node.  Here the formulas themselves are checked against the serial loop, ties included."""
import numpy as np


def serial_pass(a, lo, hi):
    """kd_tree.f90:115-146 on rows of (key, id); returns the array and ind_store."""
    a = a.copy()
    p = (lo + hi) // 2
    pv = a[p, 0]
    a[[p, hi]] = a[[hi, p]]                               # :119-124
    store = lo
    for i in range(lo, hi):                               # :128-138
        if a[i, 0] < pv:
            a[[store, i]] = a[[i, store]]
            store += 1
    a[[store, hi]] = a[[hi, store]]                       # :141-146
    return a, store


def parallel_pass(a, lo, hi):
    n = len(a)
    out = a.copy()
    pidx = (lo + hi) // 2
    virt = lambda i: hi if i == pidx else (pidx if i == hi else i)     # the array after the pivot swap
    flag = np.zeros(n, dtype=np.int64)
    for p in range(lo, hi):
        flag[p] = a[virt(p), 0] < a[pidx, 0]
    F = np.concatenate([[0], np.cumsum(flag)[:-1]])                    # exclusive prefix sum
    base = F[lo] - lo                                                  # first '>=' position: largest i with
    l, r = lo, hi                                                      # "all of [lo, i) are '<'"
    while l < r:
        mid = (l + r + 1) // 2
        if F[mid] - mid == base:
            l = mid
        else:
            r = mid - 1
    i0 = l
    store = lo + F[hi] - F[lo]
    for p in range(lo, hi + 1):                                        # every position independently
        if p < hi and flag[p]:
            out[lo + F[p] - F[lo]] = a[virt(p)]
        if p >= store:
            if p == store:
                src = pidx
            else:
                e = store if p == hi else p
                while flag[e]:
                    e = i0 + F[e] - F[i0]
                src = virt(e)
            out[p] = a[src]
    return out, store


def test_parallel_lomuto_pass_equals_the_serial_loop():
    rng = np.random.default_rng(0)
    for _ in range(4000):
        n = int(rng.integers(2, 48))
        keys = rng.integers(0, int(rng.integers(1, 9)), n)            # few distinct values: ties everywhere
        a = np.column_stack([keys, np.arange(n)])
        lo = int(rng.integers(0, n - 1)); hi = int(rng.integers(lo + 1, n))
        s, st = serial_pass(a, lo, hi)
        q, st2 = parallel_pass(a, lo, hi)
        assert st == st2 and np.array_equal(s, q), (keys, lo, hi)


def test_degenerate_passes():
    for keys in ([3, 3, 3, 3, 3], [1, 2, 3, 4, 5], [5, 4, 3, 2, 1], [2, 1], [1, 1], [9, 0, 0, 0, 0, 0, 0]):
        a = np.column_stack([np.array(keys), np.arange(len(keys))])
        s, st = serial_pass(a, 0, len(keys) - 1)
        q, st2 = parallel_pass(a, 0, len(keys) - 1)
        assert st == st2 and np.array_equal(s, q), keys

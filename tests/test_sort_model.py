"""CPU: executable model of the limb kernel's PARALLEL form of std::sort under tied scores (csrc/opp_kernels.cu
tie_sort_parallel) against the oracle's restatement of libstdc++'s introsort (itself pinned to the real std::sort in
tests/test_oracle_golden.py).

std::sort is unstable, so with equal scores the order of the candidates - and with it the order in which greedy matching
accepts connections (src/paf.cpp:151-173) - is decided by the element movement of libstdc++'s algorithm.  The kernel
keeps that movement but runs it in parallel:
  * __unguarded_partition is a Hoare partition.  With A = positions (ascending) whose element does not beat the pivot
    and B = positions (descending) the pivot does not beat, the sequential loop swaps A[k] <-> B[k] for every k with
    A[k] < B[k] (the scans before the k-th swap only cross untouched positions) and returns
    cut = A[K] if A[K] < B[K-1] else B[K-1]  (K = number of swaps; cut = A[0] when K = 0):
    the lists come from ballots / prefix sums and the swaps are independent;
  * the two sides of a cut are independent ranges: different warps take them, level by level;
  * __final_insertion_sort is an insertion sort, i.e. the unique STABLE order of the array the partition loop leaves:
    a rank sort by (score descending, position ascending) done by the whole CTA;
  * a range that exhausts the depth limit (2 floor(lg n)) is heap-sorted sequentially, as before."""
import numpy as np
import pytest

from oracle.oracle import Oracle, CAND_DT


def gt(a, b):
    return a[0] > b[0]


def sift_down(v, first, hole, length, value):
    top = hole
    child = hole
    while child < (length - 1) // 2:
        child = 2 * (child + 1)
        if gt(v[first + child], v[first + child - 1]):
            child -= 1
        v[first + hole] = v[first + child]
        hole = child
    if (length & 1) == 0 and child == (length - 2) // 2:
        child = 2 * (child + 1)
        v[first + hole] = v[first + child - 1]
        hole = child - 1
    parent = (hole - 1) // 2
    while hole > top and gt(v[first + parent], value):
        v[first + hole] = v[first + parent]
        hole = parent
        parent = (hole - 1) // 2
    v[first + hole] = value


def heap_sort_range(v, first, last):
    length = last - first
    if length >= 2:
        parent = (length - 2) // 2
        while True:
            sift_down(v, first, parent, length, v[first + parent])
            if parent == 0:
                break
            parent -= 1
    while last - first > 1:
        last -= 1
        val = v[last]
        v[last] = v[first]
        sift_down(v, first, 0, last - first, val)


def partition_parallel(v, f, l):
    """median of three to v[f], then the Hoare partition of (f, l) in its list form; returns the cut"""
    a, b, c = f + 1, f + (l - f) // 2, l - 1
    if gt(v[a], v[b]):
        m = b if gt(v[b], v[c]) else (c if gt(v[a], v[c]) else a)
    elif gt(v[a], v[c]):
        m = a
    elif gt(v[b], v[c]):
        m = c
    else:
        m = b
    v[f], v[m] = v[m], v[f]
    piv = v[f]
    A = [p for p in range(f + 1, l) if not gt(v[p], piv)]
    B = [p for p in range(l - 1, f, -1) if not gt(piv, v[p])]
    K = 0
    while K < min(len(A), len(B)) and A[K] < B[K]:
        K += 1
    for k in range(K):
        v[A[k]], v[B[k]] = v[B[k]], v[A[k]]
    if K == 0:
        return A[0]
    return A[K] if (K < len(A) and A[K] < B[K - 1]) else B[K - 1]


def sort_model(scores):
    n = len(scores)
    v = [(np.float32(s), i) for i, s in enumerate(scores)]
    if n > 16:
        lg = int(n).bit_length() - 1
        ranges = [(0, n, 2 * lg)]
        while ranges:  # one level per round; the ranges of a level go to different warps
            nxt = []
            for (f, l, depth) in ranges:
                if depth == 0:
                    heap_sort_range(v, f, l)
                    continue
                cut = partition_parallel(v, f, l)
                for (x, y) in ((cut, l), (f, cut)):
                    if y - x > 16:
                        nxt.append((x, y, depth - 1))
            ranges = nxt
    order = sorted(range(n), key=lambda t: (-float(v[t][0]), t))  # final insertion sort = stable order of this state
    return [v[t][1] for t in order]


def check(scores):
    n = len(scores)
    c = np.zeros(n, CAND_DT)
    c["idx1"] = np.arange(n)
    c["score"] = scores
    want = Oracle.std_sort_desc(c)["idx1"].tolist()
    assert sort_model(scores) == want


def test_parallel_form_of_std_sort_equals_the_sequential_one():
    rng = np.random.default_rng(3)
    for n in list(range(0, 70)) + [100, 129, 257, 400, 1000, 1024]:
        for trial in range(5):
            nd = [max(1, n), max(1, n // 3), 5, 2, 1][trial]
            check((rng.integers(0, nd, n) / 7).astype(np.float32))
    for n in (300, 1024):  # few distinct values in runs, sorted / reversed / organ-pipe inputs
        for arr in (np.arange(n), np.arange(n)[::-1], np.concatenate([np.arange(n // 2), np.arange(n // 2)[::-1]]),
                    np.repeat(np.arange(n // 8), 8), np.repeat(np.arange(n // 8)[::-1], 8)):
            check(arr.astype(np.float32))


def median_of_3_killer(n):
    """Musser's adversary for median-of-3 quicksort: drives introsort into its depth limit (heap sort of a range)."""
    k = n // 2
    a = [0] * n
    for i in range(1, k + 1):
        if i % 2:
            a[i - 1] = i
            a[i] = k + i
        a[k + i - 1] = 2 * i
    return a


def test_depth_limit_falls_back_to_heap_sort():
    for n in (128, 512, 1024):
        a = np.array(median_of_3_killer(n), np.float32)
        check(a)
        check(-a)
        check(np.floor(a / 3))

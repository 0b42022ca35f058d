"""The claim behind orbx_sort.cuh::qs_partition_warp, checked on the CPU: libstdc++'s unguarded Hoare partition
(/usr/include/c++/13/bits/stl_algo.h:1851-1870, behind __move_median_to_first) swaps the k-th "left stopper" with the k-th
"right stopper" of the array AS IT IS BEFORE THE FIRST SWAP, for as long as the former lies left of the latter, and returns
min(first left stopper that does not swap, last swapped right stopper). The device kernel computes exactly that with ballots
and parallel swaps; the GPU parity tests hold the kernel itself against the oracle (selection order depends on this sort)."""
import numpy as np


def before(a, b):          # comp(a, b) of the quadtree's sort: larger size first (src/ORBextractor.cc:640 sorts ascending and walks it backwards)
    return a[0] > b[0]


def serial_partition(a, first, last):
    """__unguarded_partition(first + 1, last, first) on a list of (size, tag); returns the cut. The pivot is a[first]."""
    a = list(a)
    lo, hi = first + 1, last
    while True:
        while before(a[lo], a[first]):
            lo += 1
        hi -= 1
        while before(a[first], a[hi]):
            hi -= 1
        if not lo < hi:
            return lo, a
        a[lo], a[hi] = a[hi], a[lo]
        lo += 1


def stopper_partition(a, first, last):
    """The parallel formulation: everything is decided on the original array."""
    a = list(a)
    p = a[first]
    ls = [i for i in range(first + 1, last) if not before(a[i], p)]
    rs = [j for j in range(last - 1, first - 1, -1) if not before(p, a[j])]
    m = 0
    while m < len(ls) and m < len(rs) and ls[m] < rs[m]:
        m += 1
    orig = list(a)
    for k in range(m):
        a[ls[k]], a[rs[k]] = orig[rs[k]], orig[ls[k]]
    cut = min(ls[m] if m < len(ls) else 1 << 30, rs[m - 1] if m > 0 else 1 << 30)
    return cut, a


def median_to_first(a, first, last):
    """__move_median_to_first(first, first + 1, mid, last - 1): what guarantees both scans a stopper inside the range."""
    a = list(a)
    r, x, y, z = first, first + 1, first + (last - first) // 2, last - 1
    if before(a[x], a[y]):
        s = y if before(a[y], a[z]) else (z if before(a[x], a[z]) else x)
    else:
        s = x if before(a[x], a[z]) else (z if before(a[y], a[z]) else y)
    a[r], a[s] = a[s], a[r]
    return a


def test_stopper_formulation_equals_the_serial_partition():
    r = np.random.RandomState(5)
    cases = 0
    for n in list(range(17, 80)) + [100, 257, 434, 1000]:
        for span in (2, 3, 5, 17, 1000):            # few distinct sizes = many ties, the case the unstable order matters for
            for rep in range(6):
                sizes = r.randint(1, span + 1, n)
                if rep == 4: sizes = np.sort(sizes)
                if rep == 5: sizes = np.sort(sizes)[::-1]
                a = [(int(s), i) for i, s in enumerate(sizes)]
                first, last = 0, n
                if rep == 3 and n > 40:             # an inner range of a larger array
                    first, last = 7, n - 9
                a = median_to_first(a, first, last)
                c1, a1 = serial_partition(a, first, last)
                c2, a2 = stopper_partition(a, first, last)
                assert c1 == c2 and a1 == a2, (n, span, rep)
                cases += 1
    assert cases > 1500

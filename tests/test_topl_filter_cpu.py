"""CPU replay of the candidate filter of `topl_warp_kernel` (simspread.jl_b200/csrc/ss_rank.cu; reference order:
`sortperm(yhat; rev=true)`, src/performance.jl:315).

The kernel keeps, per row, the L best (key, column) pairs, key = order-preserving image of the Float64 (NaN largest,
-0.0 below +0.0).  A cell has to be ranked iff its key beats the key of rank L-1.  Since round 2 the per-cell filter is
ONE floating-point compare, `!(v <= tv)` with tv the value of that key, and only the survivors are ranked with exact
key compares; the two thresholds where the numeric order and the key order differ (tv = -0.0, tv = NaN) switch the row
to the exact compare.  The filter may let extra cells through, it must never drop a qualifying one -- checked here for
every pair of a value set that contains all the special cases."""
import itertools
import struct

import numpy as np

KEY_NEGZERO = 0x7FFFFFFFFFFFFFFF
KEY_NAN = 0xFFFFFFFFFFFFFFFF


def key(v: float) -> int:
    """isless_key of ss_rank.cu."""
    if v != v:
        return KEY_NAN
    b = struct.unpack("<Q", struct.pack("<d", v))[0]
    return (~b) & 0xFFFFFFFFFFFFFFFF if b >> 63 else b | 0x8000000000000000


def kernel_filter(v: float, thr_key: int, tv: float) -> bool:
    exact = thr_key in (KEY_NEGZERO, KEY_NAN)
    if exact:
        return key(v) > thr_key
    with np.errstate(invalid="ignore"):
        return not (np.float64(v) <= np.float64(tv))


VALUES = [float("nan"), float("inf"), -float("inf"), 0.0, -0.0, 5e-324, -5e-324, 2.2250738585072014e-308, 1.0,
          np.nextafter(1.0, 2.0), np.nextafter(1.0, 0.0), -1.0, -99.0, 0.35, 1e300, -1e300, 3.0, 3.0]


def test_key_order_is_the_reference_order():
    # stable descending sort by key == Julia's sortperm(rev=true) (isless: -0.0 < 0.0, NaN largest)
    finite = [v for v in VALUES if v == v]
    by_key = sorted(finite, key=key)
    assert all(a <= b for a, b in zip(by_key, by_key[1:]))
    assert key(-0.0) < key(0.0) and key(float("nan")) > key(float("inf"))
    assert key(-0.0) == KEY_NEGZERO


def test_filter_never_drops_a_qualifying_cell():
    for v, t in itertools.product(VALUES, VALUES):
        qualifies = key(v) > key(t)
        passes = kernel_filter(v, key(t), t)
        assert passes or not qualifies, (v, t)
        # and it is tight away from the special thresholds: nothing that ties or loses gets through
        if key(t) not in (KEY_NEGZERO, KEY_NAN) and v == v:
            assert passes == qualifies, (v, t)


def test_filtered_top_l_equals_stable_sort():
    """The whole scheme (filter + exact ranking, earlier columns win ties) against a stable descending argsort."""
    rng = np.random.default_rng(3)
    L = 5
    for trial in range(200):
        row = rng.choice(np.array(VALUES + [0.0] * 6 + [0.25, 0.5]), size=40)
        keys = [key(float(x)) for x in row]
        want = sorted(range(len(row)), key=lambda i: (-keys[i], i))[:L]
        lst = []  # (key, column), best first
        for c, x in enumerate(row):
            x = float(x)
            if len(lst) >= L:
                tk = lst[L - 1][0]
                tv = struct.unpack("<d", struct.pack("<Q", (tk & 0x7FFFFFFFFFFFFFFF) if tk >> 63 else (~tk) & 0xFFFFFFFFFFFFFFFF))[0] \
                    if tk != KEY_NAN else float("nan")
                if not kernel_filter(x, tk, tv):
                    continue
            k = key(x)
            pos = sum(1 for kk, _ in lst if kk >= k)  # rank = entries with key >= k (earlier columns win ties)
            if pos >= L:
                continue
            lst.insert(pos, (k, c))
            del lst[L:]
        assert [c for _, c in lst] == want, (trial, row)

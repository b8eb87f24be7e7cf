"""Host side of the sliced (SELL-C-sigma, `lanes` lanes per row) copy of a CSR matrix: prfdd_sell_layout / prfdd_sell_fill
(include/prfdd_b200.h).  The layout must hold every entry of every row exactly once, in row order along each lane, padded
with zero values; the device kernel that reads it is compared with the CSR kernels in tests/test_gpu_kernels.py."""
import ctypes as C

import numpy as np
import pytest


def _random_csr(rng, n, m, mean_len, long_every=0):
    lens = rng.poisson(mean_len, n)
    lens[rng.integers(0, n, max(1, n // 50))] = 0          # some empty rows
    if long_every:
        lens[::long_every] += 60                            # hanging-node-like long rows among short ones
    lens = np.minimum(lens, m)
    ptr = np.zeros(n + 1, np.int32)
    ptr[1:] = np.cumsum(lens)
    col = np.concatenate([np.sort(rng.choice(m, k, replace=False)) for k in lens] + [np.zeros(0, np.int64)]).astype(np.int32)
    val = rng.standard_normal(ptr[-1])
    return ptr, col, val


@pytest.mark.parametrize("lanes", [1, 2, 4, 8, 16, 32])
@pytest.mark.parametrize("window", [0, 64, 1000])
def test_layout_holds_every_entry_once(prfdd, lanes, window):
    L = prfdd.lib()
    L.prfdd_sell_layout.restype = C.c_longlong
    rng = np.random.default_rng(lanes * 131 + window)
    n, m = 777, 500
    ptr, col, val = _random_csr(rng, n, m, 9.0, long_every=97)
    R = 32 // lanes
    S = (n + R - 1) // R
    off = np.zeros(S + 1, np.int32)
    slot_row = np.zeros(S * R, np.int32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    total = L.prfdd_sell_layout(p(ptr), C.c_int(n), C.c_int(lanes), C.c_int(window), p(off), p(slot_row))
    assert total == off[-1] and total >= ptr[-1] and total % 32 == 0
    assert sorted(slot_row[slot_row >= 0].tolist()) == list(range(n)), "every row owns exactly one slot"
    if window <= R:
        assert np.array_equal(slot_row[:n], np.arange(n))
    else:
        # rows move only inside their window
        for w0 in range(0, n, window):
            seg = slot_row[w0:min(n, w0 + window)]
            assert seg.min() >= w0 and seg.max() < w0 + window
    scol = np.full(total, -7, np.int32)
    sval = np.full(total, np.nan)
    assert L.prfdd_sell_fill(p(ptr), p(col), p(val), C.c_int(n), C.c_int(lanes), p(off), p(slot_row), p(scol), p(sval)) == 0
    assert not np.isnan(sval).any() and (scol >= 0).all() and (scol < m).all()
    # walk the layout as the kernel does and rebuild the rows
    x = rng.standard_normal(m)
    y_ref = np.array([np.dot(val[ptr[r]:ptr[r + 1]], x[col[ptr[r]:ptr[r + 1]]]) for r in range(n)])
    y = np.zeros(n)
    for s in range(S):
        width = (off[s + 1] - off[s]) // 32
        blk_c = scol[off[s]:off[s + 1]].reshape(width, R, lanes)
        blk_v = sval[off[s]:off[s + 1]].reshape(width, R, lanes)
        for q in range(R):
            r = slot_row[s * R + q]
            if r < 0:
                assert (blk_v[:, q, :] == 0).all()
                continue
            ln = ptr[r + 1] - ptr[r]
            flat_c = blk_c[:, q, :].reshape(-1)
            flat_v = blk_v[:, q, :].reshape(-1)
            assert np.array_equal(flat_c[:ln], col[ptr[r]:ptr[r + 1]]) and np.array_equal(flat_v[:ln], val[ptr[r]:ptr[r + 1]])
            assert (flat_v[ln:] == 0).all()
            y[r] = np.dot(flat_v, x[flat_c])
    assert np.allclose(y, y_ref, rtol=1e-13, atol=1e-13)
    # FP32 values: the same layout with rounded values
    sval32 = np.zeros(total, np.float32)
    scol2 = np.zeros(total, np.int32)
    assert L.prfdd_sell_fill_f32(p(ptr), p(col), p(val), C.c_int(n), C.c_int(lanes), p(off), p(slot_row), p(scol2), p(sval32)) == 0
    assert np.array_equal(scol2, scol) and np.array_equal(sval32, sval.astype(np.float32))


def test_layout_rejects_bad_arguments(prfdd):
    L = prfdd.lib()
    L.prfdd_sell_layout.restype = C.c_longlong
    ptr = np.zeros(2, np.int32)
    a = np.zeros(64, np.int32)
    p = lambda v: v.ctypes.data_as(C.c_void_p)
    assert L.prfdd_sell_layout(p(ptr), C.c_int(1), C.c_int(3), C.c_int(0), p(a), p(a)) == -6
    assert L.prfdd_sell_layout(None, C.c_int(1), C.c_int(4), C.c_int(0), p(a), p(a)) == -8

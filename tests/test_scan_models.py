"""Host-side models of three device algorithms of the scan kernels (picovdb_b200/csrc/scan_kernel.cuh, common.cuh).

No GPU: each model restates, lane by lane, what the warp does, and checks the property the kernel relies on.
The kernels themselves are checked against the oracle in tests/test_gpu_parity.py.
  * bitonic_merge_shared / bitonic_merge_regs: max(A[i], B[31 - i]) + five compare-exchange stages leave the 32
    best keys of two descending lists, descending over the lanes.
  * block_tree_merge: a binary tree over the warps' lists ends with the block's top k in warp 0.
  * the reduce-scatter of the several-queries kernel: V = NQ * R partial sums over LPR lanes; lane `sub` ends with
    the totals of values (top << (LV - NS)) | j, and every total is the butterfly's sum (same additions).
  * sparse_walk: every set bit of (active & prefilter) is scored exactly once, in full steps except the last.
"""
import numpy as np
import pytest


def _bitonic_merge(a, b):
    """a, b: descending uint64[32] (lane i holds entry i).  Returns what the 32 lanes hold afterwards."""
    lanes = np.arange(32)
    v = np.maximum(a, b[31 - lanes])
    for j in (16, 8, 4, 2, 1):
        o = v[lanes ^ j]
        keep_max = (lanes & j) == 0
        v = np.where(keep_max == (v > o), v, o)
    return v


@pytest.mark.parametrize("seed", range(8))
def test_bitonic_merge_keeps_the_32_best_in_order(seed):
    rng = np.random.default_rng(seed)
    for k in (1, 5, 10, 32):
        a = np.zeros(32, np.uint64)
        b = np.zeros(32, np.uint64)
        ka, kb = rng.integers(0, k + 1, 2)
        keys = rng.choice(np.arange(1, 10_000, dtype=np.uint64), ka + kb, replace=False)   # distinct, like real keys
        a[:ka] = np.sort(keys[:ka])[::-1]
        b[:kb] = np.sort(keys[ka:])[::-1]
        got = _bitonic_merge(a, b)
        want = np.sort(np.concatenate([a, b]))[::-1][:32]
        np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize("n_warps", [8, 16, 5])
def test_block_tree_merge_ends_in_warp_zero(n_warps):
    rng = np.random.default_rng(n_warps)
    k = 10
    keys = rng.choice(np.arange(1, 100_000, dtype=np.uint64), n_warps * k, replace=False).reshape(n_warps, k)
    L = np.zeros((n_warps, 32), np.uint64)
    L[:, :k] = np.sort(keys, axis=1)[:, ::-1]
    slist = L[:, :k].copy()                       # store_list: the first k entries of every warp
    step = 1
    while step < n_warps:
        for w in range(n_warps):                  # one round; the barrier separates rounds
            if (w & (2 * step - 1)) == 0 and w + step < n_warps:
                b = np.zeros(32, np.uint64)
                b[:k] = slist[w + step]
                L[w] = _bitonic_merge(L[w], b)
        for w in range(n_warps):
            if (w & (2 * step - 1)) == 0 and w + step < n_warps and 2 * step < n_warps:
                slist[w] = L[w][:k]
        step <<= 1
    np.testing.assert_array_equal(L[0][:k], np.sort(keys.ravel())[::-1][:k])


@pytest.mark.parametrize("lpr,r", [(32, 4), (32, 2), (32, 8), (32, 16), (16, 4), (16, 8), (8, 4), (8, 8), (8, 2)])
def test_reduce_scatter_matches_the_butterfly(lpr, r):
    nq = 4
    v_n = nq * r
    ll, lv = int(np.log2(lpr)), int(np.log2(v_n))
    ns = min(ll, lv)
    held = 1 << (lv - ns)
    rng = np.random.default_rng(lpr * 100 + r)
    part = rng.standard_normal((lpr, v_n)).astype(np.float32)     # part[lane, value]
    lanes = np.arange(lpr)
    # the single-query kernel's butterfly, per value: acc += shfl_xor(acc, o) for o = lpr/2 ... 1
    bf = part.copy()
    o = lpr // 2
    while o:
        bf = (bf + bf[lanes ^ o]).astype(np.float32)
        o //= 2
    # reduce-scatter
    vals = part.copy()
    for st in range(ll):
        m = lpr >> (st + 1)
        upper = (lanes & m) != 0
        if st < ns:
            h = v_n >> (st + 1)
            new = vals.copy()
            for i in range(h):
                send = np.where(upper, vals[:, i], vals[:, i + h])
                keep = np.where(upper, vals[:, i + h], vals[:, i])
                new[:, i] = (keep + send[lanes ^ m]).astype(np.float32)
            vals = new
        else:
            vals[:, 0] = (vals[:, 0] + vals[lanes ^ m, 0]).astype(np.float32)
    seen = set()
    for sub in range(lpr):
        top = sub >> (ll - ns)
        primary = (sub & ((1 << (ll - ns)) - 1)) == 0
        for j in range(held):
            v = (top << (lv - ns)) | j
            assert vals[sub, j] == bf[sub, v], "same additions as the butterfly: same bits"
            if primary:
                assert v not in seen
                seen.add(v)
    assert seen == set(range(v_n)), "every (query, row) total is tested exactly once"


@pytest.mark.parametrize("rpw", [1, 2, 4, 16, 32])
@pytest.mark.parametrize("density", [0.0, 0.03, 0.5, 1.0])
def test_sparse_walk_visits_every_live_row_once_in_full_steps(rpw, density):
    rng = np.random.default_rng(int(rpw * 10 + density * 7))
    n_rows, total_warps = 5000, 7
    n_words = (n_rows + 31) // 32
    bits = rng.random(n_words * 32) < density
    bits[n_rows:] = False
    words = np.packbits(bits.reshape(-1, 32), axis=1, bitorder="little").view("<u4").ravel()
    visited = []
    for warp in range(total_warps):
        ring = np.zeros(64, np.uint32)
        head = cnt = 0
        steps = []
        wi = warp
        while wi < n_words:
            w = int(words[wi])
            c = bin(w).count("1")
            assert cnt + c <= 64
            set_bits = [b for b in range(32) if (w >> b) & 1]
            for lane in range(c):                                  # lane < popc(w): the lane-th set bit
                ring[(head + cnt + lane) & 63] = (wi << 5) + set_bits[lane]
            cnt += c
            while cnt >= rpw:
                steps.append([int(ring[(head + j) & 63]) for j in range(rpw)])
                head = (head + rpw) & 63
                cnt -= rpw
            wi += total_warps
        if cnt:
            steps.append([int(ring[(head + j) & 63]) for j in range(cnt)])
        assert all(len(s) == rpw for s in steps[:-1]), "only a warp's last step may run partly empty"
        visited += [r for s in steps for r in s]
    assert sorted(visited) == np.flatnonzero(bits).tolist()

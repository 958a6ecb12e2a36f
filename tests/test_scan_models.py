"""Host-side models of three device algorithms of the scan kernels (picovdb_b200/csrc/scan_kernel.cuh, common.cuh).

No GPU: each model restates, lane by lane, what the warp does, and checks the property the kernel relies on.
The kernels themselves are checked against the oracle in tests/test_gpu_parity.py.
  * bitonic_merge_shared / bitonic_merge_regs: max(A[i], B[31 - i]) + five compare-exchange stages leave the 32
    best keys of two descending lists, descending over the lanes.
  * block_tree_merge: a binary tree over the warps' lists ends with the block's top k in warp 0.
  * the reduce-scatter of the several-queries kernel: V = NQ * R partial sums over LPR lanes; lane `sub` ends with
    the totals of values (top << (LV - NS)) | j, and every total is the butterfly's sum (same additions).
  * sparse_walk: every set bit of (active & prefilter) is scored exactly once, in full steps except the last.
  * scan_mma_topk_kernel: the three-term bf16 split carries an fp32 query exactly, and the "k index is only a
    label" operand mapping of mma.sync.m16n8k16 (one and two queries per pass) yields whole-row dot products.
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


# ------------------------------------------------------------------ the mma.sync scan's operand mapping
def _bf16_rn(x):
    """fp32 -> bf16 (round to nearest even) -> fp32, as __float2bfloat16_rn."""
    u = np.asarray(x, np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32)


def test_three_bf16_terms_carry_an_fp32_query_exactly():
    """q = hi + mid + lo with three bf16 terms (8 + 8 + 8 mantissa bits): exact for unit-vector components."""
    x = np.random.default_rng(0).standard_normal(100_000).astype(np.float32) / np.float32(20.0)
    hi = _bf16_rn(x)
    r1 = (x - hi).astype(np.float32)
    mid = _bf16_rn(r1)
    r2 = (r1 - mid).astype(np.float32)
    lo = _bf16_rn(r2)
    np.testing.assert_array_equal(((lo.astype(np.float64) + mid) + hi), x.astype(np.float64))


@pytest.mark.parametrize("nq", [1, 2])
def test_mma_fragment_mapping_scores_whole_rows(nq):
    """scan_mma_topk_kernel feeds mma.sync.m16n8k16 with 16-byte row chunks as they lie in memory: thread
    (g, t) loads columns [8t, 8t + 8) of rows g and g + 8 of a 32-column slice and hands .x/.y to one MMA and
    .z/.w to a second one, so the instruction's k index is only a LABEL for a column -- the same relabelling
    on the query side (B operand) makes C[row, n] the dot product of the row's 16 columns with part n.
    B columns: part g % 3 of query g / 3 (0-2: the lone query; 3-5: the second query of a pair).  The model
    walks the PTX fragment layout lane by lane and checks the scores the kernel extracts (thread t == 0: first
    query, t == 1: second) against float64 dot products of the three-term split."""
    rng = np.random.default_rng(nq)
    rows = _bf16_rn(rng.standard_normal((16, 32)).astype(np.float32))          # one warp step, one 32-column slice
    q = (rng.standard_normal((nq, 32)) / 6).astype(np.float32)
    parts = np.zeros((8, 32), np.float32)                                      # B column n -> its 32 query values
    for qi in range(nq):
        hi = _bf16_rn(q[qi])
        mid = _bf16_rn((q[qi] - hi).astype(np.float32))
        lo = _bf16_rn(((q[qi] - hi).astype(np.float32) - mid).astype(np.float32))
        parts[3 * qi + 0], parts[3 * qi + 1], parts[3 * qi + 2] = hi, mid, lo
    # PTX m16n8k16 fragments of thread (g, t): A a0 = (row g, k 2t..2t+1), a1 = (row g+8, same k),
    # a2 = (row g, k 2t+8..2t+9), a3 = (row g+8, same); B b0 = (k 2t..2t+1, col g), b1 = (k 2t+8..2t+9, col g);
    # C c0,c1 = (row g, cols 2t, 2t+1), c2,c3 = (row g+8, cols 2t, 2t+1).
    C = np.zeros((16, 8), np.float64)
    for half in (0, 1):                       # ca: words .x/.y (columns 8t .. 8t+3), cb: .z/.w (8t+4 .. 8t+7)
        A = np.zeros((16, 16))
        B = np.zeros((16, 8))
        for lane in range(32):
            g, t = lane >> 2, lane & 3
            cols = [8 * t + 4 * half + j for j in range(4)]          # this thread's four columns for this MMA
            for r in (g, g + 8):
                A[r, 2 * t], A[r, 2 * t + 1] = rows[r, cols[0]], rows[r, cols[1]]          # a0 / a1
                A[r, 2 * t + 8], A[r, 2 * t + 9] = rows[r, cols[2]], rows[r, cols[3]]      # a2 / a3
            B[2 * t, g], B[2 * t + 1, g] = parts[g, cols[0]], parts[g, cols[1]]            # b0 (bq[ks][2 half])
            B[2 * t + 8, g], B[2 * t + 9, g] = parts[g, cols[2]], parts[g, cols[3]]        # b1 (bq[ks][2 half + 1])
        C += A @ B
    want = rows.astype(np.float64) @ parts.astype(np.float64).T
    np.testing.assert_allclose(C, want, rtol=0, atol=1e-12)
    # what the lanes hold and how the kernel combines it: (lo + mid) + hi
    for g in range(8):
        for r in (g, g + 8):
            c = {t: (C[r, 2 * t], C[r, 2 * t + 1]) for t in range(4)}
            s0 = (c[1][0] + c[0][1]) + c[0][0]                        # t == 0: lo from lane + 1
            assert abs(s0 - rows[r].astype(np.float64) @ q[0].astype(np.float64)) < 1e-12
            if nq == 2:
                s1 = (c[2][1] + c[2][0]) + c[1][1]                    # t == 1: (mid, lo) from lane + 1
                assert abs(s1 - rows[r].astype(np.float64) @ q[1].astype(np.float64)) < 1e-12

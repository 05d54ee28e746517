"""Host model of the warp stage of reduce_grid (lmsf-slam_b200/csrc/match.cu): the transposing butterfly that replaced
30 x 5 shuffle exchanges by 31 must produce, for every quantity, the bits the plain xor butterfly produced — same
addition tree, operands swapped at most (fp addition commutes).  The device code is the authority; this pins the
argument its comment makes, with the lane / register bookkeeping written out the same way."""
import numpy as np

LM_NSUM = 30


def xor_butterfly(v):
    """v[lane] -> every lane's total after x += shfl_xor(x, d) for d = 16, 8, 4, 2, 1 (the round-1 form)."""
    x = v.copy()
    for d in (16, 8, 4, 2, 1):
        x = x + x[np.arange(32) ^ d]
    return x


def transposing_butterfly(acc):
    """acc[lane][k] -> out[k] = what lane k holds in a[0] at the end (k < 32; quantities >= LM_NSUM are zero padding)."""
    a = np.zeros((32, 32))
    a[:, :LM_NSUM] = acc
    lanes = np.arange(32)
    for d in (16, 8, 4, 2, 1):
        hi = (lanes & d) != 0
        new = a.copy()
        for i in range(d):
            send = np.where(hi, a[:, i], a[:, i + d])
            keep = np.where(hi, a[:, i + d], a[:, i])
            new[:, i] = keep + send[lanes ^ d]          # __shfl_xor_sync(send, d)
        a = new
    return a[:, 0]


def test_transposing_butterfly_is_the_xor_butterfly_bit_for_bit():
    rng = np.random.default_rng(20260005)
    for trial in range(20):
        # wide dynamic range and mixed signs: any change of the addition tree would show in the last bits
        acc = rng.standard_normal((32, LM_NSUM)) * 10.0 ** rng.integers(-12, 12, size=(32, LM_NSUM))
        got = transposing_butterfly(acc)
        for k in range(LM_NSUM):
            want = xor_butterfly(acc[:, k])
            assert np.all(want.view(np.uint64) == want.view(np.uint64)[0])    # every lane of the old form agrees
            assert got[k].view(np.uint64) == want[0].view(np.uint64), (trial, k)
        assert np.all(got[LM_NSUM:] == 0.0)


def test_final_sum_batches_keep_the_order_of_additions():
    """The last block adds the block partials of a quantity lane by lane in ascending block order; loading
    FINAL_BATCH strides before adding them must not change that order (model of the loop in reduce_grid)."""
    rng = np.random.default_rng(7)
    for nblk in (1, 31, 32, 33, 148, 296, 297, 512, 1024):
        part = rng.standard_normal(nblk) * 10.0 ** rng.integers(-8, 8, size=nblk)
        plain = np.zeros(32)
        for b in range(nblk):
            plain[b % 32] += part[b]
        batched = np.zeros(32)
        for lane in range(32):
            b0 = lane
            while b0 < nblk:
                v = [part[b0 + 32 * u] if b0 + 32 * u < nblk else 0.0 for u in range(5)]
                for u in range(5):
                    if b0 + 32 * u < nblk:
                        batched[lane] += v[u]
                b0 += 32 * 5
        assert np.array_equal(plain.view(np.uint64), batched.view(np.uint64)), nblk

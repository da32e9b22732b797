"""Host-side restatement of the shared-memory layouts `csrc/onehot_wgrad_tc.cuh` relies on (CPU; the GPU tests
`test_onehot_conv_fwd_tc` / `test_onehot_conv_wgrad_tc` check the same facts on the hardware).

Canonical UMMA layouts for 16-bit elements, un-swizzled ("INTERLEAVE"), T = 8 elements per 16 bytes (CUTLASS
cute/atom/mma_traits_sm100.hpp, "make_umma_desc"):
    K-major : ((8, m), (T, 2)) : ((1T, SBO), (1, LBO))      element (row r, k):  (r % 8) * 16 + (r // 8) * SBO + (k % 8) * 2 + (k // 8) * LBO
    MN-major: ((T, 1, m), (8, k)) : ((1, T, SBO), (1T, LBO)) element (mn,   k):  (mn % 8) * 2 + (mn // 8) * SBO + (k % 8) * 16 + (k // 8) * LBO
(byte offsets from the descriptor's start address).  The one-hot rows of a sample are a dense [272][8] bf16 array; both
Toeplitz views below therefore have overlapping core matrices, which a descriptor is free to describe."""
import numpy as np

from oracle import embracenet_oracle as O


def kmajor_offset(r, k, lbo, sbo):
    return (r % 8) * 16 + (r // 8) * sbo + (k % 8) * 2 + (k // 8) * lbo


def mnmajor_offset(mn, k, lbo, sbo):
    return (mn % 8) * 2 + (mn // 8) * sbo + (k % 8) * 16 + (k // 8) * lbo


def onehot_rows(bases, pad, dup):
    """rows[r] = onehot(base[r - pad]) in slots 0-3 (and again in 4-7 when dup), zero outside the sequence: [272, 8]."""
    rows = np.zeros((272, 8))
    for l, c in enumerate(bases):
        rows[l + pad, c] = 1
        if dup:
            rows[l + pad, 4 + c] = 1
    return rows


def test_forward_operand_is_the_kmajor_toeplitz_view():
    rs = np.random.RandomState(0)
    k, pad = 15, 7
    bases = rs.randint(0, 4, 256)
    flat = onehot_rows(bases, pad, dup=True).reshape(-1)               # element index = byte offset / 2
    for h in range(2):                                                 # two half samples of 128 positions
        start = h * 128 * 16
        for s2 in range(8):                                            # UMMA K step: 16 K elements = two taps, + 32 bytes
            for m in rs.randint(0, 128, 16):
                for kk in range(16):
                    off = start + s2 * 32 + kmajor_offset(m, kk, lbo=16, sbo=128)
                    K = s2 * 16 + kk
                    t, slot = K // 8, K % 8
                    pos = h * 128 + m + t - pad
                    want = 1.0 if (0 <= pos < 256 and bases[pos] == slot % 4) else 0.0
                    assert flat[off // 2] == want


def test_wgrad_operand_is_the_mnmajor_toeplitz_view():
    rs = np.random.RandomState(1)
    pad = 5
    bases = rs.randint(0, 4, 256)
    flat = onehot_rows(bases, pad, dup=False).reshape(-1)
    for s2 in range(16):                                               # 16 positions per UMMA K step, + 256 bytes
        for kk in range(16):
            for n in range(128):                                       # n = tap * 8 + channel slot
                off = s2 * 256 + mnmajor_offset(n, kk, lbo=128, sbo=16)
                t, c = n // 8, n % 8
                pos = s2 * 16 + kk + t - pad
                want = 1.0 if (c < 4 and 0 <= pos < 256 and bases[pos] == c) else 0.0
                assert flat[off // 2] == want


def test_weight_matrix_fill_matches_its_descriptor():
    """The kernel writes W'[o][t * 8 + slot] at (o >> 3) * 2048 + t * 128 + (o & 7) * 16 + slot * 2 and describes it with
    LBO = 128, SBO = 2048, advancing 256 bytes per K step."""
    for o in range(64):
        for K in range(128):
            t, slot = K // 8, K % 8
            assert (o >> 3) * 2048 + t * 128 + (o & 7) * 16 + slot * 2 == (K // 16) * 256 + kmajor_offset(o, K % 16, lbo=128, sbo=2048)


def test_staging_tile_is_the_128B_swizzle():
    """Row r, 16-byte chunk ch is stored at r * 128 + ((ch ^ (r & 7)) << 4): Swizzle<3,4,3> of the linear offset, i.e. what a
    TMA tensor store with CU_TENSOR_MAP_SWIZZLE_128B expects of a 1024-byte aligned tile."""
    for r in range(32):
        for ch in range(8):
            lin = r * 128 + ch * 16
            assert lin ^ (((lin >> 7) & 7) << 4) == r * 128 + ((ch ^ (r & 7)) << 4)


def test_three_term_bf16_split_is_exact():
    rs = np.random.RandomState(2)
    w = (rs.standard_normal(20000) * np.exp(rs.uniform(-8, 8, 20000))).astype(np.float32)
    hi = O.bf16_round(w.astype(np.float64)).astype(np.float32)
    r1 = (w - hi).astype(np.float32)
    mid = O.bf16_round(r1.astype(np.float64)).astype(np.float32)
    r2 = (r1 - mid).astype(np.float32)
    lo = O.bf16_round(r2.astype(np.float64)).astype(np.float32)
    assert np.array_equal(r1.astype(np.float64), w.astype(np.float64) - hi.astype(np.float64))      # the subtractions are exact
    assert np.array_equal(lo, r2)                                                                      # the last term fits in 8 bits
    assert np.array_equal(hi.astype(np.float64) + mid.astype(np.float64) + lo.astype(np.float64), w.astype(np.float64))
    two = np.abs(hi.astype(np.float64) + mid.astype(np.float64) - w)                                  # a two-term split is not
    assert two.max() > 0

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


# ---------------------------------------------------------------------------------------------------------------------
# csrc/onehot_pool_tc.cuh (inference form of the first conv layer): the position blocks stacked on the TMEM lanes through a shifted
# window of ONE tall weight array, and the pooling walk's index arithmetic.  Same facts as tests/test_gpu_infer_fused.py, on the host.
import pytest


def _ohp_params(C1, k, Lp=124):
    nb = 128 // C1
    Pq = -(-Lp // nb)
    N = -(-(2 * Pq + 8) // 16) * 16
    Kw = 64 if k <= 7 else 128
    return dict(nb=nb, Pq=Pq, N=N, Kw=Kw, sbo=Kw * 16, ksteps=(k + 1) // 2, Lp=Lp)


@pytest.mark.parametrize('C1,k', [(16, 15), (32, 5), (64, 15), (64, 5), (32, 11)])
def test_stacked_position_blocks_and_pooling_walk_of_the_fused_first_layer(C1, k):
    rs = np.random.RandomState(C1 + k)
    p = _ohp_params(C1, k)
    pad = (k - 1) // 2
    assert 2 * (p['nb'] - 1) * p['Pq'] + p['N'] + p['Kw'] // 8 <= 288 and p['N'] <= 256        # onehot_pool_tc_ok
    W = rs.standard_normal((C1, 4, k))
    bases = rs.randint(0, 4, 256)
    # shared memory as element arrays (index = byte offset / 2): the tall weight array and the sample's one-hot rows
    w_rows = 256 - C1
    wsm = np.zeros(w_rows // 8 * p['sbo'] // 2 + 4096)
    for o in range(C1):
        for t in range(p['Kw'] // 8):
            for c in range(4):
                v = W[o, c, t] if t < k else 0.0
                R = 128 - C1 + o
                off = (R >> 3) * p['sbo'] + t * 128 + (R & 7) * 16
                wsm[off // 2 + c] = 0.75 * v                               # stand-ins for the {hi, mid} split: they add up to v
                wsm[off // 2 + 4 + c] = 0.25 * v
    rows = np.zeros((288, 8))
    for l, c in enumerate(bases):
        rows[l + pad, c] = rows[l + pad, 4 + c] = 1.0
    xs = rows.reshape(-1)
    # the MMAs: D[128, N] += A(128 x 16) . B(N x 16)^T for every block q and K step s, operands read through their descriptors
    D = np.zeros((128, p['N']))
    mrow, kk = np.arange(128), np.arange(16)
    nrow = np.arange(p['N'])
    a_off = kmajor_offset(mrow[:, None], kk[None, :], lbo=128, sbo=p['sbo'])
    b_off = kmajor_offset(nrow[:, None], kk[None, :], lbo=16, sbo=128)
    for q in range(p['nb']):
        wa = ((128 - C1 - q * C1) >> 3) * p['sbo']
        xa = 2 * q * p['Pq'] * 16
        for s in range(p['ksteps']):
            A = wsm[(wa + s * 256 + a_off) // 2]
            B = xs[(xa + s * 32 + b_off) // 2]
            D += A @ B.T
    # lanes [q C1, (q + 1) C1) x columns n  ==  conv output of channel o at position 2 q Pq + n
    conv = np.zeros((C1, 256 + 64))
    for l in range(256):
        for t in range(k):
            pos = l + t - pad
            if 0 <= pos < 256:
                conv[:, l] += W[:, bases[pos], t]
    for q in range(p['nb']):
        for n in range(p['N']):
            l = 2 * q * p['Pq'] + n
            if l < 256:
                assert np.allclose(D[q * C1:(q + 1) * C1, n], conv[:, l], atol=1e-12), (q, n)
    # the pooling walk (pool_walk in conv_pool_tc.cuh) with the kernel's index arithmetic, every `parts` setting
    sc, sh = rs.standard_normal(C1), rs.standard_normal(C1)
    z = np.maximum(conv[:, :256] * sc[:, None] + sh[:, None], 0.0)
    want = np.stack([z[:, 2 * j:2 * j + 10].max(axis=1) for j in range(p['Lp'])], axis=0)          # [Lp, C1]
    for parts in (1, 2, 4):
        out = np.full((p['Lp'], C1), np.nan)
        for lane in range(128):
            qb, ch = lane // C1, lane % C1
            per = -(-p['Pq'] // parts)
            for part in range(parts):
                lo = min(part * per, p['Pq'])
                hi = min(lo + per, p['Pq'])
                valid = max(0, min(p['Pq'], p['Lp'] - qb * p['Pq']))
                n_eff = max(0, min(hi, valid) - lo)
                c_first = (2 * lo) & ~15
                kq = (c_first >> 1) - (lo + 4)
                ring = [0.0, 0.0, 0.0, 0.0]
                c_end = 2 * (hi + 4)
                assert c_end <= p['N']
                for c16 in range(c_first, c_end, 16):                      # 16-column TMEM loads; whole chunks are walked
                    for pp in range(8):
                        v0, v1 = D[lane, c16 + 2 * pp], D[lane, c16 + 2 * pp + 1]
                        pm = max(v0 * sc[ch] + sh[ch], v1 * sc[ch] + sh[ch])
                        rv = max(max(ring), max(pm, 0.0))
                        if 0 <= kq + pp < n_eff:
                            j = qb * p['Pq'] + lo + kq + pp
                            assert np.isnan(out[j, ch])                       # every pooled element is written exactly once
                            out[j, ch] = rv
                        ring[pp % 4] = pm
                    kq += 8
        assert not np.isnan(out).any()
        assert np.allclose(out, want, atol=1e-12), parts

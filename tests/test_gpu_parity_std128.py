"""GPU parity at the FULL parameter sets of BASELINE.json configs[2..4] (oracle-sized batches):
STD128 AP/DM gates, STD128 functional bootstrapping (logQ=12, N=2048, 54-bit Q, 4.8 GB key-switching table),
STD128 EvalSign/EvalDecomp (logQ=17) followed by CiphertextMulMatrix.  Marked slow: key generation dominates."""
import numpy as np
import pytest

from oracle import pyoracle as po

pytestmark = [pytest.mark.gpu, pytest.mark.slow]


def test_std128_ap_gates(keyset):
    """configs[2]: DM accumulator with its 2.1 GB (u32) bootstrapping key and data-dependent key selection."""
    ks = keyset("std128_ap")
    q = ks.p.q
    m1 = [i & 1 for i in range(8)]
    m2 = [(i >> 1) & 1 for i in range(8)]
    c1 = ks.port.encrypt_batch(ks.sk, m1, 4, q, 21)
    c2 = ks.port.encrypt_batch(ks.sk, m2, 4, q, 22)
    for gate in ("NAND", "XOR_FAST"):
        want = ks.port.eval_bin_gate(ks.bk, ks.ksk, po.GATES[gate], c1, c2, q)
        got = ks.gpu().EvalBinGate(gate, c1, c2)
        assert np.array_equal(got, want), gate
    assert ks.port.decrypt_batch(ks.sk, got, q, 4) == [a ^ b for a, b in zip(m1, m2)]


def test_std128_ap_specialised_and_generic_kernels_agree(keyset, rng):
    ks = keyset("std128_ap")
    q, n = ks.p.q, ks.p.n
    c1 = rng.integers(0, q, (21, n + 1), dtype=np.uint64)    # ragged vs the CTA group of 4
    c2 = rng.integers(0, q, (21, n + 1), dtype=np.uint64)
    c1[0, :n] = 0
    c2[0, :n] = 0                                            # every refresh digit is zero: all steps sit out
    g = ks.gpu()
    assert g.kernel_variant.startswith("dm_u32")
    a = g.EvalBinGate("NAND", c1, c2)              # 21 ciphertexts: the latency layout (one ciphertext per CTA)
    try:
        for grp in (4, 2):                         # the throughput shape (4 per CTA) and the two-per-CTA shape
            g.set_option("group", grp)
            assert np.array_equal(g.EvalBinGate("NAND", c1, c2), a), grp
    finally:
        g.set_option("group", 0)
    g.set_option("force_generic", 1)
    try:
        b = g.EvalBinGate("NAND", c1, c2)
    finally:
        g.set_option("force_generic", 0)
    assert np.array_equal(a, b)
    want = ks.port.eval_bin_gate(ks.bk, ks.ksk, po.GATES["NAND"], c1[:3], c2[:3], q)
    assert np.array_equal(a[:3], want)


def test_std128_evalfunc_logq12(keyset):
    """configs[3]: arbitrary LUT, two chained bootstraps at N=2048 / 54-bit Q / qKS=2^35."""
    ks = keyset("std128_func12")
    q = ks.p.q
    p = q // (2 * ks.p.beta)
    lut = np.array([((x // (q // p)) ** 3 % p) * (q // p) for x in range(q)], dtype=np.uint64)
    msgs = list(range(p))
    ct = ks.port.encrypt_batch(ks.sk, msgs, p, q, 23)
    want = ks.port.eval_func(ks.bk, ks.ksk, ct, q, lut)
    got = ks.gpu().EvalFunc(ct, lut)
    assert np.array_equal(got, want)
    assert ks.port.decrypt_batch(ks.sk, got, q, p) == [m ** 3 % p for m in msgs]


def test_std128_sign_decomp_mulmatrix_logq17(keyset, rng):
    """configs[4]: EvalSign (5 bootstraps) + EvalDecomp (4 bootstraps, 3 digits) at logQ=17, then the digits go through
    CiphertextMulMatrix."""
    ks = keyset("std128_sign17")
    Qin, q = 1 << 17, ks.p.q
    P = Qin // q * (q // (2 * ks.p.beta))
    msgs = [P // 2 + i - 2 for i in range(4)]
    ct = ks.port.encrypt_batch(ks.sk, msgs, P, Qin, 24)
    want = ks.port.eval_sign(ks.bk, ks.ksk, ct, Qin)
    got = ks.gpu().EvalSign(ct, Qin)
    assert np.array_equal(got, want)
    assert ks.port.decrypt_batch(ks.sk, got, q, 2) == [int(m >= P // 2) for m in msgs]
    wd, wm = ks.port.eval_decomp(ks.bk, ks.ksk, ct, Qin)
    gd, gm = ks.gpu().EvalDecomp(ct, Qin)
    assert gm == wm and np.array_equal(gd, wd)
    M = rng.integers(0, 64, (4, 6), dtype=np.int64)
    digits0 = np.ascontiguousarray(gd[:, 0, :])
    assert np.array_equal(ks.gpu().CiphertextMulMatrix(digits0, M, q), ks.port.mul_matrix(digits0, M, q))

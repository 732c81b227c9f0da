"""GPU parity: EvalFunc / EvalFloor / EvalSign / EvalDecomp / CiphertextMulMatrix through the C ABI vs the oracle.

Cases mirror the reference's UnitTestFunc.cpp (x^3 mod p LUT over all inputs, floor around p/2, sign and digit
decomposition of large-precision inputs) and examples/GEMM.cpp, at the bit level and through the batched API.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _lut_cube(q, p):
    return np.array([((x // (q // p)) ** 3 % p) * (q // p) for x in range(q)], dtype=np.uint64)


@pytest.mark.parametrize("name", ["toy_func12", "toy_func12_throw1"])
def test_eval_func_arbitrary_lut(keyset, name):
    ks = keyset(name)
    q = ks.p.q
    p = q // (2 * ks.p.beta)
    lut = _lut_cube(q, p)
    msgs = [i % p for i in range(2 * p)]
    ct = ks.port.encrypt_batch(ks.sk, msgs, p, q, 11)
    want = ks.port.eval_func(ks.bk, ks.ksk, ct, q, lut)
    got = ks.gpu().EvalFunc(ct, lut)
    assert np.array_equal(got, want)
    assert ks.gpu().last_stats.bootstraps == 2
    if name == "toy_func12":
        assert ks.port.decrypt_batch(ks.sk, got, q, p) == [m ** 3 % p for m in msgs]


def test_eval_func_negacyclic_and_periodic(keyset):
    ks = keyset("toy_func12")
    q = ks.p.q
    p = q // (2 * ks.p.beta)
    ct = ks.port.encrypt_batch(ks.sk, list(range(p)), p, q, 12)
    neg = np.array([(x % (q // 2)) + 1 for x in range(q)], dtype=np.uint64)
    neg[q // 2:] = q - neg[: q // 2]
    per = np.array([((x % (q // 2)) * 5) % q for x in range(q)], dtype=np.uint64)
    for lut, nboot in ((neg, 1), (per, 2)):
        want = ks.port.eval_func(ks.bk, ks.ksk, ct, q, lut)
        got = ks.gpu().EvalFunc(ct, lut)
        assert np.array_equal(got, want)
        assert ks.gpu().last_stats.bootstraps == nboot


def test_eval_func_lut_vec(keyset):
    """EvalFunc(vector, LUT_vec): one LUT per ciphertext (binfhe-base-scheme.cpp:791-924)."""
    ks = keyset("toy_func12")
    q = ks.p.q
    p = q // (2 * ks.p.beta)
    msgs = list(range(p))
    ct = ks.port.encrypt_batch(ks.sk, msgs, p, q, 13)
    luts = np.stack([np.array([(((x // (q // p)) * (k + 1) + k) % p) * (q // p) for x in range(q)], dtype=np.uint64)
                     for k in range(len(msgs))])
    want = ks.port.eval_func(ks.bk, ks.ksk, ct, q, luts)
    got = ks.gpu().EvalFunc(ct, luts)
    assert np.array_equal(got, want)
    assert ks.port.decrypt_batch(ks.sk, got, q, p) == [(m * (k + 1) + k) % p for k, m in enumerate(msgs)]


def test_eval_func_small_ring_logq11(keyset):
    """logQ = 11: N=1024, 27-bit Q, baseG=32 (12 digit polynomials), q = 2N (negacyclic / periodic LUTs only)."""
    ks = keyset("toy_func11")
    q = ks.p.q
    p = q // (2 * ks.p.beta)
    ct = ks.port.encrypt_batch(ks.sk, list(range(p)), p, q, 14)
    per = np.array([((x // (q // p)) % (p // 2)) * (q // p) for x in range(q)], dtype=np.uint64)
    per[q // 2:] = per[: q // 2]
    want = ks.port.eval_func(ks.bk, ks.ksk, ct, q, per)
    got = ks.gpu().EvalFunc(ct, per)
    assert np.array_equal(got, want)


def _big_inputs(ks, logQ, count):
    Qin = 1 << logQ
    q = ks.p.q
    P = Qin // q * (q // (2 * ks.p.beta))
    msgs = [P // 2 + i - 3 for i in range(count)]
    return Qin, P, msgs, ks.port.encrypt_batch(ks.sk, msgs, P, Qin, 15)


def test_eval_floor(keyset):
    ks = keyset("toy_sign17")
    Qin, P, msgs, ct = _big_inputs(ks, 17, 8)
    want = ks.port.eval_floor(ks.bk, ks.ksk, ct, Qin)
    got = ks.gpu().EvalFloor(ct, Qin)
    assert np.array_equal(got, want)
    assert ks.gpu().last_stats.bootstraps == 2


def test_eval_sign(keyset):
    ks = keyset("toy_sign17")
    Qin, P, msgs, ct = _big_inputs(ks, 17, 8)
    want = ks.port.eval_sign(ks.bk, ks.ksk, ct, Qin)
    got = ks.gpu().EvalSign(ct, Qin)
    assert np.array_equal(got, want)
    assert ks.gpu().last_stats.bootstraps == 5           # logQ=17: 2 floors (4) + final
    assert ks.port.decrypt_batch(ks.sk, got, ks.p.q, 2) == [int(m >= P // 2) for m in msgs]


def test_eval_decomp(keyset):
    ks = keyset("toy_sign17")
    Qin, P, msgs, ct = _big_inputs(ks, 17, 8)
    want, wmods = ks.port.eval_decomp(ks.bk, ks.ksk, ct, Qin)
    got, gmods = ks.gpu().EvalDecomp(ct, Qin)
    assert gmods == wmods == [4096, 4096, 512]
    assert np.array_equal(got, want)
    assert ks.gpu().last_stats.bootstraps == 4


def test_ciphertext_mul_matrix(keyset, rng):
    """examples/GEMM.cpp: random ciphertexts x matrix with entries < 64, element-wise vs the exact CPU product."""
    ks = keyset("toy_func12")
    n, qKS = ks.p.n, ks.p.qKS
    ct = rng.integers(0, qKS, (48, n + 1), dtype=np.uint64)
    M = rng.integers(0, 64, (48, 40), dtype=np.int64)
    want = ks.port.mul_matrix(ct, M, qKS)
    got = ks.gpu().CiphertextMulMatrix(ct, M, qKS)
    assert np.array_equal(got, want)
    # independent check with Python big integers
    ref = (ct.astype(object).T @ M.astype(object)) % int(qKS)
    assert np.array_equal(got.astype(object), ref.T)
    # negative entries: Euclidean residue (the reference's FP64 path is undefined there, lwe-operation.cu:123)
    M2 = rng.integers(-64, 64, (48, 5), dtype=np.int64)
    got2 = ks.gpu().CiphertextMulMatrix(ct, M2, qKS)
    ref2 = (ct.astype(object).T @ M2.astype(object)) % int(qKS)
    assert np.array_equal(got2.astype(object), ref2.T)


def test_mkmswitch_double_rounding_near_ties(keyset):
    """The first mod-switch (Q ~ 2^54 -> qKS = 2^35) must reproduce the reference's DOUBLE rounding on near-ties,
    where it differs from exact integer rounding (lwe-pke.cpp:41-46)."""
    ks = keyset("toy_func12")
    p = ks.p
    inv = pow(1 << 36, -1, p.Q)
    vals = np.array([(d * inv) % p.Q for d in range(-600, 600) if d], dtype=np.uint64)
    ext = np.resize(vals, (2, p.N + 1))
    want = ks.port.mkmswitch(ks.ksk, ext, p.q)
    got = ks.gpu().MKMSwitch(ext, p.q)
    assert np.array_equal(got, want)
    exact = [((2 * int(v) * p.qKS + p.Q) // (2 * p.Q)) % p.qKS for v in ext[0]]
    dbl = [ks.port.round_qQ(int(v), p.qKS, p.Q) for v in ext[0]]
    assert exact != dbl      # the vectors really exercise the discrepancy


@pytest.mark.parametrize("name", ["toy_func12", "toy_sign17", "toy_func12_throw1"])
def test_u64_specialised_and_generic_kernels_agree(keyset, rng, name):
    """br_cggi64 (register-resident, shuffle stage) vs br_generic<u64> on random inputs, ragged batch."""
    ks = keyset(name)
    p = ks.p
    g = ks.gpu()
    assert g.kernel_variant.startswith("cggi_u64")
    ct = rng.integers(0, p.q, (5, p.n + 1), dtype=np.uint64)
    tab = rng.integers(0, p.q, p.q, dtype=np.uint64)
    a = g.BootstrapFunc(ct, p.q, tab, p.q)
    g.set_option("force_generic", 1)
    try:
        b = g.BootstrapFunc(ct, p.q, tab, p.q)
    finally:
        g.set_option("force_generic", 0)
    assert np.array_equal(a, b)
    assert np.array_equal(a, ks.port.bootstrap_func(ks.bk, ks.ksk, ct, p.q, tab, p.q))
    # explicit-accumulator entry point (EvalAcc_CUDA contract) on the same kernel
    acc = rng.integers(0, p.Q, (3, 2, p.N), dtype=np.uint64)
    am = rng.integers(0, p.q, (3, p.n), dtype=np.uint64)
    assert np.array_equal(g.EvalAcc(am, p.q, acc), ks.port.eval_acc(ks.bk, am, p.q, acc))


def test_functional_error_behaviour(keyset):
    """Same preconditions as the reference's OPENFHE_THROWs (binfhe-base-scheme.cpp:682-686, 709-713, 1050-1054)."""
    from tfhe_gpu_b200 import TfheB200Error

    ks = keyset("toy_sign17")          # q = 2N = 4096: arbitrary LUTs are not allowed (q > N)
    g = ks.gpu()
    q, n = ks.p.q, ks.p.n
    ct = np.zeros((2, n + 1), dtype=np.uint64)
    arb = np.arange(q, dtype=np.uint64)
    arb[0], arb[q // 2] = 1, 7
    with pytest.raises(TfheB200Error, match="needs to be <= ring dimension"):
        g.EvalFunc(ct, arb)
    with pytest.raises(TfheB200Error, match="only for large precision"):
        g.EvalDecomp(ct, q)
    with pytest.raises(TfheB200Error, match="input vector is empty"):
        g.EvalSign(np.zeros((0, n + 1), dtype=np.uint64), 1 << 17)
    with pytest.raises(TfheB200Error, match="LUT length"):
        g.EvalFunc(ct, arb[: q // 2])
    with pytest.raises(TfheB200Error, match="number of rows"):
        g.CiphertextMulMatrix(ct, np.ones((3, 2), dtype=np.int64), q)
    with pytest.raises(TfheB200Error, match="unmatched with LUT size"):
        g.EvalFunc(ct, np.stack([arb, arb, arb]))


@pytest.mark.parametrize("modulus", [1 << 17, 1 << 12, 2, 3, 12289, (1 << 29) - 3, (1 << 31) - 1, 1 << 32, (1 << 32) + 15,
                                     1 << 35])
def test_ciphertext_mul_matrix_moduli(keyset, rng, modulus):
    """Every arithmetic path of the product: power-of-two and odd moduli on the 32-bit register-tiled kernel (with and
    without intermediate reductions), moduli above 2^29.5 / 2^32 on the exact 128-bit kernel; ragged tile shapes,
    entries of both signs and beyond the modulus, operands not reduced on input."""
    ks = keyset("toy_func12")
    n = ks.p.n
    rows, cols = 150, 70                                                   # not multiples of the 16 / 64 / 128 tiles
    ct = rng.integers(0, 1 << 40, (rows, n + 1), dtype=np.uint64)          # NOT reduced
    ct[0] = modulus - 1
    M = rng.integers(-(1 << 40), 1 << 40, (rows, cols), dtype=np.int64)
    M[:, 0] = modulus - 1
    ct[:, 0] = modulus - 1                                                 # worst-case accumulation: rows * (m-1)^2
    got = ks.gpu().CiphertextMulMatrix(ct, M, modulus)
    ref = (ct.astype(object).T @ M.astype(object)) % int(modulus)
    assert np.array_equal(got.astype(object), ref.T)
    assert np.array_equal(got, ks.port.mul_matrix(ct, M, modulus))


def test_large_precision_logq29(keyset):
    """UnitTestFunc.cpp:150-265 runs EvalSign / EvalDecomp at logQ = 29 (space-optimised: one key, baseG = 2^14, four
    digits -> the 4-digit instantiation of the 64-bit kernel): sign, digits and floor against the oracle, digits pinned
    by decryption."""
    ks = keyset("toy_sign29")
    Qin, P, msgs, ct = _big_inputs(ks, 29, 6)
    g = ks.gpu()
    want = ks.port.eval_sign(ks.bk, ks.ksk, ct, Qin)
    got = g.EvalSign(ct, Qin)
    assert np.array_equal(got, want)
    assert ks.port.decrypt_batch(ks.sk, got, ks.p.q, 2) == [int(m >= P // 2) for m in msgs]
    wd, wm = ks.port.eval_decomp(ks.bk, ks.ksk, ct, Qin)
    gd, gm = g.EvalDecomp(ct, Qin)
    assert gm == wm and np.array_equal(gd, wd)
    assert np.array_equal(g.EvalFloor(ct, Qin), ks.port.eval_floor(ks.bk, ks.ksk, ct, Qin))
    # the digits recompose the message (base 16 digits, the last one under the shrunken modulus)
    beta, q = ks.p.beta, ks.p.q
    for j, m in enumerate(msgs):
        total, scale = 0, 1
        for k, mod in enumerate(gm):
            pk = (q if mod == q else int(mod)) // (2 * beta)
            total += ks.port.decrypt_batch(ks.sk, np.ascontiguousarray(gd[j:j + 1, k, :]), int(mod), pk)[0] * scale
            scale *= pk
        assert total % P == m


@pytest.mark.parametrize("name", ["toy_func12", "toy_sign17", "toy_func12_throw1", "toy_sign17_throw1"])
def test_cggi64_cta_shapes_agree(keyset, rng, name):
    """The wide 64-bit kernel runs two ciphertexts per CTA (throughput) or one (batches of at most one ciphertext per SM,
    picked automatically): same bits from both shapes and from the oracle, on a ragged batch.  The thrown-digit sets (the
    reference's own timing configurations, time-estimate.cpp:59-190) take its plain path (no top-digit elimination)."""
    ks = keyset(name)
    p = ks.p
    g = ks.gpu()
    assert g.kernel_variant.startswith("cggi_u64_ntt16x128")
    ct = rng.integers(0, p.q, (7, p.n + 1), dtype=np.uint64)
    tab = rng.integers(0, p.q, p.q, dtype=np.uint64)
    outs = {}
    try:
        for grp in (0, 1, 2):
            g.set_option("group", grp)
            outs[grp] = g.BootstrapFunc(ct, p.q, tab, p.q)
    finally:
        g.set_option("group", 0)
    assert np.array_equal(outs[1], outs[2]) and np.array_equal(outs[0], outs[2])
    assert np.array_equal(outs[1], ks.port.bootstrap_func(ks.bk, ks.ksk, ct, p.q, tab, p.q))


def test_floor_logq11_one_thrown_digit(keyset, rng):
    """The reference's EvalFloor timing set (time-estimate.cpp:96-123: logQ = 11, numDigitsToThrow = 1): N = 1024, 27-bit Q,
    baseG = 32, 5 kept digits -- the 5-digit instantiation of the 32-bit kernel; floor and a raw functional bootstrap
    against the oracle, specialised and generic kernels against each other."""
    ks = keyset("toy_func11_throw1")
    p = ks.p
    g = ks.gpu()
    assert p.numDigitsToThrow == 1 and p.digitsG - p.numDigitsToThrow == 5 and g.kernel_variant == "cggi_u32_ntt32"
    ct = rng.integers(0, p.q, (9, p.n + 1), dtype=np.uint64)
    assert np.array_equal(g.EvalFloor(ct, p.q), ks.port.eval_floor(ks.bk, ks.ksk, ct, p.q))
    tab = rng.integers(0, p.q, p.q, dtype=np.uint64)
    a = g.BootstrapFunc(ct, p.q, tab, p.q)
    assert np.array_equal(a, ks.port.bootstrap_func(ks.bk, ks.ksk, ct, p.q, tab, p.q))
    g.set_option("force_generic", 1)
    try:
        assert np.array_equal(g.BootstrapFunc(ct, p.q, tab, p.q), a)
    finally:
        g.set_option("force_generic", 0)

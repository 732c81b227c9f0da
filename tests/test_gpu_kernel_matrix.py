"""Every template instantiation of the specialised kernels, on small custom rings (short LWE dimension, so the
oracle is instant): N in {512, 1024} x digitsG in {2, 3, 4, 6} for br_cggi32 (with and without top-digit elimination,
as the safety predicate decides), the AP/DM kernel, ragged batches, all-zero masks, gate and LUT accumulators, and
the explicit-accumulator operator entry point.  Keys come from the oracle's deterministic key generator."""
import numpy as np
import pytest

from oracle import pyoracle as po

pytestmark = pytest.mark.gpu

Q27 = 134215681


def _ctx(p, port, seed=5):
    from tfhe_gpu_b200 import BinFHEContextB200

    sk, bk, ksk = port.keygen(seed)
    return sk, bk, ksk, BinFHEContextB200().GPUSetup(p.as_dict(), bk, ksk, numGPUs=1)


@pytest.mark.parametrize("N", [512, 1024])
@pytest.mark.parametrize("baseG", [1 << 14, 1 << 9, 1 << 7, 1 << 5])      # digitsG = 2, 3, 4, 6
def test_cggi32_instantiations(N, baseG, rng):
    p = po.Port.params_custom(12, N, N, Q27, 128, baseG, 32, po.GINX)       # q = N (so LUTs of every class are allowed)
    port = po.Port(p)
    sk, bk, ksk, g = _ctx(p, port)
    try:
        assert g.kernel_variant.startswith("cggi_u32_ntt32")
        q, n = p.q, p.n
        c1 = rng.integers(0, q, (11, n + 1), dtype=np.uint64)              # ragged vs every CTA group size
        c2 = rng.integers(0, q, (11, n + 1), dtype=np.uint64)
        c1[0, :n] = 0
        c2[0, :n] = 0                                                      # all rotation exponents zero
        for gate in ("NAND", "XNOR_FAST"):
            want = port.eval_bin_gate(bk, ksk, po.GATES[gate], c1, c2, q)
            assert np.array_equal(g.EvalBinGate(gate, c1, c2), want), gate
        tab = rng.integers(0, q, (11, q), dtype=np.uint64)                 # per-ciphertext tables
        want = port.bootstrap_func(bk, ksk, c1, q, tab, q)
        assert np.array_equal(g.BootstrapFunc(c1, q, tab, q), want)
        half = q // 2                                                      # smaller ciphertext modulus: factor 2N/q = 4
        ch = c1 % half
        th = rng.integers(0, half, half, dtype=np.uint64)
        assert np.array_equal(g.BootstrapFunc(ch, half, th, q), port.bootstrap_func(bk, ksk, ch, half, th, q))
        acc = rng.integers(0, p.Q, (3, 2, N), dtype=np.uint64)
        am = rng.integers(0, q, (3, n), dtype=np.uint64)
        assert np.array_equal(g.EvalAcc(am, q, acc), port.eval_acc(bk, am, q, acc))
        # the generic kernel agrees as well
        g.set_option("force_generic", 1)
        assert np.array_equal(g.EvalBinGate("NAND", c1, c2), port.eval_bin_gate(bk, ksk, po.GATES["NAND"], c1, c2, q))
    finally:
        g.GPUClean()


def test_dm32_small_ring(rng):
    p = po.Port.params_custom(12, 1024, 1024, Q27, 128, 1 << 7, 32, po.AP)
    port = po.Port(p)
    sk, bk, ksk, g = _ctx(p, port)
    try:
        assert g.kernel_variant.startswith("dm_u32")
        q, n = p.q, p.n
        c1 = rng.integers(0, q, (7, n + 1), dtype=np.uint64)
        c2 = rng.integers(0, q, (7, n + 1), dtype=np.uint64)
        c1[1, :n] = 0
        c2[1, :n] = 0
        c1[2, :n] = 32                                                     # low refresh digit zero, high digit non-zero
        c2[2, :n] = 0
        want = port.eval_bin_gate(bk, ksk, po.GATES["NOR"], c1, c2, q)
        assert np.array_equal(g.EvalBinGate("NOR", c1, c2), want)
        acc = rng.integers(0, p.Q, (2, 2, 1024), dtype=np.uint64)
        am = rng.integers(0, q, (2, n), dtype=np.uint64)
        assert np.array_equal(g.EvalAcc(am, q, acc), port.eval_acc(bk, am, q, acc))
    finally:
        g.GPUClean()


Q28 = 268369921     # the 28-bit prime of MEDIUM / SIGNED_MOD_TEST (binfhecontext.cpp:140,155)
Q29 = 536813569     # the 29-bit prime of STD256


@pytest.mark.parametrize("N,Q,baseG,variant", [
    (1024, Q27, 1 << 9, "dm_u32_ntt32"),            # three digits, top digit can wrap: plain path (STD128_AP set shape)
    (512, Q27, 1 << 9, "dm_u32_ntt32"),             # TOY shape
    (1024, Q28, 1 << 10, "dm_u32_ntt32_skiptop"),   # MEDIUM shape: elimination + mid-transform sweep
    (1024, Q28, 1 << 7, "dm_u32_ntt32"),            # SIGNED_MOD_TEST shape: four digits, plain + sweep
    (2048, 134176769, 1 << 7, "dm_u32_ntt32_skiptop"),   # STD256Q shape: N = 2048, 27-bit modulus
    (2048, 536813569, 1 << 8, "dm_u32_ntt32_skiptop"),   # STD256 shape: N = 2048, 29-bit modulus (sweeps)
])
def test_dm32_variants(N, Q, baseG, variant, rng):
    """The AP/DM kernel beyond its headline shape: plain (no top-digit elimination) and 28-bit-modulus variants, on small
    custom rings, gates / explicit accumulators / zero refresh digits, against the oracle and the generic kernel."""
    p = po.Port.params_custom(6, N, N, Q, 128, baseG, 46 if N == 2048 else 32, po.AP)
    port = po.Port(p)
    sk, bk, ksk, g = _ctx(p, port)
    try:
        assert g.kernel_variant == variant
        q, n = p.q, p.n
        c1 = rng.integers(0, q, (21, n + 1), dtype=np.uint64)              # ragged vs CTAs of 2, 4 and 8
        c2 = rng.integers(0, q, (21, n + 1), dtype=np.uint64)
        c1[1, :n] = 0
        c2[1, :n] = 0
        c1[2, :n] = 32
        c2[2, :n] = 0
        for gate in ("NAND", "XNOR_FAST"):
            want = port.eval_bin_gate(bk, ksk, po.GATES[gate], c1, c2, q)
            assert np.array_equal(g.EvalBinGate(gate, c1, c2), want), gate
        Qm, QH = p.Q, p.Q >> 1
        acc = rng.integers(0, Qm, (5, 2, N), dtype=np.uint64)
        acc[0] = np.resize(np.array([0, 1, Qm - 1, QH - 1, QH, QH + 1, QH - 64, QH + 64], dtype=np.uint64), (2, N))
        acc[1] = rng.integers(QH - 300, QH + 300, (2, N), dtype=np.uint64)     # where a wrapping top digit lives
        am = rng.integers(0, q, (5, n), dtype=np.uint64)
        want = port.eval_acc(bk, am, q, acc)
        assert np.array_equal(g.EvalAcc(am, q, acc), want)
        big = 4 * 148 + 2 * 148 - 5                                            # throughput + small-batch shapes
        b1 = rng.integers(0, q, (big, n + 1), dtype=np.uint64)
        b2 = rng.integers(0, q, (big, n + 1), dtype=np.uint64)
        assert np.array_equal(g.EvalBinGate("OR", b1, b2), port.eval_bin_gate(bk, ksk, po.GATES["OR"], b1, b2, q))
        g.set_option("force_generic", 1)
        assert np.array_equal(g.EvalAcc(am, q, acc), want)
    finally:
        g.GPUClean()


@pytest.mark.parametrize("baseG", [1 << 10, 1 << 7])                          # MEDIUM / SIGNED_MOD_TEST gadgets
def test_cggi32_28bit_modulus_sweep(baseG, rng):
    """28-bit moduli on the 32-bit CGGI kernel (mid-transform reduction sweep): extreme accumulator coefficients, gate and
    LUT accumulators, against the oracle and the generic kernel."""
    p = po.Port.params_custom(12, 1024, 1024, Q28, 128, baseG, 32, po.GINX)
    port = po.Port(p)
    sk, bk, ksk, g = _ctx(p, port)
    try:
        assert g.kernel_variant.startswith("cggi_u32_ntt32")
        q, n, N = p.q, p.n, 1024
        c1 = rng.integers(0, q, (11, n + 1), dtype=np.uint64)
        c2 = rng.integers(0, q, (11, n + 1), dtype=np.uint64)
        assert np.array_equal(g.EvalBinGate("NAND", c1, c2), port.eval_bin_gate(bk, ksk, po.GATES["NAND"], c1, c2, q))
        tab = rng.integers(0, q, (11, q), dtype=np.uint64)
        assert np.array_equal(g.BootstrapFunc(c1, q, tab, q), port.bootstrap_func(bk, ksk, c1, q, tab, q))
        Qm, QH = p.Q, p.Q >> 1
        acc = rng.integers(0, Qm, (6, 2, N), dtype=np.uint64)
        acc[0] = np.resize(np.array([0, 1, Qm - 1, QH - 1, QH, QH + 1, QH - 64, QH + 64], dtype=np.uint64), (2, N))
        acc[1] = Qm - 1
        acc[2] = QH
        am = rng.integers(0, q, (6, n), dtype=np.uint64)
        want = port.eval_acc(bk, am, q, acc)
        assert np.array_equal(g.EvalAcc(am, q, acc), want)
        g.set_option("force_generic", 1)
        assert np.array_equal(g.EvalAcc(am, q, acc), want)
    finally:
        g.GPUClean()


Q37 = 137438822401  # STD192 (37-bit, 1 mod 4096)
Q35 = 34359709697   # STD192Q (35-bit)


Q50 = 1125899906826241   # STD128Q (50-bit)


@pytest.mark.parametrize("Q,q,baseG,baseR", [(Q37, 1024, 1 << 14, 32),     # STD192 shape: three digits, two per CTA
                                             (Q29, 2048, 1 << 8, 46),      # STD256 shape: four digits, 29-bit modulus
                                             (Q35, 1024, 1 << 12, 32),     # STD192Q_OPT shape
                                             (Q50, 1024, 1 << 25, 32)])    # STD128Q shape: two digits, plain path
def test_dm64w_kernel(Q, q, baseG, baseR, rng, monkeypatch):
    """AP/DM on the N = 2048 rings (br_dm64w.cu): gates, explicit accumulators with extreme coefficients, zero refresh
    digits, both CTA shapes, against the oracle and the generic kernel.  (The 29-bit ring normally runs the 32-bit DM
    kernel, test_dm32_variants; here it is kept on the 64-bit one.)"""
    monkeypatch.setenv("TFHE_B200_NO_DM32", "1")
    p = po.Port.params_custom(5, 2048, q, Q, 64, baseG, baseR, po.AP)
    port = po.Port(p)
    sk, bk, ksk, g = _ctx(p, port)
    try:
        assert g.kernel_variant.startswith("dm_u64_ntt16x128")
        n = p.n
        c1 = rng.integers(0, q, (9, n + 1), dtype=np.uint64)                   # ragged vs CTAs of 2
        c2 = rng.integers(0, q, (9, n + 1), dtype=np.uint64)
        c1[1, :n] = 0
        c2[1, :n] = 0                                                          # every refresh digit zero
        c1[2, :n] = baseR
        c2[2, :n] = 0                                                          # low digit zero, high digit non-zero
        for gate in ("NAND", "XOR_FAST"):
            want = port.eval_bin_gate(bk, ksk, po.GATES[gate], c1, c2, q)
            assert np.array_equal(g.EvalBinGate(gate, c1, c2), want), gate
        Qm, QH = p.Q, p.Q >> 1
        acc = rng.integers(0, Qm, (5, 2, 2048), dtype=np.uint64)
        acc[0] = np.resize(np.array([0, 1, Qm - 1, QH - 1, QH, QH + 1, QH - 64, QH + 64], dtype=np.uint64), (2, 2048))
        acc[1] = Qm - 1
        acc[2] = rng.integers(QH - 2000, QH, (2, 2048), dtype=np.uint64)       # where a wrapping top digit lives
        am = rng.integers(0, q, (5, n), dtype=np.uint64)
        want = port.eval_acc(bk, am, q, acc)
        assert np.array_equal(g.EvalAcc(am, q, acc), want)
        g.set_option("group", 1)                                               # one ciphertext per CTA
        assert np.array_equal(g.EvalAcc(am, q, acc), want)
        g.set_option("group", 0)
        big = 2 * 148 + 148 - 9                                                # throughput part + latency-shaped tail
        b1 = rng.integers(0, q, (big, n + 1), dtype=np.uint64)
        b2 = rng.integers(0, q, (big, n + 1), dtype=np.uint64)
        assert np.array_equal(g.EvalBinGate("OR", b1, b2), port.eval_bin_gate(bk, ksk, po.GATES["OR"], b1, b2, q))
        g.set_option("force_generic", 1)
        assert np.array_equal(g.EvalAcc(am, q, acc), want)
    finally:
        g.GPUClean()


Q27b = 134176769    # STD256Q (27-bit, 1 mod 4096)


@pytest.mark.parametrize("Q,baseG", [(Q27b, 1 << 7), (Q29, 1 << 8)])          # STD256Q / STD256 gadgets
def test_cggi32_n2048(Q, baseG, rng, monkeypatch):
    """N = 2048 on the 32-bit CGGI kernel (64 threads x 32 coefficients per polynomial, cross-lane stage; sweeps for the
    29-bit modulus): gates, per-ciphertext LUTs, explicit accumulators with extreme coefficients, ragged batches, against
    the oracle, the generic kernel and the 64-bit kernel these rings ran on before."""
    p = po.Port.params_custom(9, 2048, 2048, Q, 128, baseG, 46, po.GINX)
    port = po.Port(p)
    sk, bk, ksk, g = _ctx(p, port)
    try:
        assert g.kernel_variant == "cggi_u32_ntt32_skiptop"
        q, n, N = p.q, p.n, 2048
        c1 = rng.integers(0, q, (7, n + 1), dtype=np.uint64)                  # ragged vs CTAs of 2
        c2 = rng.integers(0, q, (7, n + 1), dtype=np.uint64)
        c1[0, :n] = 0
        c2[0, :n] = 0
        for gate in ("NAND", "XNOR_FAST"):
            want = port.eval_bin_gate(bk, ksk, po.GATES[gate], c1, c2, q)
            assert np.array_equal(g.EvalBinGate(gate, c1, c2), want), gate
        tab = rng.integers(0, q, (7, q), dtype=np.uint64)
        assert np.array_equal(g.BootstrapFunc(c1, q, tab, q), port.bootstrap_func(bk, ksk, c1, q, tab, q))
        Qm, QH = p.Q, p.Q >> 1
        acc = rng.integers(0, Qm, (5, 2, N), dtype=np.uint64)
        acc[0] = np.resize(np.array([0, 1, Qm - 1, QH - 1, QH, QH + 1, QH - 64, QH + 64], dtype=np.uint64), (2, N))
        acc[1] = Qm - 1
        acc[2] = QH
        am = rng.integers(0, q, (5, n), dtype=np.uint64)
        want = port.eval_acc(bk, am, q, acc)
        assert np.array_equal(g.EvalAcc(am, q, acc), want)
        g.set_option("force_generic", 1)
        assert np.array_equal(g.EvalAcc(am, q, acc), want)
    finally:
        g.GPUClean()
    monkeypatch.setenv("TFHE_B200_NO_CGGI32", "1")                            # the same ring in 64-bit words
    sk, bk, ksk, g = _ctx(p, port)
    try:
        assert g.kernel_variant.startswith("cggi_u64")
        assert np.array_equal(g.EvalAcc(am, q, acc), want)
    finally:
        g.GPUClean()


@pytest.mark.parametrize("N,Q,baseG", [(1024, Q27, 1 << 9), (512, Q27, 1 << 9), (1024, Q28, 1 << 7)])
def test_cggi32_wrapping_top_digit_plain_path(N, Q, baseG, rng):
    """32-bit rings whose top digit can wrap (baseG^digits barely above Q: TOY, the STD128_AP sets, SIGNED_MOD_TEST) stay
    on the plain path (all digits transformed): the reference truncates the top digit to its window, and the kernel must
    reproduce exactly that.  Accumulators are built to hit it -- every coefficient in the wrap zone, a single wrapped
    coefficient, a mix around the zone's lower edge, the extremes of the centred range.  (Top-digit elimination with an
    on-the-fly repair, as the 64-bit kernel does, was implemented and measured slower here: DESIGN.md section 9.)"""
    p = po.Port.params_custom(10, N, N, Q, 128, baseG, 32, po.GINX)
    port = po.Port(p)
    sk, bk, ksk, g = _ctx(p, port)
    try:
        assert g.kernel_variant == "cggi_u32_ntt32"
        Qm, QH = p.Q, p.Q >> 1
        B, d = baseG, p.digitsG
        off = sum((B // 2) * B**i for i in range(d))
        lo = B**d - off                                                    # smallest centred value that wraps
        assert 0 < lo < QH
        acc = rng.integers(0, Qm, (9, 2, N), dtype=np.uint64)
        acc[0] = rng.integers(lo, QH, (2, N), dtype=np.uint64)             # all wrapped
        acc[1, 0, 77] = QH - 1                                             # exactly one wrapped coefficient
        acc[1, 1, 0] = lo
        acc[2] = rng.integers(lo - 1000, lo + 1000, (2, N), dtype=np.uint64)
        acc[3, :, ::64] = QH - 5
        edge = np.array([0, 1, Qm - 1, QH - 1, QH, QH + 1, lo - 1, lo], dtype=np.uint64)
        acc[4] = np.resize(edge, (2, N))
        am = rng.integers(0, p.q, (9, p.n), dtype=np.uint64)
        want = port.eval_acc(bk, am, p.q, acc)
        assert np.array_equal(g.EvalAcc(am, p.q, acc), want)
        g.set_option("force_generic", 1)
        assert np.array_equal(g.EvalAcc(am, p.q, acc), want)
    finally:
        g.GPUClean()


def test_top_digit_elimination_extreme_coefficients(rng):
    """Accumulator coefficients at the edges of the centred range (0, 1, Q-1, QHalf-1, QHalf, QHalf+1): where a wrapping
    top digit would break the elimination identity.  STD128-like gadget (B = 2^7, 4 digits: provably safe)."""
    p = po.Port.params_custom(12, 1024, 1024, Q27, 128, 1 << 7, 32, po.GINX)
    port = po.Port(p)
    sk, bk, ksk, g = _ctx(p, port)
    try:
        assert "skiptop" in g.kernel_variant
        Q, QH = p.Q, p.Q >> 1
        edge = np.array([0, 1, Q - 1, QH - 1, QH, QH + 1, QH - 64, QH + 64], dtype=np.uint64)
        acc = np.resize(edge, (4, 2, 1024)).copy()
        acc[1] = rng.integers(QH - 200, QH + 200, (2, 1024), dtype=np.uint64)
        am = rng.integers(0, p.q, (4, p.n), dtype=np.uint64)
        assert np.array_equal(g.EvalAcc(am, p.q, acc), port.eval_acc(bk, am, p.q, acc))
    finally:
        g.GPUClean()


Q54 = 18014398509404161


@pytest.mark.parametrize("baseG", [1 << 27, 1 << 18, 1 << 14])            # digitsG = 2, 3, 4
def test_cggi64_wrapped_top_digit_repair(baseG, rng, monkeypatch):
    """54-bit rings: for B^d = 2^54 the reference's truncated top digit wraps for centred values >= B^d/2 - B/2 - ...
    (just below Q/2), so top-digit elimination needs the kernel's wrap repair.  Accumulators are built to hit it:
    every coefficient in the wrap zone, a single wrapped coefficient, a random mix around the zone's lower edge."""
    p = po.Port.params_custom(6, 2048, 4096, Q54, 64, baseG, 32, po.GINX)
    port = po.Port(p)
    sk, bk, ksk, g = _ctx(p, port)
    try:
        assert g.kernel_variant.startswith("cggi_u64") and g.kernel_variant.endswith("skiptop")
        Q, QH, N = p.Q, p.Q >> 1, 2048
        B, d = baseG, p.digitsG
        off = sum((B // 2) * B**i for i in range(d))
        lo = max(B**d - off, 0)                                            # smallest centred value that wraps
        acc = rng.integers(0, Q, (7, 2, N), dtype=np.uint64)
        if lo < QH:
            acc[0] = rng.integers(lo, QH, (2, N), dtype=np.uint64)         # all wrapped
            acc[1, 0, 777] = QH - 1                                        # exactly one wrapped coefficient
            acc[1, 1, 0] = lo
            acc[2] = rng.integers(lo - 1000, lo + 1000, (2, N), dtype=np.uint64)
            acc[3, :, ::64] = QH - 5                                       # one warp's worth of bitmap words
        edge = np.array([0, 1, Q - 1, QH - 1, QH, QH + 1, min(max(lo, 1) - 1, Q - 1), min(lo, Q - 1)], dtype=np.uint64)
        acc[4] = np.resize(edge, (2, N))
        am = rng.integers(0, p.q, (7, p.n), dtype=np.uint64)
        want = port.eval_acc(bk, am, p.q, acc)
        assert np.array_equal(g.EvalAcc(am, p.q, acc), want)               # default: the wide kernel for 2 / 3 digits
        monkeypatch.setenv("TFHE_B200_C64_NARROW", "1")                    # 64 threads x 32 coefficients
        assert np.array_equal(g.EvalAcc(am, p.q, acc), want)
        monkeypatch.delenv("TFHE_B200_C64_NARROW")
        g.set_option("force_generic", 1)
        assert np.array_equal(g.EvalAcc(am, p.q, acc), want)
    finally:
        g.GPUClean()


def test_cggi32_tma_key_ring_variant(rng, monkeypatch):
    """Opt-in variant of the headline kernel that streams the RGSW key through TMA bulk copies into a shared-memory ring
    (full / empty mbarriers): same bits as the register-staged default and as the oracle."""
    p = po.Port.params_custom(20, 1024, 1024, Q27, 128, 1 << 7, 32, po.GINX)
    port = po.Port(p)
    sk, bk, ksk, g = _ctx(p, port)
    try:
        q, n = p.q, p.n
        c1 = rng.integers(0, q, (9, n + 1), dtype=np.uint64)              # ragged last CTA
        c2 = rng.integers(0, q, (9, n + 1), dtype=np.uint64)
        want = port.eval_bin_gate(bk, ksk, po.GATES["NAND"], c1, c2, q)
        default = g.EvalBinGate("NAND", c1, c2)
        monkeypatch.setenv("TFHE_B200_TMA", "1")
        tma = g.EvalBinGate("NAND", c1, c2)
        monkeypatch.delenv("TFHE_B200_TMA")
        assert np.array_equal(default, want)
        assert np.array_equal(tma, want)
    finally:
        g.GPUClean()


@pytest.mark.parametrize("N,baseG", [(1024, 1 << 7), (512, 1 << 9)])        # headline shape (top digit eliminated) / TOY shape
def test_cggi32_persistent_variant(N, baseG, rng):
    """Persistent blind rotation (br_cggi32_kernel<..., PERS>): the groups x n rotation steps of a launch are cut into
    equal ranges, one per CTA, so groups are split between neighbouring CTAs at arbitrary steps and handed over through
    the accumulator image.  Forced here with few CTAs on a short LWE dimension: every split position, ragged last
    group, gate / per-ciphertext LUT / explicit accumulators -- all bit-exact against the oracle and the plain launch."""
    p = po.Port.params_custom(13, N, N, Q27, 128, baseG, 32, po.GINX)
    port = po.Port(p)
    sk, bk, ksk, g = _ctx(p, port)
    try:
        assert g.kernel_variant.startswith("cggi_u32_ntt32")
        q, n = p.q, p.n
        per_cta = 4 if N == 1024 else 8
        batch = 7 * per_cta - 3                                             # seven groups, the last one ragged
        c1 = rng.integers(0, q, (batch, n + 1), dtype=np.uint64)
        c2 = rng.integers(0, q, (batch, n + 1), dtype=np.uint64)
        want = port.eval_bin_gate(bk, ksk, po.GATES["NAND"], c1, c2, q)
        tab = rng.integers(0, q, (batch, q), dtype=np.uint64)
        want_f = port.bootstrap_func(bk, ksk, c1, q, tab, q)
        acc = rng.integers(0, p.Q, (batch, 2, N), dtype=np.uint64)
        want_a = port.eval_acc(bk, c1[:, :n].copy(), q, acc)
        for ctas in (2, 3, 4, 5, 6, 7):                                     # 7 x 13 steps over `ctas` ranges
            g.set_option("persistent_ctas", ctas)
            assert np.array_equal(g.EvalBinGate("NAND", c1, c2), want), ctas
            assert np.array_equal(g.BootstrapFunc(c1, q, tab, q), want_f), ctas
            assert np.array_equal(g.EvalAcc(c1[:, :n].copy(), q, acc), want_a), ctas
        g.set_option("persistent_ctas", 0)
        g.set_option("persistent", 0)
        assert np.array_equal(g.EvalBinGate("NAND", c1, c2), want)
    finally:
        g.GPUClean()


@pytest.mark.parametrize("baseG", [1 << 27, 1 << 18])                        # digitsG = 2 (EvalFunc sets), 3 (EvalSign sets)
def test_cggi64w_persistent_variant(baseG, rng):
    """Persistent variant of the wide 54-bit kernel (two ciphertexts per CTA): groups split between neighbouring CTAs
    at arbitrary steps, wrap-zone accumulators included (the wrap bitmaps are rebuilt after a hand-over)."""
    p = po.Port.params_custom(7, 2048, 4096, Q54, 64, baseG, 32, po.GINX)
    port = po.Port(p)
    sk, bk, ksk, g = _ctx(p, port)
    try:
        assert g.kernel_variant.startswith("cggi_u64") and g.kernel_variant.endswith("skiptop")
        Q, QH, N, n, q = p.Q, p.Q >> 1, 2048, p.n, p.q
        batch = 11                                                          # six groups of two, the last one ragged
        off = sum((baseG // 2) * baseG**i for i in range(p.digitsG))
        lo = max(baseG**p.digitsG - off, 0)
        acc = rng.integers(0, Q, (batch, 2, N), dtype=np.uint64)
        if lo < QH:
            acc[0] = rng.integers(lo, QH, (2, N), dtype=np.uint64)          # every coefficient in the wrap zone
            acc[5, 1, 3] = QH - 1
        am = rng.integers(0, q, (batch, n), dtype=np.uint64)
        want_a = port.eval_acc(bk, am, q, acc)
        c1 = rng.integers(0, q, (batch, n + 1), dtype=np.uint64)
        tab = rng.integers(0, q, q, dtype=np.uint64)
        want_f = port.bootstrap_func(bk, ksk, c1, q, tab, q)
        for ctas in (2, 3, 4, 5, 6):                                        # 6 x 7 steps over `ctas` ranges
            g.set_option("persistent_ctas", ctas)
            assert np.array_equal(g.EvalAcc(am, q, acc), want_a), ctas
            assert np.array_equal(g.BootstrapFunc(c1, q, tab, q), want_f), ctas
        g.set_option("persistent_ctas", 0)
        g.set_option("persistent", 0)
        assert np.array_equal(g.EvalAcc(am, q, acc), want_a)
    finally:
        g.GPUClean()

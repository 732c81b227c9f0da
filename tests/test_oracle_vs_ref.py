"""Pins the oracle port against the UNMODIFIED reference CPU implementation compiled from /root/reference
(oracle/_ref/libtfhe_ref.so): same keys (exported from the reference), same inputs, bit-for-bit equal outputs.
Skipped where the reference build is absent (the committed golden vectors cover that case)."""
import numpy as np
import pytest

from oracle import pyoracle as po

pytestmark = pytest.mark.skipif(not po.have_ref(), reason="oracle/_ref/libtfhe_ref.so not built")


def _pair(ref):
    ref.keygen()
    port = po.Port(ref.p)
    sk, bk, ksk = ref.export_keys()
    return port, sk, bk, ksk


@pytest.mark.parametrize("method", [po.GINX, po.AP])
def test_toy_gates(method):
    ref = po.Ref.named(po.TOY, method)
    assert po.Port.params_named(po.TOY, method).as_dict() == ref.p.as_dict()
    port, sk, bk, ksk = _pair(ref)
    q = ref.p.q
    m1 = [i & 1 for i in range(4)]
    m2 = [(i >> 1) & 1 for i in range(4)]
    c1, c2 = ref.encrypt_batch(m1, 4, q), ref.encrypt_batch(m2, 4, q)
    for g in ("NAND", "OR", "XOR_FAST", "XNOR"):
        want = ref.eval_bin_gate(po.GATES[g], c1, c2, q)
        assert np.array_equal(port.eval_bin_gate(bk, ksk, po.GATES[g], c1, c2, q), want), g
    # our own encryption decrypts under the reference and vice versa
    ct = port.encrypt_batch(sk, [0, 1, 2, 3], 4, q, 9)
    assert ref.decrypt_batch(ct, q, 4) == [0, 1, 2, 3]
    assert port.decrypt_batch(sk, c1, q, 4) == m1


def test_std128_ginx_nand():
    ref = po.Ref.named(po.STD128, po.GINX)
    assert po.Port.params_named(po.STD128, po.GINX).as_dict() == ref.p.as_dict()
    port, sk, bk, ksk = _pair(ref)
    q = ref.p.q
    c1, c2 = ref.encrypt_batch([0, 1, 0, 1], 4, q), ref.encrypt_batch([0, 0, 1, 1], 4, q)
    want = ref.eval_bin_gate(po.GATES["NAND"], c1, c2, q)
    assert np.array_equal(port.eval_bin_gate(bk, ksk, po.GATES["NAND"], c1, c2, q), want)
    assert ref.decrypt_batch(want, q, 4) == [1, 1, 1, 0]


def test_functional_chain_54bit():
    ref = po.Ref.func(po.TOY, True, 12)
    assert po.Port.params_func(po.TOY, True, 12).as_dict() == ref.p.as_dict()
    port, sk, bk, ksk = _pair(ref)
    q = ref.p.q
    p = q // (2 * ref.p.beta)
    lut = np.array([((x // (q // p)) ** 3 % p) * (q // p) for x in range(q)], dtype=np.uint64)
    ct = ref.encrypt_batch(list(range(p)), p, q)
    want = ref.eval_func(ct, q, lut)
    assert np.array_equal(port.eval_func(bk, ksk, ct, q, lut), want)
    assert ref.decrypt_batch(want, q, p) == [m ** 3 % p for m in range(p)]


def test_sign_and_decomp_logq17():
    ref = po.Ref.func(po.TOY, False, 17)
    port, sk, bk, ksk = _pair(ref)
    Qin, q = 1 << 17, ref.p.q
    P = Qin // q * (q // (2 * ref.p.beta))
    msgs = [P // 2 + i - 2 for i in range(4)]
    ct = ref.encrypt_batch(msgs, P, Qin)
    assert np.array_equal(port.eval_floor(bk, ksk, ct, Qin), ref.eval_floor(ct, Qin))
    want = ref.eval_sign(ct, Qin)
    assert np.array_equal(port.eval_sign(bk, ksk, ct, Qin), want)
    assert ref.decrypt_batch(want, q, 2) == [int(m >= P // 2) for m in msgs]
    w, wm = ref.eval_decomp(ct, Qin)
    g, gm = port.eval_decomp(bk, ksk, ct, Qin)
    assert gm == wm and np.array_equal(g, w)


def test_dynamic_gadget_base_sign_and_decomp_logq29():
    """timeOptimization: BTKeyGen fills the three-key map (bases 2^14, 2^18, 2^27; binfhecontext.cpp:222-247) and the
    scalar EvalSign / EvalDecomp switch key set as the modulus shrinks 2^29 -> 2^25 -> 2^21 -> 2^17 -> 2^13 -> 2^9
    (binfhe-base-scheme.cpp:342-360, 411-428): logQ = 29 walks through all three."""
    ref = po.Ref.func_dynamic(po.TOY, False, 29)
    ref.keygen()
    assert sorted(ref.key_map_bases()) == [1 << 14, 1 << 18, 1 << 27] and ref.p.baseG == 1 << 14
    km = ref.export_key_map()
    assert {b: km[b][0].digitsG for b in km} == {1 << 14: 4, 1 << 18: 3, 1 << 27: 2}
    order = [ref.p.baseG] + [b for b in sorted(km) if b != ref.p.baseG]   # the context's own set first
    ports = [po.Port(km[b][0]) for b in order]
    bks, ksks = [km[b][1] for b in order], [km[b][2] for b in order]
    Qin, q = 1 << 29, ref.p.q
    P = Qin // q * (q // (2 * ref.p.beta))
    msgs = [P // 2 + i - 2 for i in range(4)] + [3, P - 5]
    ct = ref.encrypt_batch(msgs, P, Qin)
    want = ref.eval_sign(ct, Qin)
    assert ref.decrypt_batch(want, q, 2) == [int(m >= P // 2) for m in msgs]
    assert np.array_equal(po.Port.eval_sign_dyn(ports, bks, ksks, ct, Qin), want)
    # the single-key evaluation is a different computation (other keys, other digit counts)
    assert not np.array_equal(ports[0].eval_sign(bks[0], ksks[0], ct, Qin), want)
    w, wm = ref.eval_decomp(ct, Qin)
    g, gm = po.Port.eval_decomp_dyn(ports, bks, ksks, ct, Qin)
    assert gm == wm and np.array_equal(g, w)

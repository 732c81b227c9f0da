"""SURVEY 8(f) rank 4 -- dynamic gadget base ("timeOptimization") for EvalSign / EvalDecomp.

The reference's GPU path refuses such contexts (binfhecontext.cpp:350-353); its scalar CPU path switches between the
three key sets of m_BTKey_map as the ciphertext modulus shrinks (binfhe-base-scheme.cpp:342-360, 411-428).  Here the
keys come from the REFERENCE's own BTKeyGen (oracle/_ref/libtfhe_ref.so), all three sets are loaded into our engine
(tfhe_b200_add_key_set) and the batched EvalSign / EvalDecomp must equal the reference's scalar results bit for bit --
and the oracle port's restatement of the rule (pinned against the reference in tests/test_oracle_vs_ref.py)."""
import numpy as np
import pytest

from oracle import pyoracle as po

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not po.have_ref(), reason="oracle/_ref/libtfhe_ref.so not built")]


@pytest.fixture(scope="module")
def dyn():
    from tfhe_gpu_b200 import BinFHEContextB200

    ref = po.Ref.func_dynamic(po.TOY, False, 29)
    ref.keygen()
    km = ref.export_key_map()
    own = int(ref.p.baseG)
    g = BinFHEContextB200().GPUSetup(km[own][0].as_dict(), km[own][1], km[own][2], numGPUs=1)
    yield ref, km, own, g
    g.GPUClean()


def _inputs(ref, logQ, count):
    Qin, q = 1 << logQ, ref.p.q
    P = Qin // q * (q // (2 * ref.p.beta))
    msgs = [P // 2 + i - count // 2 for i in range(count - 2)] + [1, P - 2]
    return Qin, P, msgs, ref.encrypt_batch(msgs, P, Qin)


def test_dynamic_sign_decomp_match_reference_scalar(dyn):
    from tfhe_gpu_b200 import TfheB200Error

    ref, km, own, g = dyn
    Qin, P, msgs, ct = _inputs(ref, 29, 6)
    want_sign = ref.eval_sign(ct, Qin)                      # reference scalar path over the three-key map
    want_dig, want_mods = ref.eval_decomp(ct, Qin)
    assert ref.decrypt_batch(want_sign, ref.p.q, 2) == [int(m >= P // 2) for m in msgs]

    # one key set loaded: the single-key semantics of the batched reference API (binfhe-base-scheme.cpp:989-1085)
    assert g.num_key_sets == 1
    single = g.EvalSign(ct, Qin)
    assert not np.array_equal(single, want_sign)
    port0 = po.Port(km[own][0])
    assert np.array_equal(single, port0.eval_sign(km[own][1], km[own][2], ct, Qin))

    # two sets: still no switching ("if (EKs.size() == 3)")
    others = [b for b in sorted(km) if b != own]
    g.AddKeySet(others[0], km[others[0]][1], km[others[0]][2])
    assert g.num_key_sets == 2
    assert np.array_equal(g.EvalSign(ct, Qin), single)
    with pytest.raises(TfheB200Error, match="already loaded"):
        g.AddKeySet(others[0], km[others[0]][1], km[others[0]][2])
    with pytest.raises(TfheB200Error, match="key sizes"):
        g.AddKeySet(others[1], km[others[0]][1], km[others[0]][2])

    # all three: the dynamic rule, bit-exact against the reference's scalar EvalSign / EvalDecomp
    g.AddKeySet(others[1], km[others[1]][1], km[others[1]][2])
    assert g.num_key_sets == 3
    got = g.EvalSign(ct, Qin)
    assert np.array_equal(got, want_sign)
    assert g.last_stats.bootstraps == 11                    # five floors + the final sign bootstrap
    gd, gm = g.EvalDecomp(ct, Qin)
    assert gm == want_mods and np.array_equal(gd, want_dig)
    # digits recompose the message
    beta, q = ref.p.beta, ref.p.q
    for j, m in enumerate(msgs):
        total, scale = 0, 1
        for k, mod in enumerate(gm):
            pk = int(mod) // (2 * beta)
            total += ref.decrypt(np.ascontiguousarray(gd[j, k, :]), int(mod), pk) * scale
            scale *= pk
        assert total % P == m
    # specialised kernels and the generic kernel agree on the whole dynamic chain
    g.set_option("force_generic", 1)
    try:
        assert np.array_equal(g.EvalSign(ct, Qin), want_sign)
    finally:
        g.set_option("force_generic", 0)
    # every other entry point keeps the context's own key set
    assert np.array_equal(g.EvalFloor(ct, Qin), ref.eval_floor(ct, Qin))


def test_dynamic_smaller_modulus_and_port(dyn):
    """logQ-29 context, 2^21 ciphertexts: the first floor already runs under the context's 2^14 base, then 2^27 only
    (2^21 -> 2^17 -> 2^13 -> 2^9); checked against the reference and the oracle port's tfo_eval_sign_dyn."""
    ref, km, own, g = dyn
    if g.num_key_sets != 3:
        pytest.skip("needs the three-key handle of the previous test")
    Qin, P, msgs, ct = _inputs(ref, 21, 5)
    want = ref.eval_sign(ct, Qin)
    order = [own] + [b for b in sorted(km) if b != own]
    ports = [po.Port(km[b][0]) for b in order]
    bks, ksks = [km[b][1] for b in order], [km[b][2] for b in order]
    assert np.array_equal(po.Port.eval_sign_dyn(ports, bks, ksks, ct, Qin), want)
    assert np.array_equal(g.EvalSign(ct, Qin), want)
    wd, wm = ref.eval_decomp(ct, Qin)
    gd, gm = g.EvalDecomp(ct, Qin)
    assert gm == wm and np.array_equal(gd, wd)

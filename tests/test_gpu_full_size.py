"""BASELINE.json's FULL batch sizes, checked through size-independent properties (the oracle would need minutes for
them): encrypt -> operator on the GPU -> decrypt round trips over the whole batch, a bit-exact oracle comparison on a
scattered slice, and agreement of sharding-independent paths (host buffers, pipelined, vs device tensors).
  configs[1]  STD128 CGGI, 16384 ciphertext pairs, NAND / AND / XOR
  configs[2]  STD128 AP (DM), 16384 pairs, NAND
  configs[3]  logQ = 12 EvalFunc, arbitrary LUT, 8192 inputs
  configs[4]  logQ = 17 EvalSign + EvalDecomp, 4096 inputs"""
import numpy as np
import pytest

from oracle import pyoracle as po

pytestmark = [pytest.mark.gpu, pytest.mark.slow]


def _bits(batch, salt):
    x = (np.arange(batch, dtype=np.int64) * 2654435761 + salt) >> 7
    return [int(v) & 1 for v in x]


def _slice(batch):
    return np.r_[0:4, batch // 2 - 2:batch // 2 + 2, batch - 4:batch]


@pytest.mark.parametrize("name,batch", [("std128_ginx", 16384), ("std128_ap", 16384)])
def test_gates_full_batch_round_trip(keyset, name, batch):
    import torch

    ks = keyset(name)
    q = ks.p.q
    m1, m2 = _bits(batch, 1), _bits(batch, 99)
    c1 = ks.port.encrypt_batch(ks.sk, m1, 4, q, 31)
    c2 = ks.port.encrypt_batch(ks.sk, m2, 4, q, 32)
    g = ks.gpu()
    gates = {"NAND": lambda a, b: 1 - (a & b)}
    if name == "std128_ginx":
        gates.update({"AND": lambda a, b: a & b, "XOR": lambda a, b: a ^ b})
    idx = _slice(batch)
    for gate, f in gates.items():
        got = g.EvalBinGate(gate, c1, c2)                                   # host buffers: pipelined chunks
        assert ks.port.decrypt_batch(ks.sk, got, q, 4) == [f(a, b) for a, b in zip(m1, m2)], gate
        want = ks.port.eval_bin_gate(ks.bk, ks.ksk, po.GATES[gate], c1[idx], c2[idx], q)
        assert np.array_equal(got[idx], want), gate
        if gate == "NAND":                                                   # device-resident path: same bits
            d = g.EvalBinGate(gate, torch.from_numpy(c1.astype(np.int64)).cuda(), torch.from_numpy(c2.astype(np.int64)).cuda())
            assert np.array_equal(d.cpu().numpy().astype(np.uint64), got)


def test_evalfunc_full_batch_round_trip(keyset):
    ks = keyset("std128_func12")
    q, batch = ks.p.q, 8192
    p = q // (2 * ks.p.beta)
    lut = np.array([((x // (q // p)) ** 3 % p) * (q // p) for x in range(q)], dtype=np.uint64)
    msgs = [(7 * i + 3) % p for i in range(batch)]
    ct = ks.port.encrypt_batch(ks.sk, msgs, p, q, 41)
    got = ks.gpu().EvalFunc(ct, lut)
    assert ks.port.decrypt_batch(ks.sk, got, q, p) == [m ** 3 % p for m in msgs]
    idx = _slice(batch)[:4]
    assert np.array_equal(got[idx], ks.port.eval_func(ks.bk, ks.ksk, ct[idx], q, lut))


def test_sign_decomp_full_batch_round_trip(keyset):
    ks = keyset("std128_sign17")
    Qin, q, batch = 1 << 17, ks.p.q, 4096
    P = Qin // q * (q // (2 * ks.p.beta))
    msgs = [(P // 2 + (i % 7) - 3) % P for i in range(batch)]
    ct = ks.port.encrypt_batch(ks.sk, msgs, P, Qin, 51)
    g = ks.gpu()
    sign = g.EvalSign(ct, Qin)
    assert ks.port.decrypt_batch(ks.sk, sign, q, 2) == [int(m >= P // 2) for m in msgs]
    digs, mods = g.EvalDecomp(ct, Qin)
    # the digits recompose the message: sum_k digit_k * prod(previous plaintext moduli)
    base = q // (2 * ks.p.beta)
    total = [0] * batch
    scale = 1
    for k, mod in enumerate(mods):
        pk = base if mod == q else max(2, int(mod) // (2 * ks.p.beta))
        vals = ks.port.decrypt_batch(ks.sk, np.ascontiguousarray(digs[:, k, :]), int(mod), pk)
        for j, v in enumerate(vals):
            total[j] += v * scale
        scale *= pk
    assert [t % P for t in total] == msgs
    idx = _slice(batch)[:3]
    assert np.array_equal(sign[idx], ks.port.eval_sign(ks.bk, ks.ksk, ct[idx], Qin))

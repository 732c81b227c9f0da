"""GPU parity: batched EvalBinGate through the C ABI vs the oracle, bit for bit.

Mirrors the truth-table cases of the reference's UnitTestFHEW.cpp (AND/OR/NAND/NOR/XOR/XNOR/XOR_FAST/XNOR_FAST,
AP and GINX) but through the BATCHED API (which the reference never tests) and at the bit level.
"""
import numpy as np
import pytest

from oracle import pyoracle as po

pytestmark = pytest.mark.gpu

TRUTH = {
    "AND": lambda a, b: a & b, "OR": lambda a, b: a | b, "NAND": lambda a, b: 1 - (a & b),
    "NOR": lambda a, b: 1 - (a | b), "XOR": lambda a, b: a ^ b, "XNOR": lambda a, b: 1 - (a ^ b),
    "XOR_FAST": lambda a, b: a ^ b, "XNOR_FAST": lambda a, b: 1 - (a ^ b),
}


def _inputs(ks, batch, seed=7):
    q = ks.p.q
    m1 = [(i >> 0) & 1 for i in range(batch)]
    m2 = [(i >> 1) & 1 for i in range(batch)]
    c1 = ks.port.encrypt_batch(ks.sk, m1, 4, q, seed)
    c2 = ks.port.encrypt_batch(ks.sk, m2, 4, q, seed + 1)
    return m1, m2, c1, c2


@pytest.mark.parametrize("name", ["toy_ginx", "toy_ap"])
@pytest.mark.parametrize("gate", list(TRUTH))
def test_toy_gates_bit_exact(keyset, name, gate):
    ks = keyset(name)
    m1, m2, c1, c2 = _inputs(ks, 16)
    want = ks.port.eval_bin_gate(ks.bk, ks.ksk, po.GATES[gate], c1, c2, ks.p.q)
    got = ks.gpu().EvalBinGate(gate, c1, c2)
    assert np.array_equal(got, want)
    dec = ks.port.decrypt_batch(ks.sk, got, ks.p.q, 4)
    assert dec == [TRUTH[gate](a, b) for a, b in zip(m1, m2)]


def test_toy_ginx_1024_pairs(keyset):
    """BASELINE.json configs[0]: TOY CGGI NAND on 1024 ciphertext pairs, bit-exact vs the oracle."""
    ks = keyset("toy_ginx")
    m1, m2, c1, c2 = _inputs(ks, 1024)
    want = ks.port.eval_bin_gate(ks.bk, ks.ksk, po.GATES["NAND"], c1, c2, ks.p.q)
    got = ks.gpu().EvalBinGate("NAND", c1, c2)
    assert np.array_equal(got, want)
    assert ks.gpu().last_stats.bootstraps == 1


def test_random_ciphertexts_are_data_oblivious(keyset, rng):
    """Uniformly random (a, b) -- the throughput workload -- also matches, including a_i == 0 masks."""
    ks = keyset("toy_ginx")
    q, n = ks.p.q, ks.p.n
    c1 = rng.integers(0, q, (64, n + 1), dtype=np.uint64)
    c2 = rng.integers(0, q, (64, n + 1), dtype=np.uint64)
    c1[0, :n] = 0
    c2[0, :n] = 0          # every rotation exponent is zero: blind rotation is the identity
    c1[1, :n] = q - 1
    for gate in ("NAND", "XOR_FAST"):
        want = ks.port.eval_bin_gate(ks.bk, ks.ksk, po.GATES[gate], c1, c2, q)
        got = ks.gpu().EvalBinGate(gate, c1, c2)
        assert np.array_equal(got, want), gate


@pytest.mark.parametrize("gate", ["NAND", "AND", "XOR"])
def test_std128_ginx_bit_exact(keyset, gate):
    """BASELINE.json configs[1] at an oracle-sized batch."""
    ks = keyset("std128_ginx")
    m1, m2, c1, c2 = _inputs(ks, 16)
    want = ks.port.eval_bin_gate(ks.bk, ks.ksk, po.GATES[gate], c1, c2, ks.p.q)
    got = ks.gpu().EvalBinGate(gate, c1, c2)
    assert np.array_equal(got, want)
    assert ks.port.decrypt_batch(ks.sk, got, ks.p.q, 4) == [TRUTH[gate](a, b) for a, b in zip(m1, m2)]


def test_std128_generic_and_specialised_kernels_agree(keyset, rng):
    ks = keyset("std128_ginx")
    q, n = ks.p.q, ks.p.n
    c1 = rng.integers(0, q, (37, n + 1), dtype=np.uint64)   # ragged: not a multiple of the CTA group size
    c2 = rng.integers(0, q, (37, n + 1), dtype=np.uint64)
    g = ks.gpu()
    a = g.EvalBinGate("NAND", c1, c2)
    g.set_option("force_generic", 1)
    try:
        b = g.EvalBinGate("NAND", c1, c2)
    finally:
        g.set_option("force_generic", 0)
    assert np.array_equal(a, b)
    want = ks.port.eval_bin_gate(ks.bk, ks.ksk, po.GATES["NAND"], c1[:4], c2[:4], q)
    assert np.array_equal(a[:4], want)


def test_operator_level_entry_points(keyset, rng):
    """EvalAcc_CUDA / MKMSwitch_CUDA contracts (bootstrapping.cuh:111-136) against the oracle stages."""
    ks = keyset("toy_ginx")
    p = ks.p
    batch = 8
    a = rng.integers(0, p.q, (batch, p.n), dtype=np.uint64)
    acc = np.zeros((batch, 2, p.N), dtype=np.uint64)
    for s in range(batch):
        acc[s] = ks.port.init_acc_gate(po.GATES["NAND"], int(rng.integers(0, p.q)), p.q)
    want = ks.port.eval_acc(ks.bk, a, p.q, acc)
    got = ks.gpu().EvalAcc(a, p.q, acc)
    assert np.array_equal(got, want)
    ext = rng.integers(0, p.Q, (batch, p.N + 1), dtype=np.uint64)
    want = ks.port.mkmswitch(ks.ksk, ext, p.q)
    got = ks.gpu().MKMSwitch(ext, p.q)
    assert np.array_equal(got, want)


def test_device_resident_tensors(keyset):
    import torch

    ks = keyset("toy_ginx")
    m1, m2, c1, c2 = _inputs(ks, 32)
    want = ks.port.eval_bin_gate(ks.bk, ks.ksk, po.GATES["NAND"], c1, c2, ks.p.q)
    d1 = torch.from_numpy(c1.view(np.int64)).cuda()
    d2 = torch.from_numpy(c2.view(np.int64)).cuda()
    out = ks.gpu().EvalBinGate("NAND", d1, d2)
    assert out.is_cuda
    assert np.array_equal(out.cpu().numpy().view(np.uint64), want)


def test_error_behaviour(keyset):
    from tfhe_gpu_b200 import TfheB200Error

    ks = keyset("toy_ginx")
    g = ks.gpu()
    n = ks.p.n
    empty = np.zeros((0, n + 1), dtype=np.uint64)
    with pytest.raises(TfheB200Error, match="input vector is empty"):
        g.EvalBinGate("NAND", empty, empty)
    with pytest.raises(TfheB200Error, match="size unmatched"):
        g.EvalBinGate("NAND", np.zeros((2, n + 1), dtype=np.uint64), np.zeros((3, n + 1), dtype=np.uint64))
    x = np.zeros((2, n + 1), dtype=np.uint64)
    with pytest.raises(TfheB200Error, match="independant"):
        g.EvalBinGate("NAND", x, x)


def test_host_buffer_pipeline_matches_single_shot(keyset, rng, monkeypatch):
    """Host-buffer calls with >= 4096 ciphertexts go through in chunks (upload / bootstrap / download overlapped on three
    streams); the result must equal the one-shot path and the oracle, for a ragged chunking and a composite gate."""
    ks = keyset("toy_ginx")
    q, n = ks.p.q, ks.p.n
    batch = 2 * 4096 + 37
    c1 = rng.integers(0, q, (batch, n + 1), dtype=np.uint64)
    c2 = rng.integers(0, q, (batch, n + 1), dtype=np.uint64)
    g = ks.gpu()
    for gate in ("NAND", "XOR"):
        piped = g.EvalBinGate(gate, c1, c2)
        monkeypatch.setenv("TFHE_B200_NO_PIPELINE", "1")
        single = g.EvalBinGate(gate, c1, c2)
        monkeypatch.delenv("TFHE_B200_NO_PIPELINE")
        assert np.array_equal(piped, single), gate
        idx = np.r_[0:3, 4094:4099, batch - 3:batch]
        want = ks.port.eval_bin_gate(ks.bk, ks.ksk, po.GATES[gate], c1[idx], c2[idx], q)
        assert np.array_equal(piped[idx], want), gate


def test_std128_cta_shapes_agree(keyset, rng):
    """The three CTA shapes of the STD128 CGGI kernel -- 4 ciphertexts per CTA (throughput), 2 per CTA (batches that
    cannot fill the SMs) and the latency layout (one ciphertext per CTA, one warp per digit polynomial plus two
    pointwise helper warps; picked automatically for batches <= one ciphertext per SM) -- produce the same bits, equal
    to the oracle, on ragged batches with extreme mask values."""
    ks = keyset("std128_ginx")
    q, n = ks.p.q, ks.p.n
    c1 = rng.integers(0, q, (13, n + 1), dtype=np.uint64)
    c2 = rng.integers(0, q, (13, n + 1), dtype=np.uint64)
    c1[0, :n], c2[0, :n] = 0, 0                       # every rotation exponent zero
    c1[1, :n], c2[1, :n] = q - 1, 0                   # exponent 2N/q everywhere
    c1[2, :n], c2[2, :n] = q // 2, 0                  # exponent N: X^N = -1
    g = ks.gpu()
    outs = {}
    try:
        for grp in (0, 1, 2, 4):
            g.set_option("group", grp)
            outs[grp] = {gate: g.EvalBinGate(gate, c1, c2) for gate in ("NAND", "XNOR")}
    finally:
        g.set_option("group", 0)
    for grp in (1, 2, 4):
        for gate in ("NAND", "XNOR"):
            assert np.array_equal(outs[grp][gate], outs[0][gate]), (grp, gate)
    want = ks.port.eval_bin_gate(ks.bk, ks.ksk, po.GATES["NAND"], c1[:4], c2[:4], q)
    assert np.array_equal(outs[1]["NAND"][:4], want)
    # a single ciphertext, and a batch just above one-per-SM (falls back to 2 per CTA)
    one = g.EvalBinGate("NAND", c1[:1], c2[:1])
    assert np.array_equal(one, outs[0]["NAND"][:1])
    big1 = np.tile(c1, (12, 1))[:150]
    big2 = np.tile(c2, (12, 1))[:150]
    assert np.array_equal(g.EvalBinGate("NAND", big1, big2), np.tile(outs[0]["NAND"], (12, 1))[:150])

"""Gate-graph submission (tfhe_b200_eval_circuit): a netlist evaluated in one call with
device-resident intermediates must equal, bit for bit, the same nodes evaluated one by one with the oracle's
EvalBinGate (binfhe-base-scheme.cpp:598-677) and EvalNOT (:741-745)."""
import numpy as np
import pytest

from oracle import pyoracle as po

pytestmark = pytest.mark.gpu


def _not(ct, q):
    out = (q - ct) % q                      # a -> -a
    out[:, -1] = (q // 4 + q - ct[:, -1]) % q   # b -> q/4 - b
    return out.astype(np.uint64)


def _oracle_netlist(ks, inputs, nodes, q):
    wires = [np.ascontiguousarray(x) for x in inputs]
    for g, a, b in nodes:
        if g == "NOT":
            wires.append(_not(wires[a], q))
        else:
            wires.append(ks.port.eval_bin_gate(ks.bk, ks.ksk, po.GATES[g], wires[a], wires[b], q))
    return wires


def _adder_netlist(bits):
    """Ripple-carry adder over inputs a[0..bits), b[0..bits): returns (nodes, sum wires, carry wire)."""
    nodes, wire = [], 2 * bits

    def add(g, x, y=None):
        nonlocal wire
        nodes.append((g, x, y))
        wire += 1
        return wire - 1

    sums, carry = [], None
    for i in range(bits):
        a, b = i, bits + i
        x = add("XOR", a, b)
        if carry is None:
            sums.append(x)
            carry = add("AND", a, b)
        else:
            sums.append(add("XOR", x, carry))
            t1 = add("NAND", a, b)
            t2 = add("NAND", x, carry)
            carry = add("NAND", t1, t2)          # (a & b) | (x & carry)
    return nodes, sums, carry


def test_adder_netlist_matches_gate_by_gate_oracle(keyset):
    ks = keyset("toy_ginx")
    q, bits, batch = ks.p.q, 3, 5
    vals_a, vals_b = [1, 5, 7, 2, 6], [3, 6, 7, 0, 1]
    inputs = []
    seed = 100
    for src in (vals_a, vals_b):
        for i in range(bits):
            inputs.append(ks.port.encrypt_batch(ks.sk, [(v >> i) & 1 for v in src], 4, q, seed))
            seed += 1
    inputs = np.stack(inputs)
    nodes, sums, carry = _adder_netlist(bits)
    outs = sums + [carry]
    got = ks.gpu().EvalCircuit(inputs, nodes, outs)
    wires = _oracle_netlist(ks, inputs, nodes, q)
    for k, w in enumerate(outs):
        assert np.array_equal(got[k], wires[w]), f"output {k} (wire {w})"
    total = [0] * batch
    for k in range(bits + 1):
        for j, bit in enumerate(ks.port.decrypt_batch(ks.sk, got[k], q, 4)):
            total[j] |= bit << k
    assert total == [a + b for a, b in zip(vals_a, vals_b)]
    st = ks.gpu().last_stats
    assert st.bootstraps == sum(3 if g == "XOR" else 1 for g, _, _ in nodes)


def test_every_gate_kind_and_not_chains(keyset, rng):
    ks = keyset("toy_ginx")
    q, batch = ks.p.q, 4
    ins = np.stack([ks.port.encrypt_batch(ks.sk, [int(x) for x in rng.integers(0, 2, batch)], 4, q, 200 + i)
                    for i in range(3)])
    nodes = [("OR", 0, 1), ("AND", 1, 2), ("NOR", 0, 2), ("NAND", 3, 4), ("XOR_FAST", 5, 6), ("XNOR_FAST", 0, 7),
             ("XOR", 8, 1), ("XNOR", 2, 9), ("NOT", 10, None), ("NOT", 11, None), ("NOT", 0, None), ("AND", 12, 13),
             ("OR", 14, 2)]
    outs = list(range(3, 3 + len(nodes))) + [0]          # every node output plus a pass-through input
    got = ks.gpu().EvalCircuit(ins, nodes, outs)
    wires = _oracle_netlist(ks, ins, nodes, q)
    for k, w in enumerate(outs):
        assert np.array_equal(got[k], wires[w]), f"wire {w}"


def test_wide_level_is_one_launch_per_gate_kind(keyset, rng):
    """16 independent NANDs of one level: one blind rotation over 16 x batch ciphertexts, results identical to 16
    EvalBinGate calls."""
    ks = keyset("toy_ginx")
    q, batch = ks.p.q, 3
    ins = np.stack([rng.integers(0, q, (batch, ks.p.n + 1), dtype=np.uint64) for _ in range(17)])
    nodes = [("NAND", i, i + 1) for i in range(16)]
    got = ks.gpu().EvalCircuit(ins, nodes, list(range(17, 33)))
    for i in range(16):
        want = ks.port.eval_bin_gate(ks.bk, ks.ksk, po.GATES["NAND"], ins[i], ins[i + 1], q)
        assert np.array_equal(got[i], want), i
    # 16 affine launches + ONE blind rotation + ONE key switch
    assert ks.gpu().last_stats.kernel_launches == 18


def test_circuit_errors(keyset):
    from tfhe_gpu_b200 import TfheB200Error

    ks = keyset("toy_ginx")
    g = ks.gpu()
    ins = np.zeros((2, 2, ks.p.n + 1), dtype=np.uint64)
    with pytest.raises(TfheB200Error, match="independant"):
        g.EvalCircuit(ins, [("AND", 0, 0)], [2])
    with pytest.raises(TfheB200Error, match="earlier wire"):
        g.EvalCircuit(ins, [("AND", 0, 2)], [2])
    with pytest.raises(TfheB200Error, match="out of range"):
        g.EvalCircuit(ins, [("AND", 0, 1)], [3])
    with pytest.raises(TfheB200Error, match="empty"):
        g.EvalCircuit(np.zeros((0, 2, ks.p.n + 1), dtype=np.uint64), [], [0])


def test_deep_chain_like_the_reference_long_running_tests(keyset, rng):
    """UnitTestFHEWDeep.cpp:42-316 chains thousands of NOT / AND / OR / XOR on one ciphertext and asserts every decrypt.
    Here: a 600-node dependent chain over a batch, submitted as ONE netlist, every wire returned and decrypted."""
    ks = keyset("toy_ginx")
    q, batch = ks.p.q, 6
    bits = [[int(x) for x in rng.integers(0, 2, batch)] for _ in range(3)]
    ins = np.stack([ks.port.encrypt_batch(ks.sk, b, 4, q, 300 + i) for i, b in enumerate(bits)])
    kinds = ["NAND", "NOT", "OR", "XOR_FAST", "AND", "NOT", "XOR", "NOR", "XNOR_FAST"]
    fn = {"NAND": lambda a, b: 1 - (a & b), "OR": lambda a, b: a | b, "XOR_FAST": lambda a, b: a ^ b,
          "AND": lambda a, b: a & b, "XOR": lambda a, b: a ^ b, "NOR": lambda a, b: 1 - (a | b),
          "XNOR_FAST": lambda a, b: 1 - (a ^ b)}
    nodes, plain = [], [list(b) for b in bits]
    cur = 0
    for k in range(600):
        g = kinds[k % len(kinds)]
        other = 1 + (k % 2)                                   # alternate the two side inputs
        if g == "NOT":
            nodes.append(("NOT", cur, None))
            plain.append([1 - v for v in plain[cur]])
        else:
            nodes.append((g, cur, other))
            plain.append([fn[g](a, b) for a, b in zip(plain[cur], plain[other])])
        cur = 3 + k
    outs = list(range(3, 3 + len(nodes)))
    got = ks.gpu().EvalCircuit(ins, nodes, outs)
    for k, w in enumerate(outs):
        assert ks.port.decrypt_batch(ks.sk, got[k], q, 4) == plain[w], f"node {k} ({nodes[k][0]})"

"""Host-side orchestration of the C ABI on one GPU: pinned staging of pageable buffers, the chunked three-stream
pipeline of EvalBinGate, the split of a launch into a throughput part and a latency-shaped tail (wave quantisation), the
release of the generic key layout, and the per-call statistics.  Small custom rings (LWE dimension 12) keep the oracle
instant, so whole batches are compared bit for bit (reference semantics: binfhe-base-scheme.cpp:598-677;
staging: bootstrapping.cu:904-905, 1562-1853)."""
import numpy as np
import pytest

from oracle import pyoracle as po

pytestmark = pytest.mark.gpu

Q27 = 134215681
Q54 = 18014398509404161


def _ctx(p, port, seed=5, **kw):
    from tfhe_gpu_b200 import BinFHEContextB200

    sk, bk, ksk = port.keygen(seed)
    return sk, bk, ksk, BinFHEContextB200().GPUSetup(p.as_dict(), bk, ksk, numGPUs=1, **kw)


def _sm_count():
    import torch

    return torch.cuda.get_device_properties(0).multi_processor_count


def test_pageable_pinned_and_device_buffers_agree(rng):
    """A batch of several CTA waves goes through the chunked pipeline from pageable numpy buffers (staged through the
    handle's pinned memory), from pinned buffers (direct DMA) and from device tensors: same bits, equal to the oracle."""
    import torch

    p = po.Port.params_custom(12, 1024, 1024, Q27, 128, 1 << 7, 32, po.GINX)
    port = po.Port(p)
    sk, bk, ksk, g = _ctx(p, port)
    try:
        q, n = p.q, p.n
        batch = 4 * _sm_count() * 4 + 301                    # 4 full waves of 4-ciphertext CTAs + a ragged tail
        c1 = rng.integers(0, q, (batch, n + 1), dtype=np.uint64)
        c2 = rng.integers(0, q, (batch, n + 1), dtype=np.uint64)
        want = port.eval_bin_gate(bk, ksk, po.GATES["NAND"], c1, c2, q)
        got = g.EvalBinGate("NAND", c1, c2)
        assert np.array_equal(got, want)
        again = g.EvalBinGate("NAND", c1, c2)                # second call re-uses the staging buffers
        assert np.array_equal(again, want)
        p1 = torch.from_numpy(c1.view(np.int64)).pin_memory()
        p2 = torch.from_numpy(c2.view(np.int64)).pin_memory()
        pout = torch.empty_like(p1).pin_memory()
        g.EvalBinGate("NAND", p1.numpy().view(np.uint64), p2.numpy().view(np.uint64), out=pout.numpy().view(np.uint64))
        assert np.array_equal(pout.numpy().view(np.uint64), want)
        # mixed: pinned inputs, pageable output
        mixed = g.EvalBinGate("NAND", p1.numpy().view(np.uint64), p2.numpy().view(np.uint64))
        assert np.array_equal(mixed, want)
        dout = g.EvalBinGate("NAND", p1.cuda(), p2.cuda())
        assert np.array_equal(dout.cpu().numpy().view(np.uint64), want)
        # three bootstraps per gate through the same pipeline
        xs = slice(0, 3 * _sm_count() * 4 + 5)
        assert np.array_equal(g.EvalBinGate("XOR", c1[xs], c2[xs]),
                              port.eval_bin_gate(bk, ksk, po.GATES["XOR"], c1[xs], c2[xs], q))
    finally:
        g.GPUClean()


@pytest.mark.parametrize("tail", [100, 272])
def test_cggi32_tail_launch(tail, rng, monkeypatch):
    """STD128-shaped ring: a batch of one full wave of 4-ciphertext CTAs plus a remainder.  By default it is ONE launch of
    the persistent variant (no wave quantisation); with that switched off the remainder of <= 1 (latency layout) or <= 2
    (CTAs of two) ciphertexts per SM runs as a second launch.  The bits equal the single-launch result and the oracle
    every way, in particular across the seam, for gate, per-ciphertext LUT and explicit accumulators."""
    p = po.Port.params_custom(12, 1024, 1024, Q27, 128, 1 << 7, 32, po.GINX)
    port = po.Port(p)
    sk, bk, ksk, g = _ctx(p, port)
    try:
        q, n, N = p.q, p.n, 1024
        head = _sm_count() * 4
        batch = head + tail
        c1 = rng.integers(0, q, (batch, n + 1), dtype=np.uint64)
        c2 = rng.integers(0, q, (batch, n + 1), dtype=np.uint64)
        want = port.eval_bin_gate(bk, ksk, po.GATES["NAND"], c1, c2, q)
        got = g.EvalBinGate("NAND", c1, c2)
        assert g.last_stats.kernel_launches == 3              # affine, ONE persistent blind rotation, key switch
        assert np.array_equal(got, want)
        g.set_option("persistent", 0)
        got = g.EvalBinGate("NAND", c1, c2)
        assert g.last_stats.kernel_launches == 4              # affine, two blind rotations, key switch
        assert np.array_equal(got, want)
        monkeypatch.setenv("TFHE_B200_NO_TAIL", "1")
        single = g.EvalBinGate("NAND", c1, c2)
        assert g.last_stats.kernel_launches == 3
        monkeypatch.delenv("TFHE_B200_NO_TAIL")
        assert np.array_equal(single, want)
        seam = slice(head - 5, head + 5)
        tab = rng.integers(0, q, (batch, q), dtype=np.uint64)  # per-ciphertext tables: the tail launch offsets them
        got = g.BootstrapFunc(c1, q, tab, q)
        assert np.array_equal(got[seam], port.bootstrap_func(bk, ksk, c1[seam], q, tab[seam], q))
        assert np.array_equal(got[-3:], port.bootstrap_func(bk, ksk, c1[-3:], q, tab[-3:], q))
        acc = rng.integers(0, p.Q, (batch, 2, N), dtype=np.uint64)
        am = rng.integers(0, q, (batch, n), dtype=np.uint64)
        got = g.EvalAcc(am, q, acc)
        assert np.array_equal(got[seam], port.eval_acc(bk, am[seam], q, acc[seam]))
        assert np.array_equal(got[-2:], port.eval_acc(bk, am[-2:], q, acc[-2:]))
        got_f = g.BootstrapFunc(c1, q, tab, q)
        g.set_option("persistent", 1)                          # one CTA per SM, every group boundary a potential split
        assert np.array_equal(g.EvalAcc(am, q, acc), got)
        assert np.array_equal(g.BootstrapFunc(c1, q, tab, q), got_f)
    finally:
        g.GPUClean()


def test_dm32_and_cggi64w_tail_launch(rng):
    """The same split for the AP/DM kernel (CTAs of 4 -> CTAs of 2) and the wide 64-bit kernel (CTAs of 2 -> 1)."""
    sm = _sm_count()
    p = po.Port.params_custom(6, 1024, 1024, Q27, 128, 1 << 7, 32, po.AP)
    port = po.Port(p)
    sk, bk, ksk, g = _ctx(p, port)
    try:
        assert g.kernel_variant.startswith("dm_u32")
        q, n = p.q, p.n
        batch = 4 * sm + 2 * sm - 3
        c1 = rng.integers(0, q, (batch, n + 1), dtype=np.uint64)
        c2 = rng.integers(0, q, (batch, n + 1), dtype=np.uint64)
        got = g.EvalBinGate("NAND", c1, c2)
        assert g.last_stats.kernel_launches == 4
        assert np.array_equal(got, port.eval_bin_gate(bk, ksk, po.GATES["NAND"], c1, c2, q))
    finally:
        g.GPUClean()
    p = po.Port.params_custom(4, 2048, 2048, Q54, 64, 1 << 27, 32, po.GINX)
    port = po.Port(p)
    sk, bk, ksk, g = _ctx(p, port)
    try:
        assert g.kernel_variant.startswith("cggi_u64_ntt16x128")
        q, n = p.q, p.n
        batch = 2 * sm + sm - 7
        ct = rng.integers(0, q, (batch, n + 1), dtype=np.uint64)
        tab = rng.integers(0, q, q, dtype=np.uint64)
        want = port.bootstrap_func(bk, ksk, ct, q, tab, q)
        got = g.BootstrapFunc(ct, q, tab, q)
        assert g.last_stats.kernel_launches == 2              # one persistent blind rotation, key switch
        assert np.array_equal(got, want)
        g.set_option("persistent", 0)
        got = g.BootstrapFunc(ct, q, tab, q)
        assert g.last_stats.kernel_launches == 3              # throughput shape + latency-shaped tail, key switch
        assert np.array_equal(got, want)
    finally:
        g.GPUClean()


def test_generic_key_layout_released_by_default(rng, monkeypatch):
    """A production handle frees the generic-layout key once the specialised layout exists (one copy of the key per GPU
    less); asking for the cross-check kernel then fails loudly instead of running on freed memory, results unchanged."""
    from tfhe_gpu_b200.context import TfheB200Error

    monkeypatch.delenv("TFHE_B200_KEEP_GENERIC", raising=False)
    p = po.Port.params_custom(12, 1024, 1024, Q27, 128, 1 << 7, 32, po.GINX)
    port = po.Port(p)
    sk, bk, ksk, g = _ctx(p, port)
    try:
        q, n = p.q, p.n
        c1 = rng.integers(0, q, (9, n + 1), dtype=np.uint64)
        c2 = rng.integers(0, q, (9, n + 1), dtype=np.uint64)
        want = port.eval_bin_gate(bk, ksk, po.GATES["NAND"], c1, c2, q)
        assert np.array_equal(g.EvalBinGate("NAND", c1, c2), want)
        with pytest.raises(TfheB200Error, match="generic key layout"):
            g.set_option("force_generic", 1)
        assert np.array_equal(g.EvalBinGate("NAND", c1, c2), want)
    finally:
        g.GPUClean()
    sk, bk, ksk, g = _ctx(p, port, keep_generic=True)           # the flag in tfhe_b200_params keeps it
    try:
        g.set_option("force_generic", 1)
        assert np.array_equal(g.EvalBinGate("NAND", c1, c2), want)
    finally:
        g.GPUClean()
    # a parameter set without a specialised kernel keeps (needs) the generic layout
    p2 = po.Port.params_custom(12, 256, 256, Q27, 128, 1 << 7, 32, po.GINX)
    port2 = po.Port(p2)
    sk, bk, ksk, g = _ctx(p2, port2)
    try:
        assert g.kernel_variant.startswith("generic")
        c = rng.integers(0, p2.q, (5, p2.n + 1), dtype=np.uint64)
        e = rng.integers(0, p2.q, (5, p2.n + 1), dtype=np.uint64)
        assert np.array_equal(g.EvalBinGate("AND", c, e), port2.eval_bin_gate(bk, ksk, po.GATES["AND"], c, e, p2.q))
    finally:
        g.GPUClean()


def test_stats_fields_of_unrecorded_phases_are_zero(keyset):
    """tfhe_b200_stats: phases a call does not mark stay 0 (they used to keep the previous call's events)."""
    ks = keyset("toy_ginx")
    g = ks.gpu()
    q, n = ks.p.q, ks.p.n
    rng = np.random.default_rng(3)
    c1 = rng.integers(0, q, (64, n + 1), dtype=np.uint64)
    c2 = rng.integers(0, q, (64, n + 1), dtype=np.uint64)
    g.EvalBinGate("NAND", c1, c2)
    st = g.last_stats
    assert st.blind_rotate_ms > 0 and st.keyswitch_ms > 0 and st.total_ms >= st.blind_rotate_ms
    M = rng.integers(0, 8, (64, 16)).astype(np.int64)
    g.CiphertextMulMatrix(c1, M, q)
    st = g.last_stats
    assert st.total_ms > 0 and st.blind_rotate_ms == 0 and st.keyswitch_ms == 0 and st.bootstraps == 0


def test_scattered_gate_entry_point(rng):
    """tfhe_b200_eval_bin_gate_v: one pointer per ciphertext object (what std::vector<LWECiphertext> holds); gather and
    scatter ride inside the chunk pipeline.  Same bits as the dense entry point and the oracle, for a pipelined batch
    and for a small one."""
    import ctypes as C

    from tfhe_gpu_b200.context import load_library

    p = po.Port.params_custom(12, 1024, 1024, Q27, 128, 1 << 7, 32, po.GINX)
    port = po.Port(p)
    sk, bk, ksk, g = _ctx(p, port)
    L = load_library()
    try:
        q, n = p.q, p.n
        for batch in (3 * _sm_count() * 4 + 77, 9):
            # every ciphertext its own allocation, in shuffled order, like heap-allocated LWECiphertext objects
            a1 = [rng.integers(0, q, n, dtype=np.uint64) for _ in range(batch)]
            a2 = [rng.integers(0, q, n, dtype=np.uint64) for _ in range(batch)]
            b1 = rng.integers(0, q, batch, dtype=np.uint64)
            b2 = rng.integers(0, q, batch, dtype=np.uint64)
            ao = [np.zeros(n, dtype=np.uint64) for _ in range(batch)]
            bo = np.zeros(batch, dtype=np.uint64)
            PP = C.c_void_p * batch
            rc = L.tfhe_b200_eval_bin_gate_v(g._h, 3, batch, PP(*[x.ctypes.data for x in a1]), C.c_void_p(b1.ctypes.data),
                                             PP(*[x.ctypes.data for x in a2]), C.c_void_p(b2.ctypes.data), C.c_uint64(q),
                                             PP(*[x.ctypes.data for x in ao]), C.c_void_p(bo.ctypes.data), None)
            assert rc == 0, L.tfhe_b200_last_error()
            c1 = np.concatenate([np.stack(a1), b1[:, None]], axis=1)
            c2 = np.concatenate([np.stack(a2), b2[:, None]], axis=1)
            want = port.eval_bin_gate(bk, ksk, po.GATES["NAND"], c1, c2, q)
            got = np.concatenate([np.stack(ao), bo[:, None]], axis=1)
            assert np.array_equal(got, want), batch
    finally:
        g.GPUClean()

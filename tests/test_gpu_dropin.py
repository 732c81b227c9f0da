"""Drop-in proof: the reference's UNMODIFIED host code (BinFHEContext / BinFHEScheme batched methods, compiled from
/root/reference) linked against tfhe_gpu_b200/adapter/binfhe_b200_shim.cpp instead of its own .cu files.  The
reference's batched API then runs on our engine and must equal the reference's own scalar CPU API bit for bit --
something the reference's FFT GPU path cannot do."""
import os

import numpy as np
import pytest

from oracle import pyoracle as po

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not os.path.exists(po.DROPIN_SO), reason="oracle/_ref/libtfhe_ref_dropin.so not built")]


def test_reference_batched_api_on_our_engine_toy():
    r = po.Ref.named(po.TOY, po.GINX, so=po.DROPIN_SO)
    r.keygen()
    r.gpu_setup(1)
    try:
        q = r.p.q
        m1 = [i & 1 for i in range(12)]
        m2 = [(i >> 1) & 1 for i in range(12)]
        c1, c2 = r.encrypt_batch(m1, 4, q), r.encrypt_batch(m2, 4, q)
        for g in ("NAND", "XOR", "XNOR_FAST"):
            scalar = r.eval_bin_gate(po.GATES[g], c1, c2, q)                  # reference CPU, scalar API
            batched = r.eval_bin_gate(po.GATES[g], c1, c2, q, batched=True)   # reference batched API -> our engine
            assert np.array_equal(batched, scalar), g
        assert r.decrypt_batch(batched, q, 4) == [1 - (a ^ b) for a, b in zip(m1, m2)]
    finally:
        r.gpu_clean()


def test_reference_batched_functional_api_on_our_engine():
    r = po.Ref.func(po.TOY, True, 12, so=po.DROPIN_SO)
    r.keygen()
    r.gpu_setup(1)
    try:
        q = r.p.q
        p = q // (2 * r.p.beta)
        lut = np.array([((x // (q // p)) ** 3 % p) * (q // p) for x in range(q)], dtype=np.uint64)
        ct = r.encrypt_batch(list(range(p)), p, q)
        assert np.array_equal(r.eval_func(ct, q, lut, batched=True), r.eval_func(ct, q, lut))
        n = r.p.n
        cts = np.random.default_rng(1).integers(0, r.p.qKS, (16, n + 1), dtype=np.uint64)
        M = np.random.default_rng(2).integers(0, 64, (16, 8), dtype=np.int64)
        got = r.mul_matrix(cts, r.p.qKS, M, r.p.qKS)
        ref = (cts.astype(object).T @ M.astype(object)) % int(r.p.qKS)
        assert np.array_equal(got.astype(object), ref.T)
    finally:
        r.gpu_clean()


def test_reference_batched_sign_and_decomp_on_our_engine():
    r = po.Ref.func(po.TOY, False, 17, so=po.DROPIN_SO)
    r.keygen()
    r.gpu_setup(1)
    try:
        Qin, q = 1 << 17, r.p.q
        P = Qin // q * (q // (2 * r.p.beta))
        ct = r.encrypt_batch([P // 2 + i - 2 for i in range(4)], P, Qin)
        assert np.array_equal(r.eval_sign(ct, Qin, batched=True), r.eval_sign(ct, Qin))
        a, am = r.eval_decomp(ct, Qin, batched=True)
        b, bm = r.eval_decomp(ct, Qin)
        assert am == bm and np.array_equal(a, b)
    finally:
        r.gpu_clean()


def test_fused_cpp_adapter_on_reference_objects():
    """tfhe_gpu_b200/adapter/binfhe_b200.hpp: the batched BinFHEContext surface on std::vector<LWECiphertext> through
    the FUSED C-ABI calls, constructed from the reference's own context/keys; equal to the reference scalar CPU API."""
    r = po.Ref.named(po.TOY, po.GINX, so=po.DROPIN_SO)
    r.keygen()
    r.fused_create(1)
    try:
        q = r.p.q
        m1 = [i & 1 for i in range(10)]
        m2 = [(i >> 1) & 1 for i in range(10)]
        c1, c2 = r.encrypt_batch(m1, 4, q), r.encrypt_batch(m2, 4, q)
        for g in ("NAND", "XNOR"):
            assert np.array_equal(r.fused_eval_bin_gate(po.GATES[g], c1, c2, q),
                                  r.eval_bin_gate(po.GATES[g], c1, c2, q)), g
    finally:
        r.fused_destroy()
    r = po.Ref.func(po.TOY, False, 17, so=po.DROPIN_SO)
    r.keygen()
    r.fused_create(1)
    try:
        Qin, q = 1 << 17, r.p.q
        P = Qin // q * (q // (2 * r.p.beta))
        ct = r.encrypt_batch([P // 2 + i - 2 for i in range(4)], P, Qin)
        assert np.array_equal(r.fused_eval_sign(ct, Qin), r.eval_sign(ct, Qin))
        a, am = r.fused_eval_decomp(ct, Qin)
        b, bm = r.eval_decomp(ct, Qin)
        assert am == bm and np.array_equal(a, b)
        # a periodic LUT through EvalFunc (q = 2N here, arbitrary LUTs are rejected like in the reference)
        p = q // (2 * r.p.beta)
        ct2 = r.encrypt_batch(list(range(p)), p, q)
        per = np.array([((x // (q // p)) % (p // 2)) * (q // p) for x in range(q)], dtype=np.uint64)
        per[q // 2:] = per[: q // 2]
        assert np.array_equal(r.fused_eval_func(ct2, q, per), r.eval_func(ct2, q, per))
    finally:
        r.fused_destroy()


def test_fused_cpp_adapter_time_optimization_context():
    """A timeOptimization context (three-key map, binfhecontext.cpp:222-247): the reference's own GPUSetup throws for
    it (binfhecontext.cpp:350-353); the adapter loads all three key sets and the batched EvalSign / EvalDecomp equal
    the reference's scalar CPU path, which switches gadget base as the modulus shrinks
    (binfhe-base-scheme.cpp:342-360, 411-428)."""
    r = po.Ref.func_dynamic(po.TOY, False, 29, so=po.DROPIN_SO)
    r.keygen()
    with pytest.raises(RuntimeError):
        r.gpu_setup(1)                                       # the reference API refuses the context
    r.fused_create(1)
    try:
        Qin, q = 1 << 29, r.p.q
        P = Qin // q * (q // (2 * r.p.beta))
        msgs = [P // 2 + i - 2 for i in range(4)]
        ct = r.encrypt_batch(msgs, P, Qin)
        got = r.fused_eval_sign(ct, Qin)
        assert np.array_equal(got, r.eval_sign(ct, Qin))
        assert r.decrypt_batch(got, q, 2) == [int(m >= P // 2) for m in msgs]
        a, am = r.fused_eval_decomp(ct, Qin)
        b, bm = r.eval_decomp(ct, Qin)
        assert am == bm and np.array_equal(a, b)
    finally:
        r.fused_destroy()

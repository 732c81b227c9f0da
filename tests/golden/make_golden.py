"""Generates the golden vectors in this directory by RUNNING THE UNMODIFIED REFERENCE (oracle/_ref/libtfhe_ref.so,
compiled in place from /root/reference by oracle/Makefile).  Run in the build container:

    make -C oracle ref && python tests/golden/make_golden.py

The reference draws keys and noise from a randomly seeded PRNG (distributiongenerator.h:86-130), so its own tests pin
only decrypted plaintexts; these files pin the BIT-LEVEL behaviour instead: (keys, inputs, per-stage outputs, final
outputs) of the reference's scalar CPU API.  To keep the fixtures small the full-pipeline vectors use the reference's
custom-parameter context (binfhecontext.cpp:42-49) with a tiny ring (n=8, N=32); the stage vectors (NTT, signed
digit decomposition, RoundqQ) use the real TOY / STD128 / 54-bit moduli.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import pyoracle as po  # noqa: E402

Q27 = 134215681


def pipeline(method, tag):
    r = po.Ref.custom(8, 32, 32, Q27, 8, 1 << 14, 4, method)
    r.keygen()
    sk, bk, ksk = r.export_keys()
    q = r.p.q
    rng = np.random.default_rng(1)
    m1 = [i & 1 for i in range(8)]
    m2 = [(i >> 1) & 1 for i in range(8)]
    c1, c2 = r.encrypt_batch(m1, 4, q), r.encrypt_batch(m2, 4, q)
    out = {"params": np.array(list(r.p.as_dict().values()), dtype=np.uint64), "param_names": np.array(list(r.p.as_dict())),
           "sk": sk, "bk": bk, "ksk": ksk, "c1": c1, "c2": c2, "m1": np.array(m1), "m2": np.array(m2)}
    for g in ("NAND", "AND", "OR", "NOR", "XOR_FAST", "XNOR_FAST", "XOR", "XNOR"):
        out["gate_" + g] = r.eval_bin_gate(po.GATES[g], c1, c2, q)
    # stage vectors on the same keys
    a = rng.integers(0, q, (4, r.p.n), dtype=np.uint64)
    acc = rng.integers(0, Q27, (4, 2, r.p.N), dtype=np.uint64)
    out["acc_a"], out["acc_in"], out["acc_out"] = a, acc, r.eval_acc(a, q, acc)
    ext = rng.integers(0, Q27, (4, r.p.N + 1), dtype=np.uint64)
    out["ks_in"], out["ks_out"] = ext, r.key_switch(ext)
    np.savez_compressed(os.path.join(HERE, f"pipeline_{tag}.npz"), **out)
    print(tag, {k: v.shape for k, v in out.items() if hasattr(v, "shape")})


def functional():
    # tiny functional chain: custom context, q = 32 = N so arbitrary LUTs are allowed; beta is fixed at 128 in the
    # reference (binfhecontext.h:348) which exceeds this toy q, so only bit-level agreement is meaningful here.
    r = po.Ref.custom(8, 32, 32, Q27, 8, 1 << 14, 4, po.GINX)
    r.keygen()
    sk, bk, ksk = r.export_keys()
    q = r.p.q
    rng = np.random.default_rng(2)
    ct = rng.integers(0, q, (6, r.p.n + 1), dtype=np.uint64)
    lut_arb = rng.integers(0, q, q, dtype=np.uint64)
    lut_arb[0] = 1
    lut_arb[q // 2] = 5           # neither negacyclic nor periodic
    lut_per = rng.integers(0, q, q, dtype=np.uint64)
    lut_per[q // 2:] = lut_per[: q // 2]
    lut_neg = rng.integers(1, q, q, dtype=np.uint64)
    lut_neg[q // 2:] = q - lut_neg[: q // 2]
    out = {"params": np.array(list(r.p.as_dict().values()), dtype=np.uint64), "sk": sk, "bk": bk, "ksk": ksk, "ct": ct,
           "lut_arb": lut_arb, "lut_per": lut_per, "lut_neg": lut_neg,
           "func_arb": r.eval_func(ct, q, lut_arb), "func_per": r.eval_func(ct, q, lut_per),
           "func_neg": r.eval_func(ct, q, lut_neg)}
    big = 1 << 9   # "large precision" modulus for floor / sign / decomp on the toy ring: 512 -> 2*beta*512/32 ...
    ctb = rng.integers(0, big, (4, r.p.n + 1), dtype=np.uint64)
    out["big_mod"] = np.array([big], dtype=np.uint64)
    out["ct_big"] = ctb
    out["floor"] = r.eval_floor(ctb, big)
    np.savez_compressed(os.path.join(HERE, "functional_tiny.npz"), **out)
    print("functional", {k: v.shape for k, v in out.items()})


def stages():
    out = {}
    rng = np.random.default_rng(3)
    for tag, ref in (("toy", po.Ref.named(po.TOY, po.GINX)), ("std128", po.Ref.named(po.STD128, po.GINX)),
                     ("func54", po.Ref.func(po.TOY, True, 12))):
        p = ref.p
        x = rng.integers(0, p.Q, p.N, dtype=np.uint64)
        out[f"{tag}_params"] = np.array(list(p.as_dict().values()), dtype=np.uint64)
        out[f"{tag}_ntt_in"] = x
        out[f"{tag}_ntt_fwd"] = ref.ntt(x)
        out[f"{tag}_ntt_inv"] = ref.ntt(x, True)
        x2 = rng.integers(0, p.Q, (2, p.N), dtype=np.uint64)
        x2[0, :4] = [0, 1, p.Q - 1, p.Q >> 1]            # edge values of the centred representative
        x2[1, :4] = [(p.Q >> 1) - 1, (p.Q >> 1) + 1, p.baseG // 2, p.baseG // 2 - 1]
        out[f"{tag}_sdd_in"] = x2
        out[f"{tag}_sdd_out"] = ref.signed_digit_decompose(x2)
        # RoundqQ through ModSwitch: Q -> qKS and qKS -> q
        v = rng.integers(0, p.Q, (1, 257), dtype=np.uint64)
        v[0, :4] = [0, 1, p.Q - 1, p.Q >> 1]
        out[f"{tag}_ms1_in"] = v
        out[f"{tag}_ms1_out"] = ref.mod_switch(v, p.Q, p.qKS)
        w = rng.integers(0, p.qKS, (1, 257), dtype=np.uint64)
        out[f"{tag}_ms2_in"] = w
        out[f"{tag}_ms2_out"] = ref.mod_switch(w, p.qKS, p.q)
    np.savez_compressed(os.path.join(HERE, "stages.npz"), **out)
    print("stages", len(out))


if __name__ == "__main__":
    if not po.have_ref():
        raise SystemExit("oracle/_ref/libtfhe_ref.so is missing: run `make -C oracle ref` first")
    pipeline(po.GINX, "ginx_tiny")
    pipeline(po.AP, "ap_tiny")
    functional()
    stages()

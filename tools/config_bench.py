"""Throughput of every BASELINE.json configuration on one B200 (the per-GPU share of the 8-GPU configurations), through
the operator-level API with DEVICE-resident inputs: `python tools/config_bench.py [cfg ...] > profiles/rNN_other_configs.json`.

  cfg0  TOY CGGI        EvalBinGate(NAND)               batch 1024
  cfg1  STD128 CGGI     EvalBinGate(NAND / AND / XOR)   batch 16384
  cfg2  STD128 AP (DM)  EvalBinGate(NAND)               batch 16384   (cfg2b: the named STD128_AP set, batch 4096)
  cfg3  logQ = 12       EvalFunc, arbitrary LUT x^3     batch 1024  (8192 over 8 GPUs)
  cfg4  logQ = 17       EvalSign, EvalDecomp, EvalFloor batch 512   (4096 over 8 GPUs) + CiphertextMulMatrix 1024 x 1024

Inputs are uniform random ciphertexts (the path is data-oblivious for CGGI / key switch; DM skips zero refresh digits,
~1/32 of the steps); parity for the same operators at the same parameter sets is tests/test_gpu_parity_std128.py.  The
oracle is used here for key generation only (test infrastructure; the timed path is the CUDA engine).  Each figure is
the median of 3 timed calls after 2 warm-up calls; blind-rotation share and IMAD32 roofline fraction come from the
engine's own CUDA-event statistics of the last call."""
import json
import os
import statistics
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from oracle import pyoracle as po  # noqa: E402
from tfhe_gpu_b200 import BinFHEContextB200  # noqa: E402

IMAD_PEAK = 18.5e12   # measured (profiles/r01_imad_peak.json); bench.py re-measures it live for the headline


def mm_cggi(n, N, digits):
    """modular multiplications per bootstrap, d = 2 * digitsG rows."""
    d = 2 * digits
    logN = N.bit_length() - 1
    return n * ((d + 2) * (N // 2) * logN + 4 * d * N)


def mm_dm(n, N, digits, baseR, digitsR):
    d = 2 * digits
    logN = N.bit_length() - 1
    return n * digitsR * (1 - 1 / baseR) * ((d + 2) * (N // 2) * logN + 2 * (d - 1) * N)


def timed(fn, reps=3, warm=2):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        t = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t)
    return statistics.median(ts)


def setup(p, seed=1):
    port = po.Port(p)
    t = time.time()
    sk, bk, ksk = port.keygen(seed)
    kg = time.time() - t
    t = time.time()
    ctx = BinFHEContextB200().GPUSetup(p.as_dict(), bk, ksk, numGPUs=1)
    return ctx, {"keygen_s": round(kg, 1), "gpu_setup_s": round(time.time() - t, 2), "kernel": ctx.kernel_variant}


def rand_ct(rng, batch, n, mod):
    return torch.from_numpy(rng.integers(0, mod, (batch, n + 1), dtype=np.int64)).cuda()


def op_entry(ctx, batch, dt, imad_per_bootstrap):
    st = ctx.last_stats
    boots = st.bootstraps                                   # chained bootstraps per input ciphertext
    e = {"batch": batch, "ms": round(dt * 1e3, 3), "ops_per_s": round(batch / dt, 1), "bootstraps_per_op": boots,
         "bootstraps_per_s": round(boots * batch / dt, 1), "blind_rotate_ms": round(st.blind_rotate_ms, 3),
         "keyswitch_ms": round(st.keyswitch_ms, 3), "kernel_launches": st.kernel_launches}
    # the engine's blind-rotation timer brackets ONE bootstrap of the chain; multi-bootstrap operators are charged
    # against the whole device time (key switches and glue included, so the fraction is a lower bound)
    if boots == 1 and st.blind_rotate_ms > 0:
        e["imad32_frac"] = round(imad_per_bootstrap * batch / (st.blind_rotate_ms * 1e-3) / IMAD_PEAK, 3)
    elif st.total_ms > 0:
        e["device_ms"] = round(st.total_ms, 3)
        e["imad32_frac_whole_op"] = round(imad_per_bootstrap * boots * batch / (st.total_ms * 1e-3) / IMAD_PEAK, 3)
    return e


def gates(name, p, batch, gate_list, imad):
    ctx, info = setup(p)
    rng = np.random.default_rng(0)
    c1, c2 = rand_ct(rng, batch, p.n, p.q), rand_ct(rng, batch, p.n, p.q)
    out = {"params": name, **info}
    try:
        for g in gate_list:
            dt = timed(lambda: ctx.EvalBinGate(g, c1, c2))
            out[g] = op_entry(ctx, batch, dt, imad)
    finally:
        ctx.GPUClean()
    return out


def cfg0():
    p = po.Port.params_named(po.TOY, po.GINX)
    return gates("TOY CGGI", p, 1024, ["NAND"], 3 * mm_cggi(p.n, p.N, p.digitsG))


def cfg1():
    p = po.Port.params_named(po.STD128, po.GINX)
    return gates("STD128 CGGI", p, 16384, ["NAND", "AND", "XOR"], 3 * mm_cggi(p.n, p.N, p.digitsG))


def cfg2():
    p = po.Port.params_named(po.STD128, po.AP)       # STD128 ring with the DM accumulator
    return gates("STD128, method AP (DM)", p, 16384, ["NAND"], 3 * mm_dm(p.n, p.N, p.digitsG, p.baseR, p.digitsR))


def cfg2b():
    p = po.Port.params_named(po.STD128_AP, po.AP)    # the named STD128_AP set (n = 503, baseG = 2^9): generic kernel
    return gates("STD128_AP set, method AP (DM)", p, 4096, ["NAND"], 3 * mm_dm(p.n, p.N, p.digitsG, p.baseR, p.digitsR))


def cfg3():
    p = po.Port.params_func(po.STD128, True, 12)
    ctx, info = setup(p)
    rng = np.random.default_rng(3)
    batch, q = 1024, p.q
    imad = 12 * mm_cggi(p.n, p.N, p.digitsG)
    out = {"params": "STD128 functional, logQ = 12 (N = 2048, 54-bit Q)", **info}
    try:
        ct = rand_ct(rng, batch, p.n, q)
        pt = q // (2 * p.beta)                                 # plaintext modulus; LUT given on all q inputs
        lut = torch.tensor([((x // (q // pt)) ** 3 % pt) * (q // pt) for x in range(q)], dtype=torch.int64).cuda()
        dt = timed(lambda: ctx.EvalFunc(ct, lut))              # arbitrary class: 2 chained bootstraps
        out["EvalFunc_x3_arbitrary"] = op_entry(ctx, batch, dt, imad)
        dt = timed(lambda: ctx.EvalFloor(ct, q))
        out["EvalFloor"] = op_entry(ctx, batch, dt, imad)
    finally:
        ctx.GPUClean()
    return out


def cfg4():
    p = po.Port.params_func(po.STD128, False, 17)
    ctx, info = setup(p)
    rng = np.random.default_rng(4)
    batch, Qbig = 512, 1 << 17
    imad = 12 * mm_cggi(p.n, p.N, p.digitsG)
    out = {"params": "STD128 large precision, logQ = 17 (N = 2048, 54-bit Q)", **info}
    try:
        ct = rand_ct(rng, batch, p.n, Qbig)
        dt = timed(lambda: ctx.EvalSign(ct, Qbig))
        out["EvalSign"] = op_entry(ctx, batch, dt, imad)
        dt = timed(lambda: ctx.EvalDecomp(ct, Qbig))
        out["EvalDecomp"] = op_entry(ctx, batch, dt, imad)
        dt = timed(lambda: ctx.EvalFloor(ct, Qbig))
        out["EvalFloor"] = op_entry(ctx, batch, dt, imad)
        cin = rand_ct(rng, 1024, p.n, Qbig)
        mat = torch.from_numpy(rng.integers(0, 64, (1024, 1024), dtype=np.int64)).cuda()
        dt = timed(lambda: ctx.CiphertextMulMatrix(cin, mat, Qbig))
        out["CiphertextMulMatrix_1024x1024"] = {"ms": round(dt * 1e3, 3), "in": 1024, "out": 1024,
                                                "gmacs_per_s": round(1024 * 1024 * (p.n + 1) / dt / 1e9, 1)}
    finally:
        ctx.GPUClean()
    return out


def adder_netlist(bits):
    nodes, wire = [], 2 * bits

    def add(g, x, y=None):
        nonlocal wire
        nodes.append((g, x, y))
        wire += 1
        return wire - 1

    sums, carry = [], None
    for i in range(bits):
        a, b = i, bits + i
        x = add("XOR_FAST", a, b)
        if carry is None:
            sums.append(x)
            carry = add("AND", a, b)
        else:
            sums.append(add("XOR_FAST", x, carry))
            carry = add("NAND", add("NAND", a, b), add("NAND", x, carry))
    return nodes, sums + [carry]


def circuit():
    """Section 8(f) rank 1: 16-bit ripple-carry adders and a wide one-level netlist, STD128 CGGI, batch 64 and 1024,
    one EvalCircuit submission vs the same nodes as separate EvalBinGate calls on device tensors."""
    p = po.Port.params_named(po.STD128, po.GINX)
    ctx, info = setup(p)
    rng = np.random.default_rng(5)
    out = {"params": "STD128 CGGI", **info}
    try:
        for batch in (64, 1024):
            for name, (nodes, outs, n_in) in {
                "adder16": (*adder_netlist(16), 32),
                "wide_64_nands": ([("NAND", i, i + 1) for i in range(64)], list(range(65, 129)), 65),
            }.items():
                ins = torch.from_numpy(rng.integers(0, p.q, (n_in, batch, p.n + 1), dtype=np.int64)).cuda()
                dt = timed(lambda: ctx.EvalCircuit(ins, nodes, outs))
                st = ctx.last_stats
                boots = st.bootstraps
                e = {"batch": batch, "nodes": len(nodes), "bootstraps_per_element": boots, "ms": round(dt * 1e3, 2),
                     "gates_per_s": round(boots * batch / dt, 1), "kernel_launches": st.kernel_launches}

                def one_by_one():
                    wires = [ins[i] for i in range(n_in)]
                    for g, a, b in nodes:
                        wires.append(ctx.EvalBinGate(g, wires[a], wires[b]))
                    return wires

                dt2 = timed(one_by_one, reps=2, warm=1)
                e["gate_by_gate_ms"] = round(dt2 * 1e3, 2)
                e["gate_by_gate_gates_per_s"] = round(boots * batch / dt2, 1)
                out[f"{name}_batch{batch}"] = e
    finally:
        ctx.GPUClean()
    return out


def keygen():
    """Section 8(f) rank 3: evaluation keys generated on the GPU, straight into device memory, then GPUSetup from device."""
    from tfhe_gpu_b200 import gpu_keygen

    out = {}
    for name, p in (("STD128_CGGI", po.Port.params_named(po.STD128, po.GINX)),
                    ("STD128_AP", po.Port.params_named(po.STD128, po.AP)),
                    ("STD128_logQ17", po.Port.params_func(po.STD128, False, 17))):
        r = np.random.default_rng(1)
        sk, skN = r.integers(-1, 2, p.n).astype(np.int8), r.integers(-1, 2, p.N).astype(np.int8)
        gpu_keygen(p.as_dict(), sk, skN, 1)                      # warm-up (context, allocator)
        torch.cuda.synchronize()
        t = time.perf_counter()
        bk, ksk = gpu_keygen(p.as_dict(), sk, skN, 2)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t
        t = time.perf_counter()
        ctx = BinFHEContextB200().GPUSetup(p.as_dict(), bk, ksk, numGPUs=1)
        ts = time.perf_counter() - t
        ctx.GPUClean()
        out[name] = {"gpu_keygen_s": round(dt, 3), "gpu_setup_from_device_s": round(ts, 3),
                     "bk_gb": round(bk.numel() * 8 / 1e9, 2), "ksk_gb": round(ksk.numel() * 8 / 1e9, 2)}
        del bk, ksk
        torch.cuda.empty_cache()
    return out


def dynamic():
    """Section 8(f) rank 4: EvalSign / EvalDecomp at logQ = 29 (STD128, n = 1305), one key set (baseG = 2^14, the
    batched reference semantics) against the three-key map of a timeOptimization context (2^14 -> 2^18 -> 2^27 as the
    modulus shrinks).  Keys for the three gadget bases are generated on the GPU under the same secret."""
    import math

    from tfhe_gpu_b200 import gpu_keygen

    p = po.Port.params_func(po.STD128, False, 29)
    r = np.random.default_rng(1)
    sk, skN = r.integers(-1, 2, p.n).astype(np.int8), r.integers(-1, 2, p.N).astype(np.int8)
    bk, ksk = gpu_keygen(p.as_dict(), sk, skN, 2)
    ctx = BinFHEContextB200().GPUSetup(p.as_dict(), bk, ksk, numGPUs=1)
    out = {"params": "STD128 large precision, logQ = 29 (N = 2048, 54-bit Q, baseG 2^14 / 2^18 / 2^27)",
           "kernel_own_base": ctx.kernel_variant}
    rng = np.random.default_rng(6)
    batch, Qbig = 512, 1 << 29
    try:
        ct = rand_ct(rng, batch, p.n, Qbig)
        for label in ("single_key", "three_key_map"):
            if label == "three_key_map":
                for base in (1 << 18, 1 << 27):
                    d = p.as_dict()
                    d["baseG"], d["digitsG"] = base, math.ceil(math.log(p.Q) / math.log(base))
                    bk2, ksk2 = gpu_keygen(d, sk, skN, 3 + base)
                    ctx.AddKeySet(base, bk2, ksk2)
                    del bk2, ksk2
            for op, fn in (("EvalSign", lambda: ctx.EvalSign(ct, Qbig)), ("EvalDecomp", lambda: ctx.EvalDecomp(ct, Qbig))):
                dt = timed(fn)
                st = ctx.last_stats
                out[f"{op}_{label}"] = {"batch": batch, "ms": round(dt * 1e3, 2), "ops_per_s": round(batch / dt, 1),
                                        "bootstraps_per_op": st.bootstraps,
                                        "bootstraps_per_s": round(st.bootstraps * batch / dt, 1)}
    finally:
        ctx.GPUClean()
    return out


if __name__ == "__main__":
    table = {"dynamic": dynamic, "cfg0": cfg0, "cfg1": cfg1, "cfg2": cfg2, "cfg2b": cfg2b, "cfg3": cfg3, "cfg4": cfg4, "circuit": circuit, "keygen": keygen}
    want = sys.argv[1:] or list(table)
    res = {"note": __doc__.split("\n\n")[0], "gpu": torch.cuda.get_device_name(0), "imad_peak_used": IMAD_PEAK}
    for k in want:
        t = time.time()
        try:
            res[k] = table[k]()
        except Exception as e:  # noqa: BLE001
            res[k] = {"error": repr(e)}
        res[k]["wall_s"] = round(time.time() - t, 1)
        print(k, json.dumps(res[k]), file=sys.stderr, flush=True)
    print(json.dumps(res, indent=1))

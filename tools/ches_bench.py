"""Small-batch latency with the protocol of the reference's src/binfhe/examples/CHES-experiments.cpp, through HOST
buffers (what a reference user passes):

  A  TFHE_rs_Compare (:30-62): STD128 GINX, 256 ciphertext pairs, EvalBinGate(AND) called back to back; the reference
     loops 1000 times and prints the total, here REPS calls are timed and the per-call time and the 1000-call
     extrapolation are reported.
  B  main (:64-125): (STD128, arbFunc, logQ = 12, baseG = 2^18), EvalFunc with the LUT of x^3 mod p at batch sizes
     1 .. 512, average of 5 calls after one warm-up call.

`python tools/ches_bench.py > profiles/rNN_ches.json`.  With oracle/_ref/libtfhe_ref_gpu.so present (the reference's own
CUDA path, comparison build of oracle/Makefile `refgpu`) experiment A is also run on it, on the same box.
Measurement infrastructure only: keys for our engine are generated on the GPU (tfhe_b200_keygen), inputs are uniform
random ciphertexts (the path is data-oblivious)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from oracle import pyoracle as po  # noqa: E402
from tfhe_gpu_b200 import BinFHEContextB200, gpu_keygen  # noqa: E402

REPS = int(os.environ.get("CHES_REPS", "100"))
SIZES = [1, 2, 4, 8, 16, 32, 64, 128, 148, 256, 296, 512]


def engine(p, seed=1):
    r = np.random.default_rng(seed)
    sk, skN = r.integers(-1, 2, p.n).astype(np.int8), r.integers(-1, 2, p.N).astype(np.int8)
    bk, ksk = gpu_keygen(p.as_dict(), sk, skN, 2)
    ctx = BinFHEContextB200().GPUSetup(p.as_dict(), bk, ksk, numGPUs=1)
    del bk, ksk
    torch.cuda.empty_cache()
    return ctx


def exp_a():
    p = po.Port.params_named(po.STD128, po.GINX)
    ctx = engine(p)
    rng = np.random.default_rng(0)
    c1 = rng.integers(0, p.q, (256, p.n + 1), dtype=np.uint64)
    c2 = rng.integers(0, p.q, (256, p.n + 1), dtype=np.uint64)
    out = {"params": "STD128 GINX, 256 ciphertext pairs, EvalBinGate(AND), host buffers", "reps": REPS}
    try:
        for _ in range(3):
            ctx.EvalBinGate("AND", c1, c2)
        t = time.perf_counter()
        for _ in range(REPS):
            ctx.EvalBinGate("AND", c1, c2)
        dt = (time.perf_counter() - t) / REPS
        out["ours"] = {"ms_per_call": round(dt * 1e3, 3), "us_per_1000_calls": round(dt * 1e9),
                       "gates_per_s": round(256 / dt, 1), "kernel": ctx.kernel_variant}
    finally:
        ctx.GPUClean()
    so = os.path.join(os.path.dirname(po.REF_SO), "libtfhe_ref_gpu.so")
    if os.path.exists(so):
        try:
            r = po.Ref.named(po.STD128, po.GINX, so=so)
            r.keygen()
            r.gpu_setup(1)
            reps = max(5, REPS // 10)
            for _ in range(2):
                r.eval_bin_gate(po.GATES["AND"], c1, c2, p.q, batched=True)
            t = time.perf_counter()
            for _ in range(reps):
                r.eval_bin_gate(po.GATES["AND"], c1, c2, p.q, batched=True)
            dt = (time.perf_counter() - t) / reps
            out["reference_gpu"] = {"ms_per_call": round(dt * 1e3, 3), "us_per_1000_calls": round(dt * 1e9),
                                    "gates_per_s": round(256 / dt, 1), "reps": reps,
                                    "kind": "reference CUDA path (cuFFTDx FFT, SM<900> templates on sm_100)"}
            r.gpu_clean()
        except Exception as e:  # noqa: BLE001
            out["reference_gpu"] = {"error": repr(e)}
    return out


def exp_b():
    p = po.Port.params_func(po.STD128, True, 12, 0, 1 << 18)
    ctx = engine(p)
    rng = np.random.default_rng(1)
    q = p.q
    pt = q // (2 * p.beta)
    lut = np.array([((x // (q // pt)) ** 3 % pt) * (q // pt) for x in range(q)], dtype=np.uint64)
    out = {"params": "STD128 functional, logQ = 12, baseG = 2^18 (N = 2048, 54-bit Q), EvalFunc x^3 mod p, host buffers",
           "kernel": ctx.kernel_variant, "ms_by_batch": {}}
    try:
        for b in SIZES:
            ct = rng.integers(0, q, (b, p.n + 1), dtype=np.uint64)
            ctx.EvalFunc(ct, lut)
            ts = []
            for _ in range(5):
                t = time.perf_counter()
                ctx.EvalFunc(ct, lut)
                ts.append(time.perf_counter() - t)
            out["ms_by_batch"][str(b)] = round(sum(ts) / len(ts) * 1e3, 2)
    finally:
        ctx.GPUClean()
    return out


if __name__ == "__main__":
    res = {"note": __doc__.split("\n\n")[0], "gpu": torch.cuda.get_device_name(0)}
    for name, fn in (("A_and_256", exp_a), ("B_evalfunc_latency", exp_b)):
        t = time.time()
        try:
            res[name] = fn()
        except Exception as e:  # noqa: BLE001
            res[name] = {"error": repr(e)}
        res[name]["wall_s"] = round(time.time() - t, 1)
        print(name, json.dumps(res[name]), file=sys.stderr, flush=True)
    print(json.dumps(res, indent=1))

"""The reference's own timing program, src/binfhe/examples/time-estimate.cpp, re-run against our engine through HOST
buffers: batched NAND (STD128 GINX, :31-57), EvalFunc (STD128, arbFunc, logQ = 12, numDigitsToThrow = 1, :59-94),
EvalFloor (logQ = 11, throw 1, :96-123), EvalSign (logQ = 17, throw 1, :125-156), EvalDecomp (logQ = 23, throw 1,
:158-190); the reference prints `ms / ctx` at batch 16384, here BATCH (default 4096) ciphertexts per call on one GPU.
Keys are generated on the GPU, inputs are uniform random ciphertexts.  `python tools/time_estimate.py > profiles/rNN_time_estimate.json`"""
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

BATCH = int(os.environ.get("TE_BATCH", "4096"))


def engine(p):
    r = np.random.default_rng(1)
    sk, skN = r.integers(-1, 2, p.n).astype(np.int8), r.integers(-1, 2, p.N).astype(np.int8)
    bk, ksk = gpu_keygen(p.as_dict(), sk, skN, 2)
    ctx = BinFHEContextB200().GPUSetup(p.as_dict(), bk, ksk, numGPUs=1)
    del bk, ksk
    torch.cuda.empty_cache()
    return ctx


def timed(fn):
    fn()
    ts = []
    for _ in range(3):
        t = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t)
    return sorted(ts)[1]


def run(name, p, batch, mod, call):
    ctx = engine(p)
    rng = np.random.default_rng(0)
    ct = rng.integers(0, mod, (batch, p.n + 1), dtype=np.uint64)
    try:
        dt = timed(lambda: call(ctx, ct))
        st = ctx.last_stats
        return {"params": name, "kernel": ctx.kernel_variant, "batch": batch, "ms_per_ctx": round(dt / batch * 1e3, 5),
                "ops_per_s": round(batch / dt, 1), "bootstraps_per_op": st.bootstraps,
                "bootstraps_per_s": round(st.bootstraps * batch / dt, 1)}
    finally:
        ctx.GPUClean()


def main():
    res = {"note": __doc__.split("\n\n")[0], "gpu": torch.cuda.get_device_name(0), "batch": BATCH}
    p = po.Port.params_named(po.STD128, po.GINX)
    c2 = np.random.default_rng(5).integers(0, p.q, (4 * BATCH, p.n + 1), dtype=np.uint64)
    res["NAND"] = run("STD128 GINX", p, 4 * BATCH, p.q, lambda c, ct: c.EvalBinGate("NAND", ct, c2))
    p = po.Port.params_func(po.STD128, True, 12, 0, 0, 1)
    q = p.q
    pt = q // (2 * p.beta)
    lut = np.array([((x // (q // pt)) ** 3 % pt) * (q // pt) for x in range(q)], dtype=np.uint64)
    res["EvalFunc"] = run("STD128 arbFunc logQ=12 throw=1", p, BATCH, q, lambda c, ct: c.EvalFunc(ct, lut))
    p = po.Port.params_func(po.STD128, False, 11, 0, 0, 1)
    res["EvalFloor"] = run("STD128 logQ=11 throw=1", p, BATCH, p.q, lambda c, ct: c.EvalFloor(ct, p.q))
    p17 = po.Port.params_func(po.STD128, False, 17, 0, 0, 1)
    res["EvalSign"] = run("STD128 logQ=17 throw=1", p17, BATCH, 1 << 17, lambda c, ct: c.EvalSign(ct, 1 << 17))
    p23 = po.Port.params_func(po.STD128, False, 23, 0, 0, 1)
    res["EvalDecomp"] = run("STD128 logQ=23 throw=1", p23, BATCH, 1 << 23, lambda c, ct: c.EvalDecomp(ct, 1 << 23))
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()

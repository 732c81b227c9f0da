"""Tiny driver for compute-sanitizer runs: one small call per blind-rotation kernel variant."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pyoracle as po
from tfhe_gpu_b200 import BinFHEContextB200

rng = np.random.default_rng(0)


def run(which):
    if which in ("toy_ginx", "toy_ap"):
        p = po.Port.params_named(po.TOY, po.GINX if which == "toy_ginx" else po.AP)
    elif which == "toy_func12":
        p = po.Port.params_func(po.TOY, True, 12)
    elif which == "small_std128":          # STD128 ring (N = 1024, cggi32 SKIP kernel) with a short LWE dimension
        p = po.Port.params_custom(16, 1024, 1024, 134215681, 128, 1 << 7, 32, po.GINX)
    elif which == "small_std128_ap":       # dm32 kernel
        p = po.Port.params_custom(16, 1024, 1024, 134215681, 128, 1 << 7, 32, po.AP)
    port = po.Port(p)
    sk, bk, ksk = port.keygen(3)
    ctx = BinFHEContextB200().GPUSetup(p.as_dict(), bk, ksk, numGPUs=1)
    print("variant", ctx.kernel_variant, flush=True)
    n, q = p.n, p.q
    if which == "toy_func12":
        ct = rng.integers(0, q, (5, n + 1), dtype=np.uint64)
        tab = rng.integers(0, q, q, dtype=np.uint64)
        out = ctx.BootstrapFunc(ct, q, tab, q)
        ok = np.array_equal(out, port.bootstrap_func(bk, ksk, ct, q, tab, q))
    else:
        c1 = rng.integers(0, q, (9, n + 1), dtype=np.uint64)
        c2 = rng.integers(0, q, (9, n + 1), dtype=np.uint64)
        out = ctx.EvalBinGate("NAND", c1, c2)
        ok = np.array_equal(out, port.eval_bin_gate(bk, ksk, po.GATES["NAND"], c1, c2, q))
    print("bit-exact", ok, flush=True)
    ctx.GPUClean()
    return ok


if __name__ == "__main__":
    names = sys.argv[1:] or ["toy_ginx", "small_std128", "small_std128_ap", "toy_func12"]
    oks = [run(w) for w in names]
    sys.exit(0 if all(oks) else 1)

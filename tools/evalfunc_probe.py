import sys, time, os
import numpy as np
sys.path.insert(0, os.getcwd())
import torch
from oracle import pyoracle as po
from tfhe_gpu_b200 import BinFHEContextB200, gpu_keygen
p = po.Port.params_func(po.STD128, True, 12, 0, 1 << 18)
r = np.random.default_rng(1)
sk, skN = r.integers(-1, 2, p.n).astype(np.int8), r.integers(-1, 2, p.N).astype(np.int8)
bk, ksk = gpu_keygen(p.as_dict(), sk, skN, 2)
ctx = BinFHEContextB200().GPUSetup(p.as_dict(), bk, ksk, numGPUs=1)
del bk, ksk
q = p.q; pt = q // (2 * p.beta)
lut = np.array([((x // (q // pt)) ** 3 % pt) * (q // pt) for x in range(q)], dtype=np.uint64)
rng = np.random.default_rng(0)
for b in (16, 64, 148, 256, 512):
    ct = rng.integers(0, q, (b, p.n + 1), dtype=np.uint64)
    for name, fn in (("EvalFunc", lambda: ctx.EvalFunc(ct, lut)), ("BootstrapFunc", lambda: ctx.BootstrapFunc(ct, q, lut, q)), ("EvalFloor", lambda: ctx.EvalFloor(ct, q))):
        fn()
        t = time.perf_counter(); fn(); dt = time.perf_counter() - t
        st = ctx.last_stats
        print(b, name, "wall %.2f ms" % (dt * 1e3), "total %.2f br %.2f ks %.2f h2d %.2f d2h %.2f boots %d launches %d" % (st.total_ms, st.blind_rotate_ms, st.keyswitch_ms, st.h2d_ms, st.d2h_ms, st.bootstraps, st.kernel_launches), flush=True)
ctx.GPUClean()

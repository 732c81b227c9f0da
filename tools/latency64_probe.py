import sys, time, json, os
import numpy as np
sys.path.insert(0, os.getcwd())
import torch
from oracle import pyoracle as po
from tfhe_gpu_b200 import BinFHEContextB200, gpu_keygen
for name, p in (("func12", po.Port.params_func(po.STD128, True, 12)), ("sign17", po.Port.params_func(po.STD128, False, 17))):
    r = np.random.default_rng(1)
    sk, skN = r.integers(-1, 2, p.n).astype(np.int8), r.integers(-1, 2, p.N).astype(np.int8)
    bk, ksk = gpu_keygen(p.as_dict(), sk, skN, 2)
    ctx = BinFHEContextB200().GPUSetup(p.as_dict(), bk, ksk, numGPUs=1)
    del bk, ksk
    rng = np.random.default_rng(0)
    for batch in (16, 148):
        ct = torch.from_numpy(rng.integers(0, p.q, (batch, p.n + 1), dtype=np.int64)).cuda()
        tab = torch.from_numpy(rng.integers(0, p.q, p.q, dtype=np.int64)).cuda()
        outs = {}
        for g in (2, 1, 0):
            ctx.set_option("group", g)
            for it in range(3):
                torch.cuda.synchronize(); t = time.time()
                o = ctx.BootstrapFunc(ct, p.q, tab, p.q)
                torch.cuda.synchronize(); dt = time.time() - t
            outs[g] = o.cpu().numpy()
            print(name, "batch", batch, "group", g, "ms", round(dt * 1e3, 2), ctx.kernel_variant, flush=True)
        assert (outs[1] == outs[2]).all() and (outs[0] == outs[2]).all()
    ctx.GPUClean()

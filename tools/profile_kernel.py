"""One blind rotation of a named parameter set for profiler runs: `python tools/profile_kernel.py SET METHOD BATCH`
(SET = index into the reference's paramsMap order, METHOD = GINX | AP).  Two warm-up calls, one measured call; prints
the kernel variant and the CUDA-event time.  Wrap in `ncu --set full -k regex:br_ -s 2 -c 1` (tools/ncu_traffic.py
summarises the capture).  Measurement infrastructure: keys from tfhe_b200_keygen, uniform random ciphertexts."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from oracle import pyoracle as po  # noqa: E402  (parameter probes only)
from tfhe_gpu_b200 import BinFHEContextB200, gpu_keygen  # noqa: E402

pset, method, batch = int(sys.argv[1]), {"GINX": po.GINX, "AP": po.AP}[sys.argv[2]], int(sys.argv[3])
p = po.Ref.named(pset, method).p
r = np.random.default_rng(pset)
sk, skN = r.integers(-1, 2, p.n).astype(np.int8), r.integers(-1, 2, p.N).astype(np.int8)
bk, ksk = gpu_keygen(p.as_dict(), sk, skN, seed=1)
ctx = BinFHEContextB200().GPUSetup(p.as_dict(), bk, ksk, numGPUs=1)
del bk, ksk
c1 = torch.from_numpy(r.integers(0, p.q, (batch, p.n + 1), dtype=np.int64)).cuda()
c2 = torch.from_numpy(r.integers(0, p.q, (batch, p.n + 1), dtype=np.int64)).cuda()
for _ in range(3):
    ctx.EvalBinGate("NAND", c1, c2)
st = ctx.last_stats
print(json.dumps({"set": pset, "method": sys.argv[2], "batch": batch, "kernel": ctx.kernel_variant,
                  "blind_rotate_ms": st.blind_rotate_ms, "keyswitch_ms": st.keyswitch_ms}), flush=True)
ctx.GPUClean()

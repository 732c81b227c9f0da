"""Reference-GPU comparison: times the reference's OWN CUDA path (cuFFTDx FFT kernels, patched
only to dispatch its SM<900> templates on compute capability 10.0 -- see oracle/Makefile target `refgpu`) on this box,
with the protocol of src/binfhe/examples/time-estimate.cpp:31-57 (batched EvalBinGate(NAND), STD128 GINX, ms / ctx).
Test/measurement infrastructure only."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pyoracle as po  # noqa: E402

SO = os.path.join(os.path.dirname(po.REF_SO), "libtfhe_ref_gpu.so")
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ngpus = int(sys.argv[3]) if len(sys.argv) > 3 else 1      # cc.GPUSetup(numGPUs), binfhecontext.cpp:349-360
r = po.Ref.named(po.STD128, po.GINX, so=SO)
t = time.time(); r.keygen(); t_key = time.time() - t
t = time.time(); r.gpu_setup(ngpus); t_setup = time.time() - t
q = r.p.q
m1 = [i & 1 for i in range(batch)]
m2 = [(i >> 1) & 1 for i in range(batch)]
c1, c2 = r.encrypt_batch(m1, 4, q), r.encrypt_batch(m2, 4, q)
r.eval_bin_gate(po.GATES["NAND"], c1[:512], c2[:512], q, batched=True)          # warm-up
times = []
for _ in range(reps):
    t = time.time()
    out = r.eval_bin_gate(po.GATES["NAND"], c1, c2, q, batched=True)
    times.append(time.time() - t)
dec = r.decrypt_batch(out[:256], q, 4)
ok = dec == [1 - (a & b) for a, b in zip(m1[:256], m2[:256])]
scalar = r.eval_bin_gate(po.GATES["NAND"], c1[:4], c2[:4], q)
dt = sorted(times)[len(times) // 2]
print(json.dumps({"impl": "reference GPU path (FFT, cuFFTDx SM<900> templates on sm_100)", "batch": batch, "n_gpus": ngpus,
                  "p50_s": dt, "gates_per_s": batch / dt, "ms_per_ctx": dt / batch * 1e3, "decrypt_ok": ok,
                  "bit_exact_vs_its_own_cpu_path": bool(np.array_equal(out[:4], scalar)),
                  "keygen_s": t_key, "gpu_setup_s": t_setup}), flush=True)
r.gpu_clean()

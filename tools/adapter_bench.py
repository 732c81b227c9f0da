"""Throughput of the drop-in C++ surface: tfhe_b200::BatchedBinFHE::EvalBinGate(NAND, std::vector<LWECiphertext>,
std::vector<LWECiphertext>) at batch 16384, STD128 CGGI -- the call a user of the reference's BinFHEContext makes
(binfhecontext.cpp:323-325), timed in C++ around the whole call (vector of ciphertext objects in, vector out).  The
reference's own host objects and key generator are used (oracle/_ref/libtfhe_ref_dropin.so links the UNMODIFIED
reference host code against libtfhe_b200.so); the result is checked against the reference's scalar CPU API on a slice.
Measurement infrastructure: `python tools/adapter_bench.py [batch] [reps] [num_gpus] > profiles/rNN_adapter.json`."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pyoracle as po  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 7
ngpus = int(sys.argv[3]) if len(sys.argv) > 3 else 1
SO = os.path.join(os.path.dirname(po.REF_SO), "libtfhe_ref_dropin.so")
r = po.Ref.named(po.STD128, po.GINX, so=SO)
r.keygen()
r.fused_create(ngpus)
q = r.p.q
rng = np.random.default_rng(11)
c1 = rng.integers(0, q, (batch, r.n + 1), dtype=np.uint64)
c2 = rng.integers(0, q, (batch, r.n + 1), dtype=np.uint64)
r.fused_bench_eval_bin_gate(po.GATES["NAND"], c1, c2, q, 2)                       # warm-up
secs, out = r.fused_bench_eval_bin_gate(po.GATES["NAND"], c1, c2, q, reps)
want = r.eval_bin_gate(po.GATES["NAND"], c1[:4], c2[:4], q)                       # the reference's scalar CPU path
p50 = float(np.median(secs))
print(json.dumps({"what": "BatchedBinFHE::EvalBinGate(NAND) on std::vector<LWECiphertext>, STD128 CGGI", "batch": batch,
                  "n_gpus": ngpus, "reps": reps, "p50_s": p50, "gates_per_s": batch / p50,
                  "best_gates_per_s": batch / float(secs.min()), "per_call_s": [round(float(x), 5) for x in secs],
                  "bit_exact_vs_reference_scalar_cpu": bool(np.array_equal(out[:4], want))}), flush=True)
r.fused_destroy()

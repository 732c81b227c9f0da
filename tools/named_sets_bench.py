"""Throughput of EvalBinGate(NAND) for every named parameter set of the reference's paramsMap
(binfhecontext.cpp:138-156), CGGI/GINX and (where the DM key fits) AP, on one B200 with device-resident inputs:
kernel variant, gates/s and fraction of the integer-pipe roofline by the SURVEY 8(d) convention (3 IMAD32 per modular
multiplication below 2^32, 12 above).  Keys are generated on the GPU (tfhe_b200_keygen); parity for the same sets is
tests/test_gpu_named_sets.py.  `python tools/named_sets_bench.py [batch] > profiles/rNN_named_sets.json`"""
import json
import os
import statistics
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from oracle import pyoracle as po  # noqa: E402  (parameter probes only)
from tfhe_gpu_b200 import BinFHEContextB200, gpu_keygen  # noqa: E402

IMAD_PEAK = 18.5e12
NAMES = ["TOY", "MEDIUM", "STD128_AP", "STD128_APOPT", "STD128", "STD128_OPT", "STD192", "STD192_OPT", "STD256",
         "STD256_OPT", "STD128Q", "STD128Q_OPT", "STD192Q", "STD192Q_OPT", "STD256Q", "STD256Q_OPT", "SIGNED_MOD_TEST"]
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 2048


def mm(p, method):
    logN = p.N.bit_length() - 1
    d = 2 * (p.digitsG - p.numDigitsToThrow)
    if method == po.GINX:
        return p.n * ((d + 2) * (p.N // 2) * logN + 4 * d * p.N)
    return p.n * p.digitsR * (1 - 1 / p.baseR) * ((d + 2) * (p.N // 2) * logN + 2 * (d - 1) * p.N)


out = {"note": __doc__.split("\n\n")[0], "gpu": torch.cuda.get_device_name(0), "batch": batch, "imad_peak_used": IMAD_PEAK,
       "sets": {}}
for pset, name in enumerate(NAMES):
    for method, mname in ((po.GINX, "GINX"), (po.AP, "AP")):
        p = po.Ref.named(pset, method).p if po.have_ref() else po.Port.params_named(pset, method)
        port = po.Port(p)
        key = f"{name}/{mname}"
        if port.bk_words() * 8 > (26 << 30):
            out["sets"][key] = {"skipped": f"DM key of {port.bk_words() * 8 / 2**30:.0f} GB"}
            continue
        r = np.random.default_rng(pset)
        sk, skN = r.integers(-1, 2, p.n).astype(np.int8), r.integers(-1, 2, p.N).astype(np.int8)
        try:
            bk, ksk = gpu_keygen(p.as_dict(), sk, skN, seed=1)
            ctx = BinFHEContextB200().GPUSetup(p.as_dict(), bk, ksk, numGPUs=1)
            del bk, ksk
            torch.cuda.empty_cache()
            b = batch if not ctx.kernel_variant.startswith("generic") else min(batch, 592)
            c1 = torch.from_numpy(r.integers(0, p.q, (b, p.n + 1), dtype=np.int64)).cuda()
            c2 = torch.from_numpy(r.integers(0, p.q, (b, p.n + 1), dtype=np.int64)).cuda()
            ctx.EvalBinGate("NAND", c1, c2)
            ts = []
            for _ in range(3):
                torch.cuda.synchronize()
                t = time.perf_counter()
                ctx.EvalBinGate("NAND", c1, c2)
                torch.cuda.synchronize()
                ts.append(time.perf_counter() - t)
            dt = statistics.median(ts)
            st = ctx.last_stats
            per = 3 if p.Q < (1 << 32) else 12
            e = {"kernel": ctx.kernel_variant, "n": p.n, "N": p.N, "Q_bits": int(p.Q).bit_length(), "baseG": p.baseG,
                 "batch": b, "gates_per_s": round(b / dt, 1), "blind_rotate_ms": round(st.blind_rotate_ms, 3),
                 "keyswitch_ms": round(st.keyswitch_ms, 3),
                 "imad32_frac": round(per * mm(p, method) * b / (st.blind_rotate_ms * 1e-3) / IMAD_PEAK, 3)}
            ctx.GPUClean()
        except Exception as ex:  # noqa: BLE001
            e = {"error": repr(ex)[:200]}
        out["sets"][key] = e
        print(key, json.dumps(e), file=sys.stderr, flush=True)
print(json.dumps(out, indent=1))

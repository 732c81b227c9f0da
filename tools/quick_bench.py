"""Developer micro-harness (not the contract bench): time EvalBinGate on device-resident random ciphertexts."""
import sys, time, json, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import pyoracle as po
from tfhe_gpu_b200 import BinFHEContextB200

name = sys.argv[1] if len(sys.argv) > 1 else "std128_ginx"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
groups = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0]
sets = {"std128_ginx": (po.STD128, po.GINX), "toy_ginx": (po.TOY, po.GINX), "std128_ap": (po.STD128, po.AP)}
if name == "std128_func12":
    p = po.Port.params_func(po.STD128, True, 12)
elif name == "std128_sign17":
    p = po.Port.params_func(po.STD128, False, 17)
else:
    p = po.Port.params_named(*sets[name])
port = po.Port(p)
t = time.time(); sk, bk, ksk = port.keygen(1); print("keygen %.1fs" % (time.time() - t), flush=True)
t = time.time(); ctx = BinFHEContextB200().GPUSetup(p.as_dict(), bk, ksk, numGPUs=1); print("setup %.2fs" % (time.time() - t), ctx.kernel_variant, flush=True)
rng = np.random.default_rng(0)
if name in ("std128_func12", "std128_sign17"):
    # functional path: time ONE BootstrapFunc (blind rotation + MS/KS/MS) on random ciphertexts
    ct = torch.from_numpy(rng.integers(0, p.q, (batch, p.n + 1), dtype=np.int64)).cuda()
    tab = torch.from_numpy(rng.integers(0, p.q, p.q, dtype=np.int64)).cuda()
    for it in range(2):
        torch.cuda.synchronize(); t = time.time()
        ctx.BootstrapFunc(ct, p.q, tab, p.q)
        torch.cuda.synchronize(); dt = time.time() - t
    st = ctx.last_stats
    print(json.dumps({"set": name, "batch": batch, "variant": ctx.kernel_variant, "ms": dt * 1e3,
                      "bootstraps_per_s": batch / dt, "br_ms": st.blind_rotate_ms, "ks_ms": st.keyswitch_ms}), flush=True)
    sys.exit(0)
c1 = torch.from_numpy(rng.integers(0, p.q, (batch, p.n + 1), dtype=np.int64)).cuda()
c2 = torch.from_numpy(rng.integers(0, p.q, (batch, p.n + 1), dtype=np.int64)).cuda()
for g in groups:
    if g >= 0:
        ctx.set_option("force_generic", 0); ctx.set_option("group", g)
    else:
        ctx.set_option("force_generic", 1)
    for it in range(3):
        torch.cuda.synchronize(); t = time.time()
        out = ctx.EvalBinGate("NAND", c1, c2)
        torch.cuda.synchronize(); dt = time.time() - t
        st = ctx.last_stats
    print(json.dumps({"set": name, "batch": batch, "group": g, "variant": ctx.kernel_variant, "ms": dt * 1e3,
                      "gates_per_s": batch / dt, "br_ms": st.blind_rotate_ms, "ks_ms": st.keyswitch_ms,
                      "total_ms": st.total_ms}), flush=True)

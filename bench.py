#!/usr/bin/env python
"""Benchmarks of the batched bootstrapping path.  Default: BASELINE.json configs[1] -- bootstrapped NAND gates/s,
STD128 CGGI, batch 16384.

    python bench.py --gpus N --steps K --warmup W                 # our engine (one process per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...       # the reference's own CPU path on the host cores
    python bench.py --config {toy,std128,ap,func12,sign17,decomp17,mulmatrix}   # the other BASELINE.json configs
    python bench.py --scaling strong --gpus N                     # ONE batch split over the N ranks (configs[1]:
                                                                  # "batch 16384 ... then batch-sharded at 2/4/8 GPUs")
    python bench.py --single-process --gpus N                     # the drop-in GPUSetup(numGPUs = N) path: one process,
                                                                  # one host worker thread per GPU (not under torchrun)

One "step" = one batched call (EvalBinGate / EvalFunc / EvalSign / EvalDecomp / CiphertextMulMatrix) over the rank's
share of the synthetic batch (uniform random ciphertexts: the path is data-oblivious for CGGI and the key switch).
Prints ONE JSON line (rank 0).

* value   : whole-job operations/s with the inputs already resident in HBM (device tensors through the C ABI).
* e2e     : the same metric through the C ABI with HOST (pinned) buffers: H2D of the inputs and D2H of the result are
            inside the timed region of every step.  `e2e_pageable` (N = 1): the same with plain numpy (pageable) buffers,
            which the library stages through its own pinned memory.
* roofline: blind-rotation kernel vs the integer-pipe (IMAD) peak measured live by tfhe_gpu_b200/build/imad_peak;
            roofline_hbm: the MS->KS->MS kernel vs the measured HBM copy bandwidth (MEASURED_PEAKS.json).  `traffic` is
            read from profiles/ncu_traffic.json (written by tools/ncu_traffic.py from an `ncu --set full` capture).
* cpu_baseline: the reference's scalar CPU path (oracle/_ref when present, else the oracle port) timed on a bounded
            sample on ALL host cores of this box (the thread count is set explicitly: torchrun exports
            OMP_NUM_THREADS=1), rank 0 at N = 1 only.
* reference_gpu: the reference's own CUDA path on the same GPUs (comparison build, `GPUSetup(N)`), run after our engine
            has released them.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

HOST_CORES = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
if int(os.environ.get("RANK", "0")) == 0:
    # the CPU arms use every host core; must be in the environment before any OpenMP runtime is loaded
    os.environ["OMP_NUM_THREADS"] = str(HOST_CORES)

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

IMAD_PEAK_FALLBACK = 18.5e12            # profiles/r01_imad_peak.json (sustained), this pool's B200
HBM_FALLBACK_GBS = 6650.0               # B200_PROFILING.md fallback

CONFIGS = {
    # name: (BASELINE.json configs index, kind, default global batch, description)
    "toy": (0, "gate", 1024, "TOY CGGI EvalBinGate"),
    "std128": (1, "gate", 16384, "STD128 CGGI EvalBinGate"),
    "ap": (2, "gate", 16384, "STD128 AP (DM) EvalBinGate"),
    "func12": (3, "func", 8192, "EvalFunc arbitrary LUT x^3, STD128 functional set logQ=12 (N=2048, 54-bit Q)"),
    "sign17": (4, "sign", 4096, "EvalSign, STD128 large-precision set logQ=17 (N=2048, 54-bit Q)"),
    "decomp17": (4, "decomp", 4096, "EvalDecomp, STD128 large-precision set logQ=17 (N=2048, 54-bit Q)"),
    "mulmatrix": (4, "mulmatrix", 1024, "CiphertextMulMatrix 1024 x 1024 (examples/GEMM.cpp:70-100), logQ=17 set"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="std128", choices=sorted(CONFIGS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --batch per GPU; strong: --batch split over the GPUs")
    ap.add_argument("--batch", type=int, default=0, help="ciphertexts per step (0 = the BASELINE.json batch)")
    ap.add_argument("--gate", default="NAND")
    ap.add_argument("--single-process", action="store_true",
                    help="drive all --gpus GPUs from this one process through GPUSetup(numGPUs) (not under torchrun)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="operations in the CPU sample (0 = about 10 s of work)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ref-gpu", action="store_true", help="skip the reference's own GPU path (comparison build)")
    ap.add_argument("--no-pageable", action="store_true", help="skip the pageable-host-buffer e2e leg")
    ap.add_argument("--no-imad-peak", action="store_true",
                    help="use the recorded IMAD peak instead of running the microbenchmark (profiler runs)")
    return ap.parse_args()


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": HBM_FALLBACK_GBS}, "fallback"


def ncu_traffic():
    """Per-launch DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of the kernels, as written by
    tools/ncu_traffic.py from an `ncu --set full` capture: {kernel substring: {"bytes": .., "batch": .., "source": ..}}."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return json.load(f)
    except Exception:
        return {}


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 7:
                    self.samples.append(f)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(s[0]) for s in self.samples)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[3 + i].lower().startswith("active") for s in self.samples)]
        pw = max(float(s[2]) for s in self.samples)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][1]), "power_w_max": pw,
                "reasons": reasons, "samples": len(self.samples)}


def imad_peak_live():
    exe = os.path.join(ROOT, "tfhe_gpu_b200", "build", "imad_peak")
    try:
        out = subprocess.run([exe], capture_output=True, text=True, timeout=120).stdout
        d = json.loads(out.strip().splitlines()[-1])
        return d["variants"]["imad_rrr"]["gops"] * 1e9, "measured live (imad_peak, sustained, all-register IMAD)"
    except Exception as e:  # noqa: BLE001
        return IMAD_PEAK_FALLBACK, f"fallback (profiles/r01_imad_peak.json): {e}"


def reference_gpu_arm(batch, ngpus):
    """The reference's OWN CUDA path (FFT kernels) on this box, when the comparison build exists (oracle/Makefile
    `refgpu`: patched copy that dispatches its SM<900> templates on cc 10.0), with `cc.GPUSetup(ngpus)`
    (binfhecontext.cpp:349-360) and the whole batch in one call (it shards internally, bootstrapping.cu:1616-1667).
    Runs in a subprocess AFTER our engine has released the GPUs; reported beside our numbers, never mixed into them."""
    so = os.path.join(ROOT, "oracle", "_ref", "libtfhe_ref_gpu.so")
    if not os.path.exists(so):
        return {"unavailable": "oracle/_ref/libtfhe_ref_gpu.so not built (make -C oracle refgpu)"}
    batch = min(batch, 65536)   # its pinned staging is sized for 65536 ciphertexts (bootstrapping.cu:904-905)
    try:
        env = dict(os.environ, OMP_NUM_THREADS=str(HOST_CORES))
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ref_gpu_bench.py"), str(batch), "3",
                              str(ngpus)], capture_output=True, text=True, timeout=900, env=env).stdout
        d = json.loads([ln for ln in out.splitlines() if ln.startswith("{")][-1])
        return {"value": d["gates_per_s"], "unit": "gates/s", "ms_per_ctx": d["ms_per_ctx"], "batch": d["batch"],
                "n_gpus": d.get("n_gpus", ngpus), "decrypt_ok": d["decrypt_ok"],
                "kind": "reference GPU path (cuFFTDx FFT, SM<900> templates on sm_100)"}
    except Exception as e:  # noqa: BLE001
        return {"unavailable": f"reference GPU run failed: {e}"}


# --------------------------------------------------------------------------------------------------------------
# workload description (shared by both arms)
# --------------------------------------------------------------------------------------------------------------
def mm_cggi(n, N, kept_digits):
    """modular multiplications per CGGI bootstrap, d = 2 * digits rows."""
    d = 2 * kept_digits
    logN = N.bit_length() - 1
    return n * ((d + 2) * (N // 2) * logN + 4 * d * N)


def mm_dm(n, N, digits, baseR, digitsR):
    d = 2 * digits
    logN = N.bit_length() - 1
    return n * digitsR * (1 - 1 / baseR) * ((d + 2) * (N // 2) * logN + 2 * (d - 1) * N)


class Workload:
    """Parameter set, operator and algorithmic work of one --config (test infrastructure `oracle.pyoracle` supplies the
    parameter probes, the host key generator and the checker; the timed path is the CUDA engine)."""

    def __init__(self, args):
        from oracle import pyoracle as po

        self.po, self.args, self.name = po, args, args.config
        self.cfg_index, self.kind, self.default_batch, self.desc = CONFIGS[args.config]
        if self.name == "toy":
            self.ref_args = ("named", po.TOY, po.GINX)
        elif self.name == "std128":
            self.ref_args = ("named", po.STD128, po.GINX)
        elif self.name == "ap":
            self.ref_args = ("named", po.STD128, po.AP)
        elif self.name == "func12":
            self.ref_args = ("func", po.STD128, True, 12)
        else:
            self.ref_args = ("func", po.STD128, False, 17)
        self.p = (po.Port.params_named(*self.ref_args[1:]) if self.ref_args[0] == "named"
                  else po.Port.params_func(*self.ref_args[1:]))
        p = self.p
        self.ct_mod = (1 << 17) if self.kind in ("sign", "decomp", "mulmatrix") else p.q
        self.boots = {"gate": 3 if args.gate in ("XOR", "XNOR") else 1, "func": 2, "sign": 5, "decomp": 4,
                      "mulmatrix": 0}[self.kind]
        kept = p.digitsG - p.numDigitsToThrow
        per_mm = 3 if p.Q < (1 << 32) else 12      # IMAD32 per modular multiplication (SURVEY 8d convention)
        if self.name == "ap":
            self.imad_per_boot = per_mm * mm_dm(p.n, p.N, p.digitsG, p.baseR, p.digitsR)
        else:
            self.imad_per_boot = per_mm * mm_cggi(p.n, p.N, kept)
        w = 2 if p.qKS <= (1 << 16) else (4 if p.qKS <= (1 << 32) else 8)
        self.ks_bytes_per_boot = p.N * p.dKS * (p.n + 1) * w
        self.ks_table_bytes = p.N * p.dKS * p.baseKS * (p.n + 1) * w
        self.unit = {"gate": "gates/s", "func": "EvalFunc/s", "sign": "EvalSign/s", "decomp": "EvalDecomp/s",
                     "mulmatrix": "output ciphertexts/s"}[self.kind]
        if self.name == "std128" and args.gate == "NAND":
            self.metric = "bootstrapped NAND gates/sec, STD128 CGGI"
        else:
            self.metric = f"{self.desc}{' (' + args.gate + ')' if self.kind == 'gate' else ''} per second"
        self.host_keys = self.name in ("toy", "std128")     # oracle key generator on the host; else GPU key generation

    # ---- keys ----------------------------------------------------------------------------------------------
    def make_keys(self, dev):
        """Rank 0: (bk, ksk) as host numpy arrays (oracle generator) or device tensors (tfhe_b200_keygen)."""
        if self.host_keys:
            port = self.po.Port(self.p)
            _, bk, ksk = port.keygen(20261018)
            return bk, ksk
        from tfhe_gpu_b200 import gpu_keygen

        r = np.random.default_rng(1)
        sk = r.integers(-1, 2, self.p.n).astype(np.int8)
        skN = r.integers(-1, 2, self.p.N).astype(np.int8)
        return gpu_keygen(self.p.as_dict(), sk, skN, key=bytes(range(32)), device=dev.index)   # fixed key: reproducible runs

    # ---- inputs --------------------------------------------------------------------------------------------
    def make_inputs(self, rng, count):
        """Host int64 arrays of one step for `count` units."""
        p = self.p
        if self.kind == "gate":
            return [rng.integers(0, p.q, (count, p.n + 1), dtype=np.int64) for _ in range(2)]
        if self.kind == "func":
            pt = p.q // (2 * p.beta)
            lut = np.array([((x // (p.q // pt)) ** 3 % pt) * (p.q // pt) for x in range(p.q)], dtype=np.int64)
            return [rng.integers(0, p.q, (count, p.n + 1), dtype=np.int64), lut]
        if self.kind in ("sign", "decomp"):
            return [rng.integers(0, self.ct_mod, (count, p.n + 1), dtype=np.int64)]
        # mulmatrix: `count` = number of input (= output) ciphertexts; entries < 2^6 (GEMM.cpp:70-88)
        return [rng.integers(0, self.ct_mod, (count, p.n + 1), dtype=np.int64),
                rng.integers(0, 64, (count, count), dtype=np.int64)]

    def io_bytes(self, count):
        W = (self.p.n + 1) * 8
        if self.kind == "gate":
            return 2 * count * W, count * W
        if self.kind == "func":
            return count * W + self.p.q * 8, count * W
        if self.kind == "sign":
            return count * W, count * W
        if self.kind == "decomp":
            return count * W, 3 * count * W
        return count * W + count * count * 8, count * W

    def run(self, ctx, ins, out=None):
        k = self.kind
        if k == "gate":
            return ctx.EvalBinGate(self.args.gate, ins[0], ins[1], out=out)
        if k == "func":
            return ctx.EvalFunc(ins[0], ins[1])
        if k == "sign":
            return ctx.EvalSign(ins[0], self.ct_mod)
        if k == "decomp":
            return ctx.EvalDecomp(ins[0], self.ct_mod)[0]
        return ctx.CiphertextMulMatrix(ins[0], ins[1], self.ct_mod)

    # ---- oracle (checker) ------------------------------------------------------------------------------------
    def oracle(self, port, bk, ksk, ins, sl):
        po, k, u = self.po, self.kind, (lambda a: np.ascontiguousarray(a).view(np.uint64))
        if k == "gate":
            return port.eval_bin_gate(bk, ksk, po.GATES[self.args.gate], u(ins[0][sl]), u(ins[1][sl]), self.ct_mod)
        if k == "func":
            return port.eval_func(bk, ksk, u(ins[0][sl]), self.ct_mod, u(ins[1]))
        if k == "sign":
            return port.eval_sign(bk, ksk, u(ins[0][sl]), self.ct_mod)
        if k == "decomp":
            return port.eval_decomp(bk, ksk, u(ins[0][sl]), self.ct_mod)[0]
        return port.mul_matrix(u(ins[0]), ins[1], self.ct_mod)[sl]


class CpuArm:
    """Reference scalar CPU path on the host cores.  kind 'reference' = the UNMODIFIED OpenFHE 1.0.4 / TFHE-GPU host
    code compiled from /root/reference into oracle/_ref (the scalar cc.Eval* looped over OpenMP threads, the protocol of
    BASELINE.md section 3); kind 'port' = our C restatement, used only when oracle/_ref is absent."""

    def __init__(self, wl):
        po = wl.po
        self.wl, self.po = wl, po
        self.ref = self.port = None
        # CiphertextMulMatrix has no CPU implementation in the reference (lwe-operation.cu is CUDA only; its example checks
        # against a naive CPUGEMM, GEMM.cpp:110-120): the CPU arm of that config is the oracle port's restatement
        if po.have_ref() and wl.kind != "mulmatrix":
            a = wl.ref_args
            self.ref = po.Ref.named(*a[1:]) if a[0] == "named" else po.Ref.func(*a[1:])
            self.ref.set_num_threads(HOST_CORES)
            self.ref.keygen()                      # the reference draws its own (random) keys
            self.kind, self.cores = "reference", self.ref.num_threads()
        else:
            self.port = po.Port(wl.p)
            self.port.set_num_threads(HOST_CORES)
            _, self.bk, self.ksk = self.port.keygen(7)
            self.kind, self.cores = "port", self.port.num_threads()
        self.rng = np.random.default_rng(5)
        self.unit_s = self._run(self.cores)        # warm-up (lazy NTT tables) and a first estimate: one unit per thread

    def _run(self, sample):
        wl, u = self.wl, (lambda a: np.ascontiguousarray(a).view(np.uint64))
        ins = wl.make_inputs(self.rng, sample)
        t = time.perf_counter()
        if self.ref is not None:
            r, k = self.ref, wl.kind
            if k == "gate":
                r.eval_bin_gate(self.po.GATES[wl.args.gate], u(ins[0]), u(ins[1]), wl.ct_mod)
            elif k == "func":
                r.eval_func(u(ins[0]), wl.ct_mod, u(ins[1]))
            elif k == "sign":
                r.eval_sign(u(ins[0]), wl.ct_mod)
            elif k == "decomp":
                r.eval_decomp(u(ins[0]), wl.ct_mod)
            else:
                r.mul_matrix(u(ins[0]), wl.ct_mod, ins[1], wl.ct_mod)
        else:
            wl.oracle(self.port, self.bk, self.ksk, ins, slice(None))
        return time.perf_counter() - t

    def default_sample(self):
        """About 10 s of CPU work, a whole number of units per thread."""
        if self.wl.kind == "mulmatrix":   # the work per output grows with the matrix: same shape as the GPU arm
            return self.wl.args.batch or self.wl.default_batch
        per_thread = max(1, min(64, int(round(10.0 / max(self.unit_s, 1e-3)))))
        return self.cores * per_thread

    def measure(self, sample):
        dt = self._run(sample)
        what = f"{sample} x {self.wl.desc}" + (f"({self.wl.args.gate})" if self.wl.kind == "gate" else "")
        return dt, {"value": sample / dt, "unit": self.wl.unit, "cores": self.cores, "kind": self.kind,
                    "sample": f"{what}, scalar CPU API over {self.cores} OpenMP threads, {dt:.2f} s",
                    "ms_per_op_per_thread": dt / sample * self.cores * 1e3}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on all host cores (rank 0 only)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    wl = Workload(args)
    arm = CpuArm(wl)
    sample = args.cpu_sample or arm.default_sample()
    # keep the whole --steps/--warmup run within a few minutes
    budget = 150.0 / max(1, args.steps + args.warmup)
    if not args.cpu_sample and arm.unit_s / arm.cores * sample > budget and wl.kind != "mulmatrix":
        sample = max(arm.cores, int(budget / (arm.unit_s / arm.cores)) // arm.cores * arm.cores)
    tot, base, times = 0.0, None, []
    for s in range(args.warmup + args.steps):
        dt, base = arm.measure(sample)
        if s >= args.warmup:
            tot += dt
            times.append(dt)
    value = sample * args.steps / tot
    batch = args.batch or wl.default_batch
    line = {
        "impl": "reference", "metric": wl.metric, "value": value, "unit": wl.unit,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": tot / args.steps * 1e3,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": f"{wl.desc}, reference CPU (NTT) path, bounded sample of {sample} per step (full "
                               f"workload: batch {batch}{' per GPU' if args.scaling == 'weak' else ' in total'})",
                   "baseline_config_index": wl.cfg_index, "p50_ms_per_op": statistics.median(times) / sample * 1e3},
        "cpu_baseline": {"value": value, "unit": wl.unit, "cores": base["cores"], "kind": base["kind"],
                         "sample": base["sample"]},
        "e2e": {"value": value, "unit": wl.unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    from tfhe_gpu_b200 import BinFHEContextB200
    from tfhe_gpu_b200.dist import broadcast_keys, shard_range

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback")
    if args.single_process and world > 1:
        raise SystemExit("bench.py: --single-process is not run under torchrun")
    n_gpus = args.gpus if args.single_process else world
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    wl = Workload(args)
    po, p = wl.po, wl.p
    warmup = max(args.warmup, 3)

    # ---- keys: rank 0 generates, ONE NCCL broadcast over NVLink replicates them (the only collective) ---------
    bk = ksk = None
    if rank == 0:
        bk, ksk = wl.make_keys(dev)
        if torch.is_tensor(bk):
            torch.cuda.synchronize()
    if torch.is_tensor(bk) or (rank != 0 and not wl.host_keys):
        pd = p.as_dict()
        if world > 1:
            if rank != 0:
                bk = torch.empty(po.Port(p).bk_words(), dtype=torch.int64, device=dev)
                ksk = torch.empty(po.Port(p).ksk_words(), dtype=torch.int64, device=dev)
            dist.broadcast(bk, src=0)
            dist.broadcast(ksk, src=0)
            torch.cuda.synchronize(dev)
        bk_t, ksk_t = bk, ksk
    else:
        pd, bk_t, ksk_t = broadcast_keys(p.as_dict() if rank == 0 else None, bk, ksk, dev, src=0)
    t_setup = time.perf_counter()
    ctx = BinFHEContextB200().GPUSetup(pd, bk_t, ksk_t, numGPUs=(n_gpus if args.single_process else 1),
                                       first_device=local)
    t_setup = time.perf_counter() - t_setup
    # every rank keeps a host copy of the keys for its own oracle slice check (the checker needs them on the host)
    bk_h = bk_t.cpu().numpy().view(np.uint64) if torch.is_tensor(bk_t) else bk_t
    ksk_h = ksk_t.cpu().numpy().view(np.uint64) if torch.is_tensor(ksk_t) else ksk_t
    del bk_t, ksk_t, bk, ksk
    torch.cuda.empty_cache()

    # ---- this rank's share of the synthetic batch ----------------------------------------------------------
    batch = args.batch or wl.default_batch
    if wl.kind == "mulmatrix":
        count, global_batch = batch, batch          # one matrix product per step and rank (first GPU only)
    elif args.scaling == "strong":
        _, count = shard_range(batch, world, rank)
        global_batch = batch
    else:
        count, global_batch = batch, batch * world
    if args.single_process and args.scaling == "weak":
        count = global_batch = batch * n_gpus
    rng = np.random.default_rng(1000 + rank)
    host = [torch.from_numpy(a).pin_memory() for a in wl.make_inputs(rng, count)]
    devt = [h.to(dev) for h in host]
    hnp = [h.numpy().view(np.uint64) for h in host]
    hout = None
    if wl.kind == "gate":
        hout = torch.empty((count, p.n + 1), dtype=torch.int64).pin_memory().numpy().view(np.uint64)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up -------------------------------------------------------------------------------------------------
    out = None
    for _ in range(warmup):
        out = wl.run(ctx, devt)

    # ---- correctness guard inside the bench: EVERY rank checks a slice of its own shard against the oracle ------
    port = po.Port(p)
    port.set_num_threads(max(1, HOST_CORES // max(1, world)))
    sl = slice(0, 2 if wl.boots > 1 else 4)
    want = wl.oracle(port, bk_h, ksk_h, [h.numpy() for h in host], sl)
    got = out[sl].cpu().numpy().view(np.uint64)
    ok = bool(np.array_equal(got, want))
    okt = torch.tensor([1 if ok else 0], device=dev)
    if world > 1:
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
    if int(okt[0]) != 1:
        raise SystemExit(f"bench.py: GPU output differs from the oracle on rank {rank if not ok else '?'} -- "
                         "refusing to report a number")
    del bk_h, ksk_h, port

    # ---- timed region: device-resident ---------------------------------------------------------------------------------
    sampler = ClockSampler(local)
    sampler.start()
    br_ms = ks_ms = dev_ms = 0.0
    launches = 0
    step_s = []
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ts = time.perf_counter()
        wl.run(ctx, devt)                           # synchronous C-ABI call, device-resident in/out
        step_s.append(time.perf_counter() - ts)
        st = ctx.last_stats
        br_ms += st.blind_rotate_ms; ks_ms += st.keyswitch_ms; dev_ms += st.total_ms; launches += st.kernel_launches
    barrier()
    dt = time.perf_counter() - t0

    # ---- timed region: end to end through pinned host buffers (H2D + compute + D2H inside every step) ----------------
    for _ in range(2):
        wl.run(ctx, hnp, out=hout)
    e2e_step_s = []
    barrier()
    t1 = time.perf_counter()
    for _ in range(args.steps):
        ts = time.perf_counter()
        wl.run(ctx, hnp, out=hout)
        e2e_step_s.append(time.perf_counter() - ts)
    barrier()
    dt_e2e = time.perf_counter() - t1

    # ---- N = 1: the same through PAGEABLE host buffers (plain numpy, what a std::vector caller has) ------------------
    dt_page = None
    if n_gpus == 1 and not args.no_pageable and wl.kind != "mulmatrix":
        pg = [np.array(a, copy=True) for a in hnp]
        pgout = np.empty_like(hout) if hout is not None else None
        for _ in range(2):
            wl.run(ctx, pg, out=pgout)
        nrep = max(2, min(args.steps, 5))
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        for _ in range(nrep):
            wl.run(ctx, pg, out=pgout)
        torch.cuda.synchronize()
        dt_page = (time.perf_counter() - t2) / nrep
    sampler.stop_flag = True
    sampler.join(timeout=2)

    p50, p50_e2e = statistics.median(step_s), statistics.median(e2e_step_s)
    if world > 1:
        t = torch.tensor([dt, dt_e2e, p50, p50_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt, dt_e2e, p50, p50_e2e = (float(x) for x in t)
        ln = torch.tensor([launches], dtype=torch.int64, device=dev)
        dist.all_reduce(ln, op=dist.ReduceOp.SUM)
        launches = int(ln[0])

    line = None
    if rank == 0:
        total_units = global_batch * args.steps
        value = total_units / dt
        e2e_value = total_units / dt_e2e
        peaks, peaks_src = measured_peaks()
        imad_peak, imad_src = ((IMAD_PEAK_FALLBACK, "recorded (profiles/r02_imad_peak.json)") if args.no_imad_peak
                               else imad_peak_live())
        traffic = ncu_traffic()
        h2d, d2h = wl.io_bytes(count)
        line = {
            "metric": wl.metric, "value": value, "unit": wl.unit,
            "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup, "warmup_done": warmup,
            "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "u32" if p.Q < (1 << 32) else "u64", "data": "synthetic",
            "config": {"workload": f"{wl.desc}{'(' + args.gate + ')' if wl.kind == 'gate' else ''}, batch "
                                   f"{global_batch} in total = {count} per GPU rank (n={p.n} N={p.N} Q={p.Q} "
                                   f"baseG={p.baseG} qKS={p.qKS} baseKS={p.baseKS}), bit-exact vs OpenFHE 1.0.4 CPU path",
                       "baseline_config_index": wl.cfg_index,
                       "batch_per_gpu": count if not args.single_process else global_batch // n_gpus,
                       "global_batch": global_batch,
                       "parallelism": (f"single process, GPUSetup(numGPUs={n_gpus}), one host worker per GPU"
                                       if args.single_process else f"batch-shard x{world}, one process per GPU"),
                       "kernel": ctx.kernel_variant, "bootstraps_per_op": wl.boots,
                       "l2_note": "per-step working set (inputs, extracted LWE, key-switching table) exceeds the "
                                  "126 MB L2; no explicit flush",
                       "p50_ms_per_step": p50 * 1e3,
                       "ms_per_ctx_p50": p50 * 1e3 / max(1, count),
                       "ms_per_bootstrap_p50": p50 * 1e3 / max(1, count * max(1, wl.boots)),
                       "gpu_setup_s": round(t_setup, 3),
                       "oracle_slice_bit_exact_all_ranks": True},
            "e2e": {"value": e2e_value, "unit": wl.unit, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": dt_e2e / args.steps * 1e3, "p50_ms_per_step": p50_e2e * 1e3,
                    "host_buffers": "pinned"},
            "gpu_launches": launches,
        }
        if dt_page is not None:
            line["e2e_pageable"] = {"value": count / dt_page, "unit": wl.unit, "ms_per_step": dt_page * 1e3,
                                    "host_buffers": "pageable numpy arrays, staged through the handle's pinned memory"}
        line["clocks"] = sampler.summary()
        if wl.boots >= 1:
            single = wl.boots == 1 and br_ms > 0
            t_dom = (br_ms if single else dev_ms) / args.steps * 1e-3
            achieved = wl.imad_per_boot * wl.boots * (count if not args.single_process else global_batch // n_gpus) / t_dom
            kname = {"cggi_u32": "br_cggi32_kernel", "dm_u32": "br_dm32_kernel", "cggi_u64_ntt16": "br_cggi64w_kernel",
                     "cggi_u64_ntt32": "br_cggi64_kernel"}
            kern = next((v for k, v in kname.items() if ctx.kernel_variant.startswith(k)), "br_generic_kernel")
            tr = traffic.get(kern)
            line["roofline"] = {
                "bound": "imad", "kernel": kern, "achieved": achieved / 1e12, "peak": imad_peak / 1e12,
                "unit": "TIMAD32/s", "frac": achieved / imad_peak,
                "traffic": (tr["bytes"] * count / tr["batch"]) if tr else None,
                "traffic_source": tr["source"] if tr else None,
                "peak_source": imad_src,
                "algorithmic": f"{wl.imad_per_boot:.4g} IMAD32 per bootstrap x {wl.boots} bootstraps x {count} per "
                               f"launch sequence",
                "timed": ("blind-rotation launch (CUDA events inside the C ABI)" if single else
                          "whole device time of the call (key switches and LWE glue included: a lower bound)"),
                "avg_launch_ms": t_dom * 1e3, "share_of_step": (br_ms if single else dev_ms) / max(dev_ms, 1e-9)}
            if wl.boots == 1 and ks_ms > 0:
                ks_s = ks_ms / args.steps * 1e-3
                ks_compulsory = ((p.N + 1) * 8 + (p.n + 1) * 8) * count + wl.ks_table_bytes
                hbm = peaks.get("hbm_gbs", HBM_FALLBACK_GBS)
                ktr = traffic.get("mkmswitch")
                line["roofline_hbm"] = {
                    "bound": "hbm", "kernel": "mkmswitch", "achieved": ks_compulsory / ks_s / 1e9, "peak": hbm,
                    "unit": "GB/s", "frac": ks_compulsory / ks_s / 1e9 / hbm,
                    "traffic": (ktr["bytes"] * count / ktr["batch"]) if ktr else None,
                    "dram_gbs": (ktr["bytes"] * count / ktr["batch"] / ks_s / 1e9) if ktr else None,
                    "gather_gbs": wl.ks_bytes_per_boot * count / ks_s / 1e9,
                    "peak_source": f"MEASURED_PEAKS.json ({peaks_src})",
                    "algorithmic": f"{(p.N + 1) * 8} B in + {(p.n + 1) * 8} B out per bootstrap x {count} + "
                                   f"{wl.ks_table_bytes} B table once; {wl.ks_bytes_per_boot} B gathered per bootstrap",
                    "avg_launch_ms": ks_s * 1e3, "share_of_step": ks_ms / max(dev_ms, 1e-9)}
        else:
            macs = count * count * (p.n + 1)
            t_dom = dev_ms / args.steps * 1e-3
            line["roofline"] = {"bound": "imad", "kernel": "mul_matrix32_kernel", "achieved": macs / t_dom / 1e12,
                                "peak": imad_peak / 1e12, "unit": "TIMAD32/s", "frac": macs / t_dom / imad_peak,
                                "traffic": None, "peak_source": imad_src,
                                "algorithmic": f"{count} x {count} x {p.n + 1} multiply-accumulates, one IMAD32 each",
                                "avg_launch_ms": t_dom * 1e3, "share_of_step": 1.0}
    ctx.GPUClean()
    del devt
    torch.cuda.empty_cache()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()    # the other ranks exit and release their GPUs; rank 0 runs the comparison arms
    if rank == 0:
        if not args.no_cpu_baseline and n_gpus == 1:
            arm = CpuArm(wl)
            _, line["cpu_baseline"] = arm.measure(args.cpu_sample or arm.default_sample())
        if not args.no_ref_gpu and wl.name == "std128" and args.gate == "NAND":
            line["reference_gpu"] = reference_gpu_arm(global_batch, n_gpus)
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Headline benchmark: bootstrapped NAND gates/s, STD128 CGGI (BASELINE.json configs[1]).

    python bench.py --gpus N --steps K --warmup W            # our engine (one process per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU path on the host cores

One "step" = one batched EvalBinGate(NAND) over `batch` synthetic random ciphertext pairs per GPU (uniform a, b --
the path is data-oblivious for CGGI).  Keys are a real STD128 key set (generated once with the oracle's key
generator so decrypt checks are possible).  Prints ONE JSON line (rank 0).

* value   : whole-job gates/s with the inputs already resident in HBM (device tensors through the C ABI).
* e2e     : the same metric through the C ABI with HOST (pinned) buffers: H2D of both inputs and D2H of the result
            are inside the timed region of every step.
* roofline: blind-rotation kernel vs the integer-pipe (IMAD) peak measured live by tfhe_gpu_b200/build/imad_peak;
            roofline_hbm: the MS->KS->MS kernel vs the measured HBM copy bandwidth (MEASURED_PEAKS.json).
* cpu_baseline: the reference's scalar CPU path (oracle/_ref when present, else the oracle port) timed on a bounded
            sample on this box's host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# algorithmic work per STD128 CGGI bootstrap
IMAD32_PER_BOOTSTRAP = 129.0e6          # 3 IMAD32 per modular multiplication, 42.99 M modmults
KS_BYTES_PER_BOOTSTRAP = 2_101_248      # N*dKS*(n+1)*2 B gathered from the u16 key-switching table
DRAM_TRAFFIC_PER_LAUNCH_16384 = 1.507e9 + 0.202e9      # br_cggi32 (profiles/r01_prof_cggi32_summary.md, r01e)
KS_TABLE_BYTES = 1024 * 2 * 128 * 513 * 2   # N * dKS * baseKS * (n+1) u16 entries
KS_DRAM_TRAFFIC_PER_LAUNCH_16384 = 5.30e9 + 0.07e9     # mkmswitch_packed16 (same file, prof_mkms_r01b)
IMAD_PEAK_FALLBACK = 18.5e12            # profiles/r01_imad_peak.json (sustained, power-capped), this pool's B200
HBM_FALLBACK_GBS = 6650.0               # B200_PROFILING.md fallback


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=16384, help="ciphertext pairs per GPU per step")
    ap.add_argument("--gate", default="NAND")
    ap.add_argument("--cpu-sample", type=int, default=0, help="gates in the cpu_baseline sample (0 = 4 x cores)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ref-gpu", action="store_true", help="skip the reference's own GPU path (comparison build)")
    return ap.parse_args()


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": HBM_FALLBACK_GBS}, "fallback"


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


def reference_gpu_arm(batch):
    """The reference's OWN CUDA path (FFT kernels) on this box, when the comparison build exists
    (oracle/Makefile `refgpu`: patched copy that dispatches its SM<900> templates on cc 10.0).  Runs in a subprocess
    AFTER our engine has released the GPU; reported beside our numbers, never mixed into them."""
    so = os.path.join(ROOT, "oracle", "_ref", "libtfhe_ref_gpu.so")
    if not os.path.exists(so):
        return {"unavailable": "oracle/_ref/libtfhe_ref_gpu.so not built (make -C oracle refgpu)"}
    try:
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ref_gpu_bench.py"), str(batch), "3"],
                             capture_output=True, text=True, timeout=600).stdout
        d = json.loads([l for l in out.splitlines() if l.startswith("{")][-1])
        return {"value": d["gates_per_s"], "unit": "gates/s", "ms_per_ctx": d["ms_per_ctx"], "batch": d["batch"],
                "decrypt_ok": d["decrypt_ok"], "kind": "reference GPU path (cuFFTDx FFT, SM<900> templates on sm_100)"}
    except Exception as e:  # noqa: BLE001
        return {"unavailable": f"reference GPU run failed: {e}"}


def std128_keys():
    from oracle import pyoracle as po   # key generation + CPU baselines only (test infrastructure)

    p = po.Port.params_named(po.STD128, po.GINX)
    port = po.Port(p)
    sk, bk, ksk = port.keygen(20261018)
    return po, p, port, sk, bk, ksk


class CpuArm:
    """Reference scalar CPU path on the host cores.  kind 'reference' = the UNMODIFIED OpenFHE 1.0.4 / TFHE-GPU host
    code compiled from /root/reference into oracle/_ref (the scalar cc.EvalBinGate looped over OpenMP threads, the
    protocol of BASELINE.md section 3); kind 'port' = our C restatement, used only when oracle/_ref is absent."""

    def __init__(self, po, p, port, bk, ksk, gate):
        self.po, self.p, self.port, self.bk, self.ksk, self.gate = po, p, port, bk, ksk, gate
        self.ref = None
        if po.have_ref():
            self.ref = po.Ref.named(po.STD128, po.GINX)
            self.ref.keygen()                      # the reference draws its own (random) keys
            self.kind, self.cores = "reference", self.ref.num_threads()
        else:
            self.kind, self.cores = "port", port.num_threads()
        self.rng = np.random.default_rng(5)
        self._run(self.cores)                      # warm-up: lazy NTT table precomputation

    def _run(self, sample):
        p, g = self.p, self.po.GATES[self.gate]
        c1 = self.rng.integers(0, p.q, (sample, p.n + 1), dtype=np.uint64)
        c2 = self.rng.integers(0, p.q, (sample, p.n + 1), dtype=np.uint64)
        t = time.perf_counter()
        if self.ref is not None:
            self.ref.eval_bin_gate(g, c1, c2, p.q)
        else:
            self.port.eval_bin_gate(self.bk, self.ksk, g, c1, c2, p.q)
        return time.perf_counter() - t

    def measure(self, sample):
        dt = self._run(sample)
        return dt, {"value": sample / dt, "unit": "gates/s", "cores": self.cores, "kind": self.kind,
                    "sample": f"{sample} STD128 CGGI {self.gate} gates, scalar CPU API over {self.cores} OpenMP "
                              f"threads, {dt:.2f} s",
                    "ms_per_gate_per_thread": dt / sample * self.cores * 1e3}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    po, p, port, sk, bk, ksk = std128_keys()
    arm = CpuArm(po, p, port, bk, ksk, args.gate)
    sample = args.cpu_sample or 4 * arm.cores
    tot, base = 0.0, None
    for s in range(args.warmup + args.steps):
        dt, base = arm.measure(sample)
        if s >= args.warmup:
            tot += dt
    value = sample * args.steps / tot
    line = {
        "impl": "reference", "metric": "bootstrapped NAND gates/sec, STD128 CGGI", "value": value, "unit": "gates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": tot / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": f"STD128 CGGI EvalBinGate({args.gate}), reference CPU (NTT) path, bounded sample "
                               f"of {sample} gates per step (full workload: batch {args.batch} per GPU)"},
        "cpu_baseline": {"value": value, "unit": "gates/s", "cores": base["cores"], "kind": base["kind"],
                         "sample": base["sample"]},
        "e2e": {"value": value, "unit": "gates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
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

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    # ---- keys: rank 0 generates, NCCL broadcast over NVLink replicates them (the only collective) -----------
    from tfhe_gpu_b200.dist import broadcast_keys

    po = p = port = sk = bk = ksk = pd = None
    if rank == 0:
        po, p, port, sk, bk, ksk = std128_keys()
        pd = p.as_dict()
    pd, bk_t, ksk_t = broadcast_keys(pd, bk, ksk, dev, src=0)
    ctx = BinFHEContextB200().GPUSetup(pd, bk_t, ksk_t, numGPUs=1, first_device=local)
    del bk_t, ksk_t
    torch.cuda.empty_cache()

    n, q, batch = pd["n"], pd["q"], args.batch
    N = pd["N"]
    rng = np.random.default_rng(1000 + rank)
    h1 = torch.from_numpy(rng.integers(0, q, (batch, n + 1), dtype=np.int64)).pin_memory()
    h2 = torch.from_numpy(rng.integers(0, q, (batch, n + 1), dtype=np.int64)).pin_memory()
    d1, d2 = h1.to(dev), h2.to(dev)
    hout = torch.empty((batch, n + 1), dtype=torch.int64).pin_memory()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        out = ctx.EvalBinGate(args.gate, d1, d2)       # synchronous C-ABI call, device-resident in/out
        st = ctx.last_stats
        return out, st.blind_rotate_ms, st.keyswitch_ms, st.total_ms, st.kernel_launches, st.bootstraps

    # ---- warm-up -------------------------------------------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        out, *_ = step_resident()

    # ---- correctness guard inside the bench: a small slice is checked against the oracle (rank 0) --------------------
    parity = None
    if rank == 0:
        sl = slice(0, 4)
        want = port.eval_bin_gate(bk, ksk, po.GATES[args.gate], h1[sl].numpy().view(np.uint64),
                                  h2[sl].numpy().view(np.uint64), q)
        parity = bool(np.array_equal(out[sl].cpu().numpy().view(np.uint64), want))
        if not parity:
            raise SystemExit("bench.py: GPU output differs from the oracle -- refusing to report a number")

    # ---- timed region: device-resident ---------------------------------------------------------------------------------
    sampler = ClockSampler(local)
    sampler.start()
    br_ms = ks_ms = dev_ms = 0.0
    launches = boots = 0
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        _, a, b, c, l, nb = step_resident()
        br_ms += a; ks_ms += b; dev_ms += c; launches += l; boots = nb
    barrier()
    dt = time.perf_counter() - t0

    # ---- timed region: end to end through host buffers (H2D + compute + D2H inside every step) ------------------------
    for _ in range(2):
        ctx.EvalBinGate(args.gate, h1.numpy().view(np.uint64), h2.numpy().view(np.uint64))
    barrier()
    t1 = time.perf_counter()
    e2e_launches = 0
    for _ in range(args.steps):
        ctx.EvalBinGate(args.gate, h1.numpy().view(np.uint64), h2.numpy().view(np.uint64),
                        out=hout.numpy().view(np.uint64))
        e2e_launches += ctx.last_stats.kernel_launches
    barrier()
    dt_e2e = time.perf_counter() - t1
    sampler.stop_flag = True
    sampler.join(timeout=2)

    if world > 1:
        t = torch.tensor([dt, dt_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt, dt_e2e = float(t[0]), float(t[1])

    if rank == 0:
        total_gates = batch * world * args.steps
        value = total_gates / dt
        e2e_value = total_gates / dt_e2e
        peaks, peaks_src = measured_peaks()
        imad_peak, imad_src = imad_peak_live()
        br_s = br_ms / args.steps * 1e-3          # average launch duration of the dominant kernel (CUDA events)
        ks_s = ks_ms / args.steps * 1e-3
        achieved_imad = IMAD32_PER_BOOTSTRAP * batch / br_s
        achieved_ks = KS_BYTES_PER_BOOTSTRAP * batch / ks_s / 1e9
        ks_compulsory = ((N + 1) * 8 + (n + 1) * 8) * batch + KS_TABLE_BYTES
        line = {
            "metric": "bootstrapped NAND gates/sec, STD128 CGGI", "value": value, "unit": "gates/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": f"STD128 CGGI EvalBinGate({args.gate}), batch {batch} ciphertext pairs per GPU "
                                   f"(n=512 N=1024 Q=134215681 baseG=2^7 qKS=2^14 baseKS=128), bit-exact vs "
                                   f"OpenFHE 1.0.4 CPU path",
                       "batch_per_gpu": batch, "global_batch": batch * world, "parallelism": f"batch-shard x{world}",
                       "kernel": ctx.kernel_variant,
                       "l2_note": "inputs (2 x 67 MB) + extracted LWE (134 MB) + KSK (269 MB) exceed the 126 MB L2 "
                                  "every step; no explicit flush",
                       "ms_per_bootstrap_p50": dt / args.steps * 1e3 / batch,
                       "oracle_slice_bit_exact": parity},
            "e2e": {"value": e2e_value, "unit": "gates/s", "h2d_bytes_per_step": 2 * batch * (n + 1) * 8,
                    "d2h_bytes_per_step": batch * (n + 1) * 8, "ms_per_step": dt_e2e / args.steps * 1e3},
            "gpu_launches": launches,
            "bootstraps_per_gate": boots,
            "clocks": sampler.summary(),
            "roofline": {"bound": "imad", "kernel": "br_cggi32_kernel", "achieved": achieved_imad / 1e12,
                         "peak": imad_peak / 1e12, "unit": "TIMAD32/s", "frac": achieved_imad / imad_peak,
                         "traffic": DRAM_TRAFFIC_PER_LAUNCH_16384 * batch / 16384 if args.gate == "NAND" else None,
                         "traffic_source": "ncu --set full dram__bytes_read.sum + dram__bytes_write.sum of one "
                                           "16384-ciphertext launch (profiles/r01_prof_cggi32_summary.md), bytes",
                         "peak_source": imad_src,
                         "algorithmic": f"{IMAD32_PER_BOOTSTRAP:.4g} IMAD32 per bootstrap x {batch} per launch",
                         "avg_launch_ms": br_s * 1e3, "share_of_step": br_ms / max(dev_ms, 1e-9)},
            # MS->KS->MS (1.3% of the step): a gather of N*dKS table rows per ciphertext.  The gathered bytes are served
            # mostly by L2 (the 134 MB u16 table does not quite fit, so part of it is re-read from HBM: "traffic"),
            # hence three figures: compulsory HBM bytes (in + out + table once) = "achieved", the DRAM throughput
            # ncu measured ("dram_gbs"), and the gather rate the SMs see ("gather_gbs").
            "roofline_hbm": {"bound": "hbm", "kernel": "mkmswitch_packed16_kernel",
                             "achieved": ks_compulsory / ks_s / 1e9,
                             "peak": peaks.get("hbm_gbs", HBM_FALLBACK_GBS), "unit": "GB/s",
                             "frac": ks_compulsory / ks_s / 1e9 / peaks.get("hbm_gbs", HBM_FALLBACK_GBS),
                             "traffic": KS_DRAM_TRAFFIC_PER_LAUNCH_16384 * batch / 16384,
                             "dram_gbs": KS_DRAM_TRAFFIC_PER_LAUNCH_16384 * batch / 16384 / ks_s / 1e9,
                             "gather_gbs": achieved_ks,
                             "peak_source": f"MEASURED_PEAKS.json ({peaks_src})",
                             "algorithmic": f"{(N + 1) * 8} B in + {(n + 1) * 8} B out per bootstrap x {batch} + "
                                            f"{KS_TABLE_BYTES} B table once; {KS_BYTES_PER_BOOTSTRAP} B gathered "
                                            f"(L2) per bootstrap",
                             "avg_launch_ms": ks_s * 1e3, "share_of_step": ks_ms / max(dev_ms, 1e-9)},
        }
        if not args.no_cpu_baseline:
            arm = CpuArm(po, p, port, bk, ksk, args.gate)
            _, line["cpu_baseline"] = arm.measure(args.cpu_sample or 8 * arm.cores)
        ctx.GPUClean()
        if not args.no_ref_gpu and world == 1:
            line["reference_gpu"] = reference_gpu_arm(batch)
        print(json.dumps(line), flush=True)
    ctx.GPUClean()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

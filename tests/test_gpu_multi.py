"""Single-process multi-GPU path of the drop-in API (GPUSetup(numGPUs)): keys encoded on GPU 0 and replicated
peer-to-peer, batch split contiguously, one host worker thread per GPU, no collective in the loop
(reference: bootstrapping.cu:1616-1667, binfhecontext.cpp:349-360).  Needs >= 2 GPUs (skipped otherwise).
Every result is compared with the ORACLE (whole batch for small ones, slices that straddle the shard boundaries for
large ones), not with the single-GPU engine."""
import numpy as np
import pytest

from oracle import pyoracle as po

pytestmark = pytest.mark.gpu


def _ngpu():
    try:
        import torch

        return torch.cuda.device_count()
    except Exception:
        return 0


def _boundary_slices(batch, nd, width=2):
    """Index ranges around every shard boundary of a contiguous balanced split, plus both ends."""
    base, rem = divmod(batch, nd)
    cuts = [k * base + min(k, rem) for k in range(nd + 1)]
    idx = set(range(0, min(width, batch))) | set(range(max(0, batch - width), batch))
    for c in cuts[1:-1]:
        idx |= set(range(max(0, c - width), min(batch, c + width)))
    return np.array(sorted(idx))


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
def test_two_gpu_sharding_bit_exact(keyset, rng):
    from tfhe_gpu_b200 import BinFHEContextB200

    ks = keyset("toy_ginx")
    q, n = ks.p.q, ks.p.n
    ctx = BinFHEContextB200().GPUSetup(ks.p.as_dict(), ks.bk, ks.ksk, numGPUs=2)
    try:
        assert ctx.num_gpus == 2
        c1 = rng.integers(0, q, (37, n + 1), dtype=np.uint64)   # ragged split 19 + 18
        c2 = rng.integers(0, q, (37, n + 1), dtype=np.uint64)
        want = ks.port.eval_bin_gate(ks.bk, ks.ksk, po.GATES["NAND"], c1, c2, q)
        assert np.array_equal(ctx.EvalBinGate("NAND", c1, c2), want)
        assert np.array_equal(ctx.EvalBinGate("XOR", c1, c2),
                              ks.port.eval_bin_gate(ks.bk, ks.ksk, po.GATES["XOR"], c1, c2, q))
        # device-resident inputs on GPU 0 are sharded peer-to-peer
        import torch

        d1 = torch.from_numpy(c1.view(np.int64)).cuda(0)
        d2 = torch.from_numpy(c2.view(np.int64)).cuda(0)
        out = ctx.EvalBinGate("NAND", d1, d2)
        assert np.array_equal(out.cpu().numpy().view(np.uint64), want)
        # a batch smaller than the GPU count
        assert np.array_equal(ctx.EvalBinGate("NAND", c1[:1], c2[:1]), want[:1])
        # operands on another GPU than the handle's first are refused by the host mirror
        from tfhe_gpu_b200.context import TfheB200Error

        with pytest.raises(TfheB200Error):
            ctx.EvalBinGate("NAND", d1.cuda(1), d2.cuda(1))
    finally:
        ctx.GPUClean()


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
def test_multi_gpu_pipeline_and_circuit_vs_oracle(keyset, rng):
    """Pipelined host-buffer gate path (each shard several CTA waves; pageable AND pinned buffers) and a gate netlist on
    every visible GPU, against oracle slices that straddle the shard boundaries."""
    import torch

    from tfhe_gpu_b200 import BinFHEContextB200

    ks = keyset("toy_ginx")
    q, n = ks.p.q, ks.p.n
    nd = min(_ngpu(), 8)
    ctx = BinFHEContextB200().GPUSetup(ks.p.as_dict(), ks.bk, ks.ksk, numGPUs=nd)
    try:
        assert ctx.num_gpus == nd
        batch = nd * (3 * 148 * 8 + 301)          # > 3 waves of 148 x 8 TOY ciphertexts per shard
        c1 = rng.integers(0, q, (batch, n + 1), dtype=np.uint64)
        c2 = rng.integers(0, q, (batch, n + 1), dtype=np.uint64)
        sl = _boundary_slices(batch, nd)
        want = ks.port.eval_bin_gate(ks.bk, ks.ksk, po.GATES["NAND"], c1[sl], c2[sl], q)
        got = ctx.EvalBinGate("NAND", c1, c2)                       # pageable host buffers -> pinned staging
        assert np.array_equal(got[sl], want)
        p1 = torch.from_numpy(c1.view(np.int64)).pin_memory()
        p2 = torch.from_numpy(c2.view(np.int64)).pin_memory()
        pout = torch.empty_like(p1).pin_memory()
        ctx.EvalBinGate("NAND", p1.numpy().view(np.uint64), p2.numpy().view(np.uint64),
                        out=pout.numpy().view(np.uint64))           # pinned host buffers -> direct DMA
        assert np.array_equal(pout.numpy().view(np.uint64), got)
        ins = np.stack([rng.integers(0, q, (9, n + 1), dtype=np.uint64) for _ in range(4)])
        nodes = [("XOR", 0, 1), ("NAND", 2, 3), ("NOT", 4, None), ("OR", 5, 6), ("AND", 4, 7)]
        outs = [4, 6, 7, 8]
        res = ctx.EvalCircuit(ins, nodes, outs)
        wires = [np.ascontiguousarray(x) for x in ins]
        for g, a, b in nodes:                     # the oracle evaluates the netlist gate by gate
            if g == "NOT":
                w = (q - wires[a]) % q
                w[:, -1] = (q // 4 + q - wires[a][:, -1]) % q
                wires.append(w.astype(np.uint64))
            else:
                wires.append(ks.port.eval_bin_gate(ks.bk, ks.ksk, po.GATES[g], wires[a], wires[b], q))
        for k, w in enumerate(outs):
            assert np.array_equal(res[k], wires[w]), f"wire {w}"
    finally:
        ctx.GPUClean()


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
def test_multi_gpu_functional_ops_vs_oracle(keyset):
    """EvalFunc / EvalFloor / EvalSign / EvalDecomp sharded over every visible GPU equal the oracle on the whole batch
    (the per-GPU bodies run on worker threads and never synchronise with the host between bootstraps)."""
    from tfhe_gpu_b200 import BinFHEContextB200

    nd = min(_ngpu(), 8)
    ks = keyset("toy_func12")
    q = ks.p.q
    pt = q // (2 * ks.p.beta)
    lut = np.array([((x // (q // pt)) ** 3 % pt) * (q // pt) for x in range(q)], dtype=np.uint64)
    batch = 2 * nd + 3
    ct = ks.port.encrypt_batch(ks.sk, [i % pt for i in range(batch)], pt, q, 77)
    ctx = BinFHEContextB200().GPUSetup(ks.p.as_dict(), ks.bk, ks.ksk, numGPUs=nd)
    try:
        assert np.array_equal(ctx.EvalFunc(ct, lut), ks.port.eval_func(ks.bk, ks.ksk, ct, q, lut))
        assert np.array_equal(ctx.EvalFloor(ct, q), ks.port.eval_floor(ks.bk, ks.ksk, ct, q, 0))
    finally:
        ctx.GPUClean()
    ks = keyset("toy_sign17")
    q = ks.p.q
    Qbig = 1 << 17
    pt = Qbig // q * (q // (2 * ks.p.beta))
    ct = ks.port.encrypt_batch(ks.sk, [pt // 2 + i - 3 for i in range(batch)], pt, Qbig, 78)
    ctx = BinFHEContextB200().GPUSetup(ks.p.as_dict(), ks.bk, ks.ksk, numGPUs=nd)
    try:
        assert np.array_equal(ctx.EvalSign(ct, Qbig), ks.port.eval_sign(ks.bk, ks.ksk, ct, Qbig))
        got, mods = ctx.EvalDecomp(ct, Qbig)
        want, wmods = ks.port.eval_decomp(ks.bk, ks.ksk, ct, Qbig)
        assert mods == list(wmods) and np.array_equal(got, want)
        assert ctx.last_stats.bootstraps == 4
    finally:
        ctx.GPUClean()

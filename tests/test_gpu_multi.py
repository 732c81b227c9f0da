"""Single-process multi-GPU path of the drop-in API (GPUSetup(numGPUs)): keys encoded on GPU 0 and replicated
peer-to-peer, batch split contiguously, no collective in the loop.  Needs >= 2 GPUs (skipped otherwise)."""
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
    finally:
        ctx.GPUClean()


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
def test_two_gpu_pipeline_circuit_and_functional(keyset, rng):
    """The remaining sharded paths on two GPUs: the pipelined host-buffer gate path (each shard >= 4096), a gate netlist,
    and a functional operator, all against the single-GPU engine (itself checked against the oracle elsewhere)."""
    from tfhe_gpu_b200 import BinFHEContextB200

    ks = keyset("toy_ginx")
    q, n = ks.p.q, ks.p.n
    one = ks.gpu()
    two = BinFHEContextB200().GPUSetup(ks.p.as_dict(), ks.bk, ks.ksk, numGPUs=2)
    try:
        batch = 2 * 4096 + 301
        c1 = rng.integers(0, q, (batch, n + 1), dtype=np.uint64)
        c2 = rng.integers(0, q, (batch, n + 1), dtype=np.uint64)
        assert np.array_equal(two.EvalBinGate("NAND", c1, c2), one.EvalBinGate("NAND", c1, c2))
        ins = np.stack([rng.integers(0, q, (9, n + 1), dtype=np.uint64) for _ in range(4)])
        nodes = [("XOR", 0, 1), ("NAND", 2, 3), ("NOT", 4, None), ("OR", 5, 6), ("AND", 4, 7)]
        outs = [4, 6, 7, 8]
        assert np.array_equal(two.EvalCircuit(ins, nodes, outs), one.EvalCircuit(ins, nodes, outs))
    finally:
        two.GPUClean()
    ks2 = keyset("toy_func12")
    q2 = ks2.p.q
    p2 = q2 // (2 * ks2.p.beta)
    lut = np.array([((x // (q2 // p2)) ** 3 % p2) * (q2 // p2) for x in range(q2)], dtype=np.uint64)
    ct = ks2.port.encrypt_batch(ks2.sk, [i % p2 for i in range(11)], p2, q2, 77)
    two = BinFHEContextB200().GPUSetup(ks2.p.as_dict(), ks2.bk, ks2.ksk, numGPUs=2)
    try:
        assert np.array_equal(two.EvalFunc(ct, lut), ks2.gpu().EvalFunc(ct, lut))
        assert np.array_equal(two.EvalSign(ct, q2), ks2.gpu().EvalSign(ct, q2))
    finally:
        two.GPUClean()

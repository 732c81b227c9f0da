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

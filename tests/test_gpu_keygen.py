"""evaluation keys generated on the GPU (tfhe_b200_keygen).  Key generation is randomised, so
there is nothing to compare bit for bit; instead: (1) everything evaluated under these keys decrypts correctly (gates,
CGGI and DM; functional bootstrapping), (2) the key material has the right structure and noise: RGSW(0) rows and
key-switching rows decrypt to errors of a discrete Gaussian with sigma = 3.19, masks look uniform, (3) generation is a
deterministic function of the seed."""
import numpy as np
import pytest

from oracle import pyoracle as po

pytestmark = pytest.mark.gpu


def _secrets(p, seed):
    r = np.random.default_rng(seed)
    return r.integers(-1, 2, p.n).astype(np.int8), r.integers(-1, 2, p.N).astype(np.int8)


def _sk_mod(sk, qKS):
    return np.array([(int(v) + qKS) % qKS for v in sk], dtype=np.uint64)


@pytest.mark.parametrize("pset,method,gate", [(po.TOY, po.GINX, "NAND"), (po.STD128, po.GINX, "XOR"), (po.TOY, po.AP, "NOR")])
def test_gates_decrypt_under_gpu_generated_keys(pset, method, gate):
    from tfhe_gpu_b200 import BinFHEContextB200, gpu_keygen

    p = po.Port.params_named(pset, method)
    port = po.Port(p)
    sk, skN = _secrets(p, 3)
    bk, ksk = gpu_keygen(p.as_dict(), sk, skN, seed=2024)
    ctx = BinFHEContextB200().GPUSetup(p.as_dict(), bk, ksk, numGPUs=1)
    try:
        q, batch = p.q, 64
        skm = _sk_mod(sk, p.qKS)
        m1 = [i & 1 for i in range(batch)]
        m2 = [(i >> 1) & 1 for i in range(batch)]
        c1 = port.encrypt_batch(skm, m1, 4, q, 11)
        c2 = port.encrypt_batch(skm, m2, 4, q, 12)
        truth = {"NAND": lambda a, b: 1 - (a & b), "XOR": lambda a, b: a ^ b, "NOR": lambda a, b: 1 - (a | b)}[gate]
        out = ctx.EvalBinGate(gate, c1, c2)
        assert port.decrypt_batch(skm, out, q, 4) == [truth(a, b) for a, b in zip(m1, m2)]
        # two more levels on top of the outputs (noise stays under control)
        out2 = ctx.EvalBinGate("NAND", out, c1)
        assert port.decrypt_batch(skm, out2, q, 4) == [1 - (truth(a, b) & a) for a, b in zip(m1, m2)]
    finally:
        ctx.GPUClean()


@pytest.mark.parametrize("pset", [po.TOY, po.STD128])
def test_functional_bootstrapping_under_gpu_generated_keys(pset):
    from tfhe_gpu_b200 import BinFHEContextB200, gpu_keygen

    p = po.Port.params_func(pset, True, 12)
    port = po.Port(p)
    sk, skN = _secrets(p, 5)
    bk, ksk = gpu_keygen(p.as_dict(), sk, skN, seed=7)
    ctx = BinFHEContextB200().GPUSetup(p.as_dict(), bk, ksk, numGPUs=1)
    try:
        q = p.q
        pt = q // (2 * p.beta)
        lut = np.array([((x // (q // pt)) ** 3 % pt) * (q // pt) for x in range(q)], dtype=np.uint64)
        skm = _sk_mod(sk, p.qKS)
        msgs = [i % pt for i in range(3 * pt)]
        ct = port.encrypt_batch(skm, msgs, pt, q, 13)
        assert port.decrypt_batch(skm, ctx.EvalFunc(ct, lut), q, pt) == [m ** 3 % pt for m in msgs]
    finally:
        ctx.GPUClean()


def test_key_material_noise_and_determinism():
    from tfhe_gpu_b200 import gpu_keygen

    p = po.Port.params_named(po.TOY, po.GINX)
    port = po.Port(p)
    n, N, Q, qKS = p.n, p.N, p.Q, p.qKS
    sk, skN = _secrets(p, 9)
    bk_t, ksk_t = gpu_keygen(p.as_dict(), sk, skN, seed=1)
    bk = bk_t.cpu().numpy().view(np.uint64)
    ksk = ksk_t.cpu().numpy().view(np.uint64)
    bk2, _ = gpu_keygen(p.as_dict(), sk, skN, seed=1)
    bk3, _ = gpu_keygen(p.as_dict(), sk, skN, seed=2)
    assert np.array_equal(bk, bk2.cpu().numpy().view(np.uint64))
    assert not np.array_equal(bk, bk3.cpu().numpy().view(np.uint64))
    assert bk.max() < Q and ksk.max() < qKS

    # RGSW rows: [key][i][row][comp][N].  comp1 - comp0' * NTT(sk_ring) - message = NTT(e)
    d2 = 2 * (p.digitsG - p.numDigitsToThrow)
    rows = bk.reshape(2, n, d2, 2, N)
    sk_ntt = port.ntt(np.array([(int(v) + Q) % Q for v in skN], dtype=np.uint64))
    errs, masks = [], []
    for key in range(2):
        for i in range(0, n, 5):
            has_msg = (sk[i] == 1) if key == 0 else (sk[i] == -1)
            for r in range(d2):
                G = pow(int(p.baseG), (r >> 1) + p.numDigitsToThrow, Q) if has_msg else 0
                a = rows[key, i, r, 0].astype(object)
                b = rows[key, i, r, 1].astype(object)
                if r % 2 == 0:
                    a = (a - G) % Q                       # the message sits on the mask component of even rows
                else:
                    b = (b - G) % Q
                e_ntt = np.array((b - a * sk_ntt.astype(object)) % Q, dtype=np.uint64)
                e = port.ntt(e_ntt, inverse=True).astype(np.int64)
                e[e > Q // 2] -= Q
                errs.append(e)
                masks.append(rows[key, i, r, 0].astype(np.float64) / Q)
    e = np.concatenate(errs)
    assert np.abs(e).max() <= 45, "errors must be small integers"
    assert abs(e.mean()) < 0.1 and 3.0 < e.std() < 3.4, (e.mean(), e.std())
    m = np.concatenate(masks)
    assert abs(m.mean() - 0.5) < 0.01 and abs(m.std() - 12 ** -0.5) < 0.01

    # key-switching rows [i][j][k][n+1]: b - <a, s> - s_ring[i] * j * baseKS^k = e (mod qKS)
    kr = ksk.reshape(N, p.baseKS, p.dKS, n + 1)
    s_obj = np.array([int(v) for v in sk], dtype=object)
    es = []
    for i in range(0, N, 37):
        for j in range(0, p.baseKS, 3):
            for k in range(p.dKS):
                row = kr[i, j, k].astype(object)
                msg = int(skN[i]) * j * pow(int(p.baseKS), k, qKS)
                ev = int((row[n] - (row[:n] * s_obj).sum() - msg) % qKS)
                es.append(ev - qKS if ev > qKS // 2 else ev)
    es = np.array(es, dtype=np.int64)
    assert np.abs(es).max() <= 45 and abs(es.mean()) < 0.5 and 2.8 < es.std() < 3.6, (es.mean(), es.std())


def test_secure_generator_key_material():
    """tfhe_b200_keygen proper: 256-bit generator key (explicit, or drawn from the operating system when omitted).  The
    same key reproduces the keys, fresh OS entropy does not, the TEST-ONLY seeded entry point is a different stream, and
    gates decrypt under OS-keyed keys."""
    from tfhe_gpu_b200 import BinFHEContextB200, gpu_keygen

    p = po.Port.params_named(po.TOY, po.GINX)
    port = po.Port(p)
    sk, skN = _secrets(p, 9)
    k1 = bytes(range(32))
    bk_a, ksk_a = gpu_keygen(p.as_dict(), sk, skN, key=k1)
    bk_b, ksk_b = gpu_keygen(p.as_dict(), sk, skN, key=k1)
    bk_c, ksk_c = gpu_keygen(p.as_dict(), sk, skN, key=bytes(reversed(range(32))))
    bk_os1, ksk_os1 = gpu_keygen(p.as_dict(), sk, skN)            # key from getrandom()
    bk_os2, _ = gpu_keygen(p.as_dict(), sk, skN)
    bk_seed, _ = gpu_keygen(p.as_dict(), sk, skN, seed=1)
    assert bool((bk_a == bk_b).all()) and bool((ksk_a == ksk_b).all())
    for other in (bk_c, bk_os1, bk_os2, bk_seed):
        assert not bool((bk_a == other).all())
    assert not bool((bk_os1 == bk_os2).all())
    with pytest.raises(Exception):
        gpu_keygen(p.as_dict(), sk, skN, key=b"short")
    ctx = BinFHEContextB200().GPUSetup(p.as_dict(), bk_os1, ksk_os1, numGPUs=1)
    try:
        q, batch = p.q, 32
        skm = _sk_mod(sk, p.qKS)
        m1 = [i & 1 for i in range(batch)]
        m2 = [(i >> 1) & 1 for i in range(batch)]
        c1 = port.encrypt_batch(skm, m1, 4, q, 21)
        c2 = port.encrypt_batch(skm, m2, 4, q, 22)
        out = ctx.EvalBinGate("NAND", c1, c2)
        assert port.decrypt_batch(skm, out, q, 4) == [1 - (a & b) for a, b in zip(m1, m2)]
    finally:
        ctx.GPUClean()

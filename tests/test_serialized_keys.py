"""GPUSetup straight from OpenFHE's serialized keys (tfhe_b200_setup_from_serialized).

CPU part (-m "not gpu"): the stream indexer / gatherer is host logic -- the committed streams, written by the REFERENCE's
own Serial::SerializeToFile(..., SerType::BINARY) (tests/golden/make_serialized_fixture.py, following
examples/boolean-serial-binary.cpp:76-88), must flatten to exactly the arrays the reference's GPUSetup ordering gives
(bootstrapping.cu:933-975), CGGI and DM (null a0 = 0 entries); malformed streams fail with a message; and, when
oracle/_ref is built here, freshly serialized keys of full named sets round-trip too.
GPU part (-m gpu): a handle made from the streams evaluates the same bits as one made from the flat arrays and as the
reference's scalar CPU path."""
import os

import numpy as np
import pytest

from oracle import pyoracle as po

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _fixture(name):
    z = np.load(os.path.join(GOLD, f"serialized_{name}_tiny.npz"))
    bks = open(os.path.join(GOLD, f"serialized_{name}_tiny_bk.bin"), "rb").read()
    ksks = open(os.path.join(GOLD, f"serialized_{name}_tiny_ksk.bin"), "rb").read()
    pd = {str(k): int(v) for k, v in zip(z["param_names"], z["params"])}
    return z, bks, ksks, pd


@pytest.mark.parametrize("name", ["ginx", "ap"])
def test_streams_flatten_to_the_reference_element_order(name):
    from tfhe_gpu_b200 import flatten_serialized, serialized_info

    z, bks, ksks, pd = _fixture(name)
    info = serialized_info(bks, ksks)
    assert (info.N, info.Q, info.psi, info.n, info.qKS) == (pd["N"], pd["Q"], pd["psi"], pd["n"], pd["qKS"])
    assert (info.ks_N, info.baseKS, info.dKS) == (pd["N"], pd["baseKS"], pd["dKS"])
    assert info.bk_rows == 2 * pd["digitsG"]
    dims = list(info.bk_dim)
    assert dims == ([1, 2, pd["n"]] if name == "ginx" else [pd["n"], pd["baseR"], pd["digitsR"]])
    bk, ksk = flatten_serialized(bks, ksks)
    assert np.array_equal(bk, z["bk"])
    assert np.array_equal(ksk, z["ksk"])


def test_malformed_streams_are_refused():
    from tfhe_gpu_b200 import TfheB200Error, serialized_info

    _, bks, ksks, _ = _fixture("ginx")
    with pytest.raises(TfheB200Error, match="truncated"):
        serialized_info(bks[:len(bks) // 2], ksks)
    with pytest.raises(TfheB200Error, match="truncated"):
        serialized_info(bks, ksks[:1000])
    with pytest.raises(TfheB200Error, match="trailing"):
        serialized_info(bks + b"\x00", ksks)
    with pytest.raises(TfheB200Error):
        serialized_info(ksks, bks)                       # swapped
    with pytest.raises(TfheB200Error, match="little-endian"):
        serialized_info(b"\x00" + bks[1:], ksks)
    bad = bytearray(bks)
    bad[1:5] = (0x12345678).to_bytes(4, "little")      # polymorphic id of another type
    with pytest.raises(TfheB200Error, match="polymorphic"):
        serialized_info(bytes(bad), ksks)


@pytest.mark.parametrize("name", ["ginx", "ap"])
def test_corrupted_streams_never_crash_the_reader(name):
    """Key files come from outside: random truncations, byte flips and hostile length fields (huge vector sizes) must
    end in a TfheB200Error or in a successful parse -- never in a crash, a hang or an absurd allocation."""
    from tfhe_gpu_b200 import TfheB200Error, flatten_serialized, serialized_info

    _, bks, ksks, _ = _fixture(name)
    rng = np.random.default_rng(7)

    def attempt(b, k):
        try:
            info = serialized_info(b, k)
            assert info.bk_words < (1 << 32) and info.ksk_words < (1 << 32)
            if info.bk_words * 8 <= 4 * len(b) and info.ksk_words * 8 <= 4 * len(k):
                flatten_serialized(b, k)
        except TfheB200Error:
            pass

    for _ in range(60):
        for which in (0, 1):
            src = bytearray(bks if which == 0 else ksks)
            kind = rng.integers(0, 4)
            if kind == 0:                                   # truncate
                src = src[:int(rng.integers(0, len(src)))]
            elif kind == 1:                                 # flip a few bytes in the structural head of the stream
                for pos in rng.integers(0, min(len(src), 400), 3):
                    src[int(pos)] ^= int(rng.integers(1, 256))
            elif kind == 2:                                 # flip bytes anywhere
                for pos in rng.integers(0, len(src), 4):
                    src[int(pos)] ^= int(rng.integers(1, 256))
            else:                                           # hostile 64-bit length field somewhere in the head
                pos = int(rng.integers(1, min(len(src), 300) - 8))
                src[pos:pos + 8] = int(rng.integers(1 << 40, 1 << 63)).to_bytes(8, "little")
            if which == 0:
                attempt(bytes(src), ksks)
            else:
                attempt(bks, bytes(src))


@pytest.mark.skipif(not po.have_ref(), reason="oracle/_ref/libtfhe_ref.so not built")
@pytest.mark.parametrize("method", [po.GINX, po.AP])
def test_fresh_reference_streams_round_trip(method, tmp_path):
    """A full named set (TOY), serialized by the reference just now."""
    from tfhe_gpu_b200 import flatten_serialized

    r = po.Ref.named(po.TOY, method)
    r.keygen()
    bf, kf = str(tmp_path / "bk.bin"), str(tmp_path / "ksk.bin")
    r.serialize_keys(bf, kf)
    _, bk, ksk = r.export_keys()
    gbk, gksk = flatten_serialized(np.fromfile(bf, dtype=np.uint8), np.fromfile(kf, dtype=np.uint8))
    assert np.array_equal(gbk, bk) and np.array_equal(gksk, ksk)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["ginx", "ap"])
def test_gpu_setup_from_committed_streams(name):
    from tfhe_gpu_b200 import BinFHEContextB200, TfheB200Error

    z, bks, ksks, pd = _fixture(name)
    a = BinFHEContextB200().GPUSetupFromSerialized(pd, bks, ksks, numGPUs=1)
    b = BinFHEContextB200().GPUSetup(pd, z["bk"], z["ksk"], numGPUs=1)
    try:
        got = a.EvalBinGate("NAND", z["c1"], z["c2"])
        assert np.array_equal(got, z["nand"])             # the reference's scalar CPU result, committed
        assert np.array_equal(b.EvalBinGate("NAND", z["c1"], z["c2"]), got)
        assert a.kernel_variant == b.kernel_variant
    finally:
        a.GPUClean()
        b.GPUClean()
    wrong = dict(pd, n=pd["n"] + 1)
    with pytest.raises(TfheB200Error, match="does not match"):
        BinFHEContextB200().GPUSetupFromSerialized(wrong, bks, ksks, numGPUs=1)
    with pytest.raises(TfheB200Error, match="root of unity"):
        BinFHEContextB200().GPUSetupFromSerialized(dict(pd, psi=pd["psi"] + 1), bks, ksks, numGPUs=1)


@pytest.mark.gpu
@pytest.mark.skipif(not po.have_ref(), reason="oracle/_ref/libtfhe_ref.so not built")
@pytest.mark.parametrize("pset,method", [(po.STD128, po.GINX), (po.TOY, po.AP)])
def test_gpu_setup_from_fresh_reference_streams(pset, method, tmp_path):
    """boolean-serial-binary.cpp end to end: the reference generates and serialises keys, our engine loads the files
    (memory-mapped) and evaluates gates that equal the reference's scalar CPU path bit for bit."""
    import mmap

    from tfhe_gpu_b200 import BinFHEContextB200

    r = po.Ref.named(pset, method)
    r.keygen()
    bf, kf = str(tmp_path / "refreshKey.txt"), str(tmp_path / "ksKey.txt")
    r.serialize_keys(bf, kf)
    q = r.p.q
    m1 = [i & 1 for i in range(8)]
    m2 = [(i >> 1) & 1 for i in range(8)]
    c1, c2 = r.encrypt_batch(m1, 4, q), r.encrypt_batch(m2, 4, q)
    want = r.eval_bin_gate(po.GATES["NAND"], c1, c2, q)
    with open(bf, "rb") as f1, open(kf, "rb") as f2:
        mb = mmap.mmap(f1.fileno(), 0, access=mmap.ACCESS_READ)
        mk = mmap.mmap(f2.fileno(), 0, access=mmap.ACCESS_READ)
        g = BinFHEContextB200().GPUSetupFromSerialized(r.p.as_dict(), mb, mk, numGPUs=1)
    try:
        got = g.EvalBinGate("NAND", c1, c2)
        assert np.array_equal(got, want)
        assert r.decrypt_batch(got, q, 4) == [1 - (a & b) for a, b in zip(m1, m2)]
    finally:
        g.GPUClean()

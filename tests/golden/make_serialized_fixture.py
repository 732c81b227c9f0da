"""Generates the serialized-key fixtures of tests/test_serialized_keys.py with the REFERENCE itself (oracle/_ref built
from /root/reference): tiny custom parameter sets (n = 3, N = 32), keys from the reference's BTKeyGen, written with the
reference's own Serial::SerializeToFile(..., SerType::BINARY) (examples/boolean-serial-binary.cpp:76-88), next to the
flat arrays the reference's GPUSetup ordering gives (bootstrapping.cu:933-975) and reference-evaluated gate outputs.

    python tests/golden/make_serialized_fixture.py      # needs /root/reference (run in the build container)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import pyoracle as po  # noqa: E402

Q27 = 134215681
for name, method in (("ginx", po.GINX), ("ap", po.AP)):
    r = po.Ref.custom(3, 32, 32, Q27, 8, 1 << 14, 4, method)
    r.keygen()
    bkf, kskf = os.path.join(HERE, f"serialized_{name}_tiny_bk.bin"), os.path.join(HERE, f"serialized_{name}_tiny_ksk.bin")
    r.serialize_keys(bkf, kskf)
    sk, bk, ksk = r.export_keys()
    q = r.p.q
    m1 = [i & 1 for i in range(8)]
    m2 = [(i >> 1) & 1 for i in range(8)]
    c1, c2 = r.encrypt_batch(m1, 4, q), r.encrypt_batch(m2, 4, q)
    out = r.eval_bin_gate(po.GATES["NAND"], c1, c2, q)
    pd = r.p.as_dict()
    np.savez_compressed(os.path.join(HERE, f"serialized_{name}_tiny.npz"), bk=bk, ksk=ksk, sk=sk, c1=c1, c2=c2, nand=out,
                        params=np.array([pd[k] for k in sorted(pd)], dtype=np.uint64), param_names=np.array(sorted(pd)))
    print(name, os.path.getsize(bkf), os.path.getsize(kskf), pd)

"""TEST INFRASTRUCTURE ONLY -- ctypes bindings for the two CPU checkers.

* ``Port``  : oracle/libtfhe_oracle.so, our plain-C restatement (oracle/tfhe_oracle.c).
* ``Ref``   : oracle/_ref/libtfhe_ref.so, the UNMODIFIED reference CPU implementation compiled in place from
              /root/reference by oracle/Makefile (present only where it was built; it travels to the GPU box as
              a prebuilt .so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this module.
The product package (tfhe_gpu_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "libtfhe_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libtfhe_ref.so")
DROPIN_SO = os.path.join(HERE, "_ref", "libtfhe_ref_dropin.so")

# binfhe-constants.h:46-101
TOY, STD128_AP, STD128 = 0, 2, 4
AP, GINX = 1, 2
GATES = {"OR": 0, "AND": 1, "NOR": 2, "NAND": 3, "XOR_FAST": 4, "XNOR_FAST": 5, "XOR": 6, "XNOR": 7}

u64p = C.c_void_p


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _u64(a):
    return np.ascontiguousarray(a, dtype=np.uint64)


class Params(C.Structure):
    """Mirror of tfo_params (oracle/tfhe_oracle.h)."""

    _fields_ = [
        ("n", C.c_uint32), ("N", C.c_uint32),
        ("q", C.c_uint64), ("Q", C.c_uint64), ("qKS", C.c_uint64),
        ("baseKS", C.c_uint32), ("dKS", C.c_uint32),
        ("baseG", C.c_uint32), ("digitsG", C.c_uint32), ("numDigitsToThrow", C.c_uint32),
        ("baseR", C.c_uint32), ("digitsR", C.c_uint32),
        ("method", C.c_uint32), ("reserved", C.c_uint32),
        ("psi", C.c_uint64), ("beta", C.c_uint64),
    ]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}

    @property
    def d(self):
        if self.method == GINX:
            return 2 * (self.digitsG - self.numDigitsToThrow)
        return 2 * self.digitsG


def build_port():
    subprocess.check_call(["make", "-s", "-C", HERE, "port"])


def have_ref():
    return os.path.exists(REF_SO)


class Port:
    """Our C restatement of the reference CPU path."""

    def __init__(self, params: Params):
        if not os.path.exists(PORT_SO):
            build_port()
        L = C.CDLL(PORT_SO)
        self.L = L
        L.tfo_ctx_new.restype = C.c_void_p
        L.tfo_ctx_new.argtypes = [C.c_void_p]
        L.tfo_ctx_free.argtypes = [C.c_void_p]
        L.tfo_bk_words.restype = C.c_size_t
        L.tfo_bk_words.argtypes = [C.c_void_p]
        L.tfo_ksk_words.restype = C.c_size_t
        L.tfo_ksk_words.argtypes = [C.c_void_p]
        L.tfo_round_qQ.restype = C.c_uint64
        L.tfo_round_qQ.argtypes = [C.c_uint64] * 3
        L.tfo_decrypt.restype = C.c_int64
        self.p = params
        self.h = C.c_void_p(L.tfo_ctx_new(C.byref(params)))
        self.n, self.N = params.n, params.N

    def __del__(self):
        try:
            self.L.tfo_ctx_free(self.h)
        except Exception:
            pass

    # ---- parameters ----
    @staticmethod
    def _lib():
        if not os.path.exists(PORT_SO):
            build_port()
        return C.CDLL(PORT_SO)

    @staticmethod
    def params_named(pset, method):
        p = Params()
        assert Port._lib().tfo_params_named(C.c_int(pset), C.c_int(method), C.byref(p)) == 0
        return p

    @staticmethod
    def params_func(pset, arb, logQ, N=0, baseG=0, throw=0):
        p = Params()
        assert Port._lib().tfo_params_func(C.c_int(pset), C.c_int(int(arb)), C.c_uint32(logQ), C.c_uint64(N),
                                           C.c_uint32(baseG), C.c_uint32(throw), C.byref(p)) == 0
        return p

    @staticmethod
    def params_custom(n, N, q, Q, baseKS, baseG, baseR, method):
        p = Params()
        assert Port._lib().tfo_params_custom(C.c_uint32(n), C.c_uint32(N), C.c_uint64(q), C.c_uint64(Q),
                                             C.c_uint32(baseKS), C.c_uint32(baseG), C.c_uint32(baseR),
                                             C.c_int(method), C.byref(p)) == 0
        return p

    def bk_words(self):
        return self.L.tfo_bk_words(C.byref(self.p))

    def ksk_words(self):
        return self.L.tfo_ksk_words(C.byref(self.p))

    # ---- keys ----
    def keygen(self, seed):
        sk = np.zeros(self.n, dtype=np.uint64)
        bk = np.zeros(self.bk_words(), dtype=np.uint64)
        ksk = np.zeros(self.ksk_words(), dtype=np.uint64)
        self.L.tfo_keygen(self.h, C.c_uint64(seed), _p(sk), _p(bk), _p(ksk))
        return sk, bk, ksk

    def encrypt(self, sk, m, p, mod, seed):
        ct = np.zeros(self.n + 1, dtype=np.uint64)
        self.L.tfo_encrypt(self.h, _p(sk), C.c_int64(int(m)), C.c_uint64(p), C.c_uint64(mod), C.c_uint64(seed), _p(ct))
        return ct

    def encrypt_batch(self, sk, msgs, p, mod, seed):
        return np.stack([self.encrypt(sk, m, p, mod, seed * 1000003 + i) for i, m in enumerate(msgs)])

    def decrypt(self, sk, ct, mod, p):
        ct = _u64(ct)
        return int(self.L.tfo_decrypt(self.h, _p(sk), _p(ct), C.c_uint64(mod), C.c_uint64(p)))

    def decrypt_batch(self, sk, cts, mod, p):
        return [self.decrypt(sk, c, mod, p) for c in _u64(cts)]

    # ---- stages ----
    def ntt(self, poly, inverse=False):
        a = _u64(poly).copy()
        (self.L.tfo_ntt_inverse if inverse else self.L.tfo_ntt_forward)(self.h, _p(a))
        return a

    def signed_digit_decompose(self, x):
        x = _u64(x)
        out = np.zeros((self.p.d, self.N), dtype=np.uint64)
        self.L.tfo_signed_digit_decompose(self.h, _p(x), _p(out))
        return out

    def eval_acc(self, bk, a, mod, acc):
        a = _u64(a)
        acc = _u64(acc).copy()
        self.L.tfo_eval_acc(self.h, _p(bk), C.c_int(a.shape[0]), _p(a), C.c_uint64(mod), _p(acc))
        return acc

    def round_qQ(self, v, q, Q):
        return int(self.L.tfo_round_qQ(v, q, Q))

    def mod_switch(self, x, from_mod, to_mod):
        x = _u64(x)
        out = np.zeros_like(x)
        self.L.tfo_mod_switch(C.c_int(x.shape[0]), C.c_size_t(x.shape[1]), _p(x), C.c_uint64(from_mod),
                              C.c_uint64(to_mod), _p(out))
        return out

    def key_switch(self, ksk, x):
        x = _u64(x)
        out = np.zeros((x.shape[0], self.n + 1), dtype=np.uint64)
        self.L.tfo_key_switch(self.h, _p(ksk), C.c_int(x.shape[0]), _p(x), _p(out))
        return out

    def mkmswitch(self, ksk, x, fmod):
        x = _u64(x)
        out = np.zeros((x.shape[0], self.n + 1), dtype=np.uint64)
        self.L.tfo_mkmswitch(self.h, _p(ksk), C.c_int(x.shape[0]), _p(x), C.c_uint64(fmod), _p(out))
        return out

    def init_acc_gate(self, gate, b, ctmod):
        acc = np.zeros((2, self.N), dtype=np.uint64)
        self.L.tfo_init_acc_gate(self.h, C.c_int(gate), C.c_uint64(int(b)), C.c_uint64(ctmod), _p(acc))
        return acc

    # ---- batched ops ----
    def eval_bin_gate(self, bk, ksk, gate, ct1, ct2, mod):
        ct1, ct2 = _u64(ct1), _u64(ct2)
        out = np.zeros_like(ct1)
        rc = self.L.tfo_eval_bin_gate(self.h, _p(bk), _p(ksk), C.c_int(gate), C.c_int(ct1.shape[0]), _p(ct1), _p(ct2),
                                      C.c_uint64(mod), _p(out))
        assert rc == 0
        return out

    def bootstrap_func(self, bk, ksk, ct, ctmod, table, fmod):
        ct, table = _u64(ct), _u64(table)
        out = np.zeros_like(ct)
        rc = self.L.tfo_bootstrap_func(self.h, _p(bk), _p(ksk), C.c_int(ct.shape[0]), _p(ct), C.c_uint64(ctmod),
                                       _p(table), C.c_int(int(table.ndim == 2)), C.c_uint64(fmod), _p(out))
        assert rc == 0
        return out

    def eval_func(self, bk, ksk, ct, mod, lut):
        ct, lut = _u64(ct), _u64(lut)
        out = np.zeros_like(ct)
        per_ct = int(lut.ndim == 2)
        rc = self.L.tfo_eval_func(self.h, _p(bk), _p(ksk), C.c_int(ct.shape[0]), _p(ct), C.c_uint64(mod), _p(lut),
                                  C.c_size_t(lut.shape[-1]), C.c_int(per_ct), _p(out))
        assert rc == 0, rc
        return out

    def eval_floor(self, bk, ksk, ct, mod, roundbits=0):
        ct = _u64(ct)
        out = np.zeros_like(ct)
        rc = self.L.tfo_eval_floor(self.h, _p(bk), _p(ksk), C.c_int(ct.shape[0]), _p(ct), C.c_uint64(mod),
                                   C.c_uint32(roundbits), _p(out))
        assert rc == 0
        return out

    def eval_sign(self, bk, ksk, ct, mod):
        ct = _u64(ct)
        out = np.zeros_like(ct)
        rc = self.L.tfo_eval_sign(self.h, _p(bk), _p(ksk), C.c_int(ct.shape[0]), _p(ct), C.c_uint64(mod), _p(out))
        assert rc == 0
        return out

    def eval_decomp(self, bk, ksk, ct, mod, max_digits=8):
        ct = _u64(ct)
        out = np.zeros((ct.shape[0], max_digits, self.n + 1), dtype=np.uint64)
        mods = np.zeros(max_digits, dtype=np.uint64)
        nd = self.L.tfo_eval_decomp(self.h, _p(bk), _p(ksk), C.c_int(ct.shape[0]), _p(ct), C.c_uint64(mod),
                                    C.c_int(max_digits), _p(out), _p(mods))
        assert nd > 0
        return out[:, :nd].copy(), [int(m) for m in mods[:nd]]

    @staticmethod
    def _key_arrays(ports, bks, ksks):
        nk = len(ports)
        cs = (C.c_void_p * nk)(*[pt.h for pt in ports])
        b = (C.c_void_p * nk)(*[x.ctypes.data for x in bks])
        k = (C.c_void_p * nk)(*[x.ctypes.data for x in ksks])
        return nk, cs, b, k

    @staticmethod
    def eval_sign_dyn(ports, bks, ksks, ct, mod):
        """Dynamic gadget base: ports[0] is the context's own key set (tfo_eval_sign_dyn)."""
        ct = _u64(ct)
        out = np.zeros_like(ct)
        nk, cs, b, k = Port._key_arrays(ports, bks, ksks)
        rc = ports[0].L.tfo_eval_sign_dyn(C.c_int(nk), cs, b, k, C.c_int(ct.shape[0]), _p(ct), C.c_uint64(mod), _p(out))
        assert rc == 0
        return out

    @staticmethod
    def eval_decomp_dyn(ports, bks, ksks, ct, mod, max_digits=8):
        ct = _u64(ct)
        out = np.zeros((ct.shape[0], max_digits, ports[0].n + 1), dtype=np.uint64)
        mods = np.zeros(max_digits, dtype=np.uint64)
        nk, cs, b, k = Port._key_arrays(ports, bks, ksks)
        nd = ports[0].L.tfo_eval_decomp_dyn(C.c_int(nk), cs, b, k, C.c_int(ct.shape[0]), _p(ct), C.c_uint64(mod),
                                            C.c_int(max_digits), _p(out), _p(mods))
        assert nd > 0
        return out[:, :nd].copy(), [int(m) for m in mods[:nd]]

    def mul_matrix(self, ct, M, modulus):
        ct = _u64(ct)
        M = np.ascontiguousarray(M, dtype=np.int64)
        out = np.zeros((M.shape[1], self.n + 1), dtype=np.uint64)
        rc = self.L.tfo_mul_matrix(self.h, C.c_int(M.shape[0]), C.c_int(M.shape[1]), _p(ct), _p(M),
                                   C.c_uint64(modulus), _p(out))
        assert rc == 0
        return out

    def num_threads(self):
        return int(self.L.tfo_num_threads())

    def set_num_threads(self, n):
        self.L.tfo_set_num_threads(C.c_int(int(n)))


class Ref:
    """The unmodified reference CPU implementation (scalar API looped with OpenMP)."""

    def __init__(self, kind, args, so=None):
        so = so or REF_SO
        L = C.CDLL(so)
        self.L = L
        L.ref_ctx_create.restype = C.c_void_p
        L.ref_ctx_create.argtypes = [C.c_int, C.c_void_p]
        L.ref_last_error.restype = C.c_char_p
        L.ref_bk_words.restype = C.c_uint64
        L.ref_ksk_words.restype = C.c_uint64
        L.ref_decrypt.restype = C.c_int64
        for f in ("ref_bk_words", "ref_ksk_words", "ref_keygen", "ref_ctx_destroy"):
            getattr(L, f).argtypes = [C.c_void_p]
        a = np.zeros(8, dtype=np.uint64)
        a[: len(args)] = args
        h = L.ref_ctx_create(kind, _p(a))
        if not h:
            raise RuntimeError(L.ref_last_error().decode())
        self.h = C.c_void_p(h)
        out = np.zeros(16, dtype=np.uint64)
        self._chk(L.ref_ctx_params(self.h, _p(out)))
        p = Params()
        (p.n, p.N, p.q, p.Q, p.qKS, p.baseKS, p.dKS, p.baseG, p.digitsG, p.numDigitsToThrow, p.baseR, p.digitsR,
         p.method, p.psi, p.beta) = [int(x) for x in out[:15]]
        self.p = p
        self.n, self.N = p.n, p.N

    @staticmethod
    def named(pset, method, so=None):
        return Ref(0, [pset, method], so)

    @staticmethod
    def func(pset, arb, logQ, N=0, baseG=0, throw=0, so=None):
        return Ref(1, [pset, int(arb), logQ, N, baseG, throw], so)

    @staticmethod
    def custom(n, N, q, Q, baseKS, baseG, baseR, method, so=None):
        return Ref(2, [n, N, q, Q, baseKS, baseG, baseR, method], so)

    @staticmethod
    def func_dynamic(pset, arb, logQ, N=0, so=None):
        """GenerateBinFHEContext(set, arbFunc, logQ, N, GINX, timeOptimization=true): BTKeyGen fills the three-key map."""
        return Ref(3, [pset, int(arb), logQ, N, 0, 0], so)

    def key_map_bases(self):
        out = np.zeros(8, dtype=np.uint64)
        cnt = self._chk(self.L.ref_key_map_bases(self.h, _p(out), C.c_int(8)))
        return [int(x) for x in out[:cnt]]

    def export_key_map(self):
        """{baseG: (Params, bk, ksk)} for every key set of m_BTKey_map; leaves the context on its own base."""
        own = int(self.p.baseG)
        res = {}
        try:
            for base in self.key_map_bases():
                self._chk(self.L.ref_select_key(self.h, C.c_uint64(base)))
                out = np.zeros(16, dtype=np.uint64)
                self._chk(self.L.ref_ctx_params(self.h, _p(out)))
                p = Params()
                (p.n, p.N, p.q, p.Q, p.qKS, p.baseKS, p.dKS, p.baseG, p.digitsG, p.numDigitsToThrow, p.baseR,
                 p.digitsR, p.method, p.psi, p.beta) = [int(x) for x in out[:15]]
                bk = np.zeros(self.L.ref_bk_words(self.h), dtype=np.uint64)
                ksk = np.zeros(self.L.ref_ksk_words(self.h), dtype=np.uint64)
                self._chk(self.L.ref_export_bk(self.h, _p(bk)))
                self._chk(self.L.ref_export_ksk(self.h, _p(ksk)))
                res[base] = (p, bk, ksk)
        finally:
            self._chk(self.L.ref_select_key(self.h, C.c_uint64(own)))
        return res

    def _chk(self, rc):
        if rc < 0:
            raise RuntimeError(self.L.ref_last_error().decode())
        return rc

    def keygen(self):
        self._chk(self.L.ref_keygen(self.h))

    def export_keys(self):
        sk = np.zeros(self.n, dtype=np.uint64)
        bk = np.zeros(self.L.ref_bk_words(self.h), dtype=np.uint64)
        ksk = np.zeros(self.L.ref_ksk_words(self.h), dtype=np.uint64)
        self._chk(self.L.ref_export_sk(self.h, _p(sk)))
        self._chk(self.L.ref_export_bk(self.h, _p(bk)))
        self._chk(self.L.ref_export_ksk(self.h, _p(ksk)))
        return sk, bk, ksk

    def serialize_keys(self, bk_path, ksk_path):
        """Serial::SerializeToFile(path, cc.GetRefreshKey() / cc.GetSwitchKey(), SerType::BINARY), the reference's own
        serialization (examples/boolean-serial-binary.cpp:76-88)."""
        self._chk(self.L.ref_serialize_keys(self.h, bk_path.encode(), ksk_path.encode()))

    def encrypt(self, m, p, mod):
        ct = np.zeros(self.n + 1, dtype=np.uint64)
        self._chk(self.L.ref_encrypt(self.h, C.c_int64(int(m)), C.c_uint64(p), C.c_uint64(mod), _p(ct)))
        return ct

    def encrypt_batch(self, msgs, p, mod):
        return np.stack([self.encrypt(m, p, mod) for m in msgs])

    def decrypt(self, ct, mod, p):
        ct = _u64(ct)
        return int(self.L.ref_decrypt(self.h, _p(ct), C.c_uint64(mod), C.c_uint64(p)))

    def decrypt_batch(self, cts, mod, p):
        return [self.decrypt(c, mod, p) for c in _u64(cts)]

    def _op1(self, fn, ct, mod, *extra):
        ct = _u64(ct)
        out = np.zeros_like(ct)
        self._chk(fn(self.h, C.c_int(ct.shape[0]), _p(ct), C.c_uint64(mod), *extra, _p(out)))
        return out

    def eval_bin_gate(self, gate, ct1, ct2, mod, batched=False):
        ct1, ct2 = _u64(ct1), _u64(ct2)
        out = np.zeros_like(ct1)
        fn = self.L.ref_batched_eval_bin_gate if batched else self.L.ref_eval_bin_gate
        self._chk(fn(self.h, C.c_int(gate), C.c_int(ct1.shape[0]), _p(ct1), _p(ct2), C.c_uint64(mod), _p(out)))
        return out

    def eval_func(self, ct, mod, lut, batched=False):
        lut = _u64(lut)
        fn = self.L.ref_batched_eval_func if batched else self.L.ref_eval_func
        return self._op1(fn, ct, mod, _p(lut), C.c_uint64(lut.shape[0]))

    def eval_floor(self, ct, mod, roundbits=0, batched=False):
        fn = self.L.ref_batched_eval_floor if batched else self.L.ref_eval_floor
        return self._op1(fn, ct, mod, C.c_uint32(roundbits))

    def eval_sign(self, ct, mod, batched=False):
        fn = self.L.ref_batched_eval_sign if batched else self.L.ref_eval_sign
        return self._op1(fn, ct, mod)

    def eval_decomp(self, ct, mod, max_digits=8, batched=False):
        ct = _u64(ct)
        out = np.zeros((ct.shape[0], max_digits, self.n + 1), dtype=np.uint64)
        mods = np.zeros(max_digits, dtype=np.uint64)
        fn = self.L.ref_batched_eval_decomp if batched else self.L.ref_eval_decomp
        nd = self._chk(fn(self.h, C.c_int(ct.shape[0]), _p(ct), C.c_uint64(mod), C.c_int(max_digits), _p(out),
                          _p(mods)))
        return out[:, :nd].copy(), [int(m) for m in mods[:nd]]

    def mul_matrix(self, ct, mod, M, modulus):
        ct = _u64(ct)
        M = np.ascontiguousarray(M, dtype=np.int64)
        out = np.zeros((M.shape[1], self.n + 1), dtype=np.uint64)
        self._chk(self.L.ref_batched_mul_matrix(self.h, C.c_int(M.shape[0]), C.c_int(M.shape[1]), _p(ct),
                                                C.c_uint64(mod), _p(M), C.c_uint64(modulus), _p(out)))
        return out

    # ---- stages ----
    def ntt(self, poly, inverse=False):
        a = _u64(poly).copy()
        self._chk(self.L.ref_ntt(self.h, C.c_int(int(inverse)), _p(a)))
        return a

    def signed_digit_decompose(self, x):
        x = _u64(x)
        out = np.zeros((self.p.d, self.N), dtype=np.uint64)
        self._chk(self.L.ref_signed_digit_decompose(self.h, _p(x), _p(out)))
        return out

    def eval_acc(self, a, mod, acc):
        a = _u64(a)
        acc = _u64(acc).copy()
        self._chk(self.L.ref_eval_acc(self.h, C.c_int(a.shape[0]), _p(a), C.c_uint64(mod), _p(acc)))
        return acc

    def mod_switch(self, x, from_mod, to_mod):
        x = _u64(x)
        out = np.zeros_like(x)
        self._chk(self.L.ref_mod_switch(self.h, C.c_int(x.shape[0]), C.c_uint64(x.shape[1]), _p(x),
                                        C.c_uint64(from_mod), C.c_uint64(to_mod), _p(out)))
        return out

    def key_switch(self, x):
        x = _u64(x)
        out = np.zeros((x.shape[0], self.n + 1), dtype=np.uint64)
        self._chk(self.L.ref_key_switch(self.h, C.c_int(x.shape[0]), _p(x), _p(out)))
        return out

    # ---- fused C++ adapter (tfhe_gpu_b200/adapter/binfhe_b200.hpp; only in the drop-in build) ----
    def fused_create(self, num_gpus=1):
        self.L.fused_create.restype = C.c_void_p
        self.L.fused_create.argtypes = [C.c_void_p, C.c_int]
        f = self.L.fused_create(self.h, num_gpus)
        if not f:
            raise RuntimeError(self.L.ref_last_error().decode())
        self._fused = C.c_void_p(f)

    def fused_destroy(self):
        self.L.fused_destroy.argtypes = [C.c_void_p]
        self.L.fused_destroy(self._fused)
        self._fused = None

    def fused_eval_bin_gate(self, gate, ct1, ct2, mod):
        ct1, ct2 = _u64(ct1), _u64(ct2)
        out = np.zeros_like(ct1)
        self._chk(self.L.fused_eval_bin_gate(self.h, self._fused, C.c_int(gate), C.c_int(ct1.shape[0]), _p(ct1), _p(ct2),
                                             C.c_uint64(mod), _p(out)))
        return out

    def fused_bench_eval_bin_gate(self, gate, ct1, ct2, mod, reps):
        """Times `reps` calls of the C++ adapter's EvalBinGate on std::vector<LWECiphertext> built once from the flat
        arrays; returns (per-call seconds, result of the last call)."""
        ct1, ct2 = _u64(ct1), _u64(ct2)
        out = np.zeros_like(ct1)
        secs = np.zeros(reps, dtype=np.float64)
        self._chk(self.L.fused_bench_eval_bin_gate(self.h, self._fused, C.c_int(gate), C.c_int(ct1.shape[0]), _p(ct1),
                                                   _p(ct2), C.c_uint64(mod), C.c_int(reps), _p(secs), _p(out)))
        return secs, out

    def fused_eval_func(self, ct, mod, lut):
        ct, lut = _u64(ct), _u64(lut)
        out = np.zeros_like(ct)
        self._chk(self.L.fused_eval_func(self.h, self._fused, C.c_int(ct.shape[0]), _p(ct), C.c_uint64(mod), _p(lut),
                                         C.c_uint64(lut.shape[0]), _p(out)))
        return out

    def fused_eval_sign(self, ct, mod):
        ct = _u64(ct)
        out = np.zeros_like(ct)
        self._chk(self.L.fused_eval_sign(self.h, self._fused, C.c_int(ct.shape[0]), _p(ct), C.c_uint64(mod), _p(out)))
        return out

    def fused_eval_decomp(self, ct, mod, max_digits=8):
        ct = _u64(ct)
        out = np.zeros((ct.shape[0], max_digits, self.n + 1), dtype=np.uint64)
        mods = np.zeros(max_digits, dtype=np.uint64)
        nd = self._chk(self.L.fused_eval_decomp(self.h, self._fused, C.c_int(ct.shape[0]), _p(ct), C.c_uint64(mod),
                                                C.c_int(max_digits), _p(out), _p(mods)))
        return out[:, :nd].copy(), [int(m) for m in mods[:nd]]

    def gpu_setup(self, num_gpus=1):
        self._chk(self.L.ref_gpu_setup(self.h, C.c_int(num_gpus)))

    def gpu_clean(self):
        self._chk(self.L.ref_gpu_clean(self.h))

    def num_threads(self):
        return int(self.L.ref_num_threads())

    def set_num_threads(self, n):
        self.L.ref_set_num_threads(C.c_int(int(n)))

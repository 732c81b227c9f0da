"""ctypes mirror of the reference's batched BinFHEContext API (binfhecontext.h:352-421) on top of libtfhe_b200.so.

Method names, argument meaning and error behaviour follow the reference:

=====================================  ==========================================================
reference (binfhecontext.cpp)          here
=====================================  ==========================================================
``cc.GPUSetup(numGPUs)``       :349    ``BinFHEContextB200.GPUSetup(params, bk, ksk, numGPUs)``
``cc.GPUClean()``              :362    ``GPUClean()``
``cc.EvalBinGate(g, v1, v2)``  :323    ``EvalBinGate(gate, ct1, ct2)``
``cc.EvalFunc(v, LUT)``        :327    ``EvalFunc(ct, lut)``        (2-D lut => LUT_vec overload :332)
``cc.EvalFloor(v, roundbits)`` :337    ``EvalFloor(ct, ct_mod, roundbits)``
``cc.EvalSign(v)``             :341    ``EvalSign(ct, ct_mod)``
``cc.EvalDecomp(v)``           :345    ``EvalDecomp(ct, ct_mod)``
``cc.CiphertextMulMatrix``     :319    ``CiphertextMulMatrix(ct, matrix, modulus)``
=====================================  ==========================================================

Ciphertext batches are ``uint64`` arrays of shape ``[batch, n+1]`` (``a`` then ``b``) -- numpy arrays (host) or
CUDA ``torch`` tensors (device resident; nothing is copied through the host).  API misuse raises
``TfheB200Error`` with the reference's message (the reference throws ``openfhe_error``).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

GATES = {"OR": 0, "AND": 1, "NOR": 2, "NAND": 3, "XOR_FAST": 4, "XNOR_FAST": 5, "XOR": 6, "XNOR": 7}
NOT_GATE = 8   # TFHE_B200_NOT (netlists only)
AP, GINX = 1, 2
HOST, DEVICE = 0, 1


class TfheB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[tfhe_b200 status {code}] {msg}")
        self.code = code
        self.msg = msg


class Params(C.Structure):
    """tfhe_b200_params (include/tfhe_b200.h)."""

    _fields_ = [
        ("n", C.c_uint32), ("N", C.c_uint32),
        ("q", C.c_uint64), ("Q", C.c_uint64), ("qKS", C.c_uint64),
        ("baseKS", C.c_uint32), ("dKS", C.c_uint32),
        ("baseG", C.c_uint32), ("digitsG", C.c_uint32), ("numDigitsToThrow", C.c_uint32),
        ("baseR", C.c_uint32), ("digitsR", C.c_uint32),
        ("method", C.c_uint32), ("flags", C.c_uint32),
        ("psi", C.c_uint64), ("beta", C.c_uint64),
    ]

    @classmethod
    def from_dict(cls, d):
        p = cls()
        for k, _ in cls._fields_:
            if k in d:
                setattr(p, k, int(d[k]))
        return p

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


class Stats(C.Structure):
    """tfhe_b200_stats."""

    _fields_ = [
        ("h2d_ms", C.c_float), ("prep_ms", C.c_float), ("blind_rotate_ms", C.c_float), ("keyswitch_ms", C.c_float),
        ("d2h_ms", C.c_float), ("total_ms", C.c_float), ("bootstraps", C.c_uint32), ("kernel_launches", C.c_uint32),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


def lib_path():
    # TFHE_B200_LIB: developer knob to load an experimental build of the same library (kernel A/B measurements)
    return os.environ.get("TFHE_B200_LIB") or os.path.join(_HERE, "libtfhe_b200.so")


_LIB = None

EXPORTS = [
    "tfhe_b200_setup", "tfhe_b200_clean", "tfhe_b200_last_error", "tfhe_b200_num_gpus", "tfhe_b200_bk_words",
    "tfhe_b200_ksk_words", "tfhe_b200_kernel_variant", "tfhe_b200_set_option", "tfhe_b200_eval_acc",
    "tfhe_b200_mkmswitch", "tfhe_b200_mul_matrix", "tfhe_b200_eval_bin_gate", "tfhe_b200_bootstrap_func",
    "tfhe_b200_eval_func", "tfhe_b200_eval_floor", "tfhe_b200_eval_sign", "tfhe_b200_eval_decomp",
    "tfhe_b200_eval_bin_gate_v", "tfhe_b200_eval_circuit", "tfhe_b200_keygen", "tfhe_b200_keygen_test_seed", "tfhe_b200_setup_from_serialized",
    "tfhe_b200_serialized_info", "tfhe_b200_flatten_serialized", "tfhe_b200_add_key_set", "tfhe_b200_num_key_sets",
    "tfhe_b200_persistent_plan",
]


def load_library():
    """Load libtfhe_b200.so; fails loudly when it has not been built (there is no fallback path)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise TfheB200Error(-2, f"{path} is missing: build it with `make -C tfhe_gpu_b200` "
                                "(or __graft_entry__.build()); this engine has no CPU fallback")
    L = C.CDLL(path)
    L.tfhe_b200_last_error.restype = C.c_char_p
    L.tfhe_b200_kernel_variant.restype = C.c_char_p
    L.tfhe_b200_kernel_variant.argtypes = [C.c_void_p]
    L.tfhe_b200_bk_words.restype = C.c_size_t
    L.tfhe_b200_bk_words.argtypes = [C.c_void_p]
    L.tfhe_b200_ksk_words.restype = C.c_size_t
    L.tfhe_b200_ksk_words.argtypes = [C.c_void_p]
    L.tfhe_b200_num_gpus.argtypes = [C.c_void_p]
    L.tfhe_b200_clean.argtypes = [C.c_void_p]
    L.tfhe_b200_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_int64]
    L.tfhe_b200_setup.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int, C.c_int,
                                  C.c_int, C.c_void_p]
    L.tfhe_b200_add_key_set.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                        C.c_int]
    L.tfhe_b200_num_key_sets.argtypes = [C.c_void_p]
    _LIB = L
    return L


def _is_torch(x):
    return type(x).__module__.startswith("torch")


class _Buf:
    """Uniform view (pointer, space, shape) over a numpy array or a CUDA torch tensor of 64-bit integers.

    Device tensors: the library runs on its own non-blocking streams and knows nothing about the producer's stream, so
    the tensor's device is synchronised here before its pointer crosses the C ABI (which is synchronous anyway), and
    `device` records where it lives so the caller can check it against the handle's first GPU."""

    def __init__(self, x, dtype=np.uint64):
        self.device = None
        if _is_torch(x):
            import torch

            if not x.is_cuda:
                x = x.numpy()
            else:
                assert x.dtype in (torch.int64, torch.uint64), "device tensors must be 64-bit integers"
                x = x.contiguous()
                torch.cuda.current_stream(x.device).synchronize()
                self.device = x.device.index
                self.obj, self.ptr, self.space, self.shape = x, C.c_void_p(x.data_ptr()), DEVICE, tuple(x.shape)
                return
        a = np.ascontiguousarray(x, dtype=dtype)
        self.obj, self.ptr, self.space, self.shape = a, C.c_void_p(a.ctypes.data), HOST, a.shape

    def empty_like_out(self, shape):
        if self.space == DEVICE:
            import torch

            return torch.empty(shape, dtype=self.obj.dtype, device=self.obj.device)
        return np.empty(shape, dtype=np.uint64)


def gpu_keygen(params, sk_lwe, sk_ring, key=None, device=0, seed=None):
    """tfhe_b200_keygen: evaluation keys generated on the GPU.  `sk_lwe` / `sk_ring` are
    the ternary secrets as signed values in {-1, 0, 1}.  `key`: 32 bytes of generator key material from a CSPRNG, or
    None (the engine draws them from the operating system).  An `int` selects the TEST-ONLY deterministic generator
    (tfhe_b200_keygen_test_seed: at most 64 bits of security, for fixtures and statistics tests).  Returns (bk, ksk) as
    int64 CUDA tensors in the layout GPUSetup takes -- the keys never exist on the host."""
    import torch

    if seed is not None:           # keyword form of the TEST-ONLY seeded generator
        key = int(seed)
    L = load_library()
    p = params if isinstance(params, Params) else Params.from_dict(
        params.as_dict() if hasattr(params, "as_dict") else params)
    s1 = np.ascontiguousarray(sk_lwe, dtype=np.int8)
    s2 = np.ascontiguousarray(sk_ring, dtype=np.int8)
    if s1.shape != (p.n,) or s2.shape != (p.N,):
        raise TfheB200Error(-1, "KeyGen: secret key lengths must be n and N")
    dev = torch.device("cuda", device)
    bk = torch.empty(int(L.tfhe_b200_bk_words(C.byref(p))), dtype=torch.int64, device=dev)
    ksk = torch.empty(int(L.tfhe_b200_ksk_words(C.byref(p))), dtype=torch.int64, device=dev)
    common = (C.byref(p), C.c_void_p(s1.ctypes.data), C.c_void_p(s2.ctypes.data))
    tail = (device, C.c_void_p(bk.data_ptr()), C.c_void_p(ksk.data_ptr()))
    if isinstance(key, (int, np.integer)):
        rc = L.tfhe_b200_keygen_test_seed(*common, C.c_uint64(int(key)), *tail)
    else:
        if key is not None and len(key) != 32:
            raise TfheB200Error(-1, "KeyGen: key must be 32 bytes")
        kb = (C.c_uint8 * 32).from_buffer_copy(bytes(key)) if key is not None else None
        rc = L.tfhe_b200_keygen(*common, kb, *tail)
    if rc != 0:
        raise TfheB200Error(rc, L.tfhe_b200_last_error().decode())
    return bk, ksk


class SerializedInfo(C.Structure):
    """tfhe_b200_serialized_info_t."""

    _fields_ = [("bk_dim", C.c_uint64 * 3), ("bk_rows", C.c_uint64), ("N", C.c_uint64), ("Q", C.c_uint64),
                ("psi", C.c_uint64), ("ks_N", C.c_uint64), ("baseKS", C.c_uint64), ("dKS", C.c_uint64),
                ("n", C.c_uint64), ("qKS", C.c_uint64), ("bk_words", C.c_size_t), ("ksk_words", C.c_size_t)]


def _stream(x):
    """bytes / bytearray / mmap / uint8 numpy array -> (pointer, length, keep-alive object)."""
    a = np.frombuffer(x, dtype=np.uint8) if not isinstance(x, np.ndarray) else np.ascontiguousarray(x, dtype=np.uint8)
    return C.c_void_p(a.ctypes.data), C.c_size_t(a.size), a


def serialized_info(bk_stream, ksk_stream):
    """What OpenFHE's serialized refreshing / switching key streams hold (tfhe_b200_serialized_info; host only)."""
    L = load_library()
    bp, bn, _b = _stream(bk_stream)
    kp, kn, _k = _stream(ksk_stream)
    info = SerializedInfo()
    rc = L.tfhe_b200_serialized_info(bp, bn, kp, kn, C.byref(info))
    if rc != 0:
        raise TfheB200Error(rc, L.tfhe_b200_last_error().decode())
    return info


def flatten_serialized(bk_stream, ksk_stream):
    """The serialized streams as the flat uint64 arrays GPUSetup takes (tfhe_b200_flatten_serialized; host only)."""
    L = load_library()
    info = serialized_info(bk_stream, ksk_stream)
    bp, bn, _b = _stream(bk_stream)
    kp, kn, _k = _stream(ksk_stream)
    bk = np.empty(info.bk_words, dtype=np.uint64)
    ksk = np.empty(info.ksk_words, dtype=np.uint64)
    rc = L.tfhe_b200_flatten_serialized(bp, bn, kp, kn, C.c_void_p(bk.ctypes.data), C.c_size_t(bk.size),
                                        C.c_void_p(ksk.ctypes.data), C.c_size_t(ksk.size))
    if rc != 0:
        raise TfheB200Error(rc, L.tfhe_b200_last_error().decode())
    return bk, ksk


class BinFHEContextB200:
    """GPU half of the reference's BinFHEContext.  Keys come in as the flat arrays the reference's own GPUSetup
    flattens them to (bootstrapping.cu:933-975); see include/tfhe_b200.h for the element order."""

    def __init__(self):
        self._h = None
        self.params = None
        self.first_device = 0
        self.last_stats = Stats()

    # ---- lifetime -------------------------------------------------------------------------------------
    def GPUSetup(self, params, bk, ksk, numGPUs=0, first_device=0, keep_generic=None):
        """binfhecontext.cpp:349-360.  `bk`/`ksk` may be numpy arrays or CUDA tensors (e.g. NCCL-broadcast; they must
        live on `first_device`).  `keep_generic` keeps the generic-layout key beside the specialised one so that
        set_option("force_generic") works (TFHE_B200_FLAG_KEEP_GENERIC; default: environment TFHE_B200_KEEP_GENERIC)."""
        L = load_library()
        if self._h is not None:
            self.GPUClean()  # idempotent re-setup (the reference appends and double-counts GPUs: bootstrapping.cu:762)
        if bk is None or ksk is None:
            raise TfheB200Error(-1, "ERROR: Need to call BTKeyGen before calling GPUSetup")
        p = params if isinstance(params, Params) else Params.from_dict(
            params.as_dict() if hasattr(params, "as_dict") else params)
        if keep_generic is not None:
            p = Params.from_dict(p.as_dict())
            p.flags = (p.flags | 1) if keep_generic else (p.flags & ~1)
        b, k = _Buf(bk), _Buf(ksk)
        if b.space != k.space:
            raise TfheB200Error(-1, "GPUSetup: bk and ksk must live in the same memory space")
        for t in (b, k):
            if t.space == DEVICE and t.device != first_device:
                raise TfheB200Error(-1, f"GPUSetup: device-resident keys must live on the first GPU of the handle "
                                        f"(cuda:{first_device}), got cuda:{t.device}")
        h = C.c_void_p()
        rc = L.tfhe_b200_setup(C.byref(p), b.ptr, C.c_size_t(int(np.prod(b.shape))), k.ptr,
                               C.c_size_t(int(np.prod(k.shape))), b.space, first_device, numGPUs, C.byref(h))
        if rc != 0:
            raise TfheB200Error(rc, L.tfhe_b200_last_error().decode())
        self._h, self.params, self.first_device = h, p, first_device
        return self

    def GPUSetupFromSerialized(self, params, bk_stream, ksk_stream, numGPUs=0, first_device=0, keep_generic=None):
        """GPUSetup from the byte streams OpenFHE's Serial::SerializeToFile(..., SerType::BINARY) writes for
        cc.GetRefreshKey() / cc.GetSwitchKey() (examples/boolean-serial-binary.cpp:76-88): bytes, mmap objects or uint8
        arrays.  No flat copy of the keys is made on the host (tfhe_b200_setup_from_serialized)."""
        L = load_library()
        if self._h is not None:
            self.GPUClean()
        p = params if isinstance(params, Params) else Params.from_dict(
            params.as_dict() if hasattr(params, "as_dict") else params)
        if keep_generic is not None:
            p = Params.from_dict(p.as_dict())
            p.flags = (p.flags | 1) if keep_generic else (p.flags & ~1)
        bp, bn, _b = _stream(bk_stream)
        kp, kn, _k = _stream(ksk_stream)
        h = C.c_void_p()
        rc = L.tfhe_b200_setup_from_serialized(C.byref(p), bp, bn, kp, kn, first_device, numGPUs, C.byref(h))
        if rc != 0:
            raise TfheB200Error(rc, L.tfhe_b200_last_error().decode())
        self._h, self.params, self.first_device = h, p, first_device
        return self

    def AddKeySet(self, baseG, bk, ksk):
        """Load the key set of another gadget base of a timeOptimization context (m_BTKey_map, binfhecontext.cpp:222-247);
        with three sets loaded EvalSign / EvalDecomp switch base like the reference's scalar path
        (binfhe-base-scheme.cpp:342-360, 411-428)."""
        b, k = _Buf(bk), _Buf(ksk)
        if b.space != k.space:
            raise TfheB200Error(-1, "AddKeySet: bk and ksk must live in the same memory space")
        self._call("tfhe_b200_add_key_set", self._handle(), C.c_uint32(int(baseG)), b.ptr,
                   C.c_size_t(int(np.prod(b.shape))), k.ptr, C.c_size_t(int(np.prod(k.shape))), b.space)
        return self

    @property
    def num_key_sets(self):
        return load_library().tfhe_b200_num_key_sets(self._h) if self._h else 0

    def GPUClean(self):
        """binfhecontext.cpp:362-365."""
        if self._h is not None:
            load_library().tfhe_b200_clean(self._h)
            self._h = None

    def __del__(self):
        try:
            self.GPUClean()
        except Exception:
            pass

    @property
    def num_gpus(self):
        return load_library().tfhe_b200_num_gpus(self._h) if self._h else 0

    @property
    def kernel_variant(self):
        return load_library().tfhe_b200_kernel_variant(self._h).decode() if self._h else ""

    def set_option(self, key, value):
        self._call("tfhe_b200_set_option", self._handle(), key.encode(), C.c_int64(int(value)))

    # ---- plumbing -------------------------------------------------------------------------------------
    def _handle(self):
        if self._h is None:
            raise TfheB200Error(-1, "GPUSetup has not been called")
        return self._h

    def _call(self, name, *args):
        L = load_library()
        rc = getattr(L, name)(*args)
        if rc < 0:
            raise TfheB200Error(rc, L.tfhe_b200_last_error().decode())
        return rc

    def _st(self):
        return C.byref(self.last_stats)

    def _dev(self, *bufs):
        """Device-space operands must live on the handle's first GPU (include/tfhe_b200.h, `space`)."""
        for b in bufs:
            if b.space == DEVICE and b.device != self.first_device:
                raise TfheB200Error(-1, f"device-resident operands must live on cuda:{self.first_device} "
                                        f"(the handle's first GPU), got cuda:{b.device}")

    @staticmethod
    def _batch(buf, what):
        if len(buf.shape) != 2 or buf.shape[0] == 0:
            raise TfheB200Error(-1, f"ERROR: {what}: input vector is empty")
        return buf.shape[0]

    # ---- batched operations ---------------------------------------------------------------------------
    def EvalBinGate(self, gate, ct1, ct2, ct_mod=None, out=None):
        g = GATES[gate] if isinstance(gate, str) else int(gate)
        a, b = _Buf(ct1), _Buf(ct2)
        self._dev(a, b)
        if a.shape[0] == 0 or b.shape[0] == 0:
            raise TfheB200Error(-1, "ERROR: EvalBinGate: input vector is empty")
        if a.shape != b.shape:
            raise TfheB200Error(-1, "ERROR: EvalBinGate: input ciphertexts size unmatched")
        if a.space != b.space:
            raise TfheB200Error(-1, "EvalBinGate: inputs must live in the same memory space")
        if out is None:
            out = a.empty_like_out(a.shape)
        o = _Buf(out)
        if o.obj is not out and not _is_torch(out):
            raise TfheB200Error(-1, "EvalBinGate: `out` must be a contiguous uint64 array")
        self._call("tfhe_b200_eval_bin_gate", self._handle(), g, a.shape[0], a.ptr, b.ptr,
                   C.c_uint64(ct_mod or self.params.q), o.ptr, a.space, self._st())
        return out

    def EvalCircuit(self, inputs, nodes, outputs, ct_mod=None):
        """Gate-graph submission (tfhe_b200_eval_circuit): `inputs` [n_inputs][batch][n+1]; `nodes` a list of
        (gate, in0, in1) with gate a name from GATES or "NOT" (in1 ignored), wires numbered inputs first then node
        outputs; `outputs` the wires to return -> [len(outputs)][batch][n+1].  Bit-identical to evaluating the nodes
        one by one with EvalBinGate / EvalNOT; intermediates never leave the device."""
        a = _Buf(inputs)
        self._dev(a)
        if len(a.shape) != 3 or a.shape[0] == 0 or a.shape[1] == 0:
            raise TfheB200Error(-1, "ERROR: EvalCircuit: input vector is empty")
        arr = np.zeros((len(nodes), 3), dtype=np.int32)
        for i, (g, x, y) in enumerate(nodes):
            arr[i] = (NOT_GATE if g == "NOT" else (GATES[g] if isinstance(g, str) else int(g)), x, x if y is None else y)
        ow = np.ascontiguousarray(outputs, dtype=np.int32)
        if ow.size == 0:
            raise TfheB200Error(-1, "EvalCircuit: no output wires")
        out = a.empty_like_out((int(ow.size), a.shape[1], a.shape[2]))
        self._call("tfhe_b200_eval_circuit", self._handle(), a.shape[1], a.shape[0], a.ptr,
                   C.c_uint64(ct_mod or self.params.q), len(nodes), C.c_void_p(arr.ctypes.data), int(ow.size),
                   C.c_void_p(ow.ctypes.data), _Buf(out).ptr, a.space, self._st())
        return out

    def BootstrapFunc(self, ct, ct_mod, table, fmod):
        a, t = _Buf(ct), _Buf(table)
        self._dev(a, t)
        batch = self._batch(a, "EvalFunc")
        per_ct = int(len(t.shape) == 2)
        out = a.empty_like_out(a.shape)
        self._call("tfhe_b200_bootstrap_func", self._handle(), batch, a.ptr, C.c_uint64(ct_mod), t.ptr, per_ct,
                   C.c_uint64(fmod), _Buf(out).ptr, a.space, self._st())
        return out

    def EvalFunc(self, ct, lut, ct_mod=None):
        a, t = _Buf(ct), _Buf(lut)
        self._dev(a, t)
        batch = self._batch(a, "EvalFunc")
        per_ct = int(len(t.shape) == 2)
        if per_ct and t.shape[0] != batch:
            raise TfheB200Error(-1, "ERROR: EvalFunc: input ciphertexts size unmatched with LUT size")
        out = a.empty_like_out(a.shape)
        self._call("tfhe_b200_eval_func", self._handle(), batch, a.ptr, C.c_uint64(ct_mod or self.params.q), t.ptr,
                   C.c_size_t(t.shape[-1]), per_ct, _Buf(out).ptr, a.space, self._st())
        return out

    def EvalFloor(self, ct, ct_mod, roundbits=0):
        a = _Buf(ct)
        self._dev(a)
        batch = self._batch(a, "EvalFunc")
        out = a.empty_like_out(a.shape)
        self._call("tfhe_b200_eval_floor", self._handle(), batch, a.ptr, C.c_uint64(ct_mod), C.c_uint32(roundbits),
                   _Buf(out).ptr, a.space, self._st())
        return out

    def EvalSign(self, ct, ct_mod):
        a = _Buf(ct)
        self._dev(a)
        batch = self._batch(a, "EvalFunc")
        out = a.empty_like_out(a.shape)
        self._call("tfhe_b200_eval_sign", self._handle(), batch, a.ptr, C.c_uint64(ct_mod), _Buf(out).ptr, a.space,
                   self._st())
        return out

    def EvalDecomp(self, ct, ct_mod, max_digits=8):
        a = _Buf(ct)
        self._dev(a)
        batch = self._batch(a, "EvalFunc")
        out = a.empty_like_out((batch, max_digits, a.shape[1]))
        mods = np.zeros(max_digits, dtype=np.uint64)
        nd = self._call("tfhe_b200_eval_decomp", self._handle(), batch, a.ptr, C.c_uint64(ct_mod), max_digits,
                        _Buf(out).ptr, C.c_void_p(mods.ctypes.data), a.space, self._st())
        return out[:, :nd], [int(m) for m in mods[:nd]]

    def CiphertextMulMatrix(self, ct, matrix, modulus):
        a = _Buf(ct)
        self._dev(a)
        m = _Buf(matrix, dtype=np.int64)
        self._dev(m)
        if len(a.shape) != 2 or a.shape[0] == 0:
            raise TfheB200Error(-1, "Input ciphertexts are empty.")
        if len(m.shape) != 2 or m.shape[0] == 0 or m.shape[1] == 0:
            raise TfheB200Error(-1, "Input matrix is empty.")
        if m.shape[0] != a.shape[0]:
            raise TfheB200Error(-1, "The number of rows of the matrix must be equal to the number of input ciphertexts.")
        out = a.empty_like_out((m.shape[1], a.shape[1]))
        self._call("tfhe_b200_mul_matrix", self._handle(), a.shape[0], m.shape[1], a.ptr, m.ptr, C.c_uint64(modulus),
                   _Buf(out).ptr, a.space, self._st())
        return out

    # ---- operator-level entry points (what the reference's host code calls) -----------------------------
    def EvalAcc(self, a_mask, ct_mod, acc):
        """GPUFFTBootstrap::EvalAcc_CUDA contract (bootstrapping.cuh:111-124)."""
        a, ac = _Buf(a_mask), _Buf(acc)
        self._dev(a, ac)
        out = ac.obj.clone() if ac.space == DEVICE else ac.obj.copy()
        self._call("tfhe_b200_eval_acc", self._handle(), a.shape[0], a.ptr, C.c_uint64(ct_mod), _Buf(out).ptr,
                   a.space, self._st())
        return out

    def MKMSwitch(self, ct_ext, fmod):
        """GPUFFTBootstrap::MKMSwitch_CUDA contract (bootstrapping.cuh:126-136)."""
        a = _Buf(ct_ext)
        self._dev(a)
        out = a.empty_like_out((a.shape[0], self.params.n + 1))
        self._call("tfhe_b200_mkmswitch", self._handle(), a.shape[0], a.ptr, C.c_uint64(fmod), _Buf(out).ptr, a.space,
                   self._st())
        return out

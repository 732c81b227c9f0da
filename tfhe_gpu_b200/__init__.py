"""tfhe_gpu_b200 -- B200-native batched TFHE/FHEW bootstrapping engine.

The product is ``libtfhe_b200.so`` (hand-written sm_100a CUDA kernels behind the C ABI of ``include/tfhe_b200.h``).
This package holds:

* ``csrc/``     the kernels and the C ABI,
* ``adapter/``  the C++ shim that defines the reference's ``lbcrypto::GPUFFTBootstrap`` / ``GPULWEOperation``
                operator entry points on top of the C ABI (drop-in behind ``BinFHEContext``),
* ``context.py`` a ctypes mirror of the reference's batched ``BinFHEContext`` surface over flat uint64 arrays
                (``GPUSetup`` / ``GPUClean`` / ``EvalBinGate`` / ``EvalFunc`` / ``EvalFloor`` / ``EvalSign`` /
                ``EvalDecomp`` / ``CiphertextMulMatrix``), used by the tests and the benchmark.

There is no CPU fallback: if the CUDA library is missing or no GPU is present every compute call raises.
"""
from .context import (  # noqa: F401
    BinFHEContextB200,
    TfheB200Error,
    Params,
    Stats,
    GATES,
    lib_path,
    load_library,
    gpu_keygen,
    serialized_info,
    flatten_serialized,
)

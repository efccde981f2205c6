"""cuBLAS configuration for the field MLP's contractions (the one part of the step that stays on a library).

The field MLP is fp32 (the reference trains without --fp16).  On a B200, torch's bundled cuBLAS (12.8) runs fp32 GEMMs
on the SIMT pipe (30-45 TFLOP/s measured for the [65536, 256..416] x [.., 256] shapes of this MLP).  cuBLAS 12.9 — the
one installed with the CUDA toolkit of this image, /usr/local/cuda/lib64 — can run the same fp32 GEMM on the BF16
tensor cores with its BF16x9 algorithm (CUBLAS_COMPUTE_32F_EMULATED_16BFX9: every fp32 operand is split into three
bf16 terms, the nine partial products are accumulated in fp32).  Measured on a B200 against a float64 product
(tests/perf/cublas_emulation_check.py, profiles/r1b_cublas_emulation.txt): max error 2e-7 of the result's scale with
BF16x9 against 9e-7 for the SIMT SGEMM, at 75-92 TFLOP/s.  It is an fp32-accurate replacement, not a reduced-precision
mode (nothing like TF32's 10-bit mantissa).

use_system_cublas() must run BEFORE `import torch`: it maps the toolkit's libcublasLt / libcublas into the global
symbol scope, so that torch's own calls (cublasSgemm, cublasGemmEx, cublasLtMatmul ...) bind to 12.9, and switches the
emulation on through cuBLAS' own environment variable.  If the toolkit library is absent or older than 12.9 it does
nothing and says so; the GEMMs then run as plain SGEMM.
"""
import ctypes
import os
import sys

_STATE = {"mode": "torch-bundled cuBLAS, fp32 SGEMM (SIMT)", "emulated": False, "configured": False}
_CANDIDATES = ("/usr/local/cuda/lib64", "/usr/local/cuda/targets/x86_64-linux/lib")


def use_system_cublas(emulate_fp32=True):
    """Returns True when torch will run on a cuBLAS with BF16x9 fp32 emulation enabled."""
    if os.environ.get("NSVF_NO_CUBLAS_EMULATION"):
        return False
    if _STATE["configured"]:
        return _STATE["emulated"]
    if "torch" in sys.modules:
        raise RuntimeError("nsvf_b200.blas.use_system_cublas() must be called before `import torch`")
    for d in _CANDIDATES:
        lt, bl = os.path.join(d, "libcublasLt.so.12"), os.path.join(d, "libcublas.so.12")
        if not (os.path.exists(lt) and os.path.exists(bl)):
            continue
        try:
            ctypes.CDLL(lt, mode=ctypes.RTLD_GLOBAL)
            lib = ctypes.CDLL(bl, mode=ctypes.RTLD_GLOBAL)
        except OSError:
            continue
        _STATE["configured"] = True
        if not hasattr(lib, "cublasSetEmulationStrategy"):      # < 12.9: no BF16x9
            _STATE["mode"] = "system cuBLAS without BF16x9 (< 12.9), fp32 SGEMM (SIMT)"
            return False
        if emulate_fp32:
            os.environ.setdefault("CUBLAS_EMULATE_SINGLE_PRECISION", "1")
            _STATE["emulated"] = os.environ["CUBLAS_EMULATE_SINGLE_PRECISION"] == "1"
        _STATE["mode"] = ("cuBLAS 12.9 (%s), fp32 on BF16 tensor cores via BF16x9 emulation (fp32-accurate)" % d
                          if _STATE["emulated"] else "cuBLAS 12.9 (%s), fp32 SGEMM (SIMT)" % d)
        return _STATE["emulated"]
    return False


def mode():
    return _STATE["mode"]


def emulated():
    return _STATE["emulated"]

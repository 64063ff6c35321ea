"""FP32 GEMMs on the tensor cores without giving up fp32 accuracy.

The policy's fully-connected / LSTM GEMMs are float32 like the reference's.  PyTorch 2.11 bundles cuBLAS 12.8, whose
fp32 SGEMM on B200 is a SIMT kernel (about 50-60 TFLOP/s, half of the rollout+update step).  cuBLAS 12.9 -- shipped in
this image under /usr/local/cuda -- adds FP32 *emulation* on the BF16 tensor cores (CUBLAS_COMPUTE_32F_EMULATED_16BFX9):
every fp32 input is split exactly into three bf16 values (3 x 8 significand bits = fp32's 24), the nine partial
products are accumulated in fp32, so the result has fp32 accuracy (unlike TF32, which drops 13 mantissa bits) at several
times the SIMT throughput.

cuBLAS reads CUBLAS_EMULATE_SINGLE_PRECISION when the library is loaded, and a process can only hold one
libcublas.so.12, so this must run BEFORE `import torch`:

    from active_tracking_rl_b200 import blas
    blas.enable_fp32_emulation()      # no-op (returns False) if torch is already imported or cuBLAS 12.9 is absent
    import torch
"""
import ctypes
import glob
import os
import sys

_STATE = {"enabled": False, "reason": "not requested"}


def _find(name):
    for d in (os.environ.get("TRACK2D_CUBLAS_DIR"), "/usr/local/cuda/lib64", "/usr/local/cuda/targets/x86_64-linux/lib"):
        if d:
            hits = sorted(glob.glob(os.path.join(d, name + ".12.9*"))) or sorted(glob.glob(os.path.join(d, name + ".12.[1-9][0-9]*")))
            if hits:
                return hits[-1]
    return None


def enable_fp32_emulation(strategy="performant"):
    """Load cuBLAS >= 12.9 ahead of torch's bundled copy and switch BF16x9 fp32 emulation on.  Returns True if active."""
    if _STATE["enabled"]:
        return True
    if "torch" in sys.modules:
        _STATE["reason"] = "torch was imported first"
        return False
    lt, blas = _find("libcublasLt.so"), _find("libcublas.so")
    if not lt or not blas:
        _STATE["reason"] = "cuBLAS >= 12.9 not found"
        return False
    os.environ["CUBLAS_EMULATE_SINGLE_PRECISION"] = "1"
    os.environ.setdefault("CUBLAS_EMULATION_STRATEGY", strategy)
    try:
        ctypes.CDLL(lt, mode=ctypes.RTLD_GLOBAL)   # same SONAME as torch's bundled copy: the loader will reuse these
        ctypes.CDLL(blas, mode=ctypes.RTLD_GLOBAL)
    except OSError as ex:
        _STATE["reason"] = "could not load %s: %s" % (blas, ex)
        os.environ.pop("CUBLAS_EMULATE_SINGLE_PRECISION", None)
        return False
    _STATE.update(enabled=True, reason="cuBLAS %s, BF16x9 fp32 emulation (%s)" % (os.path.basename(blas), os.environ["CUBLAS_EMULATION_STRATEGY"]))
    return True


def status():
    return dict(_STATE)

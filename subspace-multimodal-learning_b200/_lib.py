"""ctypes binding of libsml_b200.so (the C-ABI declared in include/sml_b200.h).

There is no CPU fallback: if the library is missing it is built with nvcc (csrc/build.py); if that is
impossible, importing the ops raises.  Every op launches on torch's current CUDA stream.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libsml_b200.so")

c_i32, c_i64, c_f32, c_u64, c_vp = ctypes.c_int32, ctypes.c_int64, ctypes.c_float, ctypes.c_uint64, ctypes.c_void_p


class GemmDesc(ctypes.Structure):
    _fields_ = [
        ("M", c_i32), ("N", c_i32), ("K", c_i32), ("nb1", c_i32), ("nb2", c_i32),
        ("A", c_vp), ("sa_m", c_i64), ("sa_k", c_i64), ("sa_b1", c_i64), ("sa_b2", c_i64),
        ("B", c_vp), ("sb_k", c_i64), ("sb_n", c_i64), ("sb_b1", c_i64), ("sb_b2", c_i64),
        ("C", c_vp), ("sc_m", c_i64), ("sc_n", c_i64), ("sc_b1", c_i64), ("sc_b2", c_i64),
        ("alpha", c_f32), ("beta", c_f32), ("diag", c_f32),
        ("bias", c_vp), ("relu", c_i32),
        ("A_lo", c_vp), ("B_lo", c_vp), ("Cb_hi", c_vp), ("Cb_lo", c_vp),
        ("rowtab", c_vp), ("rt_m", c_i64), ("rt_b1", c_i64), ("rt_b2", c_i64), ("rowdiv", c_i32), ("rowtab_scale", c_f32),
    ]


def _load():
    if not os.path.exists(_LIB_PATH):
        import importlib.util
        spec = importlib.util.spec_from_file_location("_sml_build", os.path.join(_HERE, "csrc", "build.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.build()
    lib = ctypes.CDLL(_LIB_PATH)
    lib.sml_last_error.restype = ctypes.c_char_p
    lib.sml_launch_count.restype = c_i64
    lib.sml_deform_pwl_krange_floats.restype = c_i64
    return lib


lib = _load()


def check(rc, what=""):
    if rc != 0:
        raise RuntimeError("libsml_b200 %s failed (%d): %s" % (what, rc, lib.sml_last_error().decode()))


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("sml_b200 ops are CUDA-only (sm_100a); got a %s tensor — there is no CPU fallback" % t.device)


def ptr(t):
    return None if t is None else c_vp(t.data_ptr())


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def stream():
    """raw cudaStream_t of torch's current stream (the fast C accessor when this torch build has it: the Python
    ``current_stream()`` path costs ~10 us, paid ~330 times per TransMIL train step)"""
    if _raw_stream is not None:
        return c_vp(_raw_stream(torch.cuda.current_device()))
    return c_vp(torch.cuda.current_stream().cuda_stream)


def launch_count():
    return int(lib.sml_launch_count())

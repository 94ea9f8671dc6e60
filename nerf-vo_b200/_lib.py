"""ctypes binding of the C-ABI library (include/nvo_b200.h).  No torch types cross the boundary: tensors are
passed as raw device pointers + sizes, the stream as the current torch CUDA stream handle.

There is NO fallback: if `libnvo_b200.so` is missing the first call raises, telling the user how to build it
(the reference raises EnvironmentError at import when its extension is unusable,
tiny_cuda_nn/bindings/torch/tinycudann/modules.py:18-19,58-59)."""
from __future__ import annotations

import ctypes
import os
import re
from typing import Optional

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libnvo_b200.so")
HEADER = os.path.join(os.path.dirname(HERE), "include", "nvo_b200.h")

NVO_MAX_LEVELS = 32
NVO_MAX_LAYERS = 6
NVO_F32, NVO_F16, NVO_F16_TMH, NVO_F32_TMF = 0, 1, 2, 3
ACT = {"none": 0, "relu": 1, "sigmoid": 2, "tanh": 3, "exponential": 4, "exp": 4, "trunc_exp": 5}


class GridDesc(ctypes.Structure):
    _fields_ = [("n_levels", ctypes.c_int32), ("log2_T", ctypes.c_int32), ("table_dtype", ctypes.c_int32), ("out_dtype", ctypes.c_int32),
                ("scalings", ctypes.c_float * NVO_MAX_LEVELS)]


class MlpDesc(ctypes.Structure):
    _fields_ = [("n_layers", ctypes.c_int32), ("in_dim", ctypes.c_int32), ("dims", ctypes.c_int32 * NVO_MAX_LAYERS),
                ("acts", ctypes.c_int32 * NVO_MAX_LAYERS)]


_lib: Optional[ctypes.CDLL] = None

# argument kinds: D descriptor pointer, S stream, p device pointer, l int64, i int32, f float, d double.
# Signatures are derived from include/nvo_b200.h itself so the binding cannot drift from the header.
def _parse_header():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    sigs = {}
    for m in re.finditer(r"\b(?:const\s+char\s*\*|int64_t|int)\s+(nvo_[a-z0-9_]+)\s*\((.*?)\)\s*;", src, flags=re.S):
        name, args = m.group(1), m.group(2)
        sig = ""
        for a in (x.strip() for x in args.split(",")):
            if not a or a == "void":
                continue
            if "nvo_grid_desc" in a or "nvo_mlp_desc" in a:
                sig += "D"
            elif re.search(r"\bstream\b", a):
                sig += "S"
            elif "*" in a:
                sig += "p"
            elif "int64_t" in a:
                sig += "l"
            elif "int32_t" in a:
                sig += "i"
            elif "double" in a:
                sig += "d"
            elif "float" in a:
                sig += "f"
            else:
                raise RuntimeError(f"nvo_b200: cannot bind argument '{a}' of {name}")
        sigs[name] = sig
    return sigs


_SIGS = _parse_header()
# functions declared `int64_t name(...)` in the header
_INT64_RETURNS = set(re.findall(r"\bint64_t\s+(nvo_[a-z0-9_]+)\s*\(", re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)))
_CT = {"D": ctypes.c_void_p, "S": ctypes.c_void_p, "p": ctypes.c_void_p, "l": ctypes.c_int64, "i": ctypes.c_int32, "f": ctypes.c_float, "d": ctypes.c_double}


def declared_symbols() -> list:
    """Every function name include/nvo_b200.h declares (used by the export test)."""
    return sorted(_SIGS)


def load() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"nvo_b200: native library not found at {LIB_PATH}. Build it with `python nerf-vo_b200/build.py` "
            "(or `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU / PyTorch fallback."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, sig in _SIGS.items():
        fn = getattr(lib, name)  # AttributeError here = header declares a symbol the library does not export
        fn.argtypes = [_CT[k] for k in sig]
        if name == "nvo_last_error":
            fn.restype = ctypes.c_char_p
        elif name in _INT64_RETURNS:
            fn.restype = ctypes.c_int64
        else:
            fn.restype = ctypes.c_int
    _lib = lib
    return lib


def launch_count() -> int:
    return int(load().nvo_launch_count())


def _ptr(x):
    if x is None:
        return None
    if isinstance(x, torch.Tensor):
        return x.data_ptr()
    return int(x)


def stream_handle() -> int:
    return torch.cuda.current_stream().cuda_stream


def call(name: str, *args):
    """Invoke a C-ABI entry point on the current torch stream; 'S' arguments are filled in automatically."""
    lib = load()
    sig = _SIGS[name]
    out, it = [], iter(args)
    for k in sig:
        if k == "S":
            out.append(stream_handle())
        elif k == "D":
            out.append(ctypes.addressof(next(it)))
        elif k == "p":
            out.append(_ptr(next(it)))
        elif k in "fd":
            out.append(float(next(it)))
        else:
            out.append(int(next(it)))
    rc = getattr(lib, name)(*out)
    if rc != 0:
        raise RuntimeError(f"{name} failed: {lib.nvo_last_error().decode()}")


def check(t: torch.Tensor, name: str, dtype=torch.float32, shape=None) -> torch.Tensor:
    """Input validation in the style of bindings.cpp:80-93 (CUDA, dtype, contiguous, sizes) -> RuntimeError."""
    if not isinstance(t, torch.Tensor):
        raise RuntimeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if t.dtype != dtype:
        raise RuntimeError(f"{name} must have dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")
    if shape is not None:
        if t.dim() != len(shape) or any(s is not None and s != d for s, d in zip(shape, t.shape)):
            raise RuntimeError(f"{name} has shape {tuple(t.shape)}, expected {tuple(shape)}")
    return t


def make_grid_desc(n_levels: int, log2_T: int, scalings, table_dtype=torch.float32, out_dtype=torch.float32) -> GridDesc:
    if not (1 <= n_levels <= NVO_MAX_LEVELS):
        raise RuntimeError(f"n_levels={n_levels} out of range [1,{NVO_MAX_LEVELS}]")
    d = GridDesc()
    d.n_levels, d.log2_T = int(n_levels), int(log2_T)
    d.table_dtype = NVO_F16 if table_dtype == torch.float16 else NVO_F32
    d.out_dtype = {"tmh": NVO_F16_TMH, "tmf": NVO_F32_TMF}.get(out_dtype) if isinstance(out_dtype, str) else (NVO_F16 if out_dtype == torch.float16 else NVO_F32)
    sc = [float(s) for s in scalings]
    for i in range(NVO_MAX_LEVELS):
        d.scalings[i] = sc[i] if i < n_levels else 0.0
    return d


def make_mlp_desc(in_dim: int, dims, acts) -> MlpDesc:
    """acts: one activation name per layer."""
    dims = [int(x) for x in dims]
    if not (1 <= len(dims) <= NVO_MAX_LAYERS):
        raise RuntimeError(f"MLP with {len(dims)} layers unsupported (max {NVO_MAX_LAYERS})")
    if len(acts) != len(dims):
        raise RuntimeError("one activation per layer required")
    d = MlpDesc()
    d.n_layers, d.in_dim = len(dims), int(in_dim)
    for i in range(NVO_MAX_LAYERS):
        d.dims[i] = dims[i] if i < len(dims) else 0
        d.acts[i] = ACT[acts[i].lower()] if i < len(dims) else 0
    return d

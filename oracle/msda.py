"""ORACLE (test infrastructure): ctypes loader for the C restatement of the MSDA core (oracle/msda_ref.c)."""
import ctypes
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "libmsda_ref.so"])


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libmsda_ref.so")
        if not os.path.exists(path):
            build()
        _LIB = ctypes.CDLL(path)
    return _LIB


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _prep(value, shapes, lsi, loc, w):
    dt = np.float64 if value.dtype == torch.float64 else np.float32
    v = np.ascontiguousarray(value.detach().cpu().numpy().astype(dt))
    lo = np.ascontiguousarray(loc.detach().cpu().numpy().astype(dt))
    ww = np.ascontiguousarray(w.detach().cpu().numpy().astype(dt))
    sh = np.ascontiguousarray(torch.as_tensor(shapes).cpu().numpy().astype(np.int64))
    ls = np.ascontiguousarray(torch.as_tensor(lsi).cpu().numpy().astype(np.int64))
    B, S, M, D = v.shape
    _, Lq, _, L, P, _ = lo.shape
    return dt, v, sh, ls, lo, ww, (B, S, M, D, L, Lq, P)


def msda_forward(value, shapes, lsi, loc, w):
    """value (B,S,M,D), shapes (L,2), lsi (L,), loc (B,Lq,M,L,P,2), w (B,Lq,M,L,P) -> (B,Lq,M*D) torch CPU."""
    dt, v, sh, ls, lo, ww, (B, S, M, D, L, Lq, P) = _prep(value, shapes, lsi, loc, w)
    out = np.empty((B, Lq, M * D), dtype=dt)
    fn = _lib().msda_ref_fwd_f64 if dt == np.float64 else _lib().msda_ref_fwd_f32
    fn(_ptr(v), _ptr(sh), _ptr(ls), _ptr(lo), _ptr(ww), _ptr(out), B, S, M, D, L, Lq, P)
    return torch.from_numpy(out)


def msda_backward(value, shapes, lsi, loc, w, grad_out):
    dt, v, sh, ls, lo, ww, (B, S, M, D, L, Lq, P) = _prep(value, shapes, lsi, loc, w)
    go = np.ascontiguousarray(grad_out.detach().cpu().numpy().astype(dt))
    gv = np.zeros_like(v)
    gl = np.zeros_like(lo)
    gw = np.zeros_like(ww)
    fn = _lib().msda_ref_bwd_f64 if dt == np.float64 else _lib().msda_ref_bwd_f32
    fn(_ptr(v), _ptr(sh), _ptr(ls), _ptr(lo), _ptr(ww), _ptr(go), _ptr(gv), _ptr(gl), _ptr(gw),
       B, S, M, D, L, Lq, P)
    return torch.from_numpy(gv), torch.from_numpy(gl), torch.from_numpy(gw)

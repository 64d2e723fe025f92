"""Drop-in for psoap.matrix_functions (the reference's Cython fills, psoap/matrix_functions.pyx).

Same names, positional signatures and in-place semantics: `mat` is caller-owned and fully written (both
triangles and the diagonal), the functions return None.  `mat` may be a float64 CUDA torch tensor (filled in
place on the device, no host traffic) or a C-contiguous float64 numpy array (filled on the device, then copied
back: N^2 doubles cross PCIe, which is what passing a host matrix to a device operator costs).
"""
import ctypes

import numpy as np

from . import _lib


def _check_mat(mat):
    torch = _lib.torch_cuda()
    if isinstance(mat, torch.Tensor):
        if mat.dtype != torch.float64 or mat.dim() != 2 or not mat.is_cuda or mat.stride(1) != 1:
            raise ValueError("mat must be a 2-D float64 CUDA tensor with unit column stride")
        return mat, None
    if not isinstance(mat, np.ndarray) or mat.dtype != np.float64 or mat.ndim != 2:
        raise ValueError("Buffer dtype mismatch, expected 'double' 2-D array")  # Cython buffer check analogue
    dev = torch.empty(mat.shape, dtype=torch.float64, device="cuda")
    return dev, mat


def _finish(dev, host):
    if host is not None:
        host[...] = dev.cpu().numpy()


def _fill_v11(mat, lwls, amps, ls):
    lib = _lib.load()
    dev, host = _check_mat(mat)
    N = dev.shape[0]
    if dev.shape[1] != N:
        raise ValueError("mat must be square")
    vecs = [_lib.dev_f64(v) for v in lwls]
    for v in vecs:
        if v.dim() != 1 or v.numel() != N:
            raise ValueError("wavelength vectors must be 1-D with len(mat) elements")
    ptrs = [_lib.ptr(v) for v in vecs] + [_lib.vp(None)] * (3 - len(vecs))
    _lib.check(lib.psoap_fill_v11(len(vecs), _lib.ptr(dev), dev.stride(0), N, ptrs[0], ptrs[1], ptrs[2],
                                  _lib.dbl_array(amps), _lib.dbl_array(ls), _lib.stream_ptr()))
    _finish(dev, host)


def fill_V11_f(mat, lwl_f, amp_f, l_f):
    """matrix_functions.pyx:21-57"""
    _fill_v11(mat, [lwl_f], [amp_f], [l_f])


def fill_V11_f_g(mat, lwl_f, lwl_g, amp_f, l_f, amp_g, l_g):
    """matrix_functions.pyx:101-144"""
    _fill_v11(mat, [lwl_f, lwl_g], [amp_f, amp_g], [l_f, l_g])


def fill_V11_f_g_h(mat, lwl_f, lwl_g, lwl_h, amp_f, l_f, amp_g, l_g, amp_h, l_h):
    """matrix_functions.pyx:151-201"""
    _fill_v11(mat, [lwl_f, lwl_g, lwl_h], [amp_f, amp_g, amp_h], [l_f, l_g, l_h])


def fill_V12_f(mat, lwl_f, lwl_predict, amp_f, l_f):
    """matrix_functions.pyx:63-94: mat is [len(lwl_f), len(lwl_predict)]."""
    lib = _lib.load()
    dev, host = _check_mat(mat)
    rows, cols = _lib.dev_f64(lwl_f), _lib.dev_f64(lwl_predict)
    M, N = rows.numel(), cols.numel()
    if dev.shape[0] < M or dev.shape[1] < N:
        raise ValueError("mat is smaller than len(lwl_f) x len(lwl_predict)")
    _lib.check(lib.psoap_fill_v12(_lib.ptr(dev), dev.stride(0), M, N, _lib.ptr(rows), _lib.ptr(cols), float(amp_f),
                                  float(l_f), _lib.stream_ptr()))
    _finish(dev, host)


def fill_V12_sum(mat_dev, lwls_rows, lwls_cols, amps, ls):
    """Sum over components of fill_V12_f into a CUDA tensor view (covariance.py:167-171, :272-278)."""
    lib = _lib.load()
    rows = [_lib.dev_f64(v) for v in lwls_rows]
    cols = [_lib.dev_f64(v) for v in lwls_cols]
    n = len(rows)
    rp = (_lib.vp * n)(*[v.data_ptr() for v in rows])
    cp = (_lib.vp * n)(*[v.data_ptr() for v in cols])
    _lib.check(lib.psoap_fill_v12n(n, _lib.ptr(mat_dev), mat_dev.stride(0), rows[0].numel(), cols[0].numel(),
                                   ctypes.cast(rp, ctypes.POINTER(_lib.vp)), ctypes.cast(cp, ctypes.POINTER(_lib.vp)),
                                   _lib.dbl_array(amps), _lib.dbl_array(ls), _lib.stream_ptr()))

"""Drop-in for the hot-path part of psoap.data: lredshift and replicate_wls (psoap/data.py:25-63).

`replicate_wls` runs on the device; it accepts the reference's arguments (masked 1-D ln-wavelengths,
velocities [n_components, n_epochs], boolean mask [n_epochs, n_pix]) and returns the same
[n_components, n_good_pix] array (numpy in, numpy out; CUDA tensor in, CUDA tensor out).
Chunk / Spectrum HDF5 IO is out of scope.
"""
import numpy as np

from . import _lib
from . import constants as C


def lredshift(lwl, v):
    """data.py:25-38 (host arithmetic; the device version is fused into the fills)."""
    return lwl + v / C.c_kms


def epoch_index(mask):
    """Epoch of every kept pixel in the row-major flattening of `mask` (what data.py:61 broadcasts)."""
    mask = np.asarray(mask, dtype=bool)
    n_epochs, n_pix = mask.shape
    return np.ascontiguousarray(np.repeat(np.arange(n_epochs, dtype=np.int32), n_pix).reshape(mask.shape)[mask])


def replicate_wls(lwls, velocities, mask):
    """data.py:40-63"""
    lib = _lib.load()
    torch = _lib.torch_cuda()
    on_dev = isinstance(lwls, torch.Tensor) and lwls.is_cuda
    lw = _lib.dev_f64(lwls)
    vel = _lib.dev_f64(velocities)
    if vel.dim() != 2:
        raise ValueError("velocities must be [n_components, n_epochs]")
    ncomp, n_epochs = vel.shape
    if isinstance(mask, torch.Tensor):
        mask = mask.cpu().numpy()
    ep = torch.from_numpy(epoch_index(mask)).cuda()
    N = lw.numel()
    if ep.numel() != N:
        raise ValueError("mask selects %d pixels but lwls has %d" % (ep.numel(), N))
    out = torch.empty((ncomp, N), dtype=torch.float64, device="cuda")
    _lib.check(lib.psoap_replicate_wls(_lib.ptr(out), _lib.ptr(lw), _lib.ptr(ep), N, _lib.ptr(vel), ncomp, n_epochs,
                                       _lib.stream_ptr()))
    return out if on_dev else out.cpu().numpy()

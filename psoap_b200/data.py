"""Drop-in for the hot-path part of psoap.data: lredshift and replicate_wls (psoap/data.py:25-63).

`replicate_wls` runs on the device; it accepts the reference's arguments (masked 1-D ln-wavelengths,
velocities [n_components, n_epochs], boolean mask [n_epochs, n_pix]) and returns the same
[n_components, n_good_pix] array (numpy in, numpy out; CUDA tensor in, CUDA tensor out).
`Chunk` is the reference's container for one spectral chunk (psoap/data.py:120-197) with the same attributes and
the same file naming; HDF5 needs h5py (not in this image), the same five datasets in an .npz are the fallback
container.  `read_chunks_dat` / `write_chunks_dat` handle the `order wl0 wl1` table (psoap/data/chunks.dat).
"""
import os

import numpy as np

from . import _lib
from . import constants as C


def redshift(wl, v):
    """data.py:10-23: relativistic Doppler shift of linear wavelengths (positive v lengthens)."""
    return wl * np.sqrt((C.c_kms + v) / (C.c_kms - v))


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
    if np.asarray(mask).shape[0] != n_epochs:
        # data.py:61 broadcasts velocities[i][:, None] against the mask: a row-count mismatch raises there too
        raise ValueError("mask has %d rows but velocities holds %d epochs" % (np.asarray(mask).shape[0], n_epochs))
    ep = torch.from_numpy(epoch_index(mask)).cuda()
    N = lw.numel()
    if ep.numel() != N:
        raise ValueError("mask selects %d pixels but lwls has %d" % (ep.numel(), N))
    out = torch.empty((ncomp, N), dtype=torch.float64, device="cuda")
    _lib.check(lib.psoap_replicate_wls(_lib.ptr(out), _lib.ptr(lw), _lib.ptr(ep), N, _lib.ptr(vel), ncomp, n_epochs,
                                       _lib.stream_ptr()))
    return out if on_dev else out.cpu().numpy()


def _have_h5py():
    try:
        import h5py  # noqa: F401
        return True
    except ImportError:
        return False


class Chunk:
    """One chunk of data, arrays of shape [n_epochs, n_pix]: wl, fl, sigma, date, mask (data.py:120-136).
    After `apply_mask()` the arrays are the flattened good pixels the likelihood consumes (data.py:138-147)."""
    DATASETS = ("wl", "fl", "sigma", "date", "mask")

    def __init__(self, wl, fl, sigma, date, mask=None):
        self.wl = wl
        self.lwl = np.log(wl)
        self.fl = fl
        self.sigma = sigma
        self.date = date
        self.date1D = date[:, 0]
        self.mask = np.ones_like(self.wl, dtype="bool") if mask is None else mask
        self.n_epochs, self.n_pix = self.wl.shape

    def apply_mask(self):
        self.wl = self.wl[self.mask]
        self.lwl = self.lwl[self.mask]
        self.fl = self.fl[self.mask]
        self.sigma = self.sigma[self.mask]
        self.date = self.date[self.mask]
        self.N = len(self.wl)

    def as_farm_chunk(self):
        """The dict `psoap_b200.farm.ChunkFarm` takes (call after apply_mask())."""
        return dict(lwl=self.lwl, fl=self.fl, sigma=self.sigma, mask=self.mask, date1D=self.date1D)

    @staticmethod
    def _fname(order, wl0, wl1, prefix):
        return prefix + C.chunk_fmt.format(order, wl0, wl1)

    @classmethod
    def open(cls, order, wl0, wl1, limit=100, prefix=""):
        """data.py:149-174: first `limit` epochs of chunk_{order}_{wl0}_{wl1}.hdf5 (or .npz), as float64."""
        base = cls._fname(order, wl0, wl1, prefix)
        if os.path.exists(base + ".hdf5"):
            if not _have_h5py():
                raise ImportError("reading %s.hdf5 needs h5py; convert it to .npz with the same dataset names" % base)
            import h5py
            with h5py.File(base + ".hdf5", "r") as f:
                arr = {k: f[k][:limit] for k in cls.DATASETS}
        else:
            with np.load(base + ".npz") as f:
                arr = {k: f[k][:limit] for k in cls.DATASETS}
        return cls(arr["wl"].astype(np.float64), arr["fl"].astype(np.float64), arr["sigma"].astype(np.float64),
                   arr["date"].astype(np.float64), np.array(arr["mask"], dtype="bool"))

    def save(self, order, wl0, wl1, prefix="", fmt=None):
        """data.py:176-197: datasets wl, fl, sigma, date (f8) and mask (bool), all of shape [n_epochs, n_pix]."""
        base = self._fname(order, wl0, wl1, prefix)
        fmt = fmt or ("hdf5" if _have_h5py() else "npz")
        if fmt == "hdf5":
            import h5py
            with h5py.File(base + ".hdf5", "w") as f:
                for k in ("wl", "fl", "sigma", "date"):
                    f.create_dataset(k, self.wl.shape, dtype="f8")[:] = getattr(self, k)
                f.create_dataset("mask", self.wl.shape, dtype="bool")[:] = self.mask
            return base + ".hdf5"
        np.savez(base + ".npz", wl=np.asarray(self.wl, dtype="f8"), fl=np.asarray(self.fl, dtype="f8"),
                 sigma=np.asarray(self.sigma, dtype="f8"), date=np.asarray(self.date, dtype="f8"),
                 mask=np.asarray(self.mask, dtype="bool"))
        return base + ".npz"


def read_chunks_dat(fname="chunks.dat"):
    """The chunk table written by psoap-generate-chunks and read with astropy.io.ascii in
    sample_parallel.py:44-46: a header line `order wl0 wl1`, then one whitespace-separated row per chunk.
    Returns a list of (order, wl0, wl1); `order` stays an int when it parses as one."""
    rows = []
    with open(fname) as f:
        lines = [ln.strip() for ln in f if ln.strip() and not ln.lstrip().startswith("#")]
    if not lines or lines[0].split() != ["order", "wl0", "wl1"]:
        raise ValueError("%s: expected the header 'order wl0 wl1'" % fname)
    for ln in lines[1:]:
        o, a, b = ln.split()
        try:
            o = int(o)
        except ValueError:
            pass
        rows.append((o, float(a), float(b)))
    return rows


def write_chunks_dat(rows, fname="chunks.dat"):
    with open(fname, "w") as f:
        f.write("order wl0 wl1\n")
        for o, a, b in rows:
            f.write("%s %s %s\n" % (o, repr(float(a)), repr(float(b))))

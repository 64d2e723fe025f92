"""Constants of the hot path (psoap/constants.py:13, :39)."""
c_kms = 2.99792458e5  # km s^-1
chunk_fmt = "chunk_{:}_{:.0f}_{:.0f}"  # order, wl0, wl1

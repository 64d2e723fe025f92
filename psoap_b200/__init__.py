"""psoap_b200 — B200-native (sm_100a) implementation of PSOAP's GP log-likelihood hot path.

Mirrors the reference's operator surface for this path:
  psoap_b200.matrix_functions  <-> psoap.matrix_functions   (fill_V11_f / _f_g / _f_g_h, fill_V12_f)
  psoap_b200.covariance        <-> psoap.covariance         (lnlike_*, lnlike, predict_*)
  psoap_b200.data              <-> psoap.data               (lredshift, replicate_wls)
  psoap_b200.orbit             <-> psoap.orbit              (SB1/SB2/ST1/ST2/ST3 get_velocities, models)
  psoap_b200.farm.ChunkFarm    <-> psoap.sample_parallel    (Worker.lnprob farm + master sum)
All compute lives in csrc/libpsoap_b200.so (hand-written CUDA behind the C ABI of include/psoap_b200.h);
there is no CPU fallback.
"""
__version__ = "0.1.0"

from . import constants  # noqa: F401

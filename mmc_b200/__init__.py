"""mmc_b200: B200-native (sm_100a) mesh-based Monte Carlo photon transport -- the photon random-walk hot path of
fangq/mmc behind a C-ABI (include/mmc_b200.h), with a pmmc-style Python front end (mmc_b200.run)."""
from .api import MMCError, Session, gpuinfo, host_seeds, lib, mesh_facenb, mesh_initelem, mesh_volumes, photon_shares, rng_selftest, run, run_multi, run_onecall, version  # noqa: F401
from . import meshgen  # noqa: F401

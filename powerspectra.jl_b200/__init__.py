"""powerspectra.jl_b200 -- B200-native (sm_100a) drop-in for the Wigner-3j hot path of
xzackli/PowerSpectra.jl: `mcm` and `coupledcov` over a `CovarianceWorkspace`.

The product is csrc/ -> libpsb200.so (C ABI in include/psb200.h) reached from Julia with
`ccall` (julia/PowerSpectraB200.jl).  This package is the host-side mirror
of the reference interface for that path, used by the tests and the benchmark because no
Julia toolchain exists in the build image.  Import name: `powerspectra_jl_b200`
(the directory name contains a dot; the alias module at the repo root maps it).
"""
from ._lib import LIB_PATH, PSB200Error, lib
from .beam import BandedSpectralMatrix, k_u, quickpolW, quickpolXi
from .covariance import (ConstantDict, CovarianceWorkspace, coupledcov, coupledcovEEEE, coupledcovTEEE,
                         coupledcovTETE, coupledcovTTEE, coupledcovTTTE, coupledcovTTTT, loop_covEEEE,
                         loop_covTEEE, loop_covTEEE_planck, loop_covTETE, loop_covTTEE, loop_covTTTE,
                         loop_covTTTT, window_function_W)
from .healpix import (CovField, HealpixMap, PolarizedHealpixMap, alm2cl_device, alm2map, effective_weight_alm, map2alm,
                      nside2lmax, nside2npix, precompute_effective_weights, weights_needed)
from .modecoupling import (Alm, alm2cl, inner_mcm00, inner_mcm02, inner_mcmmm, inner_mcmpp,
                           inner_mcmpp_mcmmm, maskedalm2spectra, maskedalm2spectra_device, mcm,
                           mcm_master, mcm_solve)
from .spectral import (BlockSpectralMatrix, SpectralArray, SpectralVector, decouple_covmat, decouple_covmat_device,
                       free_pinned, pinned_spectralzeros, spectralones, spectralzeros)

__all__ = [
    "mcm", "mcm_master", "mcm_solve", "maskedalm2spectra", "maskedalm2spectra_device", "decouple_covmat_device", "coupledcov", "CovarianceWorkspace", "window_function_W", "ConstantDict",
    "SpectralArray", "SpectralVector", "BlockSpectralMatrix", "spectralzeros", "spectralones",
    "decouple_covmat", "BandedSpectralMatrix", "quickpolXi", "quickpolW", "k_u", "Alm", "alm2cl", "HealpixMap", "PolarizedHealpixMap", "CovField", "map2alm", "alm2map", "alm2cl_device",
    "effective_weight_alm", "precompute_effective_weights", "weights_needed", "nside2lmax", "nside2npix", "lib", "LIB_PATH", "PSB200Error",
    "pinned_spectralzeros", "free_pinned",
]

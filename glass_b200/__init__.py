"""
glass_b200 -- B200-native (sm_100a) implementation of the GLASS per-shell field-generation
hot path behind GLASS's own API for that path (flat re-export like ``glass/__init__.py``).

The hot path of SURVEY.md section 8 and its "next" rows (Gaussian-spectra pre-step, catalogue
sink, per-galaxy post-processing) live here; everything else (shell construction, n(z) models,
cosmology) stays with upstream GLASS.  There is no CPU fallback: the kernels
are in ``libglassb200.so`` (``python -m glass_b200.build``) and calls raise if it is
missing.
"""

from . import algorithm, fields, galaxies, grf, harmonics, healpix, lensing, observations, points, rng, shapes, sharding, shells, transformcl, user  # noqa: F401
from .fields import (  # noqa: F401
    check_posdef_spectra,
    cls2cov,
    cltovar,
    compute_gaussian_spectra,
    cov_from_spectra,
    discretized_cls,
    effective_cls,
    enumerate_spectra,
    gaussian_fields,
    generate,
    generate_gaussian,
    generate_lognormal,
    getcl,
    glass_to_healpix_spectra,
    healpix_to_glass_spectra,
    iternorm,
    lognormal_fields,
    lognormal_gls,
    lognormal_shift_hilbert2011,
    nfields_from_nspectra,
    regularized_spectra,
    solve_gaussian_spectra,
    spectra_indices,
)
from .galaxies import galaxy_shear, gaussian_phz, redshifts, redshifts_from_bins, redshifts_from_nz  # noqa: F401
from .harmonics import multalm  # noqa: F401
from .lensing import (  # noqa: F401
    MultiPlaneConvergence,
    deflect,
    from_convergence,
    multi_plane_matrix,
    multi_plane_weights,
    shear_from_convergence,
)
from .observations import vmap_galactic_ecliptic  # noqa: F401
from .points import (  # noqa: F401
    displace,
    displacement,
    effective_bias,
    linear_bias,
    loglinear_bias,
    position_weights,
    positions_from_delta,
    uniform_positions,
)
from .shapes import ellipticity_gaussian, ellipticity_intnorm  # noqa: F401
from .shells import RadialWindow  # noqa: F401
from .user import load_cls, save_cls, write_catalog  # noqa: F401

__version__ = "0.1.0"

"""egohmr_b200 — B200-native (sm_100a) implementation of EgoHMR's diffusion-sampling hot path.

Public surface mirrors the reference's: `create_gaussian_diffusion`, `EgoHMR`, `smpl.create` / `SMPL`,
`utils.geometry.rot6d_to_rotmat`.  See DESIGN.md and include/egohmr_b200.h.
"""
from .diffusion.model_util import create_gaussian_diffusion  # noqa: F401
from .models.egohmr.egohmr import EgoHMR  # noqa: F401
from . import smpl  # noqa: F401

"""wannier.jl_b200 — B200-native (sm_100a) implementation of Wannier.jl's spread / gradient hot
path behind the reference's own function API.  See DESIGN.md / INTEGRATION.md.

The directory name contains a dot, so the package is loaded by path (see
``__graft_entry__.load_package``) under the module name ``wannier_jl_b200``.
"""
from . import lib, api, synthetic, sharding, io_w90  # noqa: F401
from .io_w90 import read_w90  # noqa: F401
from .api import (KspaceStencil, Model, Spread, SpreadCenter, SpreadPenalty, CenterSpreadPenalty,  # noqa: F401
                  Cache, omega, omega_grad, omega_center, center, compute_MU_UtMU, transform_gauge,
                  get_fg_maxloc, max_localize, get_fg_disentangle, disentangle, X_Y_to_U, X_Y_to_XY,
                  XY_to_X_Y, U_to_X_Y, orthonorm_freeze, omega_local, position_op, compute_berry_connection_kspace,
                  get_fg_rotate, opt_rotate, merge_gauge, MagModel, SpreadMag, MagCache,
                  overlap_updn, omega_updn, omega_updn_grad)
from .lib import Context, MultiContext, MagContext, StiefelContext, WB200Error  # noqa: F401

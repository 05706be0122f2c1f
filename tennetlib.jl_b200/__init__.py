"""tnl_b200: B200-native hot path of TenNetLib.jl's DMRG sweeps behind the reference's own API names.

Product code only: CUDA kernels + C ABI in `csrc/` (libtnl_b200.so) and the host-side mirror of the reference
interface (`StateEnvs` on an MPO / a vector of MPOs / a `CouplingModel`, `eig_solver`, `exp_solver`,
`update_position`, `fullsweep`, `dynamic_fullsweep`, `DMRGParams`, `dmrg_`/`dmrg1`/`dmrg2`, `TDVPEngine`, `tdvpsweep`).
Nothing here imports `oracle/`; there is no CPU fallback."""
from ._lib import EXPORTED, TnlError, load, so_path
from .couplingmodel import CouplingModel
from .dmrg import DMRGParams, dmrg, dmrg1, dmrg2, dmrg_
from .solver import eig_solver, exp_solver
from .state_envs import StateEnvs
from .sweep import SweepData, dynamic_fullsweep, fullsweep
from .tdvp import TDVPEngine, getenergy, getentropy, maxchi, sweepcount, tdvpsweep, totalerror
from .tensor import Context, DeviceTensor, HostTensor, Index
from .update_site import halfsweep_done, update_position

__all__ = ["Context", "DeviceTensor", "HostTensor", "Index", "StateEnvs", "eig_solver", "exp_solver",
           "update_position", "halfsweep_done", "fullsweep", "dynamic_fullsweep", "SweepData", "DMRGParams", "dmrg_", "dmrg", "dmrg1", "dmrg2",
           "CouplingModel", "TDVPEngine", "tdvpsweep", "sweepcount", "getenergy", "getentropy", "maxchi", "totalerror", "load", "so_path", "TnlError", "EXPORTED"]

"""`TDVPEngine`, `tdvpsweep!` -- mirror of /root/reference/src/mps/tdvp.jl:17-117,247-277."""
from __future__ import annotations

from .solver import exp_solver
from .state_envs import StateEnvs
from .sweep import SweepData, dynamic_fullsweep, fullsweep


class TDVPEngine:
    """TDVPEngine(psi, H) / TDVPEngine(psi, H, Ms; weight): state + environments, sweep history, elapsed time.
    `rlim=None`: gauge of `psi` unknown, the first sweep orthogonalises it (pass llim / rlim for a canonical MPS)."""

    def __init__(self, ctx, psi, H, Ms=None, weight: float = -1.0, llim: int = 0, rlim: int | None = None):
        self.sysenv = StateEnvs(ctx, psi, H, llim=llim, rlim=rlim, Ms=Ms, weight=weight)
        self.swdata = SweepData()
        self.abstime = 0.0

    def getpsi(self):
        return self.sysenv.getpsi()


def sweepcount(engine: TDVPEngine) -> int:
    return engine.swdata.sweepcount


def getenergy(engine: TDVPEngine) -> float:
    return engine.swdata.energy[-1]


def getentropy(engine: TDVPEngine) -> float:
    return engine.swdata.entropy[-1]


def maxchi(engine: TDVPEngine) -> int:
    return engine.swdata.maxchi[-1]


def totalerror(engine: TDVPEngine) -> float:
    return sum(engine.swdata.maxtruncerr)


def tdvpsweep(engine: TDVPEngine, time_step, nsite=2, solver=exp_solver, **kwargs):
    """`tdvpsweep!`: psi' = exp(time_step * H) psi by one second-order sweep (half a step left-to-right, half a
    step right-to-left, backward evolutions of the centre in between).  `nsite="dynamic"` mixes one- and two-site
    updates bond by bond (`dynamic_fullsweep!`); on a single MPO it starts with the Global Subspace Expansion
    (`krylov_extend!`, gse.py)."""
    if solver is not exp_solver:
        raise RuntimeError("`tdvpsweep!()`: `solver` must be `exp_solver` !!")
    if nsite == "dynamic":
        dynamic_fullsweep(engine.sysenv, solver, engine.swdata, time_step=0.5 * time_step, reverse_step=True, **kwargs)
    elif nsite in (1, 2):
        if kwargs.pop("extendat", None) is not None:
            raise RuntimeError("`tdvpsweep!()`: `extendat` must be `nothing` for `nsite == 2` or `nsite == 1`. \n"
                               " Manually call `krylov_extend!(engine.sysenv; kwargs...)` for global subspace expansion.")
        fullsweep(engine.sysenv, solver, nsite, engine.swdata, time_step=0.5 * time_step, reverse_step=True, **kwargs)
    else:
        raise RuntimeError('`tdvpsweep!()`: `nsite` must be `"dynamic"`, `2`, or `1` !!')
    engine.abstime += abs(time_step)

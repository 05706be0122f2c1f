"""`DMRGParams`, `dmrg!`, `dmrg`, `dmrg2` -- mirror of /root/reference/src/mps/dmrg.jl:26-33,72-102,148-320."""
from __future__ import annotations

from .solver import eig_solver
from .state_envs import StateEnvs
from .sweep import SweepData, fullsweep
from .update_site import FLOAT64_THRESHOLD


class DMRGParams:
    def __init__(self, *, maxdim, nsweeps, cutoff=FLOAT64_THRESHOLD, noise=0.0, noisedecay=1.0, disable_noise_after=-1):
        n = len(nsweeps)

        def vec(x, T):
            return [T(v) for v in x] if isinstance(x, (list, tuple)) else [T(x)] * n
        self.maxdim = list(maxdim)
        self.nsweeps = list(nsweeps)
        self.cutoff = vec(cutoff, float)
        self.noise = vec(noise, float)
        self.noisedecay = vec(noisedecay, float)
        self.disable_noise_after = vec(disable_noise_after, int)
        if not (len(self.maxdim) == n == len(self.cutoff) == len(self.noise) == len(self.noisedecay)
                == len(self.disable_noise_after)):
            raise ValueError("`DMRGParams()` :: Size mismatch in input vectors !! \n"
                             "Lengths of `maxdim` and `nsweeps` must be same !!")


def dmrg_(sysenv: StateEnvs, params: DMRGParams, nsite: int, **kwargs) -> SweepData:
    """`dmrg!`."""
    outputlevel = kwargs.get("outputlevel", 1)
    enerrgoal = kwargs.pop("energyErrGoal", None)
    enterrgoal = kwargs.pop("entropyErrGoal", None)
    swdata = SweepData()
    for ii in range(len(params.nsweeps)):
        errGoalMet = False
        maxdim, cutoff, noise = params.maxdim[ii], params.cutoff[ii], params.noise[ii]
        noisedecay, disable_noise_after = params.noisedecay[ii], params.disable_noise_after[ii]
        if outputlevel > 0:
            print(f"DMRG level={ii + 1} => maxdim={maxdim}, nsweeps={params.nsweeps[ii]}, cutoff={cutoff:.2E}")
            print(f"DMRG level={ii + 1} => noise={noise:.2E}, noisedecay={noisedecay:.3f}, "
                  f"disable_noise_after={disable_noise_after}", flush=True)
        for jj in range(1, params.nsweeps[ii] + 1):
            enerr, enterr = fullsweep(sysenv, eig_solver, nsite, swdata, maxdim=maxdim, cutoff=cutoff, noise=noise,
                                      **kwargs)
            if enerrgoal is not None and enterrgoal is not None:
                errGoalMet = abs(enerr) < abs(enerrgoal) and abs(enterr) < abs(enterrgoal)
            elif enerrgoal is not None:
                errGoalMet = abs(enerr) < abs(enerrgoal)
            if errGoalMet and abs(noise) < FLOAT64_THRESHOLD:
                break
            if jj == disable_noise_after:
                noise = 0.0
            noise /= noisedecay
            if noise < 100 * FLOAT64_THRESHOLD:
                noise = 0.0
    return swdata


def dmrg(ctx, psi0, H, params: DMRGParams, nsite: int, Ms=None, **kwargs):
    """dmrg(psi0, H, params, nsite) and dmrg(psi0, H, Ms, params, nsite; weight) (src/mps/dmrg.jl:231-246).
    Host tensors carry no gauge information, so unless the caller states the orthogonality limits (`llim`, `rlim`)
    the gauge is taken as unknown and the first sweep right-canonicalises psi0 (`orthogonalize!(psi, 1)`,
    src/mps/sweep.jl:100-102) -- the reference reads the limits from the MPS object itself."""
    sysenv = StateEnvs(ctx, psi0, H, llim=kwargs.pop("llim", 0), rlim=kwargs.pop("rlim", None), Ms=Ms,
                       weight=kwargs.pop("weight", -1.0))
    swdata = dmrg_(sysenv, params, nsite, **kwargs)
    return swdata.energy[-1], sysenv, swdata


def dmrg2(ctx, psi0, H, params: DMRGParams, Ms=None, **kwargs):
    return dmrg(ctx, psi0, H, params, 2, Ms=Ms, **kwargs)


def dmrg1(ctx, psi0, H, params: DMRGParams, Ms=None, **kwargs):
    return dmrg(ctx, psi0, H, params, 1, Ms=Ms, **kwargs)

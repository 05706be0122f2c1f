"""TenNetLib sweep drivers on the CPU (oracle; test infrastructure only).

Line-by-line restatement (not a copy) of the control flow of
  /root/reference/src/mps/state_envs.jl:18-27,364-378   StateEnvs, position!, product
  /root/reference/src/base/solver.jl:23-43              eig_solver defaults
  /root/reference/src/mps/update_site.jl:13-90,231-277  halfsweep_done, _update_two_site!, update_position!
  /root/reference/src/mps/sweep.jl:26-196               SweepData, fullsweep!
  /root/reference/src/mps/dmrg.jl:26-227                DMRGParams, dmrg!
plus ITensorMPS `replacebond!` / `orthogonalize!` (third-party, restated from their published
behaviour: factorize(phi, inds(M[b]); ortho, which_decomp=nothing, eigen_perturbation), then
`M[b+1] ./= norm` (ortho left) or `M[b] ./= norm` (ortho right) when normalize).
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np

from .blocksparse import BSTensor, contract, factorize, inner
from .krylov import eigsolve_lanczos, exponentiate
from .projmpo import ProjMPO, ProjMPO_MPS2, ProjMPOSum2, drho_matrices

FLOAT64_THRESHOLD = 1e-15          # src/base/global_variables.jl:10-26


class MPS:
    """Tensors + orthogonality limits (ITensorMPS `MPS`: llim/rlim)."""

    def __init__(self, tensors: Sequence[BSTensor], llim: int = 0, rlim: int | None = None):
        self.t: List[BSTensor] = list(tensors)
        self.llim = llim
        self.rlim = len(self.t) + 1 if rlim is None else rlim

    def __len__(self):
        return len(self.t)

    def __getitem__(self, j):          # 1-based like the reference
        return self.t[j - 1]

    def __setitem__(self, j, v):
        self.t[j - 1] = v

    def isortho(self):
        return self.llim + 2 == self.rlim

    def orthocenter(self):
        assert self.isortho()
        return self.llim + 1

    def copy(self):
        return MPS([x.copy() for x in self.t], self.llim, self.rlim)


def orthogonalize(psi: MPS, j: int) -> MPS:
    """ITensorMPS `orthogonalize!(psi, j)`: gauge moves without truncation."""
    N = len(psi)
    while psi.llim < j - 1:
        b = psi.llim + 1
        A = psi[b]
        L, R, _, _ = factorize(A, A.inds[:2], ortho="left", which_decomp="svd", cutoff=None, maxdim=None)
        psi[b] = L
        psi[b + 1] = contract(R, psi[b + 1])
        psi.llim = b
        if psi.rlim < b + 2:
            psi.rlim = b + 2
    while psi.rlim > j + 1:
        b = psi.rlim - 1
        A = psi[b]
        L, R, _, _ = factorize(A, A.inds[:1], ortho="right", which_decomp="svd", cutoff=None, maxdim=None)
        psi[b] = R
        psi[b - 1] = contract(psi[b - 1], L)
        psi.rlim = b
        if psi.llim > b - 2:
            psi.llim = b - 2
    return psi


def replacebond(psi: MPS, b: int, phi: BSTensor, *, maxdim, mindim, cutoff, eigen_perturbation, ortho,
                normalize, which_decomp=None, svd_alg="divide_and_conquer"):
    left = [ix for ix in psi[b].inds if ix in phi.inds]
    tags = psi[b].inds[2].tags
    L, R, spec, _ = factorize(phi, left, ortho=ortho, maxdim=maxdim, mindim=mindim, cutoff=cutoff,
                              eigen_perturbation=eigen_perturbation, which_decomp=which_decomp, tags=tags)
    psi[b] = L
    psi[b + 1] = R
    if ortho == "left":
        if psi.llim == b - 1:
            psi.llim += 1
        if psi.rlim == b + 1:
            psi.rlim += 1
        if normalize:
            psi[b + 1] = psi[b + 1].scale(1.0 / psi[b + 1].norm())
    elif ortho == "right":
        if psi.llim == b:
            psi.llim -= 1
        if psi.rlim == b + 2:
            psi.rlim -= 1
        if normalize:
            psi[b] = psi[b].scale(1.0 / psi[b].norm())
    else:
        raise ValueError(ortho)
    return spec


class StateEnvs:
    """src/mps/state_envs.jl:18-27 with PH = ProjMPO (constructor :54-60 copies psi)."""

    def __init__(self, psi: MPS, H: Sequence[BSTensor], Ms=None, weight: float = -1.0):
        self.psi = psi.copy()
        # StateEnvs(psi, H, Ms; weight)  (src/mps/state_envs.jl:86-103) when penalised states are given
        from .couplingmodel import CouplingModel, ProjCouplingModel
        Mt = [m.t if isinstance(m, MPS) else m for m in Ms] if Ms else None
        if isinstance(H, CouplingModel):
            # StateEnvs(psi, H::CouplingModel[, Ms; weight]) (src/mps/state_envs.jl:73-79,131-150)
            self.PH = ProjCouplingModel(H) if not Ms else ProjMPO_MPS2(None, Mt, weight, PH=ProjCouplingModel(H))
        elif len(H) and isinstance(H[0], (list, tuple)):
            # StateEnvs(psi, Hs::Vector{MPO}[, Ms; weight]) (src/mps/state_envs.jl:63-70,106-128)
            self.PH = ProjMPOSum2(H) if not Ms else ProjMPO_MPS2(None, Mt, weight, PH=ProjMPOSum2(H))
        else:
            self.PH = ProjMPO(H) if not Ms else ProjMPO_MPS2(H, [m.t if isinstance(m, MPS) else m for m in Ms], weight)

    def __len__(self):
        return len(self.psi)

    def nsite(self):
        return self.PH.nsite

    def set_nsite(self, n):
        self.PH.set_nsite(n)

    def position(self, pos):
        self.PH.position(self.psi.t, pos)

    def product(self, v):
        return self.PH.product(v)

    __call__ = product


def eig_solver(env, phi0, time_step=None, **kw):
    """src/base/solver.jl:23-43."""
    val, vec, info = eigsolve_lanczos(env, phi0,
                                      tol=kw.get("solver_tol", 1e-14),
                                      krylovdim=kw.get("solver_krylovdim", 5),
                                      maxiter=kw.get("solver_maxiter", 2),
                                      eager=kw.get("solver_eager", False),
                                      which=kw.get("solver_which_eigenvalue", "SR"))
    if kw.get("solver_check_convergence", False) and info["converged"] < 1:
        raise RuntimeError("`eig_solver()` not converged !!")
    return val, vec


def exp_solver(env, phi0, time_step, **kw):
    """src/base/solver.jl:66-88 (returns NaN as energy: the caller evaluates <phi|H|phi>)."""
    if time_step is None:
        raise RuntimeError("`exp_solver()` is not defined with `time_step=None` !!")
    psi, info = exponentiate(env, time_step, phi0,
                             tol=kw.get("solver_tol", 1e-12),
                             krylovdim=kw.get("solver_krylovdim", 30),
                             maxiter=kw.get("solver_maxiter", 100),
                             eager=kw.get("solver_eager", True))
    if kw.get("solver_check_convergence", False) and info["converged"] < 1:
        raise RuntimeError("`eig_solver()` not converged !!")          # message as in the reference (solver.jl:84)
    exp_solver.last_info = info
    return float("nan"), psi


def halfsweep_done(N, pos, nsite, ortho):
    if pos == 1 and ortho == "right":
        return True
    if pos == N and ortho == "left" and nsite == 1:
        return True
    if pos == N - 1 and ortho == "left" and nsite == 2:
        return True
    return False


def _kept_side(psi, b, phi, ortho):
    """Indices of the side of the bond whose density matrix is diagonalised, in the order `replacebond` /
    `factorize` group them: psi[b]'s indices for ortho left, the remaining indices of phi (in phi's order) else."""
    left = [ix for ix in psi[b].inds if ix in phi.inds]
    return left if ortho == "left" else [ix for ix in phi.inds if ix not in left]


def _update_two_site(sysenv: StateEnvs, solver, pos, ortho, time_step, normalize, maxdim, mindim, cutoff,
                     svd_alg, noise, reverse_step, **kw):
    psi = sysenv.psi
    assert 0 < pos < len(sysenv)
    assert (psi.orthocenter() == pos and ortho == "left") or (psi.orthocenter() == pos + 1 and ortho == "right")
    sysenv.set_nsite(2)
    phi = contract(psi[pos], psi[pos + 1])
    sysenv.position(pos)
    energy, phi = solver(sysenv, phi, time_step, **kw)
    if normalize:
        phi = phi.scale(1.0 / phi.norm())
    if np.isnan(energy):
        energy = float(np.real(inner(phi, sysenv.PH(phi))))
    drho = None
    if abs(noise) > FLOAT64_THRESHOLD:
        d = drho_matrices(sysenv.PH.noiseterm(phi, ortho), noise, _kept_side(psi, pos, phi, ortho))
        drho = d if ortho == "left" else {tuple(-x for x in q): M for q, M in d.items()}
    spec = replacebond(psi, pos, phi, maxdim=maxdim, mindim=mindim, cutoff=cutoff, eigen_perturbation=drho,
                       ortho=ortho, normalize=normalize, which_decomp=None, svd_alg=svd_alg)
    if reverse_step and not halfsweep_done(len(sysenv), pos, 2, ortho):
        # TDVP backward evolution of the new centre site (src/mps/update_site.jl:78-87)
        pos1 = pos + 1 if ortho == "left" else pos
        phi0 = psi[pos1]
        sysenv.set_nsite(1)
        sysenv.position(pos1)
        energy, phi0 = solver(sysenv, phi0, -time_step, **kw)
        if normalize:
            phi0 = phi0.scale(1.0 / phi0.norm())
        if np.isnan(energy):
            energy = float(np.real(inner(phi0, sysenv.PH(phi0))))
        psi[pos1] = phi0
    return energy, spec.truncerr, spec.eigs


def _update_one_site(sysenv: StateEnvs, solver, pos, ortho, time_step, normalize, maxdim, mindim, cutoff,
                     svd_alg, noise, reverse_step, **kw):
    """src/mps/update_site.jl:94-190."""
    psi = sysenv.psi
    assert 0 < pos <= len(sysenv)
    assert psi.orthocenter() == pos
    nsite = 1
    sysenv.set_nsite(nsite)
    phi = psi[pos]
    sysenv.position(pos)
    energy, phi = solver(sysenv, phi, time_step, **kw)
    if normalize:
        phi = phi.scale(1.0 / phi.norm())
    if np.isnan(energy):
        energy = float(np.real(inner(phi, sysenv.PH(phi))))
    truncerr, eigs = 0.0, np.zeros(0)
    if halfsweep_done(len(sysenv), pos, nsite, ortho):
        psi[pos] = phi
        return energy, truncerr, eigs
    posnext = pos + 1 if ortho == "left" else pos - 1
    pos0 = pos if ortho == "left" else pos - 1
    if abs(noise) > FLOAT64_THRESHOLD:
        phi2 = contract(phi, psi[posnext]) if ortho == "left" else contract(psi[posnext], phi)
        psi[pos] = phi                               # keeps the index bookkeeping of psi[pos0] for replacebond
        sysenv.set_nsite(2)
        sysenv.position(pos0)
        d = drho_matrices(sysenv.PH.noiseterm(phi2, ortho), noise, _kept_side(psi, pos0, phi2, ortho))
        drho = d if ortho == "left" else {tuple(-x for x in q): M for q, M in d.items()}
        spec = replacebond(psi, pos0, phi2, maxdim=maxdim, mindim=mindim, cutoff=cutoff, eigen_perturbation=drho,
                           ortho=ortho, normalize=normalize, which_decomp=None, svd_alg=svd_alg)
        sysenv.set_nsite(1)
        return energy, spec.truncerr, spec.eigs
    nxt = psi[posnext]
    uinds = [ix for ix in phi.inds if ix not in nxt.inds]
    if ortho == "left":
        U, R, spec, u = factorize(phi, uinds, ortho="left", maxdim=maxdim, mindim=mindim, cutoff=cutoff,
                                  which_decomp="svd", tags=psi[pos].inds[2].tags)           # U, S*V
        if normalize:
            R = R.scale(1.0 / R.norm())                                                    # normalize!(S)
        psi[pos] = U
        psi.llim, psi.rlim = pos, pos + 2
        if reverse_step:
            energy, R = _reverse_zero_site(sysenv, solver, R, pos + 1, time_step, normalize, **kw)
        psi[posnext] = contract(R, nxt)
    else:
        left = [ix for ix in phi.inds if ix not in uinds]
        L, V, spec, u = factorize(phi, left, ortho="right", maxdim=maxdim, mindim=mindim, cutoff=cutoff,
                                  which_decomp="svd", tags=psi[pos].inds[0].tags)           # U*S, V
        if normalize:
            L = L.scale(1.0 / L.norm())
        psi[pos] = V
        psi.llim, psi.rlim = pos - 2, pos
        if reverse_step:
            energy, L = _reverse_zero_site(sysenv, solver, L, pos, time_step, normalize, **kw)
        psi[posnext] = contract(nxt, L)
    return energy, spec.truncerr, spec.eigs


def _reverse_zero_site(sysenv, solver, phi0, pos1, time_step, normalize, **kw):
    """TDVP backward evolution of the bond matrix (src/mps/update_site.jl:178-185)."""
    sysenv.set_nsite(0)
    sysenv.position(pos1)
    energy, phi0 = solver(sysenv, phi0, -time_step, **kw)
    if normalize:
        phi0 = phi0.scale(1.0 / phi0.norm())
    if np.isnan(energy):
        energy = float(np.real(inner(phi0, sysenv.PH(phi0))))
    return energy, phi0


def update_position(sysenv: StateEnvs, solver, pos, nsite, ortho, **kw):
    """src/mps/update_site.jl:231-277."""
    time_step = kw.get("time_step", None)
    normalize = kw.get("normalize", True)
    maxdim = kw.get("maxdim", None)
    mindim = kw.get("mindim", 1)
    cutoff = kw.get("cutoff", FLOAT64_THRESHOLD)
    svd_alg = kw.get("svd_alg", "divide_and_conquer")
    noise = kw.get("noise", 0.0)
    reverse_step = kw.get("reverse_step", time_step is not None)
    if noise > 0 and reverse_step:
        raise RuntimeError(f"`updatePosition()` :: `noise={noise}` cannot be greater than zero "
                           f"for `reverse_step={reverse_step}` !!")
    kw2 = {k: v for k, v in kw.items() if k not in ("time_step", "normalize", "maxdim", "mindim", "cutoff",
                                                   "svd_alg", "noise", "reverse_step")}
    if nsite == 2:
        return _update_two_site(sysenv, solver, pos, ortho, time_step, normalize, maxdim, mindim, cutoff,
                                svd_alg, noise, reverse_step, **kw2)
    if nsite == 1:
        return _update_one_site(sysenv, solver, pos, ortho, time_step, normalize, maxdim, mindim, cutoff,
                                svd_alg, noise, reverse_step, **kw2)
    raise NotImplementedError(f"`update_position()` with `nsite={nsite}` not implemented !!")


class SweepData:
    def __init__(self):
        self.sweepcount = 0
        self.maxchi: List[int] = []
        self.energy: List[float] = []
        self.entropy: List[float] = []
        self.maxtruncerr: List[float] = []
        self.lasteigs: List[np.ndarray] = []


def _entropy(p):
    """src/mps/measure.jl:4-20 (von Neumann entropy of a normalised spectrum)."""
    p = np.asarray(p, dtype=np.float64)
    p = p[p > 0]
    return float(-np.sum(p * np.log(p)))


def fullsweep(sysenv: StateEnvs, solver, nsite: int, swdata: SweepData, **kw):
    """src/mps/sweep.jl:94-196."""
    psi = sysenv.psi
    if (not psi.isortho()) or psi.orthocenter() != 1:
        orthogonalize(psi, 1)
    energy = np.nan
    maxtruncerr = 0.0
    swdata.sweepcount += 1
    N = len(sysenv)
    lasteigs = [None] * (N - 1)
    for bond in range(1, N):
        energy, err, _ = update_position(sysenv, solver, bond, nsite, "left", **kw)
        maxtruncerr = max(err, maxtruncerr)
        if nsite == 1 and bond == N - 1:
            energy, _, _ = update_position(sysenv, solver, bond + 1, nsite, "left", **kw)
    for bond in range(N - 1, 0, -1):
        site = bond + 1 if nsite == 1 else bond
        energy, err, eigs = update_position(sysenv, solver, site, nsite, "right", **kw)
        lasteigs[bond - 1] = eigs
        maxtruncerr = max(err, maxtruncerr)
        if nsite == 1 and bond == 1:
            energy, _, _ = update_position(sysenv, solver, bond, nsite, "right", **kw)
    swdata.lasteigs = lasteigs
    swdata.maxchi.append(max(A.inds[2].dim for A in psi.t[:-1]))
    swdata.energy.append(energy)
    mideigs = lasteigs[N // 2 - 1]
    swdata.entropy.append(_entropy(mideigs / np.sum(mideigs)))
    swdata.maxtruncerr.append(maxtruncerr)
    if swdata.sweepcount > 1:
        return swdata.energy[-1] - swdata.energy[-2], swdata.entropy[-1] - swdata.entropy[-2]
    return np.nan, np.nan


def dynamic_fullsweep(sysenv: StateEnvs, solver, swdata: SweepData, eigthreshold: float = 1e-12, extendat=None, **kw):
    """src/mps/sweep.jl:257-382: bond by bond one-site update where the smallest kept Schmidt weight of the
    previous half sweep is below `eigthreshold` or the bond is saturated at `maxdim`, two-site update otherwise.
    The first sweep (and every `extendat`-th) is a Global Subspace Expansion (oracle/gse.py) followed by a one-site
    sweep for StateEnvs{ProjMPO}, a plain two-site sweep for every other PH."""
    from .projmpo import ProjMPO as _ProjMPO
    maxdim = kw.get("maxdim", None)
    first = swdata.sweepcount == 0 or (extendat is not None and (swdata.sweepcount + 1) % extendat == 0)
    if first:
        if type(sysenv.PH) is _ProjMPO:
            from .gse import krylov_extend
            krylov_extend(sysenv, **kw)                       # Global Subspace Expansion, then a pure one-site sweep
            return fullsweep(sysenv, solver, 1, swdata, **kw)
        return fullsweep(sysenv, solver, 2, swdata, **kw)
    psi = sysenv.psi
    if (not psi.isortho()) or psi.orthocenter() != 1:
        orthogonalize(psi, 1)
    energy = np.nan
    maxtruncerr = 0.0
    swdata.sweepcount += 1
    N = len(sysenv)
    lasteigs = [None] * (N - 1)
    big = (1 << 62) if maxdim is None else maxdim

    def pick(bond):
        return 1 if (swdata.lasteigs[bond - 1][-1] < eigthreshold or psi[bond].inds[2].dim >= big) else 2
    for bond in range(1, N):
        nsite = pick(bond)
        energy, err, eigs = update_position(sysenv, solver, bond, nsite, "left", **kw)
        lasteigs[bond - 1] = eigs
        maxtruncerr = max(err, maxtruncerr)
        if nsite == 1 and bond == N - 1:
            energy, _, _ = update_position(sysenv, solver, bond + 1, nsite, "left", **kw)
    swdata.lasteigs = lasteigs
    lasteigs = list(lasteigs)
    for bond in range(N - 1, 0, -1):
        nsite = pick(bond)
        site = bond + 1 if nsite == 1 else bond
        energy, err, eigs = update_position(sysenv, solver, site, nsite, "right", **kw)
        lasteigs[bond - 1] = eigs
        maxtruncerr = max(err, maxtruncerr)
        if nsite == 1 and bond == 1:
            energy, _, _ = update_position(sysenv, solver, bond, nsite, "right", **kw)
    swdata.lasteigs = lasteigs
    swdata.maxchi.append(max(A.inds[2].dim for A in psi.t[:-1]))
    swdata.energy.append(energy)
    mideigs = lasteigs[N // 2 - 1]
    swdata.entropy.append(_entropy(mideigs / np.sum(mideigs)))
    swdata.maxtruncerr.append(maxtruncerr)
    if swdata.sweepcount > 1:
        return swdata.energy[-1] - swdata.energy[-2], swdata.entropy[-1] - swdata.entropy[-2]
    return np.nan, np.nan


class DMRGParams:
    """src/mps/dmrg.jl:26-33,72-102."""

    def __init__(self, *, maxdim, nsweeps, cutoff=FLOAT64_THRESHOLD, noise=0.0, noisedecay=1.0,
                 disable_noise_after=-1):
        n = len(nsweeps)
        vec = lambda x, T: [T(v) for v in x] if isinstance(x, (list, tuple)) else [T(x)] * n
        self.maxdim = list(maxdim)
        self.nsweeps = list(nsweeps)
        self.cutoff = vec(cutoff, float)
        self.noise = vec(noise, float)
        self.noisedecay = vec(noisedecay, float)
        self.disable_noise_after = vec(disable_noise_after, int)
        if not (len(self.maxdim) == n == len(self.cutoff) == len(self.noise) == len(self.noisedecay)
                == len(self.disable_noise_after)):
            raise ValueError("`DMRGParams()` :: Size mismatch in input vectors !!")


def dmrg_(sysenv: StateEnvs, params: DMRGParams, nsite: int, **kw) -> SweepData:
    """`dmrg!` src/mps/dmrg.jl:148-227 (stage / sweep loop, noise decay; error goals included)."""
    enerrgoal = kw.pop("energyErrGoal", None)
    enterrgoal = kw.pop("entropyErrGoal", None)
    kw.pop("outputlevel", None)
    swdata = SweepData()
    for ii in range(len(params.nsweeps)):
        errGoalMet = False
        maxdim, cutoff, noise = params.maxdim[ii], params.cutoff[ii], params.noise[ii]
        noisedecay, disable_after = params.noisedecay[ii], params.disable_noise_after[ii]
        for jj in range(1, params.nsweeps[ii] + 1):
            enerr, enterr = fullsweep(sysenv, eig_solver, nsite, swdata, maxdim=maxdim, cutoff=cutoff,
                                      noise=noise, **kw)
            if enerrgoal is not None and enterrgoal is not None:
                errGoalMet = abs(enerr) < abs(enerrgoal) and abs(enterr) < abs(enterrgoal)
            elif enerrgoal is not None:
                errGoalMet = abs(enerr) < abs(enerrgoal)
            # (entropy-only goal: result discarded by the reference, dmrg.jl:191)
            if errGoalMet and abs(noise) < FLOAT64_THRESHOLD:
                break
            if jj == disable_after:
                noise = 0.0
            noise /= noisedecay
            if noise < 100 * FLOAT64_THRESHOLD:
                noise = 0.0
    return swdata


def dmrg2(psi0: MPS, H, params: DMRGParams, Ms=None, **kw):
    sysenv = StateEnvs(psi0, H, Ms, kw.pop("weight", -1.0))
    sw = dmrg_(sysenv, params, 2, **kw)
    return sw.energy[-1], sysenv.psi, sw


def dmrg1(psi0: MPS, H, params: DMRGParams, **kw):
    sysenv = StateEnvs(psi0, H)
    sw = dmrg_(sysenv, params, 1, **kw)
    return sw.energy[-1], sysenv.psi, sw


class TDVPEngine:
    """src/mps/tdvp.jl:17-66."""

    def __init__(self, psi: MPS, H, Ms=None, weight: float = -1.0):
        self.sysenv = StateEnvs(psi, H, Ms, weight)
        self.swdata = SweepData()
        self.abstime = 0.0

    def getpsi(self):
        return self.sysenv.psi.copy()


def tdvpsweep(engine: TDVPEngine, time_step, nsite=2, solver=exp_solver, **kw):
    """`tdvpsweep!` src/mps/tdvp.jl:247-277: psi' = exp(time_step * H) psi, second-order sweep
    (half a step left-to-right, half a step right-to-left, backward steps in between)."""
    if solver is not exp_solver:
        raise RuntimeError("`tdvpsweep!()`: `solver` must be `exp_solver` !!")
    if nsite in (1, 2):
        if kw.get("extendat", None) is not None:
            raise RuntimeError("`tdvpsweep!()`: `extendat` must be `nothing` for `nsite == 2` or `nsite == 1`.")
        kw.pop("extendat", None)
        fullsweep(engine.sysenv, solver, nsite, engine.swdata, time_step=0.5 * time_step, reverse_step=True, **kw)
    elif nsite == "dynamic":
        dynamic_fullsweep(engine.sysenv, solver, engine.swdata, time_step=0.5 * time_step, reverse_step=True, **kw)
    else:
        raise RuntimeError('`tdvpsweep!()`: `nsite` must be `"dynamic"`, `2`, or `1` !!')
    engine.abstime += abs(time_step)

"""Local updates -- mirror of /root/reference/src/mps/update_site.jl:13-277 (two-site and one-site updates,
DMRG and TDVP branches)."""
from __future__ import annotations

import math

import numpy as np

FLOAT64_THRESHOLD = 1e-15      # src/base/global_variables.jl:10-26


def halfsweep_done(N: int, pos: int, nsite: int, ortho: str) -> bool:
    if pos == 1 and ortho == "right":
        return True
    if pos == N and ortho == "left" and nsite == 1:
        return True
    if pos == N - 1 and ortho == "left" and nsite == 2:
        return True
    return False


def _update_two_site(sysenv, solver, pos, ortho, time_step, normalize, maxdim, mindim, cutoff, svd_alg, noise,
                     reverse_step, **kwargs):
    assert 0 < pos < len(sysenv)
    assert (sysenv.orthocenter() == pos and ortho == "left") or (sysenv.orthocenter() == pos + 1 and ortho == "right")
    nsite = 2
    sysenv.set_nsite(nsite)
    with sysenv.phase("make_phi"):
        phi = sysenv.make_phi(pos)                   # psi[pos] * psi[pos+1]
    with sysenv.phase("position"):
        sysenv.position(pos)
    with sysenv.phase("solver"):
        energy, phi = solver(sysenv, phi, time_step, **kwargs)
    if normalize:
        phi.scale_(1.0 / phi.norm())
    if isinstance(energy, float) and math.isnan(energy):
        energy = sysenv.expectation(phi)
    drho_noise = noise if abs(noise) > FLOAT64_THRESHOLD else 0.0
    with sysenv.phase("replacebond"):
        truncerr, eigs = sysenv.replacebond(pos, phi, maxdim=maxdim, mindim=mindim, cutoff=cutoff, noise=drho_noise,
                                            ortho=ortho, normalize=normalize, which_decomp=kwargs.get("which_decomp"),
                                            svd_alg=svd_alg)
    if reverse_step and not halfsweep_done(len(sysenv), pos, nsite, ortho):
        # TDVP backward evolution of the new centre site (update_site.jl:78-87)
        pos1 = pos + 1 if ortho == "left" else pos
        phi0 = sysenv.site_tensor(pos1).copy()
        sysenv.set_nsite(nsite - 1)
        with sysenv.phase("position"):
            sysenv.position(pos1)
        with sysenv.phase("solver"):
            energy, phi0 = solver(sysenv, phi0, -time_step, **kwargs)
        if normalize:
            phi0.scale_(1.0 / phi0.norm())
        if isinstance(energy, float) and math.isnan(energy):
            energy = sysenv.expectation(phi0)
        sysenv.set_site_tensor(pos1, phi0)
    return energy, truncerr, eigs


def _update_one_site(sysenv, solver, pos, ortho, time_step, normalize, maxdim, mindim, cutoff, svd_alg, noise,
                     reverse_step, **kwargs):
    """src/mps/update_site.jl:94-190 (DMRG branches: noise -> two-site replacebond!, else svd split)."""
    assert 0 < pos <= len(sysenv)
    assert sysenv.orthocenter() == pos
    nsite = 1
    sysenv.set_nsite(nsite)
    phi = sysenv.site_tensor(pos).copy()
    with sysenv.phase("position"):
        sysenv.position(pos)
    with sysenv.phase("solver"):
        energy, phi = solver(sysenv, phi, time_step, **kwargs)
    if normalize:
        phi.scale_(1.0 / phi.norm())
    if isinstance(energy, float) and math.isnan(energy):
        energy = sysenv.expectation(phi)
    truncerr, eigs = 0.0, np.zeros(0)
    if halfsweep_done(len(sysenv), pos, nsite, ortho):
        sysenv.set_site_tensor(pos, phi)
        return energy, truncerr, eigs
    pos0 = pos if ortho == "left" else pos - 1
    if abs(noise) > FLOAT64_THRESHOLD:
        sysenv.set_site_tensor(pos, phi)             # phi *= psi[posnext]  is formed on the device below
        sysenv.set_nsite(nsite + 1)
        phi2 = sysenv.make_phi(pos0)
        sysenv.position(pos0)
        with sysenv.phase("replacebond"):
            truncerr, eigs = sysenv.replacebond(pos0, phi2, maxdim=maxdim, mindim=mindim, cutoff=cutoff, noise=noise,
                                                ortho=ortho, normalize=normalize, which_decomp=None, svd_alg=svd_alg)
        sysenv.set_nsite(nsite)
        return energy, truncerr, eigs
    if not reverse_step:
        with sysenv.phase("replacebond"):
            truncerr, eigs = sysenv.svd_split(pos, phi, maxdim=maxdim, mindim=mindim, cutoff=cutoff, ortho=ortho,
                                              normalize=normalize, svd_alg=svd_alg)
        return energy, truncerr, eigs
    # TDVP (update_site.jl:158-186): psi[pos] = U, phi0 = S*V evolved backwards with the zero-site H_eff, then absorbed
    with sysenv.phase("replacebond"):
        truncerr, eigs, phi0 = sysenv.svd_split(pos, phi, maxdim=maxdim, mindim=mindim, cutoff=cutoff, ortho=ortho,
                                                normalize=normalize, svd_alg=svd_alg, absorb=False)
    pos1 = pos + 1 if ortho == "left" else pos
    sysenv.set_nsite(nsite - 1)
    with sysenv.phase("position"):
        sysenv.position(pos1)
    with sysenv.phase("solver"):
        energy, phi0 = solver(sysenv, phi0, -time_step, **kwargs)
    if normalize:
        phi0.scale_(1.0 / phi0.norm())
    if isinstance(energy, float) and math.isnan(energy):
        energy = sysenv.expectation(phi0)
    sysenv.absorb_bond(pos, ortho, phi0)
    return energy, truncerr, eigs


def update_position(sysenv, solver, pos: int, nsite: int, ortho: str, **kwargs):
    time_step = kwargs.get("time_step", None)
    normalize = kwargs.get("normalize", True)
    maxdim = kwargs.get("maxdim", None)
    mindim = kwargs.get("mindim", 1)
    cutoff = kwargs.get("cutoff", FLOAT64_THRESHOLD)
    svd_alg = kwargs.get("svd_alg", "divide_and_conquer")
    noise = kwargs.get("noise", 0.0)
    reverse_step = kwargs.get("reverse_step", time_step is not None)
    if noise > 0 and reverse_step:
        raise RuntimeError(f"`updatePosition()` :: `noise={noise}` cannot be greater than zero"
                           f" for `reverse_step={reverse_step}` !!")
    rest = {k: v for k, v in kwargs.items() if k not in ("time_step", "normalize", "maxdim", "mindim", "cutoff",
                                                          "svd_alg", "noise", "reverse_step")}
    if nsite == 2:
        return _update_two_site(sysenv, solver, pos, ortho, time_step, normalize, maxdim, mindim, cutoff, svd_alg,
                                noise, reverse_step, **rest)
    if nsite == 1:
        return _update_one_site(sysenv, solver, pos, ortho, time_step, normalize, maxdim, mindim, cutoff, svd_alg,
                                noise, reverse_step, **rest)
    raise NotImplementedError(f"`update_position()` with `nsite={nsite}` not implemented !!")

"""Sweep driver -- mirror of /root/reference/src/mps/sweep.jl:26-196 (`SweepData`, `fullsweep!`)."""
from __future__ import annotations

import time
from typing import List

import numpy as np

from .update_site import update_position


class SweepData:
    def __init__(self):
        self.sweepcount = 0
        self.maxchi: List[int] = []
        self.energy: List[float] = []
        self.entropy: List[float] = []
        self.maxtruncerr: List[float] = []
        self.lasteigs: List[np.ndarray] = []


def _entropy(p) -> float:
    p = np.asarray(p, dtype=np.float64)
    p = p[p > 0]
    return float(-np.sum(p * np.log(p)))


def fullsweep(sysenv, solver, nsite: int, swdata: SweepData, **kwargs):
    outputlevel = kwargs.pop("outputlevel", 1)
    noise = kwargs.get("noise", 0.0)
    if (not sysenv.isortho()) or sysenv.orthocenter() != 1:
        sysenv.orthogonalize1()
    energy = float("nan")
    maxtruncerr = 0.0
    swdata.sweepcount += 1
    N = len(sysenv)
    lasteigs = [None] * (N - 1)
    t0 = time.time()
    for bond in range(1, N):
        energy, err, _ = update_position(sysenv, solver, bond, nsite, "left", **kwargs)
        maxtruncerr = max(err, maxtruncerr)
        if outputlevel > 1:
            print(f"At left sweep {swdata.sweepcount} bond {bond} => Energy {energy}, Err {err:.2g}", flush=True)
        if nsite == 1 and bond == N - 1:
            energy, _, _ = update_position(sysenv, solver, bond + 1, nsite, "left", **kwargs)
    for bond in range(N - 1, 0, -1):
        site = bond + 1 if nsite == 1 else bond
        energy, err, eigs = update_position(sysenv, solver, site, nsite, "right", **kwargs)
        lasteigs[bond - 1] = eigs
        maxtruncerr = max(err, maxtruncerr)
        if outputlevel > 1:
            print(f"At right sweep {swdata.sweepcount} bond {bond} => Energy {energy}, Err {err:.2g}", flush=True)
        if nsite == 1 and bond == 1:
            energy, _, _ = update_position(sysenv, solver, bond, nsite, "right", **kwargs)
    sw_time = time.time() - t0
    swdata.lasteigs = lasteigs
    swdata.maxchi.append(max(sysenv.linkdims()))
    swdata.energy.append(energy)
    mideigs = lasteigs[N // 2 - 1]
    swdata.entropy.append(_entropy(mideigs / np.sum(mideigs)))
    swdata.maxtruncerr.append(maxtruncerr)
    if swdata.sweepcount > 1:
        enerr = swdata.energy[-1] - swdata.energy[-2]
        enterr = swdata.entropy[-1] - swdata.entropy[-2]
    else:
        enerr = enterr = float("nan")
    if outputlevel > 0:
        print(f"At sweep {swdata.sweepcount} => E={swdata.energy[-1]}, S={swdata.entropy[-1]}, "
              f"MaxLinkDim={swdata.maxchi[-1]}, Noise={noise:.2g}")
        print(f"At sweep {swdata.sweepcount} => dE={enerr}, dS={enterr}, MaxErr={maxtruncerr:.2g}, Time={sw_time:.3f}",
              flush=True)
    return enerr, enterr


def dynamic_fullsweep(sysenv, solver, swdata: SweepData, eigthreshold: float = 1e-12, extendat=None, **kwargs):
    """`dynamic_fullsweep!` (src/mps/sweep.jl:257-382): per bond a one-site update where the smallest kept Schmidt
    weight of the previous half sweep is below `eigthreshold` or the bond is saturated at `maxdim`, a two-site
    update otherwise.  The first sweep and every `extendat`-th one are plain two-site sweeps -- except for
    StateEnvs{ProjMPO}, where the reference does a Global Subspace Expansion followed by a one-site sweep
    (`krylov_extend!`, sweep.jl:266-282,399-555; here gse.py on the device)."""
    maxdim = kwargs.get("maxdim", None)
    outputlevel = kwargs.get("outputlevel", 1)
    noise = kwargs.get("noise", 0.0)
    if swdata.sweepcount == 0 or (extendat is not None and (swdata.sweepcount + 1) % extendat == 0):
        if sysenv.nterms == 1 and not sysenv.is_coupling_model and not sysenv.has_penalty:
            from .gse import krylov_extend
            krylov_extend(sysenv, **kwargs)
            return fullsweep(sysenv, solver, 1, swdata, **kwargs)
        return fullsweep(sysenv, solver, 2, swdata, **kwargs)
    kwargs.pop("outputlevel", None)
    if (not sysenv.isortho()) or sysenv.orthocenter() != 1:
        sysenv.orthogonalize1()
    energy = float("nan")
    maxtruncerr = 0.0
    swdata.sweepcount += 1
    N = len(sysenv)
    lasteigs = [None] * (N - 1)
    big = (1 << 62) if maxdim is None else maxdim

    def pick(bond):
        return 1 if (swdata.lasteigs[bond - 1][-1] < eigthreshold or sysenv.linkdim(bond) >= big) else 2
    t0 = time.time()
    for bond in range(1, N):
        nsite = pick(bond)
        energy, err, eigs = update_position(sysenv, solver, bond, nsite, "left", **kwargs)
        lasteigs[bond - 1] = eigs
        maxtruncerr = max(err, maxtruncerr)
        if outputlevel > 1:
            print(f"At left sweep {swdata.sweepcount} bond {bond} => Energy {energy}, Err {err:.2g}", flush=True)
        if nsite == 1 and bond == N - 1:
            energy, _, _ = update_position(sysenv, solver, bond + 1, nsite, "left", **kwargs)
    swdata.lasteigs = lasteigs
    lasteigs = list(lasteigs)
    for bond in range(N - 1, 0, -1):
        nsite = pick(bond)
        site = bond + 1 if nsite == 1 else bond
        energy, err, eigs = update_position(sysenv, solver, site, nsite, "right", **kwargs)
        lasteigs[bond - 1] = eigs
        maxtruncerr = max(err, maxtruncerr)
        if outputlevel > 1:
            print(f"At right sweep {swdata.sweepcount} bond {bond} => Energy {energy}, Err {err:.2g}", flush=True)
        if nsite == 1 and bond == 1:
            energy, _, _ = update_position(sysenv, solver, bond, nsite, "right", **kwargs)
    sw_time = time.time() - t0
    swdata.lasteigs = lasteigs
    swdata.maxchi.append(max(sysenv.linkdims()))
    swdata.energy.append(energy)
    mideigs = lasteigs[N // 2 - 1]
    swdata.entropy.append(_entropy(mideigs / np.sum(mideigs)))
    swdata.maxtruncerr.append(maxtruncerr)
    if swdata.sweepcount > 1:
        enerr = swdata.energy[-1] - swdata.energy[-2]
        enterr = swdata.entropy[-1] - swdata.entropy[-2]
    else:
        enerr = enterr = float("nan")
    if outputlevel > 0:
        print(f"At sweep {swdata.sweepcount} => E={swdata.energy[-1]}, S={swdata.entropy[-1]}, "
              f"MaxLinkDim={swdata.maxchi[-1]}, Noise={noise:.2g}")
        print(f"At sweep {swdata.sweepcount} => dE={enerr}, dS={enterr}, MaxErr={maxtruncerr:.2g}, Time={sw_time:.3f}",
              flush=True)
    return enerr, enterr

"""Generates tests/golden/*.json from the CPU oracle (run from the repo root: python tests/golden/make_golden.py).

The reference has no golden vectors and cannot run here (no Julia), so these fixtures pin the ORACLE: the CPU
suite checks the oracle still reproduces them, the GPU suite checks the CUDA path against them."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import blocksparse as ob, dmrg as od, krylov as ok, models as om  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def case_dmrg(kind, N, params, name, nsite=2):
    sites = om.siteinds(kind, N)
    H = om.heisenberg_mpo(sites)
    psi0 = od.MPS(om.neel_mps(sites))
    e, psi, sw = (od.dmrg2 if nsite == 2 else od.dmrg1)(psi0, H, od.DMRGParams(**params))
    return dict(name=name, kind=kind, N=N, params=params, energy=sw.energy, maxchi=sw.maxchi,
                maxtruncerr=sw.maxtruncerr, entropy=sw.entropy,
                linkdims=[A.inds[2].dim for A in psi.t[:-1]])


def case_excited(kind, N, params, weight):
    """Reference dmrg_ex test (test/test_MPS_DMRG.jl:68-97): ground state, then first excited state with the
    penalty weight*|gs><gs| (StateEnvs(psi0, H, [psi_gr]; weight))."""
    sites = om.siteinds(kind, N)
    H = om.heisenberg_mpo(sites)
    psi0 = od.MPS(om.neel_mps(sites))
    e0, gs, sw0 = od.dmrg2(psi0, H, od.DMRGParams(**params))
    e1, ex, sw1 = od.dmrg2(psi0, H, od.DMRGParams(**params), Ms=[gs], weight=weight)
    return dict(kind=kind, N=N, params=params, weight=weight, e0=e0, energy=sw1.energy, maxchi=sw1.maxchi)


def case_bond(kind, N, chi, seed, pos):
    """One bond of a seeded random MPS: H_eff apply, Lanczos and truncation invariants."""
    sites = om.siteinds(kind, N)
    H = om.heisenberg_mpo(sites)
    qn, dm = om.gaussian_link_sectors(chi, 1.3, 4, step=2 if kind == "S=1" else 1)
    mps = od.MPS(om.random_mps(sites, qn, dm, np.random.default_rng(seed)))
    od.orthogonalize(mps, pos)
    env = od.StateEnvs(mps, H)
    env.set_nsite(2)
    env.position(pos)
    phi = ob.contract(env.psi[pos], env.psi[pos + 1])
    phi = phi.scale(1 / phi.norm())
    ob.reset_flops()
    Hv = env.product(phi)
    flops = ob.get_flops()
    e, v, info = ok.eigsolve_lanczos(env, phi)
    out = dict(kind=kind, N=N, chi=chi, seed=seed, pos=pos, apply_flops_stored_blocks=flops,
               expectation=ob.inner(phi, Hv), Hv_norm=Hv.norm(), lanczos_energy=e, lanczos_numops=info["numops"],
               lanczos_normres=float(info["normres"]), trunc=[])
    for ortho in ("left", "right"):
        for maxdim, cutoff in ((8, 1e-14), (1000, 1e-6)):
            m2 = mps.copy()
            spec = od.replacebond(m2, pos, v.scale(1 / v.norm()), maxdim=maxdim, mindim=1, cutoff=cutoff,
                                  eigen_perturbation=None, ortho=ortho, normalize=True)
            link = m2[pos].inds[2]
            out["trunc"].append(dict(ortho=ortho, maxdim=maxdim, cutoff=cutoff, truncerr=spec.truncerr,
                                     eigs=[float(x) for x in spec.eigs], link_qns=[list(q) for q in link.qns],
                                     link_dims=list(link.dims)))
    return out


if __name__ == "__main__":
    ref = dict(nsweeps=[5], maxdim=[20], cutoff=1e-14, noise=1e-3, noisedecay=2, disable_noise_after=2)
    g = dict(
        dmrg=[case_dmrg("S=1/2", 12, ref, "reference test_MPS_DMRG.jl parameters"),
              case_dmrg("S=1/2", 12, dict(nsweeps=[6], maxdim=[64], cutoff=1e-14), "noise-free, exact bond dimension"),
              case_dmrg("S=1", 8, dict(nsweeps=[3, 4], maxdim=[30, 200], cutoff=1e-14, noise=[1e-4, 0.0]), "S=1 N=8")],
        dmrg1=[case_dmrg("S=1/2", 12, ref, "one-site DMRG, reference test parameters (dmrg_1, nsite=1)", nsite=1),
               case_dmrg("S=1", 8, dict(nsweeps=[4, 3], maxdim=[40, 40], cutoff=[1e-14, 0.0], noise=[1e-3, 0.0]),
                         "one-site S=1 N=8, noise then noise-free svd split", nsite=1)],
        excited=[case_excited("S=1/2", 12, ref, 10.0)],
        bond=[case_bond("S=1", 8, 30, 3, 4), case_bond("S=1/2", 10, 24, 5, 5)],
        ed=dict(S12_N12=-5.1420906328405, S1_N8=-10.1246372223589, S12_N20=-8.6824733343990,
                S12_N12_E1=-4.8611479370364))
    json.dump(g, open(os.path.join(OUT, "oracle_golden.json"), "w"), indent=1)
    print("wrote", os.path.join(OUT, "oracle_golden.json"))

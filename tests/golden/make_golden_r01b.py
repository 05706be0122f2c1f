"""Second set of oracle fixtures (python tests/golden/make_golden_r01b.py -> oracle_golden_r01b.json): the rows
added after the first fixture file -- exp_solver / TDVP (imaginary and real time), CouplingModel and MPO-sum DMRG,
dynamic TDVP, TTN optimisation.  As for oracle_golden.json these pin the ORACLE (the reference has no golden
vectors and cannot run here); the GPU suite compares the CUDA path with the live oracle on the same inputs."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import blocksparse as ob, couplingmodel as oc, dmrg as od, models as om, ttn as ot  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def _c(z):
    z = complex(z)
    return [z.real, z.imag]


def case_exp(kind, N, chi, seed, pos, nsite, t):
    """exp_solver on one site range of a seeded random MPS: operator count, norm and two overlaps of the result."""
    sites = om.siteinds(kind, N)
    H = om.heisenberg_mpo(sites)
    qn, dm = om.gaussian_link_sectors(chi, 1.3, 4, step=2 if kind == "S=1" else 1)
    mps = od.MPS(om.random_mps(sites, qn, dm, np.random.default_rng(seed)))
    od.orthogonalize(mps, pos)
    env = od.StateEnvs(mps, H)
    env.set_nsite(nsite)
    env.position(pos)
    phi = ob.contract(env.psi[pos], env.psi[pos + 1]) if nsite == 2 else env.psi[pos]
    _, out = od.exp_solver(env, phi, complex(*t) if t[1] else t[0])
    info = od.exp_solver.last_info
    return dict(kind=kind, N=N, chi=chi, seed=seed, pos=pos, nsite=nsite, t=t, numops=info["numops"],
                converged=info["converged"], norm=out.norm(), overlap_with_input=_c(ob.inner(phi, out)),
                expectation=_c(ob.inner(out, env.product(out))))


def case_tdvp(kind, N, dt, sched, maxdim, cutoff, model="mpo", **model_kw):
    """tdvpsweep! from the Neel state: energies, bond dimensions, truncation errors per sweep."""
    sites = om.siteinds(kind, N)
    H = om.heisenberg_mpo(sites) if model == "mpo" else oc.heisenberg_coupling_model(sites, **model_kw)
    eng = od.TDVPEngine(od.MPS(om.neel_mps(sites)), H)
    step = complex(*dt) if dt[1] else dt[0]
    for ns in sched:
        od.tdvpsweep(eng, step, ns, maxdim=maxdim, cutoff=cutoff, **({"extendat": 5} if ns == "dynamic" else {}))
    return dict(kind=kind, N=N, dt=dt, sched=sched, maxdim=maxdim, cutoff=cutoff, model=model, model_kw=model_kw,
                energy=[float(np.real(e)) for e in eng.swdata.energy], maxchi=eng.swdata.maxchi,
                maxtruncerr=eng.swdata.maxtruncerr, abstime=eng.abstime)


def case_dmrg_model(kind, N, params, model, **model_kw):
    sites = om.siteinds(kind, N)
    if model == "cm":
        H = oc.heisenberg_coupling_model(sites, **model_kw)
    else:
        H = [om.heisenberg_mpo(sites, Jz=1.0, Jxy=0.0), om.heisenberg_mpo(sites, Jz=0.0, Jxy=1.0)]
    e, psi, sw = od.dmrg2(od.MPS(om.neel_mps(sites)), H, od.DMRGParams(**params))
    return dict(kind=kind, N=N, params=params, model=model, model_kw=model_kw, energy=sw.energy, maxchi=sw.maxchi,
                maxtruncerr=sw.maxtruncerr)


def case_config0(params):
    """BASELINE.json configs[0]: S=1/2 Heisenberg chain N=20, two-site DMRG, maxdim 20 -> 64 (test_MPS_DMRG.jl scale)."""
    sites = om.siteinds("S=1/2", 20)
    H = om.heisenberg_mpo(sites)
    e, psi, sw = od.dmrg2(od.MPS(om.neel_mps(sites)), H, od.DMRGParams(**params))
    return dict(kind="S=1/2", N=20, params=params, energy=sw.energy, maxchi=sw.maxchi, maxtruncerr=sw.maxtruncerr,
                entropy=sw.entropy, linkdims=[A.inds[2].dim for A in psi.t[:-1]], ed_literature=-8.682473334399)


def case_ttn(N, h, chi0, seed, params):
    sites = ot.dense_siteinds(N)
    M = oc.tfi_coupling_model(sites, h=h)
    rng = np.random.default_rng(seed)
    psi0 = ot.default_random_ttn(sites, chi0, rng)
    E, psi, sw = ot.optimize(psi0, M, ot.OptimizeParamsTTN(**params), ot.default_sweeppath(psi0), rng=rng)
    E0 = float(np.linalg.eigvalsh(oc.coupling_model_to_dense(M))[0])
    return dict(N=N, h=h, chi0=chi0, seed=seed, params=params, energy=[float(e) for e in sw.energy], maxchi=sw.maxchi, ed=E0)


def build():
    noisy = dict(maxdim=[8, 20], nsweeps=[2, 3], cutoff=1e-14, noise=[1e-3, 0.0])
    return dict(
        exp=[case_exp("S=1", 8, 24, 3, 4, 2, [-0.1, 0.0]), case_exp("S=1", 8, 24, 3, 4, 1, [0.0, 0.05]),
             case_exp("S=1", 8, 24, 3, 4, 2, [-0.05, -0.08])],
        tdvp=[case_tdvp("S=1/2", 8, [-0.1, 0.0], [2, 2, 2, 1, 1], 16, 1e-12),
              case_tdvp("S=1/2", 8, [0.0, -0.05], [2, 2, 2, 1, 1], 16, 1e-13),
              case_tdvp("S=1/2", 8, [-0.05, 0.0], [2, 2, 1], 16, 1e-12, model="cm", merge=True, j2=0.3),
              case_tdvp("S=1/2", 8, [-0.1, 0.0], ["dynamic"] * 6, 6, 1e-12, model="cm", merge=True)],
        dmrg_models=[case_dmrg_model("S=1/2", 8, noisy, "cm", merge=True),
                     case_dmrg_model("S=1/2", 8, noisy, "cm", merge=False),
                     case_dmrg_model("S=1", 8, dict(maxdim=[10, 20], nsweeps=[2, 2], cutoff=1e-13, noise=[1e-4, 0.0]), "mposum")],
        config0=[case_config0(dict(nsweeps=[5, 5], maxdim=[20, 64], cutoff=1e-14, noise=[1e-3, 0.0], noisedecay=2,
                                   disable_noise_after=2))],
        ttn=[case_ttn(8, 1.0, 4, 1, dict(maxdim=[8, 16], nsweeps=[4, 3], cutoff=1e-14, noise=[1e-2, 0.0], noisedecay=5,
                                         disable_noise_after=3))],
    )


if __name__ == "__main__":
    g = build()
    with open(os.path.join(OUT, "oracle_golden_r01b.json"), "w") as f:
        json.dump(g, f, indent=1)
    print("wrote", os.path.join(OUT, "oracle_golden_r01b.json"))

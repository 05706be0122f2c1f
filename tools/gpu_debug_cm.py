import sys
sys.path.insert(0, ".")
import numpy as np
import tennetlib.jl_b200 as T
from oracle import blocksparse as ob, dmrg as od, models as om, couplingmodel as oc

ctx = T.Context(0)
N = 8
for kw in (dict(merge=True), dict(merge=True, j2=0.4), dict(merge=False), dict(merge=True, field=0.3)):
    sites = om.siteinds("S=1", N)
    M = oc.heisenberg_coupling_model(sites, **kw)
    qn, dm = om.gaussian_link_sectors(24, 1.3, 4, step=2)
    mps = od.MPS(om.random_mps(sites, qn, dm, np.random.default_rng(9)))
    od.orthogonalize(mps, 4)
    for ortho, pos in (("left", 4), ("right", 3)):
        for noise in (0.0, 1e-3):
            env_o = od.StateEnvs(mps, M)
            env_d = T.StateEnvs(ctx, mps.t, M, llim=3, rlim=5)
            try:
                eo, to, so = od.update_position(env_o, od.eig_solver, pos, 2, ortho, maxdim=20, cutoff=1e-13, noise=noise)
                ed, td, sd = T.update_position(env_d, T.eig_solver, pos, 2, ortho, maxdim=20, cutoff=1e-13, noise=noise)
                print(kw, ortho, noise, "dE", abs(ed - eo), "terr", td, to, "eigs", np.abs(np.array(sd) - np.array(so)).max() if len(sd) == len(so) else (len(sd), len(so)))
            except Exception as e:
                print(kw, ortho, noise, "EXC", repr(e)[:200])

"""Step-by-step oracle vs device comparison along DMRG sweeps (bring-up aid)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import tennetlib.jl_b200 as T
from oracle import blocksparse as ob, models as om, dmrg as od

ctx = T.Context()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
kind = "S=1/2"
sites = om.siteinds(kind, N)
H = om.heisenberg_mpo(sites)
def dense_state(tensors):
    acc = None
    for A in tensors:
        Ad = A.to_dense()
        acc = Ad[0] if acc is None else np.tensordot(acc, Ad, axes=([-1], [0]))
    return acc[..., 0].reshape(-1)


for noise in (1e-3,):
    print("\n##### noise", noise)
    psi0 = od.MPS(om.neel_mps(sites))
    env_o = od.StateEnvs(psi0, H)
    od.orthogonalize(env_o.psi, 1)
    env_d = T.StateEnvs(ctx, psi0.t, H, llim=0, rlim=2)
    kw = dict(maxdim=20, cutoff=1e-14, noise=noise)
    bad = 0
    for sweep in range(2):
        for ortho, bonds in (("left", range(1, N)), ("right", range(N - 1, 0, -1))):
            for b in bonds:
                eo, to, go = od.update_position(env_o, od.eig_solver, b, 2, ortho, **kw)
                ed, td, gd = T.update_position(env_d, T.eig_solver, b, 2, ortho, **kw)
                lo = env_o.psi[b].inds[2]
                ld = env_d.site_tensor(b).inds[2]
                same = (lo.qns, lo.dims) == (ld.qns, ld.dims)
                # two-site state error
                two_o = ob.contract(env_o.psi[b], env_o.psi[b + 1]).to_dense()
                A1 = env_d.site_tensor(b).to_host().to_dense(); A2 = env_d.site_tensor(b + 1).to_host().to_dense()
                two_d = np.tensordot(A1, A2, axes=([2], [0]))
                serr = min(np.abs(two_d - two_o).max(), np.abs(two_d + two_o).max()) if two_d.shape == two_o.shape else -1
                vo = dense_state(env_o.psi.t); vd = dense_state(env_d.getpsi())
                ov = abs(np.vdot(vo, vd)) / (np.linalg.norm(vo) * np.linalg.norm(vd))
                serr = 1 - ov
                flag = "" if (abs(eo - ed) < 1e-9 and same) else "   <<<<<< DIFF"
                print(f"sw{sweep} {ortho:5s} b{b}: E o={eo:.12f} d={ed:.12f} terr o={to:.3e} d={td:.3e} link o={lo.qns}{lo.dims} d={ld.qns}{ld.dims} 1-overlap {serr:.2e} eigs o={np.array2string(go[-3:],precision=3)} d={np.array2string(gd[-3:],precision=3)} {env_d.last_solver_info['numops']}{flag}")
                if flag:
                    bad += 1
                if bad > 3:
                    break
            if bad > 3:
                break
        if bad > 3:
            break

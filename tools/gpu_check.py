"""Verbose GPU-vs-oracle check used while bringing the CUDA path up (the pytest suite in tests/ is the
gate; this prints more)."""
import sys, time, traceback
import numpy as np
sys.path.insert(0, ".")
import tennetlib.jl_b200 as T
from oracle import blocksparse as ob, models as om, dmrg as od, projmpo as op, krylov as ok

ctx = T.Context()
rng = np.random.default_rng(7)
ok_all = True


def dense_of(t):
    return t.to_dense()


def rel(a, b):
    return float(np.abs(a - b).max() / max(1e-300, np.abs(b).max()))


def section(name):
    print(f"\n=== {name}", flush=True)


def run(name, fn):
    global ok_all
    section(name)
    try:
        fn()
    except Exception:
        ok_all = False
        traceback.print_exc()
        print("FAILED", name, flush=True)


def setup(kind="S=1", N=8, chi=24, seed=3):
    sites = om.siteinds(kind, N)
    H = om.heisenberg_mpo(sites)
    qns, dims = om.gaussian_link_sectors(chi, 1.3, 4, parity_offset=0, step=2 if kind == "S=1" else 1)
    psi = om.random_mps(sites, qns, dims, np.random.default_rng(seed))
    return sites, H, psi


def t_roundtrip():
    sites, H, psi = setup()
    for nrow in (1, 2, 3):
        A = psi[3]
        d = T.DeviceTensor.from_host(ctx, A, nrow=nrow)
        B = d.to_host()
        print("nrow", nrow, "roundtrip err", rel(B.to_dense(), A.to_dense()), "norm", d.norm(), A.norm())
        assert rel(B.to_dense(), A.to_dense()) == 0.0
        assert abs(d.norm() - A.norm()) < 1e-12 * A.norm()


def t_apply():
    for kind, N, chi in (("S=1", 8, 24), ("S=1/2", 10, 16), ("S=1", 6, 200)):
        sites, H, psi = setup(kind, N, chi)
        mps = od.MPS(psi)
        od.orthogonalize(mps, 1)
        env_o = od.StateEnvs(mps, H)
        env_d = T.StateEnvs(ctx, mps.t, H, llim=0, rlim=2)
        for pos in (1, N // 2, N - 1):
            env_o.set_nsite(2); env_o.position(pos)
            phi_o = ob.contract(env_o.psi[pos], env_o.psi[pos + 1])
            env_d.set_nsite(2)
            phi_d = env_d.make_phi(pos)
            e0 = rel(phi_d.to_host().to_dense(), phi_o.to_dense())
            env_d.position(pos)
            ob.reset_flops()
            Hv_o = env_o.product(phi_o)
            fl_o = ob.get_flops()
            Hv_d = env_d.product(phi_d)
            e1 = rel(Hv_d.to_host().to_dense(), Hv_o.to_dense())
            print(kind, N, chi, "pos", pos, "phi err", e0, "apply err", e1, "flops oracle", fl_o, "device", env_d.apply_flops())
            assert e0 < 1e-13 and e1 < 1e-12


def t_lanczos():
    sites, H, psi = setup("S=1", 8, 30)
    mps = od.MPS(psi); od.orthogonalize(mps, 1)
    env_o = od.StateEnvs(mps, H)
    env_d = T.StateEnvs(ctx, mps.t, H, llim=0, rlim=2)
    pos = 4
    env_o.set_nsite(2); env_o.position(pos)
    phi_o = ob.contract(env_o.psi[pos], env_o.psi[pos + 1])
    e_o, v_o, info = ok.eigsolve_lanczos(env_o, phi_o)
    env_d.set_nsite(2); phi_d = env_d.make_phi(pos); env_d.position(pos)
    e_d, v_d = T.eig_solver(env_d, phi_d)
    vd = v_d.to_host().to_dense(); vo = v_o.to_dense()
    sgn = np.sign(np.vdot(vd, vo))
    print("E oracle", e_o, "device", e_d, "diff", e_d - e_o, info, env_d.last_solver_info, "vec err", rel(sgn * vd, vo))
    assert abs(e_d - e_o) < 1e-11 * abs(e_o)
    assert rel(sgn * vd, vo) < 1e-8


def t_replacebond():
    for ortho in ("left", "right"):
        for kw in (dict(maxdim=12, cutoff=1e-14, noise=0.0), dict(maxdim=40, cutoff=1e-8, noise=0.0),
                   dict(maxdim=14, cutoff=1e-14, noise=1e-3)):
            sites, H, psi = setup("S=1", 8, 30)
            mps = od.MPS(psi); od.orthogonalize(mps, 1)
            # move centre to pos (left) or pos+1 (right) with the oracle, then hand the same state to both
            pos = 4
            od.orthogonalize(mps, pos if ortho == "left" else pos + 1)
            env_o = od.StateEnvs(mps, H)
            c = pos if ortho == "left" else pos + 1
            env_d = T.StateEnvs(ctx, mps.t, H, llim=c - 1, rlim=c + 1)
            env_o.set_nsite(2); env_o.position(pos)
            phi_o = ob.contract(env_o.psi[pos], env_o.psi[pos + 1])
            phi_o = phi_o.scale(1 / phi_o.norm())
            drho = None
            if kw["noise"]:
                d = op.drho_matrices(env_o.PH.noiseterm(phi_o, ortho), kw["noise"])
                drho = d if ortho == "left" else {tuple(-x for x in q): M for q, M in d.items()}
            spec = od.replacebond(env_o.psi, pos, phi_o, maxdim=kw["maxdim"], mindim=1, cutoff=kw["cutoff"],
                                  eigen_perturbation=drho, ortho=ortho, normalize=True)
            env_d.set_nsite(2); phi_d = env_d.make_phi(pos); env_d.position(pos)
            phi_d.scale_(1 / phi_d.norm())
            terr, eigs = env_d.replacebond(pos, phi_d, maxdim=kw["maxdim"], mindim=1, cutoff=kw["cutoff"],
                                           noise=kw["noise"], ortho=ortho, normalize=True)
            A1 = env_d.site_tensor(pos).to_host(); A2 = env_d.site_tensor(pos + 1).to_host()
            two_d = np.tensordot(A1.to_dense(), A2.to_dense(), axes=([2], [0]))
            two_o = ob.contract(env_o.psi[pos], env_o.psi[pos + 1]).to_dense()
            m_d, m_o = A1.inds[2], env_o.psi[pos].inds[2]
            print(ortho, kw, "truncerr", terr, spec.truncerr, "neigs", len(eigs), len(spec.eigs),
                  "eig err", rel(eigs, spec.eigs) if len(eigs) == len(spec.eigs) else None,
                  "link", m_d.qns == m_o.qns and m_d.dims == m_o.dims, "two-site err", rel(two_d, two_o))
            assert len(eigs) == len(spec.eigs) and (m_d.qns, m_d.dims) == (m_o.qns, m_o.dims)
            assert abs(terr - spec.truncerr) <= 1e-10 * spec.truncerr + 1e-14
            assert rel(two_d, two_o) < 1e-9


def t_dmrg():
    N = 12
    sites = om.siteinds("S=1/2", N)
    H = om.heisenberg_mpo(sites)
    psi0 = od.MPS(om.neel_mps(sites))
    kw = dict(nsweeps=[5], maxdim=[20], cutoff=1e-14, noise=1e-3, noisedecay=2, disable_noise_after=2)
    t = time.time()
    e_o, _, sw_o = od.dmrg2(psi0, H, od.DMRGParams(**kw))
    t_o = time.time() - t
    t = time.time()
    e_d, env, sw_d = T.dmrg2(ctx, psi0.t, H, T.DMRGParams(**kw), outputlevel=0)
    t_d = time.time() - t
    print("oracle", sw_o.energy, sw_o.maxchi, f"{t_o:.2f}s")
    print("device", sw_d.energy, sw_d.maxchi, f"{t_d:.2f}s")
    print("truncerr", sw_o.maxtruncerr, sw_d.maxtruncerr)
    print("ED -5.1420906328405 ; dE(device-oracle)", [a - b for a, b in zip(sw_d.energy, sw_o.energy)])
    assert sw_d.maxchi == sw_o.maxchi
    # sweeps with noise go through eigenvectors of nearly-null density-matrix directions: 1e-7; final: 1e-10
    assert max(abs(a - b) for a, b in zip(sw_d.energy, sw_o.energy)) < 1e-7 * abs(e_o)
    assert abs(sw_d.energy[-1] - sw_o.energy[-1]) < 1e-10 * abs(e_o)


run("roundtrip", t_roundtrip)
run("apply", t_apply)
run("lanczos", t_lanczos)
run("replacebond", t_replacebond)
run("dmrg", t_dmrg)
print("\ncounters", ctx.counters())
print("ALL OK" if ok_all else "SOME FAILED")
sys.exit(0 if ok_all else 1)

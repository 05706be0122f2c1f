"""Global Subspace Expansion on the device (`-m gpu`): `apply(H, psi)`, `krylov_extend!`
and the reference's own TDVP test flow (test/test_MPS_TDVP.jl:45-66) against the oracle (oracle/gse.py) and the exact
evolution."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _imports():
    import tennetlib.jl_b200 as T
    from tennetlib.jl_b200 import gse as dg, itensor as it
    from oracle import dmrg as od, gse as og, models as om
    return T, dg, it, od, og, om


def _dense(om, mps_tensors):
    """dense vector of a device MPS given as ITensors (l, s, r)"""
    from oracle import blocksparse as ob
    hs = [A.to_host() for A in mps_tensors]
    ts = [ob.BSTensor([ob.Index(ix.qns, ix.dims, dir=ix.dir, tags=ix.tags, plev=ix.plev, id=ix.id) for ix in h.inds],
                      {c: np.array(b) for c, b in h.blocks.items()},
                      np.complex128 if h.is_complex() else np.float64) for h in hs]
    return om.mps_to_dense(ts)


def test_apply_mpo_and_krylov_extend_match_oracle(ctx):
    T, dg, it, od, og, om = _imports()
    N = 8
    sites = om.siteinds("S=1/2", N)
    H = om.heisenberg_mpo(sites)
    Hd = om.mpo_to_dense(H)
    qn, dm = om.gaussian_link_sectors(12, 1.3, 4, step=1)
    mps = od.MPS(om.random_mps(sites, qn, dm, np.random.default_rng(2)))
    od.orthogonalize(mps, 1)
    v = om.mps_to_dense(mps.t)
    Hd_it = [it.ITensor.from_host(ctx, W, nrow=2) for W in H]
    psi_d = dg.DeviceMPS([it.ITensor.from_host(ctx, A) for A in mps.t], 1)
    w = _dense(om, dg.apply_mpo(Hd_it, psi_d, maxdim=None, cutoff=1e-15).t)
    assert np.abs(w - Hd @ v).max() < 1e-12 * np.abs(Hd @ v).max()
    # truncated: same bond dimensions and the same state as the oracle's apply
    phi_o = og.apply_mpo(H, mps, maxdim=8, cutoff=1e-15)
    phi_d = dg.apply_mpo(Hd_it, psi_d, maxdim=8, cutoff=1e-15)
    assert [A.inds[2].dim for A in phi_d.t[:-1]] == [A.inds[2].dim for A in phi_o.t[:-1]]
    w_o, w_d = om.mps_to_dense(phi_o.t), _dense(om, phi_d.t)
    assert abs(abs(np.vdot(w_o, w_d)) / np.linalg.norm(w_o) / np.linalg.norm(w_d) - 1.0) < 1e-10
    # krylov_extend!: state unchanged, bonds enlarged exactly as in the oracle, right-orthonormal, centre at site 1
    psi_o = od.MPS(om.neel_mps(sites))
    v0 = om.mps_to_dense(psi_o.t)
    env = T.StateEnvs(ctx, psi_o.t, H, llim=0, rlim=2)
    og.krylov_extend_mps(psi_o, H, extension_krylovdim=3)
    dg.krylov_extend(env, extension_krylovdim=3, outputlevel=0)
    assert env.linkdims() == [A.inds[2].dim for A in psi_o.t[:-1]] and max(env.linkdims()) > 1
    from helpers import to_oracle
    v1 = om.mps_to_dense([to_oracle(A) for A in env.getpsi()])
    assert abs(abs(np.vdot(v0, v1)) - 1.0) < 1e-12 and abs(np.linalg.norm(v1) - 1.0) < 1e-12
    assert env.isortho() and env.orthocenter() == 1
    for A in env.getpsi()[1:]:
        M = A.to_dense().reshape(A.inds[0].dim, -1)
        assert np.abs(M @ M.conj().T - np.eye(M.shape[0])).max() < 1e-12


@pytest.mark.parametrize("ts", [-0.02, -0.02j])
def test_dynamic_tdvp_on_a_single_mpo_matches_oracle_and_exact_evolution(ctx, ts):
    """tdvpsweep!(engine, dt; nsite = "dynamic", maxdim = 20, cutoff = 1E-12, extendat = 5) on an MPO: GSE + one-site
    sweep at sweeps 1, 5, 10, dynamic sweeps in between -- the flow of test/test_MPS_TDVP.jl:50-66."""
    import scipy.linalg as sl
    T, dg, it, od, og, om = _imports()
    from helpers import to_oracle
    N = 8
    sites = om.siteinds("S=1/2", N)
    H = om.heisenberg_mpo(sites)
    Hd = om.mpo_to_dense(H)
    psi0 = od.MPS(om.neel_mps(sites))
    v0 = om.mps_to_dense(psi0.t)
    eng_o = od.TDVPEngine(psi0, H)
    eng_d = T.TDVPEngine(ctx, psi0.t, H, llim=0, rlim=2)
    for _ in range(10):
        od.tdvpsweep(eng_o, ts, "dynamic", maxdim=20, cutoff=1e-12, extendat=5)
        T.tdvpsweep(eng_d, ts, "dynamic", maxdim=20, cutoff=1e-12, extendat=5, outputlevel=0)
    assert eng_d.swdata.maxchi == eng_o.swdata.maxchi
    assert np.allclose(eng_d.swdata.energy, eng_o.swdata.energy, rtol=1e-8, atol=0)
    v = sl.expm(10 * ts * Hd) @ v0
    v /= np.linalg.norm(v)
    w = om.mps_to_dense([to_oracle(A) for A in eng_d.getpsi()])
    assert abs(abs(np.vdot(v, w)) - 1.0) < 1e-9
    assert abs(eng_d.swdata.energy[-1] - np.real(np.vdot(v, Hd @ v))) < 1e-5

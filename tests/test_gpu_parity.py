"""CUDA path vs CPU oracle through the C ABI (ctypes) -- the parity gate.  Run on the B200 with `-m gpu`.

Tolerances: block structure / bond dimensions / work counts exact; FP64 data 1e-12 relative (different but
equivalent summation orders); energies 1e-10 relative as required by BASELINE.json north_star."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "oracle_golden.json")))


def _imports():
    import tennetlib.jl_b200 as T
    from oracle import blocksparse as ob, dmrg as od, krylov as ok, models as om, projmpo as op
    return T, ob, od, ok, om, op


def rel(a, b):
    return float(np.abs(a - b).max() / max(1e-300, np.abs(b).max()))


def _setup(om, od, kind, N, chi, seed, center):
    sites = om.siteinds(kind, N)
    H = om.heisenberg_mpo(sites)
    qn, dm = om.gaussian_link_sectors(chi, 1.3, 4, step=2 if kind == "S=1" else 1)
    mps = od.MPS(om.random_mps(sites, qn, dm, np.random.default_rng(seed)))
    od.orthogonalize(mps, center)
    return sites, H, mps


@pytest.mark.parametrize("nrow", [1, 2, 3])
def test_import_export_roundtrip_is_bit_exact(ctx, nrow):
    T, ob, od, ok, om, op = _imports()
    _, _, mps = _setup(om, od, "S=1", 8, 24, 3, 1)
    A = mps[4]
    d = T.DeviceTensor.from_host(ctx, A, nrow=nrow)
    B = d.to_host()
    assert np.array_equal(B.to_dense(), A.to_dense())
    assert [(ix.qns, ix.dims, ix.dir) for ix in B.inds] == [(ix.qns, ix.dims, ix.dir) for ix in A.inds]
    assert abs(d.norm() - A.norm()) < 1e-13 * A.norm()


def test_empty_and_single_block_tensors(ctx):
    T, ob, od, ok, om, op = _imports()
    sites = om.siteinds("S=1/2", 4)
    psi = om.neel_mps(sites)                       # product state: one 1x1x1 block per site, most blocks absent
    for A in psi:
        d = T.DeviceTensor.from_host(ctx, A, nrow=2)
        assert np.array_equal(d.to_host().to_dense(), A.to_dense())
    z = T.DeviceTensor.zeros(ctx, psi[1].inds, nrow=1)
    assert z.norm() == 0.0


def test_vector_interface(ctx):
    T, ob, od, ok, om, op = _imports()
    _, _, mps = _setup(om, od, "S=1", 8, 40, 11, 1)
    A, B = mps[4], mps[4].scale(0.5).add(mps[4], 0.25)
    x, y = T.DeviceTensor.from_host(ctx, A, 1), T.DeviceTensor.from_host(ctx, B, 1)
    assert abs(x.dot(y) - ob.inner(A, B)) < 1e-13 * abs(ob.inner(A, B))
    y.axpy_(x, -0.3).scale_(2.0)
    ref = B.add(A, -0.3).scale(2.0)
    assert rel(y.to_host().to_dense(), ref.to_dense()) < 1e-15


@pytest.mark.parametrize("kind,N,chi", [("S=1", 8, 24), ("S=1/2", 10, 16), ("S=1", 6, 200)])
def test_heff_apply_and_environments_match_oracle(ctx, kind, N, chi):
    T, ob, od, ok, om, op = _imports()
    sites, H, mps = _setup(om, od, kind, N, chi, 3, 1)
    env_o = od.StateEnvs(mps, H)
    env_d = T.StateEnvs(ctx, mps.t, H, llim=0, rlim=2)
    for pos in (1, N // 2, N - 1):
        env_o.set_nsite(2); env_o.position(pos)
        phi_o = ob.contract(env_o.psi[pos], env_o.psi[pos + 1])
        env_d.set_nsite(2)
        phi_d = env_d.make_phi(pos)
        assert rel(phi_d.to_host().to_dense(), phi_o.to_dense()) < 1e-13
        env_d.position(pos)
        Hv_o = env_o.product(phi_o)
        Hv_d = env_d.product(phi_d)
        assert rel(Hv_d.to_host().to_dense(), Hv_o.to_dense()) < 1e-12
        assert abs(env_d.expectation(phi_d) - ob.inner(phi_o, Hv_o)) < 1e-12 * abs(ob.inner(phi_o, Hv_o))
        # linearity and symmetry of the projected Hamiltonian
        w = phi_d.copy().fill_random(99)
        a = w.dot(env_d.product(phi_d)); b = env_d.product(w).dot(phi_d)
        assert abs(a - b) < 1e-10 * max(1.0, abs(a))


def test_one_site_and_zero_site_apply_match_oracle(ctx):
    T, ob, od, ok, om, op = _imports()
    sites, H, mps = _setup(om, od, "S=1", 8, 24, 3, 4)
    env_o = od.StateEnvs(mps, H)
    env_d = T.StateEnvs(ctx, mps.t, H, llim=3, rlim=5)
    env_o.set_nsite(1); env_o.position(4)
    env_d.set_nsite(1); env_d.position(4)
    v_o = env_o.psi[4]
    v_d = env_d.site_tensor(4)
    assert rel(env_d.product(v_d).to_host().to_dense(), env_o.product(v_o).to_dense()) < 1e-12
    # zero-site: bond matrix between sites 4 and 5
    L, R, spec, u = ob.factorize(v_o, v_o.inds[:2], ortho="left", which_decomp="svd", cutoff=0.0)
    env_o.psi[4] = L                                  # psi[pos] = U ; C = S*V lives on the bond (update_site.jl:170-186)
    env_d2 = T.StateEnvs(ctx, env_o.psi.t, H, llim=4, rlim=6)
    env_o.PH.lpos, env_o.PH.rpos = 0, 9
    env_o.set_nsite(0); env_o.position(5)
    env_d2.set_nsite(0); env_d2.position(5)
    C = R                                             # (u, r) bond matrix
    Cd = T.DeviceTensor.from_host(ctx, C, 1)
    ref = env_o.product(C)
    assert rel(env_d2.product(Cd).to_host().to_dense(), ref.permute(C.inds).to_dense()) < 1e-12


@pytest.mark.parametrize("k", [0, 1])
def test_bond_fixture_lanczos_and_truncation(ctx, k):
    """Golden fixture + oracle: eig_solver energy, applies count, truncation error, kept spectrum, new link sectors."""
    T, ob, od, ok, om, op = _imports()
    g = G["bond"][k]
    sites, H, mps = _setup(om, od, g["kind"], g["N"], g["chi"], g["seed"], g["pos"])
    pos = g["pos"]
    for t in g["trunc"]:
        c = pos if t["ortho"] == "left" else pos       # centre already at pos: two-site tensor is the same
        env_d = T.StateEnvs(ctx, mps.t, H, llim=pos - 1, rlim=pos + 1)
        env_d.set_nsite(2)
        phi = env_d.make_phi(pos)
        env_d.position(pos)
        phi.scale_(1 / phi.norm())
        assert abs(env_d.expectation(phi) - g["expectation"]) < 1e-11 * abs(g["expectation"])
        assert abs(env_d.product(phi).norm() - g["Hv_norm"]) < 1e-11 * g["Hv_norm"]
        e, phi = T.eig_solver(env_d, phi)
        assert abs(e - g["lanczos_energy"]) < 1e-10 * abs(g["lanczos_energy"])
        assert env_d.last_solver_info["numops"] == g["lanczos_numops"]
        phi.scale_(1 / phi.norm())
        # the fixture truncates with ortho=left bookkeeping for both directions of the same centre
        env_d.llim, env_d.rlim = (pos - 1, pos + 1) if t["ortho"] == "left" else (pos, pos + 2)
        terr, eigs = env_d.replacebond(pos, phi, maxdim=t["maxdim"], mindim=1, cutoff=t["cutoff"], noise=0.0,
                                       ortho=t["ortho"], normalize=True)
        link = env_d.site_tensor(pos).inds[2]
        assert [list(q) for q in link.qns] == t["link_qns"] and list(link.dims) == t["link_dims"]
        assert len(eigs) == len(t["eigs"])
        assert abs(terr - t["truncerr"]) < 1e-10 * t["truncerr"] + 1e-14
        assert np.abs(eigs - np.array(t["eigs"])).max() < 1e-12


@pytest.mark.parametrize("ortho", ["left", "right"])
@pytest.mark.parametrize("kw", [dict(maxdim=12, cutoff=1e-14, noise=0.0), dict(maxdim=40, cutoff=1e-8, noise=0.0),
                                dict(maxdim=14, cutoff=1e-14, noise=1e-3)])
def test_replacebond_matches_oracle(ctx, ortho, kw):
    T, ob, od, ok, om, op = _imports()
    pos = 4
    c = pos if ortho == "left" else pos + 1
    sites, H, mps = _setup(om, od, "S=1", 8, 30, 3, c)
    env_o = od.StateEnvs(mps, H)
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
    terr, eigs = env_d.replacebond(pos, phi_d, maxdim=kw["maxdim"], mindim=1, cutoff=kw["cutoff"], noise=kw["noise"],
                                   ortho=ortho, normalize=True)
    A1 = env_d.site_tensor(pos).to_host(); A2 = env_d.site_tensor(pos + 1).to_host()
    two_d = np.tensordot(A1.to_dense(), A2.to_dense(), axes=([2], [0]))
    two_o = ob.contract(env_o.psi[pos], env_o.psi[pos + 1]).to_dense()
    m_d, m_o = A1.inds[2], env_o.psi[pos].inds[2]
    assert (m_d.qns, m_d.dims) == (m_o.qns, m_o.dims)
    assert len(eigs) == len(spec.eigs) and np.abs(eigs - spec.eigs).max() < 1e-12
    assert abs(terr - spec.truncerr) <= 1e-10 * spec.truncerr + 1e-14
    assert rel(two_d, two_o) < 1e-9
    # the isometric factor is an isometry
    iso = A1.to_dense() if ortho == "left" else A2.to_dense()
    M = iso.reshape(-1, iso.shape[2]) if ortho == "left" else iso.reshape(iso.shape[0], -1).T
    assert np.abs(M.T @ M - np.eye(M.shape[1])).max() < 1e-10


def test_qr_gauge_moves_preserve_the_state(ctx):
    T, ob, od, ok, om, op = _imports()
    sites, H, mps = _setup(om, od, "S=1/2", 10, 16, 5, 1)
    raw = od.MPS(om.random_mps(sites, *om.gaussian_link_sectors(16, 1.3, 4, step=1), np.random.default_rng(5)))
    env = T.StateEnvs(ctx, raw.t, H, llim=0, rlim=11)
    v0 = om.mps_to_dense(raw.t)
    env.orthogonalize(1)
    env.orthogonalize(6)
    psi = env.getpsi()
    v1 = om.mps_to_dense(psi)
    assert rel(v1, v0) < 1e-12
    for j in range(0, 5):                 # left-canonical
        M = psi[j].to_dense(); M = M.reshape(-1, M.shape[2])
        assert np.abs(M.T @ M - np.eye(M.shape[1])).max() < 1e-12
    for j in range(6, 10):                # right-canonical
        M = psi[j].to_dense(); M = M.reshape(M.shape[0], -1)
        assert np.abs(M @ M.T - np.eye(M.shape[0])).max() < 1e-12


@pytest.mark.parametrize("k", [0, 1, 2])
def test_dmrg_matches_oracle_fixture_and_ed(ctx, k):
    """configs[0]-scale end-to-end parity: per-sweep energies, bond dimensions and truncation errors."""
    T, ob, od, ok, om, op = _imports()
    g = G["dmrg"][k]
    sites = om.siteinds(g["kind"], g["N"])
    H = om.heisenberg_mpo(sites)
    psi0 = om.neel_mps(sites)
    e, env, sw = T.dmrg2(ctx, psi0, H, T.DMRGParams(**g["params"]), outputlevel=0)
    assert sw.maxchi == g["maxchi"]
    assert env.linkdims() == g["linkdims"]
    noisy = np.array([n > 0 for n in _noise_per_sweep(g["params"])])
    de = np.abs(np.array(sw.energy) - np.array(g["energy"]))
    assert de[~noisy].max() < 1e-10 * abs(g["energy"][-1])
    assert de.max() < 1e-7 * abs(g["energy"][-1])
    te = np.abs(np.array(sw.maxtruncerr) - np.array(g["maxtruncerr"]))
    assert te[~noisy].max() < 1e-12 and te.max() < 1e-8
    ed = G["ed"]["S12_N12"] if g["kind"] == "S=1/2" else G["ed"]["S1_N8"]
    assert e > ed - 1e-11 and e - ed < 1e-8


@pytest.mark.parametrize("k", [0, 1])
def test_one_site_dmrg_matches_oracle_fixture(ctx, k):
    """`dmrg1` (src/mps/dmrg.jl:311-320 -> _update_one_site!): noise branch (two-site replacebond!) and svd split."""
    T, ob, od, ok, om, op = _imports()
    g = G["dmrg1"][k]
    sites = om.siteinds(g["kind"], g["N"])
    H = om.heisenberg_mpo(sites)
    e, env, sw = T.dmrg1(ctx, om.neel_mps(sites), H, T.DMRGParams(**g["params"]), outputlevel=0)
    assert sw.maxchi == g["maxchi"] and env.linkdims() == g["linkdims"]
    noisy = np.array([n > 0 for n in _noise_per_sweep(g["params"])])
    de = np.abs(np.array(sw.energy) - np.array(g["energy"]))
    assert de[~noisy].max() < 1e-10 * abs(g["energy"][-1]) and de.max() < 1e-7 * abs(g["energy"][-1])


def test_excited_state_dmrg_with_penalty(ctx):
    """StateEnvs(psi0, H, [psi_gr]; weight=10) (test/test_MPS_DMRG.jl:68-97): ProjMPO_MPS2 on the device."""
    T, ob, od, ok, om, op = _imports()
    g = G["excited"][0]
    sites = om.siteinds(g["kind"], g["N"])
    H = om.heisenberg_mpo(sites)
    psi0 = om.neel_mps(sites)
    e0, env0, sw0 = T.dmrg2(ctx, psi0, H, T.DMRGParams(**g["params"]), outputlevel=0)
    gs = env0.getpsi()
    e1, env1, sw1 = T.dmrg2(ctx, psi0, H, T.DMRGParams(**g["params"]), Ms=[gs], weight=g["weight"], outputlevel=0)
    assert abs(e0 - g["e0"]) < 1e-10 * abs(e0)
    assert sw1.maxchi == g["maxchi"]
    assert abs(e1 - g["energy"][-1]) < 1e-8                         # oracle fixture (its own ground state as penalty)
    assert e1 > G["ed"]["S12_N12_E1"] - 1e-9 and e1 - G["ed"]["S12_N12_E1"] < 1e-6      # exact first excited level
    with pytest.raises(ValueError):
        T.StateEnvs(ctx, psi0, H, Ms=[gs], weight=-1.0)


def _noise_per_sweep(p):
    out = []
    n = len(p["nsweeps"])
    vec = lambda x: x if isinstance(x, list) else [x] * n
    for ii in range(n):
        noise = vec(p.get("noise", 0.0))[ii]
        decay = vec(p.get("noisedecay", 1.0))[ii]
        off = vec(p.get("disable_noise_after", -1))[ii]
        for jj in range(1, p["nsweeps"][ii] + 1):
            out.append(noise)
            if jj == off:
                noise = 0.0
            noise /= decay
            if noise < 1e-13:
                noise = 0.0
    # a sweep right after noisy ones still starts from the noisy state: count it as noisy as well
    return [a if a > 0 else (out[i - 1] if i > 0 else 0.0) for i, a in enumerate(out)]


def test_large_block_properties_chi512(ctx):
    """Size-independent properties at a size where the 128x128 DMMA tiles, K tails and odd leading dimensions
    are all exercised: symmetry and linearity of H_eff, Lanczos lowers the energy, truncation bookkeeping."""
    import tennetlib.jl_b200 as T
    from tennetlib.jl_b200 import models as pm
    N, chi = 12, 512
    sites = pm.siteinds("S=1", N)
    H = pm.heisenberg_mpo(sites)
    qn, dm = pm.gaussian_link_sectors(chi, 1.3, 5)
    links = pm.random_mps_links(sites, qn, dm)
    psi = [T.DeviceTensor.zeros(ctx, [links[j].copy(dir=+1), sites[j].copy(dir=+1), links[j + 1].copy(dir=-1)], 2).fill_random(7 + j)
           for j in range(N)]
    env = T.StateEnvs(ctx, psi, H, llim=0, rlim=N + 1, copy=False)
    env.orthogonalize(1)
    env.orthogonalize(6)
    env.set_nsite(2)
    phi = env.make_phi(6)
    env.position(6)
    phi.scale_(1 / phi.norm())
    x = phi.copy().fill_random(1); y = phi.copy().fill_random(2)
    Hx, Hy = env.product(x), env.product(y)
    assert abs(y.dot(Hx) - Hy.dot(x)) < 1e-11 * abs(y.dot(Hx))
    z = x.copy().axpy_(y, 0.7)
    Hz = env.product(z)
    Hz.axpy_(Hx, -1.0).axpy_(Hy, -0.7)
    assert Hz.norm() < 1e-12 * Hx.norm()
    e0 = env.expectation(phi)
    e1, phi = T.eig_solver(env, phi)
    assert e1 < e0 and abs(phi.norm() - 1) < 1e-12
    assert abs(env.expectation(phi) - e1) < 1e-9 * abs(e1)
    terr, eigs = env.replacebond(6, phi, maxdim=300, mindim=1, cutoff=0.0, noise=0.0, ortho="left", normalize=True)
    assert len(eigs) == 300 and abs(eigs.sum() + terr - 1.0) < 1e-12 and np.all(np.diff(eigs) <= 0)
    assert sum(env.site_tensor(6).inds[2].dims) <= 300


@pytest.mark.parametrize("svd_alg", ["qr_iteration", "polar", "gram"])
def test_svd_drivers_agree(ctx, svd_alg):
    """All SVD drivers of the truncation give the same spectrum / truncation error / bond dimension; the Gram
    driver trips its accuracy guard (and falls back) when the kept spectrum spans more than 1e-10."""
    T, ob, od, ok, om, op = _imports()
    pos = 4
    sites, H, mps = _setup(om, od, "S=1", 8, 100, 3, pos)
    env_o = od.StateEnvs(mps, H)
    env_o.set_nsite(2); env_o.position(pos)
    phi_o = ob.contract(env_o.psi[pos], env_o.psi[pos + 1])
    phi_o = phi_o.scale(1 / phi_o.norm())
    for maxdim, cutoff in ((60, 1e-14), (10000, 0.0)):
        mo = env_o.psi.copy()
        spec = od.replacebond(mo, pos, phi_o, maxdim=maxdim, mindim=1, cutoff=cutoff, eigen_perturbation=None,
                              ortho="left", normalize=True)
        env_d = T.StateEnvs(ctx, mps.t, H, llim=pos - 1, rlim=pos + 1)
        env_d.set_nsite(2); phi_d = env_d.make_phi(pos); env_d.position(pos)
        phi_d.scale_(1 / phi_d.norm())
        terr, eigs = env_d.replacebond(pos, phi_d, maxdim=maxdim, mindim=1, cutoff=cutoff, noise=0.0, ortho="left",
                                       normalize=True, svd_alg=svd_alg)
        assert len(eigs) == len(spec.eigs)
        assert np.abs(eigs - spec.eigs).max() < 1e-12 and abs(terr - spec.truncerr) < 1e-12
        A1 = env_d.site_tensor(pos).to_host().to_dense(); A2 = env_d.site_tensor(pos + 1).to_host().to_dense()
        two_d = np.tensordot(A1, A2, axes=([2], [0]))
        two_o = ob.contract(mo[pos], mo[pos + 1]).to_dense()
        assert rel(two_d, two_o) < 1e-9
        M = A1.reshape(-1, A1.shape[2])
        assert np.abs(M.T @ M - np.eye(M.shape[1])).max() < 1e-10


def test_grouped_dgemm_kernel_against_naive_reference(ctx):
    """Kernel-level check: DMMA grouped GEMM vs a one-thread-per-element FP64 reference kernel, ragged sizes,
    all four operand layouts (tails in M, N and K, odd leading dimensions)."""
    for (M, N, K) in ((1, 1, 1), (7, 5, 3), (129, 65, 17), (255, 257, 33), (385, 254, 935)):
        for ta in (0, 1):
            for tb in (0, 1):
                ms, err = ctx.gemm_selftest(M, N, K, ta, tb, 1, True)
                assert 0.0 <= err < 1e-12 * K, (M, N, K, ta, tb, err)


def test_sharded_apply_two_gpus_matches_oracle():
    """world_size-2 NCCL run of the sharded apply (skipped on single-GPU boxes; tools/multi_gpu_check.py)."""
    import subprocess, sys, torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29517", "tools/multi_gpu_check.py"],
                       cwd=root, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "MULTI-GPU CHECK PASSED" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


# ---------------------------------------------------------------------------------------------- TDVP / exp_solver
@pytest.mark.parametrize("nsite,t", [(2, -0.1), (1, -0.05), (0, 0.07), (2, 0.6)])
def test_exp_solver_matches_oracle(ctx, nsite, t):
    """tnl_exponentiate vs the oracle restatement of KrylovKit.exponentiate on the same H_eff and vector."""
    T, ob, od, ok, om, op = _imports()
    sites, H, mps = _setup(om, od, "S=1", 8, 24, 3, 4)
    env_o = od.StateEnvs(mps, H)
    pos = 4
    if nsite == 0:
        v_o = env_o.psi[4]
        L, R, spec, u = ob.factorize(v_o, v_o.inds[:2], ortho="left", which_decomp="svd", cutoff=0.0)
        env_o.psi[4] = L
        phi_o, pos = R, 5
        env_d = T.StateEnvs(ctx, env_o.psi.t, H, llim=4, rlim=6)
    else:
        env_d = T.StateEnvs(ctx, mps.t, H, llim=3, rlim=5)
        phi_o = ob.contract(env_o.psi[4], env_o.psi[5]) if nsite == 2 else env_o.psi[4]
    env_o.set_nsite(nsite); env_o.position(pos)
    env_d.set_nsite(nsite); env_d.position(pos)
    phi_d = T.DeviceTensor.from_host(ctx, phi_o, 1)
    e_o, out_o = od.exp_solver(env_o, phi_o, t)
    e_d, out_d = T.exp_solver(env_d, phi_d, t)
    assert np.isnan(e_o) and np.isnan(e_d)
    info_o, info_d = od.exp_solver.last_info, env_d.last_solver_info
    assert info_d["converged"] == 1 and info_o["converged"] == 1
    ref = out_o.permute(phi_o.inds).to_dense()
    assert rel(out_d.to_host().to_dense(), ref) < 1e-10
    # the eager exit test sits at the tolerance; for growing exponentials (t > 0) the error estimate
    # |dt beta normres expH[K, K+2]| is itself at the rounding level of the small exponential
    assert abs(info_d["numops"] - info_o["numops"]) <= (1 if t < 0 else 3)
    # imaginary-time evolution is not norm preserving; the norm must match too
    assert abs(out_d.norm() - out_o.norm()) < 1e-11 * out_o.norm()


def test_exp_solver_substeps_when_the_basis_is_full(ctx):
    """krylovdim too small for one step: adaptive sub-stepping + restarts (expintegrator outer loop)."""
    T, ob, od, ok, om, op = _imports()
    sites, H, mps = _setup(om, od, "S=1", 8, 24, 5, 4)
    env_o = od.StateEnvs(mps, H)
    env_d = T.StateEnvs(ctx, mps.t, H, llim=3, rlim=5)
    env_o.set_nsite(2); env_o.position(4)
    env_d.set_nsite(2); env_d.position(4)
    phi_o = ob.contract(env_o.psi[4], env_o.psi[5])
    phi_d = env_d.make_phi(4)
    kw = dict(solver_krylovdim=8, solver_maxiter=100, solver_tol=1e-11)
    _, out_o = od.exp_solver(env_o, phi_o, -0.2, **kw)
    _, out_d = T.exp_solver(env_d, phi_d, -0.2, **kw)
    assert od.exp_solver.last_info["numiter"] > 1 and env_d.last_solver_info["numiter"] > 1
    assert od.exp_solver.last_info["converged"] == 1 and env_d.last_solver_info["converged"] == 1
    assert abs(env_d.last_solver_info["numiter"] - od.exp_solver.last_info["numiter"]) <= 2
    assert rel(out_d.to_host().to_dense(), out_o.permute(phi_o.inds).to_dense()) < 1e-9


def test_exp_solver_missing_step_and_promotion(ctx):
    T, ob, od, ok, om, op = _imports()
    sites, H, mps = _setup(om, od, "S=1", 6, 12, 5, 3)
    env_d = T.StateEnvs(ctx, mps.t, H, llim=2, rlim=4)
    env_d.set_nsite(2); env_d.position(3)
    phi = env_d.make_phi(3)
    with pytest.raises(RuntimeError):
        T.exp_solver(env_d, phi, None)
    assert not phi.is_complex()
    n0 = phi.norm()
    _, out = T.exp_solver(env_d, phi, -0.05j)            # real-time step: the vector is promoted to ComplexF64
    assert out.is_complex() and abs(out.norm() - n0) < 1e-11 * n0


@pytest.mark.parametrize("nsite", [2, 1])
def test_tdvp_imaginary_time_sweeps_match_oracle(ctx, nsite):
    """tdvpsweep! (src/mps/tdvp.jl:247-277) with a real step: energies per sweep, bond dimensions, truncation
    errors and the final state against the oracle; the energy decreases monotonically towards E0."""
    T, ob, od, ok, om, op = _imports()
    N = 8
    sites = om.siteinds("S=1/2", N)
    H = om.heisenberg_mpo(sites)
    psi0 = od.MPS(om.neel_mps(sites))
    eng_o = od.TDVPEngine(psi0, H)
    eng_d = T.TDVPEngine(ctx, psi0.t, H)
    sched = [(2, -0.1)] * 2 + [(nsite, -0.1)] * 3            # bond dimensions have to grow with two-site sweeps first
    for ns, dt in sched:
        od.tdvpsweep(eng_o, dt, ns, maxdim=16, cutoff=1e-12)
        T.tdvpsweep(eng_d, dt, ns, maxdim=16, cutoff=1e-12, outputlevel=0)
    assert eng_d.swdata.maxchi == eng_o.swdata.maxchi
    assert np.allclose(eng_d.swdata.energy, eng_o.swdata.energy, rtol=1e-10, atol=0)
    assert np.allclose(eng_d.swdata.maxtruncerr, eng_o.swdata.maxtruncerr, rtol=0, atol=1e-12)
    assert np.allclose(eng_d.swdata.entropy, eng_o.swdata.entropy, rtol=0, atol=1e-9)
    assert all(b <= a + 1e-12 for a, b in zip(eng_d.swdata.energy, eng_d.swdata.energy[1:]))
    assert abs(eng_d.abstime - 0.5) < 1e-14
    vo = om.mps_to_dense(eng_o.sysenv.psi.t)
    vd = om.mps_to_dense(eng_d.getpsi())
    assert abs(abs(np.vdot(vo, vd)) - 1.0) < 1e-10


def test_tdvp_two_site_follows_exact_imaginary_time_evolution(ctx):
    """Size-independent property: exp(-tau H)|Neel> from dense linear algebra vs TDVP on the device."""
    import scipy.linalg as sl
    T, ob, od, ok, om, op = _imports()
    N = 8
    sites = om.siteinds("S=1/2", N)
    H = om.heisenberg_mpo(sites)
    psi0 = om.neel_mps(sites)
    eng = T.TDVPEngine(ctx, psi0, H)
    for _ in range(5):
        T.tdvpsweep(eng, -0.02, 2, maxdim=16, cutoff=1e-14, outputlevel=0)
    Hd = om.mpo_to_dense(H)
    v = sl.expm(-0.1 * Hd) @ om.mps_to_dense(psi0)
    v /= np.linalg.norm(v)
    w = om.mps_to_dense(eng.getpsi())
    assert abs(abs(np.vdot(v, w)) - 1.0) < 1e-8          # O(dt^3) per step splitting error
    assert abs(T.getenergy(eng) - float(v @ Hd @ v)) < 1e-6


# ---------------------------------------------------------------------------------------------- ProjMPOSum2
def test_mpo_sum_apply_noise_and_dmrg_match_oracle(ctx):
    """StateEnvs(psi, Hs::Vector{MPO}) (src/mps/projmposum2.jl): apply, eig_solver, noisy + noise-free replacebond!
    and a full DMRG run against the oracle, with two MPOs of different bond dimension."""
    T, ob, od, ok, om, op = _imports()
    N = 8
    sites = om.siteinds("S=1", N)
    Hs = [om.heisenberg_mpo(sites, Jz=1.0, Jxy=0.0), om.heisenberg_mpo(sites, Jz=0.0, Jxy=1.0)]
    H = om.heisenberg_mpo(sites)
    qn, dm = om.gaussian_link_sectors(30, 1.3, 4)
    mps = od.MPS(om.random_mps(sites, qn, dm, np.random.default_rng(7)))
    od.orthogonalize(mps, 4)
    env_o = od.StateEnvs(mps, Hs)
    env_d = T.StateEnvs(ctx, mps.t, Hs, llim=3, rlim=5)
    env_1 = T.StateEnvs(ctx, mps.t, H, llim=3, rlim=5)
    assert env_d.nterms == 2
    env_o.set_nsite(2); env_o.position(4)
    phi_o = ob.contract(env_o.psi[4], env_o.psi[5])
    for e in (env_d, env_1):
        e.set_nsite(2)
    phi_d = env_d.make_phi(4)
    env_d.position(4); env_1.position(4)
    ref = env_o.product(phi_o).permute(phi_o.inds).to_dense()
    assert rel(env_d.product(phi_d).to_host().to_dense(), ref) < 1e-12
    assert rel(env_1.product(env_1.make_phi(4)).to_host().to_dense(), ref) < 1e-12      # sum == single MPO
    for noise in (1e-3, 0.0):
        eo, to, so = od.update_position(env_o, od.eig_solver, 4, 2, "left", maxdim=24, cutoff=1e-13, noise=noise)
        ed, td, sd = T.update_position(env_d, T.eig_solver, 4, 2, "left", maxdim=24, cutoff=1e-13, noise=noise)
        assert abs(ed - eo) < 1e-10 * abs(eo)
        assert len(sd) == len(so) and abs(td - to) < 1e-10
        assert np.allclose(sd, so, rtol=0, atol=1e-9 if noise else 1e-12)
        od.orthogonalize(env_o.psi, 4); env_d.orthogonalize(4)
        # keep the two states identical for the next round (noisy eigenvectors differ in null directions)
        env_d = T.StateEnvs(ctx, env_o.psi.t, Hs, llim=3, rlim=5)
    # full DMRG: two-site with noise, then one-site
    psi0 = od.MPS(om.neel_mps(sites))
    prm = dict(maxdim=[10, 20], nsweeps=[2, 2], cutoff=1e-13, noise=[1e-4, 0.0])
    Eo, _, swo = od.dmrg2(psi0, Hs, od.DMRGParams(**prm))
    Ed, _, swd = T.dmrg2(ctx, psi0.t, Hs, T.DMRGParams(**prm), outputlevel=0)
    assert swd.maxchi == swo.maxchi
    assert abs(Ed - Eo) < 1e-9 * abs(Eo)
    assert abs(swd.energy[-1] - swo.energy[-1]) < 1e-10 * abs(Eo)


# ---------------------------------------------------------------------------------------------- ProjCouplingModel
def _cm_setup(om, od, oc, kind, N, chi, seed, center, **model_kw):
    sites = om.siteinds(kind, N)
    M = oc.heisenberg_coupling_model(sites, **model_kw)
    qn, dm = om.gaussian_link_sectors(chi, 1.3, 4, step=2 if kind == "S=1" else 1)
    mps = od.MPS(om.random_mps(sites, qn, dm, np.random.default_rng(seed)))
    od.orthogonalize(mps, center)
    return sites, M, mps


@pytest.mark.parametrize("model_kw", [dict(merge=True), dict(merge=False), dict(merge=True, field=0.3, j2=0.5)])
def test_coupling_model_apply_matches_oracle(ctx, model_kw):
    """StateEnvs(psi, H::CouplingModel): environments and product(v) id by id (src/mps/projcouplingmodel.jl:123-383)
    for two-, one- and zero-site vectors, at the edges and mid-chain; terms that skip sites (j2) and one-site terms
    (field) included.  The sum over ids must also equal the MPO of the same Hamiltonian."""
    T, ob, od, ok, om, op = _imports()
    from oracle import couplingmodel as oc
    N = 8
    sites, M, mps = _cm_setup(om, od, oc, "S=1", N, 24, 3, 1, **model_kw)
    env_o = od.StateEnvs(mps, M)
    env_d = T.StateEnvs(ctx, mps.t, M, llim=0, rlim=2)
    assert env_d.is_coupling_model
    for pos in (1, 4, N - 1):
        env_o.set_nsite(2); env_o.position(pos)
        env_d.set_nsite(2)
        phi_o = ob.contract(env_o.psi[pos], env_o.psi[pos + 1])
        phi_d = env_d.make_phi(pos)
        env_d.position(pos)
        ref = env_o.product(phi_o).permute(phi_o.inds).to_dense()
        assert rel(env_d.product(phi_d).to_host().to_dense(), ref) < 1e-12
        w = phi_d.copy().fill_random(7)
        a = w.dot(env_d.product(phi_d)); b = env_d.product(w).dot(phi_d)
        assert abs(a - b) < 1e-10 * max(1.0, abs(a))                     # hermiticity of the summed H_eff
    if model_kw == dict(merge=True):
        # same Hamiltonian as an MPO: equal H_eff once the orthogonality centre sits inside the site range (terms
        # without a tensor on one side act as the identity there, which presumes orthonormal block bases)
        mps4 = mps.copy()
        od.orthogonalize(mps4, 4)
        env_m = T.StateEnvs(ctx, mps4.t, om.heisenberg_mpo(sites), llim=3, rlim=5)
        env_c = T.StateEnvs(ctx, mps4.t, M, llim=3, rlim=5)
        env_m.set_nsite(2); env_c.set_nsite(2)
        phi = env_m.make_phi(4)
        env_m.position(4); env_c.position(4)
        assert rel(env_c.product(phi).to_host().to_dense(), env_m.product(phi).to_host().to_dense()) < 1e-12
    # one-site and zero-site
    env_o.set_nsite(1); env_o.position(4)
    env_d.set_nsite(1); env_d.position(4)
    v_o = env_o.psi[4]
    ref = env_o.product(v_o).permute(v_o.inds).to_dense()
    assert rel(env_d.product(env_d.site_tensor(4)).to_host().to_dense(), ref) < 1e-12
    # zero-site: bond matrix between sites 4 and 5 (psi[4] = U, C = S*V as in update_site.jl:162-186)
    v4 = env_o.psi[4]
    Lf, Rf, spec, u = ob.factorize(v4, v4.inds[:2], ortho="left", which_decomp="svd", cutoff=0.0)
    env_o.psi[4] = Lf
    env_o.PH.lpos, env_o.PH.rpos = 0, N + 1
    env_d2 = T.StateEnvs(ctx, env_o.psi.t, M, llim=4, rlim=6)
    env_o.set_nsite(0); env_o.position(5)
    env_d2.set_nsite(0); env_d2.position(5)
    Cd = T.DeviceTensor.from_host(ctx, Rf, 1)
    ref = env_o.product(Rf).permute(Rf.inds).to_dense()
    assert rel(env_d2.product(Cd).to_host().to_dense(), ref) < 1e-12


@pytest.mark.parametrize("merge", [True, False])
def test_coupling_model_dmrg_matches_oracle_and_ed(ctx, merge):
    """dmrg2 / dmrg1 on a CouplingModel (the reference's own test model, test/test_MPS_DMRG.jl:28-36,104-113):
    noisy and noise-free sweeps against the oracle, final energy against exact diagonalisation."""
    T, ob, od, ok, om, op = _imports()
    from oracle import couplingmodel as oc
    N = 8
    sites = om.siteinds("S=1/2", N)
    M = oc.heisenberg_coupling_model(sites, merge=merge)
    psi0 = od.MPS(om.neel_mps(sites))
    prm = dict(maxdim=[8, 20], nsweeps=[2, 3], cutoff=1e-14, noise=[1e-3, 0.0])
    Eo, _, swo = od.dmrg2(psi0, M, od.DMRGParams(**prm))
    Ed, env_d, swd = T.dmrg2(ctx, psi0.t, M, T.DMRGParams(**prm), outputlevel=0)
    assert swd.maxchi == swo.maxchi
    assert abs(swd.energy[-1] - swo.energy[-1]) < 1e-10 * abs(Eo)
    assert abs(Ed - (-3.374932598687897)) < 1e-9
    # one-site sweeps afterwards keep the energy
    E1, _, sw1 = T.dmrg1(ctx, env_d.getpsi(), M, T.DMRGParams(maxdim=[20], nsweeps=[1], cutoff=1e-14), llim=0, rlim=2,
                         outputlevel=0)
    assert abs(E1 - Ed) < 1e-9


def test_coupling_model_noise_term_and_tdvp(ctx):
    """Noisy replacebond! (noiseterm(::ProjCouplingModel), src/mps/projcouplingmodel.jl:391-492) and TDVP sweeps on a
    CouplingModel with next-nearest-neighbour terms, against the oracle."""
    T, ob, od, ok, om, op = _imports()
    from oracle import couplingmodel as oc
    N = 8
    sites, M, mps = _cm_setup(om, od, oc, "S=1", N, 24, 9, 4, merge=True, j2=0.4)
    env_o = od.StateEnvs(mps, M)
    for ortho, pos, center in (("left", 4, 4), ("right", 3, 4)):
        od.orthogonalize(env_o.psi, center)
        env_d = T.StateEnvs(ctx, env_o.psi.t, M, llim=center - 1, rlim=center + 1)
        eo, to, so = od.update_position(env_o, od.eig_solver, pos, 2, ortho, maxdim=20, cutoff=1e-13, noise=1e-3)
        ed, td, sd = T.update_position(env_d, T.eig_solver, pos, 2, ortho, maxdim=20, cutoff=1e-13, noise=1e-3)
        assert abs(ed - eo) < 1e-10 * abs(eo)
        assert len(sd) == len(so) and abs(td - to) < 1e-10
        assert np.allclose(sd, so, rtol=0, atol=1e-9)
    sites = om.siteinds("S=1/2", N)
    M = oc.heisenberg_coupling_model(sites, merge=True, j2=0.3)
    psi0 = od.MPS(om.neel_mps(sites))
    eng_o = od.TDVPEngine(psi0, M)
    eng_d = T.TDVPEngine(ctx, psi0.t, M)
    for ns in (2, 2, 1):
        od.tdvpsweep(eng_o, -0.05, ns, maxdim=16, cutoff=1e-12)
        T.tdvpsweep(eng_d, -0.05, ns, maxdim=16, cutoff=1e-12, outputlevel=0)
    assert eng_d.swdata.maxchi == eng_o.swdata.maxchi
    assert np.allclose(eng_d.swdata.energy, eng_o.swdata.energy, rtol=1e-10, atol=0)


def test_dynamic_tdvp_and_penalty_on_coupling_model(ctx):
    """`tdvpsweep!(nsite="dynamic")` (dynamic_fullsweep!, src/mps/sweep.jl:257-382) and excited-state DMRG with
    StateEnvs(psi, H::CouplingModel, Ms; weight) (ProjCouplingModel_MPS) against the oracle."""
    T, ob, od, ok, om, op = _imports()
    from oracle import couplingmodel as oc
    sites = om.siteinds("S=1/2", 8)
    M = oc.heisenberg_coupling_model(sites, merge=True)
    psi0 = od.MPS(om.neel_mps(sites))
    eng_o, eng_d = od.TDVPEngine(psi0, M), T.TDVPEngine(ctx, psi0.t, M)
    for _ in range(6):
        od.tdvpsweep(eng_o, -0.1, "dynamic", maxdim=6, cutoff=1e-12, extendat=5)
        T.tdvpsweep(eng_d, -0.1, "dynamic", maxdim=6, cutoff=1e-12, extendat=5, outputlevel=0)
    assert eng_d.swdata.maxchi == eng_o.swdata.maxchi
    assert np.allclose(eng_d.swdata.energy, eng_o.swdata.energy, rtol=1e-10, atol=0)
    prm = dict(maxdim=[16, 32], nsweeps=[3, 3], cutoff=1e-14, noise=[1e-3, 0.0])
    E0o, p0o, _ = od.dmrg2(psi0, M, od.DMRGParams(**prm))
    E1o, _, _ = od.dmrg2(psi0, M, od.DMRGParams(**prm), Ms=[p0o], weight=10.0)
    E0d, env0, _ = T.dmrg2(ctx, psi0.t, M, T.DMRGParams(**prm), outputlevel=0)
    E1d, _, _ = T.dmrg2(ctx, psi0.t, M, T.DMRGParams(**prm), Ms=[env0.getpsi()], weight=10.0, outputlevel=0)
    assert abs(E0d - E0o) < 1e-10 * abs(E0o)
    assert abs(E1d - E1o) < 1e-8 * abs(E1o)
    assert abs(E1d - (-2.9822404877)) < 1e-6          # ED: first excited state of S=1/2 N=8 in the Sz=0 sector


# ---------------------------------------------------------------------------------------------- ComplexF64 (planar)
def _complexify(mps, seed):
    """Random complex MPS with the block structure of a real one (same indices)."""
    rng = np.random.default_rng(seed)
    for j, A in enumerate(mps.t):
        blocks = {c: b * (rng.standard_normal(b.shape) + 1j * rng.standard_normal(b.shape)) for c, b in A.blocks.items()}
        mps.t[j] = type(A)(A.inds, blocks, np.complex128)
    return mps


def test_complex_tensor_roundtrip_and_vector_interface(ctx):
    T, ob, od, ok, om, op = _imports()
    _, _, mps = _setup(om, od, "S=1", 8, 24, 3, 1)
    _complexify(mps, 5)
    A, B = mps[4], mps[4].scale(0.5 - 0.25j)
    for nrow in (1, 2, 3):
        d = T.DeviceTensor.from_host(ctx, A, nrow=nrow)
        assert d.is_complex()
        assert np.array_equal(d.to_host().to_dense(), A.to_dense())
    x, y = T.DeviceTensor.from_host(ctx, A, 1), T.DeviceTensor.from_host(ctx, B, 1)
    ref = ob.inner(A, B)
    assert abs(x.dot(y) - ref) < 1e-13 * abs(ref)
    assert abs(x.norm() - A.norm()) < 1e-13 * A.norm()
    y.axpy_(x, -0.3).scale_(2.0)
    assert rel(y.to_host().to_dense(), B.add(A, -0.3).scale(2.0).to_dense()) < 1e-15
    r = T.DeviceTensor.from_host(ctx, _setup(om, od, "S=1", 8, 24, 3, 1)[2][4], 1)
    assert not r.is_complex() and r.promote_().is_complex()


@pytest.mark.parametrize("model", ["mpo", "cm"])
def test_complex_environments_and_apply_match_oracle(ctx, model):
    """Complex MPS: environments L = L A W conj(A), H_eff apply for two-, one- and zero-site vectors, expectation."""
    T, ob, od, ok, om, op = _imports()
    from oracle import couplingmodel as oc
    N = 8
    sites, H, mps = _setup(om, od, "S=1", N, 16, 3, 1)
    if model == "cm":
        H = oc.heisenberg_coupling_model(sites, merge=True, j2=0.3)
    _complexify(mps, 7)
    od.orthogonalize(mps, 4)
    env_o = od.StateEnvs(mps, H)
    env_d = T.StateEnvs(ctx, mps.t, H, llim=3, rlim=5)
    for nsite, pos in ((2, 4), (2, 1), (2, N - 1), (1, 4)):
        od.orthogonalize(env_o.psi, pos)
        env_d = T.StateEnvs(ctx, env_o.psi.t, H, llim=pos - 1, rlim=pos + 1)
        env_o.PH.lpos, env_o.PH.rpos = 0, N + 1
        env_o.set_nsite(nsite); env_o.position(pos)
        env_d.set_nsite(nsite); env_d.position(pos)
        phi_o = ob.contract(env_o.psi[pos], env_o.psi[pos + 1]) if nsite == 2 else env_o.psi[pos]
        phi_d = env_d.make_phi(pos) if nsite == 2 else env_d.site_tensor(pos)
        assert rel(phi_d.to_host().to_dense(), phi_o.to_dense()) < 1e-13
        ref = env_o.product(phi_o).permute(phi_o.inds).to_dense()
        assert rel(env_d.product(phi_d).to_host().to_dense(), ref) < 1e-12
        e_ref = ob.inner(phi_o, env_o.product(phi_o))
        assert abs(env_d.expectation(phi_d) - e_ref.real) < 1e-11 * abs(e_ref) and abs(e_ref.imag) < 1e-10 * abs(e_ref)
    # zero-site
    od.orthogonalize(env_o.psi, 4)
    v4 = env_o.psi[4]
    Lf, Rf, spec, u = ob.factorize(v4, v4.inds[:2], ortho="left", which_decomp="svd", cutoff=0.0)
    env_o.psi[4] = Lf
    env_o.PH.lpos, env_o.PH.rpos = 0, N + 1
    env_d = T.StateEnvs(ctx, env_o.psi.t, H, llim=4, rlim=6)
    env_o.set_nsite(0); env_o.position(5)
    env_d.set_nsite(0); env_d.position(5)
    Cd = T.DeviceTensor.from_host(ctx, Rf, 1)
    assert rel(env_d.product(Cd).to_host().to_dense(), env_o.product(Rf).permute(Rf.inds).to_dense()) < 1e-12


@pytest.mark.parametrize("nsite,t", [(2, -0.1j), (1, 0.05j), (2, -0.05 - 0.08j), (0, 0.07j)])
def test_exp_solver_complex_time_matches_oracle(ctx, nsite, t):
    """Real-time evolution exp(-i dt H_eff): a real vector is promoted to ComplexF64 on the device."""
    T, ob, od, ok, om, op = _imports()
    sites, H, mps = _setup(om, od, "S=1", 8, 24, 3, 4)
    env_o = od.StateEnvs(mps, H)
    pos = 4
    if nsite == 0:
        v_o = env_o.psi[4]
        L, R, spec, u = ob.factorize(v_o, v_o.inds[:2], ortho="left", which_decomp="svd", cutoff=0.0)
        env_o.psi[4] = L
        phi_o, pos = R, 5
        env_d = T.StateEnvs(ctx, env_o.psi.t, H, llim=4, rlim=6)
    else:
        env_d = T.StateEnvs(ctx, mps.t, H, llim=3, rlim=5)
        phi_o = ob.contract(env_o.psi[4], env_o.psi[5]) if nsite == 2 else env_o.psi[4]
    env_o.set_nsite(nsite); env_o.position(pos)
    env_d.set_nsite(nsite); env_d.position(pos)
    phi_d = T.DeviceTensor.from_host(ctx, phi_o, 1)
    _, out_o = od.exp_solver(env_o, phi_o, t)
    _, out_d = T.exp_solver(env_d, phi_d, t)
    assert out_d.is_complex()
    assert env_d.last_solver_info["converged"] == 1
    assert abs(env_d.last_solver_info["numops"] - od.exp_solver.last_info["numops"]) <= 1
    assert rel(out_d.to_host().to_dense(), out_o.permute(phi_o.inds).to_dense()) < 1e-10
    if t.real == 0:
        assert abs(out_d.norm() - phi_o.norm()) < 1e-11 * phi_o.norm()          # unitary


@pytest.mark.parametrize("ortho", ["left", "right"])
def test_complex_replacebond_matches_oracle(ctx, ortho):
    """Truncation of a complex two-site tensor through the Hermitian eigenproblem of M M^+ / M^+ M: kept spectrum,
    truncation error, new link sectors; the factors are compared through their product and isometry."""
    T, ob, od, ok, om, op = _imports()
    sites, H, mps = _setup(om, od, "S=1", 8, 24, 3, 4)
    _complexify(mps, 11)
    od.orthogonalize(mps, 4)
    pos = 4 if ortho == "left" else 3
    env_d = T.StateEnvs(ctx, mps.t, H, llim=3, rlim=5)
    env_d.set_nsite(2)
    phi_d = env_d.make_phi(pos)
    phi_o = ob.contract(mps[pos], mps[pos + 1])
    phi_o = phi_o.scale(1.0 / phi_o.norm())
    phi_d.scale_(1.0 / phi_d.norm())
    psi_o = mps.copy()
    spec = od.replacebond(psi_o, pos, phi_o, maxdim=12, mindim=1, cutoff=1e-13, eigen_perturbation=None, ortho=ortho,
                          normalize=True)
    terr, eigs = env_d.replacebond(pos, phi_d, maxdim=12, mindim=1, cutoff=1e-13, noise=0.0, ortho=ortho, normalize=True)
    assert len(eigs) == len(spec.eigs) and abs(terr - spec.truncerr) < 1e-12
    assert np.allclose(eigs, spec.eigs, rtol=0, atol=1e-12)
    Ad, Bd = env_d.site_tensor(pos).to_host(), env_d.site_tensor(pos + 1).to_host()
    assert [ix.dims for ix in Ad.inds] == [ix.dims for ix in psi_o[pos].inds]
    two_d = np.tensordot(Ad.to_dense(), Bd.to_dense(), axes=([2], [0]))
    two_o = np.tensordot(psi_o[pos].to_dense(), psi_o[pos + 1].to_dense(), axes=([2], [0]))
    assert rel(two_d, two_o) < 1e-10                                          # gauge-invariant product
    iso = Ad.to_dense() if ortho == "left" else Bd.to_dense()
    m = iso.reshape(-1, iso.shape[2]) if ortho == "left" else iso.reshape(iso.shape[0], -1).T
    assert np.abs(m.conj().T @ m - np.eye(m.shape[1])).max() < 1e-12


@pytest.mark.parametrize("nsite", [2, 1])
def test_tdvp_real_time_matches_oracle_and_exact_evolution(ctx, nsite):
    """tdvpsweep!(engine, -im*dt) (the reference's TDVP test step, test/test_MPS_TDVP.jl:50-56): energies per sweep
    against the oracle, energy conservation, and the state against exp(-i t H)|Neel> from dense linear algebra."""
    import scipy.linalg as sl
    T, ob, od, ok, om, op = _imports()
    N = 8
    sites = om.siteinds("S=1/2", N)
    H = om.heisenberg_mpo(sites)
    psi0 = od.MPS(om.neel_mps(sites))
    eng_o, eng_d = od.TDVPEngine(psi0, H), T.TDVPEngine(ctx, psi0.t, H)
    sched = [2, 2] + [nsite] * 3
    dt = 0.05
    for ns in sched:
        od.tdvpsweep(eng_o, -1j * dt, ns, maxdim=16, cutoff=1e-13)
        T.tdvpsweep(eng_d, -1j * dt, ns, maxdim=16, cutoff=1e-13, outputlevel=0)
    assert eng_d.swdata.maxchi == eng_o.swdata.maxchi
    assert np.allclose(eng_d.swdata.energy, eng_o.swdata.energy, rtol=1e-10, atol=0)
    assert np.allclose(eng_d.swdata.maxtruncerr, eng_o.swdata.maxtruncerr, rtol=0, atol=1e-12)
    assert max(abs(e - eng_d.swdata.energy[0]) for e in eng_d.swdata.energy) < 1e-8      # unitary evolution
    Hd = om.mpo_to_dense(H)
    v = sl.expm(-1j * dt * len(sched) * Hd) @ om.mps_to_dense(psi0.t)
    w = om.mps_to_dense(eng_d.getpsi())
    assert abs(abs(np.vdot(v, w)) - 1.0) < 1e-8
    wo = om.mps_to_dense(eng_o.sysenv.psi.t)
    assert abs(abs(np.vdot(wo, w)) - 1.0) < 1e-10


def test_baseline_config0_matches_fixture_and_literature(ctx):
    """BASELINE.json configs[0]: S=1/2 Heisenberg chain N=20, two-site DMRG, maxdim 20 -> 64 -- per-sweep energies,
    bond dimensions and truncation errors against the committed oracle fixture, final energy against the literature
    value -8.682473334399."""
    T, ob, od, ok, om, op = _imports()
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "oracle_golden_r01b.json")))["config0"][0]
    sites = om.siteinds(g["kind"], g["N"])
    H = om.heisenberg_mpo(sites)
    e, env, sw = T.dmrg2(ctx, om.neel_mps(sites), H, T.DMRGParams(**g["params"]), outputlevel=0)
    assert sw.maxchi[-4:] == g["maxchi"][-4:] and sw.maxchi[:5] == g["maxchi"][:5]
    de = np.abs(np.array(sw.energy) - np.array(g["energy"]))
    # the first stage is truncation limited (maxdim 20) right after two noisy sweeps whose null-space eigenvectors are
    # LAPACK-build dependent; the converged second stage is reproducible to rounding
    assert de.max() < 1e-6 * abs(g["energy"][-1])
    assert de[-3:].max() < 1e-9 * abs(g["energy"][-1])
    assert abs(e - g["ed_literature"]) < 1e-9


def test_coupling_model_with_multi_site_terms_matches_oracle(ctx):
    """CouplingModel terms with three and four operators (middle tensors with two open OpLinks), terms skipping
    several sites, merged and unmerged: H_eff apply at several positions and a DMRG run against the oracle."""
    T, ob, od, ok, om, op = _imports()
    from oracle import couplingmodel as oc
    from tests.test_oracle_kat import _multi_site_model
    N = 8
    sites, os_ = _multi_site_model(om, oc, N)
    qn, dm = om.gaussian_link_sectors(8, 1.3, 4, step=1)
    mps = od.MPS(om.random_mps(sites, qn, dm, np.random.default_rng(4)))
    for merge in (True, False):
        M = oc.coupling_model(os_, sites, merge=merge)
        for nsite, pos in ((2, 3), (2, 1), (2, 7), (1, 4)):
            od.orthogonalize(mps, pos)
            env_o = od.StateEnvs(mps, M)
            env_d = T.StateEnvs(ctx, mps.t, M, llim=pos - 1, rlim=pos + 1)
            env_o.set_nsite(nsite); env_o.position(pos)
            env_d.set_nsite(nsite); env_d.position(pos)
            phi_o = ob.contract(env_o.psi[pos], env_o.psi[pos + 1]) if nsite == 2 else env_o.psi[pos]
            phi_d = env_d.make_phi(pos) if nsite == 2 else env_d.site_tensor(pos)
            ref = env_o.product(phi_o).permute(phi_o.inds).to_dense()
            assert rel(env_d.product(phi_d).to_host().to_dense(), ref) < 1e-12
    M = oc.coupling_model(os_, sites, merge=True)
    prm = dict(maxdim=[16, 32], nsweeps=[3, 3], cutoff=1e-14, noise=[1e-3, 0.0])
    Eo, _, swo = od.dmrg2(od.MPS(om.neel_mps(sites)), M, od.DMRGParams(**prm))
    Ed, _, swd = T.dmrg2(ctx, om.neel_mps(sites), M, T.DMRGParams(**prm), outputlevel=0)
    assert swd.maxchi[-1] == swo.maxchi[-1]
    assert abs(Ed - Eo) < 1e-9 * abs(Eo)

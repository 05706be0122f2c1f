"""On-device measurements and `updateH!` (`-m gpu`) against dense linear algebra on
the exported state: `bond_spectrum`, `entropy`, `measure` (src/mps/measure.jl:24-343), `updateH!` with and without
`recalcEnv` (src/mps/state_envs.jl:181-208)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _imports():
    import tennetlib.jl_b200 as T
    from tennetlib.jl_b200 import measure as dm
    from oracle import blocksparse as ob, dmrg as od, models as om
    return T, dm, ob, od, om


def _site_op(ob, om, s, name):
    ops = om.spin_ops(om.site_S2(s))
    dense = {"Sz": ops["Sz"], "S+": ops["Sp"], "S-": ops["Sm"], "Id": ops["Id"]}[name]
    step = s.qns[0][0] - s.qns[1][0]
    flux = {"Sz": 0, "Id": 0, "S+": step, "S-": -step}[name]
    return ob.BSTensor.from_dense([s.prime().copy(dir=+1), s.copy(dir=-1)], dense, flux=(flux,)), dense


def _dense_expect(v, d, N, ops):
    """<v| prod_j O_j |v> for {site: dense d x d}"""
    psi = v.reshape([d] * N)
    out = psi
    for j, O in ops.items():
        out = np.moveaxis(np.tensordot(O, out, axes=([1], [j - 1])), 0, j - 1)
    return np.vdot(psi.reshape(-1), out.reshape(-1))


@pytest.mark.parametrize("cplx", [False, True])
def test_bond_spectrum_entropy_and_measure_match_dense(ctx, cplx):
    T, dm, ob, od, om = _imports()
    from helpers import to_oracle
    N = 8
    sites = om.siteinds("S=1", N)
    H = om.heisenberg_mpo(sites)
    qn, dims = om.gaussian_link_sectors(14, 1.3, 4)
    mps = od.MPS(om.random_mps(sites, qn, dims, np.random.default_rng(4)))
    if cplx:
        rng = np.random.default_rng(5)
        for A in mps.t:
            A.dtype = np.complex128
            for c in list(A.blocks):
                A.blocks[c] = A.blocks[c] + 1j * rng.standard_normal(A.blocks[c].shape)
    env = T.StateEnvs(ctx, mps.t, H)
    env.orthogonalize(3)
    v = om.mps_to_dense([to_oracle(A) for A in env.getpsi()])
    nrm = np.linalg.norm(v)
    v = v / nrm
    # normalise the device state too (centre tensor)
    c = env.site_tensor(3).copy()
    c.scale_(1.0 / nrm)
    env.set_site_tensor(3, c)
    d = 3
    for bond in (1, 4, 7):
        s = np.linalg.svd(v.reshape(d ** bond, -1), compute_uv=False) ** 2
        s = s[s > 1e-28]
        w = dm.bond_spectrum(env, bond, sites=sites)
        assert len(w) >= len(s) and np.abs(w[:len(s)] - s).max() < 1e-13 and np.all(w[len(s):] < 1e-15)   # structural zeros: eps * sigma_max^2 through the Gram route
        assert abs(dm.entropy(env, bond, sites=sites) - float(-np.sum(s * np.log(s)))) < 1e-11
    byq = dm.bond_spectrum(env, 4, by_charge=True, sites=sites)
    allw = np.sort(np.concatenate([w for _, w in byq]))[::-1]
    ref = dm.bond_spectrum(env, 4, sites=sites)
    assert np.abs(allw[:len(ref)] - ref).max() < 1e-13
    assert len(dm.entropy(env, sites=sites)) == N - 1
    # local and multi-site expectation values
    Sz = {j: _site_op(ob, om, sites[j - 1], "Sz") for j in range(1, N + 1)}
    Sp = {j: _site_op(ob, om, sites[j - 1], "S+") for j in range(1, N + 1)}
    Sm = {j: _site_op(ob, om, sites[j - 1], "S-") for j in range(1, N + 1)}
    for j in (1, 4, 8):
        assert abs(dm.measure(env, Sz[j][0], sites=sites) - _dense_expect(v, d, N, {j: Sz[j][1]})) < 1e-12
    assert abs(dm.measure(env, Sp[2][0], sites=sites)) == 0.0                      # charged operator in a charge eigenstate
    for (i, j) in [(2, 3), (1, 8), (3, 6)]:
        zz = dm.measure(env, [Sz[i][0], Sz[j][0]], sites=sites)
        assert abs(zz - _dense_expect(v, d, N, {i: Sz[i][1], j: Sz[j][1]})) < 1e-12
        pm = dm.measure(env, [Sp[i][0], Sm[j][0]], sites=sites)
        assert abs(pm - _dense_expect(v, d, N, {i: Sp[i][1], j: Sm[j][1]})) < 1e-12
    three = dm.measure(env, [Sp[2][0], Sz[4][0], Sm[7][0], Sz[4][0]], sites=sites)
    assert abs(three - _dense_expect(v, d, N, {2: Sp[2][1], 4: Sz[4][1] @ Sz[4][1], 7: Sm[7][1]})) < 1e-12
    if not cplx:
        assert isinstance(dm.measure(env, Sz[2][0], sites=sites, real=True), float)


def test_updateH_with_and_without_recalc(ctx):
    """recalcEnv = false keeps every cached environment (only W changes); recalcEnv = true rebuilds them."""
    T, dm, ob, od, om = _imports()
    N = 8
    sites = om.siteinds("S=1/2", N)
    H = om.heisenberg_mpo(sites)
    qn, dims = om.gaussian_link_sectors(16, 1.3, 4, step=1)
    mps = od.MPS(om.random_mps(sites, qn, dims, np.random.default_rng(9)))
    env = T.StateEnvs(ctx, mps.t, H)
    env.orthogonalize(1)
    env.set_nsite(2)
    phi = env.make_phi(1)
    env.position(1)
    hv0 = env.product(phi).to_host().to_dense()

    def scaled(site, c):
        out = list(H)
        W = H[site]
        out[site] = ob.BSTensor(W.inds, {k: c * b for k, b in W.blocks.items()})
        return out
    env.updateH(scaled(0, 2.0), recalcEnv=False)           # W_1 changes: exact, the right environments do not contain it
    env.position(1)
    assert np.abs(env.product(phi).to_host().to_dense() - 2.0 * hv0).max() < 1e-13 * np.abs(hv0).max()
    env.updateH(scaled(N - 1, 3.0), recalcEnv=False)       # W_N changes: the cached right environments are REUSED (stale)
    env.position(1)
    assert np.abs(env.product(phi).to_host().to_dense() - hv0).max() < 1e-13 * np.abs(hv0).max()
    env.updateH(scaled(N - 1, 3.0), recalcEnv=True)        # fresh projected Hamiltonian
    env.position(1)
    assert np.abs(env.product(phi).to_host().to_dense() - 3.0 * hv0).max() < 1e-12 * np.abs(hv0).max()
    with pytest.raises(Exception):
        env.updateH([H, H], recalcEnv=False)
    # a sweep after updateH(recalcEnv = true) equals a sweep on a freshly built StateEnvs
    E1, _, _ = T.update_position(env, T.eig_solver, 1, 2, "left", maxdim=16, cutoff=1e-14)
    env2 = T.StateEnvs(ctx, mps.t, scaled(N - 1, 3.0))
    env2.orthogonalize(1)
    E2, _, _ = T.update_position(env2, T.eig_solver, 1, 2, "left", maxdim=16, cutoff=1e-14)
    assert abs(E1 - E2) < 1e-12 * abs(E2)

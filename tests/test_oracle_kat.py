"""Pins the CPU oracle (the parity checker) against exact diagonalisation and dense linear algebra.
The reference ships no golden vectors (test/test_MPS_DMRG.jl:100-146 has no assertions) -- these KATs are
what stands in for them."""
import numpy as np
import pytest

from oracle import blocksparse as ob, dmrg as od, krylov as ok, models as om
from tests.ed import heisenberg_dense, lowest_energies

E0_S12_N12 = -5.1420906328405      # SURVEY.md section 8c, re-derived below
E0_S1_N8 = -10.1246372223589


def test_ed_values_rederived():
    assert abs(lowest_energies(12, 1, 1)[0] - E0_S12_N12) < 1e-11
    assert abs(lowest_energies(8, 2, 1)[0] - E0_S1_N8) < 1e-11


@pytest.mark.parametrize("kind,S2,N", [("S=1/2", 1, 6), ("S=1", 2, 4)])
def test_mpo_equals_dense_hamiltonian(kind, S2, N):
    sites = om.siteinds(kind, N)
    H = om.heisenberg_mpo(sites)
    assert np.abs(om.mpo_to_dense(H) - heisenberg_dense(N, S2)).max() < 1e-14
    for W in H:   # all MPO weight sits in flux-0 blocks
        assert np.abs(ob.BSTensor.from_dense(W.inds, W.to_dense()).to_dense() - W.to_dense()).max() == 0


def test_contract_matches_dense_and_counts_flops():
    rng = np.random.default_rng(0)
    sites = om.siteinds("S=1", 4)
    qn, dm = om.gaussian_link_sectors(12, 1.3, 3)
    psi = om.random_mps(sites, qn, dm, rng)
    A, B = psi[1], psi[2]
    ob.reset_flops()
    C = ob.contract(A, B)
    assert ob.get_flops() > 0
    ref = np.tensordot(A.to_dense(), B.to_dense(), axes=([2], [0]))
    assert np.abs(C.to_dense() - ref).max() < 1e-13
    # flat NDTensors-style export / import round trip
    c, o, d = C.export_flat()
    assert np.abs(ob.BSTensor.import_flat(C.inds, c, o, d).to_dense() - C.to_dense()).max() == 0


def test_heff_apply_equals_projected_dense_hamiltonian():
    """H_eff v == P^T H P v on a small chain (invariant listed in SURVEY.md section 8c)."""
    rng = np.random.default_rng(1)
    N = 6
    sites = om.siteinds("S=1/2", N)
    H = om.heisenberg_mpo(sites)
    qn, dm = om.gaussian_link_sectors(6, 1.0, 3, step=1)
    mps = od.MPS(om.random_mps(sites, qn, dm, rng))
    od.orthogonalize(mps, 3)
    env = od.StateEnvs(mps, H)
    env.set_nsite(2)
    env.position(3)
    phi = ob.contract(env.psi[3], env.psi[4])
    Hv = env.product(phi)
    # dense: <psi(v')|H|psi(v)> with the other tensors fixed
    Hd = heisenberg_dense(N, 1)

    def full(two):
        t = [x for x in env.psi.t]
        acc = t[0].to_dense()[0]
        acc = np.tensordot(acc, t[1].to_dense(), axes=([-1], [0]))
        acc = np.tensordot(acc, two, axes=([-1], [0]))
        acc = np.tensordot(acc, t[4].to_dense(), axes=([-1], [0]))
        acc = np.tensordot(acc, t[5].to_dense(), axes=([-1], [0]))
        return acc[..., 0].reshape(-1)
    v = phi.to_dense()
    e_dense = full(v) @ Hd @ full(v)
    e_env = ob.inner(phi, Hv)
    assert abs(e_dense - e_env) < 1e-11 * abs(e_dense)
    # symmetry of H_eff
    w = ob.BSTensor.random(phi.inds, rng)
    assert abs(ob.inner(w, env.product(phi)) - ob.inner(env.product(w), phi)) < 1e-10 * abs(e_dense)


def test_lanczos_matches_eigvalsh_and_default_budget():
    rng = np.random.default_rng(2)
    n = 120
    A = rng.standard_normal((n, n)); A = (A + A.T) / 2
    w = np.linalg.eigvalsh(A)
    x0 = rng.standard_normal(n)
    e, v, info = ok.eigsolve_lanczos(lambda x: A @ x, x0, krylovdim=5, maxiter=2)
    assert info["numops"] == 7 and info["numiter"] == 2          # <= 7 applies
    assert abs(np.linalg.norm(v) - 1) < 1e-12
    e, v, info = ok.eigsolve_lanczos(lambda x: A @ x, x0, krylovdim=30, maxiter=200, tol=1e-12)
    assert info["converged"] == 1 and abs(e - w[0]) < 1e-10
    assert np.linalg.norm(A @ v - e * v) < 1e-9
    # thick restart keeps the basis orthonormal over many restarts (Householder sigma without cancellation)
    e, v, info = ok.eigsolve_lanczos(lambda x: A @ x, x0, krylovdim=5, maxiter=400, tol=1e-13)
    assert abs(e - w[0]) < 1e-9


def test_truncate_spectrum_rules():
    P = np.array([0.5, 0.3, 0.1, 0.06, 0.04])
    kept, err, docut = ob.truncate_spectrum(P, maxdim=3, mindim=1, cutoff=0.0)
    assert len(kept) == 3 and abs(err - 0.1) < 1e-15 and abs(docut - 0.08) < 1e-15
    kept, err, docut = ob.truncate_spectrum(P, maxdim=None, mindim=1, cutoff=0.05)   # relative cumulative cutoff
    assert len(kept) == 4 and abs(err - 0.04) < 1e-15
    kept, err, docut = ob.truncate_spectrum(np.array([0.4, 0.3, 0.3]), maxdim=2, mindim=1, cutoff=0.0)
    assert docut > 0.3          # degenerate pair at the cut: docut is pushed above both (ITensors quirk)
    kept, err, docut = ob.truncate_spectrum(np.array([1.0]), maxdim=1)
    assert docut == 0.5 and err == 0.0
    kept, err, docut = ob.truncate_spectrum(np.array([1.0, 0.0, 0.0]), cutoff=0.0)    # exact zeros go with <=
    assert len(kept) == 1


@pytest.mark.parametrize("ortho", ["left", "right"])
@pytest.mark.parametrize("which", ["svd", "eigen"])
def test_factorize_identity_and_isometry(ortho, which):
    rng = np.random.default_rng(3)
    sites = om.siteinds("S=1", 4)
    qn, dm = om.gaussian_link_sectors(10, 1.3, 3)
    psi = om.random_mps(sites, qn, dm, rng)
    phi = ob.contract(psi[1], psi[2])
    L, R, spec, u = ob.factorize(phi, phi.inds[:2], ortho=ortho, which_decomp=which, cutoff=0.0)
    assert np.abs(ob.contract(L, R).to_dense() - phi.to_dense()).max() < 1e-11
    iso = L if ortho == "left" else R
    other = [ix for ix in iso.inds if ix != u]
    G = ob.contract(iso, iso.dag().prime(1, [u])).to_dense()
    assert np.abs(G - np.eye(G.shape[0])).max() < 1e-10
    assert abs(spec.eigs.sum() - phi.norm() ** 2) < 1e-10 * phi.norm() ** 2
    # truncation error = discarded weight / total weight
    L2, R2, spec2, _ = ob.factorize(phi, phi.inds[:2], ortho=ortho, which_decomp=which, maxdim=5, cutoff=0.0)
    assert len(spec2.eigs) <= 5
    assert abs(spec2.truncerr - (1 - spec2.eigs.sum() / spec.eigs.sum())) < 1e-12


def test_dmrg_reference_test_case_matches_ed():
    """The reference's own smoke test (test/test_MPS_DMRG.jl:7-66: S=1/2 N=12 Neel start, 5 sweeps maxdim 20,
    cutoff 1e-14, noise 1e-3 decay 2, disabled after 2) with the assertion it lacks."""
    N = 12
    sites = om.siteinds("S=1/2", N)
    H = om.heisenberg_mpo(sites)
    psi0 = od.MPS(om.neel_mps(sites))
    p = od.DMRGParams(nsweeps=[5], maxdim=[20], cutoff=1e-14, noise=1e-3, noisedecay=2, disable_noise_after=2)
    e, psi, sw = od.dmrg2(psi0, H, p)
    assert sw.maxchi == [20] * 5
    assert e > E0_S12_N12 - 1e-12 and e - E0_S12_N12 < 1e-8          # variational, truncation-limited
    assert all(b <= a + 1e-9 for a, b in zip(sw.energy, sw.energy[1:]))
    p = od.DMRGParams(nsweeps=[6], maxdim=[64], cutoff=1e-14)
    e, psi, sw = od.dmrg2(psi0, H, p)
    assert abs(e - E0_S12_N12) < 1e-11
    assert abs(om.mps_norm(psi.t) - 1) < 1e-12


def test_dmrg_spin1_matches_ed():
    N = 8
    sites = om.siteinds("S=1", N)
    H = om.heisenberg_mpo(sites)
    psi0 = od.MPS(om.neel_mps(sites))
    p = od.DMRGParams(nsweeps=[3, 4], maxdim=[30, 200], cutoff=1e-14, noise=[1e-4, 0.0])
    e, psi, sw = od.dmrg2(psi0, H, p)
    assert abs(e - E0_S1_N8) < 1e-10


def test_excited_state_penalty_matches_ed():
    """dmrg_ex of the reference test (weight = 10): converges to E1 of the Sz=0 sector."""
    e = lowest_energies(12, 1, 2)
    assert abs(e[1] - (-4.8611479370364)) < 1e-10
    N = 12
    sites = om.siteinds("S=1/2", N)
    H = om.heisenberg_mpo(sites)
    psi0 = od.MPS(om.neel_mps(sites))
    p = od.DMRGParams(nsweeps=[5], maxdim=[20], cutoff=1e-14, noise=1e-3, noisedecay=2, disable_noise_after=2)
    e0, gs, _ = od.dmrg2(psi0, H, p)
    e1, ex, sw = od.dmrg2(psi0, H, p, Ms=[gs], weight=10.0)
    assert e1 > e[1] - 1e-9 and e1 - e[1] < 1e-6


def test_one_site_dmrg_matches_ed():
    N = 12
    sites = om.siteinds("S=1/2", N)
    H = om.heisenberg_mpo(sites)
    env = od.StateEnvs(od.MPS(om.neel_mps(sites)), H)
    od.dmrg_(env, od.DMRGParams(nsweeps=[4], maxdim=[64], cutoff=1e-14, noise=1e-3, noisedecay=2, disable_noise_after=2), 2)
    sw = od.dmrg_(env, od.DMRGParams(nsweeps=[2], maxdim=[64], cutoff=0.0), 1)       # docs/src/mps/example_dmrg.md:62-71
    assert abs(sw.energy[-1] - E0_S12_N12) < 1e-10


# ---------------------------------------------------------------------------------------------- exponentiate / TDVP
@pytest.mark.parametrize("n,kd,t,eager", [(60, 30, -0.3, True), (150, 30, -0.05j, True), (150, 10, -0.5, False),
                                          (200, 8, 0.4j, False), (50, 30, 0.2, True)])
def test_exponentiate_matches_dense_expm(n, kd, t, eager):
    """KrylovKit.exponentiate restatement vs scipy.linalg.expm on a dense symmetric matrix (one-shot eager exit,
    sub-stepping with restarts, real and imaginary steps)."""
    import scipy.linalg as sl
    from oracle.krylov import exponentiate
    rng = np.random.default_rng(n)
    M = rng.standard_normal((n, n)); M = (M + M.T) / 2
    x = rng.standard_normal(n); x /= np.linalg.norm(x)
    y, info = exponentiate(lambda v: M @ v, t, x, krylovdim=kd, maxiter=1000, eager=eager)
    ref = sl.expm(t * M) @ x
    assert info["converged"] == 1
    assert np.linalg.norm(y - ref) < 5e-12 * np.linalg.norm(ref)
    if not eager:
        assert info["numiter"] > 1           # the basis filled up: sub-steps were taken


def test_exponentiate_fixed_point_and_zero_step():
    from oracle.krylov import exponentiate
    x = np.ones(5) / np.sqrt(5)
    y, info = exponentiate(lambda v: 0.0 * v, -0.1, x)
    assert info["converged"] == 1 and np.array_equal(y, x)
    y, info = exponentiate(lambda v: v, 0.0, x)
    assert info["numops"] == 0 and np.array_equal(y, x)


@pytest.mark.parametrize("nsite,ts", [(2, -0.05), (2, -0.05j), (1, -0.05), (1, -0.05j)])
def test_tdvp_sweeps_follow_exact_evolution(nsite, ts):
    """tdvpsweep! restatement (two-site / one-site with backward steps) vs exp(t H)|Neel> for S=1/2 N=8;
    real time conserves the energy, imaginary time lowers it."""
    import scipy.linalg as sl
    from oracle import dmrg as od, models as om
    N = 8
    sites = om.siteinds("S=1/2", N)
    H = om.heisenberg_mpo(sites)
    Hd = om.mpo_to_dense(H)
    psi0 = od.MPS(om.neel_mps(sites))
    v0 = om.mps_to_dense(psi0.t)
    eng = od.TDVPEngine(psi0, H)
    nsteps = 0
    if nsite == 1:                      # one-site TDVP cannot grow the bonds: start with two-site sweeps
        for _ in range(2):
            od.tdvpsweep(eng, ts, 2, maxdim=16, cutoff=1e-14); nsteps += 1
    for _ in range(3):
        od.tdvpsweep(eng, ts, nsite, maxdim=16, cutoff=1e-14); nsteps += 1
    v = sl.expm(nsteps * ts * Hd) @ v0
    v /= np.linalg.norm(v)
    w = om.mps_to_dense(eng.sysenv.psi.t)
    assert abs(abs(np.vdot(v, w)) - 1.0) < 1e-9
    assert abs(eng.swdata.energy[-1] - np.real(np.vdot(v, Hd @ v))) < 1e-5
    assert abs(eng.abstime - nsteps * 0.05) < 1e-14
    if isinstance(ts, complex):
        assert abs(eng.swdata.energy[-1] - eng.swdata.energy[0]) < 1e-9      # energy conservation
    else:
        assert eng.swdata.energy[-1] < eng.swdata.energy[0]


def test_mpo_sum_equals_single_mpo():
    """ProjMPOSum2 restatement: H = H_zz + H_xy as two MPOs gives the same H_eff apply, noise term and DMRG energy
    as the single MPO (src/mps/projmposum2.jl:86-145)."""
    from oracle import blocksparse as ob, dmrg as od, models as om
    N = 8
    sites = om.siteinds("S=1/2", N)
    H = om.heisenberg_mpo(sites)
    Hs = [om.heisenberg_mpo(sites, Jz=1.0, Jxy=0.0), om.heisenberg_mpo(sites, Jz=0.0, Jxy=1.0)]
    qn, dm = om.gaussian_link_sectors(12, 1.3, 4, step=1)
    mps = od.MPS(om.random_mps(sites, qn, dm, np.random.default_rng(5)))
    od.orthogonalize(mps, 4)
    e1, e2 = od.StateEnvs(mps, H), od.StateEnvs(mps, Hs)
    for e in (e1, e2):
        e.set_nsite(2); e.position(4)
    phi = ob.contract(mps[4], mps[5])
    a, b = e1.product(phi), e2.product(phi)
    assert np.abs(a.to_dense() - b.permute(a.inds).to_dense()).max() < 1e-13
    na, nb = e1.PH.noiseterm(phi, "left"), e2.PH.noiseterm(phi, "left")
    # the noise term is quadratic in H: sum_k (L W_k phi)(..)^dag != (L W phi)(..)^dag in general; only its structure
    # (indices, hermiticity, positivity) is shared
    assert [ix.dim for ix in na.inds] == [ix.dim for ix in nb.inds]
    psi0 = od.MPS(om.neel_mps(sites))
    prm = od.DMRGParams(maxdim=[16, 32], nsweeps=[3, 3], cutoff=1e-14, noise=[1e-3, 0.0])
    E1, _, _ = od.dmrg2(psi0, H, prm)
    E2, _, _ = od.dmrg2(psi0, Hs, prm)
    E0 = -3.374932598687897          # ED, S=1/2 N=8 OBC
    assert abs(E1 - E0) < 1e-9 and abs(E2 - E0) < 1e-9


def _mpo_as_coupling_model(H, sites):
    """CouplingModel(mpos::MPO...) (src/base/couplingmodel.jl:352-379): one id, links retagged "OpLink"."""
    from oracle import blocksparse as ob, couplingmodel as oc
    N = len(H)
    terms = [dict() for _ in range(N)]
    for j, W in enumerate(H):
        inds = [ix.copy(tags="OpLink") if "Link" in ix.tags else ix for ix in W.inds]
        blocks = dict(W.blocks)
        if j == 0:
            inds, blocks = inds[1:], {c[1:]: b[0] for c, b in blocks.items()}
        if j == N - 1:
            inds, blocks = inds[:-1], {c[:-1]: b[..., 0] for c, b in blocks.items()}
        terms[j][7] = ob.BSTensor(inds, blocks)
    return oc.CouplingModel(sites, terms)


def test_coupling_model_restatement_against_dense_and_mpo():
    """CouplingModel / ProjCouplingModel restatement: the model equals the dense Hamiltonian (merged and unmerged
    terms, next-nearest-neighbour terms that skip a site, one-site terms), <phi|H_eff|phi> equals the dense
    expectation at several positions, and a single-id model built from an MPO reproduces the ProjMPO update
    (energy, truncation error, spectrum) including the noise term on both sweep directions."""
    from oracle import blocksparse as ob, couplingmodel as oc, dmrg as od, models as om
    N = 6
    for kind in ("S=1/2", "S=1"):
        sites = om.siteinds(kind, N)
        Hd = om.mpo_to_dense(om.heisenberg_mpo(sites))
        for merge in (True, False):
            M = oc.heisenberg_coupling_model(sites, merge=merge)
            assert np.abs(oc.coupling_model_to_dense(M) - Hd).max() < 1e-14
    M = oc.heisenberg_coupling_model(sites, merge=True, field=0.3, j2=0.5)
    Md = oc.coupling_model_to_dense(M)
    assert np.abs(Md - Md.T).max() < 1e-14
    qn, dm = om.gaussian_link_sectors(10, 1.3, 4, step=2)
    mps = od.MPS(om.random_mps(sites, qn, dm, np.random.default_rng(5)))
    v = None
    for nsite, pos in ((2, 3), (1, 3), (2, 1), (2, 5)):
        od.orthogonalize(mps, pos)
        P = oc.ProjCouplingModel(M)
        P.set_nsite(nsite); P.position(mps.t, pos)
        phi = ob.contract(mps[pos], mps[pos + 1]) if nsite == 2 else mps[pos]
        v = om.mps_to_dense(mps.t)
        assert abs(ob.inner(phi, P.product(phi)) / ob.inner(phi, phi) - v @ Md @ v / (v @ v)) < 1e-12
    H = om.heisenberg_mpo(sites)
    Mh = _mpo_as_coupling_model(H, sites)
    od.orthogonalize(mps, 3)
    for ortho, pos in (("left", 3), ("right", 2)):
        res = [od.update_position(od.StateEnvs(mps, X), od.eig_solver, pos, 2, ortho, maxdim=8, cutoff=1e-13, noise=1e-3)
               for X in (H, Mh)]
        assert abs(res[0][0] - res[1][0]) < 1e-12
        assert abs(res[0][1] - res[1][1]) < 1e-13
        assert np.abs(res[0][2] - res[1][2]).max() < 1e-13


def test_coupling_model_dmrg_and_tdvp_kat():
    """The reference's own test model (test/test_MPS_DMRG.jl:28-36,104-113) through the restated ProjCouplingModel:
    DMRG reaches the ED ground-state energy for merged and unmerged terms; TDVP lowers the energy in imaginary time."""
    from oracle import couplingmodel as oc, dmrg as od, models as om
    sites = om.siteinds("S=1/2", 8)
    for merge in (True, False):
        M = oc.heisenberg_coupling_model(sites, merge=merge)
        E, _, _ = od.dmrg2(od.MPS(om.neel_mps(sites)), M,
                           od.DMRGParams(maxdim=[8, 20], nsweeps=[2, 3], cutoff=1e-14, noise=[1e-3, 0.0]))
        assert abs(E - (-3.374932598687897)) < 1e-10
    eng = od.TDVPEngine(od.MPS(om.neel_mps(sites)), oc.heisenberg_coupling_model(sites, merge=True, j2=0.3))
    for ns in (2, 2, 1):
        od.tdvpsweep(eng, -0.05, ns, maxdim=16, cutoff=1e-12)
    assert eng.swdata.energy[0] > eng.swdata.energy[1] > eng.swdata.energy[2]


# ---------------------------------------------------------------------------------------------- TTN (row a12, oracle only)
def test_ttn_default_graph_and_sweeppath_follow_the_reference_docstring():
    """default_graph_sitenodes (src/ttn/ttn_generators.jl:50-94): the documented site -> node map for N = 32, a tree
    for sizes that are not a power of two, and a sweep path that visits every node once."""
    from oracle import ttn as ot
    graph, sitenodes = ot.default_graph_sitenodes(32)
    assert sitenodes[1] == (1, 1) and sitenodes[2] == (1, 1) and sitenodes[3] == (1, 2) and sitenodes[4] == (1, 2)
    assert sitenodes[31] == (1, 16) and sitenodes[32] == (1, 16)
    assert len(graph.nodes) == 16 + 8 + 4 + 2 and graph.isneighbor((4, 1), (4, 2))
    for N in (6, 8, 12, 20):
        g, sn = ot.default_graph_sitenodes(N)
        assert sorted(sn) == list(range(1, N + 1))
        edges = sum(len(v) for v in g.adj.values()) // 2
        assert edges == len(g.nodes) - 1                              # a tree
        psi = ot.default_random_ttn(ot.dense_siteinds(N), 3, np.random.default_rng(N))
        path = ot.default_sweeppath(psi)
        assert sorted(path) == sorted(g.nodes) and len(set(path)) == len(path)
        assert abs(np.linalg.norm(ot.ttn_to_dense(psi)) - 1.0) < 1e-12   # isometrised and normalised


def test_ttn_environments_reproduce_the_dense_expectation_value():
    """LinkTensorsTTN (src/ttn/linktensors.jl:63-262): <phi|H_eff|phi> at every node of the sweep path equals the
    dense <psi|H|psi>, with the environments moved along the tree by `position!`."""
    from oracle import blocksparse as ob, couplingmodel as oc, ttn as ot
    N = 8
    sites = ot.dense_siteinds(N)
    M = oc.tfi_coupling_model(sites, h=0.7)
    Hd = oc.coupling_model_to_dense(M)
    psi = ot.default_random_ttn(sites, 4, np.random.default_rng(2))
    env = ot.StateEnvsTTN(psi, M)
    for node in ot.default_sweeppath(psi):
        env.position(node, cutoff=-1.0)
        phi = env.psi[node]
        v = ot.ttn_to_dense(env.psi)
        e = ob.inner(phi, env.product(phi).permute(phi.inds)) / ob.inner(phi, phi)
        assert abs(e - v @ Hd @ v / (v @ v)) < 1e-12


@pytest.mark.parametrize("N", [8, 12])
def test_ttn_optimize_reaches_the_tfi_ground_state(N):
    """optimize! (src/ttn/optimize_ttn.jl:148-218) with subspace expansion on the critical transverse-field Ising
    chain (BASELINE.json configs[4] at CPU scale) against exact diagonalisation."""
    from oracle import couplingmodel as oc, ttn as ot
    sites = ot.dense_siteinds(N)
    M = oc.tfi_coupling_model(sites, h=1.0)
    E0 = np.linalg.eigvalsh(oc.coupling_model_to_dense(M))[0]
    rng = np.random.default_rng(1)
    psi0 = ot.default_random_ttn(sites, 4, rng)
    prm = ot.OptimizeParamsTTN(maxdim=[8, 16], nsweeps=[4, 3], cutoff=1e-14, noise=[1e-2, 0.0], noisedecay=5,
                               disable_noise_after=3)
    E, psi, sw = ot.optimize(psi0, M, prm, ot.default_sweeppath(psi0), rng=rng)
    assert abs(E - E0) < 1e-7 and sw.maxchi[-1] <= 16
    assert all(b <= a + 1e-9 for a, b in zip(sw.energy[3:], sw.energy[4:]))      # monotone once the noise is off
    v = ot.ttn_to_dense(psi)
    Hd = oc.coupling_model_to_dense(M)
    assert abs(v @ Hd @ v / (v @ v) - E) < 1e-10


def test_ttn_with_quantum_numbers_heisenberg_ground_state():
    """The model of the reference's TTN test (test/test_TTN.jl:7-38: S=1/2 Heisenberg CouplingModel, QN-conserving
    sites, default_randomTTN in the Sz = 0 sector, optimize! with noise): the random tree state lives in the
    right charge sector, and the optimisation reaches the ED energy up to the bond-dimension truncation."""
    from oracle import couplingmodel as oc, models as om, ttn as ot
    from tests import ed
    N = 8
    sites = om.siteinds("S=1/2", N)
    M = oc.heisenberg_coupling_model(sites, merge=True)
    rng = np.random.default_rng(3)
    psi0 = ot.default_random_ttn(sites, 6, rng)
    v = ot.ttn_to_dense(psi0)
    nz = np.nonzero(np.abs(v) > 1e-14)[0]
    assert abs(np.linalg.norm(v) - 1.0) < 1e-12 and all(bin(int(i)).count("1") == N // 2 for i in nz)
    assert any(ix.tags == "QN" for ix in psi0[psi0.orthocenter].inds)
    prm = ot.OptimizeParamsTTN(maxdim=[12, 24], nsweeps=[4, 3], cutoff=1e-14, noise=[1e-2, 0.0], noisedecay=3,
                               disable_noise_after=3)
    E, psi, sw = ot.optimize(psi0, M, prm, ot.default_sweeppath(psi0), rng=rng)
    E0 = ed.lowest_energies(N, 1, 1)[0]
    assert 0 <= E - E0 < 1e-5
    w = ot.ttn_to_dense(psi)
    assert all(bin(int(i)).count("1") == N // 2 for i in np.nonzero(np.abs(w) > 1e-12)[0])     # the sector is conserved


# ---------------------------------------------------------------------------------------------- GSE (section 8f rank 1, oracle only)
def test_apply_mpo_and_global_subspace_expansion():
    """apply(H, psi) restatement is exact without truncation and near-optimal with it; krylov_extend! leaves the state
    untouched while enlarging the bonds with right-orthonormal bases (src/mps/sweep.jl:399-555)."""
    from oracle import dmrg as od, gse, models as om
    N = 8
    sites = om.siteinds("S=1/2", N)
    H = om.heisenberg_mpo(sites)
    Hd = om.mpo_to_dense(H)
    qn, dm = om.gaussian_link_sectors(12, 1.3, 4, step=1)
    mps = od.MPS(om.random_mps(sites, qn, dm, np.random.default_rng(2)))
    od.orthogonalize(mps, 1)
    v = om.mps_to_dense(mps.t)
    w = om.mps_to_dense(gse.apply_mpo(H, mps, maxdim=None, cutoff=1e-15).t)
    assert np.abs(w - Hd @ v).max() < 1e-12 * np.abs(Hd @ v).max()
    phi = gse.apply_mpo(H, mps, maxdim=8, cutoff=1e-15)
    assert max(A.inds[2].dim for A in phi.t[:-1]) == 8
    w2 = om.mps_to_dense(phi.t)
    assert abs(np.vdot(w2, Hd @ v)) / np.linalg.norm(w2) / np.linalg.norm(Hd @ v) > 0.99
    psi = od.MPS(om.neel_mps(sites))
    v0 = om.mps_to_dense(psi.t)
    gse.krylov_extend_mps(psi, H, extension_krylovdim=3)
    v1 = om.mps_to_dense(psi.t)
    assert abs(abs(np.vdot(v0, v1)) - 1.0) < 1e-13 and abs(np.linalg.norm(v1) - 1.0) < 1e-13
    assert max(A.inds[2].dim for A in psi.t[:-1]) > 1 and psi.isortho() and psi.orthocenter() == 1
    for j in range(2, N + 1):
        M = psi[j].to_dense().reshape(psi[j].inds[0].dim, -1)
        assert np.abs(M @ M.conj().T - np.eye(M.shape[0])).max() < 1e-12


@pytest.mark.parametrize("ts", [-0.02, -0.02j])
def test_dynamic_tdvp_on_a_single_mpo_follows_exact_evolution(ts):
    """The reference's TDVP test flow (test/test_MPS_TDVP.jl:50-66: tdvpsweep!(engine, -0.01im; nsite = "dynamic",
    maxdim = 20, cutoff = 1E-12, extendat = 5) on an MPO): GSE + one-site sweep at sweeps 1, 5, 10, dynamic
    one-/two-site sweeps in between, against exp(tH)|Neel>."""
    import scipy.linalg as sl
    from oracle import dmrg as od, models as om
    N = 8
    sites = om.siteinds("S=1/2", N)
    H = om.heisenberg_mpo(sites)
    Hd = om.mpo_to_dense(H)
    psi0 = od.MPS(om.neel_mps(sites))
    v0 = om.mps_to_dense(psi0.t)
    eng = od.TDVPEngine(psi0, H)
    for _ in range(10):
        od.tdvpsweep(eng, ts, "dynamic", maxdim=20, cutoff=1e-12, extendat=5)
    v = sl.expm(10 * ts * Hd) @ v0
    v /= np.linalg.norm(v)
    w = om.mps_to_dense(eng.sysenv.psi.t)
    assert abs(abs(np.vdot(v, w)) - 1.0) < 1e-9
    assert abs(eng.swdata.energy[-1] - np.real(np.vdot(v, Hd @ v))) < 1e-5


def test_ttn_excited_state_with_projector_penalty():
    """StateEnvsTTN(psi, H, Ms; weight) (src/ttn/linkproj.jl, src/ttn/environment.jl:95-101; the second half of
    test/test_TTN.jl): first excited state of the TFI chain (no QNs) and of the Heisenberg chain in the Sz = 0 sector
    (QN tree) against exact diagonalisation."""
    from oracle import couplingmodel as oc, models as om, ttn as ot
    from tests import ed
    N = 8
    sites = ot.dense_siteinds(N)
    M = oc.tfi_coupling_model(sites, h=1.5)
    w = np.linalg.eigvalsh(oc.coupling_model_to_dense(M))
    rng = np.random.default_rng(1)
    psi0 = ot.default_random_ttn(sites, 4, rng)
    prm = ot.OptimizeParamsTTN(maxdim=[16], nsweeps=[6], cutoff=1e-14, noise=1e-2, noisedecay=5, disable_noise_after=3)
    E0, gs, _ = ot.optimize(psi0, M, prm, ot.default_sweeppath(psi0), rng=rng)
    E1, ex, _ = ot.optimize(psi0, M, prm, ot.default_sweeppath(psi0), rng=rng, Ms=[gs], weight=10.0)
    assert abs(E0 - w[0]) < 1e-9 and abs(E1 - w[1]) < 1e-8
    assert abs(ot.ttn_to_dense(gs) @ ot.ttn_to_dense(ex)) < 1e-8
    sites = om.siteinds("S=1/2", N)
    M = oc.heisenberg_coupling_model(sites, merge=True)
    rng = np.random.default_rng(3)
    psi0 = ot.default_random_ttn(sites, 6, rng)
    prm = ot.OptimizeParamsTTN(maxdim=[16], nsweeps=[6], cutoff=1e-14, noise=1e-2, noisedecay=3, disable_noise_after=3)
    E0, gs, _ = ot.optimize(psi0, M, prm, ot.default_sweeppath(psi0), rng=rng)
    E1, ex, _ = ot.optimize(psi0, M, prm, ot.default_sweeppath(psi0), rng=rng, Ms=[gs], weight=10.0)
    e = ed.lowest_energies(N, 1, 2)
    assert abs(E0 - e[0]) < 1e-9 and abs(E1 - e[1]) < 1e-7


def _multi_site_model(om, oc, N=8):
    sites = om.siteinds("S=1/2", N)
    os = []
    for j in range(1, N):
        os += [(1.0, ("Sz", j), ("Sz", j + 1)), (0.5, ("S+", j), ("S-", j + 1)), (0.5, ("S-", j), ("S+", j + 1))]
    os += [(0.7, ("S+", 1), ("Sz", 3), ("S-", 4)), (0.7, ("S-", 1), ("Sz", 3), ("S+", 4)),
           (-0.4, ("Sz", 2), ("Sz", 4), ("Sz", 5), ("Sz", 7)), (0.3, ("Sz", 3)),
           (0.2, ("S+", 2), ("S-", 6)), (0.2, ("S-", 2), ("S+", 6))]
    return sites, os


def test_coupling_model_with_three_and_four_site_terms():
    """Terms with more than two operators (middle tensors carry two OpLinks; merged terms are direct sums over both),
    terms that skip several sites and a one-site term: dense operator, projected expectation values and DMRG energy."""
    from oracle import blocksparse as ob, couplingmodel as oc, dmrg as od, models as om
    N = 8
    sites, os = _multi_site_model(om, oc, N)
    S = {"Sz": np.diag([0.5, -0.5]), "S+": np.array([[0, 1.0], [0, 0]]), "S-": np.array([[0, 0], [1.0, 0]])}

    def at(ops):
        mats = [np.eye(2)] * N
        for name, p in ops:
            mats[p - 1] = S[name]
        out = np.eye(1)
        for m in mats:
            out = np.kron(out, m)
        return out
    Hk = sum(t[0] * at(t[1:]) for t in os)
    qn, dm = om.gaussian_link_sectors(8, 1.3, 4, step=1)
    mps = od.MPS(om.random_mps(sites, qn, dm, np.random.default_rng(4)))
    for merge in (True, False):
        M = oc.coupling_model(os, sites, merge=merge)
        assert sorted({t.rank for terms in M.terms for t in terms.values()}) == [2, 3, 4]
        assert np.abs(oc.coupling_model_to_dense(M) - Hk).max() < 1e-14
        for nsite, pos in ((2, 3), (2, 1), (2, 7), (1, 4), (1, 8)):
            od.orthogonalize(mps, pos)
            P = oc.ProjCouplingModel(M)
            P.set_nsite(nsite); P.position(mps.t, pos)
            phi = ob.contract(mps[pos], mps[pos + 1]) if nsite == 2 else mps[pos]
            v = om.mps_to_dense(mps.t)
            assert abs(ob.inner(phi, P.product(phi)) / ob.inner(phi, phi) - v @ Hk @ v / (v @ v)) < 1e-12
    M = oc.coupling_model(os, sites, merge=True)
    E, _, _ = od.dmrg2(od.MPS(om.neel_mps(sites)), M,
                       od.DMRGParams(maxdim=[16, 32], nsweeps=[3, 3], cutoff=1e-14, noise=[1e-3, 0.0]))
    sel = [i for i in range(2 ** N) if bin(i).count("1") == N // 2]
    assert abs(E - np.linalg.eigvalsh(Hk[np.ix_(sel, sel)])[0]) < 1e-10

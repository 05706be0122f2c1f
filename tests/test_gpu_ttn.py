"""Parity tests (`-m gpu`) of the generic device tensor algebra and of the tree-tensor-network path built on it
 against the CPU oracle (oracle/blocksparse.py, oracle/ttn.py) and exact
diagonalisation.  Everything goes through the C ABI (tnl_tensor_contract / _permute / _dag / _directsum / _factorize,
tnl_sumop_*)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _imports():
    import tennetlib.jl_b200 as T
    from tennetlib.jl_b200 import itensor as it, ttn as dt, graph as dg
    from oracle import blocksparse as ob, couplingmodel as oc, models as om, ttn as ot
    return T, it, dt, dg, ob, oc, om, ot


def dense_in_order(host, inds):
    """dense array of a product HostTensor with its axes ordered like `inds` (matched by id and prime level)."""
    arr = host.to_dense()
    perm = [next(k for k, jx in enumerate(host.inds) if jx.id == ix.id and jx.plev == ix.plev) for ix in inds]
    return np.transpose(arr, perm)


def rel(a, b):
    return float(np.abs(a - b).max() / max(1e-300, np.abs(b).max()))


def _qn_index(ob, rng, dir, tags, nsect=3, lo=1, hi=6):
    qs = [(2 * (k - nsect // 2),) for k in range(nsect)]
    return ob.Index(qs, [int(rng.integers(lo, hi)) for _ in qs], dir=dir, tags=tags)


# ---------------------------------------------------------------------------------------------- generic algebra
@pytest.mark.parametrize("cplx", [False, True])
def test_contract_permute_dag_match_oracle(ctx, cplx):
    T, it, dt, dg, ob, oc, om, ot = _imports()
    rng = np.random.default_rng(5)
    i, j, k, l, m = (_qn_index(ob, rng, d, t) for d, t in [(+1, "i"), (+1, "j"), (-1, "k"), (+1, "l"), (-1, "m")])
    A = ob.BSTensor.random([i, j, k], rng, dtype=np.complex128 if cplx else np.float64)
    B = ob.BSTensor.random([k.dag(), l, m, j.dag()], rng, dtype=np.complex128 if cplx else np.float64)
    Ad, Bd = it.ITensor.from_host(ctx, A), it.ITensor.from_host(ctx, B)
    # A * B over (k, j): result (i, l, m)
    Co = ob.contract(A, B)
    Cd = it.contract(Ad, Bd)
    assert [ix.id for ix in Cd.inds] == [ix.id for ix in Co.inds]
    assert rel(Cd.to_host().to_dense(), Co.to_dense()) < 1e-13
    # a contraction that needs both operands permuted: B(k, l, m, j) * A(i, j, k) -> (l, m, i)
    Co2 = ob.contract(B, A)
    Cd2 = it.contract(Bd, Ad)
    assert rel(dense_in_order(Cd2.to_host(), Co2.inds), Co2.to_dense()) < 1e-13
    # dag(prime(A)) * A over nothing shared but (j): outer-like contraction with a lazily dagged operand
    Ap = A.prime(1, [i, k]).dag()
    Co3 = ob.contract(Ap, A)
    Cd3 = it.contract(Ad.prime(1, [i, k]).dag(), Ad)
    assert rel(dense_in_order(Cd3.to_host(), Co3.inds), Co3.to_dense()) < 1e-13
    # permute / materialised dag round trip
    P = Ad.permute([Ad.inds[2], Ad.inds[0], Ad.inds[1]], 2)
    assert rel(P.to_host().to_dense(), np.transpose(A.to_dense(), (2, 0, 1))) == 0.0
    D = Ad.dag().materialize()
    assert rel(D.to_host().to_dense(), np.conj(A.to_dense())) == 0.0
    assert [ix.dir for ix in D.to_host().inds] == [-ix.dir for ix in A.inds]
    # vector operations across layouts: <A|A>, A + 2 * A(permuted)
    assert abs(Ad.inner(P) - np.vdot(A.to_dense(), A.to_dense())) < 1e-12 * abs(np.vdot(A.to_dense(), A.to_dense()))
    S = Ad.add(P, 2.0)
    assert rel(S.to_host().to_dense(), 3.0 * A.to_dense()) < 1e-15


def test_directsum_matches_oracle(ctx):
    T, it, dt, dg, ob, oc, om, ot = _imports()
    rng = np.random.default_rng(6)
    i, j, k = _qn_index(ob, rng, +1, "i"), _qn_index(ob, rng, +1, "j"), _qn_index(ob, rng, -1, "Link,k")
    kp = ob.Index([(0,), (2,), (4,)], [2, 1, 3], dir=-1, tags="pad")
    A = ob.BSTensor.random([k, i, j], rng)
    P = ob.BSTensor.random([kp, i, j], rng)
    So, newo = ot._directsum(A, k, P, kp, "Link,k")
    Sd, newd = it.directsum(it.ITensor.from_host(ctx, A), k, it.ITensor.from_host(ctx, P), kp, tags="Link,k")
    assert tuple(newd.qns) == tuple(newo.qns) and tuple(newd.dims) == tuple(newo.dims) and newd.dir == newo.dir
    assert rel(Sd.to_host().to_dense(), So.to_dense()) == 0.0


@pytest.mark.parametrize("which", ["svd", "qr", "svd_trunc"])
def test_factorize_matches_oracle(ctx, which):
    T, it, dt, dg, ob, oc, om, ot = _imports()
    rng = np.random.default_rng(7)
    i, j, k, l = (_qn_index(ob, rng, d, t, hi=9) for d, t in [(+1, "i"), (+1, "j"), (-1, "k"), (-1, "l")])
    A = ob.BSTensor.random([i, k, j, l], rng)
    Ad = it.ITensor.from_host(ctx, A)
    kw = dict(maxdim=7, cutoff=1e-3) if which == "svd_trunc" else dict(maxdim=None, cutoff=None)
    Lo, Ro, spo, _ = ob.factorize(A, [i, j], ortho="left", which_decomp="svd", mindim=1, tags="Link,x", **kw)
    Ld, Rd, spd, link = it.factorize(Ad, [i, j], ortho="left", which_decomp="qr" if which == "qr" else "svd", tags="Link,x",
                                     **kw)
    full = it.contract(Ld, Rd)
    ref = ob.contract(Lo, Ro)
    assert rel(dense_in_order(full.to_host(), ref.inds), ref.to_dense()) < 1e-12
    # L is an isometry: dag(L') * L = 1 on the new link
    G = it.contract(Ld.prime(1, [link]).dag(), Ld).to_host().to_dense()
    assert np.abs(G - np.eye(G.shape[0])).max() < 1e-12
    if which != "qr":
        assert len(spd.eigs) == len(spo.eigs) and rel(spd.eigs, np.asarray(spo.eigs)) < 1e-12
        assert abs(spd.truncerr - spo.truncerr) < 1e-13
        assert sorted(zip(link.qns, link.dims)) == sorted(zip(Lo.inds[-1].qns, Lo.inds[-1].dims))


# ---------------------------------------------------------------------------------------------------------- TTN
def _device_ttn(ctx, dt, dg, psi_o):
    graph = dg.Graph((a, b) for a in psi_o.graph.adj for b in psi_o.graph.adj[a])
    return dt.TTN.from_host(ctx, psi_o.sites, graph, psi_o.tensors, psi_o.orthocenter)


def _ttn_dense(ot, ob, psi_d):
    """dense state vector of a device TTN through the oracle's helper"""
    host = psi_d.to_host()
    tens = {n: ob.BSTensor([ob.Index(ix.qns, ix.dims, dir=ix.dir, tags=ix.tags, plev=ix.plev, id=ix.id) for ix in t.inds],
                           {c: np.array(b) for c, b in t.blocks.items()}) for n, t in host.items()}
    sites = [ob.Index(s.qns, s.dims, dir=s.dir, tags=s.tags, plev=s.plev, id=s.id) for s in psi_d.sites]
    og = ot.Graph()
    for a, b in psi_d.graph.edges():
        og.addedge(a, b)
    return ot.ttn_to_dense(ot.TTN(sites, og, tens, psi_d.orthocenter))


@pytest.mark.parametrize("model", ["tfi", "heisenberg_qn"])
def test_ttn_environments_and_product_match_oracle(ctx, model):
    """LinkTensorsTTN(psi, M), position! and product at every node of the sweep path: <phi|H_eff|phi> equals the dense
    <psi|H|psi> (gauge independent), and with the oracle's gauge imported H_eff phi matches element by element."""
    T, it, dt, dg, ob, oc, om, ot = _imports()
    N = 8
    if model == "tfi":
        sites = ot.dense_siteinds(N)
        M = oc.tfi_coupling_model(sites, h=0.7)
        psi = ot.default_random_ttn(sites, 4, np.random.default_rng(2))
    else:
        sites = om.siteinds("S=1/2", N)
        M = oc.heisenberg_coupling_model(sites, merge=True)
        psi = ot.default_random_ttn(sites, 6, np.random.default_rng(3))
    Hd = oc.coupling_model_to_dense(M)
    env_o = ot.StateEnvsTTN(psi, M)
    env_d = dt.StateEnvsTTN(_device_ttn(ctx, dt, dg, psi), M)
    for node in ot.default_sweeppath(psi):
        env_o.position(node, cutoff=-1.0)
        env_d.position(node, cutoff=-1.0)
        phi = env_d.psi[node]
        e_d = phi.inner(env_d.product(phi)) / phi.inner(phi)
        v = ot.ttn_to_dense(env_o.psi)
        assert abs(e_d - v @ Hd @ v / (v @ v)) < 1e-12 * max(1.0, abs(e_d))
        w = _ttn_dense(ot, ob, env_d.psi)
        assert abs(abs(v @ w) - 1.0) < 1e-12                      # the state itself is unchanged by the gauge moves
        # same gauge on both sides: rebuild the device environments from the oracle's tensors
        fresh = dt.StateEnvsTTN(_device_ttn(ctx, dt, dg, env_o.psi), M)
        phi_o = env_o.psi[node]
        Hv_o = env_o.product(phi_o).permute(phi_o.inds)
        Hv_d = fresh.product(fresh.psi[node])
        assert rel(dense_in_order(Hv_d.to_host(), phi_o.inds), Hv_o.to_dense()) < 1e-12
        # the device eig_solver and the oracle's Lanczos agree on the node problem (same operator count)
        e1, _ = dt.eig_solver(fresh, fresh.psi[node])
        e2, _ = ot.eig_solver(env_o, phi_o, None)
        assert abs(e1 - e2) < 1e-10 * max(1.0, abs(e2))


@pytest.mark.parametrize("model", ["tfi", "heisenberg_qn"])
def test_ttn_noise_free_sweeps_match_oracle(ctx, model):
    """fullsweep! without noise is deterministic: energies after every sweep agree with the oracle to 1e-10 and the bond
    dimensions are identical."""
    T, it, dt, dg, ob, oc, om, ot = _imports()
    N = 8
    if model == "tfi":
        sites = ot.dense_siteinds(N)
        M = oc.tfi_coupling_model(sites, h=1.0)
        psi = ot.default_random_ttn(sites, 4, np.random.default_rng(11))
    else:
        sites = om.siteinds("S=1/2", N)
        M = oc.heisenberg_coupling_model(sites, merge=True)
        psi = ot.default_random_ttn(sites, 6, np.random.default_rng(12))
    path = ot.default_sweeppath(psi)
    prm_o = ot.OptimizeParamsTTN(maxdim=[16], nsweeps=[3], cutoff=1e-14, noise=0.0)
    E_o, psi_o, sw_o = ot.optimize(psi, M, prm_o, path)
    prm_d = dt.OptimizeParamsTTN(maxdim=[16], nsweeps=[3], cutoff=1e-14, noise=0.0)
    psi_d0 = _device_ttn(ctx, dt, dg, psi)
    assert dt.default_sweeppath(psi_d0) == path
    E_d, psi_d, sw_d = dt.optimize(psi_d0, M, prm_d, path)
    assert sw_d.maxchi == sw_o.maxchi
    assert np.abs(np.array(sw_d.energy) - np.array(sw_o.energy)).max() < 1e-10 * abs(E_o)
    v, w = ot.ttn_to_dense(psi_o), _ttn_dense(ot, ob, psi_d)
    assert abs(abs(v @ w) - 1.0) < 1e-8


def test_ttn_optimize_with_subspace_expansion_reaches_ed(ctx):
    """optimize! with noise (subspace_expand!: device random fill + direct sum): TFI N = 16 against the oracle's own run
    (same schedule, different random padding) and exact diagonalisation; QN Heisenberg N = 8 against ED."""
    T, it, dt, dg, ob, oc, om, ot = _imports()
    from tests import ed
    N = 12
    sites = ot.dense_siteinds(N)
    M = oc.tfi_coupling_model(sites, h=1.0)
    E0 = np.linalg.eigvalsh(oc.coupling_model_to_dense(M))[0]
    psi = ot.default_random_ttn(sites, 4, np.random.default_rng(1))
    prm = dt.OptimizeParamsTTN(maxdim=[8, 16], nsweeps=[4, 3], cutoff=1e-14, noise=[1e-2, 0.0], noisedecay=5,
                               disable_noise_after=3)
    E, psi_d, sw = dt.optimize(_device_ttn(ctx, dt, dg, psi), M, prm, ot.default_sweeppath(psi), seed=3)
    assert abs(E - E0) < 1e-7 and sw.maxchi[-1] <= 16
    assert all(b <= a + 1e-9 for a, b in zip(sw.energy[3:], sw.energy[4:]))
    w = _ttn_dense(ot, ob, psi_d)
    Hd = oc.coupling_model_to_dense(M)
    assert abs(w @ Hd @ w / (w @ w) - E) < 1e-10
    # QN tree (the model of test/test_TTN.jl)
    N = 8
    sites = om.siteinds("S=1/2", N)
    M = oc.heisenberg_coupling_model(sites, merge=True)
    psi = ot.default_random_ttn(sites, 6, np.random.default_rng(3))
    prm = dt.OptimizeParamsTTN(maxdim=[12, 24], nsweeps=[4, 3], cutoff=1e-14, noise=[1e-2, 0.0], noisedecay=3,
                               disable_noise_after=3)
    E, psi_d, sw = dt.optimize(_device_ttn(ctx, dt, dg, psi), M, prm, ot.default_sweeppath(psi), seed=5)
    assert 0 <= E - ed.lowest_energies(N, 1, 1)[0] < 1e-5
    w = _ttn_dense(ot, ob, psi_d)
    assert all(bin(int(i)).count("1") == N // 2 for i in np.nonzero(np.abs(w) > 1e-12)[0])


def test_ttn_excited_state_with_projector_penalty(ctx):
    """StateEnvsTTN(psi, H, Ms; weight) on the device: first excited states against ED (TFI without QNs, Heisenberg in
    the Sz = 0 sector) and the projector product against the oracle's LinkProjTTN."""
    T, it, dt, dg, ob, oc, om, ot = _imports()
    from tests import ed
    N = 8
    sites = ot.dense_siteinds(N)
    M = oc.tfi_coupling_model(sites, h=1.5)
    w = np.linalg.eigvalsh(oc.coupling_model_to_dense(M))
    psi = ot.default_random_ttn(sites, 4, np.random.default_rng(1))
    path = ot.default_sweeppath(psi)
    prm = dt.OptimizeParamsTTN(maxdim=[16], nsweeps=[6], cutoff=1e-14, noise=1e-2, noisedecay=5, disable_noise_after=3)
    psi_d = _device_ttn(ctx, dt, dg, psi)
    E0, gs, _ = dt.optimize(psi_d, M, prm, path, seed=1)
    E1, ex, _ = dt.optimize(psi_d, M, prm, path, Ms=[gs], weight=10.0, seed=2)
    assert abs(E0 - w[0]) < 1e-9 and abs(E1 - w[1]) < 1e-8
    assert abs(_ttn_dense(ot, ob, gs) @ _ttn_dense(ot, ob, ex)) < 1e-8
    # product with the penalty against the oracle at one node (same gauge on both sides)
    env_o = ot.StateEnvsTTN(psi, M, [psi], weight=7.0)              # penalise the start state itself: <m|v> = 1
    env_d = dt.StateEnvsTTN(psi_d, M, [psi_d], weight=7.0)
    node = psi.orthocenter
    phi_o = env_o.psi[node]
    Hv_o = env_o.product(phi_o).permute(phi_o.inds)
    Hv_d = env_d.product(env_d.psi[node])
    assert rel(dense_in_order(Hv_d.to_host(), phi_o.inds), Hv_o.to_dense()) < 1e-12
    # QN tree
    sites = om.siteinds("S=1/2", N)
    M = oc.heisenberg_coupling_model(sites, merge=True)
    psi = ot.default_random_ttn(sites, 6, np.random.default_rng(3))
    path = ot.default_sweeppath(psi)
    prm = dt.OptimizeParamsTTN(maxdim=[16], nsweeps=[6], cutoff=1e-14, noise=1e-2, noisedecay=3, disable_noise_after=3)
    psi_d = _device_ttn(ctx, dt, dg, psi)
    E0, gs, _ = dt.optimize(psi_d, M, prm, path, seed=3)
    E1, ex, _ = dt.optimize(psi_d, M, prm, path, Ms=[gs], weight=10.0, seed=4)
    e = ed.lowest_energies(N, 1, 2)
    assert abs(E0 - e[0]) < 1e-9 and abs(E1 - e[1]) < 1e-7

"""BASELINE.json configs[2..4] at a scale the oracle and exact diagonalisation can check (`-m gpu`):
  configs[2]  J1-J2 Heisenberg model on a cylinder (long-range MPO from the automaton builder)  -> two-site DMRG
  configs[3]  Fermi-Hubbard ladder, U(1) x U(1), explicit Jordan-Wigner strings                  -> real-time TDVP (ComplexF64)
  configs[4]  transverse-field Ising chain on the default binary tree                           -> TTN optimize!
The full-size versions are bench workloads (`bench.py --workload ...`)."""
import itertools

import numpy as np
import pytest

from helpers import mpo_dense, to_oracle

pytestmark = pytest.mark.gpu


def _imports():
    import tennetlib.jl_b200 as T
    from tennetlib.jl_b200 import models as pm
    from oracle import dmrg as od, models as om
    return T, pm, od, om


def _spin_half_dense(N, bonds):
    from tests import ed
    Sz, Sp, Sm = ed.spin_ops(1)

    def at(op, j):
        out = np.eye(1)
        for k in range(N):
            out = np.kron(out, op if k == j else np.eye(2))
        return out
    H = np.zeros((2 ** N, 2 ** N))
    for (i, j, J) in bonds:
        H += J * (at(Sz, i) @ at(Sz, j) + 0.5 * (at(Sp, i) @ at(Sm, j) + at(Sm, i) @ at(Sp, j)))
    return H


def test_j1j2_cylinder_dmrg_matches_ed_and_oracle(ctx):
    T, pm, od, om = _imports()
    from tests import ed
    Lx, Ly = 4, 3
    N = Lx * Ly
    bonds = pm.j1j2_cylinder_bonds(Lx, Ly, 1.0, 0.5)
    sites = pm.siteinds("S=1/2", N)
    H = pm.heisenberg_bonds_mpo(sites, bonds)
    assert max(W.inds[3].dim for W in H[:-1]) >= 3 * Ly + 2          # long-range couplings widen the MPO
    Hd = _spin_half_dense(N, bonds)
    assert np.abs(mpo_dense(H) - Hd).max() < 1e-13
    sel = ed.sz0_sector(N, 1)
    e_ed = np.linalg.eigvalsh(Hd[np.ix_(sel, sel)])[0]
    psi0 = pm.neel_mps(sites)
    prm = dict(maxdim=[16, 64, 128], nsweeps=[2, 3, 3], cutoff=1e-14, noise=[1e-3, 1e-6, 0.0])
    e_d, env, sw = T.dmrg2(ctx, psi0, H, T.DMRGParams(**prm), outputlevel=0)
    assert abs(e_d - e_ed) < 1e-8 * abs(e_ed), (e_d, e_ed)
    e_o, psi_o, sw_o = od.dmrg2(od.MPS([to_oracle(A) for A in psi0], 0, 2), [to_oracle(W) for W in H],
                                od.DMRGParams(**prm), outputlevel=0)
    assert sw.maxchi == sw_o.maxchi
    assert abs(sw.energy[-1] - sw_o.energy[-1]) < 1e-10 * abs(e_ed)
    assert np.allclose(sw.energy, sw_o.energy, rtol=1e-6, atol=0)      # noisy sweeps: 1e-6, see DESIGN section 2


def test_hubbard_ladder_real_time_tdvp_matches_oracle_and_exact(ctx):
    import scipy.linalg as sl
    T, pm, od, om = _imports()
    N = 6
    sites = pm.electron_siteinds(N)
    H = pm.hubbard_mpo(sites, pm.ladder_bonds(N // 2, 2), t=1.0, U=4.0)
    states = [1 if j % 2 == 0 else 2 for j in range(N)]
    psi0 = pm.product_mps_q(sites, states)
    Hd = mpo_dense(H)
    H_o = [to_oracle(W) for W in H]
    psi_o = od.MPS([to_oracle(A) for A in psi0], 0, 2)
    v0 = om.mps_to_dense(psi_o.t)
    eng_o = od.TDVPEngine(psi_o, H_o)
    eng_d = T.TDVPEngine(ctx, psi0, H, llim=0, rlim=2)
    dt, nst = 0.05, 6
    for _ in range(nst):
        od.tdvpsweep(eng_o, -1j * dt, 2, maxdim=64, cutoff=1e-12)
        T.tdvpsweep(eng_d, -1j * dt, 2, maxdim=64, cutoff=1e-12, outputlevel=0)
    assert eng_d.swdata.maxchi == eng_o.swdata.maxchi
    assert np.allclose(eng_d.swdata.energy, eng_o.swdata.energy, rtol=1e-10, atol=1e-12)
    v = sl.expm(-1j * dt * nst * Hd) @ v0
    w = om.mps_to_dense([to_oracle(A) for A in eng_d.getpsi()])
    assert abs(abs(np.vdot(v, w)) - 1.0) < 1e-6                       # Trotter error of the second-order sweep
    w_o = om.mps_to_dense(eng_o.sysenv.psi.t)
    assert abs(abs(np.vdot(w_o, w)) - 1.0) < 1e-10
    # the evolved state is genuinely complex and the energy is conserved
    assert np.abs(np.imag(w / w[np.argmax(np.abs(w))])).max() > 1e-3
    assert abs(eng_d.swdata.energy[-1] - np.real(np.vdot(v0, Hd @ v0))) < 1e-6


def test_ttn_tfi_n16_optimize_matches_ed(ctx):
    """configs[4] at N = 16: TFI on the default binary tree, product-package model builder and random tree."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl
    T, pm, od, om = _imports()
    from tennetlib.jl_b200 import ttn as dt
    N, h = 16, 1.0
    X = sp.csr_matrix([[0.0, 1.0], [1.0, 0.0]]); Z = sp.csr_matrix([[1.0, 0.0], [0.0, -1.0]]); I2 = sp.identity(2, format="csr")

    def at(op, j):
        out = sp.identity(1, format="csr")
        for k in range(N):
            out = sp.kron(out, op if k == j else I2, format="csr")
        return out
    Hs = sum(-1.0 * at(Z, j) @ at(Z, j + 1) for j in range(N - 1)) + sum(-h * at(X, j) for j in range(N))
    e_ed = spl.eigsh(Hs, k=1, which="SA", tol=1e-12)[0][0]
    sites = dt.dense_siteinds(N)
    M = dt.tfi_coupling_model(sites, h=h)
    psi0 = dt.default_random_ttn(ctx, sites, 4, seed=5)
    prm = dt.OptimizeParamsTTN(maxdim=[8, 16, 24], nsweeps=[3, 3, 3], cutoff=1e-14, noise=[1e-2, 1e-4, 0.0], noisedecay=5)
    E, psi, sw = dt.optimize(psi0, M, prm, dt.default_sweeppath(psi0), seed=11)
    assert sw.maxchi[-1] <= 24
    assert 0 <= E - e_ed < 5e-7 * abs(e_ed), (E, e_ed)

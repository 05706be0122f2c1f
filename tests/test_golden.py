"""Oracle vs committed golden fixtures (tests/golden/oracle_golden.json, made by tests/golden/make_golden.py)."""
import json
import os

import numpy as np
import pytest

from tests.golden.make_golden import case_bond, case_dmrg

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "oracle_golden.json")))


@pytest.mark.parametrize("k", [0, 1])
def test_oracle_reproduces_bond_fixture(k):
    g = G["bond"][k]
    r = case_bond(g["kind"], g["N"], g["chi"], g["seed"], g["pos"])
    assert r["apply_flops_stored_blocks"] == g["apply_flops_stored_blocks"]          # integer work count: exact
    for key in ("expectation", "Hv_norm", "lanczos_energy"):
        assert abs(r[key] - g[key]) < 1e-11 * abs(g[key]), key
    assert r["lanczos_numops"] == g["lanczos_numops"]
    for a, b in zip(r["trunc"], g["trunc"]):
        assert a["link_qns"] == b["link_qns"] and a["link_dims"] == b["link_dims"]     # block structure: exact
        assert abs(a["truncerr"] - b["truncerr"]) < 1e-13
        assert np.abs(np.array(a["eigs"]) - np.array(b["eigs"])).max() < 1e-13


def test_oracle_reproduces_reference_dmrg_fixture():
    g = G["dmrg"][0]
    r = case_dmrg(g["kind"], g["N"], g["params"], g["name"])
    assert r["maxchi"] == g["maxchi"] and r["linkdims"] == g["linkdims"]
    # sweeps with noise>0 pass through eigenvectors of almost-null density-matrix directions (LAPACK-build
    # dependent at the 1e-9 level); the noise-free tail is reproducible to rounding
    assert np.abs(np.array(r["energy"]) - np.array(g["energy"])).max() < 1e-7
    assert abs(r["energy"][-1] - g["energy"][-1]) < 1e-11
    assert abs(r["energy"][-1] - G["ed"]["S12_N12"]) < 1e-8


# ---- second fixture file: exp_solver / TDVP, CouplingModel and MPO-sum DMRG, dynamic TDVP, TTN
G2 = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "oracle_golden_r01b.json")))


@pytest.mark.parametrize("k", range(len(G2["exp"])))
def test_oracle_reproduces_exp_solver_fixture(k):
    from tests.golden.make_golden_r01b import case_exp
    g = G2["exp"][k]
    r = case_exp(g["kind"], g["N"], g["chi"], g["seed"], g["pos"], g["nsite"], g["t"])
    assert r["converged"] == 1 and abs(r["numops"] - g["numops"]) <= 1
    assert abs(r["norm"] - g["norm"]) < 1e-10 * g["norm"]
    for key in ("overlap_with_input", "expectation"):
        assert np.abs(np.array(r[key]) - np.array(g[key])).max() < 1e-9 * max(1.0, np.abs(np.array(g[key])).max()), key


@pytest.mark.parametrize("k", range(len(G2["tdvp"])))
def test_oracle_reproduces_tdvp_fixture(k):
    from tests.golden.make_golden_r01b import case_tdvp
    g = G2["tdvp"][k]
    r = case_tdvp(g["kind"], g["N"], g["dt"], g["sched"], g["maxdim"], g["cutoff"], g["model"], **g["model_kw"])
    assert r["maxchi"] == g["maxchi"] and abs(r["abstime"] - g["abstime"]) < 1e-14
    assert np.abs(np.array(r["energy"]) - np.array(g["energy"])).max() < 1e-9
    assert np.abs(np.array(r["maxtruncerr"]) - np.array(g["maxtruncerr"])).max() < 1e-11
    if g["dt"][0] == 0.0:                                   # real time: unitary, the energy is conserved
        assert max(abs(e - g["energy"][0]) for e in g["energy"]) < 1e-8
    else:                                                   # imaginary time: monotone decrease
        assert all(b <= a + 1e-12 for a, b in zip(g["energy"], g["energy"][1:]))


@pytest.mark.parametrize("k", range(len(G2["dmrg_models"])))
def test_oracle_reproduces_model_dmrg_fixture(k):
    from tests.golden.make_golden_r01b import case_dmrg_model
    g = G2["dmrg_models"][k]
    r = case_dmrg_model(g["kind"], g["N"], g["params"], g["model"], **g["model_kw"])
    assert r["maxchi"] == g["maxchi"]
    assert np.abs(np.array(r["energy"]) - np.array(g["energy"])).max() < 1e-7       # noisy sweeps: see above
    assert abs(r["energy"][-1] - g["energy"][-1]) < 1e-10
    ed, tol = {"S=1/2": (-3.374932598687897, 1e-8), "S=1": (G["ed"]["S1_N8"], 1e-4)}[g["kind"]]   # S=1: chi = 20 truncates
    assert 0 <= g["energy"][-1] - ed < tol


def test_oracle_reproduces_ttn_fixture():
    from tests.golden.make_golden_r01b import case_ttn
    g = G2["ttn"][0]
    r = case_ttn(g["N"], g["h"], g["chi0"], g["seed"], g["params"])
    assert r["maxchi"] == g["maxchi"]
    assert abs(r["energy"][-1] - g["energy"][-1]) < 1e-9 and abs(g["energy"][-1] - g["ed"]) < 1e-7


def test_oracle_reproduces_baseline_config0_fixture():
    """BASELINE.json configs[0] (S=1/2 N=20, maxdim 20 -> 64): the reference's CPU-runnable case, against the fixture
    and the literature ground-state energy."""
    from tests.golden.make_golden_r01b import case_config0
    g = G2["config0"][0]
    r = case_config0(g["params"])
    assert r["maxchi"] == g["maxchi"] and r["linkdims"] == g["linkdims"]
    assert np.abs(np.array(r["energy"]) - np.array(g["energy"])).max() < 1e-7
    assert abs(r["energy"][-1] - g["energy"][-1]) < 1e-11
    assert abs(g["energy"][-1] - g["ed_literature"]) < 1e-10

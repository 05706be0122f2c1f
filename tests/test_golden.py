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

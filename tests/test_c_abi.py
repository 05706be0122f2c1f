"""The drop-in boundary driven from plain C (tests/c_abi_smoke.c): what the Julia `ccall` shim does, with no Python
between the caller and libtnl_b200.so."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _exe():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    exe = os.path.join(ROOT, "tests", "_build", "c_abi_smoke")
    src = os.path.join(ROOT, "tests", "c_abi_smoke.c")
    lib = os.path.join(ROOT, "tennetlib.jl_b200", "libtnl_b200.so")
    if not os.path.exists(lib):
        import tennetlib.jl_b200 as T
        T.load(build_if_missing=True)
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(src), os.path.getmtime(lib)):
        ge.build_c_abi_smoke()
    return exe


def test_c_program_compiles_and_links_against_the_header():
    """gcc -Wall on the C caller: every entry point it uses is declared in include/tnl_b200.h with a C-compatible
    signature and exported by the shared library (link step); no compute without a GPU."""
    exe = _exe()
    assert os.access(exe, os.X_OK)
    out = subprocess.run(["nm", "-D", "--undefined-only", exe], capture_output=True, text=True).stdout
    for sym in ("tnl_ctx_create", "tnl_env_set_site_op", "tnl_tensor_import", "tnl_env_make_phi", "tnl_env_position",
                "tnl_eigsolve_lanczos", "tnl_replacebond", "tnl_tensor_export"):
        assert sym in out, sym


@pytest.mark.gpu
def test_c_program_runs_dmrg_through_the_c_abi():
    r = subprocess.run([_exe()], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "C ABI SMOKE OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]

"""Host-side logic of the product package (no GPU needed): parameter handling, sweep bookkeeping,
library loading and the exported C ABI."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import tennetlib.jl_b200 as T
from tennetlib.jl_b200 import models as pm
from tennetlib.jl_b200.update_site import halfsweep_done

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_dmrgparams_broadcast_and_validation():
    p = T.DMRGParams(maxdim=[20, 50], nsweeps=[10, 10], cutoff=1e-14, noise=1e-3, noisedecay=2, disable_noise_after=5)
    assert p.cutoff == [1e-14, 1e-14] and p.noise == [1e-3, 1e-3] and p.disable_noise_after == [5, 5]
    with pytest.raises(ValueError):
        T.DMRGParams(maxdim=[20], nsweeps=[1, 2])
    with pytest.raises(ValueError):
        T.DMRGParams(maxdim=[20, 30], nsweeps=[1, 2], cutoff=[1e-10])


def test_halfsweep_done_matches_reference_table():
    # src/mps/update_site.jl:13-23
    assert halfsweep_done(10, 1, 2, "right") and halfsweep_done(10, 1, 1, "right")
    assert halfsweep_done(10, 9, 2, "left") and not halfsweep_done(10, 9, 1, "left")
    assert halfsweep_done(10, 10, 1, "left") and not halfsweep_done(10, 5, 2, "left")


def test_update_position_rejects_noise_with_reverse_step():
    with pytest.raises(RuntimeError):
        T.update_position(None, T.eig_solver, 1, 2, "left", noise=1e-3, time_step=-0.1j)
    with pytest.raises(NotImplementedError):
        T.update_position(None, T.eig_solver, 1, 3, "left")


def test_model_builders_are_consistent_with_the_oracle():
    from oracle import models as om
    for kind, N in (("S=1/2", 5), ("S=1", 4)):
        Hp = pm.heisenberg_mpo(pm.siteinds(kind, N))
        Ho = om.heisenberg_mpo(om.siteinds(kind, N))
        for a, b in zip(Hp, Ho):
            assert np.abs(a.to_dense() - b.to_dense()).max() == 0
            assert [ix.dims for ix in a.inds] == [ix.dims for ix in b.inds]
            assert [ix.dir for ix in a.inds] == [ix.dir for ix in b.inds]
    q, d = pm.gaussian_link_sectors(4096, 1.3, 6)
    assert sum(d) == 4096 and d[6] == max(d) and len(q) == 13
    links = pm.random_mps_links(pm.siteinds("S=1", 20), q, d)
    assert links[0].dim == 1 and links[-1].dim == 1 and links[1].dim == 3 and links[10].dim == 4096


def test_library_builds_loads_and_exports_every_declared_symbol():
    lib = T.load()
    hdr = open(os.path.join(ROOT, "include", "tnl_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(tnl_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 30
    nm = subprocess.run(["nm", "-D", "--defined-only", T.so_path()], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (tnl_[a-z0-9_]+)", nm))
    missing = [s for s in declared if s not in exported]
    assert not missing, f"declared in include/tnl_b200.h but not exported: {missing}"
    assert sorted(T.EXPORTED) == declared, "ctypes binding and header disagree"
    for name in declared:
        assert hasattr(lib, name)


def test_sass_contains_fp64_tensor_ops_and_async_copies():
    """The grouped GEMM must be the DMMA path (FP64 tensor op) fed by LDGSTS, not a scalar DFMA loop."""
    out = subprocess.run(["cuobjdump", "-sass", T.so_path()], capture_output=True, text=True).stdout
    assert "DMMA" in out and "LDGSTS" in out


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(T.TnlError):
        T.Context(0)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "tennetlib.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".hpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_shard_ranges_tile_every_sector_exactly():
    """Multi-GPU partition of the right link (tnl_shard_range): for every (dim, world) the per-rank ranges are
    contiguous, disjoint and cover [0, dim); remainders rotate with the sector number so small sectors spread."""
    import ctypes as C
    lib = T.load()
    for world in (1, 2, 3, 4, 8):
        for dim in (0, 1, 2, 7, 8, 11, 88, 385, 935, 1254):
            for sector in range(5):
                seen = []
                for rank in range(world):
                    st, cnt = C.c_int32(), C.c_int32()
                    assert lib.tnl_shard_range(dim, world, sector, rank, C.byref(st), C.byref(cnt)) == 0
                    assert abs(cnt.value - dim / world) < 1
                    seen.append((st.value, cnt.value))
                seen.sort()
                pos = 0
                for st, cnt in seen:
                    assert st == pos
                    pos += cnt
                assert pos == dim
        # dim-1 sectors land on different ranks for different sector numbers
        owners = set()
        for sector in range(world):
            for rank in range(world):
                st, cnt = C.c_int32(), C.c_int32()
                lib.tnl_shard_range(1, world, sector, rank, C.byref(st), C.byref(cnt))
                if cnt.value:
                    owners.add(rank)
        assert len(owners) == world

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


def test_sass_contains_tma_loads():
    """The 128x128 tiles of the grouped GEMM are staged by the TMA unit (UTMALDG) with mbarrier completion."""
    out = subprocess.run(["cuobjdump", "-sass", T.so_path()], capture_output=True, text=True).stdout
    assert "UTMALDG.2D" in out and "SYNCS.ARRIVE.TRANS64" in out


def test_tma_swizzled_fragment_loads_are_bank_conflict_free():
    """csrc/kernels.cu gemm_tma_ws_kernel: with the k permutation  k = (0,3,12,15)[lc] ^ 2s  the 8x4 DMMA fragment
    loads of every half-warp hit 16 distinct 8-byte bank pairs in both SWIZZLE_128B tile layouts, and the four
    k-steps cover k = 0..15 exactly once."""
    ksets = [[((lc & 1) * 3 + (lc >> 1) * 12) ^ (2 * s) for lc in range(4)] for s in range(4)]
    assert sorted(k for ks in ksets for k in ks) == list(range(16))

    def pair_kfast(x, k):      # one box {16 k, 128 x}: row x = 128 bytes
        return ((x * 128 + (((k >> 1) ^ (x & 7)) << 4) + (k & 1) * 8) // 8) % 16

    def pair_xfast(x, k):      # boxes {16 x, 16 k}: box x>>4, row k = 128 bytes
        xx = x & 15
        return (((x >> 4) * 2048 + k * 128 + (((xx >> 1) ^ (k & 7)) << 4) + (xx & 1) * 8) // 8) % 16

    for f in (pair_kfast, pair_xfast):
        for ks in ksets:
            for w0 in range(0, 128, 8):
                for half in (0, 1):
                    hit = {f(w0 + lr, ks[lc]) for lr in range(4 * half, 4 * half + 4) for lc in range(4)}
                    assert len(hit) == 16

    # the kernel's closed-form fragment offsets equal the layout functions above
    def off_kfast(x, k):
        return x * 16 + (((k >> 1) ^ (x & 7)) << 1) + (k & 1)

    def off_xfast(x, k):
        xx = x & 15
        return (x >> 4) * 256 + k * 16 + (((xx >> 1) ^ (k & 7)) << 1) + (xx & 1)

    for w0 in (0, 32, 64, 96):
        for lr in range(8):
            for lc in range(4):
                kb = (lc & 1) * 3 + (lc >> 1) * 12
                d = -8 if (lc >> 1) else 8
                for s in range(4):
                    k = kb ^ (2 * s)
                    e_kf = (w0 + lr) * 16 + (((k >> 1) ^ lr) << 1) + (k & 1)
                    e_xf = (w0 >> 4) * 256 + k * 16 + 2 * ((lr >> 1) ^ (k & 3)) + (lr & 1) + 8 * ((k >> 2) & 1)
                    for i in range(4):
                        assert e_kf + i * 128 == off_kfast(w0 + 8 * i + lr, k)
                        assert e_xf + (i >> 1) * 256 + (i & 1) * (d if s < 2 else -d) == off_xfast(w0 + 8 * i + lr, k)


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


def test_coupling_model_ingestion_gives_canonical_rank4_terms():
    """`canonical_terms` (host side of tnl_env_cm_set_term): every term tensor becomes W(wl, s', s, wr) with the
    OpLink shared with the previous / next tensor of the same id, dim-1 charge-0 placeholders elsewhere, and the
    dense operator of the canonical terms equals the model's."""
    from oracle import couplingmodel as oc, models as om
    from tennetlib.jl_b200.couplingmodel import canonical_terms, is_coupling_model
    sites = om.siteinds("S=1", 6)
    M = oc.heisenberg_coupling_model(sites, merge=True, field=0.3, j2=0.5)
    assert is_coupling_model(M) and not is_coupling_model(om.heisenberg_mpo(sites))
    ct = canonical_terms(M)
    assert len(ct) == 6
    support = {}
    for j, terms in enumerate(ct):
        assert set(terms) == set(M.terms[j])
        for tid, W in terms.items():
            assert len(W.inds) == 4 and W.inds[1].plev == 1 and W.inds[2].plev == 0
            assert W.inds[0].dir == +1 and W.inds[3].dir == -1
            for c, b in W.blocks.items():
                assert b.shape == tuple(ix.dims[k] for ix, k in zip(W.inds, c))
                q = sum(ix.dir * ix.qns[k][0] for ix, k in zip(W.inds, c))
                assert q == 0                                   # flux 0 with the OpLinks carrying the charge
            support.setdefault(tid, []).append((j, W))
    # chain structure per id: first tensor has no left link, last no right link, inner links pair up
    dense = np.zeros((3 ** 6, 3 ** 6))
    for tid, lst in support.items():
        assert not lst[0][1].has_wl and not lst[-1][1].has_wr
        acc = np.ones((1, 1, 1))
        pos = 0
        for j, W in lst:
            while pos < j:                                       # sites the term skips: identity
                acc = np.einsum("rcw,xy->rxcyw", acc, np.eye(3)).reshape(acc.shape[0] * 3, acc.shape[1] * 3, acc.shape[2])
                pos += 1
            Wd = W.to_dense()
            assert Wd.shape[0] == acc.shape[2]
            acc = np.einsum("rcw,wxyv->rxcyv", acc, Wd).reshape(acc.shape[0] * 3, acc.shape[1] * 3, Wd.shape[3])
            pos += 1
        while pos < 6:
            acc = np.einsum("rcw,xy->rxcyw", acc, np.eye(3)).reshape(acc.shape[0] * 3, acc.shape[1] * 3, acc.shape[2])
            pos += 1
        assert acc.shape[2] == 1
        dense += acc[:, :, 0]
    assert np.abs(dense - oc.coupling_model_to_dense(M)).max() < 1e-13


def test_complex_host_tensors_flatten_to_interleaved_storage():
    from tennetlib.jl_b200.tensor import HostTensor, Index, flatten_blocks
    ix = [Index([(0,), (2,)], [2, 1], dir=+1), Index([(0,), (2,)], [2, 1], dir=-1)]
    rng = np.random.default_rng(0)
    t = HostTensor(ix, {(0, 0): rng.standard_normal((2, 2)) + 1j * rng.standard_normal((2, 2)), (1, 1): np.array([[0.5 - 2j]])})
    assert t.is_complex() and t.to_dense().dtype == np.complex128
    coords, offsets, data, nb = flatten_blocks(t, np.complex128)
    assert nb == 2 and list(offsets) == [0, 4] and data.dtype == np.complex128
    raw = data.view(np.float64)                                  # what tnl_tensor_import_c128 reads: (re, im) pairs
    assert raw[0] == t.blocks[(0, 0)][0, 0].real and raw[1] == t.blocks[(0, 0)][0, 0].imag
    assert raw[2] == t.blocks[(0, 0)][1, 0].real                  # column-major inside the block
    assert raw[8] == 0.5 and raw[9] == -2.0


class _FakeEnv:
    """Stand-in for StateEnvs: records the (site, nsite, ortho) sequence a sweep driver issues."""

    def __init__(self, N, lasteig, dims, kind="cm"):
        self.N, self.calls, self._dims, self._lasteig = N, [], dims, lasteig
        self.nterms, self.is_coupling_model, self.has_penalty = 1, kind == "cm", False
        self.llim, self.rlim = 0, 2

    def __len__(self):
        return self.N

    def isortho(self):
        return True

    def orthocenter(self):
        return 1

    def linkdim(self, bond):
        return self._dims[bond - 1]

    def linkdims(self):
        return list(self._dims)


def test_dynamic_fullsweep_picks_one_or_two_site_updates_like_the_reference(monkeypatch):
    """src/mps/sweep.jl:302-336: per bond nsite = 1 if the last kept Schmidt weight of the previous half sweep is
    below eigthreshold or the bond is saturated at maxdim, else 2; the site index of a one-site update on the way
    back is bond+1; the edge sites get their extra one-site update."""
    from tennetlib.jl_b200 import sweep as sw
    N = 6
    dims = [2, 4, 8, 4, 2]
    env = _FakeEnv(N, None, dims)
    data = sw.SweepData()
    data.sweepcount = 1
    data.energy, data.entropy, data.maxchi, data.maxtruncerr = [-0.5], [0.1], [4], [0.0]
    data.lasteigs = [np.array([0.9, 1e-3]), np.array([0.9, 1e-14]), np.array([0.5, 0.1]), np.array([0.9, 1e-2]),
                     np.array([0.9, 1e-13])]

    def fake_update(sysenv, solver, pos, nsite, ortho, **kw):
        sysenv.calls.append((pos, nsite, ortho))
        return -1.0, 1e-9, np.array([0.7, 0.3])
    monkeypatch.setattr(sw, "update_position", fake_update)
    sw.dynamic_fullsweep(env, T.exp_solver, data, maxdim=8, outputlevel=0, time_step=-0.05)
    left = [c for c in env.calls if c[2] == "left"]
    right = [c for c in env.calls if c[2] == "right"]
    # bond 2: tiny weight -> 1 site; bond 3: saturated (dim 8 >= maxdim) -> 1 site; bond 5: tiny -> 1 site + edge update
    assert left == [(1, 2, "left"), (2, 1, "left"), (3, 1, "left"), (4, 2, "left"), (5, 1, "left"), (6, 1, "left")]
    # on the way back the decisions use the eigs recorded on the way out (0.3 everywhere): only the saturated bond is 1-site
    assert right == [(5, 2, "right"), (4, 2, "right"), (4, 1, "right"), (2, 2, "right"), (1, 2, "right")]
    assert data.sweepcount == 2 and len(data.energy) == 2 and data.maxchi == [4, 8]
    # first sweep on a single MPO starts with the Global Subspace Expansion (src/mps/sweep.jl:266-282)
    from tennetlib.jl_b200 import gse
    seen = []
    monkeypatch.setattr(gse, "krylov_extend", lambda sysenv, **kw: seen.append(sorted(kw)))
    sw.dynamic_fullsweep(_FakeEnv(N, None, dims, kind="mpo"), T.exp_solver, sw.SweepData(), maxdim=8, outputlevel=0)
    assert len(seen) == 1


def test_tdvpsweep_argument_checks():
    from tennetlib.jl_b200 import tdvp

    class E:
        sysenv, swdata, abstime = None, None, 0.0
    with pytest.raises(RuntimeError):
        tdvp.tdvpsweep(E(), -0.1, 2, solver=T.eig_solver)
    with pytest.raises(RuntimeError):
        tdvp.tdvpsweep(E(), -0.1, 3)
    with pytest.raises(RuntimeError):
        tdvp.tdvpsweep(E(), -0.1, 2, extendat=5)
    with pytest.raises(RuntimeError):
        T.exp_solver(None, None, None)


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (CPU arm of the driver: the oracle port timed on the host cores) prints one JSON line
    with the same metric / unit / config as the GPU arm plus `impl`, `cpu_baseline` and a zero-copy `e2e`."""
    import json
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--chi", "128",
                          "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-500:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "heff_apply_fp64_tflops" and line["unit"] == "TFLOP/s"
    assert line["higher_is_better"] is True and line["value"] > 0 and line["steps"] == 1 and line["warmup"] == 1
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"] and line["config"]["chi"] == 128
    # the other ranks of a torchrun launch exit silently
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--chi", "128",
                          "--steps", "1", "--warmup", "1", "--gpus", "2"], capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_graph_and_sweeppath_match_oracle():
    from tennetlib.jl_b200 import graph as dg
    from oracle import ttn as ot
    for N in (4, 6, 8, 11, 12, 16, 27, 32, 100, 128):
        g_o, sn_o = ot.default_graph_sitenodes(N)
        g_d, sn_d = dg.default_graph_sitenodes(N)
        assert sn_o == sn_d and {k: set(v) for k, v in g_o.adj.items()} == g_d.adj
        assert dg.find_eccentric_central_node(g_d) == ot.find_eccentric_central_node(g_o)
        src = sorted(g_d.nodes)[0]
        assert dg.nodes_from_bfs(g_d, src, reverse=True) == ot.nodes_from_bfs(g_o, src, reverse=True)

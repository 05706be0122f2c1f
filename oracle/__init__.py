"""CPU oracle for the DMRG/TDVP hot path of TenNetLib.jl -- TEST INFRASTRUCTURE ONLY.

This package is a NumPy restatement of the algorithm the reference executes on its
CPU path (ITensors/NDTensors block-sparse contraction, KrylovKit Lanczos, ITensors
truncation, TenNetLib sweep drivers).  It exists to *check* the CUDA product in
`tennetlib.jl_b200/`; only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
cpu_baseline / `--impl reference` legs may import it.  The product never does.

PARITY UNPINNED: the reference (Julia, un-vendored ITensors 0.9 / ITensorMPS 0.3 /
KrylovKit 0.9|0.10, see /root/reference/Project.toml:14-21) cannot run in this image
and its tests hold no assertions or golden vectors (test/test_MPS_DMRG.jl:100-146).
The oracle is therefore pinned against exact-diagonalisation energies and dense
linear-algebra identities instead (tests/test_oracle_*.py).
"""

"""GEMM self-tests sized for compute-sanitizer (TMA kernel incl. split-K, 64x64 kernel, all transposes):
   compute-sanitizer --tool racecheck|synccheck|memcheck python tools/sanitize_gemm.py"""
import sys
sys.path.insert(0, ".")
import tennetlib.jl_b200 as T
ctx = T.Context()
for (M, N, K) in ((129, 200, 64), (300, 257, 100), (256, 256, 4096), (130, 70, 33)):
    for ta in (0, 1):
        for tb in (0, 1):
            ms, err = ctx.gemm_selftest(M, N, K, ta, tb, 1, True)
            assert err < 1e-12 * K, (M, N, K, ta, tb, err)
print("selftests ok")

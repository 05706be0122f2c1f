import sys
sys.path.insert(0, ".")
import tennetlib.jl_b200 as T
ctx = T.Context()
for (M, N, K) in ((256, 65536, 256), (65536, 256, 256), (512, 262144, 512), (262144, 512, 512), (256, 256, 65536), (4096, 4096, 256)):
    for ta in (0, 1):
        for tb in (0, 1):
            ms, _ = ctx.gemm_selftest(M, N, K, ta, tb, 5, False)
            print((M, N, K, ta, tb), f"{ms:7.3f} ms  {2.0 * M * N * K / ms / 1e9:6.2f} TFLOP/s", flush=True)

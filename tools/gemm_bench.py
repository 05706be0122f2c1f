"""Micro-benchmark / correctness sweep of the grouped DGEMM kernel at the shapes of the chi=4096 apply."""
import sys
sys.path.insert(0, ".")
import tennetlib.jl_b200 as T

ctx = T.Context()
print("== correctness (odd sizes, all transposes)")
worst = 0.0
for (M, N, K) in ((1, 1, 1), (7, 5, 3), (129, 65, 17), (255, 257, 33), (385, 1254, 935), (130, 300, 1000), (128, 128, 16),
                  (129, 200, 64), (300, 257, 100), (1000, 1000, 1001), (66, 4000, 15), (2000, 70, 333),
                  (256, 256, 20000), (129, 300, 5000), (100, 100, 4097), (400, 130, 1030)):
    for ta in (0, 1):
        for tb in (0, 1):
            ms, err = ctx.gemm_selftest(M, N, K, ta, tb, 1, True)
            worst = max(worst, err / max(1, K) ** 0.5)
print("max err/sqrt(K)", worst)
print("== speed")
shapes = [(5632, 8272, 1254, 0, 0), (8272, 1254, 5632, 0, 1), (4096, 4096, 4096, 0, 0), (8192, 8192, 8192, 0, 0),
          (1254, 3762, 3762, 1, 0)]
for (M, N, K, ta, tb) in shapes:
    ms, _ = ctx.gemm_selftest(M, N, K, ta, tb, 5, False)
    print((M, N, K, ta, tb), f"{2.0 * M * N * K / ms / 1e9:6.2f} TFLOP/s")

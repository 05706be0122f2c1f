"""Micro-benchmark / correctness sweep of the grouped DGEMM kernel variants (tuning aid)."""
import sys
sys.path.insert(0, ".")
import tennetlib.jl_b200 as T

ctx = T.Context()
print("== correctness (odd sizes, all transposes)")
for v in (3, 9, 10, 12):
    worst = 0.0
    for (M, N, K) in ((1, 1, 1), (7, 5, 3), (129, 65, 17), (255, 257, 33), (385, 1254, 935), (130, 300, 1000)):
        for ta in (0, 1):
            for tb in (0, 1):
                try:
                    ms, err = ctx.gemm_selftest(M, N, K, ta, tb, v, 1, True)
                except Exception:
                    continue
                worst = max(worst, err / max(1, K) ** 0.5)
    print("variant", v, "max err/sqrt(K)", worst)
print("== speed")
shapes = [(5632, 8272, 1254, 0, 0), (8272, 1254, 5632, 0, 1), (4096, 4096, 4096, 0, 0), (8192, 8192, 8192, 0, 0),
          (1254, 3762, 3762, 1, 0)]
for (M, N, K, ta, tb) in shapes:
    row = []
    for v in (3, 9, 10, 12):
        try:
            ms, _ = ctx.gemm_selftest(M, N, K, ta, tb, v, 5, False)
            row.append(2.0 * M * N * K / ms / 1e9)
        except Exception:
            row.append(float("nan"))
    print((M, N, K, ta, tb), " ".join(f"{x:6.2f}TF" for i, x in enumerate(row)))

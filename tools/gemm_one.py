import sys
sys.path.insert(0, ".")
import tennetlib.jl_b200 as T
ctx = T.Context()
v = int(sys.argv[1]) if len(sys.argv) > 1 else 3
ms, _ = ctx.gemm_selftest(5632, 8272, 1254, 0, 0, v, 2, False)
print("variant", v, 2.0 * 5632 * 8272 * 1254 / ms / 1e9, "TF/s")

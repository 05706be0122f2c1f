"""Step 0 of the build: measure the FP64 roofline denominators on the box.

cuBLAS DGEMM 8192^3 (burst best-of-10 and a sustained 3 s loop) and the cuSOLVER
decompositions the truncation step can choose from, at the block sizes of config 2.
Writes gpurun_out/peaks_fp64.json.  Library calls only -- nothing here is product code.
"""
import json, os, time, torch

def ev_time(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) * 1e-3)
    return best

out = {"gpu": torch.cuda.get_device_name(0)}
n = 8192
A = torch.randn(n, n, dtype=torch.float64, device="cuda")
B = torch.randn(n, n, dtype=torch.float64, device="cuda")
C = torch.empty_like(A)
t = ev_time(lambda: torch.matmul(A, B, out=C), reps=10, warm=3)
out["dgemm_8192_burst_tflops"] = 2 * n ** 3 / t / 1e12
torch.cuda.synchronize(); t0 = time.time(); k = 0
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
while time.time() - t0 < 3.0:
    torch.matmul(A, B, out=C); k += 1
    if k % 8 == 0:
        torch.cuda.synchronize()
b.record(); torch.cuda.synchronize()
out["dgemm_8192_sustained_tflops"] = 2 * n ** 3 * k / (a.elapsed_time(b) * 1e-3) / 1e12
for m in (1254, 2048, 4096):
    X = torch.randn(m, m, dtype=torch.float64, device="cuda"); Y = torch.randn(m, m, dtype=torch.float64, device="cuda")
    t = ev_time(lambda: torch.matmul(X, Y), reps=10, warm=3)
    out[f"dgemm_{m}_tflops"] = 2 * m ** 3 / t / 1e12
# HBM copy / triad in fp64
N = 1 << 28
x = torch.randn(N, dtype=torch.float64, device="cuda"); y = torch.empty_like(x)
t = ev_time(lambda: y.copy_(x), reps=10, warm=3)
out["copy_f64_gbs"] = 2 * N * 8 / t / 1e9
t = ev_time(lambda: torch.dot(x, y), reps=10, warm=3)
out["dot_f64_gbs"] = 2 * N * 8 / t / 1e9
del x, y
for m in (1254, 2805, 3762):
    M = torch.randn(m, m, dtype=torch.float64, device="cuda")
    for drv in ("gesvd", "gesvdj", "gesvda"):
        try:
            t = ev_time(lambda: torch.linalg.svd(M, full_matrices=False, driver=drv), reps=1, warm=1)
            out[f"svd_{drv}_{m}_s"] = t
        except Exception as e:  # noqa
            out[f"svd_{drv}_{m}_s"] = str(e)[:80]
    G = M @ M.T
    out[f"eigh_{m}_s"] = ev_time(lambda: torch.linalg.eigh(G), reps=2, warm=1)
    out[f"qr_{m}_s"] = ev_time(lambda: torch.linalg.qr(M), reps=2, warm=1)
    print(json.dumps(out), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/peaks_fp64.json", "w"), indent=1)
print(json.dumps(out))

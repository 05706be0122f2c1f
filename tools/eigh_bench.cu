// Micro-benchmark of the vendor Hermitian eigensolvers the truncation can call (library calls only, nothing here is
// product code): legacy cusolverDnDsyevd, 64-bit cusolverDnXsyevd, cusolverDnXsyevBatched (batch of one), the
// tridiagonalisation alone (cusolverDnDsytrd), and the concurrency of several decompositions on side streams
// (one host thread per stream, the way factorize.cu's syevd_batch drives them).
// build: nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/eigh_bench.cu -o tools/_build/eigh_bench -lcusolver
// run  : tools/_build/eigh_bench > gpurun_out/eigh_bench.json
#include <cuda_runtime.h>
#include <cusolverDn.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <thread>
#include <vector>

#define CK(x) do { auto e__ = (x); if (e__ != 0) { fprintf(stderr, "error %d at %s:%d\n", (int)e__, __FILE__, __LINE__); exit(1); } } while (0)

__global__ void fill_sym(double* A, int n, unsigned seed) {
  // A = symmetric pseudo-random matrix with a Marchenko-Pastur-like spread (diag dominant enough to be SPD-ish)
  for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < (long)n * n; e += (long)gridDim.x * blockDim.x) {
    int i = (int)(e % n), j = (int)(e / n);
    int a = i < j ? i : j, b = i < j ? j : i;
    unsigned h = seed ^ (unsigned)(a * 2654435761u) ^ (unsigned)(b * 40503u + 12345u);
    h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
    double v = (double)h / 4294967296.0 - 0.5;
    A[e] = v + (i == j ? 0.05 * n : 0.0);
  }
}

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

struct Job { int n; double *A, *A0, *W, *work; size_t lwork; int* info; cudaStream_t s; cusolverDnHandle_t h; void* hwork; size_t hbytes; };

static void run_legacy(Job& j) {
  CK(cudaMemcpyAsync(j.A, j.A0, sizeof(double) * j.n * j.n, cudaMemcpyDeviceToDevice, j.s));
  CK(cusolverDnDsyevd(j.h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, j.n, j.A, j.n, j.W, j.work, (int)j.lwork, j.info));
}

int main() {
  std::vector<int> sizes = {512, 1024, 1536, 2048, 2805, 3762};
  printf("{\n");
  cusolverDnHandle_t h;
  cudaStream_t s;
  CK(cudaStreamCreate(&s));
  CK(cusolverDnCreate(&h));
  CK(cusolverDnSetStream(h, s));
  cusolverDnParams_t params;
  CK(cusolverDnCreateParams(&params));
  int* info;
  CK(cudaMalloc(&info, 64));
  for (int n : sizes) {
    double *A, *A0, *W, *tau, *D, *E;
    CK(cudaMalloc(&A, sizeof(double) * n * n));
    CK(cudaMalloc(&A0, sizeof(double) * n * n));
    CK(cudaMalloc(&W, sizeof(double) * n));
    CK(cudaMalloc(&tau, sizeof(double) * n));
    CK(cudaMalloc(&D, sizeof(double) * n));
    CK(cudaMalloc(&E, sizeof(double) * n));
    fill_sym<<<592, 256, 0, s>>>(A0, n, 17u + n);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto timeit = [&](auto&& fn) {
      float best = 1e30f;
      for (int r = 0; r < 3; r++) {
        CK(cudaMemcpyAsync(A, A0, sizeof(double) * n * n, cudaMemcpyDeviceToDevice, s));
        CK(cudaStreamSynchronize(s));
        cudaEventRecord(e0, s);
        fn();
        cudaEventRecord(e1, s);
        CK(cudaStreamSynchronize(s));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r > 0 && ms < best) best = ms;
      }
      return best;
    };
    // legacy
    int lw = 0;
    CK(cusolverDnDsyevd_bufferSize(h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, A, n, W, &lw));
    double* work; CK(cudaMalloc(&work, sizeof(double) * lw));
    float t_legacy = timeit([&] { CK(cusolverDnDsyevd(h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, A, n, W, work, lw, info)); });
    float t_novec = timeit([&] { CK(cusolverDnDsyevd(h, CUSOLVER_EIG_MODE_NOVECTOR, CUBLAS_FILL_MODE_LOWER, n, A, n, W, work, lw, info)); });
    cudaFree(work);
    // sytrd alone
    int lt = 0;
    CK(cusolverDnDsytrd_bufferSize(h, CUBLAS_FILL_MODE_LOWER, n, A, n, D, E, tau, &lt));
    CK(cudaMalloc(&work, sizeof(double) * lt));
    float t_sytrd = timeit([&] { CK(cusolverDnDsytrd(h, CUBLAS_FILL_MODE_LOWER, n, A, n, D, E, tau, work, lt, info)); });
    cudaFree(work);
    // 64-bit API
    size_t wd = 0, wh = 0;
    CK(cusolverDnXsyevd_bufferSize(h, params, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, CUDA_R_64F, A, n, CUDA_R_64F, W,
                                   CUDA_R_64F, &wd, &wh));
    void* dw; CK(cudaMalloc(&dw, wd + 16));
    std::vector<char> hw(wh + 16);
    float t_x = timeit([&] { CK(cusolverDnXsyevd(h, params, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, CUDA_R_64F, A, n, CUDA_R_64F,
                                                 W, CUDA_R_64F, dw, wd, hw.data(), wh, info)); });
    cudaFree(dw);
    // batched API with one matrix
    float t_b = -1;
    {
      size_t bd = 0, bh = 0;
      auto st = cusolverDnXsyevBatched_bufferSize(h, params, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, CUDA_R_64F, A, n,
                                                  CUDA_R_64F, W, CUDA_R_64F, &bd, &bh, 1);
      if (st == CUSOLVER_STATUS_SUCCESS) {
        void* bw; CK(cudaMalloc(&bw, bd + 16));
        std::vector<char> bhw(bh + 16);
        bool ok = true;
        t_b = timeit([&] {
          auto s2 = cusolverDnXsyevBatched(h, params, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, CUDA_R_64F, A, n, CUDA_R_64F, W,
                                           CUDA_R_64F, bw, bd, bhw.data(), bh, info, 1);
          if (s2 != CUSOLVER_STATUS_SUCCESS) ok = false;
        });
        if (!ok) t_b = -2;
        cudaFree(bw);
      }
    }
    printf("  \"n%d\": {\"Dsyevd_ms\": %.3f, \"Dsyevd_novec_ms\": %.3f, \"Dsytrd_ms\": %.3f, \"Xsyevd_ms\": %.3f, \"XsyevBatched1_ms\": %.3f},\n", n,
           t_legacy, t_novec, t_sytrd, t_x, t_b);
    fflush(stdout);
    cudaFree(A); cudaFree(A0); cudaFree(W); cudaFree(tau); cudaFree(D); cudaFree(E);
  }
  // concurrency: the charge groups of the bench bond (rows of the two-site tensor per Sz sector)
  std::vector<int> group_n = {3762, 3412, 3412, 2325, 2325, 1341, 1341, 364, 364, 55, 55};
  for (int nstreams : {1, 2, 4, 8}) {
    std::vector<Job> jobs(group_n.size());
    std::vector<cudaStream_t> ss(nstreams);
    std::vector<cusolverDnHandle_t> hs(nstreams);
    for (int i = 0; i < nstreams; i++) { CK(cudaStreamCreateWithFlags(&ss[i], cudaStreamNonBlocking)); CK(cusolverDnCreate(&hs[i])); CK(cusolverDnSetStream(hs[i], ss[i])); }
    for (size_t k = 0; k < jobs.size(); k++) {
      Job& j = jobs[k];
      j.n = group_n[k]; j.s = ss[k % nstreams]; j.h = hs[k % nstreams];
      CK(cudaMalloc(&j.A, sizeof(double) * j.n * j.n)); CK(cudaMalloc(&j.A0, sizeof(double) * j.n * j.n)); CK(cudaMalloc(&j.W, sizeof(double) * j.n));
      CK(cudaMalloc(&j.info, 64));
      int lw = 0;
      CK(cusolverDnDsyevd_bufferSize(j.h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, j.n, j.A, j.n, j.W, &lw));
      j.lwork = lw; CK(cudaMalloc(&j.work, sizeof(double) * lw));
      fill_sym<<<592, 256, 0, j.s>>>(j.A0, j.n, 99u + (unsigned)k);
    }
    CK(cudaDeviceSynchronize());
    double best = 1e30;
    for (int rep = 0; rep < 3; rep++) {
      CK(cudaDeviceSynchronize());
      double t0 = now();
      std::vector<std::thread> th;
      for (int si = 0; si < nstreams; si++)
        th.emplace_back([&, si] {
          cudaSetDevice(0);
          for (size_t k = si; k < jobs.size(); k += nstreams) run_legacy(jobs[k]);
          cudaStreamSynchronize(ss[si]);
        });
      for (auto& t : th) t.join();
      CK(cudaDeviceSynchronize());
      double t1 = now();
      if (rep > 0 && t1 - t0 < best) best = t1 - t0;
    }
    printf("  \"groups_streams%d_ms\": %.3f,\n", nstreams, best * 1e3);
    fflush(stdout);
    for (auto& j : jobs) { cudaFree(j.A); cudaFree(j.A0); cudaFree(j.W); cudaFree(j.work); cudaFree(j.info); }
    for (int i = 0; i < nstreams; i++) { cusolverDnDestroy(hs[i]); cudaStreamDestroy(ss[i]); }
  }
  printf("  \"cusolver\": \"%d.%d.%d\"\n}\n", CUSOLVER_VER_MAJOR, CUSOLVER_VER_MINOR, CUSOLVER_VER_PATCH);
  return 0;
}

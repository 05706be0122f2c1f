// Experiment: do several cusolverDnDsyevd calls overlap when each runs in its own GREEN CONTEXT (a disjoint SM
// partition)?  On plain side streams they do not (profiles/r02c_eigh_bench.json: sytrd4_gpu is a persistent
// grid-synchronising kernel that owns the SMs), although the tridiagonalisation is latency bound (11-16 us per
// column) and would leave most of the machine idle.  Library calls only -- nothing here is product code.
// build: nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/eigh_green_bench.cu -o tools/_build/eigh_green_bench -lcusolver -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cusolverDn.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

#define CK(x) do { auto e__ = (x); if (e__ != 0) { fprintf(stderr, "error %d at %s:%d\n", (int)e__, __FILE__, __LINE__); exit(1); } } while (0)

__global__ void fill_sym(double* A, int n, unsigned seed) {
  for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < (long)n * n; e += (long)gridDim.x * blockDim.x) {
    int i = (int)(e % n), j = (int)(e / n);
    int a = i < j ? i : j, b = i < j ? j : i;
    unsigned h = seed ^ (unsigned)(a * 2654435761u) ^ (unsigned)(b * 40503u + 12345u);
    h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
    A[e] = (double)h / 4294967296.0 - 0.5 + (i == j ? 0.05 * n : 0.0);
  }
}
__global__ void check_kernel(const double* A0, const double* Z, const double* W, int n, double* out) {
  // residual of a few eigenpairs: max_i |A z_k - w_k z_k|_inf for k = blockIdx.x * stride
  int k = blockIdx.x * (n / gridDim.x);
  double worst = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double s = 0;
    for (int j = 0; j < n; j++) s += A0[(long)j * n + i] * Z[(long)k * n + j];
    worst = fmax(worst, fabs(s - W[k] * Z[(long)k * n + i]));
  }
  atomicMax((unsigned long long*)out, (unsigned long long)__double_as_longlong(worst));
}
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

struct Part { CUgreenCtx g; CUstream s; cusolverDnHandle_t h; unsigned sms; };
struct Job { int n; double *A, *A0, *W, *work; int lwork; int* info; };

int main(int argc, char** argv) {
  CK(cudaSetDevice(0));
  CK(cudaFree(0));
  CUdevice dev;
  CK(cuDeviceGet(&dev, 0));
  CUdevResource all;
  CK(cuDeviceGetDevResource(dev, &all, CU_DEV_RESOURCE_TYPE_SM));
  printf("{\n  \"sms_total\": %u,\n", all.sm.smCount);
  std::vector<int> group_n = {3412, 2722, 2718, 1348, 1341, 344, 339, 29, 26};
  std::vector<Job> jobs(group_n.size());
  for (size_t k = 0; k < jobs.size(); k++) {
    Job& j = jobs[k];
    j.n = group_n[k];
    CK(cudaMalloc(&j.A, sizeof(double) * j.n * j.n)); CK(cudaMalloc(&j.A0, sizeof(double) * j.n * j.n));
    CK(cudaMalloc(&j.W, sizeof(double) * j.n)); CK(cudaMalloc(&j.info, 64));
    fill_sym<<<592, 256>>>(j.A0, j.n, 99u + (unsigned)k);
  }
  CK(cudaDeviceSynchronize());
  const bool quick = argc > 1;
  for (int nparts : {1, 2, 3, 4}) {
    if (quick) break;
    const unsigned want = nparts == 1 ? all.sm.smCount : (all.sm.smCount / nparts) / 8 * 8;
    std::vector<CUdevResource> res(nparts);
    unsigned ng = nparts;
    CUdevResource rem;
    std::vector<Part> parts(nparts);
    if (nparts == 1) {
      res[0] = all;
    } else {
      CUresult r = cuDevSmResourceSplitByCount(res.data(), &ng, &all, &rem, 0, want);
      if (r != CUDA_SUCCESS || (int)ng < nparts) { printf("  \"parts%d\": \"split failed (%d, %u groups)\",\n", nparts, (int)r, ng); continue; }
    }
    bool ok = true;
    for (int p = 0; p < nparts && ok; p++) {
      CUdevResourceDesc desc;
      if (cuDevResourceGenerateDesc(&desc, &res[p], 1) != CUDA_SUCCESS) { ok = false; break; }
      if (cuGreenCtxCreate(&parts[p].g, desc, dev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS) { ok = false; break; }
      if (cuGreenCtxStreamCreate(&parts[p].s, parts[p].g, CU_STREAM_NON_BLOCKING, 0) != CUDA_SUCCESS) { ok = false; break; }
      parts[p].sms = res[p].sm.smCount;
      CK(cusolverDnCreate(&parts[p].h));
      CK(cusolverDnSetStream(parts[p].h, (cudaStream_t)parts[p].s));
    }
    if (!ok) { printf("  \"parts%d\": \"green context creation failed\",\n", nparts); continue; }
    // workspaces (queried per handle: the partition may change the size)
    for (size_t k = 0; k < jobs.size(); k++) {
      Job& j = jobs[k];
      int lw = 0, lwmax = 0;
      for (int p = 0; p < nparts; p++) {
        CK(cusolverDnDsyevd_bufferSize(parts[p].h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, j.n, j.A, j.n, j.W, &lw));
        lwmax = lw > lwmax ? lw : lwmax;
      }
      j.lwork = lwmax;
      CK(cudaMalloc(&j.work, sizeof(double) * lwmax));
    }
    // LPT on the measured cost model
    std::vector<std::vector<int>> mine(nparts);
    std::vector<double> load(nparts, 0.0);
    for (size_t k = 0; k < jobs.size(); k++) {
      int best = 0;
      for (int p = 1; p < nparts; p++) if (load[p] < load[best]) best = p;
      double n = jobs[k].n;
      load[best] += 0.0109 * n + 6.4e-10 * n * n * n;
      mine[best].push_back((int)k);
    }
    double best_t = 1e30;
    for (int rep = 0; rep < 3; rep++) {
      for (auto& j : jobs) CK(cudaMemcpy(j.A, j.A0, sizeof(double) * j.n * j.n, cudaMemcpyDeviceToDevice));
      CK(cudaDeviceSynchronize());
      double t0 = now();
      std::vector<std::thread> th;
      for (int p = 0; p < nparts; p++)
        th.emplace_back([&, p] {
          cudaSetDevice(0);
          for (int k : mine[p]) {
            Job& j = jobs[k];
            CK(cusolverDnDsyevd(parts[p].h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, j.n, j.A, j.n, j.W, j.work, j.lwork, j.info));
          }
          CK(cudaStreamSynchronize((cudaStream_t)parts[p].s));
        });
      for (auto& t : th) t.join();
      CK(cudaDeviceSynchronize());
      double t1 = now();
      if (rep > 0 && t1 - t0 < best_t) best_t = t1 - t0;
    }
    // correctness of the largest decomposition (residual of 8 eigenpairs)
    double* d_res; CK(cudaMalloc(&d_res, 8)); CK(cudaMemset(d_res, 0, 8));
    check_kernel<<<8, 256>>>(jobs[0].A0, jobs[0].A, jobs[0].W, jobs[0].n, d_res);
    double h_res = 0; CK(cudaMemcpy(&h_res, d_res, 8, cudaMemcpyDeviceToHost));
    int info0 = 0; CK(cudaMemcpy(&info0, jobs[0].info, 4, cudaMemcpyDeviceToHost));
    printf("  \"parts%d\": {\"sms_per_part\": %u, \"all_groups_ms\": %.3f, \"residual_n3412\": %.3e, \"info\": %d},\n", nparts, parts[0].sms,
           best_t * 1e3, h_res, info0);
    fflush(stdout);
    // single largest group on one partition
    {
      Job& j = jobs[0];
      double bt = 1e30;
      for (int rep = 0; rep < 3; rep++) {
        CK(cudaMemcpy(j.A, j.A0, sizeof(double) * j.n * j.n, cudaMemcpyDeviceToDevice));
        CK(cudaDeviceSynchronize());
        double t0 = now();
        CK(cusolverDnDsyevd(parts[0].h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, j.n, j.A, j.n, j.W, j.work, j.lwork, j.info));
        CK(cudaStreamSynchronize((cudaStream_t)parts[0].s));
        double t1 = now();
        if (rep > 0 && t1 - t0 < bt) bt = t1 - t0;
      }
      printf("  \"parts%d_single_n3412_ms\": %.3f,\n", nparts, bt * 1e3);
    }
    for (auto& j : jobs) { cudaFree(j.work); j.work = nullptr; }
    for (int p = 0; p < nparts; p++) { cusolverDnDestroy(parts[p].h); cuStreamDestroy(parts[p].s); cuGreenCtxDestroy(parts[p].g); }
    cudaFree(d_res);
  }
  // ---- unit-based partitions: split into 8-SM units and combine several units into one green context
  {
    for (unsigned mincount : {8u, 16u}) {
      std::vector<CUdevResource> res(32);
      unsigned ng = 32;
      CUdevResource rem;
      CUresult r = cuDevSmResourceSplitByCount(res.data(), &ng, &all, &rem, 0, mincount);
      printf("  \"split_min%u\": {\"rc\": %d, \"groups\": %u, \"sms_each\": %u, \"remaining\": %u},\n", mincount, (int)r, ng,
             ng ? res[0].sm.smCount : 0, r == CUDA_SUCCESS ? rem.sm.smCount : 0);
      if (r != CUDA_SUCCESS || ng < 4) continue;
      // combine: first half of the units in one context, second half in another
      unsigned h = ng / 2;
      CUdevResourceDesc d1, d2;
      CUresult r1 = cuDevResourceGenerateDesc(&d1, &res[0], h);
      CUresult r2 = cuDevResourceGenerateDesc(&d2, &res[h], ng - h);
      CUgreenCtx g1 = nullptr, g2 = nullptr;
      CUresult r3 = r1 == CUDA_SUCCESS ? cuGreenCtxCreate(&g1, d1, dev, CU_GREEN_CTX_DEFAULT_STREAM) : r1;
      CUresult r4 = r2 == CUDA_SUCCESS ? cuGreenCtxCreate(&g2, d2, dev, CU_GREEN_CTX_DEFAULT_STREAM) : r2;
      printf("  \"combine_min%u\": {\"desc1\": %d, \"desc2\": %d, \"ctx1\": %d, \"ctx2\": %d},\n", mincount, (int)r1, (int)r2, (int)r3, (int)r4);
      if (g1) cuGreenCtxDestroy(g1);
      if (g2) cuGreenCtxDestroy(g2);
    }
    // the same through cudaGetDriverEntryPoint (what the library does: no link-time libcuda dependency)
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuDevSmResourceSplitByCount", &fp, cudaEnableDefault, &q);
    printf("  \"entry_point_split\": {\"err\": %d, \"query\": %d, \"ptr\": %d},\n", (int)e, (int)q, fp != nullptr);
    e = cudaGetDriverEntryPoint("cuGreenCtxCreate", &fp, cudaEnableDefault, &q);
    printf("  \"entry_point_create\": {\"err\": %d, \"query\": %d, \"ptr\": %d},\n", (int)e, (int)q, fp != nullptr);
    e = cudaGetDriverEntryPoint("cuDeviceGetDevResource", &fp, cudaEnableDefault, &q);
    printf("  \"entry_point_getres\": {\"err\": %d, \"query\": %d, \"ptr\": %d},\n", (int)e, (int)q, fp != nullptr);
  }
  printf("  \"done\": 1\n}\n");
  return 0;
}

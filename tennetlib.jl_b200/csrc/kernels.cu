// Hand-written sm_100a kernels of the tnl_b200 hot path.
//
//  * gemm_tma_ws_kernel: grouped FP64 GEMM over charge sectors (all sectors larger than 64 x 64).  Blackwell has no
//                       tcgen05/UMMA kind for FP64; the FP64 tensor op is the warp-level DMMA.8x8x4
//                       (mma.sync.m8n8k4.f64).  Persistent, warp-specialised: a TMA producer warpgroup
//                       (cp.async.bulk.tensor + mbarriers, 128B-swizzled 128x16 operand tiles, 6 stages) feeds eight
//                       DMMA warps with 64x32 register tiles; split-K for plans that cannot fill the machine; an
//                       optional scatter epilogue stores the tiles straight into peer GPUs' staging slots (fused
//                       reduce-scatter of the sharded apply).  Replaces the per-block-pair `permutedims + BLAS.gemm!`
//                       loop NDTensors runs for every ITensor `*` on the reference hot path
//                       (src/mps/projcouplingmodel.jl:145-147,342-343; ITensorMPS ProjMPO.product).
//  * gemm_kernel      : the same product with cp.async (LDGSTS.128) staging into padded shared memory: 64x64 tiles for
//                       the small sectors (and 128x128 tiles when TNL_GEMM_TMA=0 selects the round-1 path).
//  * transform_kernel : HBM-bound regrouping pass that also applies the skinny MPO site operators
//                       (the `* W_j` steps of ProjMPO.product / `_makeL!`), one warp per output column.
//  * vec kernels      : flat Krylov-vector kernels (VectorInterface inner/add!!/scale!!/norm as used by
//                       KrylovKit.eigsolve from src/base/solver.jl:36) with warp-shuffle reductions.
#include "core.hpp"

#include <cuda.h>

#include <algorithm>

namespace tnl {

// =================================================================================================
// helpers
// =================================================================================================
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

// =================================================================================================
// grouped DGEMM
// =================================================================================================
// operand tile in shared memory.  KFAST = the operand is contiguous along k in global memory.
//   !KFAST : smem[k][x]  row stride BX+4      (x = m or n)
//    KFAST : smem[x][k]  row stride BK+4
// Both strides are == 4 (mod 16) doubles, which makes the DMMA fragment loads (8 x-values x 4 k-values per
// warp, 64-bit each) hit 16 distinct 8-byte bank pairs per half-warp: conflict free.
template <int BX, int BK, bool KFAST>
struct OpTile {
  static constexpr int LD = KFAST ? (BK + 4) : (BX + 4);
  static constexpr int ELEMS = KFAST ? BX * LD : BK * LD;
  // global element (x, k) lives at  g + x*sx + k*sk ; exactly one of sx/sk is 1
  template <int NT>
  __device__ static __forceinline__ void load(double* s, const double* g, int ld, int x0, int k0, int X, int K, int tid) {
    constexpr int CH = BX * BK / 2;   // 16-byte chunks
    if (KFAST) {
      constexpr int CPR = BK / 2;
#pragma unroll
      for (int c = tid; c < CH; c += NT) {
        int xx = c / CPR, kc = c % CPR;
        int x = x0 + xx, k = k0 + 2 * kc;
        int v = (x < X) ? max(0, min(2, K - k)) : 0;
        const double* src = v ? (g + (int64_t)x * ld + k) : g;
        cp_async16(s + xx * LD + 2 * kc, src, v * 8);
      }
    } else {
      constexpr int CPR = BX / 2;
#pragma unroll
      for (int c = tid; c < CH; c += NT) {
        int kk = c / CPR, xc = c % CPR;
        int x = x0 + 2 * xc, k = k0 + kk;
        int v = (k < K) ? max(0, min(2, X - x)) : 0;
        const double* src = v ? (g + (int64_t)k * ld + x) : g;
        cp_async16(s + kk * LD + 2 * xc, src, v * 8);
      }
    }
  }
  __device__ static __forceinline__ double frag(const double* s, int x, int k) {
    return KFAST ? s[x * LD + k] : s[k * LD + x];
  }
};

// -------------------------------------------------------------------------------------------------
// Global -> shared loader: every thread precomputes, once per tile, the source offset / validity of its 16-byte
// chunks, so that inside the k loop a load is "offset += step" + one predicate (no 64-bit multiplies, no branches).
// -------------------------------------------------------------------------------------------------
template <int BX, int BK, bool KFAST, int NT>
struct Loader {
  static constexpr int LD = OpTile<BX, BK, KFAST>::LD;
  static constexpr int CH = BX * BK / 2;
  static constexpr int NCH = CH / NT;          // chunks per thread
  static_assert(CH % NT == 0, "tile must split evenly over the threads");
  int goff[NCH];      // element offset of the chunk at k-tile 0 (relative to the operand base)
  int soff[NCH];      // shared-memory element offset inside a stage
  int kq[NCH];        // k position of the chunk inside a k-tile
  int xv[NCH];        // valid doubles along x (!KFAST: 0..2 ; KFAST: 0 or 2 meaning the row is valid)
  int kstep;          // offset advance per k-tile
  __device__ __forceinline__ void init(int ld, int x0, int X, int tid) {
#pragma unroll
    for (int i = 0; i < NCH; i++) {
      int c = tid + i * NT;
      if (KFAST) {
        constexpr int CPR = BK / 2;
        int xx = c / CPR, kc = c % CPR;
        int x = x0 + xx;
        kq[i] = 2 * kc;
        xv[i] = (x < X) ? 2 : 0;
        goff[i] = (x < X) ? x * ld + 2 * kc : 0;
        soff[i] = xx * LD + 2 * kc;
      } else {
        constexpr int CPR = BX / 2;
        int kk = c / CPR, xc = c % CPR;
        int x = x0 + 2 * xc;
        kq[i] = kk;
        xv[i] = max(0, min(2, X - x));
        goff[i] = xv[i] ? kk * ld + x : 0;
        soff[i] = kk * LD + 2 * xc;
      }
    }
    kstep = KFAST ? BK : BK * ld;
  }
  __device__ __forceinline__ void issue(double* s, const double* g, int kt, int K) const {
#pragma unroll
    for (int i = 0; i < NCH; i++) {
      int k = kt * BK + kq[i];
      int v;
      if (KFAST) v = xv[i] ? max(0, min(2, K - k)) : 0;
      else v = (k < K) ? xv[i] : 0;
      const double* src = g + (v ? (int64_t)goff[i] + (int64_t)kt * kstep : 0);
      cp_async16(s + soff[i], src, v * 8);
    }
  }
};

// Grouped DGEMM kernel: one CTA per (charge-sector problem, 128x128 or 64x64 tile); DMMA fragments double buffered
// in registers (the LDS of k-step kk+1 are issued before the DMMAs of k-step kk).
template <int BM, int BN, int BK, int WM, int WN, bool TA, bool TB, int STAGES>
__global__ void __launch_bounds__((BM / WM) * (BN / WN) * 32)
gemm_kernel(const GemmProblem* __restrict__ probs, const GemmTile* __restrict__ tiles,
               const double* __restrict__ Abase, const double* __restrict__ Bbase, double* __restrict__ Cbase,
            double alpha, int accum) {
  constexpr int NT = (BM / WM) * (BN / WN) * 32;
  using TileA = OpTile<BM, BK, TA>;
  using TileB = OpTile<BN, BK, !TB>;
  constexpr int STAGE_ELEMS = TileA::ELEMS + TileB::ELEMS;
  constexpr int MI = WM / 8, NI = WN / 8, KK = BK / 4;
  extern __shared__ __align__(16) double smem[];

  const GemmTile tile = tiles[blockIdx.x];
  const GemmProblem p = probs[tile.prob];
  const double* pA = Abase + p.a;
  const double* pB = Bbase + p.b;
  double* pC = Cbase + p.c;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm0 = (warp % (BM / WM)) * WM, wn0 = (warp / (BM / WM)) * WN;
  const int m0 = tile.m0, n0 = tile.n0;
  const int ktiles = (p.K + BK - 1) / BK;
  const int lr = lane >> 2, lc = lane & 3;

  Loader<BM, BK, TA, NT> la;
  Loader<BN, BK, !TB, NT> lb;
  la.init(p.lda, m0, p.M, tid);
  lb.init(p.ldb, n0, p.N, tid);

  double acc[MI][NI][2];
#pragma unroll
  for (int i = 0; i < MI; i++)
#pragma unroll
    for (int j = 0; j < NI; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
  for (int s = 0; s < STAGES - 1; s++) {
    if (s < ktiles) {
      double* sa = smem + s * STAGE_ELEMS;
      la.issue(sa, pA, s, p.K);
      lb.issue(sa + TileA::ELEMS, pB, s, p.K);
    }
    cp_async_commit();
  }
  // fragment element offsets of this lane inside a stage
  const int fa = TA ? (wm0 + lr) * TileA::LD + lc : lc * TileA::LD + wm0 + lr;
  const int fb = (!TB) ? (wn0 + lr) * TileB::LD + lc : lc * TileB::LD + wn0 + lr;
  constexpr int FA_I = TA ? 8 * TileA::LD : 8;            // +8 rows of m
  constexpr int FA_K = TA ? 4 : 4 * TileA::LD;            // +4 in k
  constexpr int FB_J = (!TB) ? 8 * TileB::LD : 8;
  constexpr int FB_K = (!TB) ? 4 : 4 * TileB::LD;

  // Early-sync main loop: the wait + barrier for k-tile kt+1 and the cp.async issue for k-tile kt+STAGES-1 sit
  // in front of the LAST group of 32 DMMAs of k-tile kt, so barrier latency and the first fragment loads of the
  // next tile are covered by tensor work that is already in flight.
  double a[2][MI], b[2][NI];
  cp_async_wait<STAGES - 2>();
  __syncthreads();
  {
    const double* sa = smem + fa;
    const double* sb = smem + TileA::ELEMS + fb;
#pragma unroll
    for (int i = 0; i < MI; i++) a[0][i] = sa[i * FA_I];
#pragma unroll
    for (int j = 0; j < NI; j++) b[0][j] = sb[j * FB_J];
  }
  static_assert(KK % 2 == 0, "k-steps per tile must be even (register double buffering across tiles)");
  for (int kt = 0; kt < ktiles; kt++) {
    const double* sa = smem + (kt % STAGES) * STAGE_ELEMS + fa;
    const double* sb = smem + (kt % STAGES) * STAGE_ELEMS + TileA::ELEMS + fb;
#pragma unroll
    for (int kk = 0; kk < KK; kk++) {
      const int cur = kk & 1, nxt = cur ^ 1;
      if (kk + 1 < KK) {
#pragma unroll
        for (int i = 0; i < MI; i++) a[nxt][i] = sa[i * FA_I + (kk + 1) * FA_K];
#pragma unroll
        for (int j = 0; j < NI; j++) b[nxt][j] = sb[j * FB_J + (kk + 1) * FB_K];
      } else {
        // tile kt+1 must have landed; every warp is past its last read of tile kt-1's stage
        {
          const int nk = kt + STAGES - 1;
          // groups committed so far: tiles 0 .. kt+STAGES-2  ->  allow STAGES-3 pending to have tile kt+1 complete
          cp_async_wait<(STAGES >= 3 ? STAGES - 3 : 0)>();
          __syncthreads();
          if (nk < ktiles) {
            double* st = smem + (nk % STAGES) * STAGE_ELEMS;
            la.issue(st, pA, nk, p.K);
            lb.issue(st + TileA::ELEMS, pB, nk, p.K);
          }
          cp_async_commit();
        }
        if (kt + 1 < ktiles) {
          const double* na = smem + ((kt + 1) % STAGES) * STAGE_ELEMS + fa;
          const double* nb = smem + ((kt + 1) % STAGES) * STAGE_ELEMS + TileA::ELEMS + fb;
#pragma unroll
          for (int i = 0; i < MI; i++) a[nxt][i] = na[i * FA_I];
#pragma unroll
          for (int j = 0; j < NI; j++) b[nxt][j] = nb[j * FB_J];
        }
      }
#pragma unroll
      for (int i = 0; i < MI; i++)
#pragma unroll
        for (int j = 0; j < NI; j++) dmma884(acc[i][j][0], acc[i][j][1], a[cur][i], b[cur][j]);
    }
  }
  cp_async_wait<0>();
#pragma unroll
  for (int i = 0; i < MI; i++) {
    int m = m0 + wm0 + i * 8 + lr;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < NI; j++) {
      int n = n0 + wn0 + j * 8 + 2 * lc;
      double* c = pC + (int64_t)n * p.ldc + m;
      // C = alpha * A B (+ C): alpha = +-1 and the accumulate mode serve the planar complex products
      if (n < p.N) c[0] = accum ? fma(alpha, acc[i][j][0], c[0]) : alpha * acc[i][j][0];
      if (n + 1 < p.N) c[p.ldc] = accum ? fma(alpha, acc[i][j][1], c[p.ldc]) : alpha * acc[i][j][1];
    }
  }
}

template <int BM, int BN, int BK, int WM, int WN, bool TA, bool TB, int STAGES>
static void launch_gemm_one(Ctx* ctx, const GemmProblem* probs, const GemmTile* tiles, int ntiles,
                            const double* A, const double* B, double* C, double alpha, int accum) {
  constexpr int NT = (BM / WM) * (BN / WN) * 32;
  constexpr size_t SMEM = sizeof(double) * STAGES * (OpTile<BM, BK, TA>::ELEMS + OpTile<BN, BK, !TB>::ELEMS);
  auto kern = gemm_kernel<BM, BN, BK, WM, WN, TA, TB, STAGES>;
  static bool configured = false;   // one static per template instantiation
  if (!configured) {
    CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
    configured = true;
  }
  kern<<<ntiles, NT, SMEM, ctx->stream>>>(probs, tiles, A, B, C, alpha, accum);
  CUDA_OK(cudaGetLastError());
  ctx->cnt.launches++;
  ctx->cnt.gemm_launches++;
}

template <int BM, int BN, int BK, int WM, int WN, int STAGES>
static void launch_gemm_cfg(Ctx* ctx, bool ta, bool tb, const GemmProblem* probs, const GemmTile* tiles, int ntiles,
                            const double* A, const double* B, double* C, double alpha = 1.0, int accum = 0) {
  if (ntiles == 0) return;
  if (!ta && !tb) launch_gemm_one<BM, BN, BK, WM, WN, false, false, STAGES>(ctx, probs, tiles, ntiles, A, B, C, alpha, accum);
  else if (!ta && tb) launch_gemm_one<BM, BN, BK, WM, WN, false, true, STAGES>(ctx, probs, tiles, ntiles, A, B, C, alpha, accum);
  else if (ta && !tb) launch_gemm_one<BM, BN, BK, WM, WN, true, false, STAGES>(ctx, probs, tiles, ntiles, A, B, C, alpha, accum);
  else launch_gemm_one<BM, BN, BK, WM, WN, true, true, STAGES>(ctx, probs, tiles, ntiles, A, B, C, alpha, accum);
}


// =================================================================================================
// grouped DGEMM, TMA-staged persistent warp-specialised kernel (128x128 tiles of every charge sector larger than 64x64)
// =================================================================================================
// One CTA per SM walks the tile list with stride gridDim.x.  Warpgroup 0 is the TMA producer: one elected thread issues
// cp.async.bulk.tensor (UTMALDG) for the two operand tiles (128 x 16 doubles each) of a stage and the group gives its
// registers away with setmaxnreg; warpgroups 1-2 are the eight DMMA warps (64 x 32 register tiles, 232 registers).
// Stages are handed over with a full / empty mbarrier pair per stage, so there is no CTA-wide barrier in the k loop:
// the DMMA warps wait only for data, never for each other, and the producer runs up to STAGES k-tiles ahead ACROSS
// tile boundaries -- while the 64 accumulators of tile t are stored the first k-tiles of tile t+1 are already in
// shared memory and its first fragments in registers, so the DMMA pipe does not drain between tiles.  Out-of-range
// rows / columns / k are zero-filled by the TMA unit: no predication in the kernel.  Measured on B200
// (profiles/r02e_gemm.txt): 36.3 TFLOP/s at 8192^3 (cuBLAS DGEMM: 35.5), 34.4 at the `L v` shape of the chi = 4096
// apply, against 33.4 / 31.6 for the cp.async kernel above (which paid a prologue per tile and a barrier per k-tile).
//
// Shared-memory layouts written by the TMA (SWIZZLE_128B: 16-byte chunk index ^= 128-byte row index mod 8):
//   k contiguous in global (KF): one box {16 k, 128 x}: row x = 128 bytes = the 16 k values
//   x contiguous in global     : one 3-d box {16 x, 16 k, 8 chunks} of the operand seen as (x mod 16, k, x div 16):
//                                chunk b holds x in [16b, 16b+16), row k = 128 bytes.  (The inner box extent is capped
//                                at the 128-byte swizzle span, hence the chunked view.)  The last chunk of a ragged
//                                extent reads up to 15 doubles past the row end -- values that only reach accumulator
//                                rows / columns outside the matrix, which are never stored; Ctx::alloc pads every
//                                allocation so that the over-read stays inside mapped memory.
// The DMMA fragment of k-step s takes the k values  kbase(lc) ^ 2s,  kbase = (0, 3, 12, 15)[lc]  (the same
// permutation for A and B, so the product is unchanged): with it the 8 x 4 fragment loads of a half-warp fall into
// 16 distinct 8-byte bank pairs for BOTH layouts (checked exhaustively in tests/test_host_logic.py).
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "TNL_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra TNL_DONE;\n"
      "bra TNL_WAIT;\n"
      "TNL_DONE:\n"
      "}\n" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::"r"(dst),
               "l"(tmap), "r"(bar), "r"(c0), "r"(c1)
               : "memory");
}

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n" ::"r"(dst),
               "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

struct alignas(128) TmaDesc { unsigned char b[128]; };   // CUtensorMap (opaque on the device side)

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}

// EPI = 2 (planar complex products): two operand pairs share the accumulators, C = alpha * (A1 B1 +- A2 B2) -- the k loop
// runs over the first pair's k-tiles and then over the second pair's; for "-" the accumulators are negated at the
// boundary and once more at the end.  One launch per plane of a complex x complex product instead of two, no
// read-modify-write of C, twice the k range per tile.
// EPI = 0: C = alpha * A B (+ C) in place.  EPI = 1 (multi-GPU): scatter epilogue -- every element is stored into the
// staging slot of the rank that owns its column (peer memory over NVLink), and the last CTA to finish flags "epoch
// complete" to every peer (see run_gemm_reduce_scatter).
template <bool KFA, bool KFB, int STAGES, int EPI>
__global__ void __launch_bounds__(384, 1)
gemm_tma_ws_kernel(const TmaTile* __restrict__ tiles, int ntiles, const TmaDesc* __restrict__ maps, double* __restrict__ Cbase,
                   double* __restrict__ Wbase, double alpha, int accum, ScatterArgs sargs, const TmaDesc* __restrict__ maps2,
                   int negate2) {
  constexpr int MI = 8, NI = 4;
  constexpr int TILE_ELEMS = 128 * 16;
  constexpr int STAGE_ELEMS = 2 * TILE_ELEMS;
  constexpr uint32_t STAGE_BYTES = STAGE_ELEMS * 8;
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw_u32 = smem_u32(smem_raw);
  const uint32_t base_u32 = (raw_u32 + 1023u) & ~1023u;
  const double* smem = reinterpret_cast<const double*>(smem_raw + (base_u32 - raw_u32));
  const uint32_t full0 = base_u32 + STAGES * STAGE_BYTES;
  const uint32_t empty0 = full0 + 8 * STAGES;
  const int tid = threadIdx.x;
  const int G = gridDim.x;

  if (tid == 0) {
    for (int s = 0; s < STAGES; s++) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 8); }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();

  if (tid < 128) {
    // ---------------------------------------------------------------- producer warpgroup
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;\n");
    if (tid == 0) {
      int stage = 0;
      uint32_t phase = 1;                          // a fresh mbarrier passes a wait on parity 1
      for (int ti = blockIdx.x; ti < ntiles; ti += G) {
        const int m0 = tiles[ti].m0, n0 = tiles[ti].n0, kt1 = tiles[ti].kt0 + tiles[ti].ktiles;
        for (int half = 0; half < (EPI == 2 ? 2 : 1); half++) {
          const TmaDesc* ma = (half ? maps2 : maps) + 2 * tiles[ti].prob;
          for (int kt = tiles[ti].kt0; kt < kt1; kt++) {
            mbar_wait(empty0 + 8 * stage, phase);
            const uint32_t bar = full0 + 8 * stage;
            const uint32_t dA = base_u32 + stage * STAGE_BYTES, dB = dA + TILE_ELEMS * 8;
            mbar_arrive_expect_tx(bar, STAGE_BYTES);
            if (KFA) tma_load_2d(dA, ma, bar, kt * 16, m0);
            else tma_load_3d(dA, ma, bar, 0, kt * 16, m0 >> 4);
            if (KFB) tma_load_2d(dB, ma + 1, bar, kt * 16, n0);
            else tma_load_3d(dB, ma + 1, bar, 0, kt * 16, n0 >> 4);
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
    return;
  }
  // ------------------------------------------------------------------ DMMA warpgroups
  asm volatile("setmaxnreg.inc.sync.aligned.u32 232;\n");
  const int lane = tid & 31, warp = (tid >> 5) - 4;
  const int wm0 = (warp & 1) * 64, wn0 = (warp >> 1) * 32;
  const int lr = lane >> 2, lc = lane & 3;
  const int kb = (lc & 1) * 3 + (lc >> 1) * 12;
  int ea[4], eb[4];
  const int da = KFA ? 0 : ((lc >> 1) ? -8 : 8);
  const int db = KFB ? 0 : ((lc >> 1) ? -8 : 8);
#pragma unroll
  for (int s = 0; s < 4; s++) {
    const int k = kb ^ (2 * s);
    if (KFA) ea[s] = (wm0 + lr) * 16 + ((((k >> 1) ^ lr)) << 1) + (k & 1);
    else     ea[s] = (wm0 >> 4) * 256 + k * 16 + 2 * ((lr >> 1) ^ (k & 3)) + (lr & 1) + 8 * ((k >> 2) & 1);
    if (KFB) eb[s] = (wn0 + lr) * 16 + ((((k >> 1) ^ lr)) << 1) + (k & 1);
    else     eb[s] = (wn0 >> 4) * 256 + k * 16 + 2 * ((lr >> 1) ^ (k & 3)) + (lr & 1) + 8 * ((k >> 2) & 1);
  }
#define TNL_FRAG_A(st, s, i) ((st)[ea[s] + (KFA ? (i) * 128 : ((i) >> 1) * 256 + ((i) & 1) * ((s) < 2 ? da : -da))])
#define TNL_FRAG_B(st, s, j) ((st)[TILE_ELEMS + eb[s] + (KFB ? (j) * 128 : ((j) >> 1) * 256 + ((j) & 1) * ((s) < 2 ? db : -db))])

  double acc[MI][NI][2];
  double a[2][MI], b[2][NI];
  int c_stage = 0;
  uint32_t c_phase = 0;
  mbar_wait(full0, 0);
#pragma unroll
  for (int i = 0; i < MI; i++) a[0][i] = TNL_FRAG_A(smem, 0, i);
#pragma unroll
  for (int j = 0; j < NI; j++) b[0][j] = TNL_FRAG_B(smem, 0, j);

  TmaTile cur = tiles[blockIdx.x];
  for (int ti = blockIdx.x; ti < ntiles; ti += G) {
    const int tn = ti + G;
    TmaTile nxt = cur;
    if (tn < ntiles) nxt = tiles[tn];
#pragma unroll
    for (int i = 0; i < MI; i++)
#pragma unroll
      for (int j = 0; j < NI; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
    const int nkt = EPI == 2 ? 2 * cur.ktiles : cur.ktiles;
    for (int kt = 0; kt < nkt; kt++) {
      if (EPI == 2 && negate2 && kt == cur.ktiles) {     // second operand pair enters with the opposite sign
#pragma unroll
        for (int i = 0; i < MI; i++)
#pragma unroll
          for (int j = 0; j < NI; j++) { acc[i][j][0] = -acc[i][j][0]; acc[i][j][1] = -acc[i][j][1]; }
      }
      const double* st = smem + c_stage * STAGE_ELEMS;
      const int n_stage = (c_stage + 1 == STAGES) ? 0 : c_stage + 1;
      const uint32_t n_phase = (c_stage + 1 == STAGES) ? (c_phase ^ 1u) : c_phase;
#pragma unroll
      for (int s = 0; s < 4; s++) {
        const int cb = s & 1, nb = cb ^ 1;
        if (s < 3) {
#pragma unroll
          for (int i = 0; i < MI; i++) a[nb][i] = TNL_FRAG_A(st, s + 1, i);
#pragma unroll
          for (int j = 0; j < NI; j++) b[nb][j] = TNL_FRAG_B(st, s + 1, j);
        } else if ((kt + 1 < nkt) || (tn < ntiles)) {
          // first fragments of the next k-tile of this CTA's sequence (possibly of the next output tile)
          mbar_wait(full0 + 8 * n_stage, n_phase);
          const double* ns = smem + n_stage * STAGE_ELEMS;
#pragma unroll
          for (int i = 0; i < MI; i++) a[nb][i] = TNL_FRAG_A(ns, 0, i);
#pragma unroll
          for (int j = 0; j < NI; j++) b[nb][j] = TNL_FRAG_B(ns, 0, j);
        }
#pragma unroll
        for (int i = 0; i < MI; i++)
#pragma unroll
          for (int j = 0; j < NI; j++) dmma884(acc[i][j][0], acc[i][j][1], a[cb][i], b[cb][j]);
      }
      // every fragment of this stage has been consumed by an issued DMMA: hand the stage back to the producer
      __syncwarp();
      if (lane == 0) mbar_arrive(empty0 + 8 * c_stage);
      c_stage = n_stage;
      c_phase = n_phase;
    }
    if (EPI == 2 && negate2) {
#pragma unroll
      for (int i = 0; i < MI; i++)
#pragma unroll
        for (int j = 0; j < NI; j++) { acc[i][j][0] = -acc[i][j][0]; acc[i][j][1] = -acc[i][j][1]; }
    }
    if (EPI == 0 || EPI == 2) {
      // split-K part: raw partial product into the workspace (alpha and the accumulate mode are applied by the reduction)
      double* pC = (cur.split ? Wbase : Cbase) + cur.c;
      const double al = cur.split ? 1.0 : alpha;
      const int ac = cur.split ? 0 : accum;
#pragma unroll
      for (int i = 0; i < MI; i++) {
        const int m = cur.m0 + wm0 + i * 8 + lr;
        if (m >= cur.M) continue;
#pragma unroll
        for (int j = 0; j < NI; j++) {
          const int n = cur.n0 + wn0 + j * 8 + 2 * lc;
          double* c = pC + (int64_t)n * cur.ldc + m;
          if (n < cur.N) c[0] = ac ? fma(al, acc[i][j][0], c[0]) : al * acc[i][j][0];
          if (n + 1 < cur.N) c[cur.ldc] = ac ? fma(al, acc[i][j][1], c[cur.ldc]) : al * acc[i][j][1];
        }
      }
    } else {
      const ScatterProb sp = sargs.probs[cur.prob];
      int2 ri[MI];
#pragma unroll
      for (int i = 0; i < MI; i++) {
        const int m = cur.m0 + wm0 + i * 8 + lr;
        ri[i] = m < cur.M ? sp.rowinfo[m] : make_int2(-1, 0);
      }
#pragma unroll
      for (int j = 0; j < NI; j++) {
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const int n = cur.n0 + wn0 + j * 8 + 2 * lc + e;
          if (n >= cur.N) continue;
          const int2 ci = sp.colinfo[n];
          const int seg = ci.x & 0xffff;
          double* dst = sargs.peer_slots[ci.x >> 16];
#pragma unroll
          for (int i = 0; i < MI; i++) {
            if (ri[i].x < 0) continue;
            dst[sp.T[ri[i].x * sp.nseg + seg] + ri[i].y + (int64_t)ci.y * sp.cstride[ri[i].x]] = alpha * acc[i][j][e];
          }
        }
      }
    }
    cur = nxt;
  }
#undef TNL_FRAG_A
#undef TNL_FRAG_B
  if (EPI == 1) {
    // every DMMA thread fences its own remote stores, the 256 of them meet on a named barrier (the producer
    // warpgroup has exited), and the last CTA of the grid publishes the epoch to every peer
    __threadfence_system();
    asm volatile("bar.sync 1, 256;\n" ::: "memory");
    if (tid == 128) {
      const unsigned int prev = atomicAdd(sargs.done, 1u);
      if (prev == gridDim.x - 1) {
        *sargs.done = 0;
        __threadfence_system();
        for (int k = 0; k < sargs.world; k++) {
          volatile unsigned long long* f = sargs.peer_flags[k] + sargs.rank;       // "ready" word of source `rank` on peer k
          *f = sargs.epoch;
        }
        __threadfence_system();
      }
    }
  }
}

// ---- fused reduce-scatter: flag waits and the ordered sum over the source slots -----------------------------------
// flags of one rank: words [0, 64) "ready": source s has delivered epoch e;  words [64, 128) "consumed": destination d
// has summed (and cleared) what this rank sent it in epoch e
__global__ void stage_wait_consumed_kernel(const volatile unsigned long long* my_flags, int world, unsigned long long need) {
  const int k = threadIdx.x;
  if (k < world) {
    const long long t0 = clock64();
    while (my_flags[64 + k] < need) {
      __nanosleep(100);
      if (clock64() - t0 > 120000000000ll) __trap();       // a peer died: fail instead of hanging the GPU
    }
  }
}
// out[i] = sum_s slot_s[i] (s ascending: the same order on every rank and in every run), slots cleared for the next epoch
__global__ void __launch_bounds__(256)
stage_reduce_kernel(double* __restrict__ out, double* slots, int64_t slot_cap, int64_t n, int world, volatile unsigned long long* my_flags,
                    unsigned long long* const* peer_flags, int rank, unsigned long long epoch, unsigned int* done) {
  if (threadIdx.x < world) {
    const long long t0 = clock64();
    while (my_flags[threadIdx.x] < epoch) {
      __nanosleep(100);
      if (clock64() - t0 > 120000000000ll) __trap();
    }
  }
  __syncthreads();
  __threadfence_system();
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2; i < n; i += (int64_t)gridDim.x * blockDim.x * 2) {
    double2 acc = make_double2(0.0, 0.0);
    for (int s0 = 0; s0 < world; s0 += 8) {          // all loads of up to eight slots in flight before the first add
      double2 v[8];
#pragma unroll
      for (int u = 0; u < 8; u++)
        v[u] = s0 + u < world ? __ldcv(reinterpret_cast<const double2*>(slots + (int64_t)(s0 + u) * slot_cap + i)) : make_double2(0.0, 0.0);
#pragma unroll
      for (int u = 0; u < 8; u++) {
        acc.x += v[u].x; acc.y += v[u].y;
        if (s0 + u < world) *reinterpret_cast<double2*>(slots + (int64_t)(s0 + u) * slot_cap + i) = make_double2(0.0, 0.0);
      }
    }
    *reinterpret_cast<double2*>(out + i) = acc;
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int prev = atomicAdd(done, 1u);
    if (prev == gridDim.x - 1) {
      *done = 0;
      __threadfence_system();
      for (int k = 0; k < world; k++) {
        volatile unsigned long long* f = peer_flags[k] + 64 + rank;                // "consumed" word of destination `rank` on peer k
        *f = epoch;
      }
    }
  }
}

// ---- host side: descriptor sets ----------------------------------------------------------------
typedef CUresult (*TnlEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                     const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static TnlEncodeTiledFn tma_encoder() {
  static TnlEncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    TNL_CHECK(p != nullptr && q == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled is not available in this driver");
    return (TnlEncodeTiledFn)p;
  }();
  return fn;
}
// operand of extent X (rows of op(.)) x K stored with leading dimension ld; kfast: element (x, k) at x*ld + k
static void encode_operand(CUtensorMap* m, const double* base, int X, int K, int ld, bool kfast) {
  cuuint64_t gdim[3] = {(cuuint64_t)K, (cuuint64_t)X, 1};
  cuuint64_t gstr[2] = {(cuuint64_t)ld * sizeof(double), 0};
  cuuint32_t box[3] = {16u, 128u, 1u};
  cuuint32_t estr[3] = {1u, 1u, 1u};
  if (!kfast) {                                    // (x mod 16, k, x div 16)
    gdim[0] = 16; gdim[1] = (cuuint64_t)K; gdim[2] = (cuuint64_t)(X + 15) / 16;
    gstr[1] = 16 * sizeof(double);
    box[1] = 16u; box[2] = 8u;
  }
  TNL_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (ld & 1) == 0, "GEMM operand is not 16-byte aligned");
  CUresult r = tma_encoder()(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, kfast ? 2 : 3, const_cast<double*>(base), gdim, gstr, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  TNL_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
}
static const TmaDesc* tma_maps_for(Ctx* ctx, GemmPlan& p, const double* A, const double* B) {
  const size_t per_set = 2 * p.probs.size();
  for (size_t i = 0; i < p.mapsets.size(); i++)
    if (p.mapsets[i].A == A && p.mapsets[i].B == B) return (const TmaDesc*)p.d_maps + i * per_set;
  if ((int)p.mapsets.size() == GemmPlan::MAPSETS_MAX) p.mapsets.clear();   // stream order keeps in-flight launches safe
  const size_t slot = p.mapsets.size();
  static_assert(sizeof(CUtensorMap) == 128 && sizeof(TmaDesc) == 128, "descriptor size");
  CUtensorMap* h = (CUtensorMap*)ctx->stage_pinned(per_set * sizeof(CUtensorMap));
  for (size_t i = 0; i < p.probs.size(); i++) {
    const GemmProblem& g = p.probs[i];
    encode_operand(&h[2 * i], A + g.a, g.M, g.K, g.lda, p.transA);
    encode_operand(&h[2 * i + 1], B + g.b, g.N, g.K, g.ldb, !p.transB);
  }
  TmaDesc* d = (TmaDesc*)p.d_maps + slot * per_set;
  CUDA_OK(cudaMemcpyAsync(d, h, per_set * sizeof(CUtensorMap), cudaMemcpyHostToDevice, ctx->stream));
  p.mapsets.push_back(GemmPlan::MapSet{A, B});
  return d;
}

// C (+)= alpha * sum_s ws[s], parts added in order (deterministic)
__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const SplitDesc* __restrict__ sd, int nsplit, double* __restrict__ C, const double* __restrict__ W, double alpha, int accum) {
  for (int q = blockIdx.y; q < nsplit; q += gridDim.y) {
    const SplitDesc d = sd[q];
    const int64_t tot = (int64_t)d.M * d.N, part = (int64_t)d.ldw * d.N;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (int64_t)gridDim.x * blockDim.x) {
      const int m = (int)(e % d.M);
      const int64_t n = e / d.M;
      const double* w = W + d.ws + n * d.ldw + m;
      double s = 0.0;
      for (int k = 0; k < d.S; k++) s += w[(int64_t)k * part];
      double* c = C + d.c + n * d.ldc + m;
      *c = accum ? fma(alpha, s, *c) : alpha * s;
    }
  }
}

template <bool KFA, bool KFB, int EPI = 0>
static void launch_gemm_tma_one(Ctx* ctx, GemmPlan& p, const TmaDesc* maps, double* C, double alpha, int accum,
                                const ScatterArgs& sargs = ScatterArgs(), const TmaDesc* maps2 = nullptr, int negate2 = 0) {
  constexpr int STAGES = 6;
  constexpr size_t SMEM = (size_t)STAGES * 2 * 128 * 16 * sizeof(double) + 2 * STAGES * 8 + 1024;
  auto kern = gemm_tma_ws_kernel<KFA, KFB, STAGES, EPI>;
  static bool configured = false;
  if (!configured) {
    CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
    configured = true;
  }
  const int ntiles = (int)p.tiles_tma.size();
  const int grid = std::min(ntiles, ctx->num_sms);
  kern<<<grid, 384, SMEM, ctx->stream>>>(p.d_tiles_tma, ntiles, maps, C, p.d_splitws, alpha, accum, sargs, maps2, negate2);
  if (!p.splits.empty()) {
    TNL_CHECK(EPI == 0, "split-K parts cannot be scattered");
    int64_t big = 0;
    for (auto& d : p.splits) big = std::max<int64_t>(big, (int64_t)d.M * d.N);
    dim3 rg((unsigned)std::min<int64_t>((big + 255) / 256, 1184), (unsigned)std::min<size_t>(p.splits.size(), 64));
    splitk_reduce_kernel<<<rg, 256, 0, ctx->stream>>>(p.d_splits, (int)p.splits.size(), C, p.d_splitws, alpha, accum);
    ctx->cnt.launches++;
  }
  CUDA_OK(cudaGetLastError());
  ctx->cnt.launches++;
  ctx->cnt.gemm_launches++;
}
static void launch_gemm_tma(Ctx* ctx, GemmPlan& p, const double* A, const double* B, double* C, double alpha, int accum) {
  if (p.tiles_tma.empty()) return;
  const TmaDesc* maps = tma_maps_for(ctx, p, A, B);
  const bool kfa = p.transA, kfb = !p.transB;
  if (kfa && kfb) launch_gemm_tma_one<true, true>(ctx, p, maps, C, alpha, accum);
  else if (kfa && !kfb) launch_gemm_tma_one<true, false>(ctx, p, maps, C, alpha, accum);
  else if (!kfa && kfb) launch_gemm_tma_one<false, true>(ctx, p, maps, C, alpha, accum);
  else launch_gemm_tma_one<false, false>(ctx, p, maps, C, alpha, accum);
}

// (output ranges without a contribution -- zero_fill -- need no store in the fused path: the staging slots are zero)
// Scalar all-reduce over the peers' staging headers (see comm.cpp for the layout): every rank drops its n <= 8 values
// into mailbox [epoch parity][rank] of every peer, bumps the sequence word, waits for the W sequence words of its own
// header and sums the mailboxes in rank order.  Two parities are enough: a rank can only be one all-reduce ahead of the
// slowest one, because completing epoch e needs everybody's contribution to e.
__global__ void peer_allreduce_small_kernel(double* vals, int n, unsigned long long* const* peer_hdr, unsigned long long* my_hdr,
                                            int rank, int world, unsigned long long epoch) {
  const int k = threadIdx.x;
  const int par = (int)(epoch & 1);
  if (k < world) {
    volatile double* mail = reinterpret_cast<volatile double*>(reinterpret_cast<char*>(peer_hdr[k]) + 2048) + (par * 64 + rank) * 8;
    for (int j = 0; j < n; j++) mail[j] = vals[j];
    __threadfence_system();
    volatile unsigned long long* seq = peer_hdr[k] + 128 + par * 64 + rank;
    *seq = epoch;
    volatile unsigned long long* mine = my_hdr + 128 + par * 64 + k;
    const long long t0 = clock64();
    while (*mine < epoch) {
      if (clock64() - t0 > 120000000000ll) __trap();
    }
  }
  __syncthreads();
  __threadfence_system();
  if (k < n) {
    const double* mail = reinterpret_cast<const double*>(reinterpret_cast<const char*>(my_hdr) + 2048) + par * 64 * 8;
    double s = 0.0;
    for (int src = 0; src < world; src++) s += __ldcv(mail + src * 8 + k);
    vals[k] = s;
  }
}
void peer_allreduce_small(Ctx* ctx, double* buf, int n) {
  Ctx::PeerStage& ps = ctx->pstage;
  ps.ar_epoch++;
  peer_allreduce_small_kernel<<<1, 64, 0, ctx->stream>>>(buf, n, ps.d_peer_flags, (unsigned long long*)ps.base, ctx->rank, ctx->world,
                                                          ps.ar_epoch);
  CUDA_OK(cudaGetLastError());
  ctx->cnt.launches++;
}

bool plan_all_tma(const GemmPlan& p) { return !p.tiles_tma.empty() && p.tiles_big.empty() && p.tiles_small.empty(); }

void run_gemm_reduce_scatter(Ctx* ctx, GemmPlan& p, const double* A, const double* B, const ScatterProb* d_scatter, double* out,
                             int64_t nloc) {
  Ctx::PeerStage& ps = ctx->pstage;
  TNL_CHECK(ps.ok && (size_t)nloc <= ps.slot_cap && (nloc & 1) == 0, "staging area not set up for this vector length");
  TNL_CHECK(plan_all_tma(p) && !p.transA && p.transB, "the fused reduce-scatter needs a plan of 128x128 tiles of A B^T");
  unsigned long long* my_flags = (unsigned long long*)ps.base;
  double* my_slots = (double*)((char*)ps.base + 16384);        // kStageFlagBytes (comm.cpp)
  ps.epoch++;
  Ctx::ProfRec rec{};
  const bool prof = ctx->prof_gemm;
  // nobody may still hold unsummed data of the previous epoch in the slots this rank is about to overwrite
  {
    Ctx::Scope prof_scope(ctx, 3);
    prof_scope.r.tiles = 4;
    stage_wait_consumed_kernel<<<1, 64, 0, ctx->stream>>>(my_flags, ctx->world, ps.epoch - 1);
  }
  if (prof) {
    CUDA_OK(cudaEventCreate(&rec.a));
    CUDA_OK(cudaEventCreate(&rec.b));
    CUDA_OK(cudaEventRecord(rec.a, ctx->stream));
  }
  ScatterArgs sa;
  sa.probs = d_scatter;
  sa.peer_slots = ps.d_peer_slots;
  sa.peer_flags = ps.d_peer_flags;
  sa.done = ps.d_done;
  sa.epoch = ps.epoch;
  sa.rank = ctx->rank;
  sa.world = ctx->world;
  const TmaDesc* maps = tma_maps_for(ctx, p, A, B);
  launch_gemm_tma_one<false, false, 1>(ctx, p, maps, nullptr, 1.0, 0, sa);
  if (prof) {
    CUDA_OK(cudaEventRecord(rec.b, ctx->stream));
    rec.flops = p.flops;
    rec.cat = 0;
    rec.tiles = (int)p.tiles_tma.size();
    ctx->prof_recs.push_back(rec);
  }
  ctx->cnt.gemm_flops += p.flops;
  {
    Ctx::Scope prof_scope(ctx, 3);
    prof_scope.r.tiles = 5;
    const int grid = (int)std::min<int64_t>((nloc / 2 + 255) / 256, (int64_t)ctx->num_sms * 8);
    stage_reduce_kernel<<<grid, 256, 0, ctx->stream>>>(out, my_slots, (int64_t)ps.slot_cap, nloc, ctx->world, my_flags, ps.d_peer_flags,
                                                        ctx->rank, ps.epoch, ps.d_done);
    CUDA_OK(cudaGetLastError());
    ctx->cnt.launches += 2;
    ctx->cnt.allreduce_bytes += 8.0 * nloc * ctx->world;
  }
}

// C = s1 * A1 B1 + s2 * A2 B2   (s1, s2 = +-1): the two halves of one plane of a complex x complex product.  The
// 128x128 tiles take both operand pairs in one launch (EPI = 2); small tiles and split-K plans fall back to two
// accumulating launches.
void run_gemm_dual(Ctx* ctx, GemmPlan& p, const double* A1, const double* B1, double s1, const double* A2, const double* B2, double s2,
                   double* C) {
  if (p.tiles_tma.empty() || !p.splits.empty()) {
    run_gemm(ctx, p, A1, B1, C, s1, false);
    run_gemm(ctx, p, A2, B2, C, s2, true);
    return;
  }
  for (auto& z : p.zero_fill) CUDA_OK(cudaMemsetAsync(C + z.first, 0, z.second * sizeof(double), ctx->stream));
  Ctx::ProfRec rec{};
  const bool prof = ctx->prof_gemm;
  if (prof) {
    CUDA_OK(cudaEventCreate(&rec.a));
    CUDA_OK(cudaEventCreate(&rec.b));
    CUDA_OK(cudaEventRecord(rec.a, ctx->stream));
  }
  tma_maps_for(ctx, p, A1, B1);
  const TmaDesc* m2 = tma_maps_for(ctx, p, A2, B2);
  const TmaDesc* m1 = tma_maps_for(ctx, p, A1, B1);      // (again: the second request may have recycled the cache)
  const int neg = (s2 * s1 < 0) ? 1 : 0;
  const bool kfa = p.transA, kfb = !p.transB;
  if (kfa && kfb) launch_gemm_tma_one<true, true, 2>(ctx, p, m1, C, s1, 0, ScatterArgs(), m2, neg);
  else if (kfa && !kfb) launch_gemm_tma_one<true, false, 2>(ctx, p, m1, C, s1, 0, ScatterArgs(), m2, neg);
  else if (!kfa && kfb) launch_gemm_tma_one<false, true, 2>(ctx, p, m1, C, s1, 0, ScatterArgs(), m2, neg);
  else launch_gemm_tma_one<false, false, 2>(ctx, p, m1, C, s1, 0, ScatterArgs(), m2, neg);
  // sectors below 64 in an extent: the cp.async kernel, two accumulating launches
  launch_gemm_cfg<128, 128, 16, 64, 32, 4>(ctx, p.transA, p.transB, p.d_probs, p.d_tiles_big, (int)p.tiles_big.size(), A1, B1, C, s1, 0);
  launch_gemm_cfg<128, 128, 16, 64, 32, 4>(ctx, p.transA, p.transB, p.d_probs, p.d_tiles_big, (int)p.tiles_big.size(), A2, B2, C, s2, 1);
  launch_gemm_cfg<64, 64, 16, 32, 32, 4>(ctx, p.transA, p.transB, p.d_probs, p.d_tiles_small, (int)p.tiles_small.size(), A1, B1, C, s1, 0);
  launch_gemm_cfg<64, 64, 16, 32, 32, 4>(ctx, p.transA, p.transB, p.d_probs, p.d_tiles_small, (int)p.tiles_small.size(), A2, B2, C, s2, 1);
  if (prof) {
    CUDA_OK(cudaEventRecord(rec.b, ctx->stream));
    rec.flops = 2.0 * p.flops;
    rec.cat = 0;
    rec.tiles = (int)(p.tiles_big.size() + p.tiles_small.size() + p.tiles_tma.size());
    ctx->prof_recs.push_back(rec);
  }
  ctx->cnt.gemm_flops += 2.0 * p.flops;
}

void run_gemm(Ctx* ctx, GemmPlan& p, const double* A, const double* B, double* C, double alpha, bool accum) {
  if (!accum)
    for (auto& z : p.zero_fill) CUDA_OK(cudaMemsetAsync(C + z.first, 0, z.second * sizeof(double), ctx->stream));
  Ctx::ProfRec rec{};
  const bool prof = ctx->prof_gemm && !(p.tiles_big.empty() && p.tiles_tma.empty());
  if (prof) {
    CUDA_OK(cudaEventCreate(&rec.a));
    CUDA_OK(cudaEventCreate(&rec.b));
    CUDA_OK(cudaEventRecord(rec.a, ctx->stream));
  }
  // sectors larger than 64 in both extents: persistent TMA kernel (or, with TNL_GEMM_TMA=0, 128x128x16 tiles fed by
  // 4 cp.async stages); 64x64x16 tiles, 4 warps x (32x32), for the small ones
  launch_gemm_tma(ctx, p, A, B, C, alpha, accum ? 1 : 0);
  launch_gemm_cfg<128, 128, 16, 64, 32, 4>(ctx, p.transA, p.transB, p.d_probs, p.d_tiles_big, (int)p.tiles_big.size(), A, B, C, alpha, accum ? 1 : 0);
  launch_gemm_cfg<64, 64, 16, 32, 32, 4>(ctx, p.transA, p.transB, p.d_probs, p.d_tiles_small, (int)p.tiles_small.size(), A, B, C, alpha, accum ? 1 : 0);
  if (prof) {
    CUDA_OK(cudaEventRecord(rec.b, ctx->stream));
    rec.flops = p.flops;
    rec.cat = 0;
    rec.tiles = (int)(p.tiles_big.size() + p.tiles_small.size() + p.tiles_tma.size());
    ctx->prof_recs.push_back(rec);
  }
  ctx->cnt.gemm_flops += p.flops;
}

__device__ __forceinline__ uint64_t splitmix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
struct BlkDesc { int64_t off; int64_t st[MAXR]; int d[MAXR]; int r; uint64_t key; };
__global__ void fill_random_kernel(double* __restrict__ data, BlkDesc b, int64_t n, uint64_t seed) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    int64_t t = e, addr = b.off;
    for (int k = 0; k < b.r; k++) { addr += (t % b.d[k]) * b.st[k]; t /= b.d[k]; }
    uint64_t u = splitmix64(splitmix64(seed ^ (b.key * 0xD1342543DE82EF95ull)) + (uint64_t)e);
    data[addr] = (double)(u >> 11) * (2.0 / 9007199254740992.0) - 1.0;   // uniform [-1, 1), 53 bits
  }
}
// reference kernel + self test (also the GEMM micro-benchmark used while tuning)
__global__ void ref_gemm_kernel(const double* A, const double* B, double* C, int M, int N, int K, int lda, int ldb, int ldc,
                                bool ta, bool tb) {
  int m = blockIdx.x * blockDim.x + threadIdx.x, n = blockIdx.y;
  if (m >= M || n >= N) return;
  double s = 0.0;
  for (int k = 0; k < K; k++) {
    double a = ta ? A[(int64_t)m * lda + k] : A[(int64_t)k * lda + m];
    double b = tb ? B[(int64_t)k * ldb + n] : B[(int64_t)n * ldb + k];
    s = fma(a, b, s);
  }
  C[(int64_t)n * ldc + m] = s;
}

void gemm_selftest(Ctx* ctx, int M, int N, int K, bool ta, bool tb, int reps, bool verify, double* ms, double* maxerr) {
  auto ev = [](int64_t x) { return (x + 1) & ~int64_t(1); };
  const int lda = (int)ev(ta ? K : M), ldb = (int)ev(tb ? N : K), ldc = (int)ev(M);
  const int64_t na = (int64_t)lda * (ta ? M : K), nb = (int64_t)ldb * (tb ? K : N), nc = (int64_t)ldc * N;
  double* A = (double*)ctx->alloc(na * 8);
  double* B = (double*)ctx->alloc(nb * 8);
  double* C = (double*)ctx->alloc(nc * 8);
  double* R = (double*)ctx->alloc(nc * 8);
  BlkDesc d{};
  d.r = 1; d.st[0] = 1; d.key = 1;
  d.off = 0; d.d[0] = (int)std::min<int64_t>(na, INT32_MAX);
  fill_random_kernel<<<1184, 256, 0, ctx->stream>>>(A, d, na, 11);
  d.key = 2; d.d[0] = (int)std::min<int64_t>(nb, INT32_MAX);
  fill_random_kernel<<<1184, 256, 0, ctx->stream>>>(B, d, nb, 12);
  CUDA_OK(cudaMemsetAsync(C, 0, nc * 8, ctx->stream));
  GemmProblem p{};
  p.a = p.b = p.c = 0; p.M = M; p.N = N; p.K = K; p.lda = lda; p.ldb = ldb; p.ldc = ldc;
  auto plan = plan_gemm_raw(ctx, ta, tb, {p});
  run_gemm(ctx, *plan, A, B, C);
  ctx->sync();
  cudaEvent_t e0, e1;
  CUDA_OK(cudaEventCreate(&e0)); CUDA_OK(cudaEventCreate(&e1));
  CUDA_OK(cudaEventRecord(e0, ctx->stream));
  for (int r = 0; r < reps; r++) run_gemm(ctx, *plan, A, B, C);
  CUDA_OK(cudaEventRecord(e1, ctx->stream));
  CUDA_OK(cudaEventSynchronize(e1));
  float f = 0;
  CUDA_OK(cudaEventElapsedTime(&f, e0, e1));
  *ms = f / std::max(1, reps);
  *maxerr = -1.0;
  if (verify) {
    dim3 grid((M + 127) / 128, N);
    ref_gemm_kernel<<<grid, 128, 0, ctx->stream>>>(A, B, R, M, N, K, lda, ldb, ldc, ta, tb);
    CUDA_OK(cudaGetLastError());
    std::vector<double> hc(nc), hr(nc);
    CUDA_OK(cudaMemcpyAsync(hc.data(), C, nc * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_OK(cudaMemcpyAsync(hr.data(), R, nc * 8, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->sync();
    double e = 0;
    for (int n = 0; n < N; n++)
      for (int m = 0; m < M; m++) e = std::max(e, std::fabs(hc[(int64_t)n * ldc + m] - hr[(int64_t)n * ldc + m]));
    *maxerr = e;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  ctx->free(A); ctx->free(B); ctx->free(C); ctx->free(R);
}

// =================================================================================================
// transform: Y(i, n.., p..) = sum_c sum_k X_c(i, k.., p..) * W_c(k, n)      one warp per output column
// =================================================================================================
// One warp per PASSIVE column of an output block: the X columns it needs are read once and all Na = nd0*nd1
// new-index outputs are produced from registers (up to 4 at a time), lanes run along the stride-1 index i.
// The contributions are flattened to one entry per contracted element, and the X columns of FB entries
// (FB x 4 x 32 doubles per warp) are requested before the first FMA consumes one (a first kernel with one column in
// flight per warp was latency bound at ~50 % of the HBM roofline, profiles/r01b_ncu_hbm.md).
// The per-contribution bookkeeping is staged once per work item in shared memory: lane f computes the base offset of
// X column f (64-bit strides x passive coordinates) and fetches its W row, so the streaming loop over the rows is one
// broadcast LDS for the offset, nn for the weights, four coalesced loads and 4 nn FMAs per contribution, four
// contributions (sixteen loads) in flight.  (ncu of the
// round-1 kernel, profiles/r02q_ncu_apply.md: 375 M warp instructions and 36 M load requests for 4.1 M data loads --
// the XfFlat records, the passive-offset arithmetic and the W loads were redone for every 128-row chunk.)
constexpr int XF_FMAX = 32;                 // contributions staged per pass
__global__ void __launch_bounds__(256)
transform_kernel(const XfGroup* __restrict__ groups, int ngroups, const XfBlock* __restrict__ blocks, const XfFlat* __restrict__ flats,
                    const double* __restrict__ X, double* __restrict__ Y, const double* __restrict__ W, int64_t nitems) {
  __shared__ int64_t s_x[8][XF_FMAX];       // per warp: X offset of (i = 0) of every staged contribution
  __shared__ double s_w[8][XF_FMAX][4];     // per warp: W(k_f, nc .. nc+3)
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t item = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; item < nitems; item += nwarps) {
    // work item = (group, passive column, block of the group), block fastest: neighbouring warps share their X columns
    int lo = 0, hi = ngroups - 1;
    while (lo < hi) {
      int mid = (lo + hi + 1) >> 1;
      if (groups[mid].colstart <= item) lo = mid; else hi = mid - 1;
    }
    const XfGroup g = groups[lo];
    const int64_t tg = item - g.colstart;
    const XfBlock& b = blocks[g.first + (int)(tg % g.nb)];
    int64_t t = tg / g.nb;
    int64_t ybase = b.yoff;
    int pidx[MAXP];
#pragma unroll
    for (int k = 0; k < MAXP; k++) {
      pidx[k] = (int)(t % b.pd[k]); t /= b.pd[k];
      ybase += pidx[k] * b.yps[k];
    }
    const int I = b.I, Na = b.nd0 * b.nd1;
    const int fbeg = b.fbeg, fnum = b.fnum;
    for (int nc = 0; nc < Na; nc += 4) {
      const int nn = min(4, Na - nc);
      double* yp[4];
#pragma unroll
      for (int n = 0; n < 4; n++) {
        const int na = nc + (n < nn ? n : 0);
        yp[n] = Y + ybase + (na % b.nd0) * b.yns[0] + (na / b.nd0) * b.yns[1];
      }
      for (int i0 = 0; i0 < I; i0 += 128) {
        double acc[4][4];
#pragma unroll
        for (int n = 0; n < 4; n++)
#pragma unroll
          for (int u = 0; u < 4; u++) acc[n][u] = 0.0;
        for (int fb = 0; fb < fnum; fb += XF_FMAX) {
          const int fcnt = min(XF_FMAX, fnum - fb);
          // (re)stage when the pass changes; with fnum <= XF_FMAX (the rule) once per nc chunk
          if (i0 == 0 || fnum > XF_FMAX) {
            __syncwarp();
            if (lane < fcnt) {
              const XfFlat& ff = flats[fbeg + fb + lane];
              int64_t xb = ff.xoff;
#pragma unroll
              for (int q = 0; q < MAXP; q++) xb += pidx[q] * ff.xps[q];
              s_x[wid][lane] = xb;
              const double* wr = W + ff.woff + (int64_t)ff.wst * nc;
#pragma unroll
              for (int n = 0; n < 4; n++) s_w[wid][lane][n] = n < nn ? wr[(int64_t)ff.wst * n] : 0.0;
            }
            __syncwarp();
          }
          for (int f = 0; f < fcnt; f += 4) {           // four contributions = sixteen independent loads in flight
            double v[4][4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
              const bool on = f + k < fcnt;
              const double* xp = X + s_x[wid][on ? f + k : f];
#pragma unroll
              for (int u = 0; u < 4; u++) {
                const int i = i0 + lane + 32 * u;
                v[k][u] = (on && i < I) ? xp[i] : 0.0;
              }
            }
#pragma unroll
            for (int k = 0; k < 4; k++) {
              if (f + k < fcnt) {
#pragma unroll
                for (int n = 0; n < 4; n++) {
                  const double w = s_w[wid][f + k][n];
#pragma unroll
                  for (int u = 0; u < 4; u++) acc[n][u] = fma(v[k][u], w, acc[n][u]);
                }
              }
            }
          }
        }
#pragma unroll
        for (int n = 0; n < 4; n++) {
          if (n < nn) {
#pragma unroll
            for (int u = 0; u < 4; u++) {
              const int i = i0 + lane + 32 * u;
              if (i < I) yp[n][i] = acc[n][u];
            }
          }
        }
      }
      __syncwarp();                                    // the staging area is rewritten for the next nc chunk / item
    }
  }
}

// pure relayout (no W): every output column is one input column (or zero); 8 x 32 doubles in flight per warp
__global__ void __launch_bounds__(256)
relayout_kernel(const XfBlock* __restrict__ blocks, int nblocks, const XfFlat* __restrict__ flats,
                const double* __restrict__ X, double* __restrict__ Y, int64_t ncols) {
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t col = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; col < ncols; col += nwarps) {
    int lo = 0, hi = nblocks - 1;
    while (lo < hi) {
      int mid = (lo + hi + 1) >> 1;
      if (blocks[mid].colstart <= col) lo = mid; else hi = mid - 1;
    }
    const XfBlock& b = blocks[lo];
    int64_t t = col - b.colstart;
    int64_t ybase = b.yoff;
    int64_t xb = 0;
    const bool has = b.fnum > 0;
    const XfFlat& ff = flats[has ? b.fbeg : 0];
    if (has) xb = ff.xoff;
#pragma unroll
    for (int k = 0; k < MAXP; k++) {
      const int pi = (int)(t % b.pd[k]); t /= b.pd[k];
      ybase += pi * b.yps[k];
      if (has) xb += pi * ff.xps[k];
    }
    const int I = b.I;
    const double* xp = X + xb;
    double* yp = Y + ybase;
    for (int i0 = 0; i0 < I; i0 += 256) {
      double v[8];
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const int i = i0 + lane + 32 * u;
        v[u] = (has && i < I) ? xp[i] : 0.0;
      }
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const int i = i0 + lane + 32 * u;
        if (i < I) yp[i] = v[u];
      }
    }
  }
}

void run_transform(Ctx* ctx, TransformPlan& p, const double* X, double* Y, const double* W) {
  if (p.ncols == 0) return;
  Ctx::Scope prof_scope(ctx, 1);
  int64_t warps_needed = W ? p.nitems : p.ncols;
  int64_t blocks = std::min<int64_t>((warps_needed + 7) / 8, (int64_t)ctx->num_sms * 32);
  if (!W) {
    relayout_kernel<<<(int)blocks, 256, 0, ctx->stream>>>(p.d_blocks, (int)p.blocks.size(), p.d_flats, X, Y, p.ncols);
  } else {
    transform_kernel<<<(int)blocks, 256, 0, ctx->stream>>>(p.d_groups, (int)p.groups.size(), p.d_blocks, p.d_flats, X, Y, W, p.nitems);
  }
  CUDA_OK(cudaGetLastError());
  ctx->cnt.launches++;
  ctx->cnt.xf_bytes += p.bytes;
  ctx->cnt.xf_flops += p.flops;
}

// =================================================================================================
// flat vector kernels
// =================================================================================================
constexpr int VT = 256;        // threads
constexpr int VMAXB = 148 * 8; // partial slots

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// deterministic two-level reduction: per-block partials, the last block to finish sums them in order
__global__ void __launch_bounds__(VT)
dot_kernel(const double* __restrict__ x, const double* __restrict__ y, int64_t n2, double* __restrict__ partials,
           unsigned int* __restrict__ sync, double* __restrict__ out) {
  const double2* x2 = reinterpret_cast<const double2*>(x);
  const double2* y2 = reinterpret_cast<const double2*>(y);
  double s0 = 0.0, s1 = 0.0;
  const int64_t stride = (int64_t)gridDim.x * VT;
  int64_t i = (int64_t)blockIdx.x * VT + threadIdx.x;
  for (; i + stride < n2; i += 2 * stride) {       // two independent loads in flight
    double2 a = x2[i], b = y2[i];
    double2 c = x2[i + stride], d = y2[i + stride];
    s0 = fma(a.x, b.x, s0); s0 = fma(a.y, b.y, s0);
    s1 = fma(c.x, d.x, s1); s1 = fma(c.y, d.y, s1);
  }
  if (i < n2) { double2 a = x2[i], b = y2[i]; s0 = fma(a.x, b.x, s0); s0 = fma(a.y, b.y, s0); }
  double s = warp_sum(s0 + s1);
  __shared__ double ws[VT / 32];
  __shared__ bool last;
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < VT / 32; w++) t += ws[w];
    partials[blockIdx.x] = t;
    __threadfence();
    unsigned int done = atomicAdd(sync, 1u);
    last = (done == gridDim.x - 1);
  }
  __syncthreads();
  if (last) {
    __threadfence();
    double t = 0.0;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += VT) t += partials[b];   // fixed assignment
    t = warp_sum(t);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
      double r = 0.0;
      for (int w = 0; w < VT / 32; w++) r += ws[w];
      *out = r;
      *sync = 0u;
    }
  }
}

__global__ void __launch_bounds__(VT)
axpy_kernel(double* __restrict__ y, const double* __restrict__ x, int64_t n2, const double* __restrict__ sdev, double a) {
  const double s = sdev ? a * (*sdev) : a;
  double2* y2 = reinterpret_cast<double2*>(y);
  const double2* x2 = reinterpret_cast<const double2*>(x);
  const int64_t stride = (int64_t)gridDim.x * VT;
  for (int64_t i = (int64_t)blockIdx.x * VT + threadIdx.x; i < n2; i += stride) {
    double2 a2 = x2[i], b2 = y2[i];
    b2.x = fma(s, a2.x, b2.x); b2.y = fma(s, a2.y, b2.y);
    y2[i] = b2;
  }
}

// Fused MGS step (KrylovKit ModifiedGramSchmidt2 keeps its sequential order: the coefficient of step q+1 is taken
// AFTER the correction of step q): w += s x followed by out = <y, w_new> in ONE pass over w -- 4ne bytes instead of
// the 5ne of axpy + dot (y == w: the norm^2 of the corrected vector, 3ne instead of 4ne).  s = a * (*sdev) or a.
__global__ void __launch_bounds__(VT)
axpy_dot_kernel(double* w, const double* __restrict__ x, const double* y, int64_t n2, const double* __restrict__ sdev,
                double a, double* __restrict__ partials, unsigned int* __restrict__ sync, double* __restrict__ out) {
  const double s = sdev ? a * (*sdev) : a;
  double2* w2 = reinterpret_cast<double2*>(w);
  const double2* x2 = reinterpret_cast<const double2*>(x);
  const double2* y2 = reinterpret_cast<const double2*>(y);
  const bool self = (y == w);
  double acc = 0.0;
  const int64_t stride = (int64_t)gridDim.x * VT;
  for (int64_t i = (int64_t)blockIdx.x * VT + threadIdx.x; i < n2; i += stride) {
    const double2 xv = x2[i];
    double2 b = w2[i];
    b.x = fma(s, xv.x, b.x); b.y = fma(s, xv.y, b.y);
    w2[i] = b;
    const double2 c = self ? b : y2[i];
    acc = fma(c.x, b.x, acc); acc = fma(c.y, b.y, acc);
  }
  acc = warp_sum(acc);
  __shared__ double ws[VT / 32];
  __shared__ bool last;
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int k = 0; k < VT / 32; k++) t += ws[k];
    partials[blockIdx.x] = t;
    __threadfence();
    unsigned int done = atomicAdd(sync, 1u);
    last = (done == gridDim.x - 1);
  }
  __syncthreads();
  if (last) {
    __threadfence();
    double t = 0.0;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += VT) t += partials[b];
    t = warp_sum(t);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
      double r = 0.0;
      for (int k = 0; k < VT / 32; k++) r += ws[k];
      *out = r;
      *sync = 0u;
    }
  }
}

__global__ void __launch_bounds__(VT)
scale_kernel(double* __restrict__ y, const double* __restrict__ x, int64_t n2, double a) {
  double2* y2 = reinterpret_cast<double2*>(y);
  const double2* x2 = reinterpret_cast<const double2*>(x);
  const int64_t stride = (int64_t)gridDim.x * VT;
  for (int64_t i = (int64_t)blockIdx.x * VT + threadIdx.x; i < n2; i += stride) {
    double2 v = x2[i];
    v.x *= a; v.y *= a;
    y2[i] = v;
  }
}

constexpr int LC_MAX = 32;
struct LincombArgs { const double* x[LC_MAX]; double c[LC_MAX]; int k; };
__global__ void __launch_bounds__(VT)
lincomb_kernel(double* __restrict__ y, LincombArgs a, int64_t n2) {
  double2* y2 = reinterpret_cast<double2*>(y);
  const int64_t stride = (int64_t)gridDim.x * VT;
  for (int64_t i = (int64_t)blockIdx.x * VT + threadIdx.x; i < n2; i += stride) {
    double2 acc = make_double2(0.0, 0.0);
    for (int j = 0; j < a.k; j++) {
      double2 v = reinterpret_cast<const double2*>(a.x[j])[i];
      acc.x = fma(a.c[j], v.x, acc.x); acc.y = fma(a.c[j], v.y, acc.y);
    }
    y2[i] = acc;
  }
}

static inline int vec_grid(Ctx* ctx, int64_t n2) {
  int64_t b = (n2 + VT - 1) / VT;
  return (int)std::max<int64_t>(1, std::min<int64_t>(b, (int64_t)ctx->num_sms * 8));
}

void vec_dot(Ctx* ctx, const double* x, const double* y, int64_t n, int slot) {
  Ctx::Scope prof_scope(ctx, 2);
  TNL_CHECK(n % 2 == 0, "padded vector length must be even");
  dot_kernel<<<vec_grid(ctx, n / 2), VT, 0, ctx->stream>>>(x, y, n / 2, ctx->d_partials, ctx->d_sync, ctx->d_scalars + slot);
  CUDA_OK(cudaGetLastError());
  ctx->cnt.launches++;
  ctx->cnt.vec_bytes += (x == y ? 1.0 : 2.0) * n * 8.0;
}
void vec_axpy_dev(Ctx* ctx, double* y, const double* x, int64_t n, int slot, double sign) {
  Ctx::Scope prof_scope(ctx, 2);
  axpy_kernel<<<vec_grid(ctx, n / 2), VT, 0, ctx->stream>>>(y, x, n / 2, ctx->d_scalars + slot, sign);
  CUDA_OK(cudaGetLastError());
  ctx->cnt.launches++;
  ctx->cnt.vec_bytes += 3.0 * n * 8.0;
}
void vec_axpy(Ctx* ctx, double* y, const double* x, int64_t n, double a) {
  Ctx::Scope prof_scope(ctx, 2);
  axpy_kernel<<<vec_grid(ctx, n / 2), VT, 0, ctx->stream>>>(y, x, n / 2, nullptr, a);
  CUDA_OK(cudaGetLastError());
  ctx->cnt.launches++;
  ctx->cnt.vec_bytes += 3.0 * n * 8.0;
}
// w += (slot_in >= 0 ? a * s[slot_in] : a) * x ;  s[slot_out] = <y, w>   (y may be w)
void vec_axpy_dot(Ctx* ctx, double* w, const double* x, int64_t n, int slot_in, double a, const double* y, int slot_out) {
  Ctx::Scope prof_scope(ctx, 2);
  TNL_CHECK(n % 2 == 0, "padded vector length must be even");
  axpy_dot_kernel<<<vec_grid(ctx, n / 2), VT, 0, ctx->stream>>>(w, x, y, n / 2, slot_in >= 0 ? ctx->d_scalars + slot_in : nullptr, a,
                                                               ctx->d_partials, ctx->d_sync, ctx->d_scalars + slot_out);
  CUDA_OK(cudaGetLastError());
  ctx->cnt.launches++;
  ctx->cnt.vec_bytes += (y == w ? 3.0 : 4.0) * n * 8.0;
}
void vec_scale(Ctx* ctx, double* y, int64_t n, double a) { vec_scale_to(ctx, y, y, n, a); }
void vec_scale_to(Ctx* ctx, double* y, const double* x, int64_t n, double a) {
  Ctx::Scope prof_scope(ctx, 2);
  scale_kernel<<<vec_grid(ctx, n / 2), VT, 0, ctx->stream>>>(y, x, n / 2, a);
  CUDA_OK(cudaGetLastError());
  ctx->cnt.launches++;
  ctx->cnt.vec_bytes += 2.0 * n * 8.0;
}
void vec_copy(Ctx* ctx, double* y, const double* x, int64_t n) {
  CUDA_OK(cudaMemcpyAsync(y, x, n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
}
void vec_lincomb(Ctx* ctx, double* y, const double* const* xs, const double* coef, int k, int64_t n) {
  Ctx::Scope prof_scope(ctx, 2);
  TNL_CHECK(k <= LC_MAX, "too many vectors in linear combination");
  LincombArgs a;
  a.k = k;
  for (int j = 0; j < k; j++) { a.x[j] = xs[j]; a.c[j] = coef[j]; }
  lincomb_kernel<<<vec_grid(ctx, n / 2), VT, 0, ctx->stream>>>(y, a, n / 2);
  CUDA_OK(cudaGetLastError());
  ctx->cnt.launches++;
  ctx->cnt.vec_bytes += (k + 1.0) * n * 8.0;
}
// ---- planar complex vectors: [re plane | im plane], each `n` doubles -----------------------------
// <x, y> = sum conj(x) y : out[0] = re, out[1] = im; same deterministic two-level reduction as dot_kernel
__global__ void __launch_bounds__(VT)
cdot_kernel(const double* __restrict__ x, const double* __restrict__ y, int64_t n2, int64_t plane2,
            double* __restrict__ partials, unsigned int* __restrict__ sync, double* __restrict__ out) {
  const double2* xr = reinterpret_cast<const double2*>(x);
  const double2* yr = reinterpret_cast<const double2*>(y);
  const double2* xi = xr + plane2;
  const double2* yi = yr + plane2;
  double sr = 0.0, si = 0.0;
  const int64_t stride = (int64_t)gridDim.x * VT;
  for (int64_t i = (int64_t)blockIdx.x * VT + threadIdx.x; i < n2; i += stride) {
    const double2 a = xr[i], b = xi[i], c = yr[i], d = yi[i];
    sr = fma(a.x, c.x, sr); sr = fma(a.y, c.y, sr); sr = fma(b.x, d.x, sr); sr = fma(b.y, d.y, sr);
    si = fma(a.x, d.x, si); si = fma(a.y, d.y, si); si = fma(-b.x, c.x, si); si = fma(-b.y, c.y, si);
  }
  sr = warp_sum(sr);
  si = warp_sum(si);
  __shared__ double ws[2][VT / 32];
  __shared__ bool last;
  if ((threadIdx.x & 31) == 0) { ws[0][threadIdx.x >> 5] = sr; ws[1][threadIdx.x >> 5] = si; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t0 = 0.0, t1 = 0.0;
    for (int w = 0; w < VT / 32; w++) { t0 += ws[0][w]; t1 += ws[1][w]; }
    partials[2 * blockIdx.x] = t0;
    partials[2 * blockIdx.x + 1] = t1;
    __threadfence();
    unsigned int done = atomicAdd(sync, 1u);
    last = (done == gridDim.x - 1);
  }
  __syncthreads();
  if (last) {
    __threadfence();
    double t0 = 0.0, t1 = 0.0;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += VT) { t0 += partials[2 * b]; t1 += partials[2 * b + 1]; }
    t0 = warp_sum(t0);
    t1 = warp_sum(t1);
    if ((threadIdx.x & 31) == 0) { ws[0][threadIdx.x >> 5] = t0; ws[1][threadIdx.x >> 5] = t1; }
    __syncthreads();
    if (threadIdx.x == 0) {
      double r0 = 0.0, r1 = 0.0;
      for (int w = 0; w < VT / 32; w++) { r0 += ws[0][w]; r1 += ws[1][w]; }
      out[0] = r0;
      out[1] = r1;
      *sync = 0u;
    }
  }
}

// y += s x with s = sign * (sdev[0] + i sdev[1]) or (ar + i ai)
__global__ void __launch_bounds__(VT)
caxpy_kernel(double* __restrict__ y, const double* __restrict__ x, int64_t n2, int64_t plane2,
             const double* __restrict__ sdev, double ar, double ai) {
  const double sr = sdev ? ar * sdev[0] : ar;
  const double si = sdev ? ar * sdev[1] : ai;
  double2* yr = reinterpret_cast<double2*>(y);
  double2* yi = yr + plane2;
  const double2* xr = reinterpret_cast<const double2*>(x);
  const double2* xi = xr + plane2;
  const int64_t stride = (int64_t)gridDim.x * VT;
  for (int64_t i = (int64_t)blockIdx.x * VT + threadIdx.x; i < n2; i += stride) {
    const double2 a = xr[i], b = xi[i];
    double2 c = yr[i], d = yi[i];
    c.x = fma(sr, a.x, c.x); c.x = fma(-si, b.x, c.x); c.y = fma(sr, a.y, c.y); c.y = fma(-si, b.y, c.y);
    d.x = fma(sr, b.x, d.x); d.x = fma(si, a.x, d.x); d.y = fma(sr, b.y, d.y); d.y = fma(si, a.y, d.y);
    yr[i] = c;
    yi[i] = d;
  }
}

// complex fused MGS step on planar vectors: w += s x, out = <y, w_new> = sum conj(y) w  (out[0] = re, out[1] = im)
__global__ void __launch_bounds__(VT)
caxpy_cdot_kernel(double* w, const double* __restrict__ x, const double* y, int64_t n2, int64_t plane2,
                  const double* __restrict__ sdev, double ar, double ai, double* __restrict__ partials,
                  unsigned int* __restrict__ sync, double* __restrict__ out) {
  const double sr = sdev ? ar * sdev[0] : ar;
  const double si = sdev ? ar * sdev[1] : ai;
  double2* wr = reinterpret_cast<double2*>(w);
  double2* wi = wr + plane2;
  const double2* xr = reinterpret_cast<const double2*>(x);
  const double2* xi = xr + plane2;
  const double2* yr = reinterpret_cast<const double2*>(y);
  const double2* yi = yr + plane2;
  const bool self = (y == w);
  double accr = 0.0, acci = 0.0;
  const int64_t stride = (int64_t)gridDim.x * VT;
  for (int64_t i = (int64_t)blockIdx.x * VT + threadIdx.x; i < n2; i += stride) {
    const double2 a = xr[i], b = xi[i];
    double2 c = wr[i], d = wi[i];
    c.x = fma(sr, a.x, c.x); c.x = fma(-si, b.x, c.x); c.y = fma(sr, a.y, c.y); c.y = fma(-si, b.y, c.y);
    d.x = fma(sr, b.x, d.x); d.x = fma(si, a.x, d.x); d.y = fma(sr, b.y, d.y); d.y = fma(si, a.y, d.y);
    wr[i] = c;
    wi[i] = d;
    const double2 p = self ? c : yr[i], q = self ? d : yi[i];
    accr = fma(p.x, c.x, accr); accr = fma(p.y, c.y, accr); accr = fma(q.x, d.x, accr); accr = fma(q.y, d.y, accr);
    acci = fma(p.x, d.x, acci); acci = fma(p.y, d.y, acci); acci = fma(-q.x, c.x, acci); acci = fma(-q.y, c.y, acci);
  }
  accr = warp_sum(accr);
  acci = warp_sum(acci);
  __shared__ double ws[2][VT / 32];
  __shared__ bool last;
  if ((threadIdx.x & 31) == 0) { ws[0][threadIdx.x >> 5] = accr; ws[1][threadIdx.x >> 5] = acci; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t0 = 0.0, t1 = 0.0;
    for (int k = 0; k < VT / 32; k++) { t0 += ws[0][k]; t1 += ws[1][k]; }
    partials[2 * blockIdx.x] = t0;
    partials[2 * blockIdx.x + 1] = t1;
    __threadfence();
    unsigned int done = atomicAdd(sync, 1u);
    last = (done == gridDim.x - 1);
  }
  __syncthreads();
  if (last) {
    __threadfence();
    double t0 = 0.0, t1 = 0.0;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += VT) { t0 += partials[2 * b]; t1 += partials[2 * b + 1]; }
    t0 = warp_sum(t0);
    t1 = warp_sum(t1);
    if ((threadIdx.x & 31) == 0) { ws[0][threadIdx.x >> 5] = t0; ws[1][threadIdx.x >> 5] = t1; }
    __syncthreads();
    if (threadIdx.x == 0) {
      double r0 = 0.0, r1 = 0.0;
      for (int k = 0; k < VT / 32; k++) { r0 += ws[0][k]; r1 += ws[1][k]; }
      out[0] = r0;
      out[1] = r1;
      *sync = 0u;
    }
  }
}

struct CLincombArgs { const double* x[LC_MAX]; double cr[LC_MAX]; double ci[LC_MAX]; int k; };
__global__ void __launch_bounds__(VT)
clincomb_kernel(double* __restrict__ y, CLincombArgs a, int64_t n2, int64_t plane2) {
  double2* yr = reinterpret_cast<double2*>(y);
  double2* yi = yr + plane2;
  const int64_t stride = (int64_t)gridDim.x * VT;
  for (int64_t i = (int64_t)blockIdx.x * VT + threadIdx.x; i < n2; i += stride) {
    double2 r = make_double2(0.0, 0.0), m = make_double2(0.0, 0.0);
    for (int j = 0; j < a.k; j++) {
      const double2 vr = reinterpret_cast<const double2*>(a.x[j])[i];
      const double2 vi = reinterpret_cast<const double2*>(a.x[j])[i + plane2];
      r.x = fma(a.cr[j], vr.x, r.x); r.x = fma(-a.ci[j], vi.x, r.x);
      r.y = fma(a.cr[j], vr.y, r.y); r.y = fma(-a.ci[j], vi.y, r.y);
      m.x = fma(a.cr[j], vi.x, m.x); m.x = fma(a.ci[j], vr.x, m.x);
      m.y = fma(a.cr[j], vi.y, m.y); m.y = fma(a.ci[j], vr.y, m.y);
    }
    yr[i] = r;
    yi[i] = m;
  }
}

void vec_cdot(Ctx* ctx, const double* x, const double* y, int64_t n, int slot) {
  Ctx::Scope prof_scope(ctx, 2);
  TNL_CHECK(n % 2 == 0, "padded vector length must be even");
  cdot_kernel<<<vec_grid(ctx, n / 2), VT, 0, ctx->stream>>>(x, y, n / 2, n / 2, ctx->d_partials, ctx->d_sync, ctx->d_scalars + slot);
  CUDA_OK(cudaGetLastError());
  ctx->cnt.launches++;
  ctx->cnt.vec_bytes += (x == y ? 1.0 : 2.0) * n * 16.0;
}
void vec_caxpy_dev(Ctx* ctx, double* y, const double* x, int64_t n, int slot, double sign) {
  Ctx::Scope prof_scope(ctx, 2);
  caxpy_kernel<<<vec_grid(ctx, n / 2), VT, 0, ctx->stream>>>(y, x, n / 2, n / 2, ctx->d_scalars + slot, sign, 0.0);
  CUDA_OK(cudaGetLastError());
  ctx->cnt.launches++;
  ctx->cnt.vec_bytes += 3.0 * n * 16.0;
}
void vec_caxpy(Ctx* ctx, double* y, const double* x, int64_t n, double ar, double ai) {
  Ctx::Scope prof_scope(ctx, 2);
  caxpy_kernel<<<vec_grid(ctx, n / 2), VT, 0, ctx->stream>>>(y, x, n / 2, n / 2, nullptr, ar, ai);
  CUDA_OK(cudaGetLastError());
  ctx->cnt.launches++;
  ctx->cnt.vec_bytes += 3.0 * n * 16.0;
}
// planar complex: w += (slot_in >= 0 ? ar * (s[slot_in] + i s[slot_in+1]) : ar + i ai) * x ; s[slot_out, +1] = <y, w>
void vec_caxpy_cdot(Ctx* ctx, double* w, const double* x, int64_t n, int slot_in, double ar, double ai, const double* y,
                    int slot_out) {
  Ctx::Scope prof_scope(ctx, 2);
  TNL_CHECK(n % 2 == 0, "padded vector length must be even");
  caxpy_cdot_kernel<<<vec_grid(ctx, n / 2), VT, 0, ctx->stream>>>(w, x, y, n / 2, n / 2,
                                                                 slot_in >= 0 ? ctx->d_scalars + slot_in : nullptr, ar, ai,
                                                                 ctx->d_partials, ctx->d_sync, ctx->d_scalars + slot_out);
  CUDA_OK(cudaGetLastError());
  ctx->cnt.launches++;
  ctx->cnt.vec_bytes += (y == w ? 3.0 : 4.0) * n * 16.0;
}
void vec_clincomb(Ctx* ctx, double* y, const double* const* xs, const double* cr, const double* ci, int k, int64_t n) {
  Ctx::Scope prof_scope(ctx, 2);
  TNL_CHECK(k <= LC_MAX, "too many vectors in linear combination");
  CLincombArgs a;
  a.k = k;
  for (int j = 0; j < k; j++) { a.x[j] = xs[j]; a.cr[j] = cr[j]; a.ci[j] = ci[j]; }
  clincomb_kernel<<<vec_grid(ctx, n / 2), VT, 0, ctx->stream>>>(y, a, n / 2, n / 2);
  CUDA_OK(cudaGetLastError());
  ctx->cnt.launches++;
  ctx->cnt.vec_bytes += (k + 1.0) * n * 16.0;
}

// C = op(A) op(B) for any mix of real and planar-complex operands (Ai / Bi == nullptr: real operand);
// conjA / conjB conjugate a complex operand.  Four (two, one) launches of the real grouped DGEMM:
//   Cr = Ar Br - sA sB Ai Bi ,  Ci = sB Ar Bi + sA Ai Br      (sX = -1 if conjX)
void cgemm(Ctx* ctx, GemmPlan& p, const double* Ar, const double* Ai, bool conjA, const double* Br, const double* Bi,
           bool conjB, double* Cr, double* Ci) {
  const double sA = conjA ? -1.0 : 1.0, sB = conjB ? -1.0 : 1.0;
  if (Ai && Bi && Ci && ctx->use_tma && ctx->dual_gemm) {
    // complex x complex: one dual-source launch per plane
    run_gemm_dual(ctx, p, Ar, Br, 1.0, Ai, Bi, -sA * sB, Cr);
    run_gemm_dual(ctx, p, Ar, Bi, sB, Ai, Br, sA, Ci);
    return;
  }
  run_gemm(ctx, p, Ar, Br, Cr, 1.0, false);
  if (Ai && Bi) run_gemm(ctx, p, Ai, Bi, Cr, -sA * sB, true);
  if (!Ci) { TNL_CHECK(!Ai && !Bi, "complex product needs a complex result"); return; }
  bool first = true;
  if (Bi) { run_gemm(ctx, p, Ar, Bi, Ci, sB, false); first = false; }
  if (Ai) { run_gemm(ctx, p, Ai, Br, Ci, sA, !first); first = false; }
  if (first) {   // both operands real: the imaginary plane of the result is zero on every problem
    for (const GemmProblem& q : p.probs)
      for (int n = 0; n < q.N; n++) CUDA_OK(cudaMemsetAsync(Ci + q.c + (int64_t)n * q.ldc, 0, (size_t)q.M * sizeof(double), ctx->stream));
  }
}
void cgemm(Ctx* ctx, GemmPlan& p, const Tensor& A, bool conjA, const Tensor& B, bool conjB, Tensor& C) {
  TNL_CHECK(C.cplx == (A.cplx || B.cplx), "result of a contraction is complex iff an operand is");
  cgemm(ctx, p, A.d, A.cplx ? A.im() : nullptr, conjA, B.d, B.cplx ? B.im() : nullptr, conjB, C.d, C.cplx ? C.im() : nullptr);
}
// transform of both planes (W real)
void run_transform_c(Ctx* ctx, TransformPlan& p, const Tensor& X, Tensor& Y, const double* W) {
  TNL_CHECK(X.cplx == Y.cplx, "transform keeps the element type");
  run_transform(ctx, p, X.d, Y.d, W);
  if (X.cplx) run_transform(ctx, p, X.im(), Y.im(), W);
}

void fetch_scalars(Ctx* ctx, int n) {
  CUDA_OK(cudaMemcpyAsync(ctx->h_scalars, ctx->d_scalars, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_OK(cudaStreamSynchronize(ctx->stream));
}

// =================================================================================================
// misc elementwise kernels
// =================================================================================================
void fill_random(Ctx* ctx, Tensor& t, uint64_t seed) {
  for (auto& b : t.blocks) {
    BlkDesc d;
    d.off = b.off; d.r = t.rank(); d.key = Tensor::key(b.c, t.rank());
    int64_t n = 1;
    for (int k = 0; k < t.rank(); k++) { d.st[k] = b.st[k]; d.d[k] = b.d[k]; n *= b.d[k]; }
    int grid = (int)std::min<int64_t>((n + 255) / 256, 1184);
    fill_random_kernel<<<grid, 256, 0, ctx->stream>>>(t.d, d, n, seed);
    CUDA_OK(cudaGetLastError());
    ctx->cnt.launches++;
  }
}

// T(..., i_k, ...) *= w[i_k] on one block: `w` points at the values of the block's sector of index k
__global__ void scale_index_kernel(double* __restrict__ data, BlkDesc b, int64_t n, int k, const double* __restrict__ w) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    int64_t t = e, addr = b.off;
    int ik = 0;
    for (int q = 0; q < b.r; q++) {
      const int i = (int)(t % b.d[q]);
      t /= b.d[q];
      addr += i * b.st[q];
      if (q == k) ik = i;
    }
    data[addr] *= w[ik];
  }
}
// diag(w) applied on index `which` of every block (w: one value per element of the index, sector after sector)
void scale_index(Ctx* ctx, Tensor& t, int which, const double* w_dev) {
  std::vector<int64_t> soff(t.inds[which].nsect() + 1, 0);
  for (int s = 0; s < t.inds[which].nsect(); s++) soff[s + 1] = soff[s] + t.inds[which].dims[s];
  for (int pl = 0; pl < t.planes(); pl++)
    for (auto& b : t.blocks) {
      BlkDesc d;
      d.off = b.off; d.r = t.rank(); d.key = 0;
      int64_t n = 1;
      for (int k = 0; k < t.rank(); k++) { d.st[k] = b.st[k]; d.d[k] = b.d[k]; n *= b.d[k]; }
      if (n == 0) continue;
      int grid = (int)std::min<int64_t>((n + 255) / 256, 1184);
      scale_index_kernel<<<grid, 256, 0, ctx->stream>>>(t.d + pl * t.nelem, d, n, which, w_dev + soff[b.c[which]]);
      CUDA_OK(cudaGetLastError());
      ctx->cnt.launches++;
    }
}

__global__ void scale_rc_kernel(double* __restrict__ A, int64_t ld, int64_t R, int64_t C, const double* __restrict__ s, bool rows) {
  int64_t n = R * C;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = e % R, c = e / R;
    A[c * ld + r] *= rows ? s[r] : s[c];
  }
}
void scale_rows_or_cols(Ctx* ctx, double* A, int64_t ld, int64_t R, int64_t C, const double* s, bool rows) {
  if (R * C == 0) return;
  int grid = (int)std::min<int64_t>((R * C + 255) / 256, 1184);
  scale_rc_kernel<<<grid, 256, 0, ctx->stream>>>(A, ld, R, C, s, rows);
  CUDA_OK(cudaGetLastError());
  ctx->cnt.launches++;
}

}  // namespace tnl

namespace tnl {
__global__ void transpose_kernel(double* __restrict__ dst, int64_t ldd, const double* __restrict__ src, int64_t lds, int64_t R, int64_t C) {
  __shared__ double tile[32][33];
  int64_t r0 = (int64_t)blockIdx.x * 32, c0 = (int64_t)blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int64_t r = r0 + threadIdx.x, c = c0 + j;
    if (r < R && c < C) tile[j][threadIdx.x] = src[c * lds + r];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int64_t c = c0 + threadIdx.x, r = r0 + j;
    if (r < R && c < C) dst[r * ldd + c] = tile[threadIdx.x][j];
  }
}
void transpose(Ctx* ctx, double* dst, int64_t ldd, const double* src, int64_t lds, int64_t R, int64_t C) {
  if (R * C == 0) return;
  dim3 grid((unsigned)((R + 31) / 32), (unsigned)((C + 31) / 32)), block(32, 8);
  transpose_kernel<<<grid, block, 0, ctx->stream>>>(dst, ldd, src, lds, R, C);
  CUDA_OK(cudaGetLastError());
  ctx->cnt.launches++;
}
void copy2d(Ctx* ctx, double* dst, int64_t ldd, const double* src, int64_t lds, int64_t R, int64_t C) {
  if (R * C == 0) return;
  CUDA_OK(cudaMemcpy2DAsync(dst, ldd * sizeof(double), src, lds * sizeof(double), R * sizeof(double), C,
                            cudaMemcpyDeviceToDevice, ctx->stream));
}
}  // namespace tnl

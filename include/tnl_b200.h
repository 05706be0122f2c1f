/* tnl_b200 -- C ABI of the B200-native DMRG hot path (drop-in boundary for TenNetLib.jl).
 *
 * Every entry point returns 0 on success, non-zero on error (2 = invalid argument / index
 * mismatch, 3 = CUDA, 4 = cuSOLVER, 1 = other); tnl_last_error() gives the message, which the
 * Julia shim rethrows as `error(msg)` (reference convention: src/mps/update_site.jl:39-41,247-250,
 * src/mps/projcouplingmodel.jl:367-381).  Handles are opaque pointers owned by the library; host
 * buffers passed in are only read/written during the call.  Plain pointers and sizes only -- no
 * torch / CUDA types.  Positions are 1-based like the reference.
 *
 * Host tensor format = NDTensors BlockSparse: flat `data`, per block the
 * 0-based sector number of every index (`coords`, block-major) and the 0-based element `offset`;
 * each block is dense column-major.  QN indices are described by tnl_index_t.
 */
#ifndef TNL_B200_H
#define TNL_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tnl_ctx_s* tnl_ctx_t;
typedef struct tnl_tensor_s* tnl_tensor_t;
typedef struct tnl_env_s* tnl_env_t;
typedef struct tnl_sumop_s* tnl_sumop_t;

/* One QN index: ITensors `Index{Vector{Pair{QN,Int}}}` -- sector dims, nq charges per sector, arrow. */
typedef struct {
  int32_t nsect;
  int32_t dir;          /* +1 = Out, -1 = In */
  const int32_t* dims;  /* [nsect] */
  const int32_t* qns;   /* [nsect * nq] */
} tnl_index_t;

/* ---- context ------------------------------------------------------------------------------- */
int tnl_ctx_create(int device, tnl_ctx_t* ctx);
int tnl_ctx_destroy(tnl_ctx_t ctx);
const char* tnl_last_error(tnl_ctx_t ctx); /* ctx may be NULL: last error of the calling thread */
/* out[0..7] = gemm_flops, transform_flops, vec_bytes, transform_bytes, kernel launches, gemm launches,
 * H_eff applies, all-reduce bytes.  Algorithmic counts. */
int tnl_get_counters(tnl_ctx_t ctx, double* out8);
int tnl_reset_counters(tnl_ctx_t ctx);
int tnl_ctx_sync(tnl_ctx_t ctx);
/* pre-grow the context's device memory pool by `bytes` (kept mapped for the lifetime of the process) so that the
 * per-bond allocations of a sweep never wait for the driver to map new physical memory */
int tnl_ctx_reserve(tnl_ctx_t ctx, int64_t bytes);
/* CUDA-event timing on the library's own stream (bench.py times kernels here, not on torch's stream). */
int tnl_timer_start(tnl_ctx_t ctx, int32_t slot); /* slot 0..3 */
int tnl_timer_stop(tnl_ctx_t ctx, int32_t slot, double* milliseconds);
/* kernel self test / micro-benchmark: random C = op(A) op(B) with the grouped DGEMM kernel; returns the average
 * launch time and (verify != 0) the max abs deviation from a naive FP64 reference kernel */
int tnl_gemm_selftest(tnl_ctx_t ctx, int32_t M, int32_t N, int32_t K, int32_t transA, int32_t transB, int32_t reps,
                      int32_t verify, double* ms, double* maxerr);
/* per-launch CUDA-event profile of the grouped DGEMM launches (those with 128x128 tiles): enable, run, read
 * (read synchronises, sums the launch durations and algorithmic flops, and clears the records) */
int tnl_profile_gemm(tnl_ctx_t ctx, int32_t enable);
int tnl_profile_read(tnl_ctx_t ctx, double* total_ms, int64_t* launches, double* flops, double* max_tflops);
/* by-category device time of the records consumed by the last tnl_profile_read: out[0..3] = ms of grouped DGEMM,
 * transform, Krylov-vector and collective launches, out[4..7] = their launch counts */
int tnl_profile_categories(tnl_ctx_t ctx, double* out8);
/* collectives consumed by the last tnl_profile_read, by kind: out[0..5] = ms of scalar all-reduces (Krylov inner
 * products), NCCL reduce-scatters (H_eff partial sums), all-gathers (result vectors, environment slices), large
 * all-reduces (truncation factors), and -- fused GEMM -> reduce-scatter over peer memory -- the wait for "staging slots
 * consumed" and the ordered sum over the slots (which includes the wait for the peers' tiles); out[6..11] = call counts */
int tnl_profile_collectives(tnl_ctx_t ctx, double* out12);

/* ---- multi-GPU: one process per GPU, sharded H_eff apply ----------------
 * rank 0 calls tnl_comm_unique_id, the 128 bytes are broadcast by the host (torch.distributed / MPI), every rank
 * calls tnl_comm_init.  Afterwards tnl_heff_apply / tnl_eigsolve_lanczos shard the two-site apply over the right
 * link (each rank owns tnl_shard_range of every sector).  The partial results are combined inside the last GEMM of the
 * apply: its epilogue stores every tile into the owner rank's staging slot over NVLink (CUDA-IPC peer memory) and a
 * small kernel sums the slots in rank order; if peer memory cannot be set up (or TNL_FUSED_RS=0) the same exchange is a
 * pack pass + ncclReduceScatter.  Krylov vectors stay sharded; inner products are scalar all-reduces, the result
 * vector one all-gather; the environment update shards over a link index and all-gathers the slices.
 * All ranks must hold identical replicated state and issue the same calls. */
int tnl_comm_unique_id(char* out128);
int tnl_comm_init(tnl_ctx_t ctx, const char* uid128, int32_t rank, int32_t world);
int tnl_comm_destroy(tnl_ctx_t ctx);
/* enable = 0: every rank computes the whole replicated problem without collectives (the unsharded result the
 * sharded path is checked against, bench.py `parity`); enable = 1 (default after tnl_comm_init): sharded. */
int tnl_comm_set_sharding(tnl_ctx_t ctx, int32_t enable);
int tnl_shard_range(int32_t dim, int32_t world, int32_t sector, int32_t rank, int32_t* start, int32_t* count);
/* collective micro-benchmark on the library stream: kind 0 = all-reduce of n doubles, 1 = reduce-scatter
 * (n per rank), 2 = all-gather (n per rank); returns the average milliseconds per call */
int tnl_comm_bench(tnl_ctx_t ctx, int64_t n, int32_t reps, int32_t kind, double* ms);

/* ---- tensors: ITensor <-> device (replaces nothing in the reference; it is the marshalling the shim needs) */
int tnl_tensor_import(tnl_ctx_t ctx, int32_t rank, int32_t nq, const tnl_index_t* inds, int64_t nblocks,
                      const int32_t* coords, const int64_t* offsets, const double* data, int32_t nrow,
                      tnl_tensor_t* out);
/* ComplexF64 tensors (TDVP in real time, SURVEY.md section 8 rows a1/a5/a7 "C128"): the host hands over / receives
 * NDTensors' interleaved (re, im) storage (offsets count complex elements); on the device the tensor is PLANAR --
 * two real planes in the same charge-fused layout -- so every contraction stays on the FP64 DMMA kernels.
 * tnl_tensor_export writes 2 * nelem doubles for a complex tensor.  Site operators (MPO / CouplingModel tensors)
 * are real.  eig_solver, exp_solver, the noise term, penalties and the gauge moves accept complex states.  Complex
 * vectors shard plane by plane (pack + ncclReduceScatter per plane; the fused GEMM epilogue serves Float64). */
int tnl_tensor_import_c128(tnl_ctx_t ctx, int32_t rank, int32_t nq, const tnl_index_t* inds, int64_t nblocks,
                           const int32_t* coords, const int64_t* offsets, const double* data_re_im, int32_t nrow,
                           tnl_tensor_t* out);
int tnl_tensor_is_complex(tnl_tensor_t t, int32_t* out);
int tnl_tensor_promote(tnl_tensor_t t); /* real -> complex with zero imaginary part, in place */
int tnl_tensor_create(tnl_ctx_t ctx, int32_t rank, int32_t nq, const tnl_index_t* inds, int32_t nrow,
                      tnl_tensor_t* out); /* all symmetry-allowed blocks, zero */
int tnl_tensor_free(tnl_tensor_t t);
int tnl_tensor_copy(tnl_tensor_t t, tnl_tensor_t* out);
int tnl_tensor_rank(tnl_tensor_t t, int32_t* rank, int32_t* nq);
int tnl_tensor_nrow(tnl_tensor_t t, int32_t* nrow); /* number of leading indices forming the row group of the layout */
int tnl_tensor_index(tnl_tensor_t t, int32_t which, int32_t* nsect, int32_t* dir, int32_t* dims, int32_t* qns,
                     int32_t cap);
int tnl_tensor_export_size(tnl_tensor_t t, int64_t* nblocks, int64_t* nelem);
int tnl_tensor_export(tnl_tensor_t t, int32_t* coords, int64_t* offsets, double* data);
int tnl_tensor_fill_random(tnl_tensor_t t, uint64_t seed); /* uniform [-1,1), counter based */
/* T(..., i, ...) *= values[i] along index `which` (values: dim(index) doubles, sector after sector) -- ITensor
 * contraction with a diagonal tensor, as the reference's `S * V` (src/mps/update_site.jl:172) */
int tnl_tensor_scale_index(tnl_tensor_t t, int32_t which, const double* values);

/* ---- Krylov vector interface (VectorInterface inner / norm / scale!! / add!! on ITensors) ---- */
int tnl_vec_dot(tnl_tensor_t x, tnl_tensor_t y, double* out); /* complex tensors: the real part */
int tnl_vec_dot_c(tnl_tensor_t x, tnl_tensor_t y, double* re, double* im); /* <x,y> = sum conj(x) y */
int tnl_vec_norm(tnl_tensor_t x, double* out);
int tnl_vec_scale(tnl_tensor_t x, double a);
int tnl_vec_axpy(tnl_tensor_t y, tnl_tensor_t x, double a); /* y += a*x */

/* ---- StateEnvs{ProjMPO}: src/mps/state_envs.jl:18-27,54-60 ---------------------------------- */
int tnl_env_create(tnl_ctx_t ctx, int32_t nsites, tnl_env_t* env);
int tnl_env_destroy(tnl_env_t env);
/* MPO tensor W_site(wl, s', s, wr) in host format (StateEnvs(psi, H::MPO), state_envs.jl:54-60) */
int tnl_env_set_site_op(tnl_env_t env, int32_t site, int32_t nq, const tnl_index_t* inds4, int64_t nblocks,
                        const int32_t* coords, const int64_t* offsets, const double* data);
/* StateEnvs(psi, Hs::Vector{MPO}) = ProjMPOSum2 (src/mps/state_envs.jl:63-70, src/mps/projmposum2.jl:15-145):
 * site operator of MPO number `term` (0-based; term 0 is what tnl_env_set_site_op sets).  product and noiseterm
 * become sums over the terms; every term keeps its own environments over the shared state. */
int tnl_env_set_site_op_term(tnl_env_t env, int32_t term, int32_t site, int32_t nq, const tnl_index_t* inds4,
                             int64_t nblocks, const int32_t* coords, const int64_t* offsets, const double* data);
/* updateH!(sysenv::StateEnvs{ProjMPO}, H; recalcEnv = false) (src/mps/state_envs.jl:181-208): replace W_site and keep
 * every cached environment (time-dependent Hamiltonians, docs/src/mps/example_tdvp.md:134-136).  The MPO links must
 * keep their spaces and the left watermark must be 0 (orthogonality centre at site 1), as the reference asserts.
 * recalcEnv = true is a fresh environment over the same state tensors (tnl_env_create + tnl_env_set_state). */
int tnl_env_update_site_op(tnl_env_t env, int32_t site, int32_t nq, const tnl_index_t* inds4, int64_t nblocks,
                           const int32_t* coords, const int64_t* offsets, const double* data);
/* StateEnvs(psi, H::CouplingModel) = ProjCouplingModel (src/mps/state_envs.jl:73-79, src/mps/projcouplingmodel.jl,
 * src/base/couplingmodel.jl:14-17): tensor of term `id` (the key of the IDTensors dictionary, any non-negative
 * integer) on `site`, handed over as W(wl, s', s, wr) where wl / wr are the OpLinks shared with the term's previous
 * / next tensor (has_wl / has_wr = 1); where the reference tensor has no such OpLink (first / last / only tensor of
 * a term: has_wl / has_wr = 0) the caller inserts a dim-1 charge-0 index.  Sites a term skips get no call.  An environment is either MPO-based
 * (tnl_env_set_site_op*) or CouplingModel-based, not both.  makeL!/makeR!/product/noiseterm then follow
 * projcouplingmodel.jl:123-492 id by id. */
int tnl_env_cm_set_term(tnl_env_t env, int32_t site, int64_t id, int32_t has_wl, int32_t has_wr, int32_t nq,
                        const tnl_index_t* inds4, int64_t nblocks, const int32_t* coords, const int64_t* offsets,
                        const double* data);
/* MPS tensor A_site(l, s, r); the env shares the tensor (no copy) */
int tnl_env_set_state(tnl_env_t env, int32_t site, tnl_tensor_t a);
int tnl_env_get_state(tnl_env_t env, int32_t site, tnl_tensor_t* out); /* getpsi, state_envs.jl:36 */
/* StateEnvs(psi, H, Ms; weight): add weight*|M><M| for one fixed MPS M (ProjMPO_MPS2 / ProjMPS2,
 * src/mps/state_envs.jl:86-103, src/mps/projmpo_mps2.jl:94-101, src/mps/projmps2.jl:77-224); call once per M */
int tnl_env_add_penalty(tnl_env_t env, double weight, int32_t nsites, const tnl_tensor_t* tensors);
int tnl_env_set_nsite(tnl_env_t env, int32_t nsite);                   /* set_nsite!, state_envs.jl:352-355 */
int tnl_env_position(tnl_env_t env, int32_t pos);                      /* position!, state_envs.jl:364-367 */
/* orthogonalize!: QR gauge moves of the orthogonality centre from site `from` to site `to`;
 * (N, 1) right-canonicalises an arbitrary MPS (sweep.jl:100-102) */
int tnl_env_move_center(tnl_env_t env, int32_t from, int32_t to);
int tnl_env_make_phi(tnl_env_t env, int32_t pos, tnl_tensor_t* phi); /* psi[pos]*psi[pos+1], update_site.jl:46 */
int tnl_env_apply_flops(tnl_env_t env, double* flops); /* algorithmic flops of one apply at this position */

/* product(sysenv, v) = H_eff v : state_envs.jl:376-378 -> ProjMPO.product */
int tnl_heff_apply(tnl_env_t env, tnl_tensor_t v, tnl_tensor_t* out);
/* eig_solver: src/base/solver.jl:23-43 -> KrylovKit.eigsolve(env, phi, 1, :SR; Lanczos).  phi is updated in place. */
int tnl_eigsolve_lanczos(tnl_env_t env, tnl_tensor_t phi, double tol, int32_t krylovdim, int32_t maxiter,
                         int32_t eager, double* eval, int32_t* converged, int32_t* numops, int32_t* numiter,
                         double* normres);
/* exp_solver: src/base/solver.jl:66-88 -> KrylovKit.exponentiate(env, t, phi; Lanczos): phi <- exp(t * H_eff) phi in
 * place, t = t_re + i t_im.  With t_im != 0 (real-time evolution, time_step = -im*dt) or complex environments a
 * real phi is promoted to a complex tensor first.  *err = accumulated error estimate (info.normres). */
int tnl_exponentiate(tnl_env_t env, tnl_tensor_t phi, double t_re, double t_im, double tol, int32_t krylovdim,
                     int32_t maxiter, int32_t eager, int32_t* converged, int32_t* numops, int32_t* numiter, double* err);
/* real(scalar(dag(phi) * PH(phi))) : src/mps/update_site.jl:51-57 */
int tnl_expectation(tnl_env_t env, tnl_tensor_t phi, double* e);
/* noiseterm + replacebond! : src/mps/update_site.jl:59-76.  which_decomp low 4 bits: 0 = reference rule,
 * 1 = svd, 2 = eigen; bits 4.. select the SVD driver: 0 = `divide_and_conquer` (the reference's default,
 * src/mps/update_site.jl:242) served by the Gram-matrix eigenproblem on the grouped DGEMM with deflated refinement of
 * the small singular values, 1 = cusolverDnXgesvdp (polar), 2 = alias of 0, 3 = cusolverDnDgesvd `qr_iteration`.
 * eigs receives spec.eigs (kept spectrum, descending), at most `cap` values; *neigs = number kept. */
int tnl_replacebond(tnl_env_t env, int32_t pos, tnl_tensor_t phi, int32_t ortho_left, int64_t maxdim, int64_t mindim,
                    double cutoff, double noise, int32_t normalize, int32_t which_decomp, double* truncerr,
                    double* eigs, int64_t cap, int64_t* neigs);

/* one-site update tail, src/mps/update_site.jl:158-186: U,S,V = svd(phi, uinds; maxdim, mindim, cutoff);
 * normalize!(S); psi[pos] = U; psi[posnext] = (S*V)*psi[posnext]  (posnext = pos+1 for ortho left, pos-1 for right).
 * svd_alg: as in tnl_replacebond (0 = Gram + deflated refinement, 1 = gesvdp, 3 = gesvd).
 * carry == NULL: as above.  carry != NULL (TDVP, update_site.jl:172-186): only psi[pos] = U is stored and
 * *carry = S*V (ortho left) or U*S (ortho right) is returned; after the zero-site backward evolution the caller
 * hands it to tnl_env_absorb_bond, which performs psi[posnext] = carry * psi[posnext]. */
int tnl_svd_split(tnl_env_t env, int32_t pos, tnl_tensor_t phi, int32_t ortho_left, int64_t maxdim, int64_t mindim,
                  double cutoff, int32_t normalize, int32_t svd_alg, double* truncerr, double* eigs, int64_t cap,
                  int64_t* neigs, tnl_tensor_t* carry);
int tnl_env_absorb_bond(tnl_env_t env, int32_t pos, int32_t ortho_left, tnl_tensor_t carry);

/* ---- generic block-sparse tensor algebra: the ITensor operations of the tree-tensor-network path ------------------
 * The TTN code of the reference is written as ITensor contractions over index identities
 * (src/ttn/linktensors.jl:63-118,183-221, src/ttn/ttn.jl:266-310, src/ttn/linkproj.jl:56-179,
 * src/ttn/update_site_ttn.jl:75-109); so are the Global Subspace Expansion (src/mps/sweep.jl:399-555) and the
 * measurements (src/mps/measure.jl).  An index identity is an integer LABEL chosen by the caller (the shim uses the
 * ITensor index id + prime level). */
/* out(j_0..) = t(i_perm[0]..): index k of the result is index perm[k] of t; result laid out with `nrow` row indices */
int tnl_tensor_permute(tnl_tensor_t t, const int32_t* perm, int32_t nrow, tnl_tensor_t* out);
int tnl_tensor_dag(tnl_tensor_t t, tnl_tensor_t* out); /* ITensors `dag`: arrows reversed, ComplexF64 conjugated */
/* ITensor `A * B` (or dag(A) * B, ...): contracts every label the two tensors share; result indices = free indices of
 * A then of B; labels_out must hold rank(A) + rank(B) entries.  At least one free index must remain. */
int tnl_tensor_contract(tnl_tensor_t a, const int32_t* labels_a, int32_t dag_a, tnl_tensor_t b, const int32_t* labels_b,
                        int32_t dag_b, tnl_tensor_t* out, int32_t* labels_out, int32_t* rank_out);
/* ITensors `directsum(A => ia, B => ib)` with all other indices shared and in the same order (ia == ib) */
int tnl_tensor_directsum(tnl_tensor_t a, int32_t ia, tnl_tensor_t b, int32_t ib, tnl_tensor_t* out);
/* ITensors `factorize` / `svd` / `qr` of t[(first nleft indices) | rest]: L(left..., m), R(m, right...);
 * which_decomp as in tnl_replacebond (low bits 0 = reference rule, 1 = svd, 2 = eigen, 3 = qr without truncation). */
int tnl_tensor_factorize(tnl_tensor_t t, int32_t nleft, int32_t ortho_left, int64_t maxdim, int64_t mindim, double cutoff,
                         int32_t which_decomp, tnl_tensor_t* L, tnl_tensor_t* R, double* truncerr, double* eigs, int64_t cap,
                         int64_t* neigs);
/* Effective Hamiltonian of a tree node: H v = sum_terms noprime(v * x_1 * ... * x_k) + weight * sum_m <m|v> |m>
 * (product(::LinkTensorsTTN) src/ttn/linktensors.jl:183-221; EnvCouplingModelProjTTN src/ttn/environment.jl:95-101).
 * vlabels: labels of v's indices; a term lists its operand tensors with their labels (concatenated in labels_flat);
 * relabel maps the labels of the primed output indices back to v's (noprime).  eig_solver / exp_solver run the same
 * device Krylov loops as for the MPS (src/ttn/update_site_ttn.jl:60 -> src/base/solver.jl:23-88). */
int tnl_sumop_create(tnl_ctx_t ctx, int32_t rank, const int32_t* vlabels, tnl_sumop_t* out);
int tnl_sumop_destroy(tnl_sumop_t op);
int tnl_sumop_add_term(tnl_sumop_t op, int32_t nops, const tnl_tensor_t* tensors, const int32_t* labels_flat);
int tnl_sumop_set_relabel(tnl_sumop_t op, int32_t n, const int32_t* from, const int32_t* to);
int tnl_sumop_add_projector(tnl_sumop_t op, tnl_tensor_t m, double weight);
int tnl_sumop_apply(tnl_sumop_t op, tnl_tensor_t v, tnl_tensor_t* out);
int tnl_sumop_apply_flops(tnl_sumop_t op, double* flops); /* algorithmic GEMM flops of the last apply */
int tnl_sumop_eigsolve(tnl_sumop_t op, tnl_tensor_t phi, double tol, int32_t krylovdim, int32_t maxiter, int32_t eager,
                       double* eval, int32_t* converged, int32_t* numops, int32_t* numiter, double* normres);
int tnl_sumop_exponentiate(tnl_sumop_t op, tnl_tensor_t phi, double t_re, double t_im, double tol, int32_t krylovdim,
                           int32_t maxiter, int32_t eager, int32_t* converged, int32_t* numops, int32_t* numiter, double* err);

#ifdef __cplusplus
}
#endif
#endif

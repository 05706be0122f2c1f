// Generic block-sparse tensor algebra on the device: index permutation, contraction over labelled indices, direct
// sums, and the sum-of-products operator that carries the effective Hamiltonian of a tree tensor network.
//
// The MPS path (env.cpp) knows its index roles statically; the tree tensor networks of the reference express
// everything as ITensor contractions over index identities,
//   product(::LinkTensorsTTN)        src/ttn/linktensors.jl:183-221     sum over ids of contract(v, link tensors...)
//   move_linktensors_to_next!        src/ttn/linktensors.jl:63-118      dag(prime(phi)) * tensors * phi
//   moveisometry_to_next!            src/ttn/ttn.jl:266-310             svd / qr of a node tensor, absorb into the next
//   subspace_expand!                 src/ttn/update_site_ttn.jl:75-109  directsum with random padding
//   LinkProjTTN                      src/ttn/linkproj.jl:56-179         overlaps with a fixed TTN
// and so do the Global Subspace Expansion (`apply(H, psi)`, reduced density matrices, `directsum`,
// src/mps/sweep.jl:399-555) and the measurements (src/mps/measure.jl).  Here an index identity is an integer LABEL
// supplied by the host; a contraction permutes both operands into the charge-fused [free | contracted] /
// [contracted | free] layouts (one tiled HBM pass each, skipped when the operand already has that form) and runs
// one grouped DGEMM over the charge sectors.
#pragma once
#include "env.hpp"

namespace tnl {

// Y(j_0, ..., j_{r-1}) = X(i_perm[0], ..., i_perm[r-1]): Y's index k is X's index perm[k]; result laid out with `nrow`
TensorP permute(Ctx* ctx, const Tensor& X, const std::vector<int>& perm, int nrow);

TensorP dag_copy(Ctx* ctx, const Tensor& X);   // dag(X) materialised: arrows reversed, ComplexF64 conjugated

// C = A * B over all labels the two tensors share (ITensor `*`).  dagA / dagB: the operand enters as dag(.) -- arrows
// reversed and, for ComplexF64, conjugated.  Result indices: free indices of A (in A's order) then those of B;
// `lc` receives their labels.  nrow of the result = number of free indices of A (at least one free index in total).
TensorP contract(Ctx* ctx, const Tensor& A, const std::vector<int>& la, bool dagA, const Tensor& B,
                 const std::vector<int>& lb, bool dagB, std::vector<int>* lc);

// ITensors `directsum(A => ia, B => ib)`: all other indices are shared (same order in both tensors); the new index lists
// the sectors of A's index `ia` followed by those of B's index `ib`.  Result in A's index order.
TensorP directsum(Ctx* ctx, const Tensor& A, int ia, const Tensor& B, int ib);

// H v = sum_terms noprime( v * x_1 * ... * x_k )  +  weight * sum_m <m|v> |m>
// (product(::LinkTensorsTTN), product(::EnvCouplingModelProjTTN) src/ttn/environment.jl:95-101).  Every operand is a
// tensor with integer labels; `relabel` maps the labels of the primed (output) indices back to the vector's labels.
class SumOp {
 public:
  struct Operand { TensorP t; std::vector<int> labels; };
  struct Term { std::vector<Operand> ops; };
  Ctx* ctx;
  std::vector<int> vlabels;                        // labels of the vector's indices, in the vector's index order
  std::vector<std::pair<int, int>> relabel;        // (primed label -> vector label) applied to every term's result
  std::vector<Term> terms;
  std::vector<TensorP> projs;                      // |m> in the layout of the vector
  double weight = 0.0;
  explicit SumOp(Ctx* c) : ctx(c) {}

  void apply(const Tensor& v, Tensor& out);
  double flops = 0;                                // algorithmic GEMM flops of the last apply
  // operator interface of the Krylov templates (krylov.hpp); never sharded
  void ensure_plan(const Tensor&) {}
  bool op_sharded() const { return false; }
  int64_t op_nloc() const { return 0; }
  void apply_local(const double*, double*) { throw Error(2, "SumOp is not sharded"); }
  void apply_ptr(const Tensor& proto, const double* vin, double* vout);
  void op_to_local(const double*, double*) {}
  void op_gather(const double*, double*) {}
};

}  // namespace tnl

// Device-resident mirror of the reference's StateEnvs{ProjMPO} (src/mps/state_envs.jl:18-27,54-60):
// the MPS, the MPO, the cached left/right environments with lpos/rpos watermarks
// (same contract as src/mps/projcouplingmodel.jl:123-129,212-218), the H_eff apply, the Lanczos
// eigensolver around it and the bond factorisation.
#pragma once
#include "core.hpp"

namespace tnl {

struct HostBlocks {             // NDTensors-style flat host tensor (column-major blocks)
  int rank = 0;
  std::vector<Index> inds;
  std::vector<std::vector<int>> coords;
  std::vector<int64_t> offsets;
  const double* data = nullptr;
  bool cplx = false;            // data is interleaved ComplexF64 (offsets count complex elements)
};

TensorP import_tensor(Ctx* ctx, const HostBlocks& hb, int nrow);
// natural-layout copy: blocks contiguous column-major, column-major block order
TensorP to_natural(Ctx* ctx, const Tensor& t);
TensorP relayout(Ctx* ctx, const Tensor& t, int nrow);

struct FactorizeParams {
  int ortho_left = 1;           // 1: L isometry ("left"), 0: R isometry ("right")
  int64_t maxdim = INT64_MAX;
  int64_t mindim = 1;
  double cutoff = 0.0;
  int which = 0;                // 0 = automatic (reference rule), 1 = svd, 2 = eigen, 3 = qr (no truncation)
  int svd_alg = 0;              // svd path: 0 = cusolverDnDgesvd (QR iteration), 1 = cusolverDnXgesvdp (polar)
  double noise = 0.0;           // scale of the density-matrix perturbation
  const Tensor* noiseX = nullptr;   // see factorize.cu
  std::vector<const Tensor*> noiseXmore;   // further noise operands (MPO sums): rho += noise * sum_k X_k X_k^T
  int new_dir_on_L = -1;        // arrow of the new index on the left factor
};
struct FactorizeResult {
  TensorP L, R;                 // L: (rows..., m) nrow = split ; R: (m, cols...) nrow = 1
  std::vector<double> eigs;     // kept spectrum, descending
  double truncerr = 0.0;
  std::string path;
};
// T must be laid out with nrow = split.
FactorizeResult factorize(Ctx* ctx, const Tensor& T, const FactorizeParams& prm);

struct LanczosResult {
  double eval = 0;
  int converged = 0;
  int numops = 0;
  int numiter = 0;
  double normres = 0;
};

bool is_trivial_link(const Index& ix);

struct ExpResult {
  int converged = 0;
  int numops = 0;
  int numiter = 0;
  double err = 0;               // accumulated error estimate (KrylovKit info.normres)
};
constexpr int LC_MAX_HOST = 32;   // vec_lincomb takes at most this many vectors (krylovdim + residual)

class Env {
 public:
  Ctx* ctx;
  int N;
  int nsite = 2;
  int lpos, rpos;
  std::vector<TensorP> Wlr, Wrl, Wnr; // MPO tensors as transform operands: (wl,s | s',wr), (s',wr | wl,s), (s,wr | wl,s')
  std::vector<Index> Wl, Wr;         // MPO link indices per site
  std::vector<TensorP> A_store;
  std::vector<TensorP>& A;           // site tensors (l, s, r), any nrow; terms of an MPO sum share the top env's
  // StateEnvs(psi, Hs::Vector{MPO}) = ProjMPOSum2 (src/mps/projmposum2.jl:15-145): the further MPOs H_2, H_3, ...
  // are full environments of their own (W tensors, L/R cache, watermarks, apply plan) over the SAME state;
  // product = sum of the terms' products, noiseterm = sum of the terms' noise terms.
  std::vector<std::unique_ptr<Env>> more;
  Env* parent = nullptr;
  // StateEnvs(psi, H::CouplingModel) = ProjCouplingModel (src/mps/projcouplingmodel.jl): per-id site operators
  // and environments; see the implementation notes in env.cpp
  struct CM;
  struct CMDeleter { void operator()(CM* p) const; };
  std::unique_ptr<CM, CMDeleter> cm;
  // W(wl, s', s, wr); open_l / open_r: wl / wr is a real OpLink (otherwise a dim-1 charge-0 placeholder)
  void cm_set_term(int site, int64_t id, const HostBlocks& hb, bool open_l, bool open_r);
  Env& term(int k);                  // k = 0: this env; k >= 1: more[k-1] (created on demand)
  int nterms() const { return 1 + (int)more.size(); }
  void set_nsite(int n);
  std::vector<TensorP> LR;           // LR[j] = L_{j+1} (nrow 2) or R_{j+1} (nrow 1), 0-based slot j = site j+1
  TensorP Ledge, Redge;
  // excited-state penalty  weight * sum_M |M><M|  (ProjMPO_MPS2, src/mps/projmpo_mps2.jl:94-134)
  struct Penalty {
    std::vector<TensorP> M;          // fixed MPS tensors (l, s, r)
    std::vector<TensorP> LR;         // overlap environments (psi link, M link), nrow 1
    TensorP Ledge, Redge;
    int lpos = 0, rpos = 0;
    bool dead = false;               // different total charge: overlap vanishes identically
    TensorP m;                       // |m> = dag(proj_mps) at the current position, Krylov layout
    TensorP mloc;                    // its local slice (sharded apply)
  };
  std::vector<Penalty> pens;
  double weight = 0.0;
  void add_penalty(const std::vector<TensorP>& M, double w);
  void invalidate(int lo, int hi);

  Env(Ctx* c, int n, Env* par = nullptr)
      : ctx(c), N(n), lpos(0), rpos(n + 1), Wlr(n), Wrl(n), Wnr(n), Wl(n), Wr(n), A_store(par ? 0 : n),
        A(par ? par->A_store : A_store), LR(n), parent(par) {}

  void set_site_op(int site, const HostBlocks& hb);
  void update_site_op(int site, const HostBlocks& hb);   // updateH!(...; recalcEnv = false)     // W(wl, s', s, wr), 1-based site
  void set_state(int site, TensorP a);
  void position(int pos);                               // makeL!(pos-1), makeR!(pos+nsite)
  TensorP make_phi(int pos);                            // two-site tensor (l,s1,s2,r) in Krylov layout (nrow 1)
  void apply(const Tensor& v, Tensor& out);             // out = H_eff v, both in Krylov layout
  void apply_ptr(const Tensor& proto, const double* vin, double* vout);
  void apply_local(const double* vloc, double* outloc);   // sharded core (multi-GPU)
  void ensure_plan(const Tensor& proto);
  // operator interface of the Krylov templates (krylov.hpp)
  bool op_sharded() const;
  int64_t op_nloc() const;
  void op_to_local(const double* full, double* loc);
  void op_gather(const double* loc, double* full);
  LanczosResult eigsolve(Tensor& phi, double tol, int krylovdim, int maxiter, bool eager);
  // exp_solver: phi <- exp(t * H_eff) phi  (KrylovKit.exponentiate)
  ExpResult exponentiate(Tensor& phi, double t_re, double t_im, double tol, int krylovdim, int maxiter, bool eager);
  double expectation(const Tensor& phi);
  FactorizeResult replacebond(int pos, const Tensor& phi, FactorizeParams prm, bool normalize);
  // absorb = false: psi[pos] = U only; the caller evolves the bond tensor (TDVP reverse step) and calls absorb_bond
  FactorizeResult svd_split(int pos, const Tensor& phi, FactorizeParams prm, bool normalize, bool absorb = true);
  void absorb_bond(int pos, bool left, const Tensor& carry);
  void move_center(int from, int to);                   // QR gauge moves, ITensorMPS orthogonalize!
  double apply_flops() const;                           // algorithmic flops of one H_eff apply at the current position
  bool complex_at_position() const;                     // any environment tensor bounding the site range is complex

 private:
  struct ApplyPlan;
  std::shared_ptr<ApplyPlan> ap;
  void build_apply_plan(const Tensor& v);
  void build_scatter_tables(ApplyPlan& p);   // multi-GPU: epilogue tables of the fused GEMM -> reduce-scatter
  std::shared_ptr<ApplyPlan> make_plan(const Tensor& v, TensorP L, TensorP W1, TensorP W2, TensorP R, bool allow_shard);
  void run_plan(ApplyPlan& p, const double* vin, double* vout);
  TensorP step_left(const Tensor* L, const TensorP& Asite, const Tensor& W);
  TensorP step_right(const Tensor* R, const TensorP& Asite, const Tensor& W);
  TensorP lproj();
  TensorP rproj();
  void ensure_edges();
  void makeL(int k);
  void makeR(int k);
  TensorP noise_tensor(const Tensor& phi, bool left, bool own_storage);
  TensorP noise_operand(const Tensor& phi, bool left, const Tensor* E, const Tensor& W, bool own_storage);
  void cm_makeL(int k);
  void cm_makeR(int k);
  void cm_ensure_plans(const Tensor& proto);
  void cm_apply(const Tensor& proto, const double* vin, double* vout);
  double cm_apply_flops() const;
  std::vector<TensorP> cm_noise_operands(const Tensor& phi, bool left);
  void position_penalty(Penalty& p, int pos);
  void build_penalty_vector(Penalty& p, const Tensor& proto);
};

}  // namespace tnl

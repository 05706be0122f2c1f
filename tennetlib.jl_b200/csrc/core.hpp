// tnl_b200 core data model: QN indices and charge-fused block-sparse tensors in HBM.
//
// Design: a block-sparse tensor with flux 0 and a bipartition
// (first `nrow` indices | remaining indices) is block diagonal in the fused charge of the
// row group.  We store every charge group as ONE dense column-major "super-matrix"
// [R_q x C_q] (leading dimension padded to 16 B) whose sub-blocks are the NDTensors blocks.
// Consequences:
//   * contracting two tensors over a whole row/column group is a plain dense DGEMM per charge
//     sector (grouped_gemm.cu) -- no per-block-pair GEMMs, no ragged tiny tiles;
//   * regrouping indices (and applying the skinny MPO site operators on the way) is one
//     HBM-bound pass (transform.cu);
//   * the flat buffer of a Krylov vector has no holes except zero padding, so inner products /
//     axpys are flat BLAS-1 kernels (vecops.cu).
// Semantics follow NDTensors' BlockSparse storage as consumed by the reference at
// src/mps/projcouplingmodel.jl:315-356, src/mps/update_site.jl:46-76.
#pragma once
#include <cuda_runtime.h>

#include <array>
#include <cstdint>
#include <cstdio>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

namespace tnl {

constexpr int MAXR = 6;   // max tensor rank on the hot path (T1/T2/T3 of the two-site apply are rank 5)
constexpr int MAXQ = 2;   // U(1) or U(1)xU(1)
using Charge = std::array<int, MAXQ>;

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
#define TNL_CHECK(cond, msg)                                                         \
  do {                                                                               \
    if (!(cond)) throw ::tnl::Error(2, std::string(msg) + " [" #cond "] at " __FILE__ ":" + std::to_string(__LINE__)); \
  } while (0)
#define CUDA_OK(call)                                                                \
  do {                                                                               \
    cudaError_t e__ = (call);                                                        \
    if (e__ != cudaSuccess)                                                          \
      throw ::tnl::Error(3, std::string("CUDA error: ") + cudaGetErrorString(e__) + " in " #call " at " __FILE__ ":" + \
                                std::to_string(__LINE__));                           \
  } while (0)

struct Index {
  std::vector<int> dims;      // sector dims
  std::vector<Charge> qns;    // sector charges
  int dir = +1;               // arrow
  int nq = 1;
  int nsect() const { return (int)dims.size(); }
  int64_t dim() const { int64_t d = 0; for (int x : dims) d += x; return d; }
  bool same_space(const Index& o) const { return dims == o.dims && qns == o.qns && nq == o.nq; }
  Index dag() const { Index r = *this; r.dir = -dir; return r; }
};

inline Charge charge_zero() { Charge c; c.fill(0); return c; }
inline Charge charge_neg(Charge c) { for (auto& x : c) x = -x; return c; }

// one product of sectors of an index group
struct Combo {
  int c[MAXR];       // sector number per index of the group
  int d[MAXR];       // sector dims
  int64_t dim;       // product of dims
  int64_t off;       // offset inside the group's row (or column) range
};

struct Group {
  Charge q;                 // charge of the ROW group: sum_i dir_i * qn_i
  int64_t R = 0, C = 0;     // logical rows / cols
  int64_t ld = 0;           // padded leading dimension (>= R, even)
  int64_t base = 0;         // element offset of the super-matrix in the tensor buffer
  std::vector<Combo> rows, cols;
};

struct Block {
  int c[MAXR];
  int d[MAXR];
  int64_t st[MAXR];         // element strides
  int64_t off;              // element offset of element (0,..,0)
  int group;
};

class Ctx;

// Block-sparse tensor, flux 0, fused layout with the first `nrow` indices as row group.
class Tensor {
 public:
  Ctx* ctx = nullptr;
  std::vector<Index> inds;
  int nrow = 0;
  std::vector<Group> groups;
  std::vector<Block> blocks;                       // all symmetry-allowed blocks
  std::unordered_map<uint64_t, int> lut;           // packed coords -> block number
  std::vector<uint8_t> present;                    // imported tensors: which blocks the host supplied (empty = all)
  double* d = nullptr;
  bool owns = true;                                // false: `d` points into a Ctx workspace slot
  int64_t nelem = 0;                               // padded element count of the buffer (of ONE plane)
  // ComplexF64 tensors are stored PLANAR: real plane at d, imaginary plane at d + nelem, both in the same
  // charge-fused layout -- a complex contraction is four (two) launches of the real FP64 DMMA kernel, the
  // site-operator transforms (real MPOs) run plane by plane, norms / real scalings see one flat buffer of 2*nelem
  bool cplx = false;
  double* im() const { return d + nelem; }
  int64_t planes() const { return cplx ? 2 : 1; }

  Tensor(Ctx* ctx, std::vector<Index> inds, int nrow, bool alloc = true, bool cplx = false);
  ~Tensor();
  Tensor(const Tensor&) = delete;
  Tensor& operator=(const Tensor&) = delete;

  int rank() const { return (int)inds.size(); }
  static uint64_t key(const int* c, int r) {
    uint64_t k = 0;
    for (int i = 0; i < r; i++) k = (k << 10) | (uint64_t)(c[i] & 1023);
    return k;
  }
  int find(const int* c) const {
    auto it = lut.find(key(c, rank()));
    return it == lut.end() ? -1 : it->second;
  }
  int find_group(const Charge& q) const {
    for (size_t g = 0; g < groups.size(); g++)
      if (groups[g].q == q) return (int)g;
    return -1;
  }
  int64_t logical_elems() const { int64_t n = 0; for (auto& g : groups) n += g.R * g.C; return n; }
  void zero();
};
using TensorP = std::shared_ptr<Tensor>;

// ---- device work descriptors ---------------------------------------------------------------
struct GemmProblem {          // C[M x N] = op(A)[M x K] * op(B)[K x N], column-major
  int64_t a, b, c;            // element offsets relative to the base pointers given at launch
  int M, N, K;
  int lda, ldb, ldc;
};
struct GemmTile { int prob; int m0; int n0; };
// 128x128 output tile of the persistent TMA kernel: everything a CTA needs without a second (dependent) load
// `kt0` / `ktiles`: the tile's range of 16-wide k-tiles.  Split-K tiles (`split` != 0) cover a part of K and store their
// partial result into the plan's workspace (c = offset there, ldc = its leading dimension); a reduction kernel sums
// the parts in order afterwards.
struct TmaTile { int64_t c; int m0, n0, M, N, ktiles, ldc, prob, kt0, split, pad; };
// one split problem: C[M x N] (+)= alpha * sum_s ws[s]   with ws[s] = workspace + ws + s * ldw * N
struct SplitDesc { int64_t c, ws; int M, N, ldc, ldw, S, pad; };

struct GemmPlan {
  bool transA = false, transB = false;
  bool lower_only = false;              // symmetric result (Gram matrices): only tiles that touch the lower triangle
  std::vector<GemmProblem> probs;
  std::vector<GemmTile> tiles_big, tiles_small;
  std::vector<TmaTile> tiles_tma;       // tiles of the sectors larger than 64 in both extents (TMA kernel)
  GemmProblem* d_probs = nullptr;
  GemmTile* d_tiles_big = nullptr;
  GemmTile* d_tiles_small = nullptr;
  TmaTile* d_tiles_tma = nullptr;
  // TMA descriptors (CUtensorMap, 128 bytes each; A and B of every problem).  A descriptor holds the absolute
  // global address, and the operands of a plan change from launch to launch (Krylov vectors), so the plan keeps
  // one descriptor set per distinct (A, B) base pair it has been launched with.
  static constexpr int MAPSETS_MAX = 24;
  struct MapSet { const double* A; const double* B; };
  std::vector<MapSet> mapsets;
  void* d_maps = nullptr;               // MAPSETS_MAX x 2*probs.size() descriptors
  // split-K (few output tiles, long K: environment updates of tree nodes, Gram matrices of skinny blocks)
  std::vector<SplitDesc> splits;
  SplitDesc* d_splits = nullptr;
  double* d_splitws = nullptr;
  std::vector<std::pair<int64_t, int64_t>> zero_fill;   // (offset, count) of C ranges with no contribution
  double flops = 0;
  Ctx* ctx = nullptr;
  ~GemmPlan();
};

// Scatter epilogue of the grouped DGEMM (multi-GPU, fused reduce-scatter): element (m, n) of problem `prob` goes to
// rank `rank(n)`, into the slot this rank owns there, at   T[rb(m) * nseg + seg(n)] + roff(m) + jj(n) * cstride[rb(m)]
// -- the position of the element in the destination rank's local Krylov-vector layout.
struct ScatterProb {
  const int2* rowinfo;       // [M]: (row block rb, offset of the row inside a destination block)
  const int2* colinfo;       // [N]: (segment | destination rank << 16, column inside the destination rank's share)
  const int64_t* T;          // [nrb * nseg]: offset of the destination block
  const int* cstride;        // [nrb]: stride between consecutive columns of the destination block
  int nseg, pad;
};
struct ScatterArgs {
  const ScatterProb* probs = nullptr;
  double* const* peer_slots = nullptr;          // [world]: my slot inside every rank's staging area
  unsigned long long* const* peer_flags = nullptr;
  unsigned int* done = nullptr;                 // last-CTA counter
  unsigned long long epoch = 0;
  int rank = 0, world = 1;
};

// transform: Y(i, n..., p...) = sum_contrib sum_k X(i, k..., p...) * W(k, n)
constexpr int MAXP = 4;
struct XfContrib {
  int64_t xoff;              // X element offset of (i=0, k=0, p=0)
  int64_t xps[MAXP];         // X strides of the passive dims
  int64_t ks0, ks1;          // X strides of the (up to two) contracted dims
  int kd0, kd1;              // their dims (Ka = kd0*kd1)
  int64_t woff;              // W element offset; W block is [Ka x Na] column-major
};
struct XfBlock {
  int64_t yoff;              // Y element offset of (i=0, n=0, p=0)
  int64_t yns[2];            // Y strides of the (up to two) new dims
  int nd0, nd1;              // their dims (Na = nd0*nd1)
  int64_t yps[MAXP];         // Y strides of the passive dims
  int pd[MAXP];              // passive dims
  int I;                     // leading (stride-1) dim
  int cbeg, cnum;            // contribution range
  int64_t colstart;          // first global column of this block (columns = Na * prod(pd))
  int fbeg, fnum;            // range in the flattened contribution list (one entry per contracted element)
};
// one contracted element (k0, k1) of one contribution: a column of X and the W row that multiplies it
struct XfFlat {
  int64_t xoff;              // X element offset of (i=0, this k, p=0)
  int64_t xps[MAXP];
  int64_t woff;              // W element offset of (this k, n=0)
  int wst;                   // W stride between consecutive n (= Ka of the contribution)
};
// Output blocks that read the SAME X columns (same leading and passive sectors, different new-index sectors): the
// kernel walks a group column-major over (passive column, block), so the warps of a CTA that need one X column run
// next to each other and the column comes from DRAM once (site operators feed 2-3 output blocks from each column).
// `colstart` counts work items: nb * P per group (cpw = 1: one passive column per warp; packing several columns of a
// tiny leading sector into one warp was tried and changed nothing -- the kernel was bound by per-item bookkeeping).
struct XfGroup { int64_t colstart; int64_t P; int first; int nb; int cpw; int I2; };
struct TransformPlan {
  std::vector<XfGroup> groups;
  XfGroup* d_groups = nullptr;
  std::vector<XfBlock> blocks;
  std::vector<XfContrib> contribs;
  std::vector<XfFlat> flats;
  XfBlock* d_blocks = nullptr;
  XfFlat* d_flats = nullptr;
  bool pure_copy = false;    // no W: every output column is a copy of one input column (or zero)
  int64_t ncols = 0;
  int64_t nitems = 0;        // work items of the grouped kernel (plans with W)
  double bytes = 0;          // algorithmic bytes moved (read X once + write Y once)
  double flops = 0;
  Ctx* ctx = nullptr;
  ~TransformPlan();
};

// ---- context --------------------------------------------------------------------------------
struct Counters {
  double gemm_flops = 0;          // algorithmic 2mnk summed over executed GEMM problems
  double xf_flops = 0;
  double vec_bytes = 0;           // algorithmic bytes of Krylov vector kernels
  double xf_bytes = 0;
  long long launches = 0;         // kernels of THIS library launched
  long long gemm_launches = 0;
  double apply_count = 0;
  double allreduce_bytes = 0;
  double host_plan_ms = 0;        // host wall time spent building apply plans (planner + uploads)
};

class Ctx {
 public:
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string last_error;
  Counters cnt;
  double* d_scalars = nullptr;      // device scratch for reductions
  double* h_scalars = nullptr;      // pinned mirror
  unsigned int* d_sync = nullptr;   // "last block" counters
  double* d_partials = nullptr;
  int num_sms = 148;
  // optional per-launch CUDA-event profile of the grouped GEMM kernel (bench.py roofline)
  struct ProfRec { cudaEvent_t a, b; double flops; int tiles; int cat; };   // cat: 0 gemm, 1 transform, 2 vec, 3 comm
  bool prof_gemm = false;
  std::vector<ProfRec> prof_recs;
  // scoped CUDA-event timer for the non-GEMM categories (active only while prof_gemm is on)
  struct Scope {
    Ctx* c; ProfRec r; bool on;
    Scope(Ctx* ctx, int cat) : c(ctx), on(ctx->prof_gemm) {
      if (!on) return;
      r = ProfRec{};
      r.cat = cat;
      cudaEventCreate(&r.a); cudaEventCreate(&r.b);
      cudaEventRecord(r.a, c->stream);
    }
    ~Scope() { if (on) { cudaEventRecord(r.b, c->stream); c->prof_recs.push_back(r); } }
  };
  // NCCL sharding (multi-GPU apply); rank 0 / world 1 when unused
  int rank = 0, world = 1;
  void* nccl_comm = nullptr;
  bool shard_enabled = true;        // false: every rank computes the whole (replicated, no collectives) -- the
                                    // reference path the sharded results are checked against (bench.py parity)
  int shard_world() const { return shard_enabled ? world : 1; }
  // Fused GEMM -> reduce-scatter over peer memory (comm.cpp, kernels.cu): every rank owns a staging area of `world`
  // slots (one per source rank) that its peers write straight from the epilogue of the `T3 R^T` GEMM through
  // CUDA-IPC mappings, plus two flag words per source rank ("data of epoch e complete" / "slot consumed up to e").
  struct PeerStage {
    bool tried = false, ok = false;
    void* base = nullptr;               // cudaMalloc: [flags: 2 * 64 words][world slots of slot_cap doubles]
    size_t slot_cap = 0;                // doubles per slot
    std::vector<void*> peer_base;       // the peers' bases mapped into this process (own base for rank == me)
    double** d_peer_slots = nullptr;    // device array [world]: address of MY slot inside every peer's staging area
    unsigned long long** d_peer_flags = nullptr;   // device array [world]: the peers' flag blocks
    unsigned int* d_done = nullptr;     // "last CTA" counter of the scattering GEMM
    unsigned long long epoch = 0;       // number of fused reduce-scatters issued (identical on every rank)
    unsigned long long ar_epoch = 0;    // number of scalar all-reduces issued through the mailboxes
    bool small_ar = true;               // scalar all-reduces over the mailboxes (TNL_PEER_SCALAR_AR=0: NCCL)
  } pstage;
  // cuSOLVER (opaque here; factorize.cu owns the type)
  void* cusolver = nullptr;
  void* solver_work = nullptr; size_t solver_work_bytes = 0;
  int* d_info = nullptr;

  explicit Ctx(int device);
  ~Ctx();
  void* alloc(size_t bytes);        // stream-ordered (cudaMallocAsync on the context stream)
  void free(void* p);
  // Persistent, grow-only workspaces for the big per-bond temporaries (T1..T3 of the apply, Krylov vectors,
  // factorisation scratch): allocated once at their high-water mark instead of once per bond, so the timed path
  // never waits for the driver to find / map gigabyte-sized blocks.
  enum { SLOT_T1 = 0, SLOT_T2, SLOT_T3, SLOT_P, SLOT_PACKED, SLOT_LOCIN, SLOT_LOCOUT, SLOT_VLOC, SLOT_ARENA, NSLOTS };
  struct Slot { void* p = nullptr; size_t cap = 0; };
  Slot slots[NSLOTS];
  uint64_t slot_epoch = 1;          // bumped when a slot is reallocated or reused outside an apply plan
  double* scratch(int slot, size_t nelem, bool zero);
  std::unordered_map<double*, size_t> vec_caps;
  std::vector<std::pair<double*, size_t>> vec_pool;     // free Krylov-vector buffers (ptr, capacity in doubles)
  double* vec_acquire(size_t nelem);                    // zero-filled
  void vec_release(double* p, size_t cap);
  // bump allocator inside SLOT_ARENA (factorisation temporaries); falls back to alloc() on overflow and grows
  // the arena for the next call
  size_t arena_off = 0, arena_want = 0;
  void arena_reset();
  double* arena_alloc(size_t nelem, bool* from_arena);
  // pinned staging ring for small host->device uploads that must not block the host (TMA descriptor sets)
  unsigned char* pin_ring = nullptr; size_t pin_cap = 0, pin_off = 0;
  void* stage_pinned(size_t bytes);
  bool use_tma = true;              // grouped GEMM through the TMA kernel (TNL_GEMM_TMA=0 selects the cp.async kernel)
  bool dual_gemm = true;            // complex x complex products: one dual-source launch per plane (TNL_GEMM_DUAL=0: four launches)
  void sync() { CUDA_OK(cudaStreamSynchronize(stream)); }
  template <class T> T* upload(const std::vector<T>& v) {
    if (v.empty()) return nullptr;
    T* p = (T*)alloc(v.size() * sizeof(T));
    CUDA_OK(cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, stream));
    CUDA_OK(cudaStreamSynchronize(stream));   // source vector may die right after
    return p;
  }
};

// ---- kernels / planners (implemented in the .cu / .cpp files) --------------------------------
void build_layout(Tensor& t);

// dagA / dagB: the operand enters with reversed arrows (ITensor `dag`); data is real so only the charge
// bookkeeping changes.
std::unique_ptr<GemmPlan> plan_gemm(const Tensor& A, bool transA, const Tensor& B, bool transB, Tensor& C,
                                    bool dagA = false, bool dagB = false);
// C = alpha * op(A) op(B)  (accum: += ; alpha = +-1 and accumulate serve the planar complex products)
void run_gemm(Ctx* ctx, GemmPlan& p, const double* A, const double* B, double* C, double alpha = 1.0, bool accum = false);
// Fused GEMM -> reduce-scatter over peer memory: C tiles are stored straight into the destination ranks' staging slots
// (scatter epilogue), completion is flagged to every peer, and `out` (nloc doubles) receives the sum over the source
// ranks of this rank's slots, in rank order.  The plan must consist of 128x128 tiles only (plan_all_tma).
void run_gemm_reduce_scatter(Ctx* ctx, GemmPlan& p, const double* A, const double* B, const ScatterProb* d_scatter, double* out,
                             int64_t nloc);
bool plan_all_tma(const GemmPlan& p);
void gemm_plan_force_tma(GemmPlan& plan);   // every problem on the TMA kernel (also the small sectors)
// ad-hoc problems (factorisation internals); builds tiles and uploads
// lower_only: the products are symmetric (M M^T, M^T M) and the consumer reads the lower triangle only (cuSOLVER
// CUBLAS_FILL_MODE_LOWER): tiles strictly above the diagonal are not computed
std::unique_ptr<GemmPlan> plan_gemm_raw(Ctx* ctx, bool transA, bool transB, const std::vector<GemmProblem>& probs,
                                        bool lower_only = false);
void gemm_selftest(Ctx* ctx, int M, int N, int K, bool ta, bool tb, int reps, bool verify, double* ms, double* maxerr);
void comm_unique_id(char* out128);
void comm_init(Ctx* ctx, const char* uid128, int rank, int world);
void comm_destroy(Ctx* ctx);
// staging area with at least `slot_doubles` per slot, mapped on every rank (collective call); false: fused path off
bool comm_stage_ensure(Ctx* ctx, size_t slot_doubles);
void comm_allreduce_sum(Ctx* ctx, double* buf, int64_t n);
void comm_reduce_scatter_sum(Ctx* ctx, const double* send, double* recv, int64_t n);
void comm_allgather(Ctx* ctx, const double* send, double* recv, int64_t n);
void shard_range(int d, int world, int sector, int rank, int* start, int* count);
void transpose(Ctx* ctx, double* dst, int64_t ldd, const double* src, int64_t lds, int64_t R, int64_t C);  // dst[C x R] = src[R x C]^T
void copy2d(Ctx* ctx, double* dst, int64_t ldd, const double* src, int64_t lds, int64_t R, int64_t C);

// W == nullptr: pure relayout / index permutation.  `xmap[j]` = position in X of Y's index j, or -1 for a
// new index (coming from W); `kpos` lists the X positions that are contracted with W (in W's row order);
// W's indices are ordered (contracted..., new...).
// Optional slicing of one PASSIVE index: Y's index `ypos` is a sub-range of X's index xmap[ypos];
// Y sector k corresponds to X sector orig[k], elements [start[k], start[k] + dim_Y(k)).
struct SliceMap { int ypos = -1; std::vector<int> orig; std::vector<int64_t> start; };
std::unique_ptr<TransformPlan> plan_transform(const Tensor& X, Tensor& Y, const std::vector<int>& xmap,
                                              const Tensor* W, const std::vector<int>& kpos,
                                              const SliceMap* slice = nullptr);
// inverse of a sliced transform: every block of the LOCAL tensor X is written into the sub-range of the matching
// block of the full tensor Y (index `slice.ypos`, Y sector slice.orig[k], elements from slice.start[k])
std::unique_ptr<TransformPlan> plan_scatter(const Tensor& X, Tensor& Y, const SliceMap& slice);
void finalize_transform_plan(TransformPlan& p, bool has_w);   // flattens the contributions and uploads the plan
void run_transform(Ctx* ctx, TransformPlan& p, const double* X, double* Y, const double* W);

// flat vector kernels on padded buffers of equal layout
void vec_dot(Ctx* ctx, const double* x, const double* y, int64_t n, int slot);       // d_scalars[slot] = <x,y>
void vec_axpy_dev(Ctx* ctx, double* y, const double* x, int64_t n, int slot, double sign);  // y += sign*s[slot]*x
void vec_axpy(Ctx* ctx, double* y, const double* x, int64_t n, double a);
// fused MGS step: w += (slot_in >= 0 ? a * s[slot_in] : a) * x, then s[slot_out] = <y, w> (y may alias w), one pass
void vec_axpy_dot(Ctx* ctx, double* w, const double* x, int64_t n, int slot_in, double a, const double* y, int slot_out);
void vec_caxpy_cdot(Ctx* ctx, double* w, const double* x, int64_t n, int slot_in, double ar, double ai, const double* y,
                    int slot_out);
void vec_scale(Ctx* ctx, double* y, int64_t n, double a);
void vec_scale_to(Ctx* ctx, double* y, const double* x, int64_t n, double a);         // y = a*x
void vec_copy(Ctx* ctx, double* y, const double* x, int64_t n);
void vec_lincomb(Ctx* ctx, double* y, const double* const* xs, const double* coef, int k, int64_t n);
// planar complex vectors [re | im] with planes of n doubles; scalar slots hold (re, im) pairs
void vec_cdot(Ctx* ctx, const double* x, const double* y, int64_t n, int slot);      // d_scalars[slot, slot+1] = <x,y>
void vec_caxpy_dev(Ctx* ctx, double* y, const double* x, int64_t n, int slot, double sign);
void vec_caxpy(Ctx* ctx, double* y, const double* x, int64_t n, double ar, double ai);
void vec_clincomb(Ctx* ctx, double* y, const double* const* xs, const double* cr, const double* ci, int k, int64_t n);
void cgemm(Ctx* ctx, GemmPlan& p, const double* Ar, const double* Ai, bool conjA, const double* Br, const double* Bi,
           bool conjB, double* Cr, double* Ci);
void cgemm(Ctx* ctx, GemmPlan& p, const Tensor& A, bool conjA, const Tensor& B, bool conjB, Tensor& C);
void run_transform_c(Ctx* ctx, TransformPlan& p, const Tensor& X, Tensor& Y, const double* W);
void fetch_scalars(Ctx* ctx, int n);                                                  // d_scalars -> h_scalars, sync
void fill_random(Ctx* ctx, Tensor& t, uint64_t seed);
void scale_index(Ctx* ctx, Tensor& t, int which, const double* w_dev);
void scale_rows_or_cols(Ctx* ctx, double* A, int64_t ld, int64_t R, int64_t C, const double* s, bool rows);

}  // namespace tnl

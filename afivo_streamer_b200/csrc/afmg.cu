// libafmg: host side of the B200-native FAS multigrid solver (C ABI in include/afmg.h).
//
// Mirrors the control flow of afivo/src/m_af_multigrid.f90 (mg_fas_fmg :137-180, mg_fas_vcycle
// :185-264, gsrb_boxes :648-687, update_coarse :691-738, set_coarse_phi_rhs :742-776, init_phi_rhs
// :779-799, correct_children :624-646) as sequences of kernel launches on one stream, replayed as
// CUDA graphs.  There is no CPU compute path: every cell-data operation is a kernel in
// kernels3d.cuh.
#include <cuda.h>  // driver API types only: the entry points are fetched through cudaGetDriverEntryPoint (no -lcuda)
#include <cuda_runtime.h>

#include <algorithm>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <tuple>
#include <vector>

#include "../../include/afmg.h"
#include "field.cuh"
#include "kernels2d.cuh"
#include "kernels3d.cuh"
#include "mega.cuh"

#include "afmg_builders.inc"
#include "builders_dev.cuh"

using namespace afmg;

namespace {

thread_local std::string g_create_error;

// device copy of the phase lists of the persistent-kernel segments of one cycle (mega.cuh)
struct MegaProgram {
  MegaPhase* d_phases = nullptr;
  MegaOp* d_ops = nullptr;
  int nphase = 0;
  std::vector<MegaOp> h_ops;        // host copies: labels for the per-phase profile
  std::vector<MegaPhase> h_phases;
};

struct Graph {
  cudaGraphExec_t exec = nullptr;
  int64_t launches = 0;
  std::vector<MegaProgram> progs;
};

struct ProfEntry {
  double ms = 0;
  int64_t calls = 0;
};

}  // namespace

namespace {
struct S2State;  // 2D solver state (afmg2d.inc)
struct FieldState;  // field from potential (afmg_field.inc)
}

namespace {
// one host thread per GPU of a single-process multi-GPU handle: runs the calls posted for its rank handle
struct RankWorker {
  std::thread th;
  std::mutex mu;
  std::condition_variable cv;
  std::function<int()> task;
  bool has_task = false, done = false, quit = false;
  int rc = 0;
  void loop() {
    std::unique_lock<std::mutex> lk(mu);
    for (;;) {
      cv.wait(lk, [&] { return has_task || quit; });
      if (quit) return;
      std::function<int()> t = std::move(task);
      has_task = false;
      lk.unlock();
      const int r = t();
      lk.lock();
      rc = r;
      done = true;
      cv.notify_all();
    }
  }
  void post(std::function<int()> t) {
    std::lock_guard<std::mutex> lk(mu);
    task = std::move(t);
    has_task = true;
    done = false;
    cv.notify_all();
  }
  int wait() {
    std::unique_lock<std::mutex> lk(mu);
    cv.wait(lk, [&] { return done; });
    return rc;
  }
};
}  // namespace

struct afmg_handle {
  afmg_opts o{};
  // ---- single-process multi-GPU front (afmg_opts.n_gpus > 1): no device state of its own, only the rank handles
  bool is_multi = false;
  std::vector<afmg_handle*> subs;
  std::vector<RankWorker*> workers;
  S2State* s2 = nullptr;
  FieldState* fs = nullptr;
  std::string err;
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool ev_valid = false;

  // ---- host copy of the tree, slot maps
  bool have_tree = false;
  int L = 0, nslots = 0, highest_id = 0;
  std::vector<int> lvl_off;  // [L+2]: slots of level l are [lvl_off[l], lvl_off[l+1])
  std::vector<int> id2slot, slot2id;
  std::vector<int> h_nbr, h_aux, h_nmat, h_parent, h_child0, h_coff, h_lvl;
  std::vector<int> h_ix;  // [nslots*3]
  std::vector<double> h_rmin;  // [nslots*3] box%r_min
  double* d_rmin = nullptr;
  // results of the last afmg_build_stencils_device (kept for afmg_built_stencils)
  double* d_bblob = nullptr;
  std::vector<int> b_ids, b_tags, b_meta;
  std::vector<int> npar;  // [L+2] number of boxes with children per level
  int nbc = 0, nrb = 0;
  std::vector<int> rb_lvl_off;  // [L+2] refinement-boundary faces per level (prefix)
  std::vector<int> h_rb_slot, h_rb_face;
  std::vector<char> bc_set;
  std::vector<int> h_bc_type;      // [nbc]
  std::vector<int> bc_slot, bc_face;  // [nbc]
  int box_len = 0, nc2 = 0;

  // ---- device data
  double* d_cc[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // phi, rhs, tmp, (eps: field state), field norm
  int *d_nbr = nullptr, *d_aux = nullptr, *d_nmat = nullptr, *d_parent = nullptr, *d_child0 = nullptr,
      *d_coff = nullptr, *d_lvl = nullptr, *d_rb_slot = nullptr, *d_rb_face = nullptr;
  double *d_coef = nullptr, *d_rule_c = nullptr, *d_rule_B = nullptr, *d_pcoef = nullptr;
  std::vector<double> h_bc_B, h_bc_c;  // host mirrors of the physical-boundary rows of rule_B / rule_c (afmg_set_bc)
  unsigned long long* d_scal = nullptr;  // [0] fused residual max, [1] generic max, [2] mean (double bits)
  double* d_boxsum = nullptr;
  double* d_stage = nullptr;  // two halves: chunk c uses half c & 1 (copy of c + 1 overlaps the kernel of c)
  size_t stage_bytes = 0;
  cudaStream_t copy_stream = nullptr;
  cudaStream_t side_stream = nullptr;  // the generic-stencil kernels of a level run beside the fast ones (fork / join)
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  cudaStream_t launch_stream = nullptr;  // where launch_k puts the next kernel (the solver stream unless forked)
  int side_priority = 0;
  int grad_grid = 0;             // CTAs of the persistent gradient kernel k_grad3p (field.cuh); 0: not available
  bool grad_simple = false;      // AFMG_GRAD_SIMPLE=1: one box per CTA (k_grad3) for every box
  cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_consumed[2] = {nullptr, nullptr};
  int* d_stage_slots = nullptr;
  size_t stage_slots_n = 0;
  DevCtx cx{};

  // ---- coarse solver
  bool cs_ready = false;
  CoarseCtx cs{};
  double *d_b2r = nullptr, *d_Q[3] = {nullptr, nullptr, nullptr}, *d_inveig = nullptr, *d_v0 = nullptr, *d_v1 = nullptr;
  int* d_cs_bix = nullptr;

  // ---- explicit stencils (afmg_set_stencils)
  bool have_stencils = false;
  std::vector<unsigned char> h_opk, h_pk, h_rule_flag;
  std::vector<long long> h_opoff, h_foff, h_poff;
  std::vector<int> h_tag;
  unsigned char *d_opk = nullptr, *d_pk = nullptr, *d_rule_flag = nullptr;
  long long *d_opoff = nullptr, *d_foff = nullptr, *d_poff = nullptr;
  double* d_stv = nullptr;
  std::map<int, std::pair<std::vector<double>, std::vector<double>>> l1_st;  // level-1 slot -> (7 per cell, f)
  std::vector<int> spec_off;  // [L+2] prefix over levels of this rank's boxes with an explicit operator
  int* d_spec = nullptr;
  // general coarse solve (explicit stencils on level 1): dense inverse of the BC-folded matrix
  bool cs_dense = false;
  // per-cell level-set boundary values (afmg_set_lsf_boundary_values)
  std::vector<int> bv_ids;
  double* d_bv = nullptr;
  long long* d_bvoff = nullptr;
  double *d_Ainv = nullptr, *d_lsf_fac = nullptr;
  bool cs_planes = false;  // block-tridiagonal variant of the general coarse solve (large coarse grids)
  double *d_Sinv = nullptr, *d_cs_lo = nullptr, *d_cs_up = nullptr, *d_cs_w = nullptr, *d_cs_xs = nullptr, *d_cs_t = nullptr;

  // ---- multi-GPU (one process per GPU; peers' arrays mapped through CUDA IPC)
  int nranks = 1, me = 0;
  bool connected = false;
  std::vector<int> cut;  // [(L+2) * (nranks+1)]: slots [cut[l][r], cut[l][r+1]) of level l belong to rank r
  std::vector<int> rb_cut;  // same for the refinement-boundary face list (rb rows of level l)
  std::vector<unsigned char> h_owner;
  std::vector<char> lvl_multi;   // [L+2] 1: boxes of the level are owned by more than one rank
  bool unsynced = false;         // operations ran since the last cross-GPU barrier (all of them on rank 0's levels)
  int barrier_lean = 1;          // AFMG_BARRIER_LEAN=0: the round-1 barrier with explicit system fences around the release
  int min_split_boxes = -1;      // levels with fewer boxes stay on rank 0 (-1: 4 Mi cells worth of boxes)
  unsigned char* d_owner = nullptr;
  CommBlock* d_comm = nullptr;
  CommPeers peers{};
  char* d_slab = nullptr;  // phi | rhs | tmp | box sums, one allocation so that one IPC handle covers it
  size_t slab_bytes = 0, slab_var_stride = 0;
  int slab_nvar = 3;
  char* peer_slab[AFMG_MAX_RANKS] = {};
  bool local_peers = false;  // the peers are rank handles of the same process (afmg_opts.n_gpus): no CUDA IPC
  int dev_base = 0;          // first device of the single-process group
  // slab trimmed to the owned boxes (single-process multi-GPU): a virtual range for the whole slot space, physical
  // memory mapped only under the slots this rank owns (CUDA virtual memory management API)
  bool slab_vmm = false;
  size_t slab_va_size = 0, slab_mapped_bytes = 0;
  std::vector<std::pair<size_t, size_t>> slab_maps;        // (offset, size) of the mapped pieces
  std::vector<unsigned long long> slab_map_handles;        // CUmemGenericAllocationHandle of each piece
  unsigned long long barrier_timeout_ns = 30ull * 1000000000ull;

  // ---- persistent-kernel segments (mega.cuh)
  bool mega_enabled = false;     // opt-in (AFMG_MEGA=1 / afmg_set_mega): measured slower than the graph of launches, see mega.cuh
  int mega_max_boxes = 0;        // levels with at most this many boxes run inside k_mega (0: default per n_cell)
  int mega_grid = 0;             // co-resident CTAs of k_mega on this device
  int mega_cluster = 0;          // > 0: k_mega runs as ONE thread-block cluster of this many CTAs (hardware barrier)
  size_t mega_smem = 0;
  int mega_mode = 0;             // 0 off, 1 plan (record phases, launch nothing), 2 exec (launch recorded programs)
  bool mega_open = false;        // a segment is being recorded / waiting to be launched
  std::vector<MegaOp> rec_ops;
  std::vector<MegaPhase> rec_phases;
  std::vector<MegaProgram>* progs = nullptr;  // programs of the cycle being planned / executed
  size_t prog_idx = 0;
  std::vector<MegaProgram> direct_progs;      // programs of the last direct (no graph) run
  MegaSync* d_msync = nullptr;
  bool mega_launched = false;                 // since the last check of d_msync->err
  unsigned long long mega_timeout_ns = 2ull * 1000000000ull;
  unsigned long long* d_stamps = nullptr;     // per-phase time stamps (profiling mode)
  int stamps_cap = 0;
  std::vector<std::pair<size_t, const MegaProgram*>> stamp_runs;  // (offset in d_stamps, program) of profiled launches
  int stamps_used = 0;

  // ---- state
  bool resid_fresh = false;
  std::map<std::tuple<int, int, int>, Graph> graphs;
  bool capturing = false;
  int64_t launches = 0;
  bool profiling = false;
  bool pdl = false;  // programmatic dependent launch (launch_k), AFMG_PDL=1
  bool gsrb_fused_gen = false;  // AFMG_GSRB_FUSED_GEN=1: levels with explicit-stencil boxes take ONE fused half-sweep launch
                                // (k_gsrb2g) instead of a fast and a generic one side by side; measured slower (S2e: 1.06 vs
                                // 0.95 ms per V-cycle: 80 registers and per-cell coefficient loads on the critical path)
  int n_sm = 148;
  bool gsrb_wide = true;  // AFMG_GSRB_WIDE=0: never use the 8-boxes-per-CTA half-sweep (gsrb_one_wave)
  int small_ctas = 148;  // launches of at most this many CTAs take the latency-optimised kernel variants (AFMG_SMALL_CTAS)
  bool cs_fused = true;  // single-CTA coarse solve for small separable coarse grids (AFMG_CS_FUSED=0: off)
  std::map<std::string, ProfEntry> prof;
  std::vector<std::tuple<std::string, cudaEvent_t, cudaEvent_t>> prof_pending;

  int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    err = buf;
    return code;
  }
};

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      return h->fail(AFMG_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

namespace {

// run fn(rank handle) on every GPU of a single-process multi-GPU handle, each on its own host thread (the calls block
// on device-side barriers that need all ranks in flight); first error wins
template <class F>
int multi_call(afmg_handle* h, F fn) {
  const int n = (int)h->subs.size();
  for (int r = 0; r < n; ++r) {
    afmg_handle* sub = h->subs[r];
    h->workers[r]->post([fn, sub]() -> int { return fn(sub); });
  }
  int rc = 0;
  for (int r = 0; r < n; ++r) {
    const int q = h->workers[r]->wait();
    if (q && !rc) {
      rc = q;
      h->err = "GPU " + std::to_string(r) + ": " + h->subs[r]->err;
    }
  }
  return rc;
}
#define AFMG_MULTI(h, call) \
  if ((h) && (h)->is_multi) return multi_call((h), [=](afmg_handle* sub_) -> int { return call; })

template <class T>
int dev_upload(afmg_handle* h, T** dptr, const std::vector<T>& v) {
  if (*dptr) cudaFree(*dptr);
  *dptr = nullptr;
  size_t n = std::max<size_t>(v.size(), 1);
  CK(cudaMalloc((void**)dptr, n * sizeof(T)));
  if (!v.empty()) CK(cudaMemcpy(*dptr, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return AFMG_OK;
}

inline uint64_t morton3(uint32_t x, uint32_t y, uint32_t z) {
  auto spread = [](uint64_t v) {
    v &= 0x1fffff;
    v = (v | v << 32) & 0x1f00000000ffffULL;
    v = (v | v << 16) & 0x1f0000ff0000ffULL;
    v = (v | v << 8) & 0x100f00f00f00f00fULL;
    v = (v | v << 4) & 0x10c30c30c30c30c3ULL;
    v = (v | v << 2) & 0x1249249249249249ULL;
    return v;
  };
  return spread(x) | (spread(y) << 1) | (spread(z) << 2);
}

// cyclic Jacobi eigenvalue iteration for a symmetric n x n matrix (row-major); on return A's
// diagonal holds the eigenvalues and V (row-major, V[r*n+c]) the eigenvectors as columns
void jacobi_eig(int n, std::vector<long double>& A, std::vector<long double>& V) {
  V.assign((size_t)n * n, 0.0L);
  for (int i = 0; i < n; ++i) V[(size_t)i * n + i] = 1.0L;
  for (int sweep = 0; sweep < 100; ++sweep) {
    long double off = 0, diag = 0;
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) (i == j ? diag : off) += A[(size_t)i * n + j] * A[(size_t)i * n + j];
    if (off <= 1e-40L * diag) break;
    for (int p = 0; p < n - 1; ++p)
      for (int q = p + 1; q < n; ++q) {
        long double apq = A[(size_t)p * n + q];
        if (apq == 0.0L) continue;
        long double app = A[(size_t)p * n + p], aqq = A[(size_t)q * n + q];
        long double theta = (aqq - app) / (2 * apq);
        long double t = (theta >= 0 ? 1.0L : -1.0L) / (fabsl(theta) + sqrtl(theta * theta + 1));
        long double c = 1 / sqrtl(t * t + 1), s = t * c;
        for (int k = 0; k < n; ++k) {
          long double akp = A[(size_t)k * n + p], akq = A[(size_t)k * n + q];
          A[(size_t)k * n + p] = c * akp - s * akq;
          A[(size_t)k * n + q] = s * akp + c * akq;
        }
        for (int k = 0; k < n; ++k) {
          long double apk = A[(size_t)p * n + k], aqk = A[(size_t)q * n + k];
          A[(size_t)p * n + k] = c * apk - s * aqk;
          A[(size_t)q * n + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < n; ++k) {
          long double vkp = V[(size_t)k * n + p], vkq = V[(size_t)k * n + q];
          V[(size_t)k * n + p] = c * vkp - s * vkq;
          V[(size_t)k * n + q] = s * vkp + c * vkq;
        }
      }
  }
}

// ---- kernel launch plumbing -------------------------------------------------------------------
struct Launch {
  afmg_handle* h;
  std::string name;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  Launch(afmg_handle* h_, const char* name_, int lvl = 0) : h(h_) {
    if (h->mega_mode == 1) return;  // planning pass of the persistent-kernel segments: nothing is launched
    if (h->profiling && !h->capturing) name = lvl > 0 ? std::string(name_) + "_L" + std::to_string(lvl) : std::string(name_);
    h->launches++;
    if (h->profiling && !h->capturing) {
      cudaEventCreate(&e0);
      cudaEventCreate(&e1);
      cudaEventRecord(e0, h->stream);
    }
  }
  ~Launch() {
    if (e0) {
      cudaEventRecord(e1, h->stream);
      h->prof_pending.emplace_back(name, e0, e1);
    }
  }
};

// All kernels go through this helper.  With programmatic dependent launch (PDL; AFMG_PDL=1 enables
// it) the next kernel of the stream / graph is allowed to start launching while the previous one
// drains; every kernel begins with griddepcontrol.wait (pdl_wait()), which returns once the preceding
// grid has completed and its memory is visible, so the data dependencies are the same as with plain
// stream order -- only the launch latency could overlap.  Measured on B200 inside CUDA graphs it gains
// < 1 % (S2: 0.7045 vs 0.7090 ms per V-cycle; graph replay already hides the launch), so it is off by
// default; the ~3.3 us per graph node that bound the coarse levels are kernel drain + ramp-up.
template <class... KArgs, class... Args>
void launch_k(afmg_handle* h, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, Args&&... args) {
  if (h->mega_mode == 1) return;  // planning pass (see mega_route)
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = h->launch_stream ? h->launch_stream : h->stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (h->pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (h->launch_stream) {  // side-stream kernels keep their priority as nodes of a captured graph
    attr[na].id = cudaLaunchAttributePriority;
    attr[na].val.priority = h->side_priority;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  cudaLaunchKernelEx(&cfg, kern, KArgs(std::forward<Args>(args))...);
}

void prof_resolve(afmg_handle* h) {
  for (auto& t : h->prof_pending) {
    float ms = 0;
    cudaEventSynchronize(std::get<2>(t));
    cudaEventElapsedTime(&ms, std::get<1>(t), std::get<2>(t));
    auto& e = h->prof[std::get<0>(t)];
    e.ms += ms;
    e.calls++;
    cudaEventDestroy(std::get<1>(t));
    cudaEventDestroy(std::get<2>(t));
  }
  h->prof_pending.clear();
}

#define DISPATCH_NC(h, NCVAR, ...)                                      \
  switch ((h)->o.n_cell) {                                              \
    case 4: { constexpr int NCVAR = 4; __VA_ARGS__; } break;            \
    case 8: { constexpr int NCVAR = 8; __VA_ARGS__; } break;            \
    case 16: { constexpr int NCVAR = 16; __VA_ARGS__; } break;          \
    default: break;                                                     \
  }

inline int nlev(const afmg_handle* h, int l) { return h->lvl_off[l + 1] - h->lvl_off[l]; }

// slots of level l this rank computes (all of them on a single GPU)
struct Range {
  int s0, n;
};
inline Range own(const afmg_handle* h, int l) {
  const int* c = &h->cut[(size_t)l * (h->nranks + 1)];
  return {c[h->me], c[h->me + 1] - c[h->me]};
}

// ---- persistent-kernel segments (mega.cuh) ----------------------------------------------------------
// The enq_* functions below are the single description of a cycle.  With the persistent kernel enabled a cycle is
// walked twice: a PLAN pass (mega_mode 1) in which operations on small levels are recorded as phases of a program
// instead of being launched (and nothing else is launched either), and an EXEC pass (mega_mode 2, usually under
// graph capture) in which those operations are skipped and each recorded program is launched as one k_mega where
// its segment ends (mega_flush).  Both passes take identical decisions (mega_route), so programs and launch sites
// pair up by index.
inline int mega_items(const afmg_handle* h, int kind) {
  const bool big = h->o.n_cell == 16;
  switch (kind) {
    case MK_GSRB: return big ? MegaCfg<16>::GS_BPC : MegaCfg<8>::GS_BPC;
    case MK_RB: return big ? MegaCfg<16>::RB_PER : MegaCfg<8>::RB_PER;
    case MK_EC: return big ? MegaCfg<16>::EC_PER : MegaCfg<8>::EC_PER;
    case MK_RESTRICT:
    case MK_RESID: return big ? MegaCfg<16>::RES_PER : MegaCfg<8>::RES_PER;
    case MK_CORRECT: return big ? Correct3Cfg<16>::BPC : Correct3Cfg<8>::BPC;
    default: return 1;
  }
}

inline bool mega_possible(const afmg_handle* h) {
  return h->mega_enabled && h->mega_grid > 0 && h->o.ndim == 3 && h->nranks == 1 && !h->have_stencils &&
         !h->o.subtract_mean && (h->o.n_cell == 8 || h->o.n_cell == 16);
}

void free_programs(std::vector<MegaProgram>& v) {
  for (auto& p : v) {
    cudaFree(p.d_phases);
    cudaFree(p.d_ops);
  }
  v.clear();
}

inline void mega_end_phase(afmg_handle* h) {
  if (h->mega_mode != 1) return;
  const int op0 = h->rec_phases.empty() ? 0 : h->rec_phases.back().op0 + h->rec_phases.back().nops;
  const int nops = (int)h->rec_ops.size() - op0;
  if (nops == 0) return;
  if (nops > MEGA_MAX_OPS) {  // cannot happen with the enq_* functions as written; never launch a broken program
    h->rec_ops.resize(op0);
    h->mega_enabled = false;
    return;
  }
  MegaPhase ph{op0, nops, 0, 0};
  for (int q = op0; q < op0 + nops; ++q) ph.nvb += h->rec_ops[q].nvb;
  h->rec_phases.push_back(ph);
}

// record one operation of the open phase (plan pass only); nvb < 0: ceil(n / items per block of that kind)
inline void mega_op(afmg_handle* h, int kind, int lvl, int s0, int n, int a0 = 0, int a1 = 0, int a2 = 0, int a3 = 0,
                    int nvb = -1) {
  if (h->mega_mode != 1 || n <= 0) return;
  const int per = mega_items(h, kind);
  MegaOp op{kind, lvl, s0, n, nvb >= 0 ? nvb : (n + per - 1) / per, a0, a1, a2, a3, 0};
  if (op.nvb > 0) h->rec_ops.push_back(op);
}

// an operation that only has to precede the NEXT phase joins the phase that was closed last, if there is one
inline void mega_op_prev_phase(afmg_handle* h, int kind, int lvl, int s0, int n, int a0 = 0) {
  if (h->mega_mode != 1) return;
  mega_op(h, kind, lvl, s0, n, a0);
  if (h->rec_phases.empty()) {
    mega_end_phase(h);
  } else {
    h->rec_phases.back().nops += 1;
    h->rec_phases.back().nvb += h->rec_ops.back().nvb;
  }
}

void mega_flush(afmg_handle* h) {
  if (!h->mega_open) return;
  h->mega_open = false;
  if (h->mega_mode == 1) {
    mega_end_phase(h);
    MegaProgram pr;
    pr.nphase = (int)h->rec_phases.size();
    if (pr.nphase > 0) {
      cudaMalloc((void**)&pr.d_phases, h->rec_phases.size() * sizeof(MegaPhase));
      cudaMalloc((void**)&pr.d_ops, h->rec_ops.size() * sizeof(MegaOp));
      cudaMemcpy(pr.d_phases, h->rec_phases.data(), h->rec_phases.size() * sizeof(MegaPhase), cudaMemcpyHostToDevice);
      cudaMemcpy(pr.d_ops, h->rec_ops.data(), h->rec_ops.size() * sizeof(MegaOp), cudaMemcpyHostToDevice);
    }
    pr.h_ops.swap(h->rec_ops);
    pr.h_phases.swap(h->rec_phases);
    h->rec_ops.clear();
    h->rec_phases.clear();
    h->progs->push_back(std::move(pr));
    return;
  }
  if (h->mega_mode != 2 || !h->progs || h->prog_idx >= h->progs->size()) return;
  const MegaProgram& pr = (*h->progs)[h->prog_idx++];
  if (pr.nphase == 0) return;
  int max_nvb = 1;
  for (const auto& ph : pr.h_phases) max_nvb = std::max(max_nvb, ph.nvb);
  Launch L_(h, "mega");
  unsigned long long* stamps = nullptr;
  if (h->profiling && !h->capturing && h->d_stamps && h->stamps_used + pr.nphase + 1 <= h->stamps_cap) {
    stamps = h->d_stamps + h->stamps_used;
    h->stamp_runs.emplace_back((size_t)h->stamps_used, &pr);
    h->stamps_used += pr.nphase + 1;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = h->mega_smem;
  cfg.stream = h->stream;
  cudaLaunchAttribute attr[1];
  const int cluster = h->mega_cluster;
  if (cluster > 0) {  // the grid is one cluster: co-scheduled by construction, hardware barrier between the phases
    int g = 1;
    while (g < cluster && g < max_nvb) g *= 2;
    cfg.gridDim = dim3(g);
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = g;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
  } else {
    cfg.gridDim = dim3(std::min(h->mega_grid, max_nvb));
    attr[0].id = cudaLaunchAttributeCooperative;  // all CTAs co-resident, or the launch waits: the grid barrier cannot deadlock
    attr[0].val.cooperative = 1;
  }
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  h->mega_launched = true;
  if (h->o.n_cell == 16)
    cudaLaunchKernelEx(&cfg, k_mega<16>, h->cx, h->cs, (const MegaPhase*)pr.d_phases, (const MegaOp*)pr.d_ops, pr.nphase,
                       h->d_msync, h->d_scal, h->mega_timeout_ns, stamps, cluster > 0 ? 1 : 0);
  else
    cudaLaunchKernelEx(&cfg, k_mega<8>, h->cx, h->cs, (const MegaPhase*)pr.d_phases, (const MegaOp*)pr.d_ops, pr.nphase,
                       h->d_msync, h->d_scal, h->mega_timeout_ns, stamps, cluster > 0 ? 1 : 0);
}

// true: the operation (on a level of `nboxes` boxes) belongs to the persistent-kernel segment -- the caller records it
// (plan pass) or skips it (exec pass); false: it takes the launch path, after the open segment has been closed
inline bool mega_route(afmg_handle* h, int nboxes, bool ok = true) {
  if (h->mega_mode == 0) return false;
  int lim = h->mega_max_boxes > 0 ? h->mega_max_boxes : (h->o.n_cell == 16 ? 1024 : 8192);
  if (h->mega_cluster > 0 && h->mega_max_boxes <= 0) lim = h->mega_cluster * mega_items(h, MK_GSRB);  // one pass per phase
  if (ok && nboxes <= lim) {
    h->mega_open = true;
    return true;
  }
  mega_flush(h);
  return false;
}

// Boxes with explicit stencils are handled by generic kernels next to the fast ones; the two touch disjoint boxes, so
// inside a cycle they run concurrently: the generic launch goes to a side stream between a fork and a join event (in a
// captured graph: two parallel branches).  On the launch-bound streamer trees this hides one of the two node
// latencies per half-sweep (S2e, 597 of 8905 boxes with explicit stencils: 1.20 -> 0.95 ms per V-cycle, same bits).  Not in profiling mode, where every launch is timed on its own.
struct SideLaunch {
  afmg_handle* h;
  bool on;
  SideLaunch(afmg_handle* h_, bool want) : h(h_), on(want && !h_->profiling && h_->mega_mode != 1 && h_->side_stream) {
    if (!on) return;
    cudaEventRecord(h->ev_fork, h->stream);
    cudaStreamWaitEvent(h->side_stream, h->ev_fork, 0);
  }
  void begin() { if (on) h->launch_stream = h->side_stream; }
  void end() {
    if (!on) return;
    h->launch_stream = nullptr;
    cudaEventRecord(h->ev_join, h->side_stream);
    cudaStreamWaitEvent(h->stream, h->ev_join, 0);
  }
};

// Cross-GPU barrier between dependent kernels (no-op on one GPU, where stream order suffices).
// `multi` = the operation it follows involves a level whose boxes are spread over several ranks.  Operations on
// levels that live entirely on rank 0 (the coarse grid and the small levels the partition does not split) need no
// barrier between each other -- the reference itself drops to one thread there (m_af_multigrid.f90:276-290) -- but the
// next operation that involves other ranks must wait for them (pre_sync).  Every rank evaluates the same static
// rule, so all ranks enter the same number of barriers.
inline bool lvl_multi(const afmg_handle* h, int l) { return l >= 1 && l <= h->L && h->lvl_multi[l]; }
void enq_barrier(afmg_handle* h, bool multi = true) {
  if (h->nranks == 1) return;
  if (!multi) {
    h->unsynced = true;
    return;
  }
  Launch L_(h, "barrier");
  launch_k(h, k_barrier, 1, 32, 0, h->d_comm, h->peers, h->nranks, h->me, h->barrier_timeout_ns, h->barrier_lean);
  h->unsynced = false;
}
inline void pre_sync(afmg_handle* h, bool multi = true) {
  if (h->nranks > 1 && multi && h->unsynced) enq_barrier(h, true);
}

// one half-sweep + side ghost fill on level l
template <int NC>
struct Gsrb2Cfg {  // boxes per CTA, k-splits, min CTAs per SM
  static constexpr int BPC = (NC == 16) ? 1 : (NC == 8 ? 4 : 16);
  static constexpr int KS = (NC == 16) ? 2 : (NC == 8 ? 2 : 1);
  static constexpr int MINB = (NC == 16) ? 5 : 6;
};

template <class K>
void set_max_smem(K kernel, size_t smem) {
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}

inline int nspec(const afmg_handle* h, int l) { return h->have_stencils ? h->spec_off[l + 1] - h->spec_off[l] : 0; }

// k_gsrb2s exists for 8^3 and 16^3 boxes; blocks < 0: only opt in to its shared memory size (configure_kernels)
template <int NC>
bool gsrb_small(afmg_handle* h, int blocks, int threads, size_t smem, int s0, int n, int C, int l) {
  if constexpr (NC == 8 || NC == 16) {
    using G = Gsrb2Cfg<NC>;
    if (blocks < 0) {
      set_max_smem(k_gsrb2s<NC, G::BPC, G::KS>, smem);
      return true;
    }
    if (blocks <= h->small_ctas) {
      launch_k(h, k_gsrb2s<NC, G::BPC, G::KS>, blocks, threads, smem, h->cx, s0, n, C, l);
      return true;
    }
  }
  return false;
}

// 8^3 boxes: k_gsrb2<8, 4, 2> keeps 6 CTAs of 4 boxes per SM = 3552 boxes in flight on 148 SMs.  A level slightly
// larger than that (the finest level of a streamer tree: 3904 boxes) takes a second, nearly empty wave at full
// latency.  k_gsrb2<8, 8, 1> (32 threads per box, 8 boxes per CTA, 5 CTAs per SM by shared memory = 5920 boxes) does
// such a level in one wave.
template <int NC>
bool gsrb_one_wave(afmg_handle* h, int blocks, int s0, int n, int C, int l) {
  if constexpr (NC == 8) {
    constexpr int BPC = 8, KS = 1, MINB = 5;
    const size_t smem = (size_t)BPC * (Lay3<NC>::COL + Lay3<NC>::NI) * sizeof(double);
    if (blocks < 0) {
      set_max_smem(k_gsrb2<NC, BPC, KS, MINB>, smem);
      return true;
    }
    const int cap = h->n_sm * Gsrb2Cfg<NC>::MINB;
    if (h->gsrb_wide && blocks > cap && (n + BPC - 1) / BPC <= h->n_sm * MINB) {
      launch_k(h, k_gsrb2<NC, BPC, KS, MINB>, (n + BPC - 1) / BPC, BPC * KS * NC * NC / 2, smem, h->cx, s0, n, C, l);
      return true;
    }
  }
  return false;
}

// the fast half-sweep kernel for `n` slots from `s0` on level l (boxes with explicit stencils are skipped by it)
void launch_gsrb_range(afmg_handle* h, int l, int s0, int n, int C) {
  DISPATCH_NC(h, NC, {
    using G = Gsrb2Cfg<NC>;
    constexpr int threads = G::BPC * G::KS * NC * NC / 2;
    const size_t smem = (size_t)G::BPC * (Lay3<NC>::COL + Lay3<NC>::NI) * sizeof(double);
    const int blocks = (n + G::BPC - 1) / G::BPC;
    // latency-bound launch: the variant with plain loads / stores (k_gsrb2s); a launch just above one wave of
    // resident CTAs: half as many CTAs with twice the boxes each (one wave instead of two)
    if (!gsrb_small<NC>(h, blocks, threads, smem, s0, n, C, l) && !gsrb_one_wave<NC>(h, blocks, s0, n, C, l)) {
      auto kern = k_gsrb2<NC, G::BPC, G::KS, G::MINB>;
      launch_k(h, kern, blocks, threads, smem, h->cx, s0, n, C, l);
    }
  });
}

// Every enq_* below works on the slots of a level this rank owns and ends with a cross-GPU barrier
// when its results are read, or its inputs overwritten, by kernels of other ranks.
void enq_gsrb(afmg_handle* h, int l, int redblack) {
  const Range r = own(h, l);
  pre_sync(h, lvl_multi(h, l));
  if (mega_route(h, nlev(h, l))) {
    mega_op(h, MK_GSRB, l, r.s0, r.n, redblack & 1, l);
    mega_end_phase(h);
    return;
  }
  if (r.n > 0 && nspec(h, l) > 0 && h->gsrb_fused_gen) {
    // the level holds boxes with explicit stencils: one launch sweeps all of its boxes (k_gsrb2g)
    Launch L_(h, "gsrb", l);
    DISPATCH_NC(h, NC, {
      using G = Gsrb2Cfg<NC>;
      constexpr int threads = G::BPC * G::KS * NC * NC / 2;
      const size_t smem = (size_t)G::BPC * (Lay3<NC>::COL + Lay3<NC>::NI) * sizeof(double);
      launch_k(h, k_gsrb2g<NC, G::BPC, G::KS>, (r.n + G::BPC - 1) / G::BPC, threads, smem, h->cx, r.s0, r.n, redblack & 1, l);
    });
    enq_barrier(h, lvl_multi(h, l));
    return;
  }
  SideLaunch side(h, r.n > 0 && nspec(h, l) > 0);
  if (r.n > 0) {
    Launch L_(h, "gsrb", l);
    launch_gsrb_range(h, l, r.s0, r.n, redblack & 1);
  }
  if (const int ns = nspec(h, l)) {  // boxes with an explicit stencil (skipped by the kernel above)
    side.begin();
    {
      Launch L_(h, "gsrb_gen", l);
      DISPATCH_NC(h, NC, { launch_k(h, k_gsrb_gen<NC>, ns, 256, 0, h->cx, h->d_spec + h->spec_off[l], ns, redblack & 1); });
    }
    side.end();
  }
  enq_barrier(h, lvl_multi(h, l));
}

// reads the (frozen) coarse level, writes this rank's rule rows: no barrier needed afterwards
inline Range own_rb(const afmg_handle* h, int l) {  // this rank's refinement-boundary faces of level l
  if (l < 2 || l > h->L) return {0, 0};
  const int* c = &h->rb_cut[(size_t)l * (h->nranks + 1)];
  return {c[h->me], c[h->me + 1] - c[h->me]};
}
void enq_rb_prepare(afmg_handle* h, int l) {
  const Range rb = own_rb(h, l);
  const int r0 = rb.s0, n = rb.n;
  pre_sync(h, lvl_multi(h, l) || lvl_multi(h, l - 1));  // reads the coarse neighbours, possibly on peer GPUs
  if (n == 0) return;
  if (mega_route(h, nlev(h, l))) {
    mega_op(h, MK_RB, l, r0, n, V_PHI);
    mega_end_phase(h);
    return;
  }
  Launch L_(h, "rb_prepare", l);
  DISPATCH_NC(h, NC, { launch_k(h, k_rb_prepare<NC>, n, 128, 0, h->cx, r0, n, V_PHI); });
}

// af_gc_lvl (+ parent update when mode != 0)
void enq_gc(afmg_handle* h, int l, int var, int corners, int mode) {
  const Range r = own(h, l);
  pre_sync(h, lvl_multi(h, l));
  if (mega_route(h, nlev(h, l))) {
    if (var == V_PHI && mode != 0) mega_op(h, MK_GC2, l, r.s0, r.n, corners, mode);
    else mega_op(h, MK_GC, l, r.s0, r.n, var, corners);
    mega_end_phase(h);
    return;
  }
  if (r.n > 0) {
    Launch L_(h, mode ? "gc_parent" : "gc", l);
    DISPATCH_NC(h, NC, {
      if (var == V_PHI && mode != 0)
        launch_k(h, k_gc2<NC>, r.n, 256, (size_t)2 * Lay3<NC>::COL * sizeof(double), h->cx, r.s0, r.n, corners, mode);
      else
        launch_k(h, k_gc<NC>, r.n, 256, 0, h->cx, r.s0, r.n, var, corners, mode);
    });
  }
  enq_barrier(h, lvl_multi(h, l));
}

// rb_lvl > 0: the refinement-boundary interpolation of that (finer) level rides along as extra CTAs
void enq_edges_corners(afmg_handle* h, int l, int rb_lvl = 0) {
  const Range r = own(h, l);
  const Range rb = own_rb(h, rb_lvl);
  const bool multi = lvl_multi(h, l) || (rb_lvl > 0 && lvl_multi(h, rb_lvl));
  pre_sync(h, multi);
  if (mega_route(h, nlev(h, l))) {
    mega_op(h, MK_EC, l, r.s0, r.n, V_PHI);
    mega_op(h, MK_RB, rb_lvl, rb.s0, rb.n, V_PHI);
    mega_end_phase(h);
    return;
  }
  if (r.n > 0) {
    Launch L_(h, "edges_corners", l);
    DISPATCH_NC(h, NC, { launch_k(h, k_edges_corners<NC>, r.n + rb.n, 64, 0, h->cx, r.s0, r.n, V_PHI, rb.s0, rb.n); });
  } else if (rb.n > 0) {
    enq_rb_prepare(h, rb_lvl);
  }
  enq_barrier(h, multi);
}

template <int NC>
struct OpCfg {
  static constexpr int KS = (NC == 4) ? 1 : 2;           // k-splits of k_resid3 (k-range must stay even)
  static constexpr int RES_MINB = (NC == 16) ? 3 : 6;
  static constexpr size_t TILE = (size_t)2 * Lay3<NC>::COL * sizeof(double);
};

void enq_restrict(afmg_handle* h, int l, int keep_res, int rb_lvl = 0) {
  const Range r = own(h, l);
  const Range rb = own_rb(h, rb_lvl);
  const bool multi = lvl_multi(h, l) || lvl_multi(h, l - 1) || (rb_lvl > 0 && (lvl_multi(h, rb_lvl) || lvl_multi(h, rb_lvl - 1)));
  pre_sync(h, multi);
  if (mega_route(h, nlev(h, l))) {
    mega_op(h, MK_RESTRICT, l, r.s0, r.n, keep_res);
    mega_op(h, MK_RB, rb_lvl, rb.s0, rb.n, V_PHI);
    mega_end_phase(h);
    return;
  }
  SideLaunch side(h, r.n > 0 && nspec(h, l) > 0);
  if (r.n > 0) {
    Launch L_(h, "restrict", l);
    DISPATCH_NC(h, NC, {
      constexpr int KS = OpCfg<NC>::KS;
      launch_k(h, k_resid3<NC, KS, 1, OpCfg<NC>::RES_MINB>, r.n + rb.n, KS * NC * NC / 2, OpCfg<NC>::TILE,
          h->cx, r.s0, r.n, nullptr, keep_res, rb.s0, rb.n);
    });
  } else if (rb.n > 0) {
    enq_rb_prepare(h, rb_lvl);
  }
  if (const int ns = nspec(h, l)) {
    side.begin();
    {
      Launch L_(h, "restrict_gen", l);
      DISPATCH_NC(h, NC, {
        launch_k(h, k_resid_gen<NC, 1>, ns, 256, (size_t)2 * Lay3<NC>::NI * sizeof(double),
            h->cx, h->d_spec + h->spec_off[l], ns, nullptr, keep_res);
      });
    }
    side.end();
  }
  enq_barrier(h, multi);
}

// correct_children; with push the side ghost cells of the children are filled as well (the caller
// must have run enq_rb_prepare(lp + 1) before and runs enq_edges_corners(lp + 1) after).  One CTA per
// child box (all boxes of level lp + 1 have a parent on level lp).
void enq_correct(afmg_handle* h, int lp, bool store_corr, bool push) {
  if (lp >= h->L || h->npar[lp] == 0) return;
  const Range rc = own(h, lp + 1);
  const bool multi = lvl_multi(h, lp) || lvl_multi(h, lp + 1);
  pre_sync(h, multi);
  if (mega_route(h, nlev(h, lp + 1))) {
    mega_op(h, MK_CORRECT, lp, rc.s0, rc.n, push ? 1 : 0);
    mega_end_phase(h);
  } else if (rc.n > 0) {
    Launch L_(h, "correct", lp);
    DISPATCH_NC(h, NC, {
      using CC = Correct3Cfg<NC>;
      const size_t smem = (size_t)CC::BPC * CC::SB * sizeof(double);
      launch_k(h, k_correct3<NC>, (rc.n + CC::BPC - 1) / CC::BPC, 256, smem, h->cx, rc.s0, rc.n, push ? 1 : 0);
    });
  }
  enq_barrier(h, multi);  // all children have read the old tmp of their parents
  if (store_corr) {
    const Range rp = own(h, lp);
    if (mega_route(h, nlev(h, lp))) {
      mega_op(h, MK_STORE_CORR, lp, rp.s0, rp.n);
      mega_end_phase(h);
    } else if (rp.n > 0) {
      Launch L_(h, "store_corr", lp);
      DISPATCH_NC(h, NC, { launch_k(h, k_store_corr<NC>, rp.n, 256, 0, h->cx, rp.s0, rp.n); });
    }
  }
}

// correct_children(lp) followed by af_gc_lvl(lp + 1) (m_af_multigrid.f90:219-222)
// Inside the cycles the edge / corner ghost cells written by that af_gc_lvl are dead: the upward
// gsrb_boxes that follows only reads face ghost cells and ends with its own edge / corner refresh
// (m_af_multigrid.f90:676-684), so they are skipped there (corners = false) unless n_cycle_up == 0.
void enq_correct_gc(afmg_handle* h, int lp, bool store_corr, bool corners = true, bool rb_done = false) {
  if (!rb_done) enq_rb_prepare(h, lp + 1);
  enq_correct(h, lp, store_corr, true);
  if (corners) enq_edges_corners(h, lp + 1);
}

// max over ranks of scal[idx] -> scal[4 + idx] on every rank
void enq_allmax(afmg_handle* h, int idx) {
  if (h->nranks == 1) return;
  enq_barrier(h);
  {
    Launch L_(h, "allmax");
    launch_k(h, k_allmax, 1, 1, 0, h->d_comm, h->peers, h->nranks, idx);
  }
  enq_barrier(h);  // nobody resets scal[idx] while a peer still reads it
}

// residual on levels l_lo..l_hi: purely local (reads own phi/rhs, writes own tmp)
void enq_residual(afmg_handle* h, int l_lo, int l_hi, bool with_max) {
  auto launch = [&](int s0, int n) {
    if (n == 0) return;
    Launch L_(h, "residual");
    DISPATCH_NC(h, NC, {
      constexpr int KS = OpCfg<NC>::KS;
      launch_k(h, k_resid3<NC, KS, 0, OpCfg<NC>::RES_MINB>, n, KS * NC * NC / 2, OpCfg<NC>::TILE,
          h->cx, s0, n, with_max ? h->d_scal : nullptr, 0, 0, 0);
    });
  };
  if (h->have_stencils) {
    const int s0 = h->spec_off[l_lo], ns = h->spec_off[l_hi + 1] - s0;
    if (ns > 0) {
      Launch L_(h, "residual_gen");
      DISPATCH_NC(h, NC, {
        launch_k(h, k_resid_gen<NC, 0>, ns, 256, 0, h->cx, h->d_spec + s0, ns, with_max ? h->d_scal : nullptr, 0);
      });
    }
  }
  if (h->nranks == 1) {
    const int s0 = h->lvl_off[l_lo], n = h->lvl_off[l_hi + 1] - h->lvl_off[l_lo];
    if (mega_route(h, n)) {
      mega_op(h, MK_RESID, l_hi, s0, n, with_max ? 1 : 0);
      mega_end_phase(h);
    } else {
      launch(s0, n);
    }
  } else {
    for (int l = l_lo; l <= l_hi; ++l) {
      const Range r = own(h, l);
      launch(r.s0, r.n);
    }
    if (with_max) enq_allmax(h, 0);
  }
}

// opt in to large dynamic shared memory / max carveout once per process (not a stream operation, but
// kept out of graph capture)
void configure_kernels(afmg_handle* h) {
  DISPATCH_NC(h, NC, {
    using G = Gsrb2Cfg<NC>;
    set_max_smem(k_gsrb2<NC, G::BPC, G::KS, G::MINB>, (size_t)G::BPC * (Lay3<NC>::COL + Lay3<NC>::NI) * sizeof(double));
    set_max_smem(k_gsrb2g<NC, G::BPC, G::KS>, (size_t)G::BPC * (Lay3<NC>::COL + Lay3<NC>::NI) * sizeof(double));
    gsrb_small<NC>(h, -1, 0, (size_t)G::BPC * (Lay3<NC>::COL + Lay3<NC>::NI) * sizeof(double), 0, 0, 0, 0);
    gsrb_one_wave<NC>(h, -1, 0, 0, 0, 0);
    set_max_smem(k_resid3<NC, OpCfg<NC>::KS, 0, OpCfg<NC>::RES_MINB>, OpCfg<NC>::TILE);
    set_max_smem(k_resid3<NC, OpCfg<NC>::KS, 1, OpCfg<NC>::RES_MINB>, OpCfg<NC>::TILE);
    set_max_smem(k_gc2<NC>, OpCfg<NC>::TILE);
    set_max_smem(k_correct3<NC>, (size_t)Correct3Cfg<NC>::BPC * Correct3Cfg<NC>::SB * sizeof(double));
  });
  // persistent kernel: grid = what is co-resident on this device (cooperative launch)
  h->mega_grid = 0;
  int nsm = 0, per_sm = 0;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, h->device);
  if (h->o.n_cell == 16) {
    h->mega_smem = MegaCfg<16>::smem_bytes(1024);
    set_max_smem(k_mega<16>, h->mega_smem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_mega<16>, 256, h->mega_smem);
  } else if (h->o.n_cell == 8) {
    h->mega_smem = MegaCfg<8>::smem_bytes(1024);
    set_max_smem(k_mega<8>, h->mega_smem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_mega<8>, 256, h->mega_smem);
  }
  int coop = 0;
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, h->device);
  if (coop && per_sm > 0) h->mega_grid = nsm * per_sm;
  if (const char* env = getenv("AFMG_MEGA_GRID")) {
    const int g = atoi(env);
    if (g > 0 && g < h->mega_grid) h->mega_grid = g;
  }
  if (h->mega_cluster > 8) {  // more than 8 CTAs per cluster is opt-in on the function
    cudaError_t e = (h->o.n_cell == 16) ? cudaFuncSetAttribute(k_mega<16>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1)
                                        : cudaFuncSetAttribute(k_mega<8>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) h->mega_cluster = 8;
  }
  cudaGetLastError();
}

void enq_copy_lvl(afmg_handle* h, int l, int dst, int src) {
  const Range r = own(h, l);
  const size_t n = (size_t)r.n * h->box_len;
  if (n == 0) return;
  if (mega_route(h, nlev(h, l))) {
    mega_op(h, MK_COPY, l, r.s0, r.n, dst, src);
    mega_end_phase(h);
    return;
  }
  Launch L_(h, "copy", l);
  const size_t off = (size_t)r.s0 * h->box_len;
  const int blocks = (int)std::min<size_t>((n + 1023) / 1024, 148 * 8);
  launch_k(h, k_copy, blocks, 256, 0, h->d_cc[dst] + off, h->d_cc[src] + off, n);
}

// solve_coarse_grid (m_af_multigrid.f90:266-291)
void enq_coarse(afmg_handle* h) {
  if (h->me != 0) {  // level 1 lives on rank 0; the others only keep the barrier bookkeeping in step
    enq_barrier(h, false);
    return;
  }
  const int nbox1 = nlev(h, 1);
  const int ntot = h->cs.nx[0] * h->cs.nx[1] * h->cs.nx[2];
  const int blocks = (ntot + 127) / 128;
  if (mega_route(h, nbox1, !h->cs_dense)) {
    const int nvb = (ntot + 255) / 256;
    if (ntot <= 1024 && h->cs_fused) {
      const bool with_gc = nbox1 <= 8;
      mega_op(h, MK_CS_FUSED, 1, 0, nbox1, with_gc ? 1 : 0, 0, 0, 0, 1);
      mega_end_phase(h);
      if (!with_gc) enq_gc(h, 1, V_PHI, 1, 0);
      return;
    }
    mega_op(h, MK_CS_GATHER, 1, 0, nbox1, 0, 0, 0, 0, nvb);
    mega_end_phase(h);
    for (int q = 0; q < 6; ++q) {  // three forward transforms (the last one scales by 1 / eigenvalue), three backward
      mega_op(h, MK_CS_APPLY, 1, 0, ntot, q % 3, q < 3 ? 1 : 0, q == 2 ? 1 : 0, q & 1, nvb);
      mega_end_phase(h);
    }
    mega_op(h, MK_CS_SCATTER, 1, 0, nbox1, 0, 0, 0, 0, nvb);
    mega_end_phase(h);
    enq_gc(h, 1, V_PHI, 1, 0);
    return;
  }
  // one CTA does it all, ghost cells of level 1 included; measured: 8^3 cells 20.6 -> 17 us per solve, but 16^3
  // 29 -> 80 us (one SM against 32 CTAs), so only the small coarse grids of the streamer configurations take it
  if (!h->cs_dense && ntot <= 1024 && h->cs_fused) {
    const bool with_gc = nbox1 <= 8;
    {
      Launch L_(h, "coarse");
      DISPATCH_NC(h, NC, {
        launch_k(h, k_cs_fused<NC>, 1, 1024, (size_t)2 * ntot * sizeof(double), h->cx, h->cs, nbox1, with_gc ? 1 : 0);
      });
    }
    if (with_gc) enq_barrier(h, false);
    else enq_gc(h, 1, V_PHI, 1, 0);
    return;
  }
  {
    Launch L_(h, "coarse");
    DISPATCH_NC(h, NC, { launch_k(h, k_cs_gather<NC>, blocks, 128, 0, h->cx, h->cs, nbox1); });
  }
  double *a = h->d_v0, *b = h->d_v1;
  if (h->cs_planes) {
    const int m = h->cs.m, np = h->cs.np, vb = (m + 127) / 128, mb = (m * 32 + 255) / 256;
    for (int p = 0; p < np; ++p) {  // forward: w_p = S_p^-1 (b_p - lo_p w_{p-1})
      Launch L_(h, "coarse");
      launch_k(h, k_cs_plane_rhs, vb, 128, 0, h->cs, (const double*)a, p, 0);
      launch_k(h, k_cs_plane_matvec, mb, 256, 0, h->cs, p, 0);
    }
    {
      Launch L_(h, "coarse");
      launch_k(h, k_cs_plane_rhs, vb, 128, 0, h->cs, (const double*)a, np - 1, 2);
    }
    for (int p = np - 2; p >= 0; --p) {  // backward: x_p = w_p - S_p^-1 (up_p x_{p+1})
      Launch L_(h, "coarse");
      launch_k(h, k_cs_plane_rhs, vb, 128, 0, h->cs, (const double*)a, p, 1);
      launch_k(h, k_cs_plane_matvec, mb, 256, 0, h->cs, p, 1);
    }
    {
      Launch L_(h, "coarse");
      launch_k(h, k_cs_plane_out, blocks, 128, 0, h->cs, b);
    }
    std::swap(a, b);
  } else if (h->cs_dense) {
    Launch L_(h, "coarse");
    launch_k(h, k_cs_dense, (ntot * 32 + 255) / 256, 256, 0, h->cs, a, b);
    std::swap(a, b);
  } else {
    for (int d = 0; d < 3; ++d) {
      Launch L_(h, "coarse");
      launch_k(h, k_cs_apply, blocks, 128, 0, h->cs, a, b, d, 1, d == 2);
      std::swap(a, b);
    }
    for (int d = 0; d < 3; ++d) {
      Launch L_(h, "coarse");
      launch_k(h, k_cs_apply, blocks, 128, 0, h->cs, a, b, d, 0, 0);
      std::swap(a, b);
    }
  }
  {
    Launch L_(h, "coarse");
    DISPATCH_NC(h, NC, { launch_k(h, k_cs_scatter<NC>, blocks, 128, 0, h->cx, h->cs, nbox1, a); });
  }
  enq_gc(h, 1, V_PHI, 1, 0);
}

// gsrb_boxes (m_af_multigrid.f90:648-687)
// rb_after > 0: also interpolate the refinement-boundary faces of level rb_after (= l + 1, on the way up) once
// level l has its final values; it rides on the last edges / corners launch
void enq_gsrb_boxes(afmg_handle* h, int l, bool up, int rb_after = 0) {
  const int ncyc = up ? h->o.n_cycle_up : h->o.n_cycle_down;
  bool rb_done = rb_after == 0;
  for (int n = 1; n <= 2 * ncyc; ++n) {
    enq_gsrb(h, l, n);
    const bool corners = h->o.use_corners || (up && n == 2 * ncyc);
    const bool last = n == 2 * ncyc;
    if (corners) enq_edges_corners(h, l, (last && !rb_done) ? rb_after : 0);
    if (corners && last) rb_done = true;
  }
  if (!rb_done) enq_rb_prepare(h, rb_after);
}

// update_coarse (:691-738) with_tmp = true, set_coarse_phi_rhs (:742-776) with_tmp = false
void enq_update_coarse(afmg_handle* h, int l, bool with_tmp) {
  if (!with_tmp && l == h->L) {
    enq_rb_prepare(h, l);
    enq_gc(h, l, V_PHI, 1, 0);
  }
  enq_restrict(h, l, with_tmp ? 0 : 1, l - 1);  // + the refinement-boundary interpolation of level l-1
  enq_gc(h, l - 1, V_PHI, 1, with_tmp ? 1 : 2);
}

void enq_subtract_mean(afmg_handle* h, int max_lvl);

// mg_fas_vcycle (m_af_multigrid.f90:185-264)
void enq_vcycle(afmg_handle* h, bool set_residual, int max_lvl, bool final_state = true) {
  const bool dead_corners = h->o.n_cycle_up > 0;
  for (int l = max_lvl; l >= 2; --l) {
    // below max_lvl the refinement-boundary rows of level l were interpolated by update_coarse(l + 1)
    // and level l - 1 has not changed since
    if (l == max_lvl) enq_rb_prepare(h, l);
    enq_gsrb_boxes(h, l, false);
    enq_update_coarse(h, l, true);
  }
  enq_coarse(h);
  for (int l = 2; l <= max_lvl; ++l) {
    // the interpolation for level l was issued with the upward sweeps of level l - 1 (l > 2)
    enq_correct_gc(h, l - 1, final_state && !set_residual, !dead_corners, l > 2);
    enq_gsrb_boxes(h, l, true, l < max_lvl ? l + 1 : 0);
  }
  if (set_residual) {
    const bool all = (max_lvl == h->L);
    if (all) {
      if (mega_route(h, h->lvl_off[max_lvl + 1])) mega_op_prev_phase(h, MK_CLEAR_SCAL, 0, 0, 1, 0);
      else if (h->mega_mode != 1) cudaMemsetAsync(h->d_scal, 0, sizeof(unsigned long long), h->stream);
    }
    enq_residual(h, 1, max_lvl, all);
  }
  if (h->o.subtract_mean) enq_subtract_mean(h, max_lvl);
}

__global__ void k_weighted_sum(const double* boxsum, const int* child0, const int* lvl, const double* lvl_fac, int n,
                               double inv_volume, double* out) {
  // deterministic single-CTA reduction: thread t sums slots t, t+T, ...; then a fixed tree
  pdl_wait();
  __shared__ double sh[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x)
    if (child0[i] < 0) s = s + lvl_fac[lvl[i]] * boxsum[i];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] = sh[threadIdx.x] + sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = sh[0] * inv_volume;
}

// per-box interior sums of this rank's boxes, written into the tables of all ranks
void enq_box_sums(afmg_handle* h, int var) {
  for (int l = 1; l <= h->L; ++l) {
    const Range r = (h->nranks == 1) ? Range{0, h->nslots} : own(h, l);
    if (r.n > 0) {
      Launch L_(h, "sum");
      DISPATCH_NC(h, NC, { launch_k(h, k_box_sums<NC>, r.n, 128, 0, h->cx, r.s0, r.n, var, h->d_boxsum); });
    }
    if (h->nranks == 1) break;
  }
  enq_barrier(h);
}

void enq_subtract_mean(afmg_handle* h, int max_lvl) {
  // af_tree_sum_cc over all leaves / af_total_volume, then phi -= mean on levels 1..max_lvl (full boxes)
  enq_box_sums(h, V_PHI);
  double vol = (double)nlev(h, 1);
  for (int d = 0; d < 3; ++d) vol *= h->o.n_cell * h->o.dr_base[d];
  {
    Launch L_(h, "sum");
    launch_k(h, k_weighted_sum, 1, 256, 0, h->d_boxsum, h->d_child0, h->d_lvl, h->d_coef + 8 * (h->L + 1), h->nslots,
                                            1.0 / vol, (double*)(h->d_scal + 2));
  }
  for (int l = 1; l <= max_lvl; ++l) {
    const Range r = (h->nranks == 1) ? Range{0, h->lvl_off[max_lvl + 1]} : own(h, l);
    const size_t n = (size_t)r.n * h->box_len;
    if (n > 0) {
      Launch L_(h, "sub_mean");
      const int blocks = (int)std::min<size_t>((n + 1023) / 1024, 148 * 8);
      launch_k(h, k_sub_scalar, blocks, 256, 0, h->d_cc[V_PHI] + (size_t)r.s0 * h->box_len,
                                                  (const double*)(h->d_scal + 2), n);
    }
    if (h->nranks == 1) break;
  }
  enq_barrier(h);
}

// init_phi_rhs (m_af_multigrid.f90:779-799)
void enq_init_phi_rhs(afmg_handle* h) {
  for (int l = h->L; l >= 2; --l) {
    const Range r = own(h, l);
    pre_sync(h, lvl_multi(h, l) || lvl_multi(h, l - 1));
    if (mega_route(h, nlev(h, l))) {
      mega_op(h, MK_RESTRICT_VAR, l, r.s0, r.n, V_RHS, 1);
      mega_end_phase(h);
    } else if (r.n > 0) {
      Launch L_(h, "init_phi_rhs", l);
      DISPATCH_NC(h, NC, {
        constexpr int T = (NC == 16) ? 256 : (NC == 8 ? 64 : 32);
        launch_k(h, k_restrict_var<NC>, r.n, T, 0, h->cx, r.s0, r.n, V_RHS, 1);
      });
    }
    enq_barrier(h, lvl_multi(h, l) || lvl_multi(h, l - 1));
  }
}

// mg_fas_fmg (m_af_multigrid.f90:137-180)
void enq_fmg(afmg_handle* h, bool set_residual, bool have_guess) {
  if (have_guess) {
    for (int l = h->L; l >= 2; --l) enq_update_coarse(h, l, false);
  } else {
    enq_init_phi_rhs(h);
  }
  enq_copy_lvl(h, 1, V_TMP, V_PHI);
  enq_vcycle(h, set_residual && h->L == 1, 1, h->L == 1);
  for (int l = 2; l <= h->L; ++l) {
    enq_copy_lvl(h, l, V_TMP, V_PHI);
    // the correction stored in tmp of the parents is overwritten on the way down of the next cycle
    enq_correct_gc(h, l - 1, l == h->L && !set_residual, h->o.n_cycle_up == 0);
    enq_vcycle(h, set_residual && l == h->L, l, l == h->L);
  }
}

// ---- set-up ---------------------------------------------------------------------------------------
int build_constant_stencils(afmg_handle* h) {
  // mg_box_lpl_stencil (m_af_multigrid.f90:1246-1264) per level; plus (after the L+1 rows) the volume
  // factors product(dr_base) * 0.5^(3(l-1)) used by af_tree_sum_cc (m_af_utils.f90:966-1027)
  std::vector<double> coef((size_t)(h->L + 1) * 8 + (h->L + 2), 0.0);
  for (int l = 1; l <= h->L; ++l) {
    double* c = &coef[(size_t)8 * l];
    for (int d = 0; d < 3; ++d) {
      const double dr = h->o.dr_base[d] * std::pow(0.5, l - 1);
      const double inv_dr2 = 1 / (dr * dr);
      c[1 + 2 * d] = inv_dr2;
      c[2 + 2 * d] = inv_dr2;
    }
    double s = 0.0;
    for (int m = 1; m < 7; ++m) s = s + c[m];
    c[0] = -s - h->o.helmholtz_lambda;
    c[7] = 1 / c[0];
    double fac = 1.0;
    for (int d = 0; d < 3; ++d) fac *= h->o.dr_base[d] * std::pow(0.5, l - 1);
    coef[(size_t)8 * (h->L + 1) + l] = fac;
  }
  int rc = dev_upload(h, &h->d_coef, coef);
  if (rc) return rc;
  // mg_box_prolong_linear_stencil / _sparse_stencil (m_af_multigrid.f90:1267-1304)
  std::vector<double> pc(8, 0.0);
  int pshape = 8;
  if (h->o.prolongation_type == AFMG_PROLONG_SPARSE) {
    pshape = 4;
    pc = {0.25, 0.25, 0.25, 0.25, 0, 0, 0, 0};
  } else {
    pc = {27 / 64.0, 9 / 64.0, 9 / 64.0, 3 / 64.0, 9 / 64.0, 3 / 64.0, 3 / 64.0, 1 / 64.0};
  }
  rc = dev_upload(h, &h->d_pcoef, pc);
  if (rc) return rc;
  h->cx.coef = h->d_coef;
  h->cx.pcoef = h->d_pcoef;
  h->cx.pshape = pshape;
  return AFMG_OK;
}

// unmap the peers' slabs (CUDA IPC)
void close_peers(afmg_handle* h) {
  for (int r = 0; r < AFMG_MAX_RANKS; ++r) {
    if (r != h->me && h->peer_slab[r] && !h->local_peers) cudaIpcCloseMemHandle(h->peer_slab[r]);
    h->peer_slab[r] = nullptr;
  }
  h->connected = (h->nranks == 1);
}

void mega_resolve_stamps(afmg_handle* h);

void fs_drop_graph(afmg_handle* h);  // afmg_field.inc
void drop_graphs(afmg_handle* h) {
  fs_drop_graph(h);
  for (auto& kv : h->graphs) {
    if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    free_programs(kv.second.progs);
  }
  h->graphs.clear();
}

// ---- slabs trimmed to the owned boxes ---------------------------------------------------------------------------
// Every rank addresses a box as (owner's slab base) + slot * BOX, so all ranks share one slot space; but a rank only
// ever touches its own records through its own base pointer (peers' records go through the peers' pointers).  With
// the virtual memory management API the slot space is reserved as addresses only and physical memory is mapped under
// the owned slot ranges (rounded to the 2 MB granularity): per-rank memory ~ 1 / N of the tree instead of all of it.
// Used by the single-process multi-GPU mode (afmg_opts.n_gpus), where peer access is a cuMemSetAccess away; the
// multi-process mode keeps cudaMalloc + CUDA IPC (sharing VMM allocations across processes needs file-descriptor
// passing).  AFMG_TRIM_SLABS=0 switches it off.
struct VmmApi {
  CUresult (*GetGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
  CUresult (*AddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
  CUresult (*AddressFree)(CUdeviceptr, size_t) = nullptr;
  CUresult (*Create)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
  CUresult (*Release)(CUmemGenericAllocationHandle) = nullptr;
  CUresult (*Map)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
  CUresult (*Unmap)(CUdeviceptr, size_t) = nullptr;
  CUresult (*SetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
  bool ok = false;
};
const VmmApi& vmm_api() {
  static VmmApi api = [] {
    VmmApi a;
    auto get = [](const char* name, void** fn) {
      cudaDriverEntryPointQueryResult q;
      return cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess && *fn;
    };
    a.ok = get("cuMemGetAllocationGranularity", (void**)&a.GetGranularity) && get("cuMemAddressReserve", (void**)&a.AddressReserve) &&
           get("cuMemAddressFree", (void**)&a.AddressFree) && get("cuMemCreate", (void**)&a.Create) &&
           get("cuMemRelease", (void**)&a.Release) && get("cuMemMap", (void**)&a.Map) && get("cuMemUnmap", (void**)&a.Unmap) &&
           get("cuMemSetAccess", (void**)&a.SetAccess);
    cudaGetLastError();
    return a;
  }();
  return api;
}

void slab_free(afmg_handle* h) {
  if (!h->d_slab) return;
  if (h->slab_vmm) {
    const VmmApi& a = vmm_api();
    for (size_t q = 0; q < h->slab_maps.size(); ++q) {
      a.Unmap((CUdeviceptr)(h->d_slab + h->slab_maps[q].first), h->slab_maps[q].second);
      a.Release((CUmemGenericAllocationHandle)h->slab_map_handles[q]);
    }
    a.AddressFree((CUdeviceptr)h->d_slab, h->slab_va_size);
    h->slab_maps.clear();
    h->slab_map_handles.clear();
    h->slab_vmm = false;
  } else {
    cudaFree(h->d_slab);
  }
  h->d_slab = nullptr;
  h->slab_mapped_bytes = 0;
}

// byte ranges of the slab that hold this rank's records (per variable and level) plus the box-sum table
std::vector<std::pair<size_t, size_t>> slab_owned_ranges(const afmg_handle* h) {
  std::vector<std::pair<size_t, size_t>> r;
  const size_t rec = (size_t)h->box_len * sizeof(double);
  for (int v = 0; v < h->slab_nvar; ++v)
    for (int l = 1; l <= h->L; ++l) {
      const int* c = &h->cut[(size_t)l * (h->nranks + 1)];
      if (c[h->me + 1] > c[h->me]) r.emplace_back(v * h->slab_var_stride + (size_t)c[h->me] * rec, (size_t)(c[h->me + 1] - c[h->me]) * rec);
    }
  r.emplace_back(h->slab_nvar * h->slab_var_stride, (size_t)h->nslots * sizeof(double));
  return r;
}

int slab_alloc(afmg_handle* h) {
  slab_free(h);
  bool trim = h->local_peers && h->nranks > 1 && vmm_api().ok;
  if (const char* env = getenv("AFMG_TRIM_SLABS")) trim = trim && atoi(env) != 0;
  if (!trim) {
    CK(cudaMalloc((void**)&h->d_slab, h->slab_bytes));
    CK(cudaMemset(h->d_slab, 0, h->slab_bytes));
    h->slab_mapped_bytes = h->slab_bytes;
    return AFMG_OK;
  }
  const VmmApi& a = vmm_api();
  CUmemAllocationProp prop = {};
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = h->device;
  size_t g = 0;
  if (a.GetGranularity(&g, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM) != CUDA_SUCCESS || g == 0)
    return h->fail(AFMG_ERR_CUDA, "cuMemGetAllocationGranularity failed");
  auto up = [g](size_t x) { return (x + g - 1) / g * g; };
  h->slab_va_size = up(h->slab_bytes);
  CUdeviceptr va = 0;
  if (a.AddressReserve(&va, h->slab_va_size, g, 0, 0) != CUDA_SUCCESS) return h->fail(AFMG_ERR_CUDA, "cuMemAddressReserve(%zu) failed", h->slab_va_size);
  h->d_slab = (char*)va;
  h->slab_vmm = true;
  // owned ranges, rounded outward to the granularity and merged
  auto own = slab_owned_ranges(h);
  std::vector<std::pair<size_t, size_t>> pieces;
  for (auto& r : own) pieces.emplace_back(r.first / g * g, up(r.first + r.second));  // [begin, end)
  std::sort(pieces.begin(), pieces.end());
  std::vector<std::pair<size_t, size_t>> merged;
  for (auto& p : pieces) {
    if (!merged.empty() && p.first <= merged.back().second) merged.back().second = std::max(merged.back().second, p.second);
    else merged.push_back(p);
  }
  std::vector<CUmemAccessDesc> acc(h->nranks);
  for (int r = 0; r < h->nranks; ++r) {
    acc[r].location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    acc[r].location.id = h->dev_base + r;
    acc[r].flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  }
  for (auto& m : merged) {
    const size_t off = m.first, size = m.second - m.first;
    CUmemGenericAllocationHandle mh;
    if (a.Create(&mh, size, &prop, 0) != CUDA_SUCCESS) return h->fail(AFMG_ERR_CUDA, "cuMemCreate(%zu bytes) failed (out of device memory?)", size);
    if (a.Map(va + off, size, 0, mh, 0) != CUDA_SUCCESS) {
      a.Release(mh);
      return h->fail(AFMG_ERR_CUDA, "cuMemMap failed");
    }
    h->slab_maps.emplace_back(off, size);
    h->slab_map_handles.push_back((unsigned long long)mh);
    if (a.SetAccess(va + off, size, acc.data(), acc.size()) != CUDA_SUCCESS) return h->fail(AFMG_ERR_CUDA, "cuMemSetAccess failed (peer access between the GPUs?)");
    CK(cudaMemset(h->d_slab + off, 0, size));
    h->slab_mapped_bytes += size;
  }
  return AFMG_OK;
}

// coarse_solver_initialize (m_coarse_solver.f90:71-194) + stencil_handle_boundaries (:442-491), with the
// Hypre solve replaced by the eigen-decomposition of the separable BC-folded operator
// reference layout v(ncf, i, j, k) -> device planes [m][colour][iidx]
template <int NC>
void planes_from_ref(const double* v, int ncf, std::vector<double>& out) {
  using L = Lay3<NC>;
  const size_t base = out.size();
  out.resize(base + (size_t)ncf * 2 * L::NI);
  for (int k = 1; k <= NC; ++k)
    for (int j = 1; j <= NC; ++j)
      for (int i = 1; i <= NC; ++i) {
        const int col = (i + j + k) & 1, idx = L::iidx((i - 1) >> 1, j, k);
        const double* src = v + (size_t)ncf * ((i - 1) + NC * ((j - 1) + NC * (k - 1)));
        for (int m = 0; m < ncf; ++m) out[base + (size_t)(m * 2 + col) * L::NI + idx] = src[m];
      }
}

int coarse_setup_dense(afmg_handle* h);

int coarse_setup(afmg_handle* h) {
  // the coarse buffers are reallocated below and the coarse path (separable / fused / dense) may change: cached
  // graphs hold the old pointers and launch topology by value
  drop_graphs(h);
  const int nc = h->o.n_cell, nc2 = h->nc2;
  const int nbox1 = nlev(h, 1);
  h->cs_dense = false;
  h->cs_planes = false;
  h->cs.lsf_fac = nullptr;
  h->cs.Ainv = nullptr;
  h->cs.Sinv = nullptr;
  if (!h->l1_st.empty()) return coarse_setup_dense(h);
  int nb[3], nx[3];
  for (int d = 0; d < 3; ++d) {
    nx[d] = h->o.coarse_grid_size[d];
    nb[d] = nx[d] / nc;
  }
  if (nb[0] * nb[1] * nb[2] != nbox1) return h->fail(AFMG_ERR_ARG, "coarse grid size does not match the level-1 boxes");
  // BC type per domain face must be uniform for the separable solve
  int face_type[6] = {0, 0, 0, 0, 0, 0};
  for (int r = 0; r < h->nbc; ++r) {
    if (h->h_lvl[h->bc_slot[r]] != 1) continue;
    if (!h->bc_set[r]) return h->fail(AFMG_ERR_STATE, "boundary condition not set for level-1 box %d face %d",
                                      h->slot2id[h->bc_slot[r]], h->bc_face[r] + 1);
    const int f = h->bc_face[r], ty = h->h_bc_type[r];
    if (ty != AFMG_BC_DIRICHLET && ty != AFMG_BC_NEUMANN)
      return h->fail(AFMG_ERR_UNSUPPORTED, "coarse grid: unsupported boundary condition %d (reference: error stop, "
                                           "m_coarse_solver.f90:486)", ty);
    if (face_type[f] != 0 && face_type[f] != ty) {
      // different level-1 boxes put different condition types on one domain face (the reference allows it:
      // stencil_handle_boundaries works box by box, m_coarse_solver.f90:442-491): the operator is no longer
      // separable, so the general dense coarse solve takes over
      return coarse_setup_dense(h);
    }
    face_type[f] = ty;
  }
  const double* c1 = nullptr;
  std::vector<double> coef1(8);
  {
    for (int d = 0; d < 3; ++d) {
      const double dr = h->o.dr_base[d];
      coef1[1 + 2 * d] = coef1[2 + 2 * d] = 1 / (dr * dr);
    }
    c1 = coef1.data();
  }
  // bc_to_rhs per (box, face, cell)
  std::vector<double> b2r((size_t)nbox1 * 6 * nc2, 0.0);
  for (int r = 0; r < h->nbc; ++r) {
    const int s = h->bc_slot[r], f = h->bc_face[r];
    if (h->h_lvl[s] != 1) continue;
    const double cnb = c1[f + 1];
    double v;
    if (h->h_bc_type[r] == AFMG_BC_DIRICHLET) v = -2 * cnb;
    else v = -(cnb * h->o.dr_base[f >> 1]) * ((f & 1) ? 1 : -1);
    for (int q = 0; q < nc2; ++q) b2r[((size_t)s * 6 + f) * nc2 + q] = v;
  }
  // 1D operators and their eigen-decompositions
  std::vector<std::vector<double>> Qd(3), lam(3);
  for (int d = 0; d < 3; ++d) {
    const int n = nx[d];
    const long double c = c1[1 + 2 * d];
    std::vector<long double> A((size_t)n * n, 0.0L), V;
    for (int i = 0; i < n; ++i) {
      A[(size_t)i * n + i] = -2 * c;
      if (i > 0) A[(size_t)i * n + i - 1] = c;
      if (i < n - 1) A[(size_t)i * n + i + 1] = c;
    }
    const int flo = 2 * d, fhi = 2 * d + 1;
    if (h->o.periodic[d]) {
      if (n > 2) {
        A[(size_t)0 * n + n - 1] += c;
        A[(size_t)(n - 1) * n + 0] += c;
      }
    } else {
      if (face_type[flo] == 0 || face_type[fhi] == 0)
        return h->fail(AFMG_ERR_STATE, "boundary conditions missing on domain faces of dimension %d", d + 1);
      A[0] += (face_type[flo] == AFMG_BC_DIRICHLET) ? -c : c;
      A[(size_t)(n - 1) * n + n - 1] += (face_type[fhi] == AFMG_BC_DIRICHLET) ? -c : c;
    }
    jacobi_eig(n, A, V);
    Qd[d].resize((size_t)n * n);
    lam[d].resize(n);
    for (int i = 0; i < n; ++i) lam[d][i] = (double)A[(size_t)i * n + i];
    for (size_t i = 0; i < (size_t)n * n; ++i) Qd[d][i] = (double)V[i];
  }
  const int ntot = nx[0] * nx[1] * nx[2];
  std::vector<double> inveig(ntot);
  double mu_max = 0, mu_min = 1e300;
  for (int k = 0; k < nx[2]; ++k)
    for (int j = 0; j < nx[1]; ++j)
      for (int i = 0; i < nx[0]; ++i) {
        const double mu = lam[0][i] + lam[1][j] + lam[2][k] - h->o.helmholtz_lambda;
        mu_max = std::max(mu_max, std::fabs(mu));
        mu_min = std::min(mu_min, std::fabs(mu));
        inveig[i + nx[0] * (j + nx[1] * k)] = 1 / mu;
      }
  if (mu_min < 1e-10 * mu_max)
    return h->fail(AFMG_ERR_SINGULAR, "coarse-grid operator is singular (all-Neumann/periodic without Helmholtz term)");
  std::vector<int> bix((size_t)nbox1 * 3);
  for (int s = 0; s < nbox1; ++s)
    for (int d = 0; d < 3; ++d) bix[(size_t)s * 3 + d] = h->h_ix[(size_t)s * 3 + d] - 1;
  int rc;
  if ((rc = dev_upload(h, &h->d_b2r, b2r))) return rc;
  for (int d = 0; d < 3; ++d)
    if ((rc = dev_upload(h, &h->d_Q[d], Qd[d]))) return rc;
  if ((rc = dev_upload(h, &h->d_inveig, inveig))) return rc;
  if ((rc = dev_upload(h, &h->d_cs_bix, bix))) return rc;
  std::vector<double> zeros(ntot, 0.0);
  if ((rc = dev_upload(h, &h->d_v0, zeros))) return rc;
  if ((rc = dev_upload(h, &h->d_v1, zeros))) return rc;
  for (int d = 0; d < 3; ++d) {
    h->cs.nx[d] = nx[d];
    h->cs.nb[d] = nb[d];
    h->cs.Q[d] = h->d_Q[d];
  }
  h->cs.b2r = h->d_b2r;
  h->cs.inv_eig = h->d_inveig;
  h->cs.v0 = h->d_v0;
  h->cs.v1 = h->d_v1;
  h->cs.bix = h->d_cs_bix;
  h->cs_ready = true;
  return AFMG_OK;
}

// Block-tridiagonal set-up of the general coarse solve on the device (see k_cs_plane_assemble): np dense inversions
// of m x m Schur-complement planes, in-place Gauss-Jordan without pivoting (2 small launches per pivot).
int coarse_setup_planes(afmg_handle* h, const int* nx, const int* nbx, int sdim, const std::vector<double>& cellst,
                        const std::vector<double>& b2r, const std::vector<double>& lsf_fac, bool any_f) {
  const int n = nx[0] * nx[1] * nx[2], np = nx[sdim], m = n / np, nbox1 = nlev(h, 1);
  const int gs[3] = {1, nx[0], nx[0] * nx[1]};
  std::vector<double> lo(n, 0.0), up(n, 0.0);
  for (int g = 0; g < n; ++g) {
    const int q = (g / gs[sdim]) % nx[sdim];
    if (q > 0) lo[g] = cellst[(size_t)7 * g + 1 + 2 * sdim];
    if (q < np - 1) up[g] = cellst[(size_t)7 * g + 2 + 2 * sdim];
  }
  std::vector<int> bix((size_t)nbox1 * 3), per = {h->o.periodic[0], h->o.periodic[1], h->o.periodic[2]};
  for (int sb = 0; sb < nbox1; ++sb)
    for (int d = 0; d < 3; ++d) bix[(size_t)sb * 3 + d] = h->h_ix[(size_t)sb * 3 + d] - 1;
  int rc;
  double* d_cellst = nullptr;
  int* d_per = nullptr;
  if ((rc = dev_upload(h, &d_cellst, cellst))) return rc;
  if ((rc = dev_upload(h, &d_per, per))) return rc;
  if ((rc = dev_upload(h, &h->d_b2r, b2r))) return rc;
  if ((rc = dev_upload(h, &h->d_lsf_fac, lsf_fac))) return rc;
  if ((rc = dev_upload(h, &h->d_cs_bix, bix))) return rc;
  if ((rc = dev_upload(h, &h->d_cs_lo, lo))) return rc;
  if ((rc = dev_upload(h, &h->d_cs_up, up))) return rc;
  std::vector<double> zeros(n, 0.0);
  if ((rc = dev_upload(h, &h->d_v0, zeros))) return rc;
  if ((rc = dev_upload(h, &h->d_v1, zeros))) return rc;
  if ((rc = dev_upload(h, &h->d_cs_w, zeros))) return rc;
  if ((rc = dev_upload(h, &h->d_cs_xs, zeros))) return rc;
  zeros.resize(2 * (size_t)m);
  if ((rc = dev_upload(h, &h->d_cs_t, zeros))) return rc;  // tvec [m] + pivot row / column scratch shares nothing
  if (h->d_Sinv) cudaFree(h->d_Sinv);
  h->d_Sinv = nullptr;
  CK(cudaMalloc((void**)&h->d_Sinv, (size_t)np * m * m * sizeof(double)));
  double *d_row = nullptr, *d_col = nullptr;
  CK(cudaMalloc((void**)&d_row, (size_t)m * sizeof(double)));
  CK(cudaMalloc((void**)&d_col, (size_t)m * sizeof(double)));
  for (int d = 0; d < 3; ++d) {
    h->cs.nx[d] = nx[d];
    h->cs.nb[d] = nbx[d];
    h->cs.Q[d] = nullptr;
  }
  h->cs.b2r = h->d_b2r;
  h->cs.inv_eig = nullptr;
  h->cs.v0 = h->d_v0;
  h->cs.v1 = h->d_v1;
  h->cs.bix = h->d_cs_bix;
  h->cs.Ainv = nullptr;
  h->cs.lsf_fac = any_f ? h->d_lsf_fac : nullptr;
  h->cs.sdim = sdim;
  h->cs.np = np;
  h->cs.m = m;
  h->cs.Sinv = h->d_Sinv;
  h->cs.lo = h->d_cs_lo;
  h->cs.up = h->d_cs_up;
  h->cs.w = h->d_cs_w;
  h->cs.xs = h->d_cs_xs;
  h->cs.tvec = h->d_cs_t;
  const size_t mm = (size_t)m * m;
  const int eb = (int)((mm + 255) / 256);
  for (int p = 0; p < np; ++p) {
    double* S = h->d_Sinv + (size_t)p * mm;
    k_cs_plane_assemble<<<eb, 256, 0, h->stream>>>(h->cs, d_cellst, d_per, p, S, p > 0 ? S - mm : nullptr);
    for (int q = 0; q < m; ++q) {
      k_gj_save<<<(m + 255) / 256, 256, 0, h->stream>>>(S, m, q, d_row, d_col);
      k_gj_update<<<eb, 256, 0, h->stream>>>(S, m, q, d_row, d_col);
    }
  }
  // singular operators (all-Neumann without a Helmholtz term) show up as a vanishing last pivot: non-finite entries
  cudaError_t e = cudaStreamSynchronize(h->stream);
  double probe = 0.0;
  if (e == cudaSuccess) e = cudaMemcpy(&probe, h->d_Sinv + (size_t)(np - 1) * mm + mm - 1, sizeof probe, cudaMemcpyDeviceToHost);
  cudaFree(d_cellst);
  cudaFree(d_per);
  cudaFree(d_row);
  cudaFree(d_col);
  if (e != cudaSuccess) return h->fail(AFMG_ERR_CUDA, "coarse plane set-up failed: %s", cudaGetErrorString(e));
  if (!std::isfinite(probe) || std::fabs(probe) > 1e10 * std::fabs(1.0 / cellst[0]))
    return h->fail(AFMG_ERR_SINGULAR, "coarse-grid operator is singular (all-Neumann without Helmholtz term)");
  h->cs_dense = true;
  h->cs_planes = true;
  h->cs_ready = true;
  return AFMG_OK;
}

// General coarse solve: level-1 boxes carry explicit stencils (variable eps / level set), so the operator is
// not separable.  Same matrix as coarse_solver_initialize + stencil_handle_boundaries
// (m_coarse_solver.f90:71-194, :442-491); its inverse is formed on the host (banded LU, one solve per unit
// vector, threaded) and applied on the device as a dense mat-vec (k_cs_dense).
int coarse_setup_dense(afmg_handle* h) {
  const int nc = h->o.n_cell, nc2 = h->nc2, ncell = nc * nc * nc;
  const int nbox1 = nlev(h, 1);
  int nx[3], nbx[3];
  for (int d = 0; d < 3; ++d) {
    nx[d] = h->o.coarse_grid_size[d];
    nbx[d] = nx[d] / nc;
  }
  if (nbx[0] * nbx[1] * nbx[2] != nbox1) return h->fail(AFMG_ERR_ARG, "coarse grid size does not match the level-1 boxes");
  const int n = nx[0] * nx[1] * nx[2];
  // periodic dimensions couple the first and the last cell (HYPRE_StructGridSetPeriodic, m_coarse_solver.f90:
  // 97-104): the band becomes full, so the factorisation is a dense one and is kept to small grids
  const bool any_periodic = h->o.periodic[0] || h->o.periodic[1] || h->o.periodic[2];
  // Large grids (the reference's own 3D electrode example has a 32^3 coarse grid, afivo/examples/
  // electrode_example.f90:41-45): block-tridiagonal direct solve over planes stacked along the longest non-periodic
  // dimension, dense plane inverses on the device (k_cs_plane_*); memory n^2 / np doubles
  const bool use_planes = n > 8192 || (any_periodic && n > 2048);
  int sdim = -1;
  for (int d = 0; d < 3; ++d)
    if (!h->o.periodic[d] && (sdim < 0 || nx[d] >= nx[sdim])) sdim = d;
  if (use_planes) {
    if (sdim < 0)
      return h->fail(AFMG_ERR_UNSUPPORTED, "non-separable coarse grid of %d cells, periodic in every dimension", n);
    const double gb = (double)n * (double)(n / nx[sdim]) * 8.0 / 1e9;
    if (gb > 24.0)
      return h->fail(AFMG_ERR_UNSUPPORTED, "explicit stencils on a coarse grid of %d cells: the plane inverses need %.0f GB "
                                           "(limit 24); use more, smaller level-1 boxes with a coarser level below", n, gb);
  }
  const int bw = use_planes ? 0 : (any_periodic ? n - 1 : nx[0] * nx[1]), ldab = 2 * bw + 1;
  std::vector<double> ab(use_planes ? 1 : (size_t)ldab * n, 0.0);  // ab[(bw + r - c) + ldab * c] = A(r, c)
  std::vector<double> cellst(use_planes ? (size_t)7 * n : 0, 0.0);  // BC-folded stencil of every cell, global order
  std::vector<double> b2r((size_t)nbox1 * 6 * nc2, 0.0), lsf_fac((size_t)nbox1 * ncell, 0.0);
  bool any_f = false;
  const int gstride[3] = {1, nx[0], nx[0] * nx[1]};
  std::vector<double> lvl_c(7);
  for (int d = 0; d < 3; ++d) lvl_c[1 + 2 * d] = lvl_c[2 + 2 * d] = 1 / (h->o.dr_base[d] * h->o.dr_base[d]);
  {
    double sum = 0.0;
    for (int m = 1; m < 7; ++m) sum = sum + lvl_c[m];
    lvl_c[0] = -sum - h->o.helmholtz_lambda;
  }
  for (int sb = 0; sb < nbox1; ++sb) {
    std::vector<double> full((size_t)7 * ncell);
    auto it = h->l1_st.find(sb);
    if (it != h->l1_st.end()) {
      full = it->second.first;
      if (!it->second.second.empty()) {
        any_f = true;
        std::copy(it->second.second.begin(), it->second.second.end(), lsf_fac.begin() + (size_t)sb * ncell);
      }
    } else {
      for (int q = 0; q < ncell; ++q)
        for (int m = 0; m < 7; ++m) full[(size_t)7 * q + m] = lvl_c[m];
    }
    auto lin = [&](int i, int j, int k) { return (i - 1) + nc * ((j - 1) + nc * (k - 1)); };
    for (int f = 0; f < 6; ++f) {  // stencil_handle_boundaries
      if (h->h_nbr[(size_t)sb * 6 + f] >= 0) continue;
      const int r = h->h_aux[(size_t)sb * 6 + f];
      if (!h->bc_set[r]) return h->fail(AFMG_ERR_STATE, "boundary condition not set for level-1 box %d face %d",
                                        h->slot2id[sb], f + 1);
      const int ty = h->h_bc_type[r], d = f >> 1, layer = (f & 1) ? nc : 1;
      const int ta = (d == 0) ? 1 : 0, tb = (d == 2) ? 1 : 2;
      for (int b = 1; b <= nc; ++b)
        for (int a = 1; a <= nc; ++a) {
          int q[3];
          q[d] = layer;
          q[ta] = a;
          q[tb] = b;
          double* st = &full[(size_t)7 * lin(q[0], q[1], q[2])];
          double& out = b2r[((size_t)sb * 6 + f) * nc2 + (a - 1) + (b - 1) * nc];
          if (ty == AFMG_BC_DIRICHLET) {
            st[0] = st[0] - st[f + 1];
            out = -2 * st[f + 1];
          } else if (ty == AFMG_BC_NEUMANN) {
            st[0] = st[0] + st[f + 1];
            out = -(st[f + 1] * h->o.dr_base[d]) * ((f & 1) ? 1 : -1);
          } else {
            return h->fail(AFMG_ERR_UNSUPPORTED, "coarse grid: unsupported boundary condition %d (reference: error "
                                                 "stop, m_coarse_solver.f90:486)", ty);
          }
          st[f + 1] = 0.0;
        }
    }
    const int* bix = &h->h_ix[(size_t)sb * 3];
    for (int k = 1; k <= nc; ++k)
      for (int j = 1; j <= nc; ++j)
        for (int i = 1; i <= nc; ++i) {
          const int gi[3] = {(bix[0] - 1) * nc + i - 1, (bix[1] - 1) * nc + j - 1, (bix[2] - 1) * nc + k - 1};
          const int r = gi[0] + nx[0] * (gi[1] + nx[1] * gi[2]);
          const double* st = &full[(size_t)7 * lin(i, j, k)];
          if (use_planes) {
            for (int m = 0; m < 7; ++m) cellst[(size_t)7 * r + m] = st[m];
            for (int m = 0; m < 6; ++m) {
              const int d = m >> 1, qd = gi[d] + ((m & 1) ? 1 : -1);
              if (st[m + 1] != 0.0 && (qd < 0 || qd >= nx[d]) && !h->o.periodic[d])
                return h->fail(AFMG_ERR_ARG, "coarse matrix: coupling outside the grid");
            }
            continue;
          }
          ab[(size_t)bw + (size_t)ldab * r] += st[0];
          for (int m = 0; m < 6; ++m) {
            if (st[m + 1] == 0.0) continue;
            const int d = m >> 1, sgn = (m & 1) ? 1 : -1;
            const int qd = gi[d] + sgn;
            int c = r + sgn * gstride[d];
            if (qd < 0 || qd >= nx[d]) {
              if (!h->o.periodic[d]) return h->fail(AFMG_ERR_ARG, "coarse matrix: coupling outside the grid");
              c = r - sgn * (nx[d] - 1) * gstride[d];  // wrap around
            }
            ab[(size_t)(bw + r - c) + (size_t)ldab * c] += st[m + 1];
          }
        }
  }
  if (use_planes) {
    int rc = coarse_setup_planes(h, nx, nbx, sdim, cellst, b2r, lsf_fac, any_f);
    return rc;
  }
  // banded LU without pivoting (diagonally dominant M-matrix up to sign)
  for (int c = 0; c < n; ++c) {
    const double piv = ab[(size_t)bw + (size_t)ldab * c];
    if (piv == 0.0 || !std::isfinite(piv)) return h->fail(AFMG_ERR_SINGULAR, "coarse-grid operator: zero pivot");
    const int rmax = std::min(n - 1, c + bw);
    for (int r = c + 1; r <= rmax; ++r) ab[(size_t)(bw + r - c) + (size_t)ldab * c] /= piv;
    for (int c2 = c + 1; c2 <= rmax; ++c2) {
      const double u = ab[(size_t)(bw + c - c2) + (size_t)ldab * c2];
      if (u == 0.0) continue;
      for (int r = c + 1; r <= rmax; ++r)
        ab[(size_t)(bw + r - c2) + (size_t)ldab * c2] -= ab[(size_t)(bw + r - c) + (size_t)ldab * c] * u;
    }
  }
  if (std::fabs(ab[(size_t)bw + (size_t)ldab * (n - 1)]) < 1e-10 * std::fabs(ab[(size_t)bw]))
    return h->fail(AFMG_ERR_SINGULAR, "coarse-grid operator is singular (all-Neumann without Helmholtz term)");
  // inverse, column by column (A^-1 e_c), stored row-major
  std::vector<double> Ainv((size_t)n * n);
  {
    const int nthr = std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
    std::vector<std::thread> pool;
    for (int tid = 0; tid < nthr; ++tid)
      pool.emplace_back([&, tid] {
        std::vector<double> x(n);
        for (int col = tid; col < n; col += nthr) {
          std::fill(x.begin(), x.end(), 0.0);
          x[col] = 1.0;
          for (int c = col; c < n; ++c) {  // L y = e (unit lower)
            const double xc = x[c];
            if (xc == 0.0) continue;
            const int rmax = std::min(n - 1, c + bw);
            for (int r = c + 1; r <= rmax; ++r) x[r] -= ab[(size_t)(bw + r - c) + (size_t)ldab * c] * xc;
          }
          for (int c = n - 1; c >= 0; --c) {  // U x = y
            x[c] /= ab[(size_t)bw + (size_t)ldab * c];
            const double xc = x[c];
            const int rmin = std::max(0, c - bw);
            for (int r = rmin; r < c; ++r) x[r] -= ab[(size_t)(bw + r - c) + (size_t)ldab * c] * xc;
          }
          for (int r = 0; r < n; ++r) Ainv[(size_t)r * n + col] = x[r];
        }
      });
    for (auto& th : pool) th.join();
  }
  std::vector<int> bix((size_t)nbox1 * 3);
  for (int sb = 0; sb < nbox1; ++sb)
    for (int d = 0; d < 3; ++d) bix[(size_t)sb * 3 + d] = h->h_ix[(size_t)sb * 3 + d] - 1;
  int rc;
  if ((rc = dev_upload(h, &h->d_b2r, b2r))) return rc;
  if ((rc = dev_upload(h, &h->d_Ainv, Ainv))) return rc;
  if ((rc = dev_upload(h, &h->d_lsf_fac, lsf_fac))) return rc;
  if ((rc = dev_upload(h, &h->d_cs_bix, bix))) return rc;
  std::vector<double> zeros(n, 0.0);
  if ((rc = dev_upload(h, &h->d_v0, zeros))) return rc;
  if ((rc = dev_upload(h, &h->d_v1, zeros))) return rc;
  for (int d = 0; d < 3; ++d) {
    h->cs.nx[d] = nx[d];
    h->cs.nb[d] = nbx[d];
    h->cs.Q[d] = nullptr;
  }
  h->cs.b2r = h->d_b2r;
  h->cs.inv_eig = nullptr;
  h->cs.v0 = h->d_v0;
  h->cs.v1 = h->d_v1;
  h->cs.bix = h->d_cs_bix;
  h->cs.Ainv = h->d_Ainv;
  h->cs.lsf_fac = any_f ? h->d_lsf_fac : nullptr;
  h->cs_dense = true;
  h->cs_ready = true;
  return AFMG_OK;
}

int s2_ensure_ready(afmg_handle* h);

int ensure_ready(afmg_handle* h) {
  if (!h->have_tree) return h->fail(AFMG_ERR_STATE, "afmg_set_tree has not been called");
  if (h->o.ndim == 2) return s2_ensure_ready(h);
  for (int r = 0; r < h->nbc; ++r)
    if (!h->bc_set[r])
      return h->fail(AFMG_ERR_STATE, "boundary condition not set for box %d face %d (call afmg_set_bc)",
                     h->slot2id[h->bc_slot[r]], h->bc_face[r] + 1);
  if (!h->connected)
    return h->fail(AFMG_ERR_STATE, "multi-GPU handle: call afmg_comm_export / afmg_comm_connect after afmg_set_tree");
  if (!h->cs_ready) {
    int rc = coarse_setup(h);
    if (rc) return rc;
  }
  return AFMG_OK;
}

int ensure_stage(afmg_handle* h, size_t bytes, size_t nslots) {
  if (bytes > h->stage_bytes) {
    if (h->d_stage) cudaFree(h->d_stage);
    h->d_stage = nullptr;
    CK(cudaMalloc((void**)&h->d_stage, bytes));
    h->stage_bytes = bytes;
  }
  if (nslots > h->stage_slots_n) {
    if (h->d_stage_slots) cudaFree(h->d_stage_slots);
    h->d_stage_slots = nullptr;
    CK(cudaMalloc((void**)&h->d_stage_slots, nslots * sizeof(int)));
    h->stage_slots_n = nslots;
  }
  return AFMG_OK;
}

// per-phase durations of the profiled k_mega launches (time stamps written by CTA 0 after every grid barrier)
const char* mega_kind_name(int k) {
  static const char* names[MK_COUNT] = {"none", "rb_prepare", "gsrb", "edges_corners", "restrict", "residual", "gc",
                                        "gc_parent", "coarse", "coarse", "coarse", "coarse", "correct", "store_corr",
                                        "copy", "init_phi_rhs", "clear"};
  return (k >= 0 && k < MK_COUNT) ? names[k] : "?";
}
void mega_resolve_stamps(afmg_handle* h) {
  if (h->stamp_runs.empty()) {
    h->stamps_used = 0;
    return;
  }
  std::vector<unsigned long long> st(h->stamps_used);
  cudaMemcpy(st.data(), h->d_stamps, st.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
  for (auto& run : h->stamp_runs) {
    const MegaProgram& pr = *run.second;
    for (int p = 0; p < pr.nphase; ++p) {
      const MegaOp& op = pr.h_ops[pr.h_phases[p].op0];
      const bool with_lvl = op.lvl > 0 && op.kind != MK_RESID && (op.kind < MK_CS_FUSED || op.kind >= MK_CORRECT);
      std::string name = std::string("mega:") + mega_kind_name(op.kind) + (with_lvl ? "_L" + std::to_string(op.lvl) : "");
      auto& e = h->prof[name];
      e.ms += (double)(st[run.first + p + 1] - st[run.first + p]) * 1e-6;
      e.calls++;
    }
  }
  h->stamp_runs.clear();
  h->stamps_used = 0;
}

// walk `body` once in plan mode: the persistent-kernel programs of its segments are built and uploaded into `out`
template <class F>
void mega_plan(afmg_handle* h, std::vector<MegaProgram>& out, F& body) {
  h->rec_ops.clear();
  h->rec_phases.clear();
  h->mega_open = false;
  h->progs = &out;
  h->mega_mode = 1;
  body();
  mega_flush(h);
  h->mega_mode = 0;
}
template <class F>
void mega_exec(afmg_handle* h, std::vector<MegaProgram>& progs, F& body) {
  h->progs = &progs;
  h->prog_idx = 0;
  h->mega_open = false;
  h->mega_mode = 2;
  body();
  mega_flush(h);
  h->mega_mode = 0;
  h->progs = nullptr;
}

// run `body` (a sequence of enq_* calls) without a graph: single operations of the C ABI, profiling mode
template <class F>
int run_direct(afmg_handle* h, F body) {
  if (!mega_possible(h)) {
    body();
    pre_sync(h);  // every body ends with all ranks in step (see enq_barrier)
    return AFMG_OK;
  }
  CK(cudaStreamSynchronize(h->stream));  // the programs of the previous direct run may still be in use
  mega_resolve_stamps(h);
  free_programs(h->direct_progs);
  mega_plan(h, h->direct_progs, body);
  mega_exec(h, h->direct_progs, body);
  return AFMG_OK;
}

// run `body` through a cached CUDA graph (or directly when profiling)
template <class F>
int run_graph(afmg_handle* h, std::tuple<int, int, int> key, int n_rep, F body) {
  if (h->profiling) {
    for (int i = 0; i < n_rep; ++i) {
      int rc = run_direct(h, body);
      if (rc) return rc;
    }
    CK(cudaGetLastError());
    return AFMG_OK;
  }
  auto it = h->graphs.find(key);
  if (it == h->graphs.end()) {
    cudaGraph_t g = nullptr;
    Graph gr;
    const bool mega = mega_possible(h);
    if (mega) mega_plan(h, gr.progs, body);
    const int64_t before = h->launches;
    h->capturing = true;
    CK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
    if (mega) mega_exec(h, gr.progs, body);
    else body();
    pre_sync(h);  // every cycle ends with all ranks in step (see enq_barrier)
    cudaError_t e = cudaStreamEndCapture(h->stream, &g);
    h->capturing = false;
    if (e != cudaSuccess) {
      free_programs(gr.progs);
      return h->fail(AFMG_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(e));
    }
    gr.launches = h->launches - before;
    h->launches = before;
    e = cudaGraphInstantiate(&gr.exec, g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) {
      free_programs(gr.progs);
      return h->fail(AFMG_ERR_CUDA, "graph instantiate failed: %s", cudaGetErrorString(e));
    }
    it = h->graphs.emplace(key, std::move(gr)).first;
  }
  for (int i = 0; i < n_rep; ++i) {
    CK(cudaGraphLaunch(it->second.exec, h->stream));
    h->launches += it->second.launches;
  }
  return AFMG_OK;
}

int check_lvl(afmg_handle* h, int lvl) {
  if (lvl < 1 || lvl > h->L) return h->fail(AFMG_ERR_ARG, "level %d out of range 1..%d", lvl, h->L);
  return AFMG_OK;
}

int finish_op(afmg_handle* h) {
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(h->stream));
  if (h->profiling) {
    prof_resolve(h);
    mega_resolve_stamps(h);
  }
  if (h->mega_launched) {
    h->mega_launched = false;
    unsigned long long err = 0;
    CK(cudaMemcpy(&err, &h->d_msync->err, sizeof err, cudaMemcpyDeviceToHost));
    if (err) return h->fail(AFMG_ERR_CUDA, "a grid barrier of the persistent kernel timed out (CTAs not co-resident?)");
  }
  if (h->nranks > 1) {
    unsigned long long err = 0;
    CK(cudaMemcpy(&err, &h->d_comm->err, sizeof err, cudaMemcpyDeviceToHost));
    if (err) return h->fail(AFMG_ERR_COMM, "a cross-GPU barrier timed out (a peer rank is missing or out of step)");
  }
  return AFMG_OK;
}

// device -> host copy of the records of a chunk, leaving the records of boxes this rank does not own (slot < 0)
// untouched in the caller's buffer: with several rank handles in one process they all write into the same array
cudaError_t copy_own_records(afmg_handle* h, double* hp, const double* dp, const int* slots, int m, size_t rec_len,
                                    cudaStream_t st, bool to_device = false) {
  if (to_device) {  // host -> device: only the records of boxes this rank owns travel (the others are skipped by the
                    // unpack kernel anyway; with N rank handles in one process each moves 1/N of the buffer)
    if (h->nranks == 1) return cudaMemcpyAsync(const_cast<double*>(dp), hp, (size_t)m * rec_len * sizeof(double), cudaMemcpyHostToDevice, st);
    int q = 0;
    while (q < m) {
      while (q < m && slots[q] < 0) ++q;
      int e = q;
      while (e < m && slots[e] >= 0) ++e;
      if (e > q) {
        cudaError_t rc = cudaMemcpyAsync(const_cast<double*>(dp) + (size_t)q * rec_len, hp + (size_t)q * rec_len,
                                         (size_t)(e - q) * rec_len * sizeof(double), cudaMemcpyHostToDevice, st);
        if (rc != cudaSuccess) return rc;
      }
      q = e;
    }
    return cudaSuccess;
  }
  if (h->nranks == 1) return cudaMemcpyAsync(hp, dp, (size_t)m * rec_len * sizeof(double), cudaMemcpyDeviceToHost, st);
  int q = 0;
  while (q < m) {
    while (q < m && slots[q] < 0) ++q;
    int e = q;
    while (e < m && slots[e] >= 0) ++e;
    if (e > q) {
      cudaError_t rc = cudaMemcpyAsync(hp + (size_t)q * rec_len, dp + (size_t)q * rec_len, (size_t)(e - q) * rec_len * sizeof(double),
                                       cudaMemcpyDeviceToHost, st);
      if (rc != cudaSuccess) return rc;
    }
    q = e;
  }
  return cudaSuccess;
}

#include "afmg2d.inc"
#include "afmg_field.inc"

// Shared by afmg_set_stencils (blob in host memory) and afmg_build_stencils_device (blob produced on the device by
// the builder kernels, builders_dev.cuh): bookkeeping of the per-box stencil kinds and of the coefficient pool, then
// the conversion of every variable stencil from the reference's v(n, i, j, k) to the device planes -- on the host
// for a host blob, by k_planes_from_ref for a device blob, so that in the second case no coefficient crosses PCIe.
struct StencilJob {
  int64_t src;    // offset in the blob
  long long dst;  // offset in the pool
  int ncf;        // coefficients per cell (planes); 0: plain copy of `len` doubles
  int len;
};

template <int NC>
__global__ void k_planes_from_ref(const double* __restrict__ blob, const StencilJob* __restrict__ jobs, int njobs, double* pool) {
  using L = Lay3<NC>;
  const StencilJob jb = jobs[blockIdx.x];
  const double* src = blob + jb.src;
  double* dst = pool + jb.dst;
  if (jb.ncf == 0) {
    for (int q = threadIdx.x; q < jb.len; q += blockDim.x) dst[q] = src[q];
    return;
  }
  for (int c = threadIdx.x; c < NC * NC * NC; c += blockDim.x) {
    const int i = c % NC + 1, j = (c / NC) % NC + 1, k = c / (NC * NC) + 1;
    const int col = (i + j + k) & 1, idx = L::iidx((i - 1) >> 1, j, k);
    for (int m = 0; m < jb.ncf; ++m) dst[(size_t)(m * 2 + col) * L::NI + idx] = src[(size_t)jb.ncf * c + m];
  }
}

int ingest_stencils(afmg_handle* h, int32_t n, const afmg_stencil_desc* desc, const double* blob, int64_t blob_len,
                           bool blob_on_device) {
  const int total = h->nslots, nc = h->o.n_cell, ncell = nc * nc * nc;
  const int NI2 = ncell;  // 2 colours x NI doubles per plane
  h->h_opk.assign(total, 0);
  h->h_pk.assign(total, 0);
  h->h_opoff.assign(total, 0);
  h->h_foff.assign(total, -1);
  h->h_poff.assign(total, 0);
  h->h_tag.assign(total, 0);
  h->l1_st.clear();
  std::vector<StencilJob> jobs;
  long long pool_len = 0;
  std::vector<std::pair<int, const afmg_stencil_desc*>> l1;  // level-1 boxes: the coarse solver needs them on the host
  auto need = [&](int64_t off, int64_t len) { return off >= 0 && off + len <= blob_len; };
  for (int q = 0; q < n; ++q) {
    const afmg_stencil_desc& d = desc[q];
    if (d.box_id < 1 || d.box_id > h->highest_id || h->id2slot[d.box_id] < 0)
      return h->fail(AFMG_ERR_ARG, "afmg_set_stencils: unknown box %d", d.box_id);
    const int s = h->id2slot[d.box_id];
    h->h_tag[s] = d.tag;
    if (d.op_stype == 1 || d.op_stype == 2) {
      const int64_t len = d.op_stype == 1 ? 7 : (int64_t)7 * ncell;
      if (!need(d.op_offset, len)) return h->fail(AFMG_ERR_ARG, "afmg_set_stencils: operator of box %d outside the blob", d.box_id);
      h->h_opk[s] = (unsigned char)d.op_stype;
      h->h_opoff[s] = pool_len;
      if (d.op_stype == 1) {
        jobs.push_back({d.op_offset, pool_len, 0, 7});
        pool_len += 8;
      } else {
        jobs.push_back({d.op_offset, pool_len, 7, 0});
        pool_len += (long long)7 * NI2;
      }
      if (h->h_lvl[s] == 1) l1.emplace_back(s, &d);
    } else if (d.op_stype != 0) {
      return h->fail(AFMG_ERR_UNSUPPORTED, "afmg_set_stencils: stencil type %d (sparse stencils are not implemented in the "
                                           "reference either, m_af_stencil.f90:853)", d.op_stype);
    }
    if (d.f_offset >= 0) {
      if (!need(d.f_offset, ncell)) return h->fail(AFMG_ERR_ARG, "afmg_set_stencils: f of box %d outside the blob", d.box_id);
      if (h->h_opk[s] == 0) return h->fail(AFMG_ERR_ARG, "afmg_set_stencils: box %d has f but an implicit operator", d.box_id);
      h->h_foff[s] = pool_len;
      jobs.push_back({d.f_offset, pool_len, 1, 0});
      pool_len += NI2;
    }
    if (d.prolong_shape != 0) {
      if (h->h_lvl[s] < 2) return h->fail(AFMG_ERR_ARG, "afmg_set_stencils: level-1 box %d has no prolongation", d.box_id);
      int kind;
      if (d.prolong_shape == AFMG_STENCIL_P248 && d.prolong_stype == 1) kind = 1;
      else if (d.prolong_shape == AFMG_STENCIL_P234 && d.prolong_stype == 1) kind = 2;
      else if (d.prolong_shape == AFMG_STENCIL_P234 && d.prolong_stype == 2) kind = 3;
      else return h->fail(AFMG_ERR_UNSUPPORTED, "afmg_set_stencils: prolongation shape %d / type %d of box %d",
                          d.prolong_shape, d.prolong_stype, d.box_id);
      const int64_t len = kind == 1 ? 8 : (kind == 2 ? 4 : (int64_t)4 * ncell);
      if (!need(d.prolong_offset, len))
        return h->fail(AFMG_ERR_ARG, "afmg_set_stencils: prolongation of box %d outside the blob", d.box_id);
      h->h_pk[s] = (unsigned char)kind;
      h->h_poff[s] = pool_len;
      if (kind == 3) {
        jobs.push_back({d.prolong_offset, pool_len, 4, 0});
        pool_len += (long long)4 * NI2;
      } else {
        jobs.push_back({d.prolong_offset, pool_len, 0, (int)len});
        pool_len += 8;
      }
    }
  }
  // ---- the coefficient pool
  if (h->d_stv) cudaFree(h->d_stv);
  h->d_stv = nullptr;
  CK(cudaMalloc((void**)&h->d_stv, (size_t)std::max<long long>(pool_len, 1) * sizeof(double)));
  CK(cudaMemset(h->d_stv, 0, (size_t)std::max<long long>(pool_len, 1) * sizeof(double)));
  if (blob_on_device) {
    if (!jobs.empty()) {
      StencilJob* d_jobs = nullptr;
      CK(cudaMalloc((void**)&d_jobs, jobs.size() * sizeof(StencilJob)));
      CK(cudaMemcpy(d_jobs, jobs.data(), jobs.size() * sizeof(StencilJob), cudaMemcpyHostToDevice));
      DISPATCH_NC(h, NC, { k_planes_from_ref<NC><<<(int)jobs.size(), 256, 0, h->stream>>>(blob, d_jobs, (int)jobs.size(), h->d_stv); });
      CK(cudaStreamSynchronize(h->stream));
      cudaFree(d_jobs);
    }
  } else {
    std::vector<double> pool((size_t)std::max<long long>(pool_len, 1), 0.0);
    for (const StencilJob& jb : jobs) {
      if (jb.ncf == 0) {
        std::copy(blob + jb.src, blob + jb.src + jb.len, pool.begin() + jb.dst);
      } else {
        std::vector<double> tmp;
        DISPATCH_NC(h, NC, planes_from_ref<NC>(blob + jb.src, jb.ncf, tmp));
        std::copy(tmp.begin(), tmp.end(), pool.begin() + jb.dst);
      }
    }
    CK(cudaMemcpy(h->d_stv, pool.data(), pool.size() * sizeof(double), cudaMemcpyHostToDevice));
  }
  // ---- level-1 stencils for the coarse-grid matrix (m_coarse_solver.f90:71-194), reference layout, on the host
  for (auto& pr : l1) {
    const int s = pr.first;
    const afmg_stencil_desc& d = *pr.second;
    const int64_t len = d.op_stype == 1 ? 7 : (int64_t)7 * ncell;
    std::vector<double> op(len), fv;
    if (blob_on_device) CK(cudaMemcpy(op.data(), blob + d.op_offset, len * sizeof(double), cudaMemcpyDeviceToHost));
    else std::copy(blob + d.op_offset, blob + d.op_offset + len, op.begin());
    auto& e = h->l1_st[s];
    e.first.resize((size_t)7 * ncell);
    for (int c = 0; c < ncell; ++c)
      for (int m = 0; m < 7; ++m) e.first[(size_t)7 * c + m] = d.op_stype == 1 ? op[m] : op[(size_t)7 * c + m];
    if (d.f_offset >= 0) {
      e.second.resize(ncell);
      if (blob_on_device) CK(cudaMemcpy(e.second.data(), blob + d.f_offset, (size_t)ncell * sizeof(double), cudaMemcpyDeviceToHost));
      else e.second.assign(blob + d.f_offset, blob + d.f_offset + ncell);
    }
  }
  // this rank's boxes with an explicit operator, by level
  std::vector<int> spec;
  h->spec_off.assign(h->L + 2, 0);
  bool any = false;
  for (int l = 1; l <= h->L; ++l) {
    h->spec_off[l] = (int)spec.size();
    const Range r = own(h, l);
    for (int s = r.s0; s < r.s0 + r.n; ++s)
      if (h->h_opk[s]) spec.push_back(s);
  }
  h->spec_off[h->L + 1] = (int)spec.size();
  for (int s = 0; s < total; ++s) any = any || h->h_opk[s] || h->h_pk[s] || h->h_tag[s];
  // refinement-boundary faces of variable-eps boxes use mg_sides_rb_extrap (mg_auto_rb)
  const int nrules = h->nbc + h->nrb;
  h->h_rule_flag.assign(std::max(nrules, 1), 0);
  for (int r = 0; r < h->nrb; ++r)
    if ((h->h_tag[h->h_rb_slot[r]] & h->o.operator_mask) == AFMG_TAG_VEPS_BOX) h->h_rule_flag[h->nbc + r] = 1;
  h->have_stencils = any;
  int rc;
  if ((rc = dev_upload(h, &h->d_opk, h->h_opk))) return rc;
  if ((rc = dev_upload(h, &h->d_pk, h->h_pk))) return rc;
  if ((rc = dev_upload(h, &h->d_opoff, h->h_opoff))) return rc;
  if ((rc = dev_upload(h, &h->d_foff, h->h_foff))) return rc;
  if ((rc = dev_upload(h, &h->d_poff, h->h_poff))) return rc;
  if ((rc = dev_upload(h, &h->d_spec, spec))) return rc;
  if ((rc = dev_upload(h, &h->d_rule_flag, h->h_rule_flag))) return rc;
  h->cx.opk = any ? h->d_opk : nullptr;
  h->cx.pk = any ? h->d_pk : nullptr;
  h->cx.opoff = h->d_opoff;
  h->cx.foff = h->d_foff;
  h->cx.poff = h->d_poff;
  h->cx.stv = h->d_stv;
  h->cx.rule_flag = any ? h->d_rule_flag : nullptr;
  h->cx.lsf_value_p = &h->d_comm->lsf_value;
  return AFMG_OK;
}


}  // namespace

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" {

const char* afmg_last_error(const afmg_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

// afmg_opts.n_gpus > 1: a front handle over n rank handles (one per device, one host thread each)
static int create_multi(afmg_handle** out, const afmg_opts* opts, int ng, int ndev) {
  const int base = opts->device >= 0 ? opts->device : 0;
  if (opts->ndim != 3) {
    g_create_error = "n_gpus > 1 needs a 3D tree (the 2D path runs on one GPU)";
    return AFMG_ERR_UNSUPPORTED;
  }
  if (ng > AFMG_MAX_RANKS || base + ng > ndev) {
    g_create_error = "n_gpus = " + std::to_string(ng) + " from device " + std::to_string(base) + ": only " +
                     std::to_string(ndev) + " devices visible (limit " + std::to_string(AFMG_MAX_RANKS) + ")";
    return AFMG_ERR_ARG;
  }
  afmg_handle* h = new afmg_handle();
  h->o = *opts;
  h->is_multi = true;
  h->nc2 = opts->n_cell * opts->n_cell;
  h->box_len = (opts->n_cell + 2) * (opts->n_cell + 2) * (opts->n_cell + 2);
  for (int r = 0; r < ng; ++r) {
    afmg_opts o = *opts;
    o.n_gpus = 1;
    o.device = base + r;
    afmg_handle* sub = nullptr;
    int rc = afmg_create(&sub, &o);
    if (!rc) rc = afmg_comm_init(sub, ng, r);
    if (rc) {
      if (sub) afmg_destroy(sub);
      for (auto* q : h->subs) afmg_destroy(q);
      delete h;
      return rc;
    }
    sub->local_peers = true;
    sub->dev_base = base;
    h->subs.push_back(sub);
  }
  for (int r = 0; r < ng; ++r) {
    RankWorker* w = new RankWorker();
    w->th = std::thread([w] { w->loop(); });
    h->workers.push_back(w);
  }
  *out = h;
  return AFMG_OK;
}

// peer access and peer pointers between the rank handles of one process (in place of afmg_comm_export / _connect)
static int connect_local(afmg_handle* sub, afmg_handle* front);

int afmg_create(afmg_handle** out, const afmg_opts* opts) {
  if (!out || !opts) {
    g_create_error = "null argument";
    return AFMG_ERR_ARG;
  }
  *out = nullptr;
  if (opts->ndim != 2 && opts->ndim != 3) {
    g_create_error = "ndim must be 2 or 3 (1D trees are not part of the accelerated path)";
    return AFMG_ERR_UNSUPPORTED;
  }
  if (opts->ndim == 3 && opts->n_cell != 4 && opts->n_cell != 8 && opts->n_cell != 16) {
    g_create_error = "n_cell must be 4, 8 or 16 in 3D";
    return AFMG_ERR_UNSUPPORTED;
  }
  if (opts->ndim == 2 && opts->n_cell != 4 && opts->n_cell != 8 && opts->n_cell != 16 && opts->n_cell != 32) {
    g_create_error = "n_cell must be 4, 8, 16 or 32 in 2D";
    return AFMG_ERR_UNSUPPORTED;
  }
  if (opts->ndim == 3 && opts->coord_t != AFMG_XYZ) {
    g_create_error = "3D trees must be Cartesian";
    return AFMG_ERR_ARG;
  }
  if (opts->coord_t != AFMG_XYZ && opts->coord_t != AFMG_CYL) {
    g_create_error = "unknown coordinate system";
    return AFMG_ERR_ARG;
  }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    g_create_error = std::string("no CUDA device: ") + cudaGetErrorString(e) + " (there is no CPU fallback)";
    return AFMG_ERR_CUDA;
  }
  {
    int ng = opts->n_gpus;
    if (ng == 0 && opts->ndim == 3)  // the environment only steers 3D handles (the 2D path runs on one GPU)
      if (const char* env = getenv("AFMG_N_GPUS")) ng = atoi(env);
    if (ng > 1) return create_multi(out, opts, ng, ndev);
  }
  afmg_handle* h = new afmg_handle();
  h->o = *opts;
  if (opts->device >= 0) {
    h->device = opts->device;
    e = cudaSetDevice(h->device);
  } else {
    e = cudaGetDevice(&h->device);
  }
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) {
    // highest priority: a kernel of the side stream (a cross-GPU barrier next to a long half-sweep) must get its CTA
    // before the thousands of pending CTAs of the kernel it runs beside
    int least = 0, greatest = 0;
    cudaDeviceGetStreamPriorityRange(&least, &greatest);
    h->side_priority = greatest;
    e = cudaStreamCreateWithPriority(&h->side_stream, cudaStreamNonBlocking, greatest);
  }
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming);
  for (int b = 0; b < 2 && e == cudaSuccess; ++b) {
    e = cudaEventCreateWithFlags(&h->ev_copied[b], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_consumed[b], cudaEventDisableTiming);
  }
  if (e == cudaSuccess) e = cudaEventCreate(&h->ev0);
  if (e == cudaSuccess) e = cudaEventCreate(&h->ev1);
  if (e == cudaSuccess) e = cudaMalloc((void**)&h->d_comm, sizeof(CommBlock));
  if (e == cudaSuccess) e = cudaMemset(h->d_comm, 0, sizeof(CommBlock));
  if (e == cudaSuccess) e = cudaMalloc((void**)&h->d_msync, sizeof(MegaSync));
  if (e == cudaSuccess) e = cudaMemset(h->d_msync, 0, sizeof(MegaSync));
  h->stamps_cap = 1 << 16;
  if (e == cudaSuccess) e = cudaMalloc((void**)&h->d_stamps, (size_t)h->stamps_cap * sizeof(unsigned long long));
  if (e == cudaSuccess) h->d_scal = h->d_comm->scal;  // address arithmetic only
  if (e == cudaSuccess) e = cudaMemcpy(&h->d_comm->lsf_value, &opts->lsf_boundary_value, sizeof(double), cudaMemcpyHostToDevice);
  h->peers.p[0] = h->d_comm;
  if (const char* env = getenv("AFMG_PDL")) h->pdl = atoi(env) != 0;
  if (const char* env = getenv("AFMG_CS_FUSED")) h->cs_fused = atoi(env) != 0;
  if (const char* env = getenv("AFMG_SMALL_CTAS")) h->small_ctas = atoi(env);
  if (const char* env = getenv("AFMG_GSRB_WIDE")) h->gsrb_wide = atoi(env) != 0;
  cudaDeviceGetAttribute(&h->n_sm, cudaDevAttrMultiProcessorCount, h->device);
  if (const char* env = getenv("AFMG_GSRB_FUSED_GEN")) h->gsrb_fused_gen = atoi(env) != 0;
  if (const char* env = getenv("AFMG_MIN_SPLIT_BOXES")) h->min_split_boxes = atoi(env);
  if (const char* env = getenv("AFMG_BARRIER_LEAN")) h->barrier_lean = atoi(env);
  if (const char* env = getenv("AFMG_MEGA")) h->mega_enabled = atoi(env) != 0;
  if (const char* env = getenv("AFMG_MEGA_MAX_BOXES")) h->mega_max_boxes = atoi(env);
  if (const char* env = getenv("AFMG_MEGA_CLUSTER")) h->mega_cluster = std::max(0, std::min(16, atoi(env)));
  if (const char* env = getenv("AFMG_MEGA_TIMEOUT_S")) {
    const double sec = atof(env);
    if (sec > 0) h->mega_timeout_ns = (unsigned long long)(sec * 1e9);
  }
  if (const char* env = getenv("AFMG_BARRIER_TIMEOUT_S")) {
    const double sec = atof(env);
    if (sec > 0) h->barrier_timeout_ns = (unsigned long long)(sec * 1e9);
  }
  if (e != cudaSuccess) {
    g_create_error = std::string("CUDA initialisation failed: ") + cudaGetErrorString(e);
    delete h;
    return AFMG_ERR_CUDA;
  }
  if (opts->ndim == 3) configure_kernels(h);
  h->nc2 = opts->n_cell * opts->n_cell;
  h->box_len = (opts->n_cell + 2) * (opts->n_cell + 2) * (opts->ndim == 3 ? opts->n_cell + 2 : 1);
  *out = h;
  return AFMG_OK;
}

int afmg_destroy(afmg_handle* h) {
  if (!h) return AFMG_OK;
  if (h->is_multi) {
    multi_call(h, [](afmg_handle* sub) -> int { return afmg_destroy(sub); });
    for (auto* w : h->workers) {
      {
        std::lock_guard<std::mutex> lk(w->mu);
        w->quit = true;
        w->cv.notify_all();
      }
      w->th.join();
      delete w;
    }
    delete h;
    return AFMG_OK;
  }
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  drop_graphs(h);
  free_programs(h->direct_progs);
  cudaFree(h->d_msync);
  cudaFree(h->d_stamps);
  s2_free(h);
  fs_free(h);
  close_peers(h);
  slab_free(h);
  cudaFree(h->d_owner);
  cudaFree(h->d_opk);
  cudaFree(h->d_pk);
  cudaFree(h->d_rule_flag);
  cudaFree(h->d_opoff);
  cudaFree(h->d_foff);
  cudaFree(h->d_poff);
  cudaFree(h->d_stv);
  cudaFree(h->d_spec);
  cudaFree(h->d_Ainv);
  cudaFree(h->d_rmin);
  cudaFree(h->d_bblob);
  cudaFree(h->d_Sinv);
  cudaFree(h->d_cs_lo);
  cudaFree(h->d_cs_up);
  cudaFree(h->d_cs_w);
  cudaFree(h->d_cs_xs);
  cudaFree(h->d_cs_t);
  cudaFree(h->d_lsf_fac);
  cudaFree(h->d_bv);
  cudaFree(h->d_bvoff);
  int* ip[] = {h->d_nbr, h->d_aux, h->d_nmat, h->d_parent, h->d_child0, h->d_coff, h->d_lvl, h->d_rb_slot,
               h->d_rb_face, h->d_stage_slots, h->d_cs_bix};
  for (auto p : ip) cudaFree(p);
  double* dp[] = {h->d_coef, h->d_rule_c, h->d_rule_B, h->d_pcoef, h->d_stage, h->d_b2r, h->d_Q[0],
                  h->d_Q[1], h->d_Q[2], h->d_inveig, h->d_v0, h->d_v1};
  for (auto p : dp) cudaFree(p);
  cudaFree(h->d_comm);
  cudaEventDestroy(h->ev0);
  cudaEventDestroy(h->ev1);
  for (int b = 0; b < 2; ++b) {
    if (h->ev_copied[b]) cudaEventDestroy(h->ev_copied[b]);
    if (h->ev_consumed[b]) cudaEventDestroy(h->ev_consumed[b]);
  }
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  if (h->side_stream) cudaStreamDestroy(h->side_stream);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  cudaStreamDestroy(h->stream);
  delete h;
  return AFMG_OK;
}

int afmg_set_tree(afmg_handle* h, const afmg_tree* t) {
  if (h && h->is_multi) {
    if (!t) return AFMG_ERR_ARG;
    int rc = multi_call(h, [=](afmg_handle* sub_) -> int { return afmg_set_tree(sub_, t); });
    if (rc) return rc;
    h->have_tree = true;
    h->L = h->subs[0]->L;
    h->nslots = h->subs[0]->nslots;
    return multi_call(h, [=](afmg_handle* sub_) -> int { return connect_local(sub_, h); });  // all slabs exist now
  }
  if (!h || !t) return AFMG_ERR_ARG;
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  drop_graphs(h);
  fs_free(h);  // field data (fc, norm, eps, level-set distances) belongs to the previous tree
  h->have_tree = false;
  h->cs_ready = false;
  h->resid_fresh = false;
  const int L = t->highest_lvl, N = t->highest_id;
  if (L < 1 || N < 1) return h->fail(AFMG_ERR_ARG, "empty tree");
  if (h->o.ndim == 2) return s2_set_tree(h, t);
  int total = 0;
  for (int l = 0; l < L; ++l) total += t->lvl_counts[l];
  h->L = L;
  h->highest_id = N;
  h->nslots = total;
  h->lvl_off.assign(L + 2, 0);
  for (int l = 1; l <= L; ++l) h->lvl_off[l + 1] = h->lvl_off[l] + t->lvl_counts[l - 1];
  h->lvl_off[1] = 0;
  h->id2slot.assign(N + 1, -1);
  h->slot2id.assign(total, 0);
  // ownership: contiguous Morton ranges per level, cut at sibling groups (afmg_partition); depends on the box
  // counts only
  std::vector<int> rel((size_t)L * (h->nranks + 1));
  {
    int min_split = h->min_split_boxes;
    if (min_split < 0) min_split = (4 << 20) / (h->o.n_cell * h->o.n_cell * h->o.n_cell);  // 4 Mi cells
    afmg_partition_min(h->nranks, L, t->lvl_counts, min_split, rel.data());
  }
  // slots: level 1 in list order, finer levels in Morton order of ix-1
  {
    int p = 0;
    for (int l = 1; l <= L; ++l) {
      const int n = t->lvl_counts[l - 1];
      std::vector<std::pair<uint64_t, int>> keyed(n);
      for (int q = 0; q < n; ++q) {
        const int id = t->lvl_ids[p + q];
        if (id < 1 || id > N) return h->fail(AFMG_ERR_ARG, "box id %d out of range", id);
        if (t->lvl[id] != l) return h->fail(AFMG_ERR_ARG, "box %d listed on level %d but has lvl %d", id, l, t->lvl[id]);
        const int* ix = t->ix + (size_t)id * 3;
        keyed[q] = {l == 1 ? (uint64_t)q : morton3(ix[0] - 1, ix[1] - 1, ix[2] - 1), id};
      }
      std::sort(keyed.begin(), keyed.end());
      for (int q = 0; q < n; ++q) {
        const int id = keyed[q].second;
        if (h->id2slot[id] != -1) return h->fail(AFMG_ERR_ARG, "box %d listed twice", id);
        h->id2slot[id] = h->lvl_off[l] + q;
        h->slot2id[h->lvl_off[l] + q] = id;
      }
      p += n;
    }
  }
  auto slot_of = [&](int id) { return id > 0 ? h->id2slot[id] : (id == 0 ? -2 : -1); };
  h->h_nbr.assign((size_t)total * 6, 0);
  h->h_aux.assign((size_t)total * 6, -1);
  h->h_nmat.assign((size_t)total * 27, 0);
  h->h_parent.assign(total, -1);
  h->h_child0.assign(total, -1);
  h->h_coff.assign(total, 0);
  h->h_lvl.assign(total, 0);
  h->h_ix.assign((size_t)total * 3, 0);
  h->h_rmin.assign((size_t)total * 3, 0.0);
  h->npar.assign(L + 2, 0);
  h->bc_slot.clear();
  h->bc_face.clear();
  h->h_rb_slot.clear();
  h->h_rb_face.clear();
  h->rb_lvl_off.assign(L + 2, 0);
  for (int s = 0; s < total; ++s) {
    const int id = h->slot2id[s];
    const int l = t->lvl[id];
    h->h_lvl[s] = l;
    for (int d = 0; d < 3; ++d) h->h_ix[(size_t)s * 3 + d] = t->ix[(size_t)id * 3 + d];
    for (int d = 0; d < 3; ++d)  // box%r_min; without it: r_base + (ix - 1) * n_cell * dr (exact for power-of-two sizes)
      h->h_rmin[(size_t)s * 3 + d] = t->r_min ? t->r_min[(size_t)id * 3 + d]
                                              : h->o.r_base[d] + (t->ix[(size_t)id * 3 + d] - 1) * (h->o.n_cell * h->o.dr_base[d] * std::pow(0.5, l - 1));
    if (l > 1) {
      const int p = t->parent[id];
      if (p < 1 || p > N || h->id2slot[p] < 0) return h->fail(AFMG_ERR_ARG, "box %d has invalid parent %d", id, p);
      h->h_parent[s] = h->id2slot[p];
    }
    const int* ix = t->ix + (size_t)id * 3;
    h->h_coff[s] = ((ix[0] - 1) & 1) | (((ix[1] - 1) & 1) << 1) | (((ix[2] - 1) & 1) << 2);
    const int* ch = t->children + (size_t)id * 8;
    if (ch[0] != 0) {
      for (int c = 0; c < 8; ++c)
        if (ch[c] < 1 || ch[c] > N || h->id2slot[ch[c]] < 0)
          return h->fail(AFMG_ERR_ARG, "box %d has invalid child %d", id, ch[c]);
      h->h_child0[s] = h->id2slot[ch[0]];
      for (int c = 0; c < 8; ++c)
        if (h->id2slot[ch[c]] != h->h_child0[s] + c)
          return h->fail(AFMG_ERR_ARG, "children of box %d are not a complete sibling group", id);
      h->npar[l]++;
    }
    for (int m = 0; m < 27; ++m) {
      const int v = t->neighbor_mat[(size_t)id * 27 + m];
      if (v > N) return h->fail(AFMG_ERR_ARG, "neighbor_mat of box %d out of range", id);
      h->h_nmat[(size_t)s * 27 + m] = slot_of(v);
    }
  }
  // faces: neighbour slot, or rule row (physical faces first, then refinement boundaries by level)
  for (int s = 0; s < total; ++s) {
    const int id = h->slot2id[s];
    for (int f = 0; f < 6; ++f) {
      const int v = t->neighbors[(size_t)id * 6 + f];
      if (v > 0) {
        if (v > N || h->id2slot[v] < 0) return h->fail(AFMG_ERR_ARG, "box %d has invalid neighbour %d", id, v);
        h->h_nbr[(size_t)s * 6 + f] = h->id2slot[v];
      } else if (v < 0) {
        h->h_nbr[(size_t)s * 6 + f] = -1;
        h->h_aux[(size_t)s * 6 + f] = (int)h->bc_slot.size();
        h->bc_slot.push_back(s);
        h->bc_face.push_back(f);
      } else {
        if (h->h_lvl[s] == 1) return h->fail(AFMG_ERR_ARG, "level-1 box %d has a refinement boundary", id);
        h->h_nbr[(size_t)s * 6 + f] = -1;
      }
    }
  }
  h->nbc = (int)h->bc_slot.size();
  for (int s = 0; s < total; ++s)
    for (int f = 0; f < 6; ++f)
      if (h->h_nbr[(size_t)s * 6 + f] == -1 && h->h_aux[(size_t)s * 6 + f] == -1) {
        const int p = h->h_parent[s];
        if (h->h_nbr[(size_t)p * 6 + f] < 0)
          return h->fail(AFMG_ERR_ARG, "tree is not 2:1 balanced at box %d face %d", h->slot2id[s], f + 1);
        h->h_aux[(size_t)s * 6 + f] = h->nbc + (int)h->h_rb_slot.size();
        h->h_rb_slot.push_back(s);
        h->h_rb_face.push_back(f);
        h->rb_lvl_off[h->h_lvl[s] + 1]++;
      }
  for (int l = 1; l <= L; ++l) h->rb_lvl_off[l + 1] += h->rb_lvl_off[l];
  h->nrb = (int)h->h_rb_slot.size();
  h->bc_set.assign(h->nbc, 0);
  h->h_bc_type.assign(h->nbc, 0);

  int rc;
  if ((rc = dev_upload(h, &h->d_nbr, h->h_nbr))) return rc;
  if ((rc = dev_upload(h, &h->d_aux, h->h_aux))) return rc;
  if ((rc = dev_upload(h, &h->d_nmat, h->h_nmat))) return rc;
  if ((rc = dev_upload(h, &h->d_parent, h->h_parent))) return rc;
  if ((rc = dev_upload(h, &h->d_child0, h->h_child0))) return rc;
  if ((rc = dev_upload(h, &h->d_coff, h->h_coff))) return rc;
  if ((rc = dev_upload(h, &h->d_lvl, h->h_lvl))) return rc;
  if ((rc = dev_upload(h, &h->d_rb_slot, h->h_rb_slot))) return rc;
  if ((rc = dev_upload(h, &h->d_rb_face, h->h_rb_face))) return rc;
  if ((rc = dev_upload(h, &h->d_rmin, h->h_rmin))) return rc;
  const int nrules = h->nbc + h->nrb;
  std::vector<double> rule_c((size_t)std::max(nrules, 1) * 3, 0.0);
  for (int r = h->nbc; r < nrules; ++r) {  // mg_sides_rb: 0.5*gc + 0.75*x1 - 0.25*x2
    rule_c[(size_t)r * 3 + 0] = 0.5;
    rule_c[(size_t)r * 3 + 1] = 0.75;
    rule_c[(size_t)r * 3 + 2] = -0.25;
  }
  if ((rc = dev_upload(h, &h->d_rule_c, rule_c))) return rc;
  std::vector<double> rule_B((size_t)std::max(nrules, 1) * h->nc2, 0.0);
  if ((rc = dev_upload(h, &h->d_rule_B, rule_B))) return rc;
  h->h_bc_B.assign((size_t)h->nbc * h->nc2, 0.0);  // what the device rows hold now
  h->h_bc_c.assign((size_t)h->nbc * 3, 0.0);
  // one slab per rank: phi | rhs | tmp | box sums (a single CUDA IPC handle covers all of it)
  close_peers(h);
  slab_free(h);
  h->slab_var_stride = (((size_t)total * h->box_len * sizeof(double)) + 255) / 256 * 256;
  // multi-GPU: the field norm lives in the slab too, so that the peers can read its halo (af_gc_tree of the norm);
  // on one GPU it is allocated on first use (afmg_field.inc)
  h->slab_nvar = (h->nranks > 1) ? 4 : 3;
  h->slab_bytes = h->slab_nvar * h->slab_var_stride + (size_t)total * sizeof(double);

  // ownership: contiguous Morton ranges per level, cut at sibling groups (afmg_partition)
  {
    h->cut.assign((size_t)(L + 2) * (h->nranks + 1), 0);
    h->rb_cut.assign((size_t)(L + 2) * (h->nranks + 1), 0);
    h->h_owner.assign(total, 0);
    for (int l = 1; l <= L; ++l)
      for (int r = 0; r <= h->nranks; ++r) h->cut[(size_t)l * (h->nranks + 1) + r] = h->lvl_off[l] + rel[(size_t)(l - 1) * (h->nranks + 1) + r];
    for (int l = 1; l <= L; ++l)
      for (int r = 0; r < h->nranks; ++r)
        for (int q = h->cut[(size_t)l * (h->nranks + 1) + r]; q < h->cut[(size_t)l * (h->nranks + 1) + r + 1]; ++q)
          h->h_owner[q] = (unsigned char)r;
    // refinement-boundary faces are listed by level, then by slot: the faces of a rank are contiguous
    for (int l = 1; l <= L; ++l) {
      int* c = &h->rb_cut[(size_t)l * (h->nranks + 1)];
      int q = h->rb_lvl_off[l];
      for (int r = 0; r < h->nranks; ++r) {
        c[r] = q;
        while (q < h->rb_lvl_off[l + 1] && h->h_owner[h->h_rb_slot[q]] == r) ++q;
      }
      c[h->nranks] = q;
      if (q != h->rb_lvl_off[l + 1]) return h->fail(AFMG_ERR_ARG, "internal: refinement-boundary faces not sorted by owner");
    }
    if ((rc = dev_upload(h, &h->d_owner, h->h_owner))) return rc;
    // the slab: all of it on one GPU / in the multi-process mode, the owned slot ranges in the single-process mode
    if ((rc = slab_alloc(h))) return rc;
    for (int v = 0; v < 3; ++v) h->d_cc[v] = (double*)(h->d_slab + v * h->slab_var_stride);
    h->d_cc[3] = nullptr;
    h->d_cc[4] = (h->slab_nvar == 4) ? (double*)(h->d_slab + 3 * h->slab_var_stride) : nullptr;
    h->d_boxsum = (double*)(h->d_slab + h->slab_nvar * h->slab_var_stride);
    h->peer_slab[h->me] = h->d_slab;
    h->lvl_multi.assign(L + 2, 0);
    for (int l = 1; l <= L; ++l) {
      int owners = 0;
      for (int r = 0; r < h->nranks; ++r)
        owners += h->cut[(size_t)l * (h->nranks + 1) + r + 1] > h->cut[(size_t)l * (h->nranks + 1) + r] ? 1 : 0;
      h->lvl_multi[l] = owners > 1;
    }
    h->unsynced = false;
  }
  h->connected = (h->nranks == 1);
  // explicit stencils belong to the previous tree: the shim re-ships them (afmg_set_stencils)
  h->have_stencils = false;
  h->l1_st.clear();
  h->spec_off.assign(L + 2, 0);
  h->cx.opk = nullptr;
  h->cx.pk = nullptr;
  h->cx.opoff = h->cx.foff = h->cx.poff = nullptr;
  h->cx.stv = nullptr;
  h->cx.rule_flag = nullptr;
  h->cx.lsf_value_p = &h->d_comm->lsf_value;
  h->cx.bvoff = nullptr;  // the boundary-value list belongs to the previous tree
  h->cx.bv = nullptr;
  h->bv_ids.clear();

  for (int v = 0; v < 5; ++v) h->cx.cc[v] = h->d_cc[v];
  h->cx.nranks = h->nranks;
  h->cx.me = h->me;
  h->cx.owner = h->d_owner;
  for (int r = 0; r < AFMG_MAX_RANKS; ++r) {
    for (int v = 0; v < 5; ++v) h->cx.ccr[r][v] = nullptr;
    h->cx.bsum[r] = nullptr;
  }
  for (int v = 0; v < 5; ++v) h->cx.ccr[h->me][v] = h->d_cc[v];
  h->cx.bsum[h->me] = h->d_boxsum;
  h->cx.nbr = h->d_nbr;
  h->cx.aux = h->d_aux;
  h->cx.nmat = h->d_nmat;
  h->cx.parent = h->d_parent;
  h->cx.child0 = h->d_child0;
  h->cx.coff = h->d_coff;
  h->cx.lvl = h->d_lvl;
  h->cx.rule_c = h->d_rule_c;
  h->cx.rule_B = h->d_rule_B;
  h->cx.rb_slot = h->d_rb_slot;
  h->cx.rb_face = h->d_rb_face;
  h->cx.rb_row0 = h->nbc;
  if ((rc = build_constant_stencils(h))) return rc;
  h->have_tree = true;
  return AFMG_OK;
}

int afmg_set_bc(afmg_handle* h, int32_t n_faces, const int32_t* box_id, const int32_t* nb, const int32_t* bc_type,
                const double* bc_val) {
  AFMG_MULTI(h, afmg_set_bc(sub_, n_faces, box_id, nb, bc_type, bc_val));
  if (!h) return AFMG_ERR_ARG;
  if (!h->have_tree) return h->fail(AFMG_ERR_STATE, "afmg_set_tree has not been called");
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  if (h->o.ndim == 2) return s2_set_bc(h, n_faces, box_id, nb, bc_type, bc_val);
  const int nc2 = h->nc2;
  bool types_changed = false;
  // The rows go into the host mirrors of the physical-boundary part of rule_B / rule_c, and the touched range of
  // rows goes up in ONE copy each: the call precedes every solve (the applied voltage changes with time,
  // src/m_field.f90:590-610), and a copy per face -- the faces arrive in the caller's order, not in row order -- cost
  // 3 ms on a streamer-like tree, five V-cycles' worth.
  for (int q = 0; q < n_faces; ++q) {  // all arguments are checked before anything changes
    const int id = box_id[q], f = nb[q] - 1;
    if (id < 1 || id > h->highest_id || h->id2slot[id] < 0) return h->fail(AFMG_ERR_ARG, "afmg_set_bc: unknown box %d", id);
    if (f < 0 || f > 5) return h->fail(AFMG_ERR_ARG, "afmg_set_bc: invalid neighbour direction %d", nb[q]);
    const int s = h->id2slot[id];
    if (h->h_nbr[(size_t)s * 6 + f] != -1 || h->h_aux[(size_t)s * 6 + f] >= h->nbc)
      return h->fail(AFMG_ERR_ARG, "afmg_set_bc: box %d face %d is not a physical boundary", id, nb[q]);
    if (bc_type[q] != AFMG_BC_DIRICHLET && bc_type[q] != AFMG_BC_NEUMANN && bc_type[q] != AFMG_BC_CONTINUOUS &&
        bc_type[q] != AFMG_BC_DIRICHLET_COPY)
      return h->fail(AFMG_ERR_ARG, "afmg_set_bc: unknown boundary condition %d", bc_type[q]);
  }
  int rmin = h->nbc, rmax = -1;
  for (int q = 0; q < n_faces; ++q) {
    const int f = nb[q] - 1, s = h->id2slot[box_id[q]];
    const int r = h->h_aux[(size_t)s * 6 + f];
    double c0 = 0, c1 = 0, c2 = 0;
    switch (bc_type[q]) {  // bc_to_gc (m_af_ghostcell.f90:192-214)
      case AFMG_BC_DIRICHLET: c0 = 2; c1 = -1; c2 = 0; break;
      case AFMG_BC_NEUMANN:
        c0 = h->o.dr_base[f >> 1] * std::pow(0.5, h->h_lvl[s] - 1) * ((f & 1) ? 1 : -1);
        c1 = 1;
        c2 = 0;
        break;
      case AFMG_BC_CONTINUOUS: c0 = 0; c1 = 2; c2 = -1; break;
      default: c0 = 1; c1 = 0; c2 = 0; break;  // AFMG_BC_DIRICHLET_COPY
    }
    if (!h->bc_set[r] || h->h_bc_type[r] != bc_type[q]) types_changed = true;
    h->bc_set[r] = 1;
    h->h_bc_type[r] = bc_type[q];
    h->h_bc_c[(size_t)r * 3 + 0] = c0;
    h->h_bc_c[(size_t)r * 3 + 1] = c1;
    h->h_bc_c[(size_t)r * 3 + 2] = c2;
    std::memcpy(&h->h_bc_B[(size_t)r * nc2], bc_val + (size_t)q * nc2, nc2 * sizeof(double));
    rmin = std::min(rmin, r);
    rmax = std::max(rmax, r);
  }
  if (rmax >= rmin) {
    const size_t nr = (size_t)(rmax - rmin + 1);
    CK(cudaMemcpy(h->d_rule_B + (size_t)rmin * nc2, &h->h_bc_B[(size_t)rmin * nc2], nr * nc2 * sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(h->d_rule_c + (size_t)rmin * 3, &h->h_bc_c[(size_t)rmin * 3], nr * 3 * sizeof(double), cudaMemcpyHostToDevice));
  }
  if (types_changed) h->cs_ready = false;
  return AFMG_OK;
}

int afmg_set_helmholtz_lambda(afmg_handle* h, double lambda) {
  AFMG_MULTI(h, afmg_set_helmholtz_lambda(sub_, lambda));
  if (!h) return AFMG_ERR_ARG;
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  h->o.helmholtz_lambda = lambda;
  h->cs_ready = false;
  drop_graphs(h);
  if (h->have_tree) return h->o.ndim == 2 ? s2_constant_stencils(h) : build_constant_stencils(h);
  return AFMG_OK;
}

int afmg_set_lsf_boundary_value(afmg_handle* h, double value) {
  AFMG_MULTI(h, afmg_set_lsf_boundary_value(sub_, value));
  if (!h) return AFMG_ERR_ARG;
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  h->o.lsf_boundary_value = value;
  // the kernels read it from device memory (DevCtx::lsf_value_p), so the cached graphs stay valid: the value
  // changes with the applied voltage on every time step (src/m_field.f90:481-487)
  CK(cudaMemcpyAsync(&h->d_comm->lsf_value, &h->o.lsf_boundary_value, sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  h->resid_fresh = false;
  return AFMG_OK;
}

int afmg_set_lsf_boundary_values(afmg_handle* h, int32_t n, const int32_t* box_id, const double* values) {
  AFMG_MULTI(h, afmg_set_lsf_boundary_values(sub_, n, box_id, values));
  if (!h) return AFMG_ERR_ARG;
  if (!h->have_tree) return h->fail(AFMG_ERR_STATE, "afmg_set_tree has not been called");
  if (n < 0 || (n > 0 && (!box_id || !values))) return h->fail(AFMG_ERR_ARG, "afmg_set_lsf_boundary_values: null argument");
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  const int nd = h->o.ndim, nc = h->o.n_cell;
  const int ncell = (nd == 3) ? nc * nc * nc : nc * nc;
  for (int q = 0; q < n; ++q)
    if (box_id[q] < 1 || box_id[q] > h->highest_id || h->id2slot[box_id[q]] < 0)
      return h->fail(AFMG_ERR_ARG, "afmg_set_lsf_boundary_values: unknown box %d", box_id[q]);
  // device order of the values: the colour-split interior [colour][NI] in 3D (like stencil%f), cell order in 2D
  std::vector<double> dev((size_t)std::max(n, 1) * ncell, 0.0);
  for (int q = 0; q < n; ++q) {
    const double* v = values + (size_t)q * ncell;
    double* o = dev.data() + (size_t)q * ncell;
    if (nd == 2) {
      std::copy(v, v + ncell, o);
    } else {
      const int H = nc / 2, NI = nc * nc * H;
      for (int k = 1; k <= nc; ++k)
        for (int j = 1; j <= nc; ++j)
          for (int i = 1; i <= nc; ++i)
            o[((i + j + k) & 1) * NI + ((k - 1) * nc + (j - 1)) * H + ((i - 1) >> 1)] = v[(i - 1) + nc * ((j - 1) + nc * (k - 1))];
    }
  }
  const bool same_list = (int)h->bv_ids.size() == n && std::equal(h->bv_ids.begin(), h->bv_ids.end(), box_id) && n > 0 &&
                         h->d_bv != nullptr;
  h->resid_fresh = false;
  if (same_list) {  // new values for the same boxes (the voltage changed): the pointers in the cached graphs stay valid
    CK(cudaMemcpy(h->d_bv, dev.data(), dev.size() * sizeof(double), cudaMemcpyHostToDevice));
    return AFMG_OK;
  }
  drop_graphs(h);
  h->bv_ids.assign(box_id, box_id + n);
  int rc;
  if (n == 0) {
    h->cx.bvoff = nullptr;
    h->cx.bv = nullptr;
    if (h->s2) {
      h->s2->cx.bvoff = nullptr;
      h->s2->cx.bv = nullptr;
    }
    return AFMG_OK;
  }
  std::vector<long long> off(std::max(h->nslots, 1), -1);
  for (int q = 0; q < n; ++q) off[h->id2slot[box_id[q]]] = (long long)q * ncell;
  if ((rc = dev_upload(h, &h->d_bv, dev))) return rc;
  if ((rc = dev_upload(h, &h->d_bvoff, off))) return rc;
  h->cx.bvoff = h->d_bvoff;
  h->cx.bv = h->d_bv;
  if (h->s2) {
    h->s2->cx.bvoff = h->d_bvoff;
    h->s2->cx.bv = h->d_bv;
  }
  return AFMG_OK;
}

int afmg_set_stencils(afmg_handle* h, int32_t n, const afmg_stencil_desc* desc, const double* blob, int64_t blob_len) {
  AFMG_MULTI(h, afmg_set_stencils(sub_, n, desc, blob, blob_len));
  if (!h) return AFMG_ERR_ARG;
  if (!h->have_tree) return h->fail(AFMG_ERR_STATE, "afmg_set_tree has not been called");
  if (n < 0 || (n > 0 && (!desc || !blob))) return h->fail(AFMG_ERR_ARG, "afmg_set_stencils: null argument");
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  drop_graphs(h);
  h->cs_ready = false;
  h->resid_fresh = false;
  if (h->fs) h->fs->veps_valid = false;  // box tags change
  if (h->o.ndim == 2) return s2_set_stencils(h, n, desc, blob, blob_len);
  return ingest_stencils(h, n, desc, blob, blob_len, false);
}

// mg_set_operators_tree on the device (builders_dev.cuh): tags, operator and prolongation stencils of every box from
// the resident permittivity (afmg_upload(AFMG_EPS)) and / or one of the built-in electrode shapes, then the same
// ingestion as afmg_set_stencils -- without the coefficients crossing PCIe -- and the level-set distance stencils
// of the field computation (afmg_set_lsf_distances).
int afmg_build_stencils_device(afmg_handle* h, const afmg_electrode* electrode, const afmg_lsf_opts* lsf_opts) {
  if (!h) return AFMG_ERR_ARG;
  if (h->is_multi || h->nranks > 1) return h->fail(AFMG_ERR_UNSUPPORTED, "afmg_build_stencils_device: single-GPU handles only (use the host builders)");
  if (!h->have_tree) return h->fail(AFMG_ERR_STATE, "afmg_set_tree has not been called");
  if (h->o.ndim != 3) return h->fail(AFMG_ERR_UNSUPPORTED, "afmg_build_stencils_device: 3D only (2D boxes are built on the host)");
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  const double* d_eps = (h->fs && h->fs->d_eps) ? h->fs->d_eps : nullptr;
  if (!d_eps && !electrode) return h->fail(AFMG_ERR_ARG, "afmg_build_stencils_device: neither AFMG_EPS uploaded nor an electrode given");
  BuildCtx bc{};
  bc.eps = d_eps;
  bc.rmin = h->d_rmin;
  bc.lvl = h->d_lvl;
  bc.parent = h->d_parent;
  bc.coff = h->d_coff;
  for (int d = 0; d < 3; ++d) bc.dr_base[d] = h->o.dr_base[d];
  bc.has_el = electrode ? 1 : 0;
  if (electrode) bc.el = *electrode;
  if (lsf_opts) bc.lo = *lsf_opts;
  else afmg_lsf_opts_default(&bc.lo);
  bc.operator_mask = h->o.operator_mask;
  bc.prolong_auto = h->o.prolongation_type == AFMG_PROLONG_AUTO ? 1 : 0;
  const int total = h->nslots, nc = h->o.n_cell, ncell = nc * nc * nc;
  int *d_tag = nullptr, *d_nb = nullptr, *d_list = nullptr, *d_meta = nullptr;
  CK(cudaMalloc((void**)&d_tag, (size_t)total * sizeof(int)));
  CK(cudaMalloc((void**)&d_nb, (size_t)total * sizeof(int)));
  auto cleanup = [&] {
    cudaFree(d_tag);
    cudaFree(d_nb);
    cudaFree(d_list);
    cudaFree(d_meta);
  };
  {
    Launch L_(h, "build_tags");
    DISPATCH_NC(h, NC, { k_dev_tags<NC><<<total, 256, 0, h->stream>>>(bc, total, d_tag, d_nb); });
  }
  std::vector<int> tags(total), nb(total);
  cudaError_t e = cudaMemcpyAsync(tags.data(), d_tag, (size_t)total * sizeof(int), cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(nb.data(), d_nb, (size_t)total * sizeof(int), cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  if (e != cudaSuccess) {
    cleanup();
    return h->fail(AFMG_ERR_CUDA, "afmg_build_stencils_device: %s", cudaGetErrorString(e));
  }
  if (electrode) {  // check_coarse_representation_lsf (m_af_multigrid.f90:2142-2161): the reference stops here
    bool on_coarse = false;
    for (int s2 = 0; s2 < nlev(h, 1); ++s2) on_coarse = on_coarse || (tags[s2] & AFMG_TAG_LSF_BOX);
    if (!on_coarse) {
      cleanup();
      return h->fail(AFMG_ERR_ARG, "level set function not resolved on coarse grid: no roots found on level 1, use a finer coarse grid");
    }
  }
  std::vector<int> list;
  for (int s2 = 0; s2 < total; ++s2)
    if (tags[s2] != 0) list.push_back(s2);
  const int n = (int)list.size();
  size_t stride = 0;
  DISPATCH_NC(h, NC, stride = BuildBlob<NC>::STRIDE);
  if (h->d_bblob) cudaFree(h->d_bblob);
  h->d_bblob = nullptr;
  std::vector<int> meta((size_t)std::max(n, 1) * 4, 0);
  if (n > 0) {
    e = cudaMalloc((void**)&h->d_bblob, (size_t)n * stride * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_list, (size_t)n * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_meta, (size_t)n * 4 * sizeof(int));
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_list, list.data(), (size_t)n * sizeof(int), cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) {
      Launch L_(h, "build_stencils");
      DISPATCH_NC(h, NC, { k_dev_build<NC><<<n, 256, 0, h->stream>>>(bc, d_list, d_tag, n, h->d_bblob, d_meta); });
      e = cudaMemcpyAsync(meta.data(), d_meta, (size_t)n * 4 * sizeof(int), cudaMemcpyDeviceToHost, h->stream);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) {
      cleanup();
      return h->fail(AFMG_ERR_CUDA, "afmg_build_stencils_device: %s", cudaGetErrorString(e));
    }
  }
  cleanup();
  // ---- descriptors over the device blob, then the common ingestion
  size_t oF = 0, oPV = 0, oDD = 0;
  DISPATCH_NC(h, NC, { oF = BuildBlob<NC>::F; oPV = BuildBlob<NC>::PV; oDD = BuildBlob<NC>::DD; });
  std::vector<afmg_stencil_desc> desc(std::max(n, 1));
  h->b_ids.assign(n, 0);
  h->b_tags.assign(n, 0);
  h->b_meta = meta;
  for (int q = 0; q < n; ++q) {
    afmg_stencil_desc& d = desc[q];
    const int64_t base = (int64_t)q * (int64_t)stride;
    d.box_id = h->slot2id[list[q]];
    d.tag = tags[list[q]];
    d.op_stype = meta[4 * q];
    d.op_offset = base;
    d.f_offset = meta[4 * q + 1] ? base + (int64_t)oF : -1;
    d.prolong_shape = meta[4 * q + 2] ? AFMG_STENCIL_P234 : 0;
    d.prolong_stype = meta[4 * q + 2];
    d.prolong_offset = base + (int64_t)oPV;
    d.cylindrical_gradient = 0;
    h->b_ids[q] = d.box_id;
    h->b_tags[q] = d.tag;
  }
  drop_graphs(h);
  h->cs_ready = false;
  h->resid_fresh = false;
  if (h->fs) h->fs->veps_valid = false;
  int rc = ingest_stencils(h, n, desc.data(), h->d_bblob, (int64_t)n * (int64_t)stride, true);
  if (rc) return rc;
  // ---- the sparse distance stencils of the field computation at the electrode (mg_box_lpllsf_gradient): only the
  // boxes with an internal boundary, a few KB each
  if (electrode) {
    std::vector<int32_t> ids, nent, cells;
    std::vector<double> dds, vals, ddbox((size_t)6 * ncell);
    for (int q = 0; q < n; ++q) {
      if (!(tags[list[q]] & AFMG_TAG_LSF_BOX)) continue;
      CK(cudaMemcpy(ddbox.data(), h->d_bblob + (size_t)q * stride + oDD, ddbox.size() * sizeof(double), cudaMemcpyDeviceToHost));
      const int s2 = list[q];
      const double fac = std::pow(0.5, h->h_lvl[s2] - 1);
      int ne = 0;
      for (int c = 0; c < ncell; ++c) {
        bool any = false;
        for (int m = 0; m < 6; ++m) any = any || ddbox[(size_t)6 * c + m] < 1.0;
        if (!any) continue;
        const int ijk[3] = {c % nc + 1, (c / nc) % nc + 1, c / (nc * nc) + 1};
        double r[3];
        for (int d = 0; d < 3; ++d) {
          cells.push_back(ijk[d]);
          r[d] = h->h_rmin[(size_t)s2 * 3 + d] + (ijk[d] - 0.5) * (h->o.dr_base[d] * fac);
        }
        for (int m = 0; m < 6; ++m) dds.push_back(ddbox[(size_t)6 * c + m]);
        vals.push_back(builders::electrode_lsf_value(electrode, r));
        ++ne;
      }
      ids.push_back(h->slot2id[s2]);
      nent.push_back(ne);
    }
    if (!ids.empty() && (rc = afmg_set_lsf_distances(h, (int)ids.size(), ids.data(), nent.data(), cells.data(), dds.data(), vals.data())))
      return rc;
  }
  return AFMG_OK;
}

// what the last afmg_build_stencils_device produced, in the reference's order (for checks against the host builders):
// per tagged box its id, tag, meta (op_stype, has_f, prolongation stype, 0) and the record v(7, cells) | f(cells) |
// pv(4, cells) | dd(6, cells).  Pass null pointers to query *n only.
int afmg_built_stencils(afmg_handle* h, int32_t* n, int32_t* box_id, int32_t* tag, int32_t* meta, double* blob) {
  if (!h || !n) return AFMG_ERR_ARG;
  *n = (int32_t)h->b_ids.size();
  if (box_id) std::copy(h->b_ids.begin(), h->b_ids.end(), box_id);
  if (tag) std::copy(h->b_tags.begin(), h->b_tags.end(), tag);
  if (meta) std::copy(h->b_meta.begin(), h->b_meta.begin() + 4 * h->b_ids.size(), meta);
  if (blob && !h->b_ids.empty()) {
    CK(cudaSetDevice(h->device));
    size_t stride = 0;
    DISPATCH_NC(h, NC, stride = BuildBlob<NC>::STRIDE);
    CK(cudaMemcpy(blob, h->d_bblob, h->b_ids.size() * stride * sizeof(double), cudaMemcpyDeviceToHost));
  }
  return AFMG_OK;
}

int afmg_update_operator_stencil(afmg_handle* h) {
  AFMG_MULTI(h, afmg_update_operator_stencil(sub_));
  if (!h) return AFMG_ERR_ARG;
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  h->cs_ready = false;
  drop_graphs(h);
  if (h->have_tree) return h->o.ndim == 2 ? s2_constant_stencils(h) : build_constant_stencils(h);
  return AFMG_OK;
}

static int transfer(afmg_handle* h, int var, int n, const int32_t* box_id, double* packed, bool up, bool device_ptr,
                    bool interior = false) {
  if (!h) return AFMG_ERR_ARG;
  if (!h->have_tree) return h->fail(AFMG_ERR_STATE, "afmg_set_tree has not been called");
  if (var == AFMG_EPS || var == AFMG_FLD || var == AFMG_PHOTO) {  // extra variables of the field computation (afmg_field.inc)
    if (device_ptr || interior) return h->fail(AFMG_ERR_UNSUPPORTED, "AFMG_EPS / AFMG_FLD / AFMG_PHOTO move through afmg_upload / afmg_download only");
    if (var == AFMG_EPS && up && h->fs) h->fs->veps_valid = false;
    return fs_transfer(h, var, n, box_id, packed, up);
  }
  if (var < 0 || var > 2) return h->fail(AFMG_ERR_ARG, "invalid variable %d", var);
  if (n == 0) return AFMG_OK;
  CK(cudaSetDevice(h->device));
  if (h->o.ndim == 2) return s2_transfer(h, var, n, box_id, packed, up, device_ptr, interior);
  // slot < 0 marks a box owned by another rank: its packed record is skipped (k_pack / k_unpack)
  std::vector<int> slots(n);
  for (int q = 0; q < n; ++q) {
    const int id = box_id[q];
    if (id < 1 || id > h->highest_id || h->id2slot[id] < 0) return h->fail(AFMG_ERR_ARG, "unknown box id %d", id);
    slots[q] = h->id2slot[id];
    if (h->nranks > 1 && h->h_owner[slots[q]] != h->me) slots[q] = -1;
  }
  const size_t rec_len = interior ? (size_t)h->o.n_cell * h->o.n_cell * h->o.n_cell : (size_t)h->box_len;
  const size_t box_bytes = rec_len * sizeof(double);
  // host transfers go through a two-halves staging area: the PCIe copy of chunk c + 1 (copy stream) overlaps the
  // pack / unpack kernel of chunk c (solver stream)
  const int chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)n, ((size_t)128 << 20) / box_bytes));
  int rc = ensure_stage(h, device_ptr ? 16 : (size_t)2 * chunk * box_bytes, (size_t)n);
  if (rc) return rc;
  CK(cudaMemcpyAsync(h->d_stage_slots, slots.data(), (size_t)n * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  int c = 0;
  bool used[2] = {false, false};
  for (int q0 = 0; q0 < n; q0 += chunk, ++c) {
    const int m = std::min(chunk, n - q0), hb = c & 1;
    double* hp = packed + (size_t)q0 * rec_len;
    double* dp = device_ptr ? hp : h->d_stage + (size_t)hb * chunk * rec_len;
    if (up) {
      if (!device_ptr) {
        if (used[hb]) CK(cudaStreamWaitEvent(h->copy_stream, h->ev_consumed[hb], 0));
        CK(copy_own_records(h, hp, dp, slots.data() + q0, m, rec_len, h->copy_stream, true));
        CK(cudaEventRecord(h->ev_copied[hb], h->copy_stream));
        CK(cudaStreamWaitEvent(h->stream, h->ev_copied[hb], 0));
      }
      {
        Launch L_(h, "unpack");
        if (interior) {
          DISPATCH_NC(h, NC, { launch_k(h, k_unpack_interior<NC>, m, 256, 0, h->d_cc[var], h->d_stage_slots + q0, m, dp); });
        } else {
          DISPATCH_NC(h, NC, { launch_k(h, k_unpack<NC>, m, 256, 0, h->d_cc[var], h->d_stage_slots + q0, m, dp); });
        }
      }
      if (!device_ptr) CK(cudaEventRecord(h->ev_consumed[hb], h->stream));
    } else {
      if (!device_ptr && used[hb]) CK(cudaStreamWaitEvent(h->stream, h->ev_consumed[hb], 0));
      {
        Launch L_(h, "pack");
        if (interior) {
          DISPATCH_NC(h, NC, { launch_k(h, k_pack_interior<NC>, m, 256, 0, h->d_cc[var], h->d_stage_slots + q0, m, dp); });
        } else {
          DISPATCH_NC(h, NC, { launch_k(h, k_pack<NC>, m, 256, 0, h->d_cc[var], h->d_stage_slots + q0, m, dp); });
        }
      }
      if (!device_ptr) {
        CK(cudaEventRecord(h->ev_copied[hb], h->stream));
        CK(cudaStreamWaitEvent(h->copy_stream, h->ev_copied[hb], 0));
        CK(copy_own_records(h, hp, dp, slots.data() + q0, m, rec_len, h->copy_stream));
        CK(cudaEventRecord(h->ev_consumed[hb], h->copy_stream));
      }
    }
    used[hb] = true;
  }
  if (up) h->resid_fresh = false;
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(h->copy_stream));
  CK(cudaStreamSynchronize(h->stream));  // slots vector and (pageable) host buffers must outlive the copies
  return AFMG_OK;
}

int afmg_upload(afmg_handle* h, int32_t var, int32_t n, const int32_t* box_id, const double* packed) {
  AFMG_MULTI(h, afmg_upload(sub_, var, n, box_id, packed));
  return transfer(h, var, n, box_id, const_cast<double*>(packed), true, false);
}
int afmg_upload_interior(afmg_handle* h, int32_t var, int32_t n, const int32_t* box_id, const double* packed) {
  AFMG_MULTI(h, afmg_upload_interior(sub_, var, n, box_id, packed));
  return transfer(h, var, n, box_id, const_cast<double*>(packed), true, false, true);
}
int afmg_download(afmg_handle* h, int32_t var, int32_t n, const int32_t* box_id, double* packed) {
  AFMG_MULTI(h, afmg_download(sub_, var, n, box_id, packed));
  return transfer(h, var, n, box_id, packed, false, false);
}
int afmg_download_interior(afmg_handle* h, int32_t var, int32_t n, const int32_t* box_id, double* packed) {
  AFMG_MULTI(h, afmg_download_interior(sub_, var, n, box_id, packed));
  return transfer(h, var, n, box_id, packed, false, false, true);
}
// page-locked host memory for the caller's packed buffers: copies from / to it are true DMA transfers (a pageable
// buffer is staged by the driver at a fraction of the PCIe rate)
void* afmg_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return p;
}
void afmg_host_free(void* p) {
  if (p) cudaFreeHost(p);
}
int afmg_upload_device(afmg_handle* h, int32_t var, int32_t n, const int32_t* box_id, const double* packed) {
  AFMG_MULTI(h, afmg_upload_device(sub_, var, n, box_id, packed));
  return transfer(h, var, n, box_id, const_cast<double*>(packed), true, true);
}
int afmg_download_device(afmg_handle* h, int32_t var, int32_t n, const int32_t* box_id, double* packed) {
  AFMG_MULTI(h, afmg_download_device(sub_, var, n, box_id, packed));
  return transfer(h, var, n, box_id, packed, false, true);
}

// field_set_rhs (src/m_field.f90:406-444) with the accumulation on the device: the species densities go up (or are
// already resident: on_device) one after the other, chunk by chunk through the two-halves staging area, and each is
// added with its charge factor in the reference's order, so the result has the reference's bits.
int afmg_field_set_rhs(afmg_handle* h, int32_t n, const int32_t* box_id, int32_t n_species, const double* charges,
                       const double* const* densities, int32_t on_device) {
  AFMG_MULTI(h, afmg_field_set_rhs(sub_, n, box_id, n_species, charges, densities, on_device));
  if (!h) return AFMG_ERR_ARG;
  if (!h->have_tree) return h->fail(AFMG_ERR_STATE, "afmg_set_tree has not been called");
  if (n < 0 || n_species < 1 || !charges || !densities || (n > 0 && !box_id))
    return h->fail(AFMG_ERR_ARG, "afmg_field_set_rhs: invalid arguments");
  for (int sp = 0; sp < n_species; ++sp)
    if (!densities[sp]) return h->fail(AFMG_ERR_ARG, "afmg_field_set_rhs: density %d is null", sp + 1);
  if (n == 0) return AFMG_OK;
  CK(cudaSetDevice(h->device));
  std::vector<int> slots(n);
  for (int q = 0; q < n; ++q) {
    const int id = box_id[q];
    if (id < 1 || id > h->highest_id || h->id2slot[id] < 0) return h->fail(AFMG_ERR_ARG, "unknown box id %d", id);
    slots[q] = h->id2slot[id];
    if (h->nranks > 1 && h->h_owner[slots[q]] != h->me) slots[q] = -1;
  }
  const size_t rec_len = (size_t)h->box_len, box_bytes = rec_len * sizeof(double);
  const int chunk = (int)std::max<size_t>(1, std::min<size_t>((size_t)n, ((size_t)128 << 20) / box_bytes));
  int rc = ensure_stage(h, on_device ? 16 : (size_t)2 * chunk * box_bytes, (size_t)n);
  if (rc) return rc;
  CK(cudaMemcpyAsync(h->d_stage_slots, slots.data(), (size_t)n * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  double* rhs = h->o.ndim == 2 ? h->s2->cx.cc[AFMG_RHS] : h->d_cc[AFMG_RHS];
  int c = 0;
  bool used[2] = {false, false};
  for (int q0 = 0; q0 < n; q0 += chunk) {
    const int m = std::min(chunk, n - q0);
    for (int sp = 0; sp < n_species; ++sp, ++c) {
      const int hb = c & 1;
      const double* src = densities[sp] + (size_t)q0 * rec_len;
      const double* dp = src;
      if (!on_device) {
        double* stage = h->d_stage + (size_t)hb * chunk * rec_len;
        if (used[hb]) CK(cudaStreamWaitEvent(h->copy_stream, h->ev_consumed[hb], 0));
        CK(cudaMemcpyAsync(stage, src, (size_t)m * box_bytes, cudaMemcpyHostToDevice, h->copy_stream));
        CK(cudaEventRecord(h->ev_copied[hb], h->copy_stream));
        CK(cudaStreamWaitEvent(h->stream, h->ev_copied[hb], 0));
        dp = stage;
      }
      {
        Launch L_(h, "set_rhs");
        if (h->o.ndim == 2) {
          launch_k(h, afmg2::k2_axpy_boxes, m, 64, 0, rhs, h->d_stage_slots + q0, m, dp, (int)rec_len, charges[sp], sp == 0 ? 1 : 0);
        } else {
          DISPATCH_NC(h, NC, {
            launch_k(h, k_unpack_axpy<NC>, m, 256, 0, rhs, h->d_stage_slots + q0, m, dp, charges[sp], sp == 0 ? 1 : 0);
          });
        }
      }
      if (!on_device) CK(cudaEventRecord(h->ev_consumed[hb], h->stream));
      used[hb] = true;
    }
  }
  h->resid_fresh = false;
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(h->copy_stream));
  return finish_op(h);
}

int afmg_clear(afmg_handle* h, int32_t var) {
  AFMG_MULTI(h, afmg_clear(sub_, var));
  if (!h) return AFMG_ERR_ARG;
  if (!h->have_tree) return h->fail(AFMG_ERR_STATE, "afmg_set_tree has not been called");
  if (var < 0 || var > 2) return h->fail(AFMG_ERR_ARG, "invalid variable %d", var);
  CK(cudaSetDevice(h->device));
  if (h->o.ndim == 3 && h->nranks > 1) {  // only the records this rank owns (the others may not even be mapped)
    for (int l = 1; l <= h->L; ++l) {
      const Range r = own(h, l);
      if (r.n > 0)
        CK(cudaMemsetAsync(h->d_cc[var] + (size_t)r.s0 * h->box_len, 0, (size_t)r.n * h->box_len * sizeof(double), h->stream));
    }
  } else {
    CK(cudaMemsetAsync(h->o.ndim == 2 ? h->s2->cx.cc[var] : h->d_cc[var], 0, (size_t)h->nslots * h->box_len * sizeof(double),
                       h->stream));
  }
  h->resid_fresh = false;
  return finish_op(h);
}

int afmg_fas_vcycle_async(afmg_handle* h, int32_t set_residual, int32_t highest_lvl, int32_t n_cycles) {
  AFMG_MULTI(h, afmg_fas_vcycle_async(sub_, set_residual, highest_lvl, n_cycles));
  if (!h) return AFMG_ERR_ARG;
  int rc = ensure_ready(h);
  if (rc) return rc;
  CK(cudaSetDevice(h->device));
  const int max_lvl = highest_lvl > 0 ? highest_lvl : h->L;
  if ((rc = check_lvl(h, max_lvl))) return rc;
  CK(cudaEventRecord(h->ev0, h->stream));
  rc = run_graph(h, {0, set_residual ? 1 : 0, max_lvl}, n_cycles, [&] {
    if (h->o.ndim == 2) s2_vcycle(h, set_residual != 0, max_lvl);
    else enq_vcycle(h, set_residual != 0, max_lvl);
  });
  if (rc) return rc;
  CK(cudaEventRecord(h->ev1, h->stream));
  h->ev_valid = true;
  h->resid_fresh = set_residual && max_lvl == h->L;
  return AFMG_OK;
}

int afmg_fas_fmg_async(afmg_handle* h, int32_t set_residual, int32_t have_guess, int32_t n_cycles) {
  AFMG_MULTI(h, afmg_fas_fmg_async(sub_, set_residual, have_guess, n_cycles));
  if (!h) return AFMG_ERR_ARG;
  int rc = ensure_ready(h);
  if (rc) return rc;
  CK(cudaSetDevice(h->device));
  CK(cudaEventRecord(h->ev0, h->stream));
  rc = run_graph(h, {1, set_residual ? 1 : 0, have_guess ? 1 : 0}, n_cycles,
                 [&] {
                   if (h->o.ndim == 2) s2_fmg(h, set_residual != 0, have_guess != 0);
                   else enq_fmg(h, set_residual != 0, have_guess != 0);
                 });
  if (rc) return rc;
  CK(cudaEventRecord(h->ev1, h->stream));
  h->ev_valid = true;
  h->resid_fresh = set_residual != 0;
  return AFMG_OK;
}

int afmg_sync(afmg_handle* h) {
  AFMG_MULTI(h, afmg_sync(sub_));
  if (!h) return AFMG_ERR_ARG;
  CK(cudaSetDevice(h->device));
  return finish_op(h);
}

int afmg_fas_vcycle(afmg_handle* h, int32_t set_residual, int32_t highest_lvl, int32_t standalone) {
  AFMG_MULTI(h, afmg_fas_vcycle(sub_, set_residual, highest_lvl, standalone));
  (void)standalone;  // mg_use / done_with_mg bookkeeping lives in the Fortran shim
  int rc = afmg_fas_vcycle_async(h, set_residual, highest_lvl, 1);
  if (rc) return rc;
  return afmg_sync(h);
}

int afmg_fas_fmg(afmg_handle* h, int32_t set_residual, int32_t have_guess) {
  AFMG_MULTI(h, afmg_fas_fmg(sub_, set_residual, have_guess));
  int rc = afmg_fas_fmg_async(h, set_residual, have_guess, 1);
  if (rc) return rc;
  return afmg_sync(h);
}

int afmg_field_solve(afmg_handle* h, int32_t have_guess, double residual_threshold, double max_residual, int32_t max_fmg,
                     int32_t n_vcycles, double* residuals, int32_t* n_fmg, int32_t* n_vc) {
  if (h && h->is_multi) {
    if (!residuals || !n_fmg || !n_vc) return AFMG_ERR_ARG;
    const int ng = (int)h->subs.size(), cap = std::max(1, max_fmg + n_vcycles);
    std::vector<double> res((size_t)ng * cap, 0.0);
    std::vector<int32_t> nf(ng, 0), nv(ng, 0);
    const int rc = multi_call(h, [&, have_guess, residual_threshold, max_residual, max_fmg, n_vcycles](afmg_handle* sub_) -> int {
      const int r = sub_->me;
      return afmg_field_solve(sub_, have_guess, residual_threshold, max_residual, max_fmg, n_vcycles, &res[(size_t)r * cap], &nf[r], &nv[r]);
    });
    std::copy(res.begin(), res.begin() + cap, residuals);  // identical on every rank
    *n_fmg = nf[0];
    *n_vc = nv[0];
    return rc;
  }
  if (!h || !residuals || !n_fmg || !n_vc) return AFMG_ERR_ARG;
  *n_fmg = 0;
  *n_vc = 0;
  int rc, k = 0;
  if (!have_guess) {
    bool ok = false;
    for (int i = 0; i < max_fmg; ++i) {
      if ((rc = afmg_fas_fmg_async(h, 1, 1, 1))) return rc;
      if ((rc = afmg_max_abs(h, AFMG_TMP, &residuals[k]))) return rc;
      ++k;
      ++*n_fmg;
      if (residuals[k - 1] < residual_threshold) {
        ok = true;
        break;
      }
      if (i >= 2) {  // i > 2 in the reference's 1-based loop
        const double lo = std::min({residuals[k - 3], residuals[k - 2], residuals[k - 1]});
        const double hi = std::max({residuals[k - 3], residuals[k - 2], residuals[k - 1]});
        const double ratio = lo / hi;
        if (ratio < 2.0 && ratio > 0.5 && residuals[k - 1] < max_residual) {
          ok = true;
          break;
        }
      }
    }
    if (!ok) return h->fail(AFMG_ERR_NOT_CONVERGED, "no convergence in initial field computation after %d FMG cycles", max_fmg);
  }
  for (int i = 0; i < n_vcycles; ++i) {
    if ((rc = afmg_fas_vcycle_async(h, 1, 0, 1))) return rc;
    if ((rc = afmg_max_abs(h, AFMG_TMP, &residuals[k]))) return rc;
    ++k;
    ++*n_vc;
    if (residuals[k - 1] < residual_threshold) break;
  }
  return afmg_sync(h);
}

// ---- single operations ------------------------------------------------------------------------------
#define SINGLE_OP_PROLOGUE()          \
  if (!h) return AFMG_ERR_ARG;        \
  {                                   \
    int rc_ = ensure_ready(h);        \
    if (rc_) return rc_;              \
  }                                   \
  CK(cudaSetDevice(h->device));       \
  h->resid_fresh = false;

int afmg_gsrb_boxes(afmg_handle* h, int32_t lvl, int32_t type_cycle) {
  AFMG_MULTI(h, afmg_gsrb_boxes(sub_, lvl, type_cycle));
  SINGLE_OP_PROLOGUE();
  int rc = check_lvl(h, lvl);
  if (rc) return rc;
  if (type_cycle != 1 && type_cycle != 3) return h->fail(AFMG_ERR_ARG, "gsrb_boxes: invalid cycle type");
  if (h->o.ndim == 2) {
    s2_rb_prepare(h, lvl);
    s2_gsrb_boxes(h, lvl, type_cycle == 3);
    return finish_op(h);
  }
  if (int rc_ = run_direct(h, [&] {
        enq_rb_prepare(h, lvl);
        enq_gsrb_boxes(h, lvl, type_cycle == 3);
      }))
    return rc_;
  return finish_op(h);
}

int afmg_gsrb_halfsweep(afmg_handle* h, int32_t lvl, int32_t redblack) {
  AFMG_MULTI(h, afmg_gsrb_halfsweep(sub_, lvl, redblack));
  SINGLE_OP_PROLOGUE();
  int rc = check_lvl(h, lvl);
  if (rc) return rc;
  if (h->o.ndim == 2) {
    s2_rb_prepare(h, lvl);
    s2_gsrb(h, lvl, redblack);
    s2_gc(h, lvl, 0);
    return finish_op(h);
  }
  if (int rc_ = run_direct(h, [&] {
        enq_rb_prepare(h, lvl);
        enq_gsrb(h, lvl, redblack);
      }))
    return rc_;
  return finish_op(h);
}

int afmg_gc_lvl(afmg_handle* h, int32_t lvl, int32_t var, int32_t corners) {
  AFMG_MULTI(h, afmg_gc_lvl(sub_, lvl, var, corners));
  SINGLE_OP_PROLOGUE();
  int rc = check_lvl(h, lvl);
  if (rc) return rc;
  if (var != AFMG_PHI) return h->fail(AFMG_ERR_UNSUPPORTED, "ghost cells are only defined for phi on this path");
  if (h->o.ndim == 2) {
    s2_rb_prepare(h, lvl);
    s2_gc(h, lvl, corners);
    return finish_op(h);
  }
  if (int rc_ = run_direct(h, [&] {
        enq_rb_prepare(h, lvl);
        enq_gc(h, lvl, var, corners, 0);
      }))
    return rc_;
  return finish_op(h);
}

int afmg_update_coarse(afmg_handle* h, int32_t lvl, int32_t with_tmp) {
  AFMG_MULTI(h, afmg_update_coarse(sub_, lvl, with_tmp));
  SINGLE_OP_PROLOGUE();
  int rc = check_lvl(h, lvl);
  if (rc) return rc;
  if (lvl < 2) return h->fail(AFMG_ERR_ARG, "update_coarse needs lvl >= 2");
  if (h->o.ndim == 2) s2_update_coarse(h, lvl, with_tmp != 0);
  else if (int rc_ = run_direct(h, [&] { enq_update_coarse(h, lvl, with_tmp != 0); })) return rc_;
  return finish_op(h);
}

int afmg_correct_children(afmg_handle* h, int32_t lvl_parents) {
  AFMG_MULTI(h, afmg_correct_children(sub_, lvl_parents));
  SINGLE_OP_PROLOGUE();
  int rc = check_lvl(h, lvl_parents);
  if (rc) return rc;
  if (h->o.ndim == 2) s2_correct(h, lvl_parents, true);
  else if (int rc_ = run_direct(h, [&] { enq_correct(h, lvl_parents, true, false); })) return rc_;
  return finish_op(h);
}

int afmg_correct_children_gc(afmg_handle* h, int32_t lvl_parents) {
  AFMG_MULTI(h, afmg_correct_children_gc(sub_, lvl_parents));
  SINGLE_OP_PROLOGUE();
  int rc = check_lvl(h, lvl_parents);
  if (rc) return rc;
  if (lvl_parents >= h->L) return h->fail(AFMG_ERR_ARG, "no level above %d", lvl_parents);
  if (h->o.ndim == 2) s2_correct_gc(h, lvl_parents, true);
  else if (int rc_ = run_direct(h, [&] { enq_correct_gc(h, lvl_parents, true); })) return rc_;
  return finish_op(h);
}

int afmg_residual_lvl(afmg_handle* h, int32_t lvl) {
  AFMG_MULTI(h, afmg_residual_lvl(sub_, lvl));
  SINGLE_OP_PROLOGUE();
  int rc = check_lvl(h, lvl);
  if (rc) return rc;
  if (h->o.ndim == 2) s2_residual(h, lvl, lvl, false);
  else if (int rc_ = run_direct(h, [&] { enq_residual(h, lvl, lvl, false); })) return rc_;
  return finish_op(h);
}

int afmg_solve_coarse_grid(afmg_handle* h) {
  AFMG_MULTI(h, afmg_solve_coarse_grid(sub_));
  SINGLE_OP_PROLOGUE();
  if (h->o.ndim == 2) s2_coarse(h);
  else if (int rc_ = run_direct(h, [&] { enq_coarse(h); })) return rc_;
  return finish_op(h);
}

int afmg_init_phi_rhs(afmg_handle* h) {
  AFMG_MULTI(h, afmg_init_phi_rhs(sub_));
  SINGLE_OP_PROLOGUE();
  if (h->o.ndim == 2) s2_init_phi_rhs(h);
  else if (int rc_ = run_direct(h, [&] { enq_init_phi_rhs(h); })) return rc_;
  return finish_op(h);
}

int afmg_max_abs(afmg_handle* h, int32_t var, double* out) {
  if (h && h->is_multi) {
    if (!out) return AFMG_ERR_ARG;
    std::vector<double> v(h->subs.size(), 0.0);
    const int rc = multi_call(h, [&v, var](afmg_handle* sub_) -> int { return afmg_max_abs(sub_, var, &v[sub_->me]); });
    *out = v[0];  // the combined value, identical on every rank
    return rc;
  }
  if (!h || !out) return AFMG_ERR_ARG;
  if (!h->have_tree) return h->fail(AFMG_ERR_STATE, "afmg_set_tree has not been called");
  if (var < 0 || var > 2) return h->fail(AFMG_ERR_ARG, "invalid variable %d", var);
  CK(cudaSetDevice(h->device));
  unsigned long long bits = 0;
  const int comb = (h->nranks > 1) ? 4 : 0;  // multi-GPU: the value combined over all ranks
  if (var == AFMG_TMP && h->resid_fresh) {
    CK(cudaMemcpyAsync(&bits, h->d_scal + comb, sizeof bits, cudaMemcpyDeviceToHost, h->stream));
  } else {
    if (h->nranks > 1 && !h->connected) return h->fail(AFMG_ERR_STATE, "multi-GPU handle is not connected");
    CK(cudaMemsetAsync(h->d_scal + 1, 0, sizeof(unsigned long long), h->stream));
    if (h->o.ndim == 2) {
      Launch L_(h, "maxabs");
      DISPATCH_NC2(h, NC, { launch_k(h, afmg2::k2_maxabs<NC>, h->nslots, 64, 0, h->s2->cx, 0, h->nslots, var, h->d_scal + 1); });
    }
    for (int l = 1; l <= h->L && h->o.ndim == 3; ++l) {
      const Range r = (h->nranks == 1) ? Range{0, h->nslots} : own(h, l);
      if (r.n > 0) {
        Launch L_(h, "maxabs");
        DISPATCH_NC(h, NC, { launch_k(h, k_maxabs<NC>, r.n, 256, 0, h->cx, r.s0, r.n, var, h->d_scal + 1); });
      }
      if (h->nranks == 1) break;
    }
    enq_allmax(h, 1);
    CK(cudaMemcpyAsync(&bits, h->d_scal + 1 + comb, sizeof bits, cudaMemcpyDeviceToHost, h->stream));
  }
  int rc = finish_op(h);
  if (rc) return rc;
  std::memcpy(out, &bits, sizeof bits);
  return AFMG_OK;
}

int afmg_tree_sum(afmg_handle* h, int32_t var, double* out) {
  if (h && h->is_multi) {
    if (!out) return AFMG_ERR_ARG;
    std::vector<double> v(h->subs.size(), 0.0);
    const int rc = multi_call(h, [&v, var](afmg_handle* sub_) -> int { return afmg_tree_sum(sub_, var, &v[sub_->me]); });
    *out = v[0];
    return rc;
  }
  if (!h || !out) return AFMG_ERR_ARG;
  if (!h->have_tree) return h->fail(AFMG_ERR_STATE, "afmg_set_tree has not been called");
  if (var < 0 || var > 2) return h->fail(AFMG_ERR_ARG, "invalid variable %d", var);
  CK(cudaSetDevice(h->device));
  if (h->nranks > 1 && !h->connected) return h->fail(AFMG_ERR_STATE, "multi-GPU handle is not connected");
  if (h->o.ndim == 2) s2_box_sums(h, var);
  else enq_box_sums(h, var);
  std::vector<double> sums(h->nslots);
  CK(cudaMemcpyAsync(sums.data(), h->o.ndim == 2 ? h->s2->d_boxsum : h->d_boxsum, (size_t)h->nslots * sizeof(double),
                     cudaMemcpyDeviceToHost, h->stream));
  int rc = finish_op(h);
  if (rc) return rc;
  // af_tree_sum_cc: my_sum += fac(lvl) * box_sum over leaves, in level / list order
  double s = 0.0;
  for (int l = 1; l <= h->L; ++l) {
    double fac = 1.0;
    for (int d = 0; d < h->o.ndim; ++d) fac *= h->o.dr_base[d] * std::pow(0.5, l - 1);
    const std::vector<int>& child0 = h->o.ndim == 2 ? h->s2->h_child0 : h->h_child0;
    for (int q = h->lvl_off[l]; q < h->lvl_off[l + 1]; ++q)
      if (child0[q] < 0) s = s + fac * sums[q];
  }
  *out = s;
  return AFMG_OK;
}

int afmg_checksum(afmg_handle* h, int32_t var, uint64_t* sum_out, uint64_t* xor_out) {
  if (h && h->is_multi) {
    if (!sum_out || !xor_out) return AFMG_ERR_ARG;
    std::vector<uint64_t> a(h->subs.size(), 0), b(h->subs.size(), 0);
    const int rc = multi_call(h, [&a, &b, var](afmg_handle* sub_) -> int { return afmg_checksum(sub_, var, &a[sub_->me], &b[sub_->me]); });
    *sum_out = 0;
    *xor_out = 0;
    for (size_t r = 0; r < a.size(); ++r) {  // every rank covers the boxes it owns
      *sum_out += a[r];
      *xor_out ^= b[r];
    }
    return rc;
  }
  if (!h || !sum_out || !xor_out) return AFMG_ERR_ARG;
  if (!h->have_tree) return h->fail(AFMG_ERR_STATE, "afmg_set_tree has not been called");
  if (var < 0 || var > 2) return h->fail(AFMG_ERR_ARG, "invalid variable %d", var);
  CK(cudaSetDevice(h->device));
  unsigned long long* d_out = h->d_scal + 6;  // scal[6], scal[7]: not used by the reductions
  CK(cudaMemsetAsync(d_out, 0, 2 * sizeof(unsigned long long), h->stream));
  const double* base = h->o.ndim == 2 ? h->s2->cx.cc[var] : h->d_cc[var];
  for (int l = 1; l <= h->L; ++l) {
    const Range r = (h->nranks == 1) ? Range{0, h->nslots} : own(h, l);
    const size_t n = (size_t)r.n * h->box_len;
    if (n > 0) {
      Launch L_(h, "checksum");
      const int blocks = (int)std::min<size_t>((n + 1023) / 1024, 148 * 8);
      launch_k(h, k_checksum, blocks, 256, 0, base + (size_t)r.s0 * h->box_len, n, d_out);
    }
    if (h->nranks == 1) break;
  }
  unsigned long long out[2] = {0, 0};
  CK(cudaMemcpyAsync(out, d_out, sizeof out, cudaMemcpyDeviceToHost, h->stream));
  int rc = finish_op(h);
  if (rc) return rc;
  *sum_out = out[0];
  *xor_out = out[1];
  return AFMG_OK;
}

int afmg_set_mega(afmg_handle* h, int32_t enabled, int32_t max_boxes) {
  AFMG_MULTI(h, afmg_set_mega(sub_, enabled, max_boxes));
  if (!h) return AFMG_ERR_ARG;
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  drop_graphs(h);  // the cached cycles were built for the previous setting
  h->mega_enabled = enabled != 0;
  h->mega_max_boxes = max_boxes > 0 ? max_boxes : 0;
  return AFMG_OK;
}

// device memory of the cell-data slab per GPU: bytes physically mapped and bytes of the full slot space
int afmg_slab_bytes(afmg_handle* h, int32_t cap, int64_t* mapped, int64_t* full, int32_t* n) {
  if (!h || !n) return AFMG_ERR_ARG;
  std::vector<afmg_handle*> hs = h->is_multi ? h->subs : std::vector<afmg_handle*>{h};
  *n = (int32_t)hs.size();
  for (int r = 0; r < (int)hs.size() && r < cap; ++r) {
    if (mapped) mapped[r] = (int64_t)hs[r]->slab_mapped_bytes;
    if (full) full[r] = (int64_t)hs[r]->slab_bytes;
  }
  return AFMG_OK;
}

int32_t afmg_mega_active(const afmg_handle* h) { return (h && mega_possible(h)) ? h->mega_grid : 0; }

int64_t afmg_kernel_launches(const afmg_handle* h) {
  if (!h) return 0;
  int64_t n = h->launches;
  for (const afmg_handle* sub : h->subs) n += sub->launches;
  return n;
}

int afmg_last_cycle_ms(afmg_handle* h, double* ms) {
  if (h && h->is_multi) {
    if (!ms) return AFMG_ERR_ARG;
    std::vector<double> v(h->subs.size(), 0.0);
    const int rc = multi_call(h, [&v](afmg_handle* sub_) -> int { return afmg_last_cycle_ms(sub_, &v[sub_->me]); });
    *ms = *std::max_element(v.begin(), v.end());  // device time of the slowest GPU
    return rc;
  }
  if (!h || !ms) return AFMG_ERR_ARG;
  if (!h->ev_valid) return h->fail(AFMG_ERR_STATE, "no cycle has been run");
  CK(cudaEventSynchronize(h->ev1));
  float f = 0;
  CK(cudaEventElapsedTime(&f, h->ev0, h->ev1));
  *ms = f;
  return AFMG_OK;
}

int afmg_set_profiling(afmg_handle* h, int32_t on) {
  AFMG_MULTI(h, afmg_set_profiling(sub_, on));
  if (!h) return AFMG_ERR_ARG;
  CK(cudaStreamSynchronize(h->stream));
  prof_resolve(h);
  mega_resolve_stamps(h);
  h->profiling = on != 0;
  if (on) h->prof.clear();
  return AFMG_OK;
}

int afmg_profile(afmg_handle* h, int32_t cap, char (*names)[32], double* ms, int64_t* calls, int32_t* n) {
  if (h && h->is_multi) return afmg_profile(h->subs[0], cap, names, ms, calls, n);  // rank 0's kernels
  if (!h || !n) return AFMG_ERR_ARG;
  prof_resolve(h);
  mega_resolve_stamps(h);
  int k = 0;
  for (auto& kv : h->prof) {
    if (k >= cap) break;
    std::snprintf(names[k], 32, "%s", kv.first.c_str());
    ms[k] = kv.second.ms;
    calls[k] = kv.second.calls;
    ++k;
  }
  *n = k;
  return AFMG_OK;
}

int afmg_cell_updates(afmg_handle* h, int32_t highest_lvl, int32_t fmg, double* out) {
  if (h && h->is_multi) return afmg_cell_updates(h->subs[0], highest_lvl, fmg, out);  // whole tree, host arithmetic
  if (!h || !out) return AFMG_ERR_ARG;
  if (!h->have_tree) return h->fail(AFMG_ERR_STATE, "afmg_set_tree has not been called");
  const int maxl = highest_lvl > 0 ? highest_lvl : h->L;
  const double ncell = (double)h->o.n_cell * h->o.n_cell * (h->o.ndim == 3 ? h->o.n_cell : 1);
  auto vc = [&](int m) {
    double s = 0;
    for (int l = 2; l <= m; ++l) s += (double)(h->o.n_cycle_down + h->o.n_cycle_up) * nlev(h, l) * ncell;
    return s;
  };
  double s = 0;
  if (fmg) for (int m = 2; m <= h->L; ++m) s += vc(m);
  else s = vc(maxl);
  *out = s;
  return AFMG_OK;
}

int32_t afmg_layout_offset(int32_t ndim, int32_t nc, int32_t i, int32_t j, int32_t k) {
  if (ndim == 3) {
    switch (nc) {
      case 4: return Lay3<4>::cell(i, j, k);
      case 8: return Lay3<8>::cell(i, j, k);
      case 16: return Lay3<16>::cell(i, j, k);
      case 32: return Lay3<32>::cell(i, j, k);
    }
  } else if (ndim == 2) {
    // 2D box records keep the reference's own order cc(0:nc+1, 0:nc+1), first index fastest
    if (nc == 4 || nc == 8 || nc == 16 || nc == 32) return i + (nc + 2) * j;
  }
  return -1;
}

int32_t afmg_layout_box_len(int32_t ndim, int32_t nc) {
  return ndim == 3 ? (nc + 2) * (nc + 2) * (nc + 2) : (nc + 2) * (nc + 2);
}

// Morton (Z-order) key with x in the lowest bit, afivo's convention (m_morton.f90; known answers
// afivo/tests/answers/test_morton_2d, _3d).  The device slot order inside a level and the multi-GPU cuts follow
// this key of ix - 1 (in 2D the 3D interleave with z = 0 is used internally: same order, different key values).
int64_t afmg_morton_key(int32_t ndim, int32_t ix, int32_t iy, int32_t iz) {
  if (ndim == 3) return (int64_t)morton3((uint32_t)ix, (uint32_t)iy, (uint32_t)iz);
  uint64_t r = 0;
  for (int b = 0; b < 31; ++b) {
    r |= (uint64_t)(((uint32_t)ix >> b) & 1u) << (2 * b);
    r |= (uint64_t)(((uint32_t)iy >> b) & 1u) << (2 * b + 1);
  }
  return (int64_t)r;
}

int32_t afmg_slot_of_box(const afmg_handle* h, int32_t box_id) {
  if (h && h->is_multi) return afmg_slot_of_box(h->subs[0], box_id);
  if (!h || !h->have_tree || box_id < 1 || box_id > h->highest_id) return -1;
  return h->id2slot[box_id];
}

// Contiguous Morton ranges per level, cut at sibling groups of 2^ndim boxes so that the children of a
// box never straddle two ranks; level 1 (the coarse grid) stays on rank 0.  cuts[(l-1)*(n_ranks+1)+r] =
// first box (position in the level's Morton order) of rank r on level l; the last entry is the count.
int afmg_partition(int32_t n_ranks, int32_t highest_lvl, const int32_t* lvl_counts, int32_t* cuts) {
  return afmg_partition_min(n_ranks, highest_lvl, lvl_counts, 0, cuts);
}

// min_split_boxes: levels with fewer boxes are not split (they stay on rank 0 with the coarse grid): below a few
// million cells a level is launch-latency bound, splitting it buys nothing and costs a cross-GPU barrier per operation
// (measured on S3, 8 GPUs: 512 boxes of 16^3 take 0.096 ms per gsrb_boxes on one GPU and 0.103 ms on eight)
int afmg_partition_min(int32_t n_ranks, int32_t highest_lvl, const int32_t* lvl_counts, int32_t min_split_boxes,
                       int32_t* cuts) {
  if (n_ranks < 1 || n_ranks > AFMG_MAX_RANKS || highest_lvl < 1 || !lvl_counts || !cuts) return AFMG_ERR_ARG;
  for (int l = 1; l <= highest_lvl; ++l) {
    int32_t* c = cuts + (size_t)(l - 1) * (n_ranks + 1);
    const int n = lvl_counts[l - 1];
    if (l == 1 || n % 8 != 0 || n < min_split_boxes) {
      c[0] = 0;
      for (int r = 1; r <= n_ranks; ++r) c[r] = n;
      continue;
    }
    const long long groups = n / 8;
    // ceil: levels with fewer sibling groups than ranks fill the low ranks (next to the coarse grid)
    for (int r = 0; r <= n_ranks; ++r) c[r] = (int32_t)(8 * ((groups * r + n_ranks - 1) / n_ranks));
  }
  return AFMG_OK;
}

int afmg_comm_init(afmg_handle* h, int32_t n_ranks, int32_t rank) {
  if (h && h->is_multi) return h->fail(AFMG_ERR_STATE, "afmg_comm_init: this handle drives its GPUs itself (afmg_opts.n_gpus)");
  if (!h) return AFMG_ERR_ARG;
  if (n_ranks < 1 || n_ranks > AFMG_MAX_RANKS || rank < 0 || rank >= n_ranks)
    return h->fail(AFMG_ERR_ARG, "afmg_comm_init: need 1 <= n_ranks <= %d and 0 <= rank < n_ranks", AFMG_MAX_RANKS);
  if (h->have_tree) return h->fail(AFMG_ERR_STATE, "afmg_comm_init must be called before afmg_set_tree");
  h->nranks = n_ranks;
  h->me = rank;
  for (auto& p : h->peers.p) p = nullptr;
  h->peers.p[rank] = h->d_comm;
  h->connected = (n_ranks == 1);
  return AFMG_OK;
}

namespace {
struct CommBlob {  // what one rank tells the others (AFMG_COMM_BLOB_BYTES)
  cudaIpcMemHandle_t slab, comm;
  uint64_t slab_bytes;
  int32_t rank, nslots, box_len, pad;
};
static_assert(sizeof(CommBlob) <= AFMG_COMM_BLOB_BYTES, "blob size");
}  // namespace

int afmg_comm_export(afmg_handle* h, void* blob) {
  if (h && h->is_multi) return h->fail(AFMG_ERR_STATE, "afmg_comm_export: this handle drives its GPUs itself (afmg_opts.n_gpus)");
  if (!h || !blob) return AFMG_ERR_ARG;
  if (!h->have_tree) return h->fail(AFMG_ERR_STATE, "afmg_comm_export: call afmg_set_tree first");
  CK(cudaSetDevice(h->device));
  CommBlob b{};
  CK(cudaIpcGetMemHandle(&b.slab, h->d_slab));
  CK(cudaIpcGetMemHandle(&b.comm, h->d_comm));
  b.slab_bytes = h->slab_bytes;
  b.rank = h->me;
  b.nslots = h->nslots;
  b.box_len = h->box_len;
  std::memset(blob, 0, AFMG_COMM_BLOB_BYTES);
  std::memcpy(blob, &b, sizeof b);
  return AFMG_OK;
}

int afmg_comm_connect(afmg_handle* h, const void* blobs) {
  if (h && h->is_multi) return h->fail(AFMG_ERR_STATE, "afmg_comm_connect: this handle drives its GPUs itself (afmg_opts.n_gpus)");
  if (!h || !blobs) return AFMG_ERR_ARG;
  if (!h->have_tree) return h->fail(AFMG_ERR_STATE, "afmg_comm_connect: call afmg_set_tree first");
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  drop_graphs(h);
  for (int r = 0; r < h->nranks; ++r) {
    CommBlob b;
    std::memcpy(&b, (const char*)blobs + (size_t)r * AFMG_COMM_BLOB_BYTES, sizeof b);
    if (b.rank != r || b.nslots != h->nslots || b.box_len != h->box_len || b.slab_bytes != h->slab_bytes)
      return h->fail(AFMG_ERR_ARG, "afmg_comm_connect: blob %d does not describe the same tree (rank %d, %d slots)", r,
                     b.rank, b.nslots);
    if (r == h->me) continue;
    if (!h->peer_slab[r]) {
      void* p = nullptr;
      CK(cudaIpcOpenMemHandle(&p, b.slab, cudaIpcMemLazyEnablePeerAccess));
      h->peer_slab[r] = (char*)p;
    }
    if (!h->peers.p[r]) {
      void* p = nullptr;
      CK(cudaIpcOpenMemHandle(&p, b.comm, cudaIpcMemLazyEnablePeerAccess));
      h->peers.p[r] = (CommBlock*)p;
    }
  }
  for (int r = 0; r < h->nranks; ++r) {
    for (int v = 0; v < 3; ++v) h->cx.ccr[r][v] = (double*)(h->peer_slab[r] + v * h->slab_var_stride);
    h->cx.ccr[r][V_FLD] = (h->slab_nvar == 4) ? (double*)(h->peer_slab[r] + 3 * h->slab_var_stride) : nullptr;
    h->cx.bsum[r] = (double*)(h->peer_slab[r] + h->slab_nvar * h->slab_var_stride);
  }
  h->connected = true;
  return AFMG_OK;
}

static int connect_local(afmg_handle* sub, afmg_handle* front) {
  afmg_handle* h = sub;
  CK(cudaSetDevice(h->device));
  CK(cudaStreamSynchronize(h->stream));
  drop_graphs(h);
  for (int r = 0; r < h->nranks; ++r) {
    afmg_handle* q = front->subs[r];
    if (q->nslots != h->nslots || q->slab_bytes != h->slab_bytes)
      return h->fail(AFMG_ERR_ARG, "internal: rank handles hold different trees");
    if (r == h->me) continue;
    int can = 0;
    CK(cudaDeviceCanAccessPeer(&can, h->device, q->device));
    if (!can) return h->fail(AFMG_ERR_UNSUPPORTED, "GPU %d cannot access the memory of GPU %d (no NVLink / P2P path)", h->device, q->device);
    cudaError_t e = cudaDeviceEnablePeerAccess(q->device, 0);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
      return h->fail(AFMG_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d): %s", q->device, cudaGetErrorString(e));
    cudaGetLastError();
    h->peer_slab[r] = q->d_slab;
    h->peers.p[r] = q->d_comm;
  }
  for (int r = 0; r < h->nranks; ++r) {
    for (int v = 0; v < 3; ++v) h->cx.ccr[r][v] = (double*)(h->peer_slab[r] + v * h->slab_var_stride);
    h->cx.ccr[r][V_FLD] = (h->slab_nvar == 4) ? (double*)(h->peer_slab[r] + 3 * h->slab_var_stride) : nullptr;
    h->cx.bsum[r] = (double*)(h->peer_slab[r] + h->slab_nvar * h->slab_var_stride);
  }
  h->connected = true;
  return AFMG_OK;
}

int32_t afmg_owner_of_box(const afmg_handle* h, int32_t box_id) {
  if (h && h->is_multi) return afmg_owner_of_box(h->subs[0], box_id);
  if (!h || !h->have_tree || box_id < 1 || box_id > h->highest_id || h->id2slot[box_id] < 0) return -1;
  return h->h_owner[h->id2slot[box_id]];
}

#include "afmg_field_api.inc"

}  // extern "C"


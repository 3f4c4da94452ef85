// k_mega: one persistent, cooperatively launched kernel that executes a whole SEGMENT of a multigrid cycle -- the
// operations on the small levels of a tree, or the complete cycle of a streamer-like tree (nc = 8, many levels of a
// few hundred boxes, SURVEY config C3/C4) -- as a list of phases separated by grid-wide barriers.
//
// Why: on such levels a kernel of the launch path costs 5-7 us for < 2 us of work (launch gap + prologue + the
// dependent loads of a cold start), and a V-cycle of gsrb_boxes / update_coarse / correct_children
// (afivo/src/m_af_multigrid.f90:185-264, :648-687) is ~110 of them.  Here a phase boundary costs one barrier
// (~1 us: one atomic per CTA on a single counter + one acquire spin), the CTAs stay resident, and the mbarrier / TMA
// plumbing is set up once.
//
// What a phase is: exactly what one kernel launch is on the launch path (same device functions: gsrb2_vblock,
// resid3_box, correct3_box, gc2_box, ...; kernels3d.cuh), so the results are bit-identical to it.  A phase holds up
// to a few operations whose virtual blocks are concatenated and distributed over the CTAs grid-stride.
//
// Memory ordering between phases (writer side -> barrier -> reader side):
//   * every thread waits for the completion of its own TMA bulk stores (cp.async.bulk.wait_group 0, not .read);
//   * __syncthreads(); thread 0: red.release.gpu on the counter, then it spins with ld.acquire.gpu (which also drops
//     stale L1 lines of the SM); __syncthreads() releases the CTA (see mega_grid_barrier for why one fence suffices);
//   * data that is written during the kernel is never read through the non-coherent path (LDG = false variants).
// The barrier has the same time-out escape as k_barrier: a CTA that waits longer than `timeout_ns` sets sync->err,
// later barriers fall through, and the host reports AFMG_ERR_CUDA instead of hanging the device.
#pragma once
#include "kernels3d.cuh"

namespace afmg {

enum MegaKind : int {
  MK_NONE = 0,
  MK_RB,            // rb_prepare_face            s0 = first face, a0 = var
  MK_GSRB,          // gsrb2_vblock               a0 = colour, a1 = level
  MK_EC,            // gc_edges_corners           a0 = var
  MK_RESTRICT,      // resid3_box<MODE 1>         a0 = keep_res
  MK_RESID,         // resid3_box<MODE 0>         a0 = with max-norm
  MK_GC,            // gc_sides (+ edges/corners) a0 = var, a1 = corners
  MK_GC2,           // gc2_box                    a0 = corners, a1 = mode
  MK_CS_FUSED,      // cs_fused_body              a0 = with_gc, n = level-1 boxes
  MK_CS_GATHER,     // cs_gather_cell             n = level-1 boxes
  MK_CS_APPLY,      // cs_apply_cell              a0 = d, a1 = trans, a2 = scale, a3 = 0: v0 -> v1, 1: v1 -> v0
  MK_CS_SCATTER,    // cs_scatter_cell            a3 = source buffer (0: v0, 1: v1)
  MK_CORRECT,       // correct3_box               a0 = push
  MK_STORE_CORR,    // tmp = phi - tmp on boxes with children
  MK_COPY,          // whole records: var a0 <- var a1
  MK_RESTRICT_VAR,  // k_restrict_var             a0 = var, a1 = clear_phi
  MK_CLEAR_SCAL,    // scal[a0] = 0
  MK_COUNT
};

struct MegaOp {
  int kind, lvl;  // lvl: for labels only
  int s0, n;      // first item (slot / face / ...) and number of items
  int nvb;        // virtual blocks = ceil(n / items per block)
  int a0, a1, a2, a3;
  int pad;
};
struct MegaPhase {
  int op0, nops;  // ops[op0 .. op0 + nops)
  int nvb;        // sum of their virtual blocks
  int pad;
};
struct MegaSync {
  unsigned long long count;  // arrivals, never reset
  unsigned long long base;   // value of count when the running launch started (updated by CTA 0 at its end)
  unsigned long long err;    // != 0 after a barrier time-out
  unsigned long long pad;
};

template <int NC>
struct MegaCfg {
  static constexpr int THREADS = 256;
  static constexpr int MINB = 3;  // CTAs per SM the register budget is set for (85 registers)
  // items per virtual block
  static constexpr int GS_BPC = (NC == 16) ? 1 : 4;
  static constexpr int GS_KS = 2;
  static constexpr int RES_KS = 2;
  static constexpr int RES_TPB = RES_KS * NC * NC / 2;   // threads per box in resid3_box
  static constexpr int RES_PER = THREADS / RES_TPB;      // 4 (nc = 8) or 1 (nc = 16)
  static constexpr int RB_NT = NC * NC;
  static constexpr int RB_PER = THREADS / RB_NT;
  static constexpr int EC_NT = (NC == 16) ? 256 : 128;
  static constexpr int EC_PER = THREADS / EC_NT;
  static constexpr int W = NC / 2 + 2;
  static constexpr size_t smem_bytes(int coarse_cells) {
    size_t a = (size_t)GS_BPC * (Lay3<NC>::COL + Lay3<NC>::NI);
    const size_t b = (size_t)RES_PER * 2 * Lay3<NC>::COL;
    const size_t c = (size_t)Correct3Cfg<NC>::BPC * Correct3Cfg<NC>::SB;
    const size_t d = (size_t)2 * coarse_cells;
    a = a > b ? a : b;
    a = a > c ? a : c;
    a = a > d ? a : d;
    return a * sizeof(double);
  }
  static_assert(GS_BPC * GS_KS * NC * NC / 2 == THREADS, "half-sweep block shape");
  static_assert(RES_TPB % 32 == 0 && THREADS % RES_TPB == 0, "residual sub-groups are whole warps");
};

__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void red_release_gpu_inc(unsigned long long* p) {
  asm volatile("red.release.gpu.global.add.u64 [%0], 1;" ::"l"(p) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ unsigned long long ld_relaxed_gpu(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// All CTAs of the (co-resident) grid; `target` = value the arrival counter reaches when every CTA has arrived.
// Cost matters here: a cycle has ~100 of these.  One GPU-scope release per CTA (red.release.gpu = MEMBAR.GPU + REDG;
// measured: every additional GPU-scope fence -- __threadfence(), fence.proxy.async -- adds ~0.6 us to EVERY phase) and
// acquire polls (LDG.STRONG.GPU + CCTL.IVALL, no membar).  Why no proxy fence is needed around it: the async-proxy
// writes of this CTA (bulk stores) have been waited for to completion by their issuing threads before the
// __syncthreads() (cp.async.bulk.wait_group 0), so they are ordinary visible writes that the release publishes; and
// the async-proxy reads of the next phase (TMA loads) go to L2, where the release / acquire pair has made every
// generic write of the other CTAs visible (the L1 lines an SM could still hold are dropped by the acquire).
__device__ __forceinline__ void mega_grid_barrier(MegaSync* sync, unsigned long long target, unsigned long long timeout_ns) {
  __syncthreads();
  if (threadIdx.x == 0) {
    red_release_gpu_inc(&sync->count);
    unsigned long long t0 = 0;
    unsigned spins = 0;
    while (ld_acquire_gpu(&sync->count) < target) {
      if ((++spins & 255u) == 0) {
        if (ld_relaxed_gpu(&sync->err) != 0) break;  // an earlier barrier timed out: fall through
        const unsigned long long now = globaltimer_ns();
        if (t0 == 0) t0 = now;
        if (now - t0 > timeout_ns) {
          sync->err = 1;
          __threadfence();
          break;
        }
      }
    }
  }
  __syncthreads();
}

// The same phase boundary when the whole grid is ONE thread-block cluster (<= 16 CTAs: the levels of at most a few
// dozen boxes at the bottom of a cycle): the hardware cluster barrier with release / acquire semantics, executed by
// every thread -- no atomics, no polling, no __syncthreads pair around it.  Cluster scope covers the global-memory
// writes of the CTAs of the cluster; the bulk stores have been waited for by their issuing threads as above.
__device__ __forceinline__ void mega_cluster_barrier() {
  asm volatile(
      "barrier.cluster.arrive.release.aligned;\n"
      "barrier.cluster.wait.acquire.aligned;\n" ::
          : "memory");
}

template <int NC>
__device__ __forceinline__ void mega_run_op(const DevCtx& cx, const CoarseCtx& cs, const MegaOp& op, int v, double* smem,
                                            uint64_t* bar, uint32_t& par, unsigned long long* scal) {
  using L = Lay3<NC>;
  using M = MegaCfg<NC>;
  constexpr int BOX = L::BOX, COL = L::COL;
  const int tid = threadIdx.x;
  switch (op.kind) {
    case MK_GSRB:
      gsrb2_vblock<NC, M::GS_BPC, M::GS_KS>(cx, op.s0, op.n, op.a0, op.a1, v, smem, bar, par);
      break;
    case MK_RB: {
      const int g = tid / M::RB_NT, q = v * M::RB_PER + g;
      if (q < op.n) rb_prepare_face<NC>(cx, op.s0 + q, op.a0, tid % M::RB_NT, M::RB_NT);
    } break;
    case MK_EC: {
      const int g = tid / M::EC_NT, q = v * M::EC_PER + g;
      if (q < op.n) gc_edges_corners<NC>(cx, op.s0 + q, op.a0, tid % M::EC_NT, M::EC_NT);
    } break;
    case MK_RESTRICT:
    case MK_RESID: {
      const int q0 = v * M::RES_PER, nhere = min(M::RES_PER, op.n - q0);
      if (tid == 0) {
        mbar_expect_tx(bar, (uint32_t)(nhere * 2 * COL * 8));
        for (int b = 0; b < nhere; ++b)
          bulk_g2s(smem + (size_t)b * 2 * COL, cx.cc[V_PHI] + (size_t)(op.s0 + q0 + b) * BOX, 2 * COL * 8, bar);
      }
      const int g = tid / M::RES_TPB;
      if (g < nhere) {
        if (op.kind == MK_RESTRICT)
          resid3_box<NC, M::RES_KS, 1, false>(cx, op.s0 + q0 + g, tid % M::RES_TPB, smem + (size_t)g * 2 * COL, nullptr,
                                              op.a0, bar, par);
        else
          resid3_box<NC, M::RES_KS, 0, false>(cx, op.s0 + q0 + g, tid % M::RES_TPB, smem + (size_t)g * 2 * COL,
                                              op.a0 ? scal : nullptr, 0, bar, par);
      }
      par ^= 1u;
    } break;
    case MK_GC: {
      gc_sides<NC>(cx, op.s0 + v, op.a0);
      if (op.a1) {
        __syncthreads();
        gc_edges_corners<NC>(cx, op.s0 + v, op.a0);
      }
    } break;
    case MK_GC2:
      gc2_box<NC>(cx, op.s0 + v, op.a0, op.a1, smem, bar, par);
      break;
    case MK_CS_FUSED:
      cs_fused_body<NC>(cx, cs, op.n, op.a0, smem);
      break;
    case MK_CS_GATHER:
      cs_gather_cell<NC>(cx, cs, op.n, v * M::THREADS + tid);
      break;
    case MK_CS_APPLY:
      cs_apply_cell(cs, op.a3 ? cs.v1 : cs.v0, op.a3 ? cs.v0 : cs.v1, op.a0, op.a1, op.a2, v * M::THREADS + tid);
      break;
    case MK_CS_SCATTER:
      cs_scatter_cell<NC>(cx, cs, op.n, op.a3 ? cs.v1 : cs.v0, v * M::THREADS + tid);
      break;
    case MK_CORRECT:
    {
      constexpr int CB = Correct3Cfg<NC>::BPC;
      correct3_box<NC, false>(cx, op.s0 + v * CB, min(CB, op.n - v * CB), op.a0, smem, bar, par);
    }
      break;
    case MK_STORE_CORR: {
      const int slot = op.s0 + v;
      if (cx.child0[slot] >= 0) {
        const double* phi = cx.cc[V_PHI] + (size_t)slot * BOX;
        double* tmp = cx.cc[V_TMP] + (size_t)slot * BOX;
        for (int q = tid; q < BOX; q += M::THREADS) tmp[q] = phi[q] - tmp[q];
      }
    } break;
    case MK_COPY: {
      const double* src = cx.cc[op.a1] + (size_t)(op.s0 + v) * BOX;
      double* dst = cx.cc[op.a0] + (size_t)(op.s0 + v) * BOX;
      for (int q = tid; q < BOX; q += M::THREADS) dst[q] = src[q];
    } break;
    case MK_RESTRICT_VAR: {
      // k_restrict_var (init_phi_rhs, m_af_multigrid.f90:779-799)
      constexpr int H = L::H;
      const int slot = op.s0 + v;
      const double* src = cx.cc[op.a0] + (size_t)slot * BOX;
      const int p = cx.parent[slot], cof = cx.coff[slot];
      double* dst = cx.at<BOX>(op.a0, p);
      const int ox = (cof & 1) * H, oy = ((cof >> 1) & 1) * H, oz = ((cof >> 2) & 1) * H;
      if (op.a1) {
        double* phi = cx.cc[V_PHI] + (size_t)slot * BOX;
        for (int q = tid; q < BOX; q += M::THREADS) phi[q] = 0.0;
      }
      for (int n = tid; n < H * H * H; n += M::THREADS) {
        const int ic = n % H + 1, jc = (n / H) % H + 1, kc = n / (H * H) + 1;
        double s = 0.0;
#pragma unroll
        for (int dk = 0; dk < 2; ++dk)
#pragma unroll
          for (int dj = 0; dj < 2; ++dj)
#pragma unroll
            for (int di = 0; di < 2; ++di) s = s + src[L::interior(2 * ic - 1 + di, 2 * jc - 1 + dj, 2 * kc - 1 + dk)];
        dst[L::interior(ox + ic, oy + jc, oz + kc)] = 0.125 * s;
      }
    } break;
    case MK_CLEAR_SCAL:
      if (tid == 0) scal[op.a0] = 0ull;
      break;
    default: break;
  }
}

constexpr int MEGA_MAX_OPS = 4;  // operations per phase (the host never records more)

template <int NC>
__global__ void __launch_bounds__(MegaCfg<NC>::THREADS, MegaCfg<NC>::MINB)
    k_mega(DevCtx cx, CoarseCtx cs, const MegaPhase* __restrict__ phases, const MegaOp* __restrict__ ops, int nphase,
           MegaSync* sync, unsigned long long* scal, unsigned long long timeout_ns, unsigned long long* stamps,
           int cluster) {
  extern __shared__ __align__(128) double smem[];
  __shared__ uint64_t bar;
  __shared__ unsigned long long s_base;
  // phase descriptors are double-buffered in shared memory: the one of phase p + 1 is fetched while phase p runs,
  // so that no dependent global loads sit between a barrier and the first TMA load of the next phase
  __shared__ MegaPhase s_ph[2];
  __shared__ MegaOp s_op[2][MEGA_MAX_OPS];
  const int tid = threadIdx.x;
  auto fetch = [&](int p) {  // threads 32 .. 32 + MEGA_MAX_OPS of the CTA
    const int q = tid - 32;
    if (q >= 0 && q < MEGA_MAX_OPS && p < nphase) {
      const MegaPhase ph = phases[p];
      if (q == 0) s_ph[p & 1] = ph;
      if (q < ph.nops) s_op[p & 1][q] = ops[ph.op0 + q];
    }
  };
  if (tid == 0) {
    mbar_init(&bar, 1);
    s_base = cluster ? 0ull : ld_acquire_gpu(&sync->base);
    if (stamps && blockIdx.x == 0) stamps[0] = globaltimer_ns();
  }
  fetch(0);
  __syncthreads();
  uint32_t par = 0;
  unsigned long long target = s_base;
  for (int p = 0; p < nphase; ++p) {
    fetch(p + 1);  // visible after the barrier's __syncthreads
    const MegaPhase ph = s_ph[p & 1];
    for (int vb = blockIdx.x; vb < ph.nvb; vb += gridDim.x) {
      int o = 0, v = vb;
      while (v >= s_op[p & 1][o].nvb) {
        v -= s_op[p & 1][o].nvb;
        ++o;
      }
      const MegaOp op = s_op[p & 1][o];
      mega_run_op<NC>(cx, cs, op, v, smem, &bar, par, scal);
      // the block's bulk stores are complete and its shared memory is free for the next block
      bulk_commit();
      bulk_wait_all();
      fence_async_smem();
      __syncthreads();
    }
    target += gridDim.x;
    if (cluster) mega_cluster_barrier();
    else mega_grid_barrier(sync, target, timeout_ns);
    if (stamps && blockIdx.x == 0 && tid == 0) stamps[p + 1] = globaltimer_ns();
  }
  // every CTA has read `base` before it arrived at the first barrier, and the last barrier above is complete
  if (!cluster && blockIdx.x == 0 && tid == 0 && nphase > 0) sync->base = target;
}

}  // namespace afmg

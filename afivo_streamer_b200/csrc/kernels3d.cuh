// sm_100a kernels of the 3D FAS multigrid path.  Everything is fp64 and bandwidth bound; no
// tensor cores.  Compiled with -fmad=false so that every expression rounds exactly like the
// reference's gfortran -O2 build (no FMA contraction on x86-64): results are bit-identical to the
// CPU oracle except downstream of the coarse-grid solve.
//
// Conventions: "slot" = position of a box in the device arrays (level-major, Morton order inside a
// level), box record layout = layout.cuh.  Face ghost cells of a box whose neighbour exists are
// written by the NEIGHBOUR's half-sweep kernel ("push"); ghost cells on physical / refinement
// boundaries are computed by the box's own CTA from a rule row  ghost = (c0*B + c1*x1) + c2*x2,
// where B is the boundary value (bc_val) or the interpolated coarse value (k_rb_prepare).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

#include "layout.cuh"

namespace afmg {

enum { V_PHI = 0, V_RHS = 1, V_TMP = 2, V_EPS = 3, V_FLD = 4 };
#define AFMG_MAX_RANKS 8

struct DevCtx {
  double* cc[5];         // per variable: nslots * BOX doubles (V_EPS unused here; V_FLD: field norm, may be null)
  const int* nbr;        // [nslots*6]  >= 0: neighbour slot; -1: own-ghost rule row in aux
  const int* aux;        // [nslots*6]  rule row
  const int* nmat;       // [nslots*27] >= 0 slot, -1 physical boundary, -2 no box (coarser there)
  const int* parent;     // [nslots]    parent slot (-1 on level 1)
  const int* child0;     // [nslots]    slot of first child, -1 for leaves
  const int* coff;       // [nslots]    child index 0..7 inside the parent (bit d: upper half in dim d)
  const int* lvl;        // [nslots]
  const double* coef;    // [(L+1)*8]   per level: c1..c7 of the constant 7-point stencil, 1/c1
  const double* rule_c;  // [nrules*3]  c0, c1, c2
  double* rule_B;        // [nrules*NC2] boundary values / interpolated coarse values, index (a-1)+(b-1)*nc
  const int* rb_slot;    // [nrb] fine slot of refinement-boundary face r (rule row = rb_row0 + r)
  const int* rb_face;    // [nrb] face 0..5
  int rb_row0;           // first rule row that is a refinement-boundary face
  const double* pcoef;   // [8] constant prolongation coefficients
  int pshape;            // 8 = stencil_prolong_248 (linear), 4 = stencil_prolong_234 (sparse)
  // ---- per-box stencils shipped by the host (afmg_set_stencils); all null / 0 when every box uses the
  // implicit constant Laplacian and the default prolongation
  const unsigned char* opk;  // [nslots] operator: 0 implicit constant (coef), 1 explicit constant, 2 variable
  const long long* opoff;    // [nslots] offset in stv: 7 doubles (kind 1) or 7 planes [m][colour][NI] (kind 2)
  const long long* foff;     // [nslots] offset of f [colour][NI] (bc_correction = f * lsf_value), -1: none
  const unsigned char* pk;   // [nslots] prolongation of a child box: 0 default, 1 constant p248 (8), 2 constant
                             //          p234 (4), 3 variable p234 (4 planes [m][colour][NI])
  const long long* poff;     // [nslots] offset in stv
  const double* stv;         // coefficient pool
  const double* lsf_value_p;  // mg%lsf_boundary_value, in device memory (CommBlock::lsf_value): it changes every
                              // time step (the applied voltage), and cached graphs must stay valid
  __device__ __forceinline__ double lsf_value() const { return *lsf_value_p; }
  // mg%lsf_boundary_function as data (afmg_set_lsf_boundary_values): per-cell boundary values [colour][NI] of the
  // boxes listed there, -1 / null: the scalar lsf_value
  const long long* bvoff;    // [nslots] offset in bv, or null
  const double* bv;
  const unsigned char* rule_flag;  // [nrules] 1: refinement-boundary face of a variable-eps box (mg_sides_rb_extrap)
  // ---- multi-GPU (one process per GPU): every rank allocates the same slot-indexed arrays and
  // maps its peers' arrays through CUDA IPC, so a box is addressed as (owner rank, slot) on every
  // GPU.  nranks == 1: ccr is unused.
  int nranks, me;
  const unsigned char* owner;  // [nslots] rank that owns (computes) the box
  double* ccr[AFMG_MAX_RANKS][5];  // phi / rhs / tmp (/ field norm) base pointers of every rank (own entry == cc)
  double* bsum[AFMG_MAX_RANKS];    // per-box sums (k_box_sums) of every rank

  // base of the record of box `slot` for variable `var` in the memory of the rank that owns it
  template <int BOX>
  __device__ __forceinline__ double* at(int var, int slot) const {
    if (nranks == 1) return cc[var] + (size_t)slot * BOX;
    return ccr[owner[slot]][var] + (size_t)slot * BOX;
  }
  __device__ __forceinline__ bool remote(int slot) const { return nranks > 1 && owner[slot] != me; }
  // per-cell level-set boundary values of a box (mg_lsf_boundary_value, m_coarse_solver.f90:493-510), or null
  __device__ __forceinline__ const double* bv_of(int slot) const {
    return (bvoff && bvoff[slot] >= 0) ? bv + bvoff[slot] : nullptr;
  }
};

// ---------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + 1D TMA bulk copy (cp.async.bulk -> SASS UBLKCP)
// ---------------------------------------------------------------------------------------------
// programmatic dependent launch: wait for the preceding grid (and its memory) before touching data
// (AFMG_PDL=1).  launch_dependents first: the NEXT kernel of the stream / graph may be scheduled as soon as every CTA of
// this one has started, so its launch latency overlaps with this kernel's execution; its own griddepcontrol.wait still
// blocks until this grid has completed and flushed.  Both instructions are no-ops without the launch attribute.
__device__ __forceinline__ void pdl_wait() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

__device__ __forceinline__ void atomic_max_nonneg(unsigned long long* addr, double v) {
  atomicMax(addr, (unsigned long long)__double_as_longlong(v));
}

// ---------------------------------------------------------------------------------------------
// Multi-GPU synchronisation over peer memory (NVLink).  Every rank owns one CommBlock and maps the
// blocks of its peers (CUDA IPC).  k_barrier is enqueued between dependent kernels in place of the
// stream order a single GPU gives for free: it publishes this rank's barrier count into every peer's
// block (system-scope release after a system fence, so the previous kernels' peer stores are visible
// first) and waits until every peer has published the same count.  A wait that exceeds the time-out
// sets `err` (reported by the host as AFMG_ERR_COMM) and later barriers fall through, so a lost peer
// can never hang the GPU.
// ---------------------------------------------------------------------------------------------
struct CommBlock {
  unsigned long long flags[AFMG_MAX_RANKS];  // flags[r] = number of barriers rank r has entered
  unsigned long long epoch;                  // number of barriers this rank has entered
  unsigned long long err;                    // != 0 after a barrier time-out
  unsigned long long scal[8];                // reduction scratch: [0] residual max, [1] generic max,
                                             // [2] mean, [4],[5] = [0],[1] combined over all ranks
  double lsf_value;                          // mg%lsf_boundary_value (DevCtx::lsf_value_p points here)
};
struct CommPeers {
  CommBlock* p[AFMG_MAX_RANKS];
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// lean (the default): the release store IS the system-scope fence (it is cumulative over the writes of the preceding
// kernels, which happen before this one in stream order), and the acquire polls order the peers' writes before the
// kernels that follow; the explicit fences of the first version (lean = 0) cost several microseconds per barrier.
__global__ void k_barrier(CommBlock* mine, CommPeers peers, int nranks, int me, unsigned long long timeout_ns, int lean) {
  pdl_wait();
  __shared__ unsigned long long ep;
  const int t = threadIdx.x;
  if (t == 0) {
    ep = mine->epoch + 1;
    mine->epoch = ep;
  }
  __syncthreads();
  const unsigned long long e = ep;
  if (t < nranks && t != me) {
    if (!lean) __threadfence_system();
    st_release_sys(&peers.p[t]->flags[me], e);
    if (lean ? (mine->err == 0) : (ld_acquire_sys(&mine->err) == 0)) {
      unsigned long long t0 = 0;
      unsigned spins = 0;
      while (ld_acquire_sys(&mine->flags[t]) < e) {
        if (lean && (++spins & 63u)) continue;
        const unsigned long long now = globaltimer_ns();
        if (t0 == 0) t0 = now;
        if (now - t0 > timeout_ns) {
          mine->err = 1;
          break;
        }
      }
    }
  }
  if (!lean) __threadfence_system();
}

// scal[4 + idx] = max over ranks of scal[idx] (after a barrier); non-negative doubles order like uint64
__global__ void k_allmax(CommBlock* mine, CommPeers peers, int nranks, int idx) {
  pdl_wait();
  unsigned long long m = 0;
  for (int r = 0; r < nranks; ++r) {
    const unsigned long long v = ld_acquire_sys(&peers.p[r]->scal[idx]);
    m = v > m ? v : m;
  }
  mine->scal[4 + idx] = m;
}

// neighbour offset tables
__device__ __forceinline__ int nmat_index(int dx, int dy, int dz) { return (dx + 1) + 3 * (dy + 1) + 9 * (dz + 1); }

// ---------------------------------------------------------------------------------------------
// k_gsrb2: same operation as k_gsrb, restructured for issue rate and bytes in flight:
//   - TMA bulk copies bring in the opposite colour block of phi (interior + faces) AND the rhs of
//     colour C; the rhs buffer is overwritten in place with the new phi values and leaves by one TMA
//     bulk store (smem -> global), so the inner loop is LDS / DADD / DMUL / STS only
//   - KS threads share one (m, j) column (k split in KS ranges) for more warps per box
//   - ghost pushes / boundary rules run after the sweep as cooperative, coalesced copies out of smem;
//     the two z faces are contiguous in smem and leave as bulk copies too
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Ghost-cell epilogue shared by the kernels that hold freshly computed interior values of a box in
// shared memory (I0 / I1 = interior colour blocks 0 / 1, NI doubles each).  For the colours in `mask`
// the boundary layers are pushed into the opposite ghost faces of the same-level neighbours
// (copy_from_nb, m_af_ghostcell.f90:654-669); on physical / refinement faces all nc^2 ghost cells of
// the box are recomputed from their rule (bc_to_gc :173-279; mg_sides_rb m_af_multigrid.f90:383-459).
// t = index among the TPB threads working on this box.  The z faces are contiguous in shared memory
// and leave as TMA bulk copies issued by t == 0; the CALLER commits and waits for the bulk group.
// Face metadata of one box (neighbour slots, rule rows, rule coefficients), optionally prefetched into shared
// memory while the box data is still in flight: on small levels the three dependent global loads of the epilogue
// (nbr -> aux -> rule_c) are otherwise a quarter of a half-sweep's duration.
struct FaceMeta {
  int nb[6];
  int row[6];
  double rc[6][3];
};
__device__ __forceinline__ void prefetch_face_meta(const DevCtx& cx, int slot, FaceMeta* fm, int f) {
  const int nb = cx.nbr[slot * 6 + f], row = cx.aux[slot * 6 + f];
  fm->nb[f] = nb;
  fm->row[f] = row;
  if (nb < 0) {
    fm->rc[f][0] = cx.rule_c[3 * row];
    fm->rc[f][1] = cx.rule_c[3 * row + 1];
    fm->rc[f][2] = cx.rule_c[3 * row + 2];
  }
}

template <int NC, int TPB, bool SRC_SMEM = true>
__device__ __forceinline__ void epilogue_faces(const DevCtx& cx, int slot, const double* I0, const double* I1, int mask,
                                               int t, const FaceMeta* fm = nullptr) {
  using L = Lay3<NC>;
  constexpr int H = L::H, NI = L::NI, NF = L::NF, COL = L::COL, BOX = L::BOX;
  double* const gbox = cx.cc[V_PHI] + (size_t)slot * BOX;
  const int* nbp = fm ? fm->nb : cx.nbr + slot * 6;
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    if (!((mask >> c) & 1)) continue;
    const double* Ic = c ? I1 : I0;
#pragma unroll
    for (int f = 4; f < 6; ++f) {
      const int nb = nbp[f];
      if (nb < 0) continue;
      double* dst = cx.at<BOX>(V_PHI, nb) + c * COL + NI + (f ^ 1) * NF;
      const double* srcz = (f == 4) ? Ic : Ic + (NC - 1) * NC * H;
      if (SRC_SMEM && !cx.remote(nb)) {
        if (t == 0) bulk_s2g(dst, srcz, NF * 8);
      } else {  // peer GPU (plain coalesced stores over NVLink) or source not in shared memory
        for (int fi = t; fi < NF; fi += TPB) dst[fi] = srcz[fi];
      }
    }
#pragma unroll
    for (int f = 0; f < 4; ++f) {
      const int nb = nbp[f];
      if (nb < 0) continue;
      double* dst = cx.at<BOX>(V_PHI, nb) + c * COL + NI + (f ^ 1) * NF;
      for (int fi = t; fi < NF; fi += TPB) {
        const int k = fi / H + 1, ah = fi % H;
        int src;
        if (f >= 2) {
          src = L::iidx(ah, (f & 1) ? NC : 1, k);
        } else {
          // cells i = 1 (f = 0) or i = NC (f = 1) of colour c: j has parity (c + i + k) & 1
          const int i = (f & 1) ? NC : 1;
          const int j = 2 * ah + 2 - ((c + i + k) & 1);
          src = L::iidx((i - 1) >> 1, j, k);
        }
        dst[fi] = Ic[src];
      }
    }
  }
#pragma unroll
  for (int f = 0; f < 6; ++f) {
    if (nbp[f] >= 0) continue;
    const int row = fm ? fm->row[f] : cx.aux[slot * 6 + f];
    const double* rc = fm ? fm->rc[f] : cx.rule_c + 3 * row;
    const double r0 = rc[0], r1 = rc[1], r2 = rc[2];
    const double* B = cx.rule_B + (size_t)row * L::NC2;
    const int d = f >> 1, hi = f & 1;
    // mg_sides_rb_extrap (m_af_multigrid.f90:468-621): the second point is the DIAGONAL layer-2 cell
    const bool diag = cx.rule_flag && cx.rule_flag[row];
    for (int n = t; n < L::NC2; n += TPB) {
      const int a = n % NC + 1, bb = n / NC + 1;
      const int l1 = hi ? NC : 1, l2 = hi ? NC - 1 : 2;
      const int a2 = diag ? a - 1 + 2 * (a & 1) : a, b2 = diag ? bb - 1 + 2 * (bb & 1) : bb;
      // (x, y, z) of the layer-1 / layer-2 cells: dim d takes the layer, the others (a, bb) in order
      const int px1 = (d == 0) ? l1 : a, py1 = (d == 0) ? a : (d == 1 ? l1 : bb), pz1 = (d == 2) ? l1 : bb;
      const int px2 = (d == 0) ? l2 : a2, py2 = (d == 0) ? a2 : (d == 1 ? l2 : b2), pz2 = (d == 2) ? l2 : b2;
      const int col1 = (px1 + py1 + pz1) & 1;  // colour of the layer-1 cell; ghost has colour 1 - col1
      const int i1 = L::iidx((px1 - 1) >> 1, py1, pz1), i2 = L::iidx((px2 - 1) >> 1, py2, pz2);
      const double x1 = (col1 ? I1 : I0)[i1];
      const double x2 = (col1 ? I0 : I1)[i2];
      gbox[(1 - col1) * COL + L::fidx(f, a, bb)] = (r0 * B[n] + r1 * x1) + r2 * x2;
    }
  }
}

// One virtual block (BPC boxes starting at slot0 + vb * BPC) of a half-sweep.  `bar` is an initialised mbarrier of
// the CTA (count 1) and `par` the parity of its next completion (flipped here), so that a persistent kernel can
// call this repeatedly (k_mega); k_gsrb2 calls it once with vb = blockIdx.x.
template <int NC, int BPC, int KS>
__device__ __forceinline__ void gsrb2_vblock(const DevCtx& cx, int slot0, int nbox, int C, int lvl, int vb, double* smem,
                                             uint64_t* bar, uint32_t& par) {
  using L = Lay3<NC>;
  constexpr int H = L::H, NI = L::NI, NF = L::NF, COL = L::COL, BOX = L::BOX;
  constexpr int TPB = H * NC * KS;  // threads per box
  constexpr int KL = NC / KS;       // k-steps per thread
  constexpr int SBOX = COL + NI;    // smem doubles per box: phi block of the other colour, rhs/out
  const int tid = threadIdx.x;
  const int box0 = vb * BPC;
  const int nhere = min(BPC, nbox - box0);
  double* const phi = cx.cc[V_PHI];
  // boxes with an explicit (variable-coefficient / level-set) stencil are left to k_gsrb_gen
  if (tid == 0) {
    int nload = 0;
    for (int b = 0; b < nhere; ++b) nload += (cx.opk && cx.opk[slot0 + box0 + b]) ? 0 : 1;
    mbar_expect_tx(bar, (uint32_t)(nload * SBOX * 8));
    for (int b = 0; b < nhere; ++b) {
      if (cx.opk && cx.opk[slot0 + box0 + b]) continue;
      const size_t base = (size_t)(slot0 + box0 + b) * BOX;
      bulk_g2s(smem + b * SBOX, phi + base + (1 - C) * COL, COL * 8, bar);
      bulk_g2s(smem + b * SBOX + COL, cx.cc[V_RHS] + base + C * COL, NI * 8, bar);
    }
  }
  const int b = tid / TPB, t = tid % TPB;
  const bool active = b < nhere && !(cx.opk && cx.opk[slot0 + box0 + (b < nhere ? b : 0)]);
  const int slot = slot0 + box0 + (b < nhere ? b : 0);
  const double* const S = smem + b * SBOX;
  double* const R = smem + b * SBOX + COL;
  const double* cf = cx.coef + 8 * lvl;
  const double c2 = cf[1], c3 = cf[2], c4 = cf[3], c5 = cf[4], c6 = cf[5], c7 = cf[6], inv = cf[7];
  __shared__ FaceMeta fmeta[BPC];
  if (active && t < 6) prefetch_face_meta(cx, slot, &fmeta[b], t);  // visible after the __syncthreads below
  mbar_wait(bar, par);
  par ^= 1u;

  if (active) {
    const int m = t % H, j = (t / H) % NC + 1, ks = t / (H * NC);
    const int k0 = ks * KL + 1;
    double s_km1 = (k0 == 1) ? S[NI + 4 * NF + (j - 1) * H + m] : S[L::iidx(m, j, k0 - 1)];
    double s_k = S[L::iidx(m, j, k0)];
#pragma unroll
    for (int kk = 0; kk < KL; ++kk) {
      const int k = k0 + kk;
      const int idx = L::iidx(m, j, k);
      const int pi = (C + j + k) & 1;  // 1: i = 2m+1, 0: i = 2m+2
      const double s_kp1 = (k < NC) ? S[idx + NC * H] : S[NI + 5 * NF + (j - 1) * H + m];
      const double ym = (j > 1) ? S[idx - H] : S[NI + 2 * NF + (k - 1) * H + m];
      const double yp = (j < NC) ? S[idx + H] : S[NI + 3 * NF + (k - 1) * H + m];
      const int fx = (k - 1) * H + ((j - 1) >> 1);
      double xm, xp;
      if (pi) {
        xm = (m > 0) ? S[idx - 1] : S[NI + 0 * NF + fx];
        xp = s_k;
      } else {
        xm = s_k;
        xp = (m < H - 1) ? S[idx + 1] : S[NI + 1 * NF + fx];
      }
      double acc = R[idx];
      acc = acc - c2 * xm;
      acc = acc - c3 * xp;
      acc = acc - c4 * ym;
      acc = acc - c5 * yp;
      acc = acc - c6 * s_km1;
      acc = acc - c7 * s_kp1;
      R[idx] = acc * inv;
      s_km1 = s_k;
      s_k = s_kp1;
    }
  }
  fence_async_smem();
  __syncthreads();
  if (active) {
    // ---- epilogue: new colour-C values are in R (layout of an interior colour block)
    if (t == 0) bulk_s2g(phi + (size_t)slot * BOX + C * COL, R, NI * 8);
    epilogue_faces<NC, TPB>(cx, slot, C ? S : R, C ? R : S, 1 << C, t, &fmeta[b]);
    if (t == 0) bulk_commit();
    if (t == 0) bulk_wait_read0();
  }
}

// k_gsrb2s: the same half-sweep for SMALL launches (levels of a few hundred boxes, where a kernel is a chain of
// dependent latencies and not a bandwidth problem; a dependent graph node costs ~1 us on this part,
// profiles/r02q_graphnode.txt, so the rest of a ~6 us node is the kernel's own critical path).  Differences to k_gsrb2,
// none of them in the arithmetic: the box data arrives by plain 16-byte loads issued by all threads at once instead
// of TMA bulk copies behind an mbarrier (lower latency for a few KB), the boundary values of the rule faces (rule_B)
// are fetched into registers while the box data is in flight instead of after the sweep, and results leave by plain
// stores (no bulk-group wait at the end).  Used when the launch has at most SMALL_CTAS blocks.
template <int NC, int BPC, int KS>
__global__ void __launch_bounds__(BPC* KS* NC* NC / 2, 2) k_gsrb2s(DevCtx cx, int slot0, int nbox, int C, int lvl) {
  pdl_wait();
  using L = Lay3<NC>;
  constexpr int H = L::H, NI = L::NI, NF = L::NF, COL = L::COL, BOX = L::BOX;
  constexpr int TPB = H * NC * KS, KL = NC / KS, SBOX = COL + NI;
  static_assert(L::NC2 == TPB, "one rule-face cell per thread");
  extern __shared__ __align__(128) double smem[];
  const int tid = threadIdx.x;
  const int box0 = blockIdx.x * BPC;
  const int nhere = min(BPC, nbox - box0);
  const int b = tid / TPB, t = tid % TPB;
  const int slot = slot0 + box0 + (b < nhere ? b : 0);
  const bool active = b < nhere && !(cx.opk && cx.opk[slot]);
  double* const S = smem + b * SBOX;
  double* const R = S + COL;
  double* const gphi = cx.cc[V_PHI] + (size_t)slot * BOX;
  // ---- everything this block needs from global memory, issued back to back
  if (active) {
    const double2* src = reinterpret_cast<const double2*>(gphi + (1 - C) * COL);
    double2* dst = reinterpret_cast<double2*>(S);
#pragma unroll
    for (int q = t; q < COL / 2; q += TPB) dst[q] = src[q];
    const double2* rsrc = reinterpret_cast<const double2*>(cx.cc[V_RHS] + (size_t)slot * BOX + C * COL);
    double2* rdst = reinterpret_cast<double2*>(R);
#pragma unroll
    for (int q = t; q < NI / 2; q += TPB) rdst[q] = rsrc[q];
  }
  const double* cf = cx.coef + 8 * lvl;
  const double c2 = cf[1], c3 = cf[2], c4 = cf[3], c5 = cf[4], c6 = cf[5], c7 = cf[6], inv = cf[7];
  int nbf[6], rowf[6];
  double Bpre[6], rc0[6], rc1[6], rc2[6];
#pragma unroll
  for (int f = 0; f < 6; ++f) {
    nbf[f] = active ? cx.nbr[slot * 6 + f] : 0;
    rowf[f] = active ? cx.aux[slot * 6 + f] : 0;
  }
#pragma unroll
  for (int f = 0; f < 6; ++f) {
    Bpre[f] = rc0[f] = rc1[f] = rc2[f] = 0.0;
    if (active && nbf[f] < 0) {
      Bpre[f] = cx.rule_B[(size_t)rowf[f] * L::NC2 + t];
      rc0[f] = cx.rule_c[3 * rowf[f]];
      rc1[f] = cx.rule_c[3 * rowf[f] + 1];
      rc2[f] = cx.rule_c[3 * rowf[f] + 2];
    }
  }
  __syncthreads();
  if (active) {
    const int m = t % H, j = (t / H) % NC + 1, ks = t / (H * NC);
    const int k0 = ks * KL + 1;
    double s_km1 = (k0 == 1) ? S[NI + 4 * NF + (j - 1) * H + m] : S[L::iidx(m, j, k0 - 1)];
    double s_k = S[L::iidx(m, j, k0)];
#pragma unroll
    for (int kk = 0; kk < KL; ++kk) {
      const int k = k0 + kk;
      const int idx = L::iidx(m, j, k);
      const int pi = (C + j + k) & 1;
      const double s_kp1 = (k < NC) ? S[idx + NC * H] : S[NI + 5 * NF + (j - 1) * H + m];
      const double ym = (j > 1) ? S[idx - H] : S[NI + 2 * NF + (k - 1) * H + m];
      const double yp = (j < NC) ? S[idx + H] : S[NI + 3 * NF + (k - 1) * H + m];
      const int fx = (k - 1) * H + ((j - 1) >> 1);
      double xm, xp;
      if (pi) {
        xm = (m > 0) ? S[idx - 1] : S[NI + 0 * NF + fx];
        xp = s_k;
      } else {
        xm = s_k;
        xp = (m < H - 1) ? S[idx + 1] : S[NI + 1 * NF + fx];
      }
      double acc = R[idx];
      acc = acc - c2 * xm;
      acc = acc - c3 * xp;
      acc = acc - c4 * ym;
      acc = acc - c5 * yp;
      acc = acc - c6 * s_km1;
      acc = acc - c7 * s_kp1;
      R[idx] = acc * inv;
      s_km1 = s_k;
      s_k = s_kp1;
    }
  }
  __syncthreads();
  if (!active) return;
  // ---- new colour-C values out: own interior, neighbours' ghost faces, own rule faces
  {
    const double2* src = reinterpret_cast<const double2*>(R);
    double2* dst = reinterpret_cast<double2*>(gphi + C * COL);
#pragma unroll
    for (int q = t; q < NI / 2; q += TPB) dst[q] = src[q];
  }
  const double* I0 = C ? S : R;
  const double* I1 = C ? R : S;
  const double* Ic = R;
#pragma unroll
  for (int f = 0; f < 6; ++f) {
    const int nb = nbf[f];
    if (nb < 0) continue;
    double* dstf = cx.at<BOX>(V_PHI, nb) + C * COL + NI + (f ^ 1) * NF;
    for (int fi = t; fi < NF; fi += TPB) {
      int src;
      if (f >= 4) {
        src = (f == 4 ? 0 : (NC - 1) * NC * H) + fi;
      } else {
        const int k = fi / H + 1, ah = fi % H;
        if (f >= 2) {
          src = L::iidx(ah, (f & 1) ? NC : 1, k);
        } else {
          const int i = (f & 1) ? NC : 1;
          const int j = 2 * ah + 2 - ((C + i + k) & 1);
          src = L::iidx((i - 1) >> 1, j, k);
        }
      }
      dstf[fi] = Ic[src];
    }
  }
#pragma unroll
  for (int f = 0; f < 6; ++f) {
    if (nbf[f] >= 0) continue;
    const int d = f >> 1, hi = f & 1;
    const bool diag = cx.rule_flag && cx.rule_flag[rowf[f]];
    const int n = t;
    const int a = n % NC + 1, bb = n / NC + 1;
    const int l1 = hi ? NC : 1, l2 = hi ? NC - 1 : 2;
    const int a2 = diag ? a - 1 + 2 * (a & 1) : a, b2 = diag ? bb - 1 + 2 * (bb & 1) : bb;
    const int px1 = (d == 0) ? l1 : a, py1 = (d == 0) ? a : (d == 1 ? l1 : bb), pz1 = (d == 2) ? l1 : bb;
    const int px2 = (d == 0) ? l2 : a2, py2 = (d == 0) ? a2 : (d == 1 ? l2 : b2), pz2 = (d == 2) ? l2 : b2;
    const int col1 = (px1 + py1 + pz1) & 1;
    const int i1 = L::iidx((px1 - 1) >> 1, py1, pz1), i2 = L::iidx((px2 - 1) >> 1, py2, pz2);
    const double x1 = (col1 ? I1 : I0)[i1];
    const double x2 = (col1 ? I0 : I1)[i2];
    gphi[(1 - col1) * COL + L::fidx(f, a, bb)] = (rc0[f] * Bpre[f] + rc1[f] * x1) + rc2[f] * x2;
  }
}

template <int NC, int BPC, int KS, int MINB>
__global__ void __launch_bounds__(BPC* KS* NC* NC / 2, MINB) k_gsrb2(DevCtx cx, int slot0, int nbox, int C, int lvl) {
  pdl_wait();
  extern __shared__ __align__(128) double smem[];
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) mbar_init(&bar, 1);
  __syncthreads();
  uint32_t par = 0;
  gsrb2_vblock<NC, BPC, KS>(cx, slot0, nbox, C, lvl, blockIdx.x, smem, &bar, par);
}

// ---------------------------------------------------------------------------------------------
// Generic cell access on a box record in global memory (non-hot kernels)
// ---------------------------------------------------------------------------------------------
template <int NC>
__device__ __forceinline__ double ldcell(const double* box, int i, int j, int k) {
  return box[Lay3<NC>::cell(i, j, k)];
}

// L phi at interior cell (i,j,k): stencil_apply_357 (m_af_stencil.f90:462-487), left to right
template <int NC>
__device__ __forceinline__ double apply357(const double* box, const double* cf, double c1, int i, int j, int k) {
  double acc = c1 * ldcell<NC>(box, i, j, k);
  acc = acc + cf[1] * ldcell<NC>(box, i - 1, j, k);
  acc = acc + cf[2] * ldcell<NC>(box, i + 1, j, k);
  acc = acc + cf[3] * ldcell<NC>(box, i, j - 1, k);
  acc = acc + cf[4] * ldcell<NC>(box, i, j + 1, k);
  acc = acc + cf[5] * ldcell<NC>(box, i, j, k - 1);
  acc = acc + cf[6] * ldcell<NC>(box, i, j, k + 1);
  return acc;
}

// ---------------------------------------------------------------------------------------------
// Shared-memory versions of the operator kernels.  S points at both colour blocks of one box
// (2*COL doubles, as laid out in global memory) staged by one TMA bulk copy.
// ---------------------------------------------------------------------------------------------
// L phi at interior cell (i,j,k) from the staged box: stencil_apply_357 (m_af_stencil.f90:462-487)
template <int NC>
__device__ __forceinline__ double apply357_smem(const double* S, const double* cf, double c1, int i, int j, int k) {
  using L = Lay3<NC>;
  constexpr int H = L::H, NI = L::NI, NF = L::NF, COL = L::COL;
  const int c = (i + j + k) & 1, m = (i - 1) >> 1;
  const double* Sn = S + (1 - c) * COL;
  const int idx = L::iidx(m, j, k);
  const int fx = (k - 1) * H + ((j - 1) >> 1);
  double xm, xp;
  if (i & 1) {
    xm = (m > 0) ? Sn[idx - 1] : Sn[NI + 0 * NF + fx];
    xp = Sn[idx];
  } else {
    xm = Sn[idx];
    xp = (m < H - 1) ? Sn[idx + 1] : Sn[NI + 1 * NF + fx];
  }
  const double ym = (j > 1) ? Sn[idx - H] : Sn[NI + 2 * NF + (k - 1) * H + m];
  const double yp = (j < NC) ? Sn[idx + H] : Sn[NI + 3 * NF + (k - 1) * H + m];
  const double zm = (k > 1) ? Sn[idx - NC * H] : Sn[NI + 4 * NF + (j - 1) * H + m];
  const double zp = (k < NC) ? Sn[idx + NC * H] : Sn[NI + 5 * NF + (j - 1) * H + m];
  double acc = c1 * S[c * COL + idx];
  acc = acc + cf[1] * xm;
  acc = acc + cf[2] * xp;
  acc = acc + cf[3] * ym;
  acc = acc + cf[4] * yp;
  acc = acc + cf[5] * zm;
  acc = acc + cf[6] * zp;
  return acc;
}

// k_resid3: residual of one box per CTA with the x-pair formulation: thread (m, j, ks) owns the two
// cells i = 2m+1 (A) and i = 2m+2 (B) of row (j, k) -- one of each colour, same index in their colour
// blocks -- and walks its k-range, so that z neighbours chain through registers (8 LDS per 2 cells).
//   MODE 0: tmp = rhs - L(phi)  (+ max |tmp| over leaves)          residual_box / af_tree_maxabs_cc
//   MODE 1: child part of update_coarse / set_coarse_phi_rhs: the residual and phi are averaged over
//           2x2x2 cells in the reference's summation order (m_af_restrict.f90:120-133) and written into
//           the parent's tmp / phi; odd-j lanes accumulate, the even-j row arrives by warp shuffle.
template <int NC>
__device__ __forceinline__ void rb_prepare_face(const DevCtx& cx, int r, int var, int t0 = -1, int nt = 0);

// One box of k_resid3: `t` = index among the KS * NC * NC / 2 threads working on it, S = its staged record (2 * COL
// doubles, TMA load in flight on `bar`, whose completion parity is `par`).  LDG: read rhs through the read-only path
// (only valid when rhs is not written during the kernel: not in the persistent k_mega).
template <int NC, int KS, int MODE, bool LDG>
__device__ __forceinline__ void resid3_box(const DevCtx& cx, int slot, int t, const double* S,
                                           unsigned long long* maxabs_bits, int keep_res, uint64_t* bar, uint32_t par) {
  using L = Lay3<NC>;
  constexpr int H = L::H, NI = L::NI, NF = L::NF, COL = L::COL, BOX = L::BOX, KL = NC / KS;
  static_assert(KL % 2 == 0, "z pairs must stay inside one thread");
  const double* grhs = cx.cc[V_RHS] + (size_t)slot * BOX;
  double* gtmp = cx.cc[V_TMP] + (size_t)slot * BOX;
  const int m = t % H, j = (t / H) % NC + 1, ks = t / (H * NC);
  const int k0 = ks * KL + 1;
  // colour of cell A at (j, k0); it alternates with k
  const int cA0 = (1 + j + k0) & 1;
  double rA[KL], rB[KL];
#pragma unroll
  for (int kk = 0; kk < KL; ++kk) {
    const int cA = (cA0 + kk) & 1;
    const int idx = L::iidx(m, j, k0 + kk);
    rA[kk] = (LDG ? __ldg(grhs + cA * COL + idx) : grhs[cA * COL + idx]);
    rB[kk] = (LDG ? __ldg(grhs + (1 - cA) * COL + idx) : grhs[(1 - cA) * COL + idx]);
  }
  const double* cf = cx.coef + 8 * cx.lvl[slot];
  const double c1 = cf[0], c2 = cf[1], c3 = cf[2], c4 = cf[3], c5 = cf[4], c6 = cf[5], c7 = cf[6];
  int ox = 0, oy = 0, oz = 0;
  double *ptmp = nullptr, *pphi = nullptr;  // parent records (possibly on a peer GPU)
  if (MODE == 1) {
    const int p = cx.parent[slot];
    ptmp = cx.at<BOX>(V_TMP, p);
    pphi = cx.at<BOX>(V_PHI, p);
    const int cof = cx.coff[slot];
    ox = (cof & 1) * H;
    oy = ((cof >> 1) & 1) * H;
    oz = ((cof >> 2) & 1) * H;
  }
  const unsigned lanes = __activemask();  // a box of 4^3 cells has fewer than 32 threads
  mbar_wait(bar, par);
  double mx = 0.0, sr = 0.0, sp = 0.0;
  // chain registers: a0/b0 = centre values at k, azm/bzm = values below
  const double* SA = S + cA0 * COL;        // block holding cell A at k0
  const double* SB = S + (1 - cA0) * COL;  // block holding cell B at k0
  double a0 = SA[L::iidx(m, j, k0)], b0 = SB[L::iidx(m, j, k0)];
  double azm = (k0 == 1) ? SB[NI + 4 * NF + (j - 1) * H + m] : SB[L::iidx(m, j, k0 - 1)];
  double bzm = (k0 == 1) ? SA[NI + 4 * NF + (j - 1) * H + m] : SA[L::iidx(m, j, k0 - 1)];
#pragma unroll
  for (int kk = 0; kk < KL; ++kk) {
    const int k = k0 + kk;
    const int idx = L::iidx(m, j, k);
    const int fx = (k - 1) * H + ((j - 1) >> 1);
    const int fy = NI + (k - 1) * H + m, fz = NI + (j - 1) * H + m;
    const double azp = (k < NC) ? SB[idx + NC * H] : SB[fz + 5 * NF];
    const double bzp = (k < NC) ? SA[idx + NC * H] : SA[fz + 5 * NF];
    const double axm = (m > 0) ? SB[idx - 1] : SB[NI + 0 * NF + fx];
    const double bxp = (m < H - 1) ? SA[idx + 1] : SA[NI + 1 * NF + fx];
    const double aym = (j > 1) ? SB[idx - H] : SB[fy + 2 * NF];
    const double ayp = (j < NC) ? SB[idx + H] : SB[fy + 3 * NF];
    const double bym = (j > 1) ? SA[idx - H] : SA[fy + 2 * NF];
    const double byp = (j < NC) ? SA[idx + H] : SA[fy + 3 * NF];
    double la = c1 * a0;
    la = la + c2 * axm;
    la = la + c3 * b0;
    la = la + c4 * aym;
    la = la + c5 * ayp;
    la = la + c6 * azm;
    la = la + c7 * azp;
    double lb = c1 * b0;
    lb = lb + c2 * a0;
    lb = lb + c3 * bxp;
    lb = lb + c4 * bym;
    lb = lb + c5 * byp;
    lb = lb + c6 * bzm;
    lb = lb + c7 * bzp;
    const double resA = rA[kk] - la, resB = rB[kk] - lb;
    const int cA = (cA0 + kk) & 1;
    if (MODE == 0 || keep_res) {
      gtmp[cA * COL + idx] = resA;
      gtmp[(1 - cA) * COL + idx] = resB;
    }
    if (MODE == 0) {
      mx = fmax(mx, fmax(fabs(resA), fabs(resB)));
    } else {
      // rows j (own) and j+1 (lane + H); only odd j accumulates
      const double nA = __shfl_down_sync(lanes, resA, H), nB = __shfl_down_sync(lanes, resB, H);
      if ((kk & 1) == 0) {
        sr = 0.0;
        sp = 0.0;
      }
      sr = sr + resA;
      sr = sr + resB;
      sr = sr + nA;
      sr = sr + nB;
      sp = sp + a0;
      sp = sp + b0;
      sp = sp + ayp;  // phi(2m+1, j+1, k)
      sp = sp + byp;  // phi(2m+2, j+1, k)
      if ((kk & 1) == 1 && (j & 1)) {
        const int qp = L::interior(ox + m + 1, oy + ((j + 1) >> 1), oz + (k >> 1));
        ptmp[qp] = 0.125 * sr;
        pphi[qp] = 0.125 * sp;
      }
    }
    // next step: colours swap, so A/B blocks swap roles
    azm = a0;
    bzm = b0;
    a0 = azp;
    b0 = bzp;
    const double* tsw = SA;
    SA = SB;
    SB = tsw;
  }
  if (MODE == 0 && maxabs_bits && cx.child0[slot] < 0) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(lanes, mx, o, 32));
    if ((t & 31) == 0 && mx > 0.0) atomic_max_nonneg(maxabs_bits, mx);
  }
}

// CTAs beyond nbox (rb_n of them) interpolate refinement-boundary faces rb_r0.. of ANOTHER level (k_rb_prepare's
// work, independent of this kernel's): one graph node less per level on the launch-bound small levels.
template <int NC, int KS, int MODE, int MINB>
__global__ void __launch_bounds__(KS* NC* NC / 2, MINB)
    k_resid3(DevCtx cx, int slot0, int nbox, unsigned long long* maxabs_bits, int keep_res, int rb_r0, int rb_n) {
  pdl_wait();
  if ((int)blockIdx.x >= nbox) {
    rb_prepare_face<NC>(cx, rb_r0 + (int)blockIdx.x - nbox, V_PHI);
    return;
  }
  using L = Lay3<NC>;
  extern __shared__ __align__(128) double smem[];
  __shared__ uint64_t bar;
  const int slot = slot0 + blockIdx.x;
  const int t = threadIdx.x;
  if (cx.opk && cx.opk[slot]) return;  // explicit stencil: k_resid_gen
  if (t == 0) mbar_init(&bar, 1);
  __syncthreads();
  if (t == 0) {
    mbar_expect_tx(&bar, 2 * L::COL * 8);
    bulk_g2s(smem, cx.cc[V_PHI] + (size_t)slot * L::BOX, 2 * L::COL * 8, &bar);
  }
  resid3_box<NC, KS, MODE, true>(cx, slot, t, smem, maxabs_bits, keep_res, &bar, 0);
}

// ---------------------------------------------------------------------------------------------
// Boxes with an explicit operator stencil (afmg_set_stencils: variable-epsilon boxes, level-set /
// electrode boxes, constant stencils that differ from the level's Laplacian).  They are a small
// fraction of a tree (near electrodes and dielectrics), are listed per level and handled by the
// generic kernels below; the fast kernels skip them.
// ---------------------------------------------------------------------------------------------
// coefficient m (0 = centre, 1..6 = -x,+x,-y,+y,-z,+z) of the operator at interior cell (col, idx)
template <int NC>
__device__ __forceinline__ double op_coef(const DevCtx& cx, int kind, const double* sv, const double* cf, int m, int col,
                                          int idx) {
  if (kind == 2) return sv[(size_t)(m * 2 + col) * Lay3<NC>::NI + idx];
  return kind == 1 ? sv[m] : cf[m];
}

// L phi at interior cell (i,j,k) of box record `box` (global memory), any stencil kind, including the
// "- bc_correction" of stencil_apply_357 (m_af_stencil.f90:462-493)
template <int NC>
__device__ __forceinline__ double apply_gen(const DevCtx& cx, int kind, const double* sv, const double* fv,
                                            const double* cf, const double* box, int i, int j, int k,
                                            const double* bvp = nullptr) {
  using L = Lay3<NC>;
  const int col = (i + j + k) & 1, idx = L::iidx((i - 1) >> 1, j, k);
  double acc = op_coef<NC>(cx, kind, sv, cf, 0, col, idx) * box[col * L::COL + idx];
  acc = acc + op_coef<NC>(cx, kind, sv, cf, 1, col, idx) * ldcell<NC>(box, i - 1, j, k);
  acc = acc + op_coef<NC>(cx, kind, sv, cf, 2, col, idx) * ldcell<NC>(box, i + 1, j, k);
  acc = acc + op_coef<NC>(cx, kind, sv, cf, 3, col, idx) * ldcell<NC>(box, i, j - 1, k);
  acc = acc + op_coef<NC>(cx, kind, sv, cf, 4, col, idx) * ldcell<NC>(box, i, j + 1, k);
  acc = acc + op_coef<NC>(cx, kind, sv, cf, 5, col, idx) * ldcell<NC>(box, i, j, k - 1);
  acc = acc + op_coef<NC>(cx, kind, sv, cf, 6, col, idx) * ldcell<NC>(box, i, j, k + 1);
  if (fv) acc = acc - fv[col * L::NI + idx] * (bvp ? bvp[col * L::NI + idx] : cx.lsf_value());
  return acc;
}

// k_gsrb_gen: half-sweep of colour C + side ghost fill for the listed boxes (stencil_gsrb_357,
// m_af_stencil.f90:838-998, all stencil kinds).  With a bc_correction the reference adds it to rhs
// before the sweep and subtracts it afterwards on ALL interior cells (:856-859, :993-996): the rounding
// of (rhs + b) - b is reproduced.  One CTA per box; phi of the other colour is read in place (it is not
// written by this half-sweep).
template <int NC>
__global__ void __launch_bounds__(256) k_gsrb_gen(DevCtx cx, const int* list, int nbox, int C) {
  pdl_wait();
  using L = Lay3<NC>;
  constexpr int NI = L::NI, COL = L::COL, BOX = L::BOX;
  const int slot = list[blockIdx.x];
  const int t = threadIdx.x;
  const int kind = cx.opk[slot];
  const double* sv = cx.stv + cx.opoff[slot];
  const double* fv = cx.foff[slot] >= 0 ? cx.stv + cx.foff[slot] : nullptr;
  const double* bvp = cx.bv_of(slot);
  const double* cf = cx.coef + 8 * cx.lvl[slot];
  double* gbox = cx.cc[V_PHI] + (size_t)slot * BOX;
  double* grhs = cx.cc[V_RHS] + (size_t)slot * BOX;
  const double inv_c1 = (kind == 2) ? 0.0 : 1 / op_coef<NC>(cx, kind, sv, cf, 0, 0, 0);
  for (int n = t; n < 2 * NI; n += 256) {
    const int col = n / NI, idx = n % NI;
    double r = grhs[col * COL + idx];
    double bc = 0.0;
    if (fv) {
      bc = fv[col * NI + idx] * (bvp ? bvp[col * NI + idx] : cx.lsf_value());
      r = r + bc;
    }
    if (col == C) {
      int i, j, k;
      L::uncell(col * COL + idx, i, j, k);
      double acc = r;
      acc = acc - op_coef<NC>(cx, kind, sv, cf, 1, col, idx) * ldcell<NC>(gbox, i - 1, j, k);
      acc = acc - op_coef<NC>(cx, kind, sv, cf, 2, col, idx) * ldcell<NC>(gbox, i + 1, j, k);
      acc = acc - op_coef<NC>(cx, kind, sv, cf, 3, col, idx) * ldcell<NC>(gbox, i, j - 1, k);
      acc = acc - op_coef<NC>(cx, kind, sv, cf, 4, col, idx) * ldcell<NC>(gbox, i, j + 1, k);
      acc = acc - op_coef<NC>(cx, kind, sv, cf, 5, col, idx) * ldcell<NC>(gbox, i, j, k - 1);
      acc = acc - op_coef<NC>(cx, kind, sv, cf, 6, col, idx) * ldcell<NC>(gbox, i, j, k + 1);
      gbox[col * COL + idx] = (kind == 2) ? acc / op_coef<NC>(cx, kind, sv, cf, 0, col, idx) : acc * inv_c1;
    }
    if (fv) grhs[col * COL + idx] = r - bc;
  }
  __syncthreads();
  epilogue_faces<NC, 256, false>(cx, slot, gbox, gbox + COL, 1 << C, t);
}

// k_gsrb2g: the fused half-sweep of k_gsrb2 for a level that CONTAINS boxes with explicit stencils: every box of the
// level goes through the same TMA-staged sweep; a box takes its coefficients from the level's constant Laplacian
// (kind 0, the arithmetic of k_gsrb2), from its own 7 constants (kind 1) or per cell from its coefficient planes
// (kind 2: `/ c(1)` as in stencil_gsrb_357, m_af_stencil.f90:974-990), with the bc_correction of level-set boxes
// added to the right-hand side before and subtracted after the sweep on ALL cells as the reference does (:856-859,
// :993-996; the rounding of (rhs + b) - b is reproduced and written back).  One launch per half-sweep instead of a
// fast and a generic one (k_gsrb_gen, which stays as the unfused reference implementation): on the launch-bound
// streamer trees the generic launch had cost as much as the fast one.  A box's threads are whole warps, so the kind
// branch is warp-uniform.
template <int NC, int BPC, int KS>
__global__ void __launch_bounds__(BPC* KS* NC* NC / 2, 3) k_gsrb2g(DevCtx cx, int slot0, int nbox, int C, int lvl) {
  pdl_wait();
  using L = Lay3<NC>;
  constexpr int H = L::H, NI = L::NI, NF = L::NF, COL = L::COL, BOX = L::BOX;
  constexpr int TPB = H * NC * KS, KL = NC / KS, SBOX = COL + NI;
  extern __shared__ __align__(128) double smem[];
  __shared__ uint64_t bar;
  __shared__ FaceMeta fmeta[BPC];
  const int tid = threadIdx.x;
  const int box0 = blockIdx.x * BPC;
  const int nhere = min(BPC, nbox - box0);
  double* const phi = cx.cc[V_PHI];
  if (tid == 0) mbar_init(&bar, 1);
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(&bar, (uint32_t)(nhere * SBOX * 8));
    for (int b = 0; b < nhere; ++b) {
      const size_t base = (size_t)(slot0 + box0 + b) * BOX;
      bulk_g2s(smem + b * SBOX, phi + base + (1 - C) * COL, COL * 8, &bar);
      bulk_g2s(smem + b * SBOX + COL, cx.cc[V_RHS] + base + C * COL, NI * 8, &bar);
    }
  }
  const int b = tid / TPB, t = tid % TPB;
  const bool active = b < nhere;
  const int slot = slot0 + box0 + (active ? b : 0);
  const double* const S = smem + b * SBOX;
  double* const R = smem + b * SBOX + COL;
  const double* cf = cx.coef + 8 * lvl;
  const int kind = (active && cx.opk) ? cx.opk[slot] : 0;
  const double* sv = kind ? cx.stv + cx.opoff[slot] : nullptr;
  const double* fv = (kind && cx.foff[slot] >= 0) ? cx.stv + cx.foff[slot] : nullptr;
  const double* bvp = fv ? cx.bv_of(slot) : nullptr;
  double* const grhs = cx.cc[V_RHS] + (size_t)slot * BOX;
  if (active && t < 6) prefetch_face_meta(cx, slot, &fmeta[b], t);
  // constant kinds: coefficients once per thread (kind 0: the level's, with the stored reciprocal of c1)
  const double* cc7 = kind == 1 ? sv : cf;
  const double c2 = cc7[1], c3 = cc7[2], c4 = cc7[3], c5 = cc7[4], c6 = cc7[5], c7 = cc7[6];
  const double inv = kind == 1 ? 1 / sv[0] : cf[7];
  const double lsfv = fv ? cx.lsf_value() : 0.0;
  mbar_wait(&bar, 0);
  if (active) {
    const int m = t % H, j = (t / H) % NC + 1, ks = t / (H * NC);
    const int k0 = ks * KL + 1;
    double s_km1 = (k0 == 1) ? S[NI + 4 * NF + (j - 1) * H + m] : S[L::iidx(m, j, k0 - 1)];
    double s_k = S[L::iidx(m, j, k0)];
#pragma unroll
    for (int kk = 0; kk < KL; ++kk) {
      const int k = k0 + kk;
      const int idx = L::iidx(m, j, k);
      const int pi = (C + j + k) & 1;
      const double s_kp1 = (k < NC) ? S[idx + NC * H] : S[NI + 5 * NF + (j - 1) * H + m];
      const double ym = (j > 1) ? S[idx - H] : S[NI + 2 * NF + (k - 1) * H + m];
      const double yp = (j < NC) ? S[idx + H] : S[NI + 3 * NF + (k - 1) * H + m];
      const int fx = (k - 1) * H + ((j - 1) >> 1);
      double xm, xp;
      if (pi) {
        xm = (m > 0) ? S[idx - 1] : S[NI + 0 * NF + fx];
        xp = s_k;
      } else {
        xm = s_k;
        xp = (m < H - 1) ? S[idx + 1] : S[NI + 1 * NF + fx];
      }
      double r = R[idx], bc = 0.0;
      if (fv) {
        bc = fv[C * NI + idx] * (bvp ? bvp[C * NI + idx] : lsfv);
        r = r + bc;
      }
      double acc = r;
      if (kind == 2) {
        const double* q = sv + (size_t)C * NI + idx;  // plane m of colour C: sv[(m * 2 + C) * NI + idx]
        acc = acc - q[(size_t)2 * NI] * xm;
        acc = acc - q[(size_t)4 * NI] * xp;
        acc = acc - q[(size_t)6 * NI] * ym;
        acc = acc - q[(size_t)8 * NI] * yp;
        acc = acc - q[(size_t)10 * NI] * s_km1;
        acc = acc - q[(size_t)12 * NI] * s_kp1;
        R[idx] = acc / q[0];
      } else {
        acc = acc - c2 * xm;
        acc = acc - c3 * xp;
        acc = acc - c4 * ym;
        acc = acc - c5 * yp;
        acc = acc - c6 * s_km1;
        acc = acc - c7 * s_kp1;
        R[idx] = acc * inv;
      }
      if (fv) {  // rhs = (rhs + bc) - bc on both colours, like the reference's add-before / subtract-after
        grhs[C * COL + idx] = r - bc;
        const double bo = fv[(1 - C) * NI + idx] * (bvp ? bvp[(1 - C) * NI + idx] : lsfv);
        const double ro = grhs[(1 - C) * COL + idx];
        grhs[(1 - C) * COL + idx] = (ro + bo) - bo;
      }
      s_km1 = s_k;
      s_k = s_kp1;
    }
  }
  fence_async_smem();
  __syncthreads();
  if (active) {
    if (t == 0) bulk_s2g(phi + (size_t)slot * BOX + C * COL, R, NI * 8);
    epilogue_faces<NC, TPB>(cx, slot, C ? S : R, C ? R : S, 1 << C, t, &fmeta[b]);
    if (t == 0) bulk_commit();
    if (t == 0) bulk_wait_read0();
  }
}

// k_resid_gen: residual_box (MODE 0, + leaf max-norm) or the child part of update_coarse /
// set_coarse_phi_rhs (MODE 1) for the listed boxes, any stencil kind.  One thread per fine cell; for
// MODE 1 the residuals go through shared memory and one thread per coarse cell adds the 2x2x2 values in
// the reference's order (m_af_restrict.f90:120-133).
template <int NC, int MODE>
__global__ void __launch_bounds__(256) k_resid_gen(DevCtx cx, const int* list, int nbox, unsigned long long* maxabs_bits,
                                                   int keep_res) {
  pdl_wait();
  using L = Lay3<NC>;
  constexpr int H = L::H, BOX = L::BOX, NI = L::NI, COL = L::COL;
  extern __shared__ __align__(16) double sres[];  // MODE 1: 2 * NI residuals in interior-block order
  const int slot = list[blockIdx.x];
  const int kind = cx.opk[slot];
  const double* sv = cx.stv + cx.opoff[slot];
  const double* fv = cx.foff[slot] >= 0 ? cx.stv + cx.foff[slot] : nullptr;
  const double* cf = cx.coef + 8 * cx.lvl[slot];
  const double* phi = cx.cc[V_PHI] + (size_t)slot * BOX;
  const double* rhs = cx.cc[V_RHS] + (size_t)slot * BOX;
  double* tmp = cx.cc[V_TMP] + (size_t)slot * BOX;
  double mx = 0.0;
  for (int n = threadIdx.x; n < 2 * NI; n += blockDim.x) {
    const int col = n / NI, idx = n - col * NI;
    const int o = col * COL + idx;
    int i, j, k;
    L::uncell(o, i, j, k);
    const double res = rhs[o] - apply_gen<NC>(cx, kind, sv, fv, cf, phi, i, j, k, cx.bv_of(slot));
    if (MODE == 0 || keep_res) tmp[o] = res;
    if (MODE == 1) sres[n] = res;
    mx = fmax(mx, fabs(res));
  }
  if (MODE == 1) {
    __syncthreads();
    const int p = cx.parent[slot], cof = cx.coff[slot];
    double* ptmp = cx.at<BOX>(V_TMP, p);
    double* pphi = cx.at<BOX>(V_PHI, p);
    const int ox = (cof & 1) * H, oy = ((cof >> 1) & 1) * H, oz = ((cof >> 2) & 1) * H;
    for (int n = threadIdx.x; n < H * H * H; n += blockDim.x) {
      const int ic = n % H + 1, jc = (n / H) % H + 1, kc = n / (H * H) + 1;
      double sr = 0.0, sp = 0.0;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int i = 2 * ic - 1 + (q & 1), j = 2 * jc - 1 + ((q >> 1) & 1), k = 2 * kc - 1 + (q >> 2);
        const int col = (i + j + k) & 1, idx = L::iidx((i - 1) >> 1, j, k);
        sr = sr + sres[col * NI + idx];
        sp = sp + phi[col * COL + idx];
      }
      const int qp = L::interior(ox + ic, oy + jc, oz + kc);
      ptmp[qp] = 0.125 * sr;
      pphi[qp] = 0.125 * sp;
    }
  }
  if (MODE == 0 && maxabs_bits && cx.child0[slot] < 0) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0 && mx > 0.0) atomic_max_nonneg(maxabs_bits, mx);
  }
}

// k_correct3: correct_children with the child's interior staged in shared memory by TMA (bulk load, update
// in place, bulk store).  With push != 0 it also performs the side ghost fill of the af_gc_lvl that
// follows correct_children in the cycle (m_af_multigrid.f90:222, :171): boundary layers of both
// colours are pushed to the neighbours, rule faces are recomputed (epilogue_faces); edges / corners
// are done by k_edges_corners afterwards.
// TPB threads work on one child box and a CTA of 256 threads holds 256 / TPB of them (16^3 boxes: one, with the
// register-sliding prolongation below; 8^3 boxes: four groups of 64 threads, which also take the sliding path --
// with 256 threads per 8^3 box they had fallen to the generic one and needed 6.6 waves of CTAs on the finest level of
// a streamer tree).  cslot0 = first child of the block, nhere = how many of them exist.
template <int NC>
struct Correct3Cfg {
  static constexpr int TPB = (NC == 16) ? 256 : 64;
  static constexpr int BPC = 256 / TPB;
  static constexpr int W = NC / 2 + 2;
  static constexpr int SB = 2 * Lay3<NC>::NI + W * W * W;  // smem doubles per box: I0[NI], I1[NI], sub[W^3]
};

template <int NC, bool LDG>
__device__ __forceinline__ void correct3_box(const DevCtx& cx, int cslot0, int nhere, int push, double* smem, uint64_t* bar,
                                             uint32_t& par) {
  using L = Lay3<NC>;
  constexpr int H = L::H, W = H + 2, NI = L::NI, COL = L::COL, BOX = L::BOX;
  constexpr int TPB = Correct3Cfg<NC>::TPB, BPC = Correct3Cfg<NC>::BPC, SB = Correct3Cfg<NC>::SB;
  const int g = threadIdx.x / TPB, t = threadIdx.x % TPB;
  const bool active = g < nhere;
  const int cslot = cslot0 + (active ? g : 0);
  double* I0 = smem + (size_t)g * SB;
  double* I1 = I0 + NI;
  double* sub = I0 + 2 * NI;
  // cslot: child box (this rank's); its parent may live on a peer GPU
  const int slot = cx.parent[cslot], ch = cx.coff[cslot];
  double* cphi = cx.cc[V_PHI] + (size_t)cslot * BOX;
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, (uint32_t)(nhere * 2 * NI * 8));
    for (int q = 0; q < nhere; ++q) {
      double* cq = cx.cc[V_PHI] + (size_t)(cslot0 + q) * BOX;
      bulk_g2s(smem + (size_t)q * SB, cq, NI * 8, bar);
      bulk_g2s(smem + (size_t)q * SB + NI, cq + COL, NI * 8, bar);
    }
  }
  const double* phi = cx.at<BOX>(V_PHI, slot);
  const double* tmp = cx.at<BOX>(V_TMP, slot);
  const int ox = (ch & 1) * H, oy = ((ch >> 1) & 1) * H, oz = ((ch >> 2) & 1) * H;
  __shared__ FaceMeta fmeta[BPC];
  if (push && active && t < 6) prefetch_face_meta(cx, cslot, &fmeta[g], t);
  if (active) {
    // all loads of the window are issued before the first use (one round trip instead of NW)
    constexpr int NW = (W * W * W + TPB - 1) / TPB;
    double pv[NW], tv[NW];
#pragma unroll
    for (int u = 0; u < NW; ++u) {
      const int n = t + u * TPB;
      if (n < W * W * W) {
        const int a = n % W, b = (n / W) % W, c = n / (W * W);
        const int q = L::cell(ox + a, oy + b, oz + c);
        pv[u] = LDG ? __ldg(phi + q) : phi[q];
        tv[u] = LDG ? __ldg(tmp + q) : tmp[q];
      }
    }
#pragma unroll
    for (int u = 0; u < NW; ++u) {
      const int n = t + u * TPB;
      if (n < W * W * W) sub[n] = pv[u] - tv[u];
    }
  }
  mbar_wait(bar, par);
  par ^= 1u;
  __syncthreads();
  // prolongation stencil of this child: the default one, or what the host shipped for the box
  // (constant p248 / p234, or variable p234 for variable-epsilon boxes, m_af_multigrid.f90:1308-1388)
  const int pkind = cx.pk ? cx.pk[cslot] : 0;
  const double* pv = pkind ? cx.stv + cx.poff[cslot] : cx.pcoef;
  double pc[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) pc[q] = (pkind == 3 || (pkind == 2 && q >= 4)) ? 0.0 : pv[q];
  const int pshape = (pkind == 0) ? cx.pshape : (pkind == 1 ? 8 : 4);
  constexpr int TPBX = H * NC;  // threads per k-range
  constexpr int KSX = (TPB / TPBX < NC) ? TPB / TPBX : NC;
  constexpr int KLX = NC / KSX;
  if (active && t < TPBX * KSX && pkind == 0 && pshape == 8 && (KLX % 2) == 0) {
    // Default stencil_prolong_248 (m_af_stencil.f90:766-813), the common case.  A thread owns the cell pair
    // i = 2m+1, 2m+2 of row j and walks k: the 3 x 2 x 3 coarse values it needs per fine k-pair stay in
    // registers and slide along k (6 LDS per k-pair instead of 32); the coefficients are the literals of
    // mg_box_prolong_linear_stencil (m_af_multigrid.f90:1282), which is what cx.pcoef holds here.
    constexpr double c27 = 27 / 64.0, c9 = 9 / 64.0, c3 = 3 / 64.0, c1 = 1 / 64.0;
    const int m = t % H, j = (t / H) % NC + 1, ks = t / TPBX;
    const int j1 = (j + 1) >> 1, j2 = j1 + 1 - 2 * (j & 1);
    const int K0 = ks * (KLX / 2);  // coarse plane below the first one this thread centres on
    double A[2][3], B[2][3], C[2][3];
    auto load_plane = [&](double (&P)[2][3], int kz) {
      const double* r1 = sub + (kz * W + j1) * W + m;
      const double* r2 = sub + (kz * W + j2) * W + m;
#pragma unroll
      for (int x = 0; x < 3; ++x) {
        P[0][x] = r1[x];
        P[1][x] = r2[x];
      }
    };
    load_plane(A, K0);
    load_plane(B, K0 + 1);
    const int jpar = j & 1;
#pragma unroll
    for (int kc = 0; kc < KLX / 2; ++kc) {
      load_plane(C, K0 + kc + 2);
#pragma unroll
      for (int ko = 0; ko < 2; ++ko) {  // ko = 0: odd fine k (k2 = plane below), 1: even k (plane above)
        const int k = ks * KLX + 2 * kc + ko + 1;
        const double(&Q)[2][3] = ko ? C : A;
#pragma unroll
        for (int p = 0; p < 2; ++p) {  // p = 1: i = 2m+1 (i2 = i1 - 1), p = 0: i = 2m+2 (i2 = i1 + 1)
          const int x2 = p ? 0 : 2;
          const int c = (p + jpar + ko + 1) & 1;  // colour of that cell: (i + j + k) & 1
          double* Ic = c ? I1 : I0;
          const int idx = L::iidx(m, j, k);
          double acc = Ic[idx];
          acc = acc + c27 * B[0][1];
          acc = acc + c9 * B[0][x2];
          acc = acc + c9 * B[1][1];
          acc = acc + c3 * B[1][x2];
          acc = acc + c9 * Q[0][1];
          acc = acc + c3 * Q[0][x2];
          acc = acc + c3 * Q[1][1];
          acc = acc + c1 * Q[1][x2];
          Ic[idx] = acc;
        }
      }
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int x = 0; x < 3; ++x) {
          A[r][x] = B[r][x];
          B[r][x] = C[r][x];
        }
    }
  } else if (active && t < TPBX * KSX) {
    const int m = t % H, j = (t / H) % NC + 1, ks = t / TPBX;
    const int j1 = (j + 1) >> 1, j2 = j1 + 1 - 2 * (j & 1);
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      double* Ic = c ? I1 : I0;
#pragma unroll
      for (int kk = 0; kk < KLX; ++kk) {
        const int k = ks * KLX + kk + 1;
        const int i = 2 * m + 2 - ((c + j + k) & 1);
        const int i1 = (i + 1) >> 1, i2 = i1 + 1 - 2 * (i & 1);
        const int k1 = (k + 1) >> 1, k2 = k1 + 1 - 2 * (k & 1);
        const double* r11 = sub + (k1 * W + j1) * W;
        const double* r21 = sub + (k1 * W + j2) * W;
        const double* r12 = sub + (k2 * W + j1) * W;
        const double* r22 = sub + (k2 * W + j2) * W;
        const int idx = L::iidx(m, j, k);
        double acc = Ic[idx];
        if (pshape == 8) {
          acc = acc + pc[0] * r11[i1];
          acc = acc + pc[1] * r11[i2];
          acc = acc + pc[2] * r21[i1];
          acc = acc + pc[3] * r21[i2];
          acc = acc + pc[4] * r12[i1];
          acc = acc + pc[5] * r12[i2];
          acc = acc + pc[6] * r22[i1];
          acc = acc + pc[7] * r22[i2];
        } else if (pkind != 3) {
          acc = acc + pc[0] * r11[i1];
          acc = acc + pc[1] * r11[i2];
          acc = acc + pc[2] * r21[i1];
          acc = acc + pc[3] * r12[i1];
        } else {
          acc = acc + pv[(size_t)(0 * 2 + c) * NI + idx] * r11[i1];
          acc = acc + pv[(size_t)(1 * 2 + c) * NI + idx] * r11[i2];
          acc = acc + pv[(size_t)(2 * 2 + c) * NI + idx] * r21[i1];
          acc = acc + pv[(size_t)(3 * 2 + c) * NI + idx] * r12[i1];
        }
        Ic[idx] = acc;
      }
    }
  }
  fence_async_smem();
  __syncthreads();
  if (active && t == 0) {
    bulk_s2g(cphi, I0, NI * 8);
    bulk_s2g(cphi + COL, I1, NI * 8);
  }
  if (push && active) epilogue_faces<NC, TPB>(cx, cslot, I0, I1, 3, t, &fmeta[g]);
  if (active && t == 0) {
    bulk_commit();
    bulk_wait_read0();
  }
}

template <int NC>
__global__ void __launch_bounds__(256, 4) k_correct3(DevCtx cx, int slot0, int nbox, int push) {
  pdl_wait();
  extern __shared__ __align__(128) double smem[];
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) mbar_init(&bar, 1);
  __syncthreads();
  uint32_t par = 0;
  constexpr int BPC = Correct3Cfg<NC>::BPC;
  const int b0 = blockIdx.x * BPC;
  correct3_box<NC, true>(cx, slot0 + b0, min(BPC, nbox - b0), push, smem, &bar, par);
}

template <int NC>
__device__ void gc_sides(const DevCtx& cx, int slot, int var, double* stage = nullptr);
template <int NC>
__device__ void gc_edges_corners(const DevCtx& cx, int slot, int var, int t0 = -1, int nt = 0);

// k_gc2: af_gc_lvl for one level, and for boxes with children the parent part of update_coarse
// (rhs = L phi + tmp; tmp = phi on the full record, m_af_multigrid.f90:722-736) computed from a shared-memory copy
// of the box: the two interior colour blocks arrive by TMA while the ghost faces are being gathered, and the
// gathered values go straight into the copy (no read-back of what the CTA just wrote).
template <int NC>
__device__ __forceinline__ void gc2_box(const DevCtx& cx, int slot, int corners, int mode, double* smem, uint64_t* bar,
                                        uint32_t& par) {
  using L = Lay3<NC>;
  constexpr int H = L::H, NI = L::NI, NF = L::NF, COL = L::COL, BOX = L::BOX;
  // smem: 2 * COL doubles
  const int t = threadIdx.x;
  double* gphi = cx.cc[V_PHI] + (size_t)slot * BOX;
  const bool upd = mode != 0 && cx.child0[slot] >= 0;
  if (upd && t == 0) {
    mbar_expect_tx(bar, 2 * NI * 8);
    bulk_g2s(smem, gphi, NI * 8, bar);
    bulk_g2s(smem + COL, gphi + COL, NI * 8, bar);
  }
  gc_sides<NC>(cx, slot, V_PHI, upd ? smem : nullptr);
  __syncthreads();
  if (corners) gc_edges_corners<NC>(cx, slot, V_PHI);
  if (!upd) return;
  mbar_wait(bar, par);
  par ^= 1u;
  __syncthreads();
  double* rhs = cx.cc[V_RHS] + (size_t)slot * BOX;
  double* tmp = cx.cc[V_TMP] + (size_t)slot * BOX;
  const double* cf = cx.coef + 8 * cx.lvl[slot];
  const double c1 = cf[0];
  // explicit stencil of this box, if any (apply_gen only reads interior + face cells: both are staged)
  const int okind = cx.opk ? cx.opk[slot] : 0;
  const double* osv = okind ? cx.stv + cx.opoff[slot] : nullptr;
  const double* ofv = (okind && cx.foff[slot] >= 0) ? cx.stv + cx.foff[slot] : nullptr;
#pragma unroll
  for (int c = 0; c < 2; ++c)
    for (int idx = t; idx < NI; idx += 256) {
      const int m = idx % H, j = (idx / H) % NC + 1, k = idx / (H * NC) + 1;
      const int i = 2 * m + 2 - ((c + j + k) & 1);
      const int q = c * COL + idx;
      const double lp = okind ? apply_gen<NC>(cx, okind, osv, ofv, cf, smem, i, j, k, cx.bv_of(slot))
                              : apply357_smem<NC>(smem, cf, c1, i, j, k);
      rhs[q] = lp + tmp[q];
      if (mode == 1) tmp[q] = smem[q];
    }
  if (mode == 1) {  // tmp = phi on the ghost cells too: faces from the copy, edges / corners from global
    for (int n = t; n < 2 * 6 * NF; n += 256) {
      const int q = (n / (6 * NF)) * COL + NI + n % (6 * NF);
      tmp[q] = smem[q];
    }
    for (int q = 2 * COL + t; q < BOX; q += 256) tmp[q] = gphi[q];
  }
}

template <int NC>
__global__ void __launch_bounds__(256) k_gc2(DevCtx cx, int slot0, int nbox, int corners, int mode) {
  pdl_wait();
  extern __shared__ __align__(128) double smem[];  // 2*COL
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) mbar_init(&bar, 1);
  __syncthreads();
  uint32_t par = 0;
  gc2_box<NC>(cx, slot0 + blockIdx.x, corners, mode, smem, &bar, par);
}

// tmp_p = phi_p - tmp_p on the full record of every box of [slot0, slot0+nbox) that has children
// (first statement of correct_children, m_af_multigrid.f90:636-637)
template <int NC>
__global__ void k_store_corr(DevCtx cx, int slot0, int nbox) {
  pdl_wait();
  using L = Lay3<NC>;
  const int slot = slot0 + blockIdx.x;
  if (cx.child0[slot] < 0) return;
  const double* phi = cx.cc[V_PHI] + (size_t)slot * L::BOX;
  double* tmp = cx.cc[V_TMP] + (size_t)slot * L::BOX;
  for (int q = threadIdx.x; q < L::BOX; q += blockDim.x) tmp[q] = phi[q] - tmp[q];
}

// ---------------------------------------------------------------------------------------------
// k_rb_prepare: for every refinement-boundary face of a level, interpolate the coarse neighbour's
// boundary layer to the fine face (first half of mg_sides_rb, m_af_multigrid.f90:294-380).  The
// coarse level is frozen while the fine level is smoothed, so this runs once per gsrb_boxes /
// af_gc_lvl group.  One CTA per face, one thread per fine face cell.
// ---------------------------------------------------------------------------------------------
// t0 / nt: index of the calling thread among the nt threads that share this face (default: the whole CTA)
template <int NC>
__device__ __forceinline__ void rb_prepare_face(const DevCtx& cx, int r, int var, int t0, int nt) {
  using L = Lay3<NC>;
  constexpr int H = L::H;
  if (t0 < 0) {
    t0 = threadIdx.x;
    nt = blockDim.x;
  }
  const int s = cx.rb_slot[r], f = cx.rb_face[r];
  const int p = cx.parent[s];
  const int pn = cx.nbr[p * 6 + f];  // coarse neighbour (exists by 2:1 balance)
  const int d = f >> 1;
  const int layer = (f & 1) ? 1 : NC;
  const int cof = cx.coff[s];
  const int ta = (d == 0) ? 1 : 0, tb = (d == 2) ? 1 : 2;
  const int coa = ((cof >> ta) & 1) * H, cob = ((cof >> tb) & 1) * H;
  const double* cb = cx.at<L::BOX>(var, pn);
  auto T = [&](int x, int y) {
    int q[3];
    q[d] = layer;
    q[ta] = coa + x;
    q[tb] = cob + y;
    return ldcell<NC>(cb, q[0], q[1], q[2]);
  };
  double* out = cx.rule_B + (size_t)(cx.rb_row0 + r) * L::NC2;
  if (cx.rule_flag && cx.rule_flag[cx.rb_row0 + r]) {
    // mg_sides_rb_extrap: the ghost cell first takes the value of the PARENT's cell covering it
    // (af_gc_prolong_copy, m_af_ghostcell.f90:378-390: i_c1 = offset + (i+1)/2 with i the ghost index)
    const double* pb = cx.at<L::BOX>(var, p);
    const int g = (f & 1) ? NC + 1 : 0;
    const int cod = ((cof >> d) & 1) * H;
    for (int n = t0; n < L::NC2; n += nt) {
      const int a = n % NC + 1, b = n / NC + 1;
      int q[3];
      q[d] = cod + ((g + 1) >> 1);
      q[ta] = coa + ((a + 1) >> 1);
      q[tb] = cob + ((b + 1) >> 1);
      out[n] = ldcell<NC>(pb, q[0], q[1], q[2]);
    }
    return;
  }
  for (int n = t0; n < L::NC2; n += nt) {
    const int a = n % NC + 1, b = n / NC + 1;
    const int ia = (a + 1) >> 1, ib = (b + 1) >> 1;
    const double t0 = T(ia, ib);
    const double g1 = 0.125 * (T(ia + 1, ib) - T(ia - 1, ib));
    const double g2 = 0.125 * (T(ia, ib + 1) - T(ia, ib - 1));
    double v = (a & 1) ? (t0 - g1) : (t0 + g1);
    v = (b & 1) ? (v - g2) : (v + g2);
    out[n] = v;
  }
}
template <int NC>
__global__ void k_rb_prepare(DevCtx cx, int r0, int nr, int var) {
  pdl_wait();
  if ((int)blockIdx.x >= nr) return;
  rb_prepare_face<NC>(cx, r0 + blockIdx.x, var);
}

// ---------------------------------------------------------------------------------------------
// Ghost cells by gathering (af_gc_box, m_af_ghostcell.f90:64-170): sides, then edges and corners.
// ---------------------------------------------------------------------------------------------
template <int NC>
__device__ void gc_sides(const DevCtx& cx, int slot, int var, double* stage) {
  using L = Lay3<NC>;
  double* box = cx.cc[var] + (size_t)slot * L::BOX;
  for (int n = threadIdx.x; n < 6 * L::NC2; n += blockDim.x) {
    const int f = n / L::NC2, rr = n % L::NC2;
    const int a = rr % NC + 1, b = rr / NC + 1;
    const int d = f >> 1, hi = f & 1;
    const int ta = (d == 0) ? 1 : 0, tb = (d == 2) ? 1 : 2;
    int q[3];
    q[ta] = a;
    q[tb] = b;
    const int nb = cx.nbr[slot * 6 + f];
    double v;
    if (nb >= 0) {  // copy_from_nb: ghost (g) <- neighbour cell g - dnb*nc
      q[d] = hi ? 1 : NC;
      v = ldcell<NC>(cx.at<L::BOX>(var, nb), q[0], q[1], q[2]);
    } else {
      const int row = cx.aux[slot * 6 + f];
      const double* rc = cx.rule_c + 3 * row;
      const double B = cx.rule_B[(size_t)row * L::NC2 + rr];
      q[d] = hi ? NC : 1;
      const double x1 = ldcell<NC>(box, q[0], q[1], q[2]);
      q[d] = hi ? NC - 1 : 2;
      if (cx.rule_flag && cx.rule_flag[row]) {  // mg_sides_rb_extrap: diagonal second point
        q[ta] = a - 1 + 2 * (a & 1);
        q[tb] = b - 1 + 2 * (b & 1);
      }
      const double x2 = ldcell<NC>(box, q[0], q[1], q[2]);
      v = (rc[0] * B + rc[1] * x1) + rc[2] * x2;
    }
    box[L::face(f, a, b)] = v;
    if (stage) stage[L::face(f, a, b)] = v;  // the caller's shared-memory copy of the box
  }
}

// af_gc_box_corner (m_af_ghostcell.f90:125-170): needs the box's own face ghosts (call after a
// __syncthreads following gc_sides, or in a later kernel)
template <int NC>
__device__ void gc_edges_corners(const DevCtx& cx, int slot, int var, int t0, int nt) {
  using L = Lay3<NC>;
  double* box = cx.cc[var] + (size_t)slot * L::BOX;
  if (t0 < 0) {
    t0 = threadIdx.x;
    nt = blockDim.x;
  }
  for (int n = t0; n < 12 * NC + 8; n += nt) {
    if (n < 12 * NC) {
      const int e = n / NC, pos = n % NC + 1, dim = e >> 2;
      const int o1 = (dim == 0) ? 1 : 0, o2 = (dim == 2) ? 1 : 2;
      int dir[3] = {0, 0, 0}, q[3];
      dir[o1] = (e & 1) ? 1 : -1;
      dir[o2] = (e & 2) ? 1 : -1;
      q[dim] = pos;
      q[o1] = (e & 1) ? NC + 1 : 0;
      q[o2] = (e & 2) ? NC + 1 : 0;
      const int nb = cx.nmat[slot * 27 + nmat_index(dir[0], dir[1], dir[2])];
      double v;
      if (nb >= 0) {
        v = ldcell<NC>(cx.at<L::BOX>(var, nb), q[0] - dir[0] * NC, q[1] - dir[1] * NC, q[2] - dir[2] * NC);
      } else {  // af_edge_gc_extrap (:885-924): a + b - c
        int qa[3] = {q[0], q[1], q[2]}, qb[3] = {q[0], q[1], q[2]}, qc[3] = {q[0], q[1], q[2]};
        qa[o1] -= dir[o1];
        qb[o2] -= dir[o2];
        qc[o1] -= dir[o1];
        qc[o2] -= dir[o2];
        // the reference orders the two face-ghost terms by (dim+1, dim+2) cyclically
        const int c1 = (dim + 1) % 3;
        const double va = ldcell<NC>(box, qa[0], qa[1], qa[2]);
        const double vb = ldcell<NC>(box, qb[0], qb[1], qb[2]);
        const double vc = ldcell<NC>(box, qc[0], qc[1], qc[2]);
        v = (c1 == o1) ? (va + vb - vc) : (vb + va - vc);
      }
      box[L::edge(e, pos)] = v;
    } else {
      const int c = n - 12 * NC;
      const int dx = (c & 1) ? 1 : -1, dy = (c & 2) ? 1 : -1, dz = (c & 4) ? 1 : -1;
      const int qi = (c & 1) ? NC + 1 : 0, qj = (c & 2) ? NC + 1 : 0, qk = (c & 4) ? NC + 1 : 0;
      const int nb = cx.nmat[slot * 27 + nmat_index(dx, dy, dz)];
      double v;
      if (nb >= 0) {
        v = ldcell<NC>(cx.at<L::BOX>(var, nb), qi - dx * NC, qj - dy * NC, qk - dz * NC);
      } else {  // af_corner_gc_extrap (:860-879)
        v = ldcell<NC>(box, qi, qj - dy, qk - dz) + ldcell<NC>(box, qi - dx, qj, qk - dz) +
            ldcell<NC>(box, qi - dx, qj - dy, qk) - 2 * ldcell<NC>(box, qi - dx, qj - dy, qk - dz);
      }
      box[L::corner(c)] = v;
    }
  }
}

// k_gc: af_gc_lvl (m_af_ghostcell.f90:49-61) for boxes [slot0, slot0+nbox).  One CTA per box.  (The
// variant that also performs the parent part of update_coarse is k_gc2.)
template <int NC>
__global__ void k_gc(DevCtx cx, int slot0, int nbox, int var, int corners, int mode) {
  pdl_wait();
  using L = Lay3<NC>;
  const int slot = slot0 + blockIdx.x;
  if ((int)blockIdx.x >= nbox) return;
  gc_sides<NC>(cx, slot, var);
  if (corners) {
    __syncthreads();
    gc_edges_corners<NC>(cx, slot, var);
  }
}

// edges + corners only (after the last half-sweep of an upward gsrb_boxes, or every half-sweep when
// mg%use_corners, m_af_multigrid.f90:676-684)
template <int NC>
__global__ void k_edges_corners(DevCtx cx, int slot0, int nbox, int var, int rb_r0, int rb_n) {
  pdl_wait();
  const int slot = slot0 + blockIdx.x;
  if ((int)blockIdx.x >= nbox) {  // piggy-backed k_rb_prepare work of the next finer level (see k_resid3)
    if ((int)blockIdx.x < nbox + rb_n) rb_prepare_face<NC>(cx, rb_r0 + (int)blockIdx.x - nbox, V_PHI);
    return;
  }
  gc_edges_corners<NC>(cx, slot, var);
}

// restriction of one variable only (init_phi_rhs, m_af_multigrid.f90:779-799: phi = 0, restrict rhs)
template <int NC>
__global__ void k_restrict_var(DevCtx cx, int slot0, int nbox, int var, int clear_phi) {
  pdl_wait();
  using L = Lay3<NC>;
  constexpr int H = L::H;
  const int slot = slot0 + blockIdx.x;
  if ((int)blockIdx.x >= nbox) return;
  const double* src = cx.cc[var] + (size_t)slot * L::BOX;
  const int p = cx.parent[slot], cof = cx.coff[slot];
  double* dst = cx.at<L::BOX>(var, p);
  const int ox = (cof & 1) * H, oy = ((cof >> 1) & 1) * H, oz = ((cof >> 2) & 1) * H;
  if (clear_phi) {
    double* phi = cx.cc[V_PHI] + (size_t)slot * L::BOX;
    for (int q = threadIdx.x; q < L::BOX; q += blockDim.x) phi[q] = 0.0;
  }
  for (int n = threadIdx.x; n < H * H * H; n += blockDim.x) {
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
}

// max |var| over the interior of leaves
template <int NC>
__global__ void k_maxabs(DevCtx cx, int slot0, int nbox, int var, unsigned long long* maxabs_bits) {
  pdl_wait();
  using L = Lay3<NC>;
  const int slot = slot0 + blockIdx.x;
  if ((int)blockIdx.x >= nbox || cx.child0[slot] >= 0) return;
  const double* v = cx.cc[var] + (size_t)slot * L::BOX;
  double mx = 0.0;
  for (int n = threadIdx.x; n < 2 * L::NI; n += blockDim.x) {
    const int q = (n < L::NI) ? n : (L::COL + n - L::NI);
    mx = fmax(mx, fabs(v[q]));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0 && mx > 0.0) atomic_max_nonneg(maxabs_bits, mx);
}

// Order-independent bitwise checksum of whole box records (interior + ghost cells): wrapping sum and XOR of the
// 64-bit patterns.  Used to prove that N-GPU solves are bit-identical to the 1-GPU solve (bench.py, mgpu_check).
__global__ void k_checksum(const double* base, size_t n, unsigned long long* out) {
  pdl_wait();
  unsigned long long s = 0, x = 0;
  const unsigned long long* p = reinterpret_cast<const unsigned long long*>(base);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const unsigned long long v = p[i];
    s += v;
    x ^= v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    x ^= __shfl_xor_sync(0xffffffffu, x, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(out, s);
    atomicXor(out + 1, x);
  }
}

// per-leaf-box sums of the interior in (k,j,i) order (af_tree_sum_cc, m_af_utils.f90:966-1027); the
// host adds fac(lvl) * sum in box order.  One warp per box is plenty (only used by subtract_mean).
template <int NC>
__global__ void k_box_sums(DevCtx cx, int slot0, int nbox, int var, double* out) {
  pdl_wait();
  using L = Lay3<NC>;
  const int slot = slot0 + blockIdx.x;
  if ((int)blockIdx.x >= nbox) return;
  const double* v = cx.cc[var] + (size_t)slot * L::BOX;
  __shared__ double part[NC * NC];
  for (int n = threadIdx.x; n < NC * NC; n += blockDim.x) {  // row sums in i order
    const int j = n % NC + 1, k = n / NC + 1;
    double s = 0.0;
    for (int i = 1; i <= NC; ++i) s = s + v[L::interior(i, j, k)];
    part[n] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int n = 0; n < NC * NC; ++n) s = s + part[n];
    out[slot] = s;
    for (int r = 0; r < cx.nranks; ++r)  // multi-GPU: every rank keeps the complete table
      if (r != cx.me && cx.nranks > 1) cx.bsum[r][slot] = s;
  }
}

// ---------------------------------------------------------------------------------------------
// streaming helpers on whole box records: copy (af_boxes_copy_cc, m_af_utils.f90:553-563), clear
// (af_box_clear_cc :385), add constant (subtract_mean, m_af_multigrid.f90:247-257)
// ---------------------------------------------------------------------------------------------
__global__ void k_copy(double* __restrict__ dst, const double* __restrict__ src, size_t n) {
  pdl_wait();
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) dst[i] = src[i];
}
__global__ void k_fill(double* dst, double v, size_t n) {
  pdl_wait();
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) dst[i] = v;
}
__global__ void k_sub_scalar(double* dst, const double* scalar, size_t n) {
  pdl_wait();
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const double s = *scalar;
  for (; i < n; i += stride) dst[i] = dst[i] - s;
}

// ---------------------------------------------------------------------------------------------
// pack / unpack between the reference's cc(0:nc+1,0:nc+1,0:nc+1) order and the device layout
// ---------------------------------------------------------------------------------------------
template <int NC>
__global__ void k_unpack(double* var_base, const int* slots, int n, const double* packed) {
  pdl_wait();
  using L = Lay3<NC>;
  constexpr int N2 = NC + 2;
  if ((int)blockIdx.x >= n || slots[blockIdx.x] < 0) return;
  double* box = var_base + (size_t)slots[blockIdx.x] * L::BOX;
  const double* src = packed + (size_t)blockIdx.x * L::BOX;
  for (int q = threadIdx.x; q < L::BOX; q += blockDim.x) {
    int i, j, k;
    L::uncell(q, i, j, k);
    box[q] = src[(k * N2 + j) * N2 + i];
  }
}
// field_set_rhs (src/m_field.f90:406-444): rhs = 0; rhs = rhs + q_n * density_n, one call per species with the
// packed density of the listed boxes (reference order); first != 0 starts from the reference's 0.0
template <int NC>
__global__ void k_unpack_axpy(double* var_base, const int* slots, int n, const double* packed, double q, int first) {
  pdl_wait();
  using L = Lay3<NC>;
  constexpr int N2 = NC + 2;
  if ((int)blockIdx.x >= n || slots[blockIdx.x] < 0) return;
  double* box = var_base + (size_t)slots[blockIdx.x] * L::BOX;
  const double* src = packed + (size_t)blockIdx.x * L::BOX;
  for (int p = threadIdx.x; p < L::BOX; p += blockDim.x) {
    int i, j, k;
    L::uncell(p, i, j, k);
    const double prev = first ? 0.0 : box[p];
    box[p] = prev + q * src[(k * N2 + j) * N2 + i];
  }
}
// interior cells only: packed holds cc(1:nc, 1:nc, 1:nc) per box (ghost cells of the record are kept)
template <int NC>
__global__ void k_unpack_interior(double* var_base, const int* slots, int n, const double* packed) {
  pdl_wait();
  using L = Lay3<NC>;
  if ((int)blockIdx.x >= n || slots[blockIdx.x] < 0) return;
  double* box = var_base + (size_t)slots[blockIdx.x] * L::BOX;
  const double* src = packed + (size_t)blockIdx.x * (NC * NC * NC);
  for (int q = threadIdx.x; q < 2 * L::NI; q += blockDim.x) {
    const int o = (q < L::NI) ? q : (L::COL + q - L::NI);
    int i, j, k;
    L::uncell(o, i, j, k);
    box[o] = src[((k - 1) * NC + (j - 1)) * NC + (i - 1)];
  }
}
// interior cells only: packed receives cc(1:nc, 1:nc, 1:nc) per box
template <int NC>
__global__ void k_pack_interior(const double* var_base, const int* slots, int n, double* packed) {
  pdl_wait();
  using L = Lay3<NC>;
  if ((int)blockIdx.x >= n || slots[blockIdx.x] < 0) return;
  const double* box = var_base + (size_t)slots[blockIdx.x] * L::BOX;
  double* dst = packed + (size_t)blockIdx.x * (NC * NC * NC);
  // one thread per output cell (coalesced stores; the two colour blocks are read with stride 2)
  for (int q = threadIdx.x; q < NC * NC * NC; q += blockDim.x) {
    const int i = q % NC + 1, j = (q / NC) % NC + 1, k = q / (NC * NC) + 1;
    dst[q] = box[L::interior(i, j, k)];
  }
}
template <int NC>
__global__ void k_pack(const double* var_base, const int* slots, int n, double* packed) {
  pdl_wait();
  using L = Lay3<NC>;
  constexpr int N2 = NC + 2;
  if ((int)blockIdx.x >= n || slots[blockIdx.x] < 0) return;
  const double* box = var_base + (size_t)slots[blockIdx.x] * L::BOX;
  double* dst = packed + (size_t)blockIdx.x * L::BOX;
  for (int q = threadIdx.x; q < L::BOX; q += blockDim.x) {
    int i, j, k;
    L::uncell(q, i, j, k);
    dst[(k * N2 + j) * N2 + i] = box[q];
  }
}

// ---------------------------------------------------------------------------------------------
// Coarse grid (replaces Hypre, m_coarse_solver.f90): the BC-folded level-1 operator of a constant
// coefficient Laplace/Helmholtz problem is separable, A = Tx (x) I (x) I + I (x) Ty (x) I + I (x) I (x) Tz
// (- lambda), so  x = Q diag(1/(lx+ly+lz-lambda)) Q^T b  with the 1D eigenvectors Q = Qx (x) Qy (x) Qz:
// an exact direct solve ("fast diagonalisation").
// ---------------------------------------------------------------------------------------------
struct CoarseCtx {
  int nx[3];            // coarse grid size in cells
  int nb[3];            // level-1 boxes per dim
  const int* bix;       // [nbox1*3] 0-based position of each level-1 box in the coarse grid (box%ix - 1)
  const double* b2r;    // [nbox1][6][NC2] bc_to_rhs (stencil_handle_boundaries, m_coarse_solver.f90:442-491)
  const double* Q[3];   // eigenvectors, Q[d][row*n + col], column = eigenvector
  const double* inv_eig;  // [n] 1 / (lx(i)+ly(j)+lz(k) - lambda)
  double* v0;           // [n] work vectors
  double* v1;
  const double* lsf_fac;  // [nbox1][nc^3] stencil%f of level-1 boxes (rhs += f * lsf_boundary_value), or null
  const double* Ainv;     // [n][n] dense inverse (general path: explicit stencils on level 1), or null
  // block-tridiagonal path for large non-separable coarse grids (k_cs_plane_*): planes stacked along `sdim`
  int sdim, np, m;        // stacking dimension, number of planes, cells per plane
  const double* Sinv;     // [np][m][m] inverses of the Schur-complement planes
  const double* lo;       // [n] coupling of every cell to the plane below (0 on the first plane), global cell order
  const double* up;       // [n] coupling to the plane above
  double* w;              // [n] work vectors in (plane, in-plane) order
  double* xs;
  double* tvec;           // [m]
};

// (plane, in-plane index) of global coarse cell g and back: the in-plane index runs over the two other dimensions
// in increasing dimension order
__device__ __forceinline__ void cs_plane_of(const CoarseCtx& cs, int g, int& p, int& i) {
  const int q[3] = {g % cs.nx[0], (g / cs.nx[0]) % cs.nx[1], g / (cs.nx[0] * cs.nx[1])};
  const int a = cs.sdim == 0 ? 1 : 0, b = cs.sdim == 2 ? 1 : 2;
  p = q[cs.sdim];
  i = q[a] + cs.nx[a] * q[b];
}
__device__ __forceinline__ int cs_global_of(const CoarseCtx& cs, int p, int i) {
  const int a = cs.sdim == 0 ? 1 : 0, b = cs.sdim == 2 ? 1 : 2;
  int q[3];
  q[cs.sdim] = p;
  q[a] = i % cs.nx[a];
  q[b] = i / cs.nx[a];
  return q[0] + cs.nx[0] * (q[1] + cs.nx[1] * q[2]);
}

// coarse_solver_set_rhs_phi (m_coarse_solver.f90:286-338): b = rhs + bc_to_rhs * bc_val per face
template <int NC>
__device__ __forceinline__ void cs_gather_cell(const DevCtx& cx, const CoarseCtx& cs, int nbox1, int n);
template <int NC>
__global__ void k_cs_gather(DevCtx cx, CoarseCtx cs, int nbox1) {
  pdl_wait();
  cs_gather_cell<NC>(cx, cs, nbox1, blockIdx.x * blockDim.x + threadIdx.x);
}
template <int NC>
__device__ __forceinline__ void cs_gather_cell(const DevCtx& cx, const CoarseCtx& cs, int nbox1, int n) {
  using L = Lay3<NC>;
  const int ncell = NC * NC * NC;
  if (n >= nbox1 * ncell) return;
  const int bx = n / ncell, r = n % ncell;
  const int i = r % NC + 1, j = (r / NC) % NC + 1, k = r / (NC * NC) + 1;
  const double* rhs = cx.cc[V_RHS] + (size_t)bx * L::BOX;  // level-1 boxes occupy slots 0..nbox1-1
  double t = rhs[L::interior(i, j, k)];
  const int q[3] = {i, j, k};
#pragma unroll
  for (int f = 0; f < 6; ++f) {
    const int d = f >> 1;
    if (cx.nbr[bx * 6 + f] >= 0) continue;
    if (q[d] != ((f & 1) ? NC : 1)) continue;
    const int ta = (d == 0) ? 1 : 0, tb = (d == 2) ? 1 : 2;
    const int fi = (q[ta] - 1) + (q[tb] - 1) * NC;
    const int row = cx.aux[bx * 6 + f];
    t = t + cs.b2r[((size_t)bx * 6 + f) * L::NC2 + fi] * cx.rule_B[(size_t)row * L::NC2 + fi];
  }
  // level-set boundary inside the coarse grid (m_coarse_solver.f90:320-324)
  if (cs.lsf_fac) {
    const double* bvp = cx.bv_of(bx);
    t = t + cs.lsf_fac[(size_t)bx * ncell + r] * (bvp ? bvp[((i + j + k) & 1) * L::NI + L::iidx((i - 1) >> 1, j, k)] : cx.lsf_value());
  }
  const int gi = cs.bix[bx * 3] * NC + i - 1, gj = cs.bix[bx * 3 + 1] * NC + j - 1, gk = cs.bix[bx * 3 + 2] * NC + k - 1;
  cs.v0[gi + cs.nx[0] * (gj + cs.nx[1] * gk)] = t;
}

// out = (M applied along dimension d) in, M = Q^T (trans = 1) or Q (trans = 0); optional scaling of
// the result by inv_eig (fused into the last forward transform)
__device__ __forceinline__ void cs_apply_cell(const CoarseCtx& cs, const double* in, double* out, int d, int trans,
                                              int scale, int n);
__global__ void k_cs_apply(CoarseCtx cs, const double* in, double* out, int d, int trans, int scale) {
  pdl_wait();
  cs_apply_cell(cs, in, out, d, trans, scale, blockIdx.x * blockDim.x + threadIdx.x);
}
__device__ __forceinline__ void cs_apply_cell(const CoarseCtx& cs, const double* in, double* out, int d, int trans,
                                              int scale, int n) {
  const int ntot = cs.nx[0] * cs.nx[1] * cs.nx[2];
  if (n >= ntot) return;
  int q[3] = {n % cs.nx[0], (n / cs.nx[0]) % cs.nx[1], n / (cs.nx[0] * cs.nx[1])};
  const int stride = (d == 0) ? 1 : (d == 1 ? cs.nx[0] : cs.nx[0] * cs.nx[1]);
  const int nd = cs.nx[d], o = q[d];
  const double* Q = cs.Q[d];
  const int base = n - o * stride;
  double s = 0.0;
  for (int p = 0; p < nd; ++p) {
    const double mval = trans ? Q[p * nd + o] : Q[o * nd + p];
    s = s + mval * in[base + p * stride];
  }
  if (scale) s = s * cs.inv_eig[n];
  out[n] = s;
}

// general path: x = A^-1 b with the dense inverse computed on the host at set-up; one warp per row
__global__ void k_cs_dense(CoarseCtx cs, const double* in, double* out) {
  pdl_wait();
  const int n = cs.nx[0] * cs.nx[1] * cs.nx[2];
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= n) return;
  const double* a = cs.Ainv + (size_t)row * n;
  double s = 0.0;
  for (int c = lane; c < n; c += 32) s = s + a[c] * in[c];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s = s + __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[row] = s;
}

// ---- block-tridiagonal direct solve for coarse grids too large for a dense inverse ------------------------------
// A = blocktridiag(L_p, D_p, U_p) over planes p of m cells (L, U diagonal: the couplings across the planes).  Set-up
// (once per operator): S_0 = D_0, S_p = D_p - L_p S_{p-1}^{-1} U_{p-1}, all S_p^{-1} stored dense.  Solve:
//   forward   w_p = S_p^{-1} (b_p - lo_p * w_{p-1})           backward  x_p = w_p - S_p^{-1} (up_p * x_{p+1})
// i.e. 2 np - 1 dense mat-vecs of m x m.  Exact (a block LU without pivoting of a diagonally dominant M-matrix),
// like the banded LU of the oracle and the tolerance -> 0 limit of the reference's PFMG.
//
// S_p[i][j] from the BC-folded 7-point stencils (cellst: [n][7], global cell order) and the previous inverse
__global__ void k_cs_plane_assemble(CoarseCtx cs, const double* cellst, const int* periodic, int p, double* S,
                                    const double* Sinv_prev) {
  const int m = cs.m;
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (size_t)m * m) return;
  const int i = (int)(e / m), j = (int)(e % m);
  const int gi = cs_global_of(cs, p, i), gj = cs_global_of(cs, p, j);
  const double* st = cellst + (size_t)7 * gi;
  double v = 0.0;
  if (i == j) v = st[0];
  const int q[3] = {gi % cs.nx[0], (gi / cs.nx[0]) % cs.nx[1], gi / (cs.nx[0] * cs.nx[1])};
  const int gs[3] = {1, cs.nx[0], cs.nx[0] * cs.nx[1]};
  for (int f = 0; f < 6; ++f) {
    const int d = f >> 1, sgn = (f & 1) ? 1 : -1;
    if (d == cs.sdim || st[f + 1] == 0.0) continue;
    int qd = q[d] + sgn, g2 = gi + sgn * gs[d];
    if (qd < 0 || qd >= cs.nx[d]) {
      if (!periodic[d]) continue;
      g2 = gi - sgn * (cs.nx[d] - 1) * gs[d];
    }
    if (g2 == gj) v = v + st[f + 1];
  }
  if (Sinv_prev) v = v - cs.lo[gi] * Sinv_prev[e] * cs.up[cs_global_of(cs, p - 1, j)];
  S[e] = v;
}
// in-place Gauss-Jordan inversion without pivoting, pivot step p: first the pivot row and column are saved ...
__global__ void k_gj_save(const double* S, int m, int p, double* rowp, double* colp) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= m) return;
  rowp[t] = S[(size_t)p * m + t];
  colp[t] = S[(size_t)t * m + p];
}
// ... then every element is updated from them
__global__ void k_gj_update(double* S, int m, int p, const double* rowp, const double* colp) {
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (size_t)m * m) return;
  const int r = (int)(e / m), c = (int)(e % m);
  const double inv = 1.0 / rowp[p];
  if (r == p) S[e] = (c == p) ? inv : rowp[c] * inv;
  else if (c == p) S[e] = -(colp[r] * inv);
  else S[e] = S[e] - (colp[r] * inv) * rowp[c];
}
// forward / backward substitution step on plane p; one warp per row of S_p^{-1}
// mode 0: w_p = Sinv_p (b_p - lo_p * w_{p-1})       (b in global order, w / x in plane order)
// mode 1: x_p = w_p - Sinv_p (up_p * x_{p+1})       (x_{np-1} = w_{np-1} is a plain copy: mode 2)
__global__ void k_cs_plane_rhs(CoarseCtx cs, const double* b, int p, int mode) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cs.m) return;
  const int g = cs_global_of(cs, p, i);
  if (mode == 0) cs.tvec[i] = (p == 0) ? b[g] : b[g] - cs.lo[g] * cs.w[(size_t)(p - 1) * cs.m + i];
  else if (mode == 1) cs.tvec[i] = cs.up[g] * cs.xs[(size_t)(p + 1) * cs.m + i];
  else cs.xs[(size_t)p * cs.m + i] = cs.w[(size_t)p * cs.m + i];
}
__global__ void k_cs_plane_matvec(CoarseCtx cs, int p, int mode) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= cs.m) return;
  const double* a = cs.Sinv + ((size_t)p * cs.m + row) * cs.m;
  double s = 0.0;
  for (int c = lane; c < cs.m; c += 32) s = s + a[c] * cs.tvec[c];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s = s + __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) {
    if (mode == 0) cs.w[(size_t)p * cs.m + row] = s;
    else cs.xs[(size_t)p * cs.m + row] = cs.w[(size_t)p * cs.m + row] - s;
  }
}
// x (plane order) -> out (global order), the layout k_cs_scatter reads
__global__ void k_cs_plane_out(CoarseCtx cs, double* out) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= cs.nx[0] * cs.nx[1] * cs.nx[2]) return;
  int p, i;
  cs_plane_of(cs, g, p, i);
  out[g] = cs.xs[(size_t)p * cs.m + i];
}

// coarse_solver_get_phi (m_coarse_solver.f90:341-358)
template <int NC>
__device__ __forceinline__ void cs_scatter_cell(const DevCtx& cx, const CoarseCtx& cs, int nbox1, const double* x, int n);
template <int NC>
__global__ void k_cs_scatter(DevCtx cx, CoarseCtx cs, int nbox1, const double* x) {
  pdl_wait();
  cs_scatter_cell<NC>(cx, cs, nbox1, x, blockIdx.x * blockDim.x + threadIdx.x);
}
template <int NC>
__device__ __forceinline__ void cs_scatter_cell(const DevCtx& cx, const CoarseCtx& cs, int nbox1, const double* x, int n) {
  using L = Lay3<NC>;
  const int ncell = NC * NC * NC;
  if (n >= nbox1 * ncell) return;
  const int bx = n / ncell, r = n % ncell;
  const int i = r % NC + 1, j = (r / NC) % NC + 1, k = r / (NC * NC) + 1;
  const int gi = cs.bix[bx * 3] * NC + i - 1, gj = cs.bix[bx * 3 + 1] * NC + j - 1, gk = cs.bix[bx * 3 + 2] * NC + k - 1;
  double* phi = cx.cc[V_PHI] + (size_t)bx * L::BOX;
  phi[L::interior(i, j, k)] = x[gi + cs.nx[0] * (gj + cs.nx[1] * gk)];
}

// The whole separable coarse solve (and the af_gc_lvl(1) that follows it, m_af_multigrid.f90:289) in ONE CTA for
// coarse grids of up to 1024 cells (8^3, the streamer default): gather, three forward transforms, eigenvalue scaling, three backward
// transforms, scatter, ghost cells -- nine launches of ~3 us each otherwise, on the critical path of every cycle.
// Same loops and summation order as k_cs_gather / k_cs_apply / k_cs_scatter: bit-identical results.
template <int NC>
__device__ __forceinline__ void cs_fused_body(const DevCtx& cx, const CoarseCtx& cs, int nbox1, int with_gc, double* sv) {
  using L = Lay3<NC>;
  // sv: two work vectors of ntot doubles
  const int ntot = cs.nx[0] * cs.nx[1] * cs.nx[2];
  double* a = sv;
  double* b = sv + ntot;
  const int ncell = NC * NC * NC;
  for (int n = threadIdx.x; n < nbox1 * ncell; n += blockDim.x) {
    const int bx = n / ncell, r = n % ncell;
    const int i = r % NC + 1, j = (r / NC) % NC + 1, k = r / (NC * NC) + 1;
    const double* rhs = cx.cc[V_RHS] + (size_t)bx * L::BOX;
    double t = rhs[L::interior(i, j, k)];
    const int q[3] = {i, j, k};
#pragma unroll
    for (int f = 0; f < 6; ++f) {
      const int d = f >> 1;
      if (cx.nbr[bx * 6 + f] >= 0) continue;
      if (q[d] != ((f & 1) ? NC : 1)) continue;
      const int ta = (d == 0) ? 1 : 0, tb = (d == 2) ? 1 : 2;
      const int fi = (q[ta] - 1) + (q[tb] - 1) * NC;
      const int row = cx.aux[bx * 6 + f];
      t = t + cs.b2r[((size_t)bx * 6 + f) * L::NC2 + fi] * cx.rule_B[(size_t)row * L::NC2 + fi];
    }
    if (cs.lsf_fac) {
      const double* bvp = cx.bv_of(bx);
      t = t + cs.lsf_fac[(size_t)bx * ncell + r] * (bvp ? bvp[((i + j + k) & 1) * L::NI + L::iidx((i - 1) >> 1, j, k)] : cx.lsf_value());
    }
    const int gi = cs.bix[bx * 3] * NC + i - 1, gj = cs.bix[bx * 3 + 1] * NC + j - 1, gk = cs.bix[bx * 3 + 2] * NC + k - 1;
    a[gi + cs.nx[0] * (gj + cs.nx[1] * gk)] = t;
  }
  __syncthreads();
  for (int pass = 0; pass < 6; ++pass) {
    const int d = pass % 3, trans = pass < 3, scale = pass == 2;
    const int stride = (d == 0) ? 1 : (d == 1 ? cs.nx[0] : cs.nx[0] * cs.nx[1]);
    const int nd = cs.nx[d];
    const double* Q = cs.Q[d];
    for (int n = threadIdx.x; n < ntot; n += blockDim.x) {
      const int o = (d == 0) ? n % cs.nx[0] : (d == 1 ? (n / cs.nx[0]) % cs.nx[1] : n / (cs.nx[0] * cs.nx[1]));
      const int base = n - o * stride;
      double s = 0.0;
      for (int p = 0; p < nd; ++p) {
        const double mval = trans ? __ldg(Q + p * nd + o) : __ldg(Q + o * nd + p);
        s = s + mval * a[base + p * stride];
      }
      if (scale) s = s * cs.inv_eig[n];
      b[n] = s;
    }
    __syncthreads();
    double* tsw = a;
    a = b;
    b = tsw;
  }
  for (int n = threadIdx.x; n < nbox1 * ncell; n += blockDim.x) {
    const int bx = n / ncell, r = n % ncell;
    const int i = r % NC + 1, j = (r / NC) % NC + 1, k = r / (NC * NC) + 1;
    const int gi = cs.bix[bx * 3] * NC + i - 1, gj = cs.bix[bx * 3 + 1] * NC + j - 1, gk = cs.bix[bx * 3 + 2] * NC + k - 1;
    double* phi = cx.cc[V_PHI] + (size_t)bx * L::BOX;
    phi[L::interior(i, j, k)] = a[gi + cs.nx[0] * (gj + cs.nx[1] * gk)];
  }
  if (!with_gc) return;
  // af_gc_lvl(1): sides of every level-1 box first, then edges and corners (they read the neighbours' interiors
  // and the box's own face ghosts, all written by this CTA)
  __syncthreads();
  for (int bx = 0; bx < nbox1; ++bx) gc_sides<NC>(cx, bx, V_PHI);
  __syncthreads();
  for (int bx = 0; bx < nbox1; ++bx) gc_edges_corners<NC>(cx, bx, V_PHI);
}

template <int NC>
__global__ void __launch_bounds__(1024) k_cs_fused(DevCtx cx, CoarseCtx cs, int nbox1, int with_gc) {
  pdl_wait();
  extern __shared__ __align__(16) double sv[];
  cs_fused_body<NC>(cx, cs, nbox1, with_gc, sv);
}

}  // namespace afmg

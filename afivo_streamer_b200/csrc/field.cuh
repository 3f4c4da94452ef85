// Field from potential on the device: mg_compute_phi_gradient, mg_box_lpl_gradient,
// mg_box_lpllsf_gradient, mg_box_field_norm (afivo/src/m_af_multigrid.f90:1857-2137) and the af_gc_tree of
// the field norm that field_from_potential issues afterwards (src/m_field.f90:531-548; cc_methods of
// i_electric_fld = af_bc_neumann_zero + af_gc_interp, src/m_field.f90:392-393; af_gc_interp
// m_af_ghostcell.f90:394-498).  fp64, -fmad=false: bit-identical to the CPU oracle.
//
// The face-centred field keeps the reference's own record: fc(nc+1, nc+1[, nc+1], NDIM) per box, first index
// fastest (m_af_core.f90:552).  The field norm is one more cell-centred variable in the box layout of
// layout.cuh (3D) / the reference order (2D).
#pragma once
#include "kernels2d.cuh"
#include "kernels3d.cuh"

namespace afmg {

#define AFMG_MAX_LVL 31  // lvls(1:30) in the reference (m_af_types.f90:13)

struct FieldCtx {
  double* fc;                 // [nslots * ND*(NC+1)^ND]
  double* fld;                // [nslots * BOX] field norm (i_norm)
  const double* eps;          // [nslots * BOX] tree%mg_i_eps, or null
  const unsigned char* veps;  // [nslots] 1: iand(box%tag, operator_mask) == mg_veps_box; may be null
  const double* bc_c;         // [nbc*3]   boundary rule of the field norm per physical face (c0, c1, c2)
  const double* bc_B;         // [nbc*NF]  its boundary values (NF = nc^2 in 3D, nc in 2D)
  const double* lsf_value_p;  // mg%lsf_boundary_value (device memory, see DevCtx::lsf_value_p)
  __device__ __forceinline__ double lsf_value() const { return *lsf_value_p; }
  double inv_dr[AFMG_MAX_LVL][3];  // fac / box%dr per level
};

// ---------------------------------------------------------------------------------------------
// 3D
// ---------------------------------------------------------------------------------------------
template <int NC>
struct Fc3 {
  static constexpr int N1 = NC + 1, PER = N1 * N1 * N1, LEN = 3 * PER;
  static AFMG_HD int at(int i, int j, int k, int d) { return d * PER + (i - 1) + N1 * ((j - 1) + N1 * (k - 1)); }
};

// mg_box_lpl_gradient (+ mg_box_field_norm) for boxes [slot0, slot0+nbox): one CTA per box.  phi (interior and
// face ghost cells) is first copied into shared memory in the reference's natural order P(0:nc+1)^3, so that the
// three difference passes and the norm index it with plain strides and fc leaves in its own record order
// (rows of nc+1 / nc consecutive doubles).  Boundary faces of variable-eps boxes (:1938-1997) read eps from
// global memory (few boxes).
template <int NC>
__global__ void __launch_bounds__(256) k_grad3(DevCtx cx, FieldCtx fx, int slot0, int nbox, int with_norm, int only_veps) {
  using L = Lay3<NC>;
  using F = Fc3<NC>;
  constexpr int BOX = L::BOX, N1 = NC + 1, N2 = NC + 2;
  extern __shared__ __align__(16) double P[];  // N2^3, natural order (i fastest); edges / corners unused
  const int slot = slot0 + blockIdx.x;
  const int t = threadIdx.x;
  if (only_veps && !(fx.veps && fx.veps[slot])) return;  // the other boxes have been done by k_grad3p
  const double* phi = cx.cc[V_PHI] + (size_t)slot * BOX;
  for (int n = t; n < NC * NC * NC; n += 256) {
    const int i = n % NC + 1, j = (n / NC) % NC + 1, k = n / (NC * NC) + 1;
    P[(k * N2 + j) * N2 + i] = phi[L::interior(i, j, k)];
  }
  for (int n = t; n < 6 * NC * NC; n += 256) {
    const int f = n / (NC * NC), a = n % NC + 1, b = (n / NC) % NC + 1;
    const int g = (f & 1) ? N1 : 0;
    const int i = (f < 2) ? g : a, j = (f < 2) ? a : (f < 4 ? g : b), k = (f < 4) ? b : g;
    P[(k * N2 + j) * N2 + i] = phi[L::face(f, a, b)];
  }
  const int lv = cx.lvl[slot];
  const double idr0 = fx.inv_dr[lv][0], idr1 = fx.inv_dr[lv][1], idr2 = fx.inv_dr[lv][2];
  const double* E = (fx.veps && fx.veps[slot]) ? fx.eps + (size_t)slot * BOX : nullptr;
  double* fcb = fx.fc + (size_t)slot * F::LEN;
  double* out = fx.fld + (size_t)slot * BOX;
  __syncthreads();
  // thread (i, j, ks) walks k: the centre value and its z neighbours slide through registers, x / y neighbours
  // come from shared memory; every face value is inv_dr(d) * (phi(high cell) - phi(low cell)), the eps-weighted
  // form on the boundary faces of a variable-eps box
  constexpr int KS = (256 / (NC * NC) < NC) ? ((256 / (NC * NC) > 0) ? 256 / (NC * NC) : 1) : NC;
  constexpr int KL = NC / KS;
  // Entries of the record outside a component's index range (e.g. fc(nc+1, j, k, 2)) are never assigned by the
  // reference and keep the zero af_init_box gave them (m_af_core.f90:555).  Writing that zero here completes
  // every 32-byte sector of the record, which spares the memory system a read-modify-write per hole.
  {
    constexpr int NA = N1 * N1, NH = NA + NC * N1;
    for (int n = t; n < 3 * NH; n += 256) {
      const int d = n / NH;
      int r = n - d * NH;
      const int o1 = (d == 0) ? 1 : 0, o2 = (d == 2) ? 1 : 2;
      int q[3];
      if (r < NA) {
        q[o1] = N1;
        q[d] = r % N1 + 1;
        q[o2] = r / N1 + 1;
      } else {
        r -= NA;
        q[o2] = N1;
        q[d] = r % N1 + 1;
        q[o1] = r / N1 + 1;
      }
      fcb[F::at(q[0], q[1], q[2], d)] = 0.0;
    }
  }
  if (t >= NC * NC * KS) return;
  const int i = t % NC + 1, j = (t / NC) % NC + 1, ks = t / (NC * NC);
  auto face = [&](int d, double idr, double hi, double lo, int ih, int jh, int kh) -> double {
    if (E) {  // (ih, jh, kh) = high cell of the face; pos = its index along d
      const int pos = (d == 0) ? ih : (d == 1 ? jh : kh);
      if (pos == 1 || pos == N1) {
        const double eh = E[L::cell(ih, jh, kh)];
        const double el = E[L::cell(ih - (d == 0), jh - (d == 1), kh - (d == 2))];
        const double eg = (pos == 1) ? el : eh;
        return 2 * idr * (hi - lo) * eg / (eh + el);
      }
    }
    return idr * (hi - lo);
  };
  const int k0 = ks * KL + 1;
  const double* Pc = P + (k0 * N2 + j) * N2 + i;
  double zm = Pc[-N2 * N2], c = Pc[0];
#pragma unroll 4
  for (int kk = 0; kk < KL; ++kk) {
    const int k = k0 + kk;
    const double xm = Pc[-1], xp = Pc[1], ym = Pc[-N2], yp = Pc[N2], zp = Pc[N2 * N2];
    const double fxl = face(0, idr0, c, xm, i, j, k), fxh = face(0, idr0, xp, c, i + 1, j, k);
    const double fyl = face(1, idr1, c, ym, i, j, k), fyh = face(1, idr1, yp, c, i, j + 1, k);
    const double fzl = face(2, idr2, c, zm, i, j, k), fzh = face(2, idr2, zp, c, i, j, k + 1);
    fcb[F::at(i, j, k, 0)] = fxl;
    fcb[F::at(i, j, k, 1)] = fyl;
    fcb[F::at(i, j, k, 2)] = fzl;
    if (i == NC) fcb[F::at(N1, j, k, 0)] = fxh;
    if (j == NC) fcb[F::at(i, N1, k, 1)] = fyh;
    if (k == NC) fcb[F::at(i, j, N1, 2)] = fzh;
    if (with_norm) {
      const double a = fxl + fxh, b = fyl + fyh, cc = fzl + fzh;
      out[L::interior(i, j, k)] = 0.5 * sqrt(a * a + b * b + cc * cc);
    }
    zm = c;
    c = zp;
    Pc += N2 * N2;
  }
}

// k_grad3p: the same operation as k_grad3 for the boxes WITHOUT eps-weighted boundary faces (all boxes of a tree
// without dielectrics), organised for bandwidth: this kernel moves 45 B per cell (8 (nc+2)^3 / nc^3 read, 3 fc
// components and the norm written) and is the second-largest consumer of a time step after the multigrid cycles.
//   * persistent CTAs (as many as are resident), boxes taken grid-stride;
//   * the record's two colour blocks arrive by ONE TMA bulk copy (2 COL doubles), and the copy of the NEXT box is
//     issued as soon as the current one has been unpacked into the natural-order array P, so loads overlap the
//     arithmetic and the stores of the current box;
//   * the fc record is written in its own linear order, consecutive threads -> consecutive doubles over the whole
//     component (entries outside a component's range are the zeros af_init_box left there, m_af_core.f90:555), and
//     the norm in the linear order of its colour blocks: every store instruction covers whole 32-byte sectors.
// Same expressions as k_grad3 (face value = inv_dr * (hi - lo); norm = 0.5 sqrt(a^2 + b^2 + c^2)): identical bits.
template <int NC>
struct Grad3Cfg {
  static constexpr int THREADS = (NC == 16) ? 512 : 128;
  static constexpr size_t SMEM = ((size_t)2 * Lay3<NC>::COL + (size_t)(NC + 2) * (NC + 2) * (NC + 2)) * sizeof(double);
};

template <int NC>
__global__ void __launch_bounds__(Grad3Cfg<NC>::THREADS) k_grad3p(DevCtx cx, FieldCtx fx, int slot0, int nbox, int with_norm) {
  using L = Lay3<NC>;
  using F = Fc3<NC>;
  constexpr int THREADS = Grad3Cfg<NC>::THREADS;
  constexpr int BOX = L::BOX, COL = L::COL, NI = L::NI, NF = L::NF, H = L::H, N1 = NC + 1, N2 = NC + 2, PER = F::PER;
  constexpr int NIT = (PER + THREADS - 1) / THREADS;  // fc entries of one component per thread
  constexpr int CIT = NI / THREADS;                   // interior cells of one colour per thread
  constexpr int KSTEP = THREADS / (H * NC);           // k advances by this much from one of them to the next
  static_assert(NI % THREADS == 0 && THREADS % (H * NC) == 0 && KSTEP % 2 == 0, "cell <-> thread map of the interior");
  extern __shared__ __align__(128) double smem[];
  double* const raw = smem;          // the record's colour blocks as they lie in global memory
  double* const P = smem + 2 * COL;  // N2^3, natural order (i fastest); edges / corners unused
  __shared__ uint64_t bar;
  const int t = threadIdx.x;
  if (t == 0) mbar_init(&bar, 1);
  __syncthreads();
  // ---- index maps, the same for every box: computed once, kept in registers (the loops below are fully unrolled).
  // fc entry n = t + it * THREADS of a component <-> cell (i, j, k) of the (nc+1)^3 record: offset of P(i, j, k) and,
  // per component, whether the entry exists (bit it + 10 d)
  int cidx[NIT];
  unsigned vmask = 0;
#pragma unroll
  for (int it = 0; it < NIT; ++it) {
    const int n = t + it * THREADS;
    const int i = n % N1 + 1, r = n / N1, j = r % N1 + 1, k = r / N1 + 1;
    cidx[it] = (k * N2 + j) * N2 + i;
    if (n < PER) {
      if (j <= NC && k <= NC) vmask |= 1u << it;
      if (i <= NC && k <= NC) vmask |= 1u << (it + 10);
      if (i <= NC && j <= NC) vmask |= 1u << (it + 20);
    }
  }
  static_assert(NIT <= 10, "validity bits");
  // interior cell r = t + it * THREADS of colour block c: (m, j) fixed, k = k0 + KSTEP * it, and since KSTEP is even
  // the x index depends on the colour only
  const int cm = t % H, cj = (t / H) % NC + 1, ck0 = t / (H * NC) + 1;
  const int pbase0 = (ck0 * N2 + cj) * N2 + 2 * cm + 2 - ((0 + cj + ck0) & 1);
  const int pbase1 = (ck0 * N2 + cj) * N2 + 2 * cm + 2 - ((1 + cj + ck0) & 1);
  const double* const phi = cx.cc[V_PHI];
  int b = blockIdx.x;
  if (t == 0 && b < nbox) {
    mbar_expect_tx(&bar, (uint32_t)(2 * COL * 8));
    bulk_g2s(raw, phi + (size_t)(slot0 + b) * BOX, 2 * COL * 8, &bar);
  }
  uint32_t par = 0;
  for (; b < nbox; b += gridDim.x) {
    const int slot = slot0 + b;
    const bool veps = fx.veps && fx.veps[slot];
    const int lv = cx.lvl[slot];
    const double idr0 = fx.inv_dr[lv][0], idr1 = fx.inv_dr[lv][1], idr2 = fx.inv_dr[lv][2];
    mbar_wait(&bar, par);
    par ^= 1u;
    // ---- unpack: interior of both colours, then the six ghost faces of both colours
#pragma unroll
    for (int it = 0; it < CIT; ++it) {
      P[pbase0 + it * KSTEP * N2 * N2] = raw[t + it * THREADS];
      P[pbase1 + it * KSTEP * N2 * N2] = raw[COL + t + it * THREADS];
    }
    for (int n = t; n < 12 * NF; n += THREADS) {
      const int c = n / (6 * NF), r = n - c * 6 * NF;
      int i, j, k;
      L::uncell(c * COL + NI + r, i, j, k);
      P[(k * N2 + j) * N2 + i] = raw[c * COL + NI + r];
    }
    __syncthreads();
    const int nb = b + gridDim.x;
    if (t == 0 && nb < nbox) {  // raw is free: fetch the next record while this one is worked on
      fence_async_smem();
      mbar_expect_tx(&bar, (uint32_t)(2 * COL * 8));
      bulk_g2s(raw, phi + (size_t)(slot0 + nb) * BOX, 2 * COL * 8, &bar);
    }
    if (!veps) {
      double* const fcb = fx.fc + (size_t)slot * F::LEN + t;
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const double idr = (d == 0) ? idr0 : (d == 1 ? idr1 : idr2);
        const int st = (d == 0) ? 1 : (d == 1 ? N2 : N2 * N2);
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
          if (it * THREADS + t >= PER) break;  // only the last iteration can be partial
          double v = 0.0;
          if ((vmask >> (it + 10 * d)) & 1u) v = idr * (P[cidx[it]] - P[cidx[it] - st]);
          fcb[d * PER + it * THREADS] = v;
        }
      }
      if (with_norm) {
        double* const out = fx.fld + (size_t)slot * BOX + t;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
#pragma unroll
          for (int it = 0; it < CIT; ++it) {
            const double* Pc = P + (c ? pbase1 : pbase0) + it * KSTEP * N2 * N2;
            const double ctr = Pc[0];
            const double fxl = idr0 * (ctr - Pc[-1]), fxh = idr0 * (Pc[1] - ctr);
            const double fyl = idr1 * (ctr - Pc[-N2]), fyh = idr1 * (Pc[N2] - ctr);
            const double fzl = idr2 * (ctr - Pc[-N2 * N2]), fzh = idr2 * (Pc[N2 * N2] - ctr);
            const double a = fxl + fxh, bb = fyl + fyh, cc = fzl + fzh;
            out[c * COL + it * THREADS] = 0.5 * sqrt(a * a + bb * bb + cc * cc);
          }
        }
      }
    }
    __syncthreads();  // P is rewritten by the next box
  }
}

// mg_box_field_norm (:2023-2051) from the stored fc, boxes [slot0, slot0+nbox)
template <int NC>
__device__ __forceinline__ void norm_from_fc3(const double* fcb, double* out, int t, int nt) {
  using L = Lay3<NC>;
  using F = Fc3<NC>;
  for (int n = t; n < 2 * L::NI; n += nt) {
    const int q = (n < L::NI) ? n : L::COL + (n - L::NI);
    int i, j, k;
    L::uncell(q, i, j, k);
    const double a = fcb[F::at(i, j, k, 0)] + fcb[F::at(i + 1, j, k, 0)];
    const double b = fcb[F::at(i, j, k, 1)] + fcb[F::at(i, j + 1, k, 1)];
    const double c = fcb[F::at(i, j, k, 2)] + fcb[F::at(i, j, k + 1, 2)];
    out[q] = 0.5 * sqrt(a * a + b * b + c * c);
  }
}
template <int NC>
__global__ void __launch_bounds__(256) k_norm3(FieldCtx fx, int slot0, int nbox) {
  const int slot = slot0 + blockIdx.x;
  norm_from_fc3<NC>(fx.fc + (size_t)slot * Fc3<NC>::LEN, fx.fld + (size_t)slot * Lay3<NC>::BOX, threadIdx.x, 256);
}

// mg_box_lpllsf_gradient (:2055-2137) on the leaves that hold a level-set boundary: entries e of box b =
// the box's sparse distance stencil (cells in IJK order with any(dd < 1)).  The reference's sequential loop
// lets the low-side write of cell i+1 win over the high-side write of cell i on their common face; two
// phases separated by a barrier give the same result in parallel.  Then the norm of the box is redone.
template <int NC>
__global__ void __launch_bounds__(256) k_lsf_fix3(DevCtx cx, FieldCtx fx, const int* slots, const int* eoff,
                                                  const int* ecell, const double* edd, const double* elsf, int with_norm) {
  using L = Lay3<NC>;
  using F = Fc3<NC>;
  const int slot = slots[blockIdx.x];
  const int e0 = eoff[blockIdx.x], e1 = eoff[blockIdx.x + 1];
  const int lv = cx.lvl[slot];
  const double* phi = cx.cc[V_PHI] + (size_t)slot * L::BOX;
  double* fcb = fx.fc + (size_t)slot * F::LEN;
  const double* bvp = cx.bv_of(slot);
  for (int phase = 0; phase < 2; ++phase) {
    for (int n = threadIdx.x; n < 3 * (e1 - e0); n += 256) {
      const int e = e0 + n / 3, d = n % 3;
      if (!(elsf[e] >= 0)) continue;
      int q[3] = {ecell[3 * e], ecell[3 * e + 1], ecell[3 * e + 2]};
      const double p = phi[L::interior(q[0], q[1], q[2])];
      const double bc = bvp ? bvp[((q[0] + q[1] + q[2]) & 1) * L::NI + L::iidx((q[0] - 1) >> 1, q[1], q[2])] : fx.lsf_value();
      const double idr = fx.inv_dr[lv][d];
      if (phase == 0) {
        const double dd = edd[6 * e + 2 * d + 1];
        q[d] += 1;
        if (dd < 1) fcb[F::at(q[0], q[1], q[2], d)] = idr * (bc - p) / dd;
      } else {
        const double dd = edd[6 * e + 2 * d];
        if (dd < 1) fcb[F::at(q[0], q[1], q[2], d)] = idr * (p - bc) / dd;
      }
    }
    __syncthreads();
  }
  if (with_norm) norm_from_fc3<NC>(fcb, fx.fld + (size_t)slot * L::BOX, threadIdx.x, 256);
}

// af_gc_box for the field norm (variable V_FLD): neighbour copy, bc_to_gc with the variable's own boundary rule,
// af_gc_interp on refinement boundaries; then edges and corners.  Neighbours and the parent's neighbour may live
// on a peer GPU (cx.at).
template <int NC>
__global__ void __launch_bounds__(256) k_gc_fld3(DevCtx cx, FieldCtx fx, int slot0, int nbox, int corners) {
  using L = Lay3<NC>;
  constexpr int H = L::H;
  const int slot = slot0 + blockIdx.x;
  double* box = cx.cc[V_FLD] + (size_t)slot * L::BOX;
  const double third = 1 / 3.0, sixth = 1 / 6.0;
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
    if (nb >= 0) {
      q[d] = hi ? 1 : NC;
      v = ldcell<NC>(cx.at<L::BOX>(V_FLD, nb), q[0], q[1], q[2]);
    } else {
      const int row = cx.aux[slot * 6 + f];
      if (row < cx.rb_row0) {  // physical boundary: bc_to_gc
        const double* rc = fx.bc_c + 3 * row;
        const double B = fx.bc_B[(size_t)row * L::NC2 + rr];
        q[d] = hi ? NC : 1;
        const double x1 = ldcell<NC>(box, q[0], q[1], q[2]);
        q[d] = hi ? NC - 1 : 2;
        const double x2 = ldcell<NC>(box, q[0], q[1], q[2]);
        v = (rc[0] * B + rc[1] * x1) + rc[2] * x2;
      } else {  // af_gc_interp
        const int p = cx.parent[slot], cof = cx.coff[slot];
        const double* P = cx.at<L::BOX>(V_FLD, cx.nbr[p * 6 + f]);
        const int a1 = ((cof >> ta) & 1) * H + ((a + 1) >> 1), a2 = a1 + 1 - 2 * (a & 1);
        const int b1 = ((cof >> tb) & 1) * H + ((b + 1) >> 1), b2 = b1 + 1 - 2 * (b & 1);
        int c[3];
        c[d] = hi ? 1 : NC;
        c[ta] = a1;
        c[tb] = b1;
        const double v11 = ldcell<NC>(P, c[0], c[1], c[2]);
        c[ta] = a2;
        const double v21 = ldcell<NC>(P, c[0], c[1], c[2]);
        c[ta] = a1;
        c[tb] = b2;
        const double v12 = ldcell<NC>(P, c[0], c[1], c[2]);
        q[d] = hi ? NC : 1;
        const double vf = ldcell<NC>(box, q[0], q[1], q[2]);
        // order of the two sixth terms in the reference: dims 1, 2: (c2,c1) then (c1,c2); dim 3: (c1,c2) first
        v = (d == 2) ? (third * v11 + sixth * v12 + sixth * v21 + third * vf)
                     : (third * v11 + sixth * v21 + sixth * v12 + third * vf);
      }
    }
    box[L::face(f, a, b)] = v;
  }
  if (corners) {
    __syncthreads();
    gc_edges_corners<NC>(cx, slot, V_FLD);
  }
}

// dst = dst - c * src on the full records of the leaves (photoi_helmh_compute, src/m_photoi_helmh.f90:192-201)
__global__ void k_axpy_leaves(double* dst, const double* src, const int* child0, double c, int box_len) {
  const int slot = blockIdx.x;
  if (child0[slot] >= 0) return;
  double* a = dst + (size_t)slot * box_len;
  const double* b = src + (size_t)slot * box_len;
  for (int q = threadIdx.x; q < box_len; q += blockDim.x) a[q] = a[q] - c * b[q];
}

// plain records (fc) to / from a packed buffer in box order
__global__ void k_rec_copy(double* base, const int* slots, int n, double* packed, int rec_len, int to_device) {
  const int s = slots[blockIdx.x];
  if (s < 0) return;
  double* a = base + (size_t)s * rec_len;
  double* b = packed + (size_t)blockIdx.x * rec_len;
  for (int q = threadIdx.x; q < rec_len; q += blockDim.x) {
    if (to_device) a[q] = b[q];
    else b[q] = a[q];
  }
}

}  // namespace afmg

// ---------------------------------------------------------------------------------------------
// 2D (box records in the reference order cc(0:nc+1, 0:nc+1))
// ---------------------------------------------------------------------------------------------
namespace afmg2 {

template <int NC>
struct Fc2 {
  static constexpr int N1 = NC + 1, PER = N1 * N1, LEN = 2 * PER;
  __host__ __device__ static int at(int i, int j, int d) { return d * PER + (i - 1) + N1 * (j - 1); }
};

template <int NC>
__device__ __forceinline__ double face_val2(const double* S, const double* E, double idr, int d, int i, int j) {
  using B = B2<NC>;
  const int qh = B::at(i, j), ql = (d == 0) ? B::at(i - 1, j) : B::at(i, j - 1);
  const double hi = S[qh], lo = S[ql];
  if (E) {
    const int pos = (d == 0) ? i : j;
    if (pos == 1 || pos == NC + 1) {
      const double eh = E[qh], el = E[ql];
      const double eg = (pos == 1) ? el : eh;
      return 2 * idr * (hi - lo) * eg / (eh + el);
    }
  }
  return idr * (hi - lo);
}

template <int NC>
__device__ __forceinline__ void norm_from_fc2(const double* fcb, double* out, int t, int nt) {
  using B = B2<NC>;
  using F = Fc2<NC>;
  for (int n = t; n < NC * NC; n += nt) {
    const int i = n % NC + 1, j = n / NC + 1;
    const double a = fcb[F::at(i, j, 0)] + fcb[F::at(i + 1, j, 0)];
    const double b = fcb[F::at(i, j, 1)] + fcb[F::at(i, j + 1, 1)];
    out[B::at(i, j)] = 0.5 * sqrt(a * a + b * b);
  }
}

template <int NC>
__global__ void k2_grad(Ctx cx, afmg::FieldCtx fx, int slot0, int nbox, int with_norm) {
  using B = B2<NC>;
  using F = Fc2<NC>;
  constexpr int N1 = NC + 1;
  const int slot = slot0 + blockIdx.x;
  const double* S = cx.cc[V_PHI] + (size_t)slot * B::BOX;
  const int lv = cx.lvl[slot];
  const double* E = (fx.veps && fx.veps[slot]) ? fx.eps + (size_t)slot * B::BOX : nullptr;
  double* fcb = fx.fc + (size_t)slot * F::LEN;
  for (int d = 0; d < 2; ++d) {
    const int ni = (d == 0) ? N1 : NC, nj = (d == 1) ? N1 : NC;
    const double idr = fx.inv_dr[lv][d];
    for (int n = threadIdx.x; n < ni * nj; n += blockDim.x) {
      const int i = n % ni + 1, j = n / ni + 1;
      fcb[F::at(i, j, d)] = face_val2<NC>(S, E, idr, d, i, j);
    }
  }
  if (!with_norm) return;
  __syncthreads();
  norm_from_fc2<NC>(fcb, fx.fld + (size_t)slot * B::BOX, threadIdx.x, blockDim.x);
}

template <int NC>
__global__ void k2_norm(afmg::FieldCtx fx, int slot0, int nbox) {
  const int slot = slot0 + blockIdx.x;
  norm_from_fc2<NC>(fx.fc + (size_t)slot * Fc2<NC>::LEN, fx.fld + (size_t)slot * B2<NC>::BOX, threadIdx.x, blockDim.x);
}

template <int NC>
__global__ void k2_lsf_fix(Ctx cx, afmg::FieldCtx fx, const int* slots, const int* eoff, const int* ecell, const double* edd,
                           const double* elsf, int with_norm) {
  using B = B2<NC>;
  using F = Fc2<NC>;
  const int slot = slots[blockIdx.x];
  const int e0 = eoff[blockIdx.x], e1 = eoff[blockIdx.x + 1];
  const int lv = cx.lvl[slot];
  const double* phi = cx.cc[V_PHI] + (size_t)slot * B::BOX;
  double* fcb = fx.fc + (size_t)slot * F::LEN;
  const double* bvp = (cx.bvoff && cx.bvoff[slot] >= 0) ? cx.bv + cx.bvoff[slot] : nullptr;
  for (int phase = 0; phase < 2; ++phase) {
    for (int n = threadIdx.x; n < 2 * (e1 - e0); n += blockDim.x) {
      const int e = e0 + n / 2, d = n % 2;
      if (!(elsf[e] >= 0)) continue;
      int q[2] = {ecell[2 * e], ecell[2 * e + 1]};
      const double p = phi[B::at(q[0], q[1])];
      const double bc = bvp ? bvp[(q[0] - 1) + NC * (q[1] - 1)] : fx.lsf_value();
      const double idr = fx.inv_dr[lv][d];
      if (phase == 0) {
        const double dd = edd[4 * e + 2 * d + 1];
        q[d] += 1;
        if (dd < 1) fcb[F::at(q[0], q[1], d)] = idr * (bc - p) / dd;
      } else {
        const double dd = edd[4 * e + 2 * d];
        if (dd < 1) fcb[F::at(q[0], q[1], d)] = idr * (p - bc) / dd;
      }
    }
    __syncthreads();
  }
  if (with_norm) norm_from_fc2<NC>(fcb, fx.fld + (size_t)slot * B::BOX, threadIdx.x, blockDim.x);
}

// af_gc_box of the field norm in 2D: af_gc_interp (m_af_ghostcell.f90:429-447) on refinement boundaries
template <int NC>
__global__ void k2_gc_fld(Ctx cx, afmg::FieldCtx fx, int slot0, int nbox, int corners) {
  using B = B2<NC>;
  constexpr int H = NC / 2;
  const int slot = slot0 + blockIdx.x;
  double* vb = fx.fld;
  double* box = vb + (size_t)slot * B::BOX;
  const double third = 1 / 3.0, sixth = 1 / 6.0;
  for (int n = threadIdx.x; n < 4 * NC; n += blockDim.x) {
    const int f = n / NC, a = n % NC + 1;
    const int d = f >> 1, hi = f & 1, td = 1 - d;
    int q[2];
    q[td] = a;
    const int nb = cx.nbr[slot * 4 + f];
    double v;
    if (nb >= 0) {
      q[d] = hi ? 1 : NC;
      v = vb[(size_t)nb * B::BOX + B::at(q[0], q[1])];
    } else {
      const int row = cx.aux[slot * 4 + f];
      if (row < cx.rb_row0) {
        const double* rc = fx.bc_c + 3 * row;
        const double Bv = fx.bc_B[(size_t)row * NC + (a - 1)];
        q[d] = hi ? NC : 1;
        const double x1 = box[B::at(q[0], q[1])];
        q[d] = hi ? NC - 1 : 2;
        const double x2 = box[B::at(q[0], q[1])];
        v = rc[0] * Bv + rc[1] * x1 + rc[2] * x2;
      } else {
        const int p = cx.parent[slot], cof = cx.coff[slot];
        const double* P = vb + (size_t)cx.nbr[p * 4 + f] * B::BOX;
        const int a1 = ((cof >> td) & 1) * H + ((a + 1) >> 1), a2 = a1 + 1 - 2 * (a & 1);
        int c[2];
        c[d] = hi ? 1 : NC;
        c[td] = a1;
        const double v1 = P[B::at(c[0], c[1])];
        c[td] = a2;
        const double v2 = P[B::at(c[0], c[1])];
        q[d] = hi ? NC : 1;
        v = 0.5 * v1 + sixth * v2 + third * box[B::at(q[0], q[1])];
      }
    }
    q[d] = hi ? NC + 1 : 0;
    box[B::at(q[0], q[1])] = v;
  }
  if (!corners) return;
  __syncthreads();
  if (threadIdx.x < 4) {
    const int c = threadIdx.x;
    const int dx = (c & 1) ? 1 : -1, dy = (c & 2) ? 1 : -1;
    const int qi = (c & 1) ? NC + 1 : 0, qj = (c & 2) ? NC + 1 : 0;
    const int nb = cx.nmat[slot * 9 + (dx + 1) + 3 * (dy + 1)];
    double v;
    if (nb >= 0) v = vb[(size_t)nb * B::BOX + B::at(qi - dx * NC, qj - dy * NC)];
    else v = box[B::at(qi - dx, qj)] + box[B::at(qi, qj - dy)] - box[B::at(qi - dx, qj - dy)];
    box[B::at(qi, qj)] = v;
  }
}

}  // namespace afmg2

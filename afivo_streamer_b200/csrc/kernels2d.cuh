// sm_100a kernels of the 2D (Cartesian and cylindrical) FAS multigrid path.
// 2D boxes are tiny (nc^2 = 64 cells for the streamer default nc = 8), the whole 2D tree of the
// reference's CPU-runnable configuration fits in L2, and the path is launch-latency bound: the kernels
// are plain one-CTA-per-box kernels on box records in the reference's own (nc+2)^2 order (first index
// fastest), replayed as CUDA graphs.  All stencil kinds are handled by the same code: implicit constant
// Laplacian (mg_box_lpl_stencil), its cylindrical form (stencil%cylindrical_gradient, af_cyl_flux_factors
// m_af_types.f90:1199-1211), explicit constant / variable stencils and level-set boxes.
// Expression order follows afivo/src/m_af_stencil.f90 (2D branches) so that results are bit-identical
// to the CPU oracle (-fmad=false).
#pragma once
#include <cuda_runtime.h>

namespace afmg2 {

enum { V_PHI = 0, V_RHS = 1, V_TMP = 2 };

// (AFMG_PDL=1).  launch_dependents first: the NEXT kernel of the stream / graph may be scheduled as soon as every CTA of
// this one has started, so its launch latency overlaps with this kernel's execution; its own griddepcontrol.wait still
// blocks until this grid has completed and flushed.  Both instructions are no-ops without the launch attribute.
__device__ __forceinline__ void pdl_wait() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

struct Ctx {
  double* cc[3];         // per variable: nslots * (nc+2)^2 doubles, box record = cc(0:nc+1, 0:nc+1)
  const int* nbr;        // [nslots*4]  >= 0 neighbour slot; -1: own-ghost rule row in aux
  const int* aux;        // [nslots*4]
  const int* nmat;       // [nslots*9]  >= 0 slot, -1 physical boundary, -2 no box
  const int* parent;     // [nslots]
  const int* child0;     // [nslots]  first child slot or -1
  const int* coff;       // [nslots]  bit d: upper half of the parent in dim d
  const int* lvl;        // [nslots]
  const double* coef;    // [(L+1)*8]  per level: c1..c5 of the constant 5-point stencil (Cartesian form)
  const double* drx;     // [L+1]      dr(1) per level (cylindrical factors)
  const double* rmin;    // [nslots]   box%r_min(1)
  int cyl;               // tree%coord_t == af_cyl
  const double* rule_c;  // [nrules*3]
  double* rule_B;        // [nrules*nc]
  const unsigned char* rule_flag;  // [nrules] 1: mg_sides_rb_extrap face (variable-eps box), may be null
  const int* rb_slot;    // [nrb]
  const int* rb_face;    // [nrb]
  int rb_row0;
  const double* pcoef;   // [4] default prolongation coefficients
  int pshape;            // 4 = p248 (2D: 4 points), 3 = p234 (2D: 3 points)
  // explicit stencils (afmg_set_stencils), reference layout v(5, i, j) / v(3, i, j), first index fastest
  const unsigned char* opk;  // [nslots] bits 0-1: 0 implicit, 1 explicit constant, 2 variable; bit 2: cylindrical_gradient
  const long long* opoff;
  const long long* foff;     // -1: none
  const unsigned char* pk;   // [nslots] 0 default, 1 constant p248, 2 constant p234, 3 variable p234
  const long long* poff;
  const double* stv;
  const double* lsf_value_p;  // see DevCtx::lsf_value_p (kernels3d.cuh)
  __device__ __forceinline__ double lsf_value() const { return *lsf_value_p; }
  const long long* bvoff;    // [nslots] per-cell level-set boundary values (afmg_set_lsf_boundary_values), or null
  const double* bv;          // nc^2 per listed box, cell order (i fastest)
  double two_pi;         // 2 * acos(-1) as the host computes it (af_tree_sum_cc in cylindrical coordinates)
};

template <int NC>
struct B2 {
  static constexpr int N2 = NC + 2, BOX = N2 * N2, H = NC / 2;
  __host__ __device__ static int at(int i, int j) { return i + N2 * j; }
};

// effective 5-point coefficients at cell (i, j): q[0] centre, q[1..4] = -x, +x, -y, +y.
// Cylindrical form: m_af_stencil.f90:886-925 (cc_cyl), af_cyl_flux_factors m_af_types.f90:1199-1211.
template <int NC>
__device__ __forceinline__ void coefs2(const Ctx& cx, int slot, int i, int j, double* q, bool& variable, bool& cyl_form) {
  const int k = cx.opk ? cx.opk[slot] : 0;
  const int kind = k & 3;
  const double* c;
  if (kind == 0) c = cx.coef + 8 * cx.lvl[slot];
  else if (kind == 1) c = cx.stv + cx.opoff[slot];
  else c = cx.stv + cx.opoff[slot] + 5 * ((i - 1) + NC * (j - 1));
  variable = (kind == 2);
  cyl_form = (kind == 0) ? (cx.cyl != 0) : ((k & 4) != 0);
  if (cyl_form) {
    const double dr = cx.drx[cx.lvl[slot]];
    const double r = cx.rmin[slot] + (i - 0.5) * dr;
    const double inv_r = 1 / r;
    const double rf0 = (r - 0.5 * dr) * inv_r, rf1 = (r + 0.5 * dr) * inv_r;
    q[1] = rf0 * c[1];
    q[2] = rf1 * c[2];
    q[0] = c[0] - (q[1] - c[1]) - (q[2] - c[2]);
    q[3] = c[3];
    q[4] = c[4];
  } else {
#pragma unroll
    for (int m = 0; m < 5; ++m) q[m] = c[m];
  }
}

// bc_correction(i, j) = f(i, j) * lsf_boundary_value, or 0 (flag has = false)
template <int NC>
__device__ __forceinline__ double bc_corr2(const Ctx& cx, int slot, int i, int j, bool& has) {
  has = cx.opk && cx.foff[slot] >= 0;
  if (!has) return 0.0;
  const double V = (cx.bvoff && cx.bvoff[slot] >= 0) ? cx.bv[cx.bvoff[slot] + (i - 1) + NC * (j - 1)] : cx.lsf_value();
  return cx.stv[cx.foff[slot] + (i - 1) + NC * (j - 1)] * V;
}

// stencil_apply_357, 2D (m_af_stencil.f90:367-460, :490-493)
template <int NC>
__device__ __forceinline__ double apply2(const Ctx& cx, int slot, const double* box, int i, int j) {
  using B = B2<NC>;
  double q[5];
  bool var, cylf, has;
  coefs2<NC>(cx, slot, i, j, q, var, cylf);
  const int n = B::at(i, j);
  double acc = q[0] * box[n];
  acc = acc + q[1] * box[n - 1];
  acc = acc + q[2] * box[n + 1];
  acc = acc + q[3] * box[n - B::N2];
  acc = acc + q[4] * box[n + B::N2];
  const double bc = bc_corr2<NC>(cx, slot, i, j, has);
  if (has) acc = acc - bc;
  return acc;
}

// one red-black half-sweep (stencil_gsrb_357, 2D: m_af_stencil.f90:880-954); ghost cells are refreshed by
// k2_gc afterwards (gsrb_boxes, m_af_multigrid.f90:648-687)
template <int NC>
__global__ void k2_gsrb(Ctx cx, int slot0, int nbox, int C) {
  pdl_wait();
  using B = B2<NC>;
  const int slot = slot0 + blockIdx.x;
  double* phi = cx.cc[V_PHI] + (size_t)slot * B::BOX;
  double* rhs = cx.cc[V_RHS] + (size_t)slot * B::BOX;
  for (int n = threadIdx.x; n < NC * NC; n += blockDim.x) {
    const int i = n % NC + 1, j = n / NC + 1;
    const int o = B::at(i, j);
    bool has;
    const double bc = bc_corr2<NC>(cx, slot, i, j, has);
    double r = rhs[o];
    if (has) r = r + bc;
    if (((i + j) & 1) == C) {
      double q[5];
      bool var, cylf;
      coefs2<NC>(cx, slot, i, j, q, var, cylf);
      double acc = r;
      acc = acc - q[1] * phi[o - 1];
      acc = acc - q[2] * phi[o + 1];
      acc = acc - q[3] * phi[o - B::N2];
      acc = acc - q[4] * phi[o + B::N2];
      phi[o] = var ? acc / q[0] : acc * (1 / q[0]);
    }
    if (has) rhs[o] = r - bc;
  }
}

// first half of mg_sides_rb in 2D (m_af_multigrid.f90:294-369) / af_gc_prolong_copy for extrapolated faces
template <int NC>
__global__ void k2_rb_prepare(Ctx cx, int r0, int nr, int var) {
  pdl_wait();
  using B = B2<NC>;
  constexpr int H = B::H;
  const int r = r0 + blockIdx.x;
  const int s = cx.rb_slot[r], f = cx.rb_face[r];
  const int p = cx.parent[s];
  const int d = f >> 1, td = 1 - d;
  const int cof = cx.coff[s];
  const int cot = ((cof >> td) & 1) * H;
  double* out = cx.rule_B + (size_t)(cx.rb_row0 + r) * NC;
  if (cx.rule_flag && cx.rule_flag[cx.rb_row0 + r]) {
    const double* pb = cx.cc[var] + (size_t)p * B::BOX;
    const int g = (f & 1) ? NC + 1 : 0;
    const int cod = ((cof >> d) & 1) * H;
    for (int a = threadIdx.x + 1; a <= NC; a += blockDim.x) {
      int q[2];
      q[d] = cod + ((g + 1) >> 1);
      q[td] = cot + ((a + 1) >> 1);
      out[a - 1] = pb[B::at(q[0], q[1])];
    }
    return;
  }
  const int pn = cx.nbr[p * 4 + f];
  const double* cb = cx.cc[var] + (size_t)pn * B::BOX;
  const int layer = (f & 1) ? 1 : NC;
  auto T = [&](int x) {
    int q[2];
    q[d] = layer;
    q[td] = cot + x;
    return cb[B::at(q[0], q[1])];
  };
  for (int a = threadIdx.x + 1; a <= NC; a += blockDim.x) {
    const int ia = (a + 1) >> 1;
    const double t0 = T(ia);
    const double g1 = 0.125 * (T(ia + 1) - T(ia - 1));
    out[a - 1] = (a & 1) ? (t0 - g1) : (t0 + g1);
  }
}

// af_gc_box in 2D (m_af_ghostcell.f90:64-170): sides, then the four corners
template <int NC>
__global__ void k2_gc(Ctx cx, int slot0, int nbox, int var, int corners) {
  pdl_wait();
  using B = B2<NC>;
  const int slot = slot0 + blockIdx.x;
  double* vb = cx.cc[var];
  double* box = vb + (size_t)slot * B::BOX;
  for (int n = threadIdx.x; n < 4 * NC; n += blockDim.x) {
    const int f = n / NC, a = n % NC + 1;
    const int d = f >> 1, hi = f & 1, td = 1 - d;
    int q[2];
    q[td] = a;
    const int nb = cx.nbr[slot * 4 + f];
    double v;
    if (nb >= 0) {  // copy_from_nb
      q[d] = hi ? 1 : NC;
      v = vb[(size_t)nb * B::BOX + B::at(q[0], q[1])];
    } else {
      const int row = cx.aux[slot * 4 + f];
      const double* rc = cx.rule_c + 3 * row;
      const double Bv = cx.rule_B[(size_t)row * NC + (a - 1)];
      const int l1 = hi ? NC : 1, l2 = hi ? NC - 1 : 2;
      q[d] = l1;
      const double x1 = box[B::at(q[0], q[1])];
      if (cx.rule_flag && cx.rule_flag[row]) {
        // mg_sides_rb_extrap, 2D (m_af_multigrid.f90:509-512): bilinear extrapolation with 4 points
        const int da = -1 + 2 * (a & 1);
        int qn[2], qt[2], qd[2];
        qn[d] = l2; qn[td] = a;            // one cell further inside
        qt[d] = l1; qt[td] = a + da;       // transverse neighbour
        qd[d] = l2; qd[td] = a + da;       // diagonal
        const double xn = box[B::at(qn[0], qn[1])], xt = box[B::at(qt[0], qt[1])], xd = box[B::at(qd[0], qd[1])];
        // the reference adds cc(i+di, j) + cc(i, j+dj): x step first, y step second
        const double s = (d == 0) ? (xn + xt) : (xt + xn);
        v = 0.5 * Bv + 1.125 * x1 - 0.375 * s + 0.125 * xd;
      } else {
        q[d] = l2;
        const double x2 = box[B::at(q[0], q[1])];
        v = rc[0] * Bv + rc[1] * x1 + rc[2] * x2;
      }
    }
    q[d] = hi ? NC + 1 : 0;
    box[B::at(q[0], q[1])] = v;
  }
  if (!corners) return;
  __syncthreads();
  if (threadIdx.x < 4) {  // af_gc_box_corner (:155-169), af_corner_gc_extrap (:860-870)
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

// residual_box (MODE 0, + leaf max-norm) / child part of update_coarse and set_coarse_phi_rhs (MODE 1):
// residual into shared memory, then af_restrict_box in 2D (m_af_restrict.f90:84-117): phi without, the
// residual with the cylindrical child weights (af_cyl_child_weights, m_af_types.f90:1187-1196)
template <int NC, int MODE>
__global__ void k2_resid(Ctx cx, int slot0, int nbox, unsigned long long* maxabs_bits, int keep_res) {
  pdl_wait();
  using B = B2<NC>;
  constexpr int H = B::H;
  __shared__ double sres[NC * NC];
  const int slot = slot0 + blockIdx.x;
  const double* phi = cx.cc[V_PHI] + (size_t)slot * B::BOX;
  const double* rhs = cx.cc[V_RHS] + (size_t)slot * B::BOX;
  double* tmp = cx.cc[V_TMP] + (size_t)slot * B::BOX;
  double mx = 0.0;
  for (int n = threadIdx.x; n < NC * NC; n += blockDim.x) {
    const int i = n % NC + 1, j = n / NC + 1;
    const int o = B::at(i, j);
    const double res = rhs[o] - apply2<NC>(cx, slot, phi, i, j);
    if (MODE == 0 || keep_res) tmp[o] = res;
    if (MODE == 1) sres[n] = res;
    mx = fmax(mx, fabs(res));
  }
  if (MODE == 1) {
    __syncthreads();
    const int p = cx.parent[slot], cof = cx.coff[slot];
    double* ptmp = cx.cc[V_TMP] + (size_t)p * B::BOX;
    double* pphi = cx.cc[V_PHI] + (size_t)p * B::BOX;
    const int ox = (cof & 1) * H, oy = ((cof >> 1) & 1) * H;
    for (int n = threadIdx.x; n < H * H; n += blockDim.x) {
      const int ic = n % H + 1, jc = n / H + 1;
      const int i = 2 * ic - 1, j = 2 * jc - 1;
      double sp = 0.0;
      sp = sp + phi[B::at(i, j)];
      sp = sp + phi[B::at(i + 1, j)];
      sp = sp + phi[B::at(i, j + 1)];
      sp = sp + phi[B::at(i + 1, j + 1)];
      auto R = [&](int a, int b) { return sres[(a - 1) + NC * (b - 1)]; };
      double rr;
      if (cx.cyl) {
        const double drp = cx.drx[cx.lvl[p]];
        const double rc = cx.rmin[p] + (ox + ic - 0.5) * drp;
        const double t = 0.25 * drp / rc;
        const double w1 = 1 - t, w2 = 1 + t;
        double s1 = 0.0, s2 = 0.0;
        s1 = s1 + R(i, j);
        s1 = s1 + R(i, j + 1);
        s2 = s2 + R(i + 1, j);
        s2 = s2 + R(i + 1, j + 1);
        rr = 0.25 * (w1 * s1 + w2 * s2);
      } else {
        double s = 0.0;
        s = s + R(i, j);
        s = s + R(i + 1, j);
        s = s + R(i, j + 1);
        s = s + R(i + 1, j + 1);
        rr = 0.25 * s;
      }
      ptmp[B::at(ox + ic, oy + jc)] = rr;
      pphi[B::at(ox + ic, oy + jc)] = 0.25 * sp;
    }
  }
  if (MODE == 0 && maxabs_bits && cx.child0[slot] < 0) {
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0 && mx > 0.0) atomicMax(maxabs_bits, (unsigned long long)__double_as_longlong(mx));
  }
}

// parent part of update_coarse (m_af_multigrid.f90:724-737): rhs = L(phi) + tmp (interior); mode 1: tmp = phi
// (full record); mode 2 (set_coarse_phi_rhs :769-774): rhs only
template <int NC>
__global__ void k2_parent(Ctx cx, int slot0, int nbox, int mode) {
  pdl_wait();
  using B = B2<NC>;
  const int slot = slot0 + blockIdx.x;
  if (cx.child0[slot] < 0) return;
  const double* phi = cx.cc[V_PHI] + (size_t)slot * B::BOX;
  double* rhs = cx.cc[V_RHS] + (size_t)slot * B::BOX;
  double* tmp = cx.cc[V_TMP] + (size_t)slot * B::BOX;
  for (int n = threadIdx.x; n < NC * NC; n += blockDim.x) {
    const int i = n % NC + 1, j = n / NC + 1;
    const int o = B::at(i, j);
    rhs[o] = apply2<NC>(cx, slot, phi, i, j) + tmp[o];
  }
  if (mode == 1) {
    __syncthreads();
    for (int q = threadIdx.x; q < B::BOX; q += blockDim.x) tmp[q] = phi[q];
  }
}

// correct_children (m_af_multigrid.f90:624-646), one CTA per child: phi_c += P(phi_p - tmp_p) with
// stencil_prolong_248 / _234 in 2D (m_af_stencil.f90:610-648, :715-764)
template <int NC>
__global__ void k2_correct(Ctx cx, int slot0, int nbox) {
  pdl_wait();
  using B = B2<NC>;
  constexpr int H = B::H, W = H + 2;
  __shared__ double sub[W * W];
  const int cslot = slot0 + blockIdx.x;
  const int p = cx.parent[cslot], ch = cx.coff[cslot];
  const double* pphi = cx.cc[V_PHI] + (size_t)p * B::BOX;
  const double* ptmp = cx.cc[V_TMP] + (size_t)p * B::BOX;
  const int ox = (ch & 1) * H, oy = ((ch >> 1) & 1) * H;
  for (int n = threadIdx.x; n < W * W; n += blockDim.x) {
    const int a = n % W, b = n / W;
    const int q = B::at(ox + a, oy + b);
    sub[n] = pphi[q] - ptmp[q];
  }
  __syncthreads();
  const int pkind = cx.pk ? cx.pk[cslot] : 0;
  const double* pv = pkind ? cx.stv + cx.poff[cslot] : cx.pcoef;
  const int pshape = (pkind == 0) ? cx.pshape : (pkind == 1 ? 4 : 3);
  double* cphi = cx.cc[V_PHI] + (size_t)cslot * B::BOX;
  for (int n = threadIdx.x; n < NC * NC; n += blockDim.x) {
    const int i = n % NC + 1, j = n / NC + 1;
    const int i1 = (i + 1) >> 1, i2 = i1 + 1 - 2 * (i & 1);
    const int j1 = (j + 1) >> 1, j2 = j1 + 1 - 2 * (j & 1);
    const double* c = (pkind == 3) ? pv + 3 * n : pv;
    double acc = cphi[B::at(i, j)];
    acc = acc + c[0] * sub[i1 + W * j1];
    acc = acc + c[1] * sub[i2 + W * j1];
    acc = acc + c[2] * sub[i1 + W * j2];
    if (pshape == 4) acc = acc + c[3] * sub[i2 + W * j2];
    cphi[B::at(i, j)] = acc;
  }
}

// tmp_p = phi_p - tmp_p on the full record of boxes with children (m_af_multigrid.f90:636-637)
template <int NC>
__global__ void k2_store_corr(Ctx cx, int slot0, int nbox) {
  pdl_wait();
  using B = B2<NC>;
  const int slot = slot0 + blockIdx.x;
  if (cx.child0[slot] < 0) return;
  const double* phi = cx.cc[V_PHI] + (size_t)slot * B::BOX;
  double* tmp = cx.cc[V_TMP] + (size_t)slot * B::BOX;
  for (int q = threadIdx.x; q < B::BOX; q += blockDim.x) tmp[q] = phi[q] - tmp[q];
}

// init_phi_rhs (m_af_multigrid.f90:779-799): phi = 0 on the box, restrict rhs into the parent (with the
// cylindrical weights: mg_box_rstr_lpl uses geometry for every variable but phi)
template <int NC>
__global__ void k2_restrict_var(Ctx cx, int slot0, int nbox, int var, int clear_phi) {
  pdl_wait();
  using B = B2<NC>;
  constexpr int H = B::H;
  const int slot = slot0 + blockIdx.x;
  const double* src = cx.cc[var] + (size_t)slot * B::BOX;
  const int p = cx.parent[slot], cof = cx.coff[slot];
  double* dst = cx.cc[var] + (size_t)p * B::BOX;
  const int ox = (cof & 1) * H, oy = ((cof >> 1) & 1) * H;
  if (clear_phi) {
    double* phi = cx.cc[V_PHI] + (size_t)slot * B::BOX;
    for (int q = threadIdx.x; q < B::BOX; q += blockDim.x) phi[q] = 0.0;
  }
  for (int n = threadIdx.x; n < H * H; n += blockDim.x) {
    const int ic = n % H + 1, jc = n / H + 1;
    const int i = 2 * ic - 1, j = 2 * jc - 1;
    double rr;
    if (cx.cyl && var != V_PHI) {
      const double drp = cx.drx[cx.lvl[p]];
      const double rc = cx.rmin[p] + (ox + ic - 0.5) * drp;
      const double t = 0.25 * drp / rc;
      const double w1 = 1 - t, w2 = 1 + t;
      double s1 = 0.0, s2 = 0.0;
      s1 = s1 + src[B::at(i, j)];
      s1 = s1 + src[B::at(i, j + 1)];
      s2 = s2 + src[B::at(i + 1, j)];
      s2 = s2 + src[B::at(i + 1, j + 1)];
      rr = 0.25 * (w1 * s1 + w2 * s2);
    } else {
      double s = 0.0;
      s = s + src[B::at(i, j)];
      s = s + src[B::at(i + 1, j)];
      s = s + src[B::at(i, j + 1)];
      s = s + src[B::at(i + 1, j + 1)];
      rr = 0.25 * s;
    }
    dst[B::at(ox + ic, oy + jc)] = rr;
  }
}

// max |var| over the interior of leaves (af_tree_maxabs_cc)
template <int NC>
__global__ void k2_maxabs(Ctx cx, int slot0, int nbox, int var, unsigned long long* maxabs_bits) {
  pdl_wait();
  using B = B2<NC>;
  const int slot = slot0 + blockIdx.x;
  if (cx.child0[slot] >= 0) return;
  const double* v = cx.cc[var] + (size_t)slot * B::BOX;
  double mx = 0.0;
  for (int n = threadIdx.x; n < NC * NC; n += blockDim.x) mx = fmax(mx, fabs(v[B::at(n % NC + 1, n / NC + 1)]));
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0 && mx > 0.0) atomicMax(maxabs_bits, (unsigned long long)__double_as_longlong(mx));
}

// per-box interior sums in (j, i) order, weighted by the cell radius and 2 pi in cylindrical coordinates
// (af_tree_sum_cc, m_af_utils.f90:966-1027); one thread per box (2D boxes are tiny)
template <int NC>
__global__ void k2_box_sums(Ctx cx, int nslots, int var, double* out) {
  pdl_wait();
  using B = B2<NC>;
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= nslots) return;
  const double* v = cx.cc[var] + (size_t)slot * B::BOX;
  double s = 0.0;
  if (cx.cyl) {
    const double dr = cx.drx[cx.lvl[slot]];
    for (int j = 1; j <= NC; ++j)
      for (int i = 1; i <= NC; ++i) s = s + v[B::at(i, j)] * (cx.rmin[slot] + (i - 0.5) * dr);
    s = s * cx.two_pi;
  } else {
    for (int j = 1; j <= NC; ++j)
      for (int i = 1; i <= NC; ++i) s = s + v[B::at(i, j)];
  }
  out[slot] = s;
}

// box records are already in the reference's order: upload / download are row copies
__global__ void k2_copy_boxes(double* var_base, const int* slots, int n, double* packed, int box_len, int to_device) {
  pdl_wait();
  const int q = blockIdx.x;
  if (q >= n || slots[q] < 0) return;
  double* a = var_base + (size_t)slots[q] * box_len;
  double* b = packed + (size_t)q * box_len;
  for (int t = threadIdx.x; t < box_len; t += blockDim.x) {
    if (to_device) a[t] = b[t];
    else b[t] = a[t];
  }
}

// field_set_rhs (src/m_field.f90:406-444), one species: rhs = (first ? 0 : rhs) + q * density on whole records
__global__ void k2_axpy_boxes(double* var_base, const int* slots, int n, const double* packed, int box_len, double q,
                              int first) {
  pdl_wait();
  const int b = blockIdx.x;
  if (b >= n || slots[b] < 0) return;
  double* a = var_base + (size_t)slots[b] * box_len;
  const double* s = packed + (size_t)b * box_len;
  for (int t = threadIdx.x; t < box_len; t += blockDim.x) a[t] = (first ? 0.0 : a[t]) + q * s[t];
}

// interior cells only: packed holds cc(1:nc, 1:nc) per box
__global__ void k2_unpack_interior(double* var_base, const int* slots, int n, double* packed, int nc, int up) {
  pdl_wait();
  const int q = blockIdx.x;
  if (q >= n || slots[q] < 0) return;
  double* a = var_base + (size_t)slots[q] * (nc + 2) * (nc + 2);
  double* b = packed + (size_t)q * nc * nc;
  for (int t = threadIdx.x; t < nc * nc; t += blockDim.x) {
    double* cell = a + (t % nc + 1) + (nc + 2) * (t / nc + 1);
    if (up) *cell = b[t];
    else b[t] = *cell;
  }
}

// ---- coarse grid: dense inverse of the BC-folded level-1 operator (cylindrical and variable stencils
// are not separable); same matrix as coarse_solver_initialize (m_coarse_solver.f90:71-194, :442-491)
struct Coarse2 {
  int nx[2];
  const int* bix;      // [nbox1*2] box%ix - 1
  const double* b2r;   // [nbox1][4][nc]
  const double* lsf_fac;  // [nbox1][nc^2] or null
  const double* Ainv;  // [n][n]
  double* v0;
  double* v1;
};

template <int NC>
__global__ void k2_cs_gather(Ctx cx, Coarse2 cs, int nbox1) {
  pdl_wait();
  using B = B2<NC>;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= nbox1 * NC * NC) return;
  const int bx = n / (NC * NC), r = n % (NC * NC);
  const int i = r % NC + 1, j = r / NC + 1;
  double t = cx.cc[V_RHS][(size_t)bx * B::BOX + B::at(i, j)];
  const int q[2] = {i, j};
#pragma unroll
  for (int f = 0; f < 4; ++f) {
    const int d = f >> 1;
    if (cx.nbr[bx * 4 + f] >= 0) continue;
    if (q[d] != ((f & 1) ? NC : 1)) continue;
    const int fi = q[1 - d] - 1;
    const int row = cx.aux[bx * 4 + f];
    t = t + cs.b2r[((size_t)bx * 4 + f) * NC + fi] * cx.rule_B[(size_t)row * NC + fi];
  }
  if (cs.lsf_fac) {
    const double V = (cx.bvoff && cx.bvoff[bx] >= 0) ? cx.bv[cx.bvoff[bx] + r] : cx.lsf_value();
    t = t + cs.lsf_fac[(size_t)bx * NC * NC + r] * V;
  }
  const int gi = cs.bix[bx * 2] * NC + i - 1, gj = cs.bix[bx * 2 + 1] * NC + j - 1;
  cs.v0[gi + cs.nx[0] * gj] = t;
}

__global__ void k2_cs_dense(Coarse2 cs, const double* in, double* out) {
  pdl_wait();
  const int n = cs.nx[0] * cs.nx[1];
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= n) return;
  const double* a = cs.Ainv + (size_t)row * n;
  double s = 0.0;
  for (int c = lane; c < n; c += 32) s = s + a[c] * in[c];
  for (int o = 16; o > 0; o >>= 1) s = s + __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[row] = s;
}

template <int NC>
__global__ void k2_cs_scatter(Ctx cx, Coarse2 cs, int nbox1, const double* x) {
  pdl_wait();
  using B = B2<NC>;
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= nbox1 * NC * NC) return;
  const int bx = n / (NC * NC), r = n % (NC * NC);
  const int i = r % NC + 1, j = r / NC + 1;
  const int gi = cs.bix[bx * 2] * NC + i - 1, gj = cs.bix[bx * 2 + 1] * NC + j - 1;
  cx.cc[V_PHI][(size_t)bx * B::BOX + B::at(i, j)] = x[gi + cs.nx[0] * gj];
}

}  // namespace afmg2

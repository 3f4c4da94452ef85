// Device box layout (3D).  One box record holds the same (nc+2)^3 values as the reference's
// cc(0:nc+1,0:nc+1,0:nc+1) (afivo/src/m_af_core.f90:551) but permuted so that every piece a kernel
// streams is contiguous and 16-byte aligned (one cp.async.bulk per piece):
//
//   [ colour 0 : interior(k,j,m)  | face 0 | face 1 | ... | face 5 ]      COL doubles
//   [ colour 1 : interior(k,j,m)  | face 0 | face 1 | ... | face 5 ]      COL doubles
//   [ 12 edges x nc ] [ 8 corners ]
//
// colour = (i+j+k) & 1 with the reference's box-local 1-based indices (globally consistent since
// nc is even, m_af_core.f90:161).  A red-black half-sweep (m_af_stencil.f90:956-973) that updates
// colour C reads only the other colour block and writes only block C.
//   interior cell (i,j,k), 1..nc : offset  C*COL + ((k-1)*nc + (j-1))*H + ((i-1)>>1),   H = nc/2
//   face f = nb-1 (lowx,highx,lowy,highy,lowz,highz), ghost index g = 0 or nc+1 in dim f/2,
//     transverse indices (a,b) in increasing dimension order (same as bc_val in bc_to_gc,
//     m_af_ghostcell.f90:173-279):  C*COL + NI + f*NF + (b-1)*H + ((a-1)>>1),  C = (g+a+b)&1
//   edge e (af_edge_dim/af_edge_dir numbering, m_af_types.f90:217-235), cell n along it:
//     OFF_E + e*nc + (n-1);  corner c (af_child_dix numbering): OFF_C + c
#pragma once
#include <cstdint>

#ifdef __CUDACC__
#define AFMG_HD __host__ __device__ __forceinline__
#else
#define AFMG_HD inline
#endif

template <int NC>
struct Lay3 {
  static constexpr int H = NC / 2;
  static constexpr int NI = NC * NC * H;      // interior cells of one colour
  static constexpr int NF = NC * H;           // cells of one colour on one face
  static constexpr int COL = NI + 6 * NF;     // doubles per colour block
  static constexpr int OFF_E = 2 * COL;
  static constexpr int OFF_C = OFF_E + 12 * NC;
  static constexpr int BOX = OFF_C + 8;       // == (NC+2)^3
  static constexpr int NC2 = NC * NC;
  static_assert(BOX == (NC + 2) * (NC + 2) * (NC + 2), "layout must be a permutation of the box");
  static_assert((COL * 8) % 16 == 0 && (BOX * 8) % 16 == 0, "bulk-copy alignment");

  // index inside a colour block of interior cell (m = (i-1)>>1, j, k)
  static AFMG_HD int iidx(int m, int j, int k) { return ((k - 1) * NC + (j - 1)) * H + m; }
  // index inside a colour block of face cell
  static AFMG_HD int fidx(int f, int a, int b) { return NI + f * NF + (b - 1) * H + ((a - 1) >> 1); }

  static AFMG_HD int interior(int i, int j, int k) {
    return ((i + j + k) & 1) * COL + iidx((i - 1) >> 1, j, k);
  }
  static AFMG_HD int face(int f, int a, int b) {
    int g = (f & 1) ? NC + 1 : 0;
    return ((g + a + b) & 1) * COL + fidx(f, a, b);
  }
  static AFMG_HD int edge(int e, int n) { return OFF_E + e * NC + (n - 1); }
  static AFMG_HD int corner(int c) { return OFF_C + c; }

  // general cell (0 <= i,j,k <= NC+1)
  static AFMG_HD int cell(int i, int j, int k) {
    const bool bi = (i == 0) | (i == NC + 1), bj = (j == 0) | (j == NC + 1), bk = (k == 0) | (k == NC + 1);
    const int nb = (int)bi + (int)bj + (int)bk;
    if (nb == 0) return interior(i, j, k);
    if (nb == 1) {
      if (bi) return face(i ? 1 : 0, j, k);
      if (bj) return face(j ? 3 : 2, i, k);
      return face(k ? 5 : 4, i, j);
    }
    if (nb == 2) {
      // edge along the one interior dimension; e = 4*dim + (low other dim high) + 2*(high other dim high)
      if (!bi) return edge(0 + (j ? 1 : 0) + (k ? 2 : 0), i);
      if (!bj) return edge(4 + (i ? 1 : 0) + (k ? 2 : 0), j);
      return edge(8 + (i ? 1 : 0) + (j ? 2 : 0), k);
    }
    return corner((i ? 1 : 0) + (j ? 2 : 0) + (k ? 4 : 0));
  }

  // inverse map: layout offset q -> (i,j,k)
  static AFMG_HD void uncell(int q, int& i, int& j, int& k) {
    if (q < OFF_E) {
      const int c = q >= COL ? 1 : 0;
      int r = q - c * COL;
      if (r < NI) {
        const int m = r % H;
        r /= H;
        j = r % NC + 1;
        k = r / NC + 1;
        i = 2 * m + 2 - ((c + j + k) & 1);
        return;
      }
      r -= NI;
      const int f = r / NF;
      r -= f * NF;
      const int ah = r % H, b = r / H + 1;
      const int g = (f & 1) ? NC + 1 : 0;
      const int a = 2 * ah + 2 - ((c + g + b) & 1);
      if (f < 2) { i = g; j = a; k = b; }
      else if (f < 4) { i = a; j = g; k = b; }
      else { i = a; j = b; k = g; }
      return;
    }
    if (q < OFF_C) {
      const int r = q - OFF_E, e = r / NC, n = r % NC + 1, dim = e >> 2;
      const int lo = (e & 1) ? NC + 1 : 0, hi = (e & 2) ? NC + 1 : 0;
      if (dim == 0) { i = n; j = lo; k = hi; }
      else if (dim == 1) { i = lo; j = n; k = hi; }
      else { i = lo; j = hi; k = n; }
      return;
    }
    const int c = q - OFF_C;
    i = (c & 1) ? NC + 1 : 0;
    j = (c & 2) ? NC + 1 : 0;
    k = (c & 4) ? NC + 1 : 0;
  }
};

// 2D boxes keep the reference's own (nc+2)^2 order (kernels2d.cuh): they are tiny and the 2D path is
// launch-latency bound, so there is nothing to gain from a colour split.

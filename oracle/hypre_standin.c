/* Stand-in for the part of Hypre's Fortran Struct interface that afivo's coarse-grid solver calls
 * (afivo/src/m_coarse_solver.f90:82-88,148-149,187-192,202-218,248-281,331-335,355,371-387,399-414,428-435).
 *
 * TEST / BASELINE INFRASTRUCTURE, not part of the product: with this library on the link line in place of
 * Hypre 2.31.0 (which /root/reference does not vendor), the UNMODIFIED afivo sources link on a machine that has a
 * Fortran compiler, and their coarse solve becomes the tolerance -> 0 limit of PFMG: an exact banded LU of the same
 * matrix the reference assembled.  That is the coarse solve the oracle (oracle/afmg_oracle.cpp) and the GPU library
 * use, so a reference built this way is comparable to both at 1e-10 instead of at PFMG's 1e-6.
 *
 * Conventions (gfortran, implicit interfaces): symbol = lower-case name + '_'; every argument by reference; object
 * handles are type(c_ptr) variables, i.e. void** here; integers are default (32-bit) integers; the MPI communicator
 * is an ignored integer (m_coarse_solver.f90:41-44 stubs MPI).
 *
 * Semantics kept from Hypre: index space [ilower, iupper] per dimension, first dimension fastest in box values;
 * matrix box values are ordered (entry fastest, then cells); stencil entries that reach outside a non-periodic grid
 * are ignored; SetSymmetric(1) stores the centre and the "upper" entries only and mirrors them.  SMG, PFMG and CycRed
 * all map to the same direct solve; GetNumIterations returns 1.
 *
 * A singular matrix (all-Neumann / fully periodic Poisson) gets its null pivot pinned (that unknown = 0), which
 * selects one member of the solution family; callers of such systems subtract the mean afterwards
 * (mg%subtract_mean, m_af_multigrid.f90:256-258).
 *
 * Build: make -C oracle hypre   ->  oracle/_hypre/libHYPRE.so
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define MAXD 3

typedef struct {
  int ndim, lo[MAXD], hi[MAXD], period[MAXD], n[MAXD];
  long ncell;
} grid_t;

typedef struct {
  int ndim, size;
  int (*off)[MAXD];
} stencil_t;

typedef struct {
  grid_t g;
  stencil_t st;
  int symmetric;
  double* val; /* [ncell][st.size] */
  long version;
} matrix_t;

typedef struct {
  grid_t g;
  double* val;
} vector_t;

typedef struct {
  int max_iter, n_pre, n_post, n_iter;
  double tol;
  /* factorisation */
  long n, bw, version;
  const matrix_t* of;
  char* pinned; /* rows whose null pivot was replaced by the equation x = 0 */
  double* lu; /* banded, row p holds columns p-bw .. p+bw at lu[p*(2bw+1) + (q-p+bw)] */
} solver_t;

static long cell_index(const grid_t* g, const int* ijk) {
  long p = 0, stride = 1;
  for (int d = 0; d < g->ndim; ++d) {
    p += (long)(ijk[d] - g->lo[d]) * stride;
    stride *= g->n[d];
  }
  return p;
}

/* neighbour of cell ijk at offset o; -1 when it falls outside a non-periodic direction */
static long neighbour(const grid_t* g, const int* ijk, const int* o) {
  int q[MAXD];
  for (int d = 0; d < g->ndim; ++d) {
    int x = ijk[d] + o[d];
    if (x < g->lo[d] || x > g->hi[d]) {
      if (g->period[d] <= 0) return -1;
      x = g->lo[d] + ((x - g->lo[d]) % g->n[d] + g->n[d]) % g->n[d];
    }
    q[d] = x;
  }
  return cell_index(g, q);
}

static int box_ok(const grid_t* g, const int* lo, const int* hi) {
  for (int d = 0; d < g->ndim; ++d)
    if (lo[d] < g->lo[d] || hi[d] > g->hi[d] || hi[d] < lo[d]) return 0;
  return 1;
}

/* ---- library lifetime */
void hypre_initialize_(int* ierr) { *ierr = 0; }
void hypre_finalize_(int* ierr) { *ierr = 0; }

/* ---- grid */
void hypre_structgridcreate_(const int* comm, const int* ndim, void** grid, int* ierr) {
  (void)comm;
  *ierr = 1;
  if (*ndim < 1 || *ndim > MAXD) return;
  grid_t* g = (grid_t*)calloc(1, sizeof *g);
  if (!g) return;
  g->ndim = *ndim;
  *grid = g;
  *ierr = 0;
}

void hypre_structgridsetextents_(void** grid, const int* ilower, const int* iupper, int* ierr) {
  grid_t* g = (grid_t*)*grid;
  g->ncell = 1;
  for (int d = 0; d < g->ndim; ++d) {
    g->lo[d] = ilower[d];
    g->hi[d] = iupper[d];
    g->n[d] = iupper[d] - ilower[d] + 1;
    g->ncell *= g->n[d];
  }
  *ierr = g->ncell > 0 ? 0 : 1;
}

void hypre_structgridsetperiodic_(void** grid, const int* period, int* ierr) {
  grid_t* g = (grid_t*)*grid;
  *ierr = 0;
  for (int d = 0; d < g->ndim; ++d) {
    g->period[d] = period[d];
    if (period[d] != 0 && period[d] != g->n[d]) *ierr = 1; /* afivo only asks for the grid's own size */
  }
}

void hypre_structgridassemble_(void** grid, int* ierr) {
  (void)grid;
  *ierr = 0;
}

void hypre_structgriddestroy_(void** grid, int* ierr) {
  free(*grid);
  *grid = NULL;
  *ierr = 0;
}

/* ---- stencil */
void hypre_structstencilcreate_(const int* ndim, const int* size, void** stencil, int* ierr) {
  *ierr = 1;
  if (*ndim < 1 || *ndim > MAXD || *size < 1) return;
  stencil_t* s = (stencil_t*)calloc(1, sizeof *s);
  if (!s) return;
  s->ndim = *ndim;
  s->size = *size;
  s->off = calloc((size_t)*size, sizeof *s->off);
  *stencil = s;
  *ierr = 0;
}

void hypre_structstencilsetelement_(void** stencil, const int* index, const int* offset, int* ierr) {
  stencil_t* s = (stencil_t*)*stencil;
  *ierr = 1;
  if (*index < 0 || *index >= s->size) return;
  for (int d = 0; d < s->ndim; ++d) s->off[*index][d] = offset[d];
  *ierr = 0;
}

void hypre_structstencildestroy_(void** stencil, int* ierr) {
  stencil_t* s = (stencil_t*)*stencil;
  if (s) free(s->off);
  free(s);
  *stencil = NULL;
  *ierr = 0;
}

/* ---- matrix */
void hypre_structmatrixcreate_(const int* comm, void** grid, void** stencil, void** matrix, int* ierr) {
  (void)comm;
  *ierr = 1;
  const grid_t* g = (const grid_t*)*grid;
  const stencil_t* s = (const stencil_t*)*stencil;
  matrix_t* A = (matrix_t*)calloc(1, sizeof *A);
  if (!A) return;
  A->g = *g; /* copies: the caller destroys the stencil right after (m_coarse_solver.f90:387) */
  A->st.ndim = s->ndim;
  A->st.size = s->size;
  A->st.off = calloc((size_t)s->size, sizeof *A->st.off);
  memcpy(A->st.off, s->off, (size_t)s->size * sizeof *s->off);
  *matrix = A;
  *ierr = 0;
}

void hypre_structmatrixsetsymmetric_(void** matrix, const int* symmetric, int* ierr) {
  ((matrix_t*)*matrix)->symmetric = *symmetric;
  *ierr = 0;
}

void hypre_structmatrixinitialize_(void** matrix, int* ierr) {
  matrix_t* A = (matrix_t*)*matrix;
  free(A->val);
  A->val = (double*)calloc((size_t)A->g.ncell * A->st.size, sizeof(double));
  *ierr = A->val ? 0 : 1;
}

void hypre_structmatrixsetboxvalues_(void** matrix, const int* ilower, const int* iupper, const int* nentries,
                                     const int* entries, const double* values, int* ierr) {
  matrix_t* A = (matrix_t*)*matrix;
  *ierr = 1;
  if (!A->val || !box_ok(&A->g, ilower, iupper)) return;
  for (int e = 0; e < *nentries; ++e)
    if (entries[e] < 0 || entries[e] >= A->st.size) return;
  int ijk[MAXD] = {0, 0, 0}, lo[MAXD] = {0, 0, 0}, hi[MAXD] = {0, 0, 0};
  for (int d = 0; d < A->g.ndim; ++d) lo[d] = ilower[d], hi[d] = iupper[d];
  long cnt = 0;
  for (ijk[2] = lo[2]; ijk[2] <= hi[2]; ++ijk[2])
    for (ijk[1] = lo[1]; ijk[1] <= hi[1]; ++ijk[1])
      for (ijk[0] = lo[0]; ijk[0] <= hi[0]; ++ijk[0]) {
        const long p = cell_index(&A->g, ijk);
        for (int e = 0; e < *nentries; ++e) A->val[p * A->st.size + entries[e]] = values[cnt++];
      }
  A->version++;
  *ierr = 0;
}

void hypre_structmatrixassemble_(void** matrix, int* ierr) {
  (void)matrix;
  *ierr = 0;
}

void hypre_structmatrixdestroy_(void** matrix, int* ierr) {
  matrix_t* A = (matrix_t*)*matrix;
  if (A) {
    free(A->val);
    free(A->st.off);
  }
  free(A);
  *matrix = NULL;
  *ierr = 0;
}

/* ---- vector */
void hypre_structvectorcreate_(const int* comm, void** grid, void** vec, int* ierr) {
  (void)comm;
  *ierr = 1;
  vector_t* v = (vector_t*)calloc(1, sizeof *v);
  if (!v) return;
  v->g = *(const grid_t*)*grid;
  *vec = v;
  *ierr = 0;
}

void hypre_structvectorinitialize_(void** vec, int* ierr) {
  vector_t* v = (vector_t*)*vec;
  free(v->val);
  v->val = (double*)calloc((size_t)v->g.ncell, sizeof(double));
  *ierr = v->val ? 0 : 1;
}

void hypre_structvectorassemble_(void** vec, int* ierr) {
  (void)vec;
  *ierr = 0;
}

static void vector_box(vector_t* v, const int* ilower, const int* iupper, double* values, int get, int* ierr) {
  *ierr = 1;
  if (!v->val || !box_ok(&v->g, ilower, iupper)) return;
  int ijk[MAXD] = {0, 0, 0}, lo[MAXD] = {0, 0, 0}, hi[MAXD] = {0, 0, 0};
  for (int d = 0; d < v->g.ndim; ++d) lo[d] = ilower[d], hi[d] = iupper[d];
  long cnt = 0;
  for (ijk[2] = lo[2]; ijk[2] <= hi[2]; ++ijk[2])
    for (ijk[1] = lo[1]; ijk[1] <= hi[1]; ++ijk[1])
      for (ijk[0] = lo[0]; ijk[0] <= hi[0]; ++ijk[0]) {
        const long p = cell_index(&v->g, ijk);
        if (get) values[cnt++] = v->val[p];
        else v->val[p] = values[cnt++];
      }
  *ierr = 0;
}

void hypre_structvectorsetboxvalues_(void** vec, const int* ilower, const int* iupper, const double* values, int* ierr) {
  vector_box((vector_t*)*vec, ilower, iupper, (double*)values, 0, ierr);
}

void hypre_structvectorgetboxvalues_(void** vec, const int* ilower, const int* iupper, double* values, int* ierr) {
  vector_box((vector_t*)*vec, ilower, iupper, values, 1, ierr);
}

void hypre_structvectordestroy_(void** vec, int* ierr) {
  vector_t* v = (vector_t*)*vec;
  if (v) free(v->val);
  free(v);
  *vec = NULL;
  *ierr = 0;
}

/* ---- the solver: banded LU without pivoting (the matrices of this path are M-matrices up to sign) */
static void for_each_entry(const matrix_t* A, void (*fn)(long p, long q, double v, void* ctx), void* ctx) {
  const grid_t* g = &A->g;
  int ijk[MAXD] = {0, 0, 0}, lo[MAXD] = {0, 0, 0}, hi[MAXD] = {0, 0, 0};
  for (int d = 0; d < g->ndim; ++d) lo[d] = g->lo[d], hi[d] = g->hi[d];
  for (ijk[2] = lo[2]; ijk[2] <= hi[2]; ++ijk[2])
    for (ijk[1] = lo[1]; ijk[1] <= hi[1]; ++ijk[1])
      for (ijk[0] = lo[0]; ijk[0] <= hi[0]; ++ijk[0]) {
        const long p = cell_index(g, ijk);
        for (int e = 0; e < A->st.size; ++e) {
          const long q = neighbour(g, ijk, A->st.off[e]);
          if (q < 0) continue;
          const double v = A->val[p * A->st.size + e];
          fn(p, q, v, ctx);
          if (A->symmetric && q != p) fn(q, p, v, ctx);
        }
      }
}

static void band_width(long p, long q, double v, void* ctx) {
  long* bw = (long*)ctx;
  if (v != 0.0 && labs(p - q) > *bw) *bw = labs(p - q);
}

static void band_add(long p, long q, double v, void* ctx) {
  solver_t* s = (solver_t*)ctx;
  if (v != 0.0) s->lu[p * (2 * s->bw + 1) + (q - p + s->bw)] += v;
}

static int factorise(solver_t* s, const matrix_t* A) {
  free(s->lu);
  free(s->pinned);
  s->lu = NULL;
  s->n = A->g.ncell;
  s->pinned = (char*)calloc((size_t)s->n, 1);
  s->bw = 0;
  for_each_entry(A, band_width, &s->bw);
  const long w = 2 * s->bw + 1, n = s->n, bw = s->bw;
  s->lu = (double*)calloc((size_t)n * w, sizeof(double));
  if (!s->lu || !s->pinned) return 1;
  for_each_entry(A, band_add, s);
  double dmax = 0;
  for (long p = 0; p < n; ++p) dmax = fmax(dmax, fabs(s->lu[p * w + bw]));
  for (long k = 0; k < n; ++k) {
    double* rk = s->lu + k * w;
    if (fabs(rk[bw]) <= 1e-13 * dmax) { /* null pivot of a singular system: pin this unknown to zero */
      for (long c = 0; c < w; ++c) rk[c] = 0;
      rk[bw] = 1;
      s->pinned[k] = 1;
      for (long i = k + 1; i <= k + bw && i < n; ++i) s->lu[i * w + (k - i + bw)] = 0;
      continue;
    }
    const double inv = 1.0 / rk[bw];
    const long jmax = (k + bw < n - 1) ? k + bw : n - 1;
    for (long i = k + 1; i <= jmax; ++i) {
      double* ri = s->lu + i * w;
      const double l = ri[k - i + bw] * inv;
      if (l == 0.0) continue;
      ri[k - i + bw] = l;
      for (long j = k + 1; j <= jmax; ++j) ri[j - i + bw] -= l * rk[j - k + bw];
    }
  }
  s->of = A;
  s->version = A->version;
  return 0;
}

static int solve(solver_t* s, const matrix_t* A, const vector_t* b, vector_t* x) {
  if (s->of != A || s->version != A->version || !s->lu)
    if (factorise(s, A)) return 1;
  const long w = 2 * s->bw + 1, n = s->n, bw = s->bw;
  if (b->g.ncell != n || x->g.ncell != n) return 1;
  double* y = x->val;
  for (long p = 0; p < n; ++p) y[p] = b->val[p];
  for (long i = 0; i < n; ++i) { /* forward: L has unit diagonal */
    const long j0 = i - bw > 0 ? i - bw : 0;
    double acc = y[i];
    for (long j = j0; j < i; ++j) acc -= s->lu[i * w + (j - i + bw)] * y[j];
    y[i] = s->pinned[i] ? 0.0 : acc;
  }
  for (long i = n - 1; i >= 0; --i) {
    const long j1 = i + bw < n - 1 ? i + bw : n - 1;
    double acc = y[i];
    for (long j = i + 1; j <= j1; ++j) acc -= s->lu[i * w + (j - i + bw)] * y[j];
    y[i] = acc / s->lu[i * w + bw];
  }
  s->n_iter = 1;
  return 0;
}

static void solver_create(void** solver, int* ierr) {
  solver_t* s = (solver_t*)calloc(1, sizeof *s);
  *solver = s;
  *ierr = s ? 0 : 1;
}

static void solver_destroy(void** solver, int* ierr) {
  solver_t* s = (solver_t*)*solver;
  if (s) {
    free(s->lu);
    free(s->pinned);
  }
  free(s);
  *solver = NULL;
  *ierr = 0;
}

#define SOLVER_COMMON(name)                                                                                  \
  void hypre_struct##name##create_(const int* comm, void** solver, int* ierr) {                              \
    (void)comm;                                                                                              \
    solver_create(solver, ierr);                                                                             \
  }                                                                                                          \
  void hypre_struct##name##destroy_(void** solver, int* ierr) { solver_destroy(solver, ierr); }              \
  void hypre_struct##name##setup_(void** solver, void** A, void** b, void** x, int* ierr) {                  \
    (void)b;                                                                                                 \
    (void)x;                                                                                                 \
    *ierr = factorise((solver_t*)*solver, (const matrix_t*)*A);                                              \
  }                                                                                                          \
  void hypre_struct##name##solve_(void** solver, void** A, void** b, void** x, int* ierr) {                  \
    *ierr = solve((solver_t*)*solver, (const matrix_t*)*A, (const vector_t*)*b, (vector_t*)*x);              \
  }

#define SOLVER_ITERATIVE(name)                                                                               \
  void hypre_struct##name##setmaxiter_(void** solver, const int* v, int* ierr) {                             \
    ((solver_t*)*solver)->max_iter = *v;                                                                     \
    *ierr = 0;                                                                                               \
  }                                                                                                          \
  void hypre_struct##name##settol_(void** solver, const double* v, int* ierr) {                              \
    ((solver_t*)*solver)->tol = *v;                                                                          \
    *ierr = 0;                                                                                               \
  }                                                                                                          \
  void hypre_struct##name##setnumprerelax_(void** solver, const int* v, int* ierr) {                         \
    ((solver_t*)*solver)->n_pre = *v;                                                                        \
    *ierr = 0;                                                                                               \
  }                                                                                                          \
  void hypre_struct##name##setnumpostrelax_(void** solver, const int* v, int* ierr) {                        \
    ((solver_t*)*solver)->n_post = *v;                                                                       \
    *ierr = 0;                                                                                               \
  }

SOLVER_COMMON(cycred)
SOLVER_COMMON(smg)
SOLVER_COMMON(pfmg)
SOLVER_ITERATIVE(smg)
SOLVER_ITERATIVE(pfmg)

void hypre_structsmggetnumiterations_(void** solver, int* n, int* ierr) {
  *n = ((solver_t*)*solver)->n_iter;
  *ierr = 0;
}

/* sic: Hypre's Fortran name has no final 's' (m_coarse_solver.f90:435) */
void hypre_structpfmggetnumiteration_(void** solver, int* n, int* ierr) {
  *n = ((solver_t*)*solver)->n_iter;
  *ierr = 0;
}

"""CPU: the Hypre stand-in of oracle/hypre_standin.c (second boundary), driven through the Fortran-mangled
symbols in the order afivo/src/m_coarse_solver.f90 calls them (hypre_create_grid :238-265, hypre_create_vector
:268-282, hypre_create_matrix :361-389, hypre_set_matrix :104-194 with stencil_handle_boundaries :442-491,
coarse_solver_set_rhs_phi :286-340, coarse_solver :421-439, coarse_solver_get_phi :343-358), checked against an
independent scipy solve and against the oracle's coarse-grid solve of the same level-1 problem."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spl

from afivo_streamer_b200 import mg as M
from afivo_streamer_b200 import tree as T
from afivo_streamer_b200 import workloads as W
from oracle.oracle import Oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_hypre", "libHYPRE.so")

# every Hypre routine m_coarse_solver.f90 calls, as gfortran names it
SYMBOLS = """hypre_initialize hypre_finalize
hypre_structgridcreate hypre_structgridsetextents hypre_structgridsetperiodic hypre_structgridassemble
hypre_structgriddestroy hypre_structstencilcreate hypre_structstencilsetelement hypre_structstencildestroy
hypre_structmatrixcreate hypre_structmatrixsetsymmetric hypre_structmatrixinitialize hypre_structmatrixsetboxvalues
hypre_structmatrixassemble hypre_structmatrixdestroy hypre_structvectorcreate hypre_structvectorinitialize
hypre_structvectorassemble hypre_structvectorsetboxvalues hypre_structvectorgetboxvalues hypre_structvectordestroy
hypre_structcycredcreate hypre_structcycredsetup hypre_structcycredsolve hypre_structcycreddestroy
hypre_structsmgcreate hypre_structsmgsetmaxiter hypre_structsmgsettol hypre_structsmgsetnumprerelax
hypre_structsmgsetnumpostrelax hypre_structsmgsetup hypre_structsmgsolve hypre_structsmggetnumiterations
hypre_structsmgdestroy hypre_structpfmgcreate hypre_structpfmgsetmaxiter hypre_structpfmgsettol
hypre_structpfmgsetnumprerelax hypre_structpfmgsetnumpostrelax hypre_structpfmgsetup hypre_structpfmgsolve
hypre_structpfmggetnumiteration hypre_structpfmgdestroy""".split()

OFFSETS = {2: [(0, 0), (-1, 0), (1, 0), (0, -1), (0, 1)],
           3: [(0, 0, 0), (-1, 0, 0), (1, 0, 0), (0, -1, 0), (0, 1, 0), (0, 0, -1), (0, 0, 1)]}  # stencil_offsets :27-39


@pytest.fixture(scope="module")
def hy():
    if not os.path.exists(LIB):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s", "hypre"])
    return C.CDLL(LIB)


def test_exports_every_routine_the_reference_calls(hy):
    for s in SYMBOLS:
        assert hasattr(hy, s + "_"), s
    src = "/root/reference/afivo/src/m_coarse_solver.f90"  # only in the build container; the list above is the contract
    if os.path.exists(src):
        import re
        text = "\n".join(ln.split("!")[0] for ln in open(src).read().splitlines())  # comments hold a dead call
        called = {m.lower() for m in re.findall(r"call\s+(hypre_\w+)\s*\(", text, flags=re.I)}
        local = {m.lower() for m in re.findall(r"^\s*subroutine\s+(hypre_\w+)", text, flags=re.I | re.M)
                 if m.lower() != "hypre_structmatrixsetboxvalues"}  # that one is an interface block
        assert called - local <= set(SYMBOLS), sorted(called - local - set(SYMBOLS))


class Hypre:
    """The call sequences of m_coarse_solver.f90, argument for argument (everything by reference)."""

    def __init__(self, lib, nx, periodic, stencil_ix, symmetric, solver="pfmg"):
        self.L, self.nd, self.nx = lib, len(nx), list(nx)
        self.ierr = C.c_int(0)
        self.comm = C.c_int(0)
        self.stencil_ix = list(stencil_ix)  # 1-based rows of the full stencil, as in hypre_set_matrix
        self.solver_name = solver
        self.call("hypre_initialize")
        # hypre_create_grid
        self.grid = C.c_void_p()
        self.call("hypre_structgridcreate", self.comm, C.c_int(self.nd), self.grid)
        self.call("hypre_structgridsetextents", self.grid, self.iv([1] * self.nd), self.iv(nx))
        self.call("hypre_structgridsetperiodic", self.grid, self.iv([n if p else 0 for n, p in zip(nx, periodic)]))
        self.call("hypre_structgridassemble", self.grid)
        self.phi, self.rhs = self.vector(), self.vector()
        # hypre_create_matrix
        st = C.c_void_p()
        size = len(self.stencil_ix)
        self.call("hypre_structstencilcreate", C.c_int(self.nd), C.c_int(size), st)
        for i, six in enumerate(self.stencil_ix):
            self.call("hypre_structstencilsetelement", st, C.c_int(i), self.iv(OFFSETS[self.nd][six - 1]))
        self.A = C.c_void_p()
        self.call("hypre_structmatrixcreate", self.comm, self.grid, st, self.A)
        self.call("hypre_structmatrixsetsymmetric", self.A, C.c_int(symmetric))
        self.call("hypre_structmatrixinitialize", self.A)
        self.call("hypre_structstencildestroy", st)
        self.solver = None

    def iv(self, a):
        return (C.c_int * len(a))(*[int(x) for x in a])

    def call(self, name, *args):
        refs = [a if isinstance(a, (C.Array, C._Pointer)) else C.byref(a) for a in args]  # scalars and handles by reference
        getattr(self.L, name + "_")(*refs, C.byref(self.ierr))
        assert self.ierr.value == 0, name

    def vector(self):
        v = C.c_void_p()
        self.call("hypre_structvectorcreate", self.comm, self.grid, v)
        self.call("hypre_structvectorinitialize", v)
        self.call("hypre_structvectorassemble", v)
        return v

    def dp(self, a):
        self._keep = np.ascontiguousarray(a, np.float64)
        return self._keep.ctypes.data_as(C.POINTER(C.c_double))

    def set_matrix_box(self, ilo, ihi, full_coeffs):
        """full_coeffs: (ncell, 2D+1) in IJK order (i fastest)."""
        size = len(self.stencil_ix)
        coeffs = full_coeffs[:, [s - 1 for s in self.stencil_ix]]
        self.call("hypre_structmatrixsetboxvalues", self.A, self.iv(ilo), self.iv(ihi), C.c_int(size),
                  self.iv(range(size)), self.dp(coeffs))

    def prepare_solve(self):
        self.call("hypre_structmatrixassemble", self.A)
        self.solver = C.c_void_p()
        n = self.solver_name
        self.call(f"hypre_struct{n}create", self.comm, self.solver)
        if n != "cycred":
            self.call(f"hypre_struct{n}setmaxiter", self.solver, C.c_int(50))
            self.call(f"hypre_struct{n}settol", self.solver, C.c_double(1e-6))
            self.call(f"hypre_struct{n}setnumprerelax", self.solver, C.c_int(1))
            self.call(f"hypre_struct{n}setnumpostrelax", self.solver, C.c_int(1))
        self.call(f"hypre_struct{n}setup", self.solver, self.A, self.rhs, self.phi)

    def set_box(self, vec, ilo, ihi, vals):
        self.call("hypre_structvectorsetboxvalues", vec, self.iv(ilo), self.iv(ihi), self.dp(vals))

    def get_box(self, vec, ilo, ihi):
        out = np.zeros(int(np.prod(np.array(ihi) - np.array(ilo) + 1)))
        self.call("hypre_structvectorgetboxvalues", vec, self.iv(ilo), self.iv(ihi),
                  out.ctypes.data_as(C.POINTER(C.c_double)))
        return out

    def solve(self):
        n = self.solver_name
        self.call(f"hypre_struct{n}solve", self.solver, self.A, self.rhs, self.phi)
        if n != "cycred":
            it = C.c_int(0)
            self.call("hypre_structsmggetnumiterations" if n == "smg" else "hypre_structpfmggetnumiteration",
                      self.solver, it)
            assert it.value == 1

    def destroy(self):
        self.call("hypre_structgriddestroy", self.grid)
        self.call("hypre_structmatrixdestroy", self.A)
        self.call("hypre_structvectordestroy", self.rhs)
        self.call("hypre_structvectordestroy", self.phi)
        self.call(f"hypre_struct{self.solver_name}destroy", self.solver)
        self.call("hypre_finalize")


def scipy_matrix(nx, periodic, full):
    """Independent assembly: full[(cells in IJK order), 2D+1]; entries that leave a non-periodic grid are dropped."""
    nd = len(nx)
    n = int(np.prod(nx))
    idx = np.arange(n)
    ijk = [(idx // int(np.prod(nx[:d]))) % nx[d] for d in range(nd)]
    rows, cols, vals = [], [], []
    for e, off in enumerate(OFFSETS[nd]):
        q = [ijk[d] + off[d] for d in range(nd)]
        ok = np.ones(n, bool)
        for d in range(nd):
            if periodic[d]:
                q[d] = q[d] % nx[d]
            else:
                ok &= (q[d] >= 0) & (q[d] < nx[d])
        col = sum(q[d] * int(np.prod(nx[:d])) for d in range(nd))
        rows.append(idx[ok]); cols.append(col[ok]); vals.append(full[ok, e])
    return sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n))


def laplacian_dirichlet(nx, h, periodic, lam=0.0):
    """Full 2D+1 coefficients of a Dirichlet-0-folded Laplacian (- lam) on a single box grid, IJK order."""
    nd = len(nx)
    n = int(np.prod(nx))
    full = np.zeros((n, 2 * nd + 1))
    full[:, 1:] = 1 / h ** 2
    full[:, 0] = -2 * nd / h ** 2 - lam
    idx = np.arange(n)
    for d in range(nd):
        if periodic[d]:
            continue
        x = (idx // int(np.prod(nx[:d]))) % nx[d]
        for side, e in ((0, 1 + 2 * d), (nx[d] - 1, 2 + 2 * d)):
            m = x == side
            full[m, 0] -= full[m, e]
            full[m, e] = 0
    return full


@pytest.mark.parametrize("nx,periodic,solver", [((8, 8, 8), (0, 0, 0), "pfmg"), ((16, 8), (0, 0), "smg"),
                                                 ((8, 8, 4), (1, 0, 0), "pfmg"), ((8, 8), (0, 1), "cycred")])
def test_solves_the_assembled_system(hy, nx, periodic, solver):
    nd = len(nx)
    rng = np.random.default_rng(7)
    full = laplacian_dirichlet(nx, 1.0 / nx[0], periodic, lam=3.0)
    full[:, 1:] *= 1 + 0.3 * rng.random((full.shape[0], 1))  # cylindrical-like asymmetry between rows
    b = rng.standard_normal(full.shape[0])
    H = Hypre(hy, nx, periodic, range(1, 2 * nd + 2), symmetric=0, solver=solver)
    H.set_matrix_box([1] * nd, nx, full)
    H.prepare_solve()
    H.set_box(H.rhs, [1] * nd, nx, b)
    H.set_box(H.phi, [1] * nd, nx, np.zeros_like(b))
    H.solve()
    x = H.get_box(H.phi, [1] * nd, nx)
    ref = spl.spsolve(scipy_matrix(nx, periodic, full).tocsc(), b)
    assert np.max(np.abs(x - ref)) <= 1e-11 * np.max(np.abs(ref))
    H.destroy()


def test_symmetric_storage_mirrors_the_upper_entries(hy):
    nx, per = (8, 8, 8), (0, 0, 0)
    full = laplacian_dirichlet(nx, 0.125, per)
    b = np.random.default_rng(3).standard_normal(full.shape[0])
    out = []
    for sym, six in ((0, [1, 2, 3, 4, 5, 6, 7]), (1, [1, 3, 5, 7])):  # hypre_set_matrix :119-143
        H = Hypre(hy, nx, per, six, symmetric=sym)
        H.set_matrix_box([1, 1, 1], nx, full)
        H.prepare_solve()
        H.set_box(H.rhs, [1, 1, 1], nx, b)
        H.solve()
        out.append(H.get_box(H.phi, [1, 1, 1], nx))
        H.destroy()
    assert np.max(np.abs(out[0] - out[1])) <= 1e-12 * np.max(np.abs(out[0]))


def test_singular_neumann_system_returns_a_member_of_the_family(hy):
    nx, per = (8, 8), (1, 1)
    full = laplacian_dirichlet(nx, 0.125, per)
    rng = np.random.default_rng(5)
    b = rng.standard_normal(64)
    b -= b.mean()
    H = Hypre(hy, nx, per, range(1, 6), symmetric=0)
    H.set_matrix_box([1, 1], nx, full)
    H.prepare_solve()
    H.set_box(H.rhs, [1, 1], nx, b)
    H.solve()
    x = H.get_box(H.phi, [1, 1], nx)
    A = scipy_matrix(nx, per, full)
    assert np.max(np.abs(A @ x - b)) < 1e-10 * np.max(np.abs(b)) * 64
    H.destroy()


def test_reference_coarse_solve_sequence_matches_the_oracle(hy):
    """Level 1 = 2 x 2 x 2 boxes of 8^3; Dirichlet (non-zero) in z, Neumann (non-zero) in x, y; the per-box loop of
    hypre_set_matrix / coarse_solver_set_rhs_phi restated here, the solve by the stand-in, against the oracle's
    solve_coarse_grid (its own restatement of the same folding with a banded LU)."""
    tree = T.uniform_tree(3, 8, 16, 2)
    nc = tree.nc

    def sides(nb, c):
        if (nb - 1) // 2 == 2:
            return W.AF_BC_DIRICHLET, 1.0 + c[..., 0] * c[..., 1] + (nb == 6)
        return W.AF_BC_NEUMANN, 0.5 * c[..., 2] - 0.25

    bc = W.bc_table(tree, sides)
    orc = Oracle(tree)
    orc.set_bc(bc)
    orc.mg_init()
    l1 = np.asarray(tree.lvl_ids[0], np.int32)
    rng = np.random.default_rng(11)
    rhs = np.zeros((len(l1),) + (nc + 2,) * 3)
    rhs[:, 1:-1, 1:-1, 1:-1] = rng.standard_normal((len(l1), nc, nc, nc))
    orc.set_cc(M.I_RHS, l1, rhs)
    orc.set_cc(M.I_PHI, l1, np.zeros_like(rhs))
    orc.solve_coarse_grid()
    want = orc.get_cc(M.I_PHI, l1).reshape(rhs.shape)[:, 1:-1, 1:-1, 1:-1]

    nx = [int(v) for v in tree.coarse_grid_size]
    H = Hypre(hy, nx, (0, 0, 0), [1, 3, 5, 7], symmetric=1)  # Cartesian, no lsf: symmetric storage (:110-128)
    bcmap = {(int(i), int(n)): (int(t), v) for i, n, t, v in zip(bc.ids, bc.nbs, bc.types, bc.vals)}
    bc_to_rhs = {}
    for b in l1:
        dr = tree.dr[b]
        full = np.zeros((nc, nc, nc, 7))  # [k, j, i, entry]; mg_box_lpl_stencil (m_af_multigrid.f90:1246-1264)
        for d in range(3):
            full[..., 1 + 2 * d] = full[..., 2 + 2 * d] = 1 / dr[d] ** 2
        full[..., 0] = -full[..., 1:].sum(axis=-1)
        for nb in range(1, 7):  # stencil_handle_boundaries (:442-491)
            if tree.neighbors[b, nb - 1] >= 0:
                continue
            d, hi = (nb - 1) // 2, (nb - 1) % 2
            sl = [slice(None)] * 3
            sl[2 - d] = nc - 1 if hi else 0
            sl = tuple(sl)
            t, _ = bcmap[(int(b), nb)]
            if t == W.AF_BC_DIRICHLET:
                full[sl + (0,)] -= full[sl + (nb,)]
                bc_to_rhs[(int(b), nb)] = (-2 * full[sl + (nb,)]).ravel()
            else:
                full[sl + (0,)] += full[sl + (nb,)]
                bc_to_rhs[(int(b), nb)] = -(full[sl + (nb,)] * dr[d]).ravel() * (1 if hi else -1)
            full[sl + (nb,)] = 0
        ilo = (tree.ix[b] - 1) * nc + 1
        H.set_matrix_box(ilo, ilo + nc - 1, full.reshape(-1, 7))
    H.prepare_solve()
    for n, b in enumerate(l1):  # coarse_solver_set_rhs_phi (:286-340)
        tmp = rhs[n, 1:-1, 1:-1, 1:-1].copy()
        for nb in range(1, 7):
            if tree.neighbors[b, nb - 1] >= 0:
                continue
            d, hi = (nb - 1) // 2, (nb - 1) % 2
            sl = [slice(None)] * 3
            sl[2 - d] = nc - 1 if hi else 0
            tmp[tuple(sl)] += (bc_to_rhs[(int(b), nb)] * bcmap[(int(b), nb)][1]).reshape(nc, nc)
        ilo = (tree.ix[b] - 1) * nc + 1
        H.set_box(H.rhs, ilo, ilo + nc - 1, tmp)
        H.set_box(H.phi, ilo, ilo + nc - 1, np.zeros(nc ** 3))
    H.solve()
    for n, b in enumerate(l1):  # coarse_solver_get_phi (:343-358)
        ilo = (tree.ix[b] - 1) * nc + 1
        got = H.get_box(H.phi, ilo, ilo + nc - 1).reshape(nc, nc, nc)
        assert np.max(np.abs(got - want[n])) <= 1e-12 * np.max(np.abs(want)), (n, np.max(np.abs(got - want[n])))
    H.destroy()


def test_cylindrical_coarse_solve_sequence_matches_the_oracle(hy):
    """2D cylindrical: the reference switches to the non-symmetric 5-entry storage (hypre_set_matrix :110-115) and
    af_stencil_get_box expands the constant stencil with the flux factors (r -+ dr/2) / r (SURVEY a10:
    cc_cyl(2:3) = rfac * c(2:3), cc_cyl(1) = c(1) - (cc_cyl(2) - c(2)) - (cc_cyl(3) - c(3)))."""
    tree = T.build_tree(2, 8, [16, 8], 2, None, coord_t=T.AF_CYL, r_max=[2.0, 1.0])
    nc = tree.nc

    def sides(nb, c):
        if nb == 1:
            return W.AF_BC_NEUMANN, 0.0
        if nb == 2:
            return W.AF_BC_NEUMANN, 0.3 + c[..., 1]
        return W.AF_BC_DIRICHLET, 1.0 + c[..., 0] * (nb == 4)

    bc = W.bc_table(tree, sides)
    orc = Oracle(tree)
    orc.set_bc(bc)
    orc.mg_init()
    l1 = np.asarray(tree.lvl_ids[0], np.int32)
    assert len(l1) == 2
    rng = np.random.default_rng(13)
    rhs = np.zeros((len(l1), nc + 2, nc + 2))
    rhs[:, 1:-1, 1:-1] = rng.standard_normal((len(l1), nc, nc))
    orc.set_cc(M.I_RHS, l1, rhs)
    orc.set_cc(M.I_PHI, l1, np.zeros_like(rhs))
    orc.solve_coarse_grid()
    want = orc.get_cc(M.I_PHI, l1).reshape(rhs.shape)[:, 1:-1, 1:-1]

    nx = [int(v) for v in tree.coarse_grid_size]
    H = Hypre(hy, nx, (0, 0), [1, 2, 3, 4, 5], symmetric=0)
    bcmap = {(int(i), int(n)): (int(t), v) for i, n, t, v in zip(bc.ids, bc.nbs, bc.types, bc.vals)}
    bc_to_rhs = {}
    for b in l1:
        dr = tree.dr[b]
        c = np.array([0.0, 1 / dr[0] ** 2, 1 / dr[0] ** 2, 1 / dr[1] ** 2, 1 / dr[1] ** 2])
        c[0] = -c[1:].sum()
        r = tree.r_min[b, 0] + (np.arange(1, nc + 1) - 0.5) * dr[0]
        full = np.zeros((nc, nc, 5))  # [j, i, entry]
        full[..., 3:] = c[3:]
        full[..., 1] = (r - 0.5 * dr[0]) / r * c[1]
        full[..., 2] = (r + 0.5 * dr[0]) / r * c[2]
        full[..., 0] = c[0] - (full[..., 1] - c[1]) - (full[..., 2] - c[2])
        for nb in range(1, 5):
            if tree.neighbors[b, nb - 1] >= 0:
                continue
            d, hi = (nb - 1) // 2, (nb - 1) % 2
            sl = [slice(None)] * 2
            sl[1 - d] = nc - 1 if hi else 0
            sl = tuple(sl)
            if bcmap[(int(b), nb)][0] == W.AF_BC_DIRICHLET:
                full[sl + (0,)] -= full[sl + (nb,)]
                bc_to_rhs[(int(b), nb)] = (-2 * full[sl + (nb,)]).ravel()
            else:
                full[sl + (0,)] += full[sl + (nb,)]
                bc_to_rhs[(int(b), nb)] = -(full[sl + (nb,)] * dr[d]).ravel() * (1 if hi else -1)
            full[sl + (nb,)] = 0
        ilo = (tree.ix[b] - 1) * nc + 1
        H.set_matrix_box(ilo, ilo + nc - 1, full.reshape(-1, 5))
    H.prepare_solve()
    for n, b in enumerate(l1):
        tmp = rhs[n, 1:-1, 1:-1].copy()
        for nb in range(1, 5):
            if tree.neighbors[b, nb - 1] >= 0:
                continue
            d, hi = (nb - 1) // 2, (nb - 1) % 2
            sl = [slice(None)] * 2
            sl[1 - d] = nc - 1 if hi else 0
            tmp[tuple(sl)] += bc_to_rhs[(int(b), nb)] * bcmap[(int(b), nb)][1]
        ilo = (tree.ix[b] - 1) * nc + 1
        H.set_box(H.rhs, ilo, ilo + nc - 1, tmp)
        H.set_box(H.phi, ilo, ilo + nc - 1, np.zeros(nc ** 2))
    H.solve()
    for n, b in enumerate(l1):
        ilo = (tree.ix[b] - 1) * nc + 1
        got = H.get_box(H.phi, ilo, ilo + nc - 1).reshape(nc, nc)
        assert np.max(np.abs(got - want[n])) <= 1e-12 * np.max(np.abs(want)), (n, np.max(np.abs(got - want[n])))
    H.destroy()

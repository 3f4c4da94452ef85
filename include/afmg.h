/* afmg.h -- C ABI of the B200-native FAS multigrid solver that replaces afivo's m_af_multigrid.
 *
 * The reference exposes no C ABI on this path; its replaceable unit is the set of module-level
 * Fortran entry points of afivo/src/m_af_multigrid.f90 (mg_init :43, mg_destroy :111, mg_use :118,
 * mg_fas_fmg :137, mg_fas_vcycle :185, mg_update_operator_stencil :1188) acting on af_t / box_t
 * (afivo/src/m_af_types.f90:286-393) and mg_t (:572-665).  A Fortran shim module with those same
 * signatures (fortran/m_af_multigrid_gpu.f90) binds the functions below through
 * ISO_C_BINDING.  Every entry point says which reference interface it stands in for.
 *
 * Conventions: all integers are int32, reals are fp64, logicals are int (0/1).  Box ids, level
 * numbers and neighbour directions are the reference's (1-based ids, af_no_box = 0,
 * af_phys_boundary = -1; nb = 1..2*ndim = lowx, highx, lowy, highy, lowz, highz;
 * m_af_types.f90:38-41, 186-214).  Cell data crosses the boundary in the reference's own box layout:
 * cc(0:nc+1, 0:nc+1 [, 0:nc+1]) per box, first index fastest (m_af_core.f90:551).
 * All pointers are HOST pointers unless the function name ends in _device.
 * Every function returns AFMG_OK (0) or a negative error code; nothing aborts the process
 * (the reference uses `error stop`; the shim turns non-zero codes into that).
 * A handle is not re-entrant; use one handle per mg_t.  There is no CPU fallback: without a CUDA
 * device afmg_create fails with AFMG_ERR_CUDA.
 */
#ifndef AFMG_H
#define AFMG_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct afmg_handle afmg_handle;

enum {
  AFMG_OK = 0,
  AFMG_ERR_ARG = -1,          /* invalid argument / inconsistent tree                      */
  AFMG_ERR_CUDA = -2,         /* CUDA runtime error or no device (see afmg_last_error)     */
  AFMG_ERR_UNSUPPORTED = -3,  /* valid in the reference, not (yet) supported here          */
  AFMG_ERR_STATE = -4,        /* call order violated (e.g. solve before afmg_set_tree)     */
  AFMG_ERR_SINGULAR = -5,     /* coarse-grid operator is singular (all-Neumann, lambda=0)   */
  AFMG_ERR_NOT_CONVERGED = -7, /* afmg_field_solve: residual criterion not met                  */
  AFMG_ERR_COMM = -6           /* a peer GPU did not reach a barrier in time               */
};

/* cell-centred variables of the solver: mg%i_phi, mg%i_rhs, mg%i_tmp, tree%mg_i_eps
 * (m_af_types.f90:574-584) */
enum { AFMG_PHI = 0, AFMG_RHS = 1, AFMG_TMP = 2, AFMG_EPS = 3,
       AFMG_FLD = 4 /* cell-centred field norm: i_norm of mg_compute_phi_gradient, the callers' i_electric_fld
                       (src/m_streamer.f90:302) */,
       AFMG_PHOTO = 5 /* photoionization source accumulated by afmg_helmholtz_compute (i_photo) */ };

/* boundary condition types (m_af_types.f90:58-69) */
enum {
  AFMG_BC_DIRICHLET = -10,
  AFMG_BC_NEUMANN = -11,
  AFMG_BC_CONTINUOUS = -12,
  AFMG_BC_DIRICHLET_COPY = -13
};

/* mg%prolongation_type (m_af_types.f90:523-544) */
enum { AFMG_PROLONG_LINEAR = 17, AFMG_PROLONG_SPARSE = 18, AFMG_PROLONG_AUTO = 19 };

/* coordinate systems (m_af_types.f90:45-48) */
enum { AFMG_XYZ = 1, AFMG_CYL = 2 };

/* POD copy of the scalar members of mg_t (m_af_types.f90:572-665) and of the af_t members the
 * solver needs (m_af_types.f90:326-393).  Filled by mg_init in the shim. */
typedef struct afmg_opts {
  int32_t ndim;                 /* NDIM (2 or 3)                                              */
  int32_t n_cell;               /* tree%n_cell: cells per box side (even)                     */
  int32_t coord_t;              /* tree%coord_t: AFMG_XYZ / AFMG_CYL                          */
  int32_t n_cycle_down;         /* mg%n_cycle_down (default 2)                                */
  int32_t n_cycle_up;           /* mg%n_cycle_up (default 2)                                  */
  int32_t use_corners;          /* mg%use_corners                                             */
  int32_t subtract_mean;        /* mg%subtract_mean                                           */
  int32_t prolongation_type;    /* mg%prolongation_type                                       */
  int32_t operator_mask;        /* mg%operator_mask (-1 = all bits)                           */
  int32_t has_eps;              /* tree%mg_i_eps > 0                                          */
  int32_t device;               /* CUDA device ordinal, -1 = current device                   */
  int32_t n_gpus;               /* 0 / 1: this handle drives ONE GPU (`device`).  N > 1: single-process multi-GPU --
                                 * the handle drives devices device .. device + N - 1 (device = -1: 0 .. N - 1) of one
                                 * NVLink domain with one host thread per GPU inside the library, boxes partitioned as
                                 * by afmg_partition_min, halos through peer memory (cudaDeviceEnablePeerAccess); every
                                 * call of this API then acts on the whole tree, exactly as with N = 1.  This is the mode
                                 * for a single-process caller such as the reference (one OpenMP process,
                                 * afivo/documentation/parallelization.md).  0 also reads AFMG_N_GPUS.               */
  double helmholtz_lambda;      /* mg%helmholtz_lambda (lambda^2 of L phi - lambda phi = f)   */
  double lsf_boundary_value;    /* mg%lsf_boundary_value                                      */
  int32_t coarse_grid_size[3];  /* tree%coarse_grid_size (cells)                              */
  int32_t periodic[3];          /* tree%periodic                                              */
  double dr_base[3];            /* tree%dr_base                                               */
  double r_base[3];             /* tree%r_base                                                */
} afmg_opts;

/* Borrowed, read-only flat copy of the tree topology (box_t members lvl, ix, parent, children,
 * neighbors, neighbor_mat, m_af_types.f90:286-300; level lists lvls(l)%ids, :76-80).
 * Per-box arrays have (highest_id + 1) rows; row 0 is unused, so reference ids index directly.
 * lvl_ids concatenates lvls(1)%ids ... lvls(highest_lvl)%ids, lvl_counts gives their sizes. */
typedef struct afmg_tree {
  int32_t highest_lvl;
  int32_t highest_id;
  const int32_t* lvl_counts;   /* (highest_lvl)                                               */
  const int32_t* lvl_ids;      /* (sum lvl_counts)                                            */
  const int32_t* lvl;          /* (n+1)                                                       */
  const int32_t* ix;           /* (n+1, ndim)  1-based spatial index on the level grid        */
  const int32_t* parent;       /* (n+1)                                                       */
  const int32_t* children;     /* (n+1, 2^ndim)                                               */
  const int32_t* neighbors;    /* (n+1, 2*ndim)                                               */
  const int32_t* neighbor_mat; /* (n+1, 3^ndim), first offset fastest                         */
  const double* r_min;         /* (n+1, ndim)  box%r_min (used by cylindrical stencils)       */
} afmg_tree;

/* ---- lifecycle: mg_init (m_af_multigrid.f90:43-109), mg_destroy (:111-115) -------------------- */
int afmg_create(afmg_handle** out, const afmg_opts* opts);
int afmg_destroy(afmg_handle* h);
/* text of the last error raised on this handle (or by afmg_create when h == NULL) */
const char* afmg_last_error(const afmg_handle* h);

/* ---- topology: called after mg_init and whenever af_adjust_refinement (m_af_core.f90:697) changed
 * the tree.  Copies the arrays; rebuilds slot maps, ghost-cell plans, constant stencils
 * (mg_set_operators_tree, m_af_multigrid.f90:1216-1224) and the coarse-grid factorisation
 * (coarse_solver_initialize, m_coarse_solver.f90:71-194).  Cell data on the device is reset to zero: the
 * caller uploads phi (the previous solution, prolonged onto new boxes by the reference's refinement) and rhs
 * again, as the shim's mg_gpu_fas_* do before every solve. */
int afmg_set_tree(afmg_handle* h, const afmg_tree* tree);

/* ---- boundary conditions as data: one row per physical face.  Replaces the mg%sides_bc callback
 * (m_af_types.f90:401-420; af_bc_dirichlet_zero etc. m_af_ghostcell.f90:615-652;
 * field_bc_homogeneous src/m_field.f90:590-610): the shim evaluates the callback on the host for
 * every face with neighbors(nb) == af_phys_boundary.  bc_val has n_faces rows of nc^(ndim-1) values,
 * indexed like bc_val in bc_to_gc (m_af_ghostcell.f90:173-279).  May be called before every solve
 * (e.g. when the applied voltage changed). */
int afmg_set_bc(afmg_handle* h, int32_t n_faces, const int32_t* box_id, const int32_t* nb,
                const int32_t* bc_type, const double* bc_val);

/* ---- operator updates: mg_update_operator_stencil (m_af_multigrid.f90:1188-1214) and the scalar
 * members that enter the stencils */
int afmg_set_helmholtz_lambda(afmg_handle* h, double lambda);
int afmg_set_lsf_boundary_value(afmg_handle* h, double value);
/* mg%lsf_boundary_function as data (m_af_types.f90:628; several electrodes at different potentials, e.g.
 * rod_rod_get_potential src/m_field.f90:802-825): values holds, for each listed box, mg_lsf_boundary_value(box, mg)
 * (m_coarse_solver.f90:493-510) = the function at the nc^ndim cell centres (first index fastest).  They replace the
 * scalar mg%lsf_boundary_value in bc_correction = f * value (m_af_multigrid.f90:1171-1174), in the coarse-grid
 * right-hand side (m_coarse_solver.f90:320-326) and in mg_box_lpllsf_gradient; boxes not listed keep the scalar.
 * Call it with the same box list before every solve whose electrode potentials changed (values only are updated);
 * n = 0 returns to the scalar everywhere.  afmg_set_tree drops the list. */
int afmg_set_lsf_boundary_values(afmg_handle* h, int32_t n, const int32_t* box_id, const double* values);
/* Rebuild the implicit constant stencils and the coarse-grid factorisation after lambda changed
 * (mg_update_operator_stencil, m_af_multigrid.f90:1188-1214); explicit stencils are re-shipped with
 * afmg_set_stencils by the shim. */
int afmg_update_operator_stencil(afmg_handle* h);

/* ---- operator / prolongation stencils as data.  The builders stay on the host in the reference
 * (mg_set_operators_lvl m_af_multigrid.f90:1147-1185: mg_box_lpld_stencil :1493-1532 for variable eps,
 * mg_box_lsf_stencil :1782-1854 for level-set boxes, mg_box_prolong_eps_stencil :1308-1388; they call user
 * callbacks), so the shim ships what they stored in box%stencils (stencil_t, m_af_types.f90:260-282) for
 * every box whose operator is not the plain constant Laplacian or whose prolongation is not the default:
 *   op_stype       0 = implicit (library derives 1/dr^2, lambda), 1 = stencil_constant: c(2*ndim+1) at
 *                  op_offset, 2 = stencil_variable: v(2*ndim+1, nc, nc[, nc]) at op_offset (first index fastest)
 *   f_offset       >= 0: stencil%f (nc^3) -- the library applies bc_correction = f * lsf_boundary_value
 *                  (m_af_multigrid.f90:1171-1174); -1: none
 *   prolong_shape  0 = default (mg%prolongation_type), AFMG_STENCIL_P248 (2^ndim coefficients) or
 *                  AFMG_STENCIL_P234 (ndim+1); prolong_stype 1 = constant c(n) / 2 = variable v(ndim+1, cells)
 *   tag            box%tag (mg_lsf_box = 1, mg_veps_box = 2, mg_ceps_box = 4, m_af_types.f90:497-508); boxes
 *                  with iand(tag, operator_mask) == mg_veps_box get mg_sides_rb_extrap ghost cells on
 *                  refinement boundaries (mg_auto_rb, :926-940)
 * Offsets count doubles inside coeff_blob.  The call replaces all previously shipped stencils; boxes
 * not listed use the implicit Laplacian and the default prolongation.  n = 0 clears. */
enum { AFMG_STENCIL_P234 = 2, AFMG_STENCIL_P248 = 3 };
enum { AFMG_TAG_LSF_BOX = 1, AFMG_TAG_VEPS_BOX = 2, AFMG_TAG_CEPS_BOX = 4 };
typedef struct afmg_stencil_desc {
  int32_t box_id;
  int32_t op_stype;
  int32_t prolong_shape;
  int32_t prolong_stype;
  int32_t tag;
  int32_t cylindrical_gradient; /* stencil%cylindrical_gradient (2D cylindrical trees, m_af_types.f90:272) */
  int64_t op_offset;
  int64_t f_offset;
  int64_t prolong_offset;
} afmg_stencil_desc;
int afmg_set_stencils(afmg_handle* h, int32_t n, const afmg_stencil_desc* desc, const double* coeff_blob,
                      int64_t blob_len);

/* ---- host-side stencil builders: what mg_set_operators_lvl (m_af_multigrid.f90:1147-1185) stores in
 * box%stencils, for hosts that are not the Fortran reference (include/afmg.hpp, afivo_streamer_b200/stencils.py).  The
 * Fortran shim does not need them: it ships the stencils the reference itself built.  Pure host functions (no handle,
 * no device, usable without a GPU); same expression order as the reference, results identical bit for bit.  Cell
 * arrays: cc = (nc+2)^ndim doubles incl. ghost cells, per-cell outputs in IJK order (first index fastest).  Sizes and
 * exact meaning: see afivo_streamer_b200/csrc/afmg_builders.inc. */
/* mg%lsf: the level-set function (m_af_types.f90:667-722 mg_func_lsf), user = caller context */
typedef double (*afmg_lsf_fn)(const double* r, void* user);
enum { AFMG_LSF_DIST_LINEAR = 0, AFMG_LSF_DIST_GSS = 1 }; /* mg%lsf_dist => mg_lsf_dist_linear (:1627) / _gss (:1651) */
typedef struct afmg_lsf_opts {
  int32_t dist_method;
  double gradient_safety_factor; /* mg%lsf_gradient_safety_factor = 1.5   (m_af_types.f90:607) */
  double length_scale;           /* mg%lsf_length_scale = 1e100           (:610)               */
  double tol;                    /* mg%lsf_tol = 1e-8                     (:613)               */
  double min_rel_distance;       /* mg%lsf_min_rel_distance = 1e-4        (:616)               */
} afmg_lsf_opts;
void afmg_lsf_opts_default(afmg_lsf_opts* o);
/* mg_set_box_tag (:1100-1145): returns the tag (>= 0) or AFMG_ERR_ARG; eps_cc may be NULL (no mg_i_eps) */
int32_t afmg_build_box_tag(int32_t ndim, int32_t nc, const double* eps_cc, int32_t has_lsf);
/* mg_store_operator_stencil (:823-859) of a box with iand(tag, operator_mask) /= mg_normal_box: mg_box_lpld_stencil
 * (:1493-1532), mg_box_lsf_stencil (:1782-1854), mg_box_lpld_lsf_stencil (:1535-1623) */
int afmg_build_box_operator(int32_t ndim, int32_t nc, int32_t coord_t, int32_t masked_tag, const double* dr,
                            const double* r_min, const double* eps_cc, const double* lsf_dd, double* v, double* f,
                            int32_t* stype, int32_t* has_f, int32_t* cylindrical_gradient);
/* mg_store_prolongation_stencil (:862-903), variable cases: mg_box_prolong_eps_stencil (:1308-1388) or
 * mg_box_prolong_lsf_stencil (:1392-1482) */
int afmg_build_box_prolongation(int32_t ndim, int32_t nc, int32_t masked_tag, const int32_t* ix,
                                const double* eps_parent_cc, const double* lsf_pdd, double* v, int32_t* stype,
                                int32_t* shape);
/* get_possible_lsf_root_mask + store_lsf_distance_matrix (:954-1097) with mg_lsf_dist_linear / _gss, bisection, gss
 * and numerical_gradient (:1627-1776, :2164-2190) */
int afmg_build_box_lsf_distances(int32_t ndim, int32_t nc, const double* r_min, const double* dr, afmg_lsf_fn lsf,
                                 void* user, const afmg_lsf_opts* opts, const double* lsf_cc, uint8_t* root_mask,
                                 double* dd, int32_t* n_boundary);
/* the coarse-point distances mg_box_prolong_lsf_stencil evaluates (:1392-1482) */
int afmg_build_box_lsf_prolong_distances(int32_t ndim, int32_t nc, const double* r_min, const double* dr,
                                         const int32_t* ix, const double* r_min_p, const double* dr_p, afmg_lsf_fn lsf,
                                         void* user, const afmg_lsf_opts* opts, const uint8_t* root_mask, double* pdd);

/* The built-in electrode shapes of the streamer code (field_electrode_type, src/m_field.f90:254-362; level-set
 * functions :686-904 on GM_dist_vec_line, src/m_geometry.f90:23-51) as afmg_lsf_fn-compatible host functions, so that
 * the distance search above runs without a callback into the host language.  Fill the parameters the shape uses (the
 * reference's field_rod_r0, field_rod_r1, field_rod_radius, cone_tip_radius, cone_length_frac, the same for rod2;
 * domain_center(1:2) for coaxial), call afmg_electrode_prepare (argument checks = the reference's error stops; derived
 * cone parameters, get_conical_rod_properties :698-719), then pass afmg_electrode_lsf with user = the struct.
 * afmg_electrode_potential is mg%lsf_boundary_function (*_get_potential), or the scalar mg%lsf_boundary_value of the
 * single-electrode shapes (:482-487). */
enum {
  AFMG_ELECTRODE_SPHERE = 1,
  AFMG_ELECTRODE_ROD = 2,
  AFMG_ELECTRODE_ROD_CONE_TOP = 3,
  AFMG_ELECTRODE_ROD_ROD = 4,
  AFMG_ELECTRODE_SPHERE_ROD = 5,
  AFMG_ELECTRODE_TWO_ROD_CONE = 6,
  AFMG_ELECTRODE_COAXIAL = 7
};
typedef struct afmg_electrode {
  int32_t type, ndim;
  int32_t electrode_grounded, electrode2_grounded;
  double current_voltage;
  double rod_r0[3], rod_r1[3], rod_radius, cone_tip_radius, cone_length_frac;
  double rod2_r0[3], rod2_r1[3], rod2_radius, cone2_tip_radius, cone2_length_frac;
  double domain_center[3];
  /* derived by afmg_electrode_prepare */
  double cone_tip_center[3], cone_tip_r_curvature, cone2_tip_center[3], cone2_tip_r_curvature;
} afmg_electrode;
int afmg_electrode_prepare(afmg_electrode* e);
/* mg_set_operators_tree ON THE DEVICE (csrc/builders_dev.cuh): box tags, operator and prolongation
 * stencils of every box from the permittivity resident on the device (afmg_upload(AFMG_EPS, all boxes incl. ghost cells)
 * and / or one of the built-in electrode shapes above (NULL: none), with the level-set options of mg_t (NULL: defaults),
 * followed by what afmg_set_stencils and afmg_set_lsf_distances do -- but no coefficient crosses PCIe.  The
 * arithmetic is the host builders' (same per-cell functions), hence the reference's.  3D, single-GPU handles;
 * mg%lsf_use_custom_prolongation and user level-set callbacks stay with the host builders.  Errors like the reference:
 * an electrode that level 1 does not resolve (check_coarse_representation_lsf) is AFMG_ERR_ARG. */
int afmg_build_stencils_device(afmg_handle* h, const afmg_electrode* electrode, const afmg_lsf_opts* lsf_opts);
/* the records it built, in the reference's order, for checks against afmg_build_box_* (see csrc/afmg.cu) */
int afmg_built_stencils(afmg_handle* h, int32_t* n, int32_t* box_id, int32_t* tag, int32_t* meta, double* blob);
double afmg_electrode_lsf(const double* r, void* electrode);
double afmg_electrode_potential(const double* r, void* electrode);

/* ---- cell data: box%cc(:, :, :, iv) of n boxes, (nc+2)^ndim doubles each, packed in the order of
 * `box_id` (m_af_types.f90:302).  upload/download take host memory; the _device variants take
 * device memory (for callers that keep rhs / phi resident). */
int afmg_upload(afmg_handle* h, int32_t var, int32_t n, const int32_t* box_id, const double* packed);
int afmg_download(afmg_handle* h, int32_t var, int32_t n, const int32_t* box_id, double* packed);
/* upload of interior cells only: n boxes of nc^ndim doubles, cc(1:nc, 1:nc[, 1:nc]); ghost cells on the
 * device are left as they are.  For the right-hand side, whose ghost cells are never read (callers set rhs
 * on the interior of leaves only, src/m_field.f90:422-435): 30 % fewer bytes over PCIe for nc = 16. */
int afmg_upload_interior(afmg_handle* h, int32_t var, int32_t n, const int32_t* box_id, const double* packed);
/* download of interior cells only (nc^ndim doubles per box): what a caller needs when it refills the ghost cells
 * itself or only reads cell centres (af_loop_box kernels over cc(1:nc, ...), e.g. the particle / fluid updates that
 * follow field_compute); the ghost cells of phi stay valid on the device. */
int afmg_download_interior(afmg_handle* h, int32_t var, int32_t n, const int32_t* box_id, double* packed);
/* Page-locked host buffers for `packed`: upload / download split a transfer into chunks and overlap the PCIe copy
 * of one chunk with the pack / unpack kernel of the previous one; with page-locked memory the copies are DMA
 * transfers at the full PCIe rate (pageable memory is staged by the driver).  The Fortran shim packs box%cc into
 * such a buffer (c_f_pointer) instead of an allocatable temporary.  NULL on failure. */
void* afmg_host_alloc(size_t bytes);
void afmg_host_free(void* p);
int afmg_upload_device(afmg_handle* h, int32_t var, int32_t n, const int32_t* box_id_host,
                       const double* packed_device);
int afmg_download_device(afmg_handle* h, int32_t var, int32_t n, const int32_t* box_id_host,
                         double* packed_device);
/* field_set_rhs (src/m_field.f90:406-444) with the sum on the device: rhs = 0, then rhs = rhs + charges[s] *
 * densities[s] for s = 0 .. n_species-1 on the complete records of the listed boxes (the reference loops over the leaves),
 * in that order, so the bits are the reference's.  charges[s] = charged_species_charge(s) * (-UC_elem_charge / UC_eps0);
 * densities[s] = packed cc(:, :, :, charged_species_itree(s) + s_in) of the boxes in the order of box_id, (nc+2)^ndim
 * doubles each, in host memory (page-locked for speed) or, with on_device != 0, in device memory -- the case this
 * exists for: a caller that keeps its densities resident ships nothing per solve.  The surface-charge term of
 * dielectrics (surface_surface_charge_to_rhs) stays with the caller: pass it as one more "species" with charge 1. */
int afmg_field_set_rhs(afmg_handle* h, int32_t n, const int32_t* box_id, int32_t n_species, const double* charges,
                       const double* const* densities, int32_t on_device);
/* af_tree_clear_cc / af_box_clear_cc (m_af_utils.f90:385): set a variable to zero on all boxes */
int afmg_clear(afmg_handle* h, int32_t var);

/* ---- the two solver entry points -------------------------------------------------------------
 * mg_fas_fmg(tree, mg, set_residual, have_guess)               m_af_multigrid.f90:137-180
 * mg_fas_vcycle(tree, mg, set_residual, highest_lvl, standalone) m_af_multigrid.f90:185-264
 * highest_lvl <= 0 means tree%highest_lvl.  Blocking: return after the device finished. */
int afmg_fas_fmg(afmg_handle* h, int32_t set_residual, int32_t have_guess);
int afmg_fas_vcycle(afmg_handle* h, int32_t set_residual, int32_t highest_lvl, int32_t standalone);
/* Asynchronous variants: enqueue n_cycles back-to-back cycles on the handle's stream and return;
 * afmg_sync waits.  Used by callers that keep data resident and test convergence every few cycles
 * (field_compute, src/m_field.f90:491-524). */
int afmg_fas_fmg_async(afmg_handle* h, int32_t set_residual, int32_t have_guess, int32_t n_cycles);
int afmg_fas_vcycle_async(afmg_handle* h, int32_t set_residual, int32_t highest_lvl, int32_t n_cycles);
int afmg_sync(afmg_handle* h);

/* ---- the caller's convergence loop, run next to the device.  field_compute
 * (src/m_field.f90:491-524): without a guess, repeat mg_fas_fmg(set_residual = T, have_guess = T) up to
 * max_fmg times until the max-norm of the residual over the leaves is below residual_threshold, or has
 * stalled (ratio of min / max over the last three cycles within (0.5, 2)) below max_residual; then up to
 * n_vcycles mg_fas_vcycle(set_residual = T), stopping at the threshold.  Only the 8-byte max-norm crosses
 * PCIe per cycle.  residuals receives the max-norms in order (capacity max_fmg + n_vcycles); n_fmg / n_vc
 * the number of cycles run.  Returns AFMG_ERR_NOT_CONVERGED if the FMG loop ends without meeting either
 * criterion (the reference: error stop "No convergence in initial field computation"). */
int afmg_field_solve(afmg_handle* h, int32_t have_guess, double residual_threshold, double max_residual,
                     int32_t max_fmg, int32_t n_vcycles, double* residuals, int32_t* n_fmg, int32_t* n_vc);

/* ---- field from potential on the device -------------------------------------
 * The step every caller takes right after a solve (field_from_potential, src/m_field.f90:531-548).
 * Extra device variables are allocated on first use and dropped by afmg_set_tree:
 *   fc       box%fc(nc+1, nc+1[, nc+1], NDIM, i_fc) (m_af_core.f90:552), one face-centred variable, moved with
 *            afmg_upload_fc / afmg_download_fc in exactly that record (first index fastest);
 *   AFMG_FLD the cell-centred norm, AFMG_EPS tree%mg_i_eps (needed only when variable-eps boxes exist; default
 *            1), both moved with afmg_upload / afmg_download in the box layout cc(0:nc+1, ...).
 * On a multi-GPU handle every rank computes / moves the boxes it owns; the norm's halo travels through peer memory.
 *
 * mg_compute_phi_gradient(tree, mg, i_fc, fac, i_norm) m_af_multigrid.f90:1857-1898: fc = fac/dr * (phi difference)
 * on every box (mg_box_lpl_gradient :1901-1999, with the eps-weighted boundary faces of mg_veps_box boxes),
 * mg_box_lpllsf_gradient (:2055-2137) on leaves listed by afmg_set_lsf_distances, and, if with_norm, the norm
 * (mg_box_field_norm :2023-2051) on the interior of every box. */
int afmg_compute_phi_gradient(afmg_handle* h, double fac, int32_t with_norm);
/* mg_compute_field_norm (m_af_multigrid.f90:2002-2020) from the current fc, e.g. after the host corrected fc for
 * surface charge (surface_correct_field_fc, src/m_field.f90:537-541) and sent it back with afmg_upload_fc */
int afmg_compute_field_norm(afmg_handle* h);
/* af_gc_tree(tree, [var], corners) m_af_ghostcell.f90:25-46.  AFMG_PHI: the multigrid's own methods
 * (mg%sides_bc, mg_sides_rb).  AFMG_FLD: the methods the callers register for the field norm
 * (af_set_cc_methods(tree, i_electric_fld, af_bc_neumann_zero, af_gc_interp), src/m_field.f90:392-393):
 * af_gc_interp (m_af_ghostcell.f90:394-498) on refinement boundaries and the boundary condition given by
 * afmg_set_fld_bc (default af_bc_neumann_zero). */
int afmg_gc_tree(afmg_handle* h, int32_t var, int32_t corners);
/* field_from_potential without dielectric (src/m_field.f90:543-547): gradient with norm, then
 * af_gc_tree(tree, [i_electric_fld]) */
int afmg_field_from_potential(afmg_handle* h, double fac);
/* boundary condition of AFMG_FLD on physical faces, same row format as afmg_set_bc */
int afmg_set_fld_bc(afmg_handle* h, int32_t n_faces, const int32_t* box_id, const int32_t* nb, const int32_t* bc_type,
                    const double* bc_val);
/* The sparse level-set distance stencils (mg_lsf_distance_key, store_lsf_distance_matrix
 * m_af_multigrid.f90:977-1097) of n_boxes boxes: box b has n_entries[b] entries, stored back to back in the
 * order of stencil%sparse_ix: cell_ix (ndim per entry, 1-based), dd = sparse_v (2*ndim per entry) and lsf =
 * cc(IJK, mg%i_lsf) of that cell (its sign decides whether the face is corrected; NULL: all >= 0).  The
 * boundary value is mg%lsf_boundary_value (afmg_set_lsf_boundary_value).  Replaces the previous list. */
int afmg_set_lsf_distances(afmg_handle* h, int32_t n_boxes, const int32_t* box_id, const int32_t* n_entries,
                           const int32_t* cell_ix, const double* dd, const double* lsf);
int afmg_upload_fc(afmg_handle* h, int32_t n, const int32_t* box_id, const double* packed);
int afmg_download_fc(afmg_handle* h, int32_t n, const int32_t* box_id, double* packed);

/* ---- Helmholtz photoionization on the device: photoi_helmh_compute
 * (src/m_photoi_helmh.f90:162-204).  modes[n] is the handle of mg_helm(n) (own helmholtz_lambda, own phi =
 * i_modes(n), boundary conditions photoi_helmh_bc shipped with afmg_set_bc), all on the same tree and
 * device; the right-hand side is the AFMG_RHS of modes[0] (leaves) and is copied device to device to the
 * other modes, as i_rhs is one shared variable in the reference.  Per mode: up to max_fmg_cycles
 * mg_fas_fmg(set_residual = T, have_guess = T) until max|residual| / max(max|rhs|, sqrt(epsilon)) <
 * max_rel_residual, then i_photo = i_photo - coeffs(n) * i_modes(n) on the leaves.  The result is AFMG_PHOTO of
 * modes[0] (afmg_download).  n_cycles / residuals (may be NULL) receive the FMG count and last residual
 * max-norm of every mode.  Only three doubles per cycle cross PCIe. */
int afmg_helmholtz_compute(afmg_handle* const* modes, int32_t n_modes, const double* coeffs, int32_t max_fmg_cycles,
                           double max_rel_residual, int32_t* n_cycles, double* residuals);

/* ---- single operations (the mg_t per-level building blocks; exported for parity tests) -------- */
int afmg_gsrb_boxes(afmg_handle* h, int32_t lvl, int32_t type_cycle /*1 = down, 3 = up*/); /* :648-687 */
int afmg_gsrb_halfsweep(afmg_handle* h, int32_t lvl, int32_t redblack);   /* mg%box_gsrb on a level + side ghost fill */
int afmg_gc_lvl(afmg_handle* h, int32_t lvl, int32_t var, int32_t corners); /* af_gc_lvl, m_af_ghostcell.f90:49-61 */
int afmg_update_coarse(afmg_handle* h, int32_t lvl, int32_t with_tmp);    /* :691-738 / :742-776 */
int afmg_correct_children(afmg_handle* h, int32_t lvl_parents);           /* :624-646 */
/* correct_children(lvl_parents) followed by af_gc_lvl(lvl_parents + 1), as in the cycles (:219-222) */
int afmg_correct_children_gc(afmg_handle* h, int32_t lvl_parents);
int afmg_residual_lvl(afmg_handle* h, int32_t lvl);                       /* residual_box :801-810 */
int afmg_solve_coarse_grid(afmg_handle* h);                               /* :266-291 */
int afmg_init_phi_rhs(afmg_handle* h);                                    /* :779-799 */

/* ---- reductions: af_tree_maxabs_cc (m_af_utils.f90:773-785, leaves, interior cells) and
 * af_tree_sum_cc (m_af_utils.f90:966-1027, volume-weighted leaf sum) */
int afmg_max_abs(afmg_handle* h, int32_t var, double* out);
int afmg_tree_sum(afmg_handle* h, int32_t var, double* out);
/* Order-independent bitwise checksum of a variable over the complete records (interior and ghost cells) of the
 * boxes this rank owns: wrapping sum and XOR of the 64-bit patterns.  Combining the ranks' values (sum mod 2^64,
 * XOR) gives the checksum of the whole tree, equal for every number of GPUs when the solves are bit-identical
 * (the reference has no such routine; it is the evidence tool for "ghost-cell and index mapping bit-exact"). */
int afmg_checksum(afmg_handle* h, int32_t var, uint64_t* sum_out, uint64_t* xor_out);

/* ---- instrumentation ------------------------------------------------------------------------- */
/* number of kernel launches the handle issued since creation (graph nodes counted per replay) */
int64_t afmg_kernel_launches(const afmg_handle* h);
/* device time in ms of the most recent (group of) cycle(s), measured with CUDA events on the
 * handle's stream */
int afmg_last_cycle_ms(afmg_handle* h, double* ms);
/* per-kernel-family accumulated device time; fills up to cap entries, returns count via *n.
 * Only collected while profiling is enabled (it serialises launches). */
/* Persistent-kernel segments (csrc/mega.cuh): operations on levels of at most `max_boxes` boxes (0 = default: 8192
 * for 8^3 boxes, 1024 for 16^3) run inside ONE cooperatively launched kernel with grid-wide barriers between them
 * instead of one launch each -- the launch-bound regime of the streamer trees (nc = 8, many small levels).  Results are
 * bit-identical to the launch path.  On by default for 3D, single-GPU handles without explicit stencils; AFMG_MEGA=0
 * or enabled = 0 selects the launch path.  afmg_mega_active: number of CTAs of the persistent kernel, 0 = not used. */
/* Device memory of the cell-data slab (phi, rhs, tmp [, field norm]) per GPU of the handle: bytes physically mapped and
 * bytes of the full slot space.  A single-process multi-GPU handle (afmg_opts.n_gpus) reserves the slot space as
 * virtual addresses and maps memory only under the boxes each GPU owns (~ 1 / N of the tree per GPU, so a tree larger
 * than one GPU's memory can be partitioned); single-GPU and multi-process handles map all of it. */
int afmg_slab_bytes(afmg_handle* h, int32_t cap, int64_t* mapped, int64_t* full, int32_t* n);
int afmg_set_mega(afmg_handle* h, int32_t enabled, int32_t max_boxes);
int32_t afmg_mega_active(const afmg_handle* h);
int afmg_set_profiling(afmg_handle* h, int32_t on);
int afmg_profile(afmg_handle* h, int32_t cap, char (*names)[32], double* ms, int64_t* calls, int32_t* n);
/* cell-updates performed by one V-cycle to highest_lvl (<=0: all) and by one FMG */
int afmg_cell_updates(afmg_handle* h, int32_t highest_lvl, int32_t fmg, double* out);

/* ---- device layout (exported so host-side tests can check the index maps without a GPU) -------
 * Offset (in doubles) of cell (i, j, k), 0 <= i,j,k <= nc+1 (k ignored in 2D), inside the device
 * box record; a bijection onto 0 .. (nc+2)^ndim - 1.  See DESIGN.md "data layout". */
int32_t afmg_layout_offset(int32_t ndim, int32_t nc, int32_t i, int32_t j, int32_t k);
int32_t afmg_layout_box_len(int32_t ndim, int32_t nc);
/* Morton (Z-order) key, x in the lowest bit as in afivo's m_morton (morton_from_ix2 / _ix3; known answers
 * afivo/tests/answers/test_morton_2d, _3d).  Boxes of a level are stored, and cut into per-GPU ranges, in the order
 * of this key of box%ix - 1. */
int64_t afmg_morton_key(int32_t ndim, int32_t ix, int32_t iy, int32_t iz);
/* slot (position in the device arrays) of a box id, -1 if unknown */
int32_t afmg_slot_of_box(const afmg_handle* h, int32_t box_id);

/* ---- multi-GPU: one process per GPU of one NVLink / NVSwitch domain.  Boxes are
 * partitioned by contiguous Morton ranges per level (afmg_partition); every rank passes the same full
 * tree and boundary conditions, computes only the boxes it owns, and reads / writes the halo boxes of
 * its peers directly through NVLink peer memory (CUDA IPC): ghost-cell pushes, restriction into and
 * prolongation from remote parents and the max-norm reduction are fused into the kernels, with
 * device-side barriers in place of collectives.  Call order on every rank:
 *     afmg_create; afmg_comm_init(h, n, rank); afmg_set_tree; afmg_comm_export(h, blob);
 *     <all-gather the n blobs with MPI / torch.distributed / ...>; afmg_comm_connect(h, blobs); ...
 * and again export / all-gather / connect after every later afmg_set_tree.  All ranks must then make
 * the same sequence of solver calls.  upload / download only touch boxes owned by the calling rank
 * (other boxes' records in `packed` are ignored / left untouched); afmg_max_abs and afmg_tree_sum
 * return the global value on every rank.  A peer that stops responding makes the calls of the others
 * fail with AFMG_ERR_COMM after AFMG_BARRIER_TIMEOUT_S (default 30) seconds instead of hanging. */
#define AFMG_MAX_RANKS 8
#define AFMG_COMM_BLOB_BYTES 192
int afmg_comm_init(afmg_handle* h, int32_t n_ranks, int32_t rank);
int afmg_comm_export(afmg_handle* h, void* blob /* AFMG_COMM_BLOB_BYTES */);
int afmg_comm_connect(afmg_handle* h, const void* blobs /* n_ranks * AFMG_COMM_BLOB_BYTES, rank order */);
/* which rank owns a box (-1 if unknown) */
int32_t afmg_owner_of_box(const afmg_handle* h, int32_t box_id);
/* The partition rule itself (pure host function, no handle or GPU needed): cuts has
 * highest_lvl * (n_ranks + 1) entries; rank r owns positions [cuts[(l-1)*(n_ranks+1)+r],
 * cuts[(l-1)*(n_ranks+1)+r+1]) of level l's boxes in Morton order of box%ix. */
int afmg_partition(int32_t n_ranks, int32_t highest_lvl, const int32_t* lvl_counts, int32_t* cuts);
/* The rule the handles use: as afmg_partition, but levels with fewer than min_split_boxes boxes are not split -- they
 * stay on rank 0 next to the coarse grid, and operations between such levels need no cross-GPU barrier (default of a
 * handle: 4 Mi cells worth of boxes, i.e. 1024 boxes of 16^3 or 8192 of 8^3; AFMG_MIN_SPLIT_BOXES overrides). */
int afmg_partition_min(int32_t n_ranks, int32_t highest_lvl, const int32_t* lvl_counts, int32_t min_split_boxes,
                       int32_t* cuts);

#ifdef __cplusplus
}
#endif
#endif /* AFMG_H */

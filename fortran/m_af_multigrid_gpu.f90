#include "cpp_macros.h"
!> Drop-in replacement for the solver entry points of afivo's m_af_multigrid
!> (afivo/src/m_af_multigrid.f90:43, :111, :137, :185, :1188) that forwards to libafmg.so through
!> ISO_C_BINDING.  Same signatures; the tree (af_t / box_t) stays the reference's own.
!>
!> What happens on every solve (mg_gpu_fas_fmg / mg_gpu_fas_vcycle / mg_gpu_field_solve), in this order:
!>   1. mg_use(tree, mg) of the reference: boxes created by af_adjust_refinement (tag == af_init_tag) are tagged and
!>      get their operator / prolongation stencils on the host (m_af_multigrid.f90:118-126, 1147-1185);
!>   2. topology -> afmg_set_tree, only when a hash of the complete box lists / parents / children / neighbours
!>      changed (or after mg_gpu_invalidate); with it the explicit stencils are re-shipped and phi is marked stale;
!>   3. mg%sides_bc evaluated for every physical face -> afmg_set_bc, and mg%lsf_boundary_value ->
!>      afmg_set_lsf_boundary_value: EVERY solve, because both carry the applied voltage, which changes in time
!>      (src/m_field.f90:481-487, 590-610).  Neither invalidates the cached cycle graphs when only values change;
!>   4. rhs of the leaves up; phi up only if the device copy is stale (tree changed, mg_gpu_mark_phi_dirty);
!>   5. the cycle(s); phi down (with ghost cells: the reference's post-condition), i_tmp down when set_residual
!>      (switch off with mg_gpu_set_download_tmp when only the max-norm is needed: mg_gpu_tree_maxabs_tmp).
!> Packed buffers are page-locked memory from afmg_host_alloc, kept between calls.
!>
!> NOTE: neither this image nor the GPU box has a Fortran compiler (profiles/r02a_fortran_probe.txt), so this file is
!> shipped as source and has not been compiled.  Build: add it to afivo/src (it needs cpp_macros.h like every afivo
!> source that uses DTIMES / NDIM), list it in afivo/src/definitions.make after m_af_multigrid, link with -lafmg.
module m_af_multigrid_gpu
  use iso_c_binding
  use m_af_types
  use m_af_stencil, only: af_stencil_index, af_stencil_none, stencil_constant, stencil_variable
  use m_af_multigrid, only: mg_use
  implicit none
  private

  type, bind(c) :: afmg_opts
     integer(c_int32_t) :: ndim, n_cell, coord_t, n_cycle_down, n_cycle_up
     integer(c_int32_t) :: use_corners, subtract_mean, prolongation_type, operator_mask
     integer(c_int32_t) :: has_eps, device, n_gpus
     real(c_double)     :: helmholtz_lambda, lsf_boundary_value
     integer(c_int32_t) :: coarse_grid_size(3), periodic(3)
     real(c_double)     :: dr_base(3), r_base(3)
  end type afmg_opts

  !> afmg_stencil_desc (include/afmg.h): one box whose operator is not the implicit constant Laplacian
  type, bind(c) :: afmg_stencil_desc
     integer(c_int32_t) :: box_id, op_stype, prolong_shape, prolong_stype, tag, cylindrical_gradient
     integer(c_int64_t) :: op_offset, f_offset, prolong_offset
  end type afmg_stencil_desc

  type, bind(c) :: afmg_tree
     integer(c_int32_t) :: highest_lvl, highest_id
     type(c_ptr) :: lvl_counts, lvl_ids, lvl, ix, parent, children, neighbors, neighbor_mat, r_min
  end type afmg_tree

  interface
     integer(c_int) function afmg_create(h, opts) bind(c, name="afmg_create")
       import; type(c_ptr), intent(out) :: h; type(afmg_opts), intent(in) :: opts
     end function
     integer(c_int) function afmg_destroy(h) bind(c, name="afmg_destroy")
       import; type(c_ptr), value :: h
     end function
     integer(c_int) function afmg_set_tree(h, t) bind(c, name="afmg_set_tree")
       import; type(c_ptr), value :: h; type(afmg_tree), intent(in) :: t
     end function
     integer(c_int) function afmg_set_bc(h, n, ids, nbs, types, vals) bind(c, name="afmg_set_bc")
       import; type(c_ptr), value :: h; integer(c_int32_t), value :: n
       integer(c_int32_t), intent(in) :: ids(*), nbs(*), types(*); real(c_double), intent(in) :: vals(*)
     end function
     integer(c_int) function afmg_set_stencils(h, n, desc, blob, blob_len) bind(c, name="afmg_set_stencils")
       import; type(c_ptr), value :: h; integer(c_int32_t), value :: n
       type(afmg_stencil_desc), intent(in) :: desc(*); real(c_double), intent(in) :: blob(*)
       integer(c_int64_t), value :: blob_len
     end function
     integer(c_int) function afmg_set_lsf_boundary_value(h, v) bind(c, name="afmg_set_lsf_boundary_value")
       import; type(c_ptr), value :: h; real(c_double), value :: v
     end function
     integer(c_int) function afmg_set_lsf_boundary_values(h, n, ids, values) bind(c, name="afmg_set_lsf_boundary_values")
       import; type(c_ptr), value :: h; integer(c_int32_t), value :: n
       integer(c_int32_t), intent(in) :: ids(*); real(c_double), intent(in) :: values(*)
     end function
     integer(c_int) function afmg_upload(h, var, n, ids, packed) bind(c, name="afmg_upload")
       import; type(c_ptr), value :: h; integer(c_int32_t), value :: var, n
       integer(c_int32_t), intent(in) :: ids(*); real(c_double), intent(in) :: packed(*)
     end function
     integer(c_int) function afmg_download(h, var, n, ids, packed) bind(c, name="afmg_download")
       import; type(c_ptr), value :: h; integer(c_int32_t), value :: var, n
       integer(c_int32_t), intent(in) :: ids(*); real(c_double), intent(out) :: packed(*)
     end function
     type(c_ptr) function afmg_host_alloc(bytes) bind(c, name="afmg_host_alloc")
       import; integer(c_size_t), value :: bytes
     end function
     subroutine afmg_host_free(p) bind(c, name="afmg_host_free")
       import; type(c_ptr), value :: p
     end subroutine
     integer(c_int) function afmg_fas_fmg(h, set_residual, have_guess) bind(c, name="afmg_fas_fmg")
       import; type(c_ptr), value :: h; integer(c_int32_t), value :: set_residual, have_guess
     end function
     integer(c_int) function afmg_fas_vcycle(h, set_residual, highest_lvl, standalone) &
          bind(c, name="afmg_fas_vcycle")
       import; type(c_ptr), value :: h; integer(c_int32_t), value :: set_residual, highest_lvl, standalone
     end function
     integer(c_int) function afmg_set_helmholtz_lambda(h, lambda) bind(c, name="afmg_set_helmholtz_lambda")
       import; type(c_ptr), value :: h; real(c_double), value :: lambda
     end function
     integer(c_int) function afmg_update_operator_stencil(h) bind(c, name="afmg_update_operator_stencil")
       import; type(c_ptr), value :: h
     end function
     integer(c_int) function afmg_compute_phi_gradient(h, fac, with_norm) bind(c, name="afmg_compute_phi_gradient")
       import; type(c_ptr), value :: h; real(c_double), value :: fac; integer(c_int32_t), value :: with_norm
     end function
     integer(c_int) function afmg_compute_field_norm(h) bind(c, name="afmg_compute_field_norm")
       import; type(c_ptr), value :: h
     end function
     integer(c_int) function afmg_gc_tree(h, var, corners) bind(c, name="afmg_gc_tree")
       import; type(c_ptr), value :: h; integer(c_int32_t), value :: var, corners
     end function
     integer(c_int) function afmg_set_lsf_distances(h, n_boxes, ids, n_entries, cell_ix, dd, lsf) &
          bind(c, name="afmg_set_lsf_distances")
       import; type(c_ptr), value :: h; integer(c_int32_t), value :: n_boxes
       integer(c_int32_t), intent(in) :: ids(*), n_entries(*), cell_ix(*); real(c_double), intent(in) :: dd(*), lsf(*)
     end function
     integer(c_int) function afmg_upload_fc(h, n, ids, packed) bind(c, name="afmg_upload_fc")
       import; type(c_ptr), value :: h; integer(c_int32_t), value :: n
       integer(c_int32_t), intent(in) :: ids(*); real(c_double), intent(in) :: packed(*)
     end function
     integer(c_int) function afmg_download_fc(h, n, ids, packed) bind(c, name="afmg_download_fc")
       import; type(c_ptr), value :: h; integer(c_int32_t), value :: n
       integer(c_int32_t), intent(in) :: ids(*); real(c_double), intent(out) :: packed(*)
     end function
     integer(c_int) function afmg_helmholtz_compute(modes, n_modes, coeffs, max_fmg, max_rel, n_cycles, residuals) &
          bind(c, name="afmg_helmholtz_compute")
       import; type(c_ptr), intent(in) :: modes(*); integer(c_int32_t), value :: n_modes, max_fmg
       real(c_double), intent(in) :: coeffs(*); real(c_double), value :: max_rel
       integer(c_int32_t), intent(out) :: n_cycles(*); real(c_double), intent(out) :: residuals(*)
     end function
     integer(c_int) function afmg_field_solve(h, have_guess, threshold, max_residual, max_fmg, n_vcycles, &
          residuals, n_fmg, n_vc) bind(c, name="afmg_field_solve")
       import; type(c_ptr), value :: h; integer(c_int32_t), value :: have_guess, max_fmg, n_vcycles
       real(c_double), value :: threshold, max_residual
       real(c_double), intent(out) :: residuals(*); integer(c_int32_t), intent(out) :: n_fmg, n_vc
     end function
     integer(c_int) function afmg_max_abs(h, var, val) bind(c, name="afmg_max_abs")
       import; type(c_ptr), value :: h; integer(c_int32_t), value :: var; real(c_double), intent(out) :: val
     end function
  end interface

  integer, parameter :: afmg_phi = 0, afmg_rhs = 1, afmg_tmp = 2, afmg_eps = 3, afmg_fld = 4, afmg_photo = 5

  !> One GPU solver per mg_t; looked up by the address-independent key mg%i_phi/mg%i_rhs/lambda slot
  type gpu_state_t
     type(c_ptr)        :: h = c_null_ptr
     integer(c_int64_t) :: tree_hash = -1          ! hash of the topology the device holds
     logical            :: phi_current = .false.   ! the device copy of phi equals the host copy
     logical            :: download_tmp = .true.   ! bring i_tmp back after a solve with set_residual
     type(c_ptr)        :: pinned = c_null_ptr     ! page-locked packing buffer (afmg_host_alloc)
     integer(c_size_t)  :: pinned_len = 0          ! its size in doubles
  end type gpu_state_t

  integer, parameter :: max_solvers = 16
  type(gpu_state_t), save :: solvers(max_solvers)

  public :: mg_gpu_init, mg_gpu_destroy, mg_gpu_fas_fmg, mg_gpu_fas_vcycle
  public :: mg_gpu_update_operator_stencil, mg_gpu_tree_maxabs_tmp
  public :: mg_gpu_compute_phi_gradient, mg_gpu_compute_field_norm, mg_gpu_gc_tree_norm, photoi_gpu_helmh_compute
  public :: mg_gpu_field_solve
  public :: mg_gpu_invalidate, mg_gpu_mark_phi_dirty, mg_gpu_set_download_tmp

contains

  subroutine check(rc, what)
    integer(c_int), intent(in) :: rc
    character(len=*), intent(in) :: what
    if (rc /= 0) then
       print *, "libafmg error ", rc, " in ", what
       error stop "m_af_multigrid_gpu"
    end if
  end subroutine check

  !> mg_init (m_af_multigrid.f90:43-109): options -> afmg_create, topology -> afmg_set_tree
  subroutine mg_gpu_init(tree, mg, slot)
    type(af_t), intent(inout) :: tree
    type(mg_t), intent(inout) :: mg
    integer, intent(in)       :: slot !< which of the (up to 16) solvers: 1 = field, 2.. = Helmholtz modes
    type(afmg_opts) :: o

    if (.not. associated(mg%sides_bc)) error stop "mg_init: sides_bc not set"
    o%ndim = NDIM; o%n_cell = tree%n_cell; o%coord_t = tree%coord_t
    o%n_cycle_down = mg%n_cycle_down; o%n_cycle_up = mg%n_cycle_up
    o%use_corners = merge(1, 0, mg%use_corners); o%subtract_mean = merge(1, 0, mg%subtract_mean)
    o%prolongation_type = mg%prolongation_type; o%operator_mask = mg%operator_mask
    o%has_eps = merge(1, 0, tree%mg_i_eps > 0); o%device = -1; o%n_gpus = 0   ! 0: one GPU, or AFMG_N_GPUS GPUs of one node
    o%helmholtz_lambda = mg%helmholtz_lambda; o%lsf_boundary_value = mg%lsf_boundary_value
    o%coarse_grid_size = 1; o%periodic = 0; o%dr_base = 0; o%r_base = 0
    o%coarse_grid_size(1:NDIM) = tree%coarse_grid_size(1:NDIM)
    o%periodic(1:NDIM) = merge(1, 0, tree%periodic(1:NDIM))
    o%dr_base(1:NDIM) = tree%dr_base; o%r_base(1:NDIM) = tree%r_base
    call check(afmg_create(solvers(slot)%h, o), "afmg_create")
    mg%initialized = .true.
    solvers(slot)%tree_hash = -1
    call prepare_solve(tree, mg, slot)
  end subroutine mg_gpu_init

  !> Force the next solve to re-send the topology (e.g. from a wrapper of af_adjust_refinement when
  !> ref_info%n_add + ref_info%n_rm > 0); the hash below makes this optional, not required
  subroutine mg_gpu_invalidate(slot)
    integer, intent(in) :: slot
    solvers(slot)%tree_hash = -1
    solvers(slot)%phi_current = .false.
  end subroutine mg_gpu_invalidate

  !> The host changed mg%i_phi since the last solve (anything but af_adjust_refinement, which is detected)
  subroutine mg_gpu_mark_phi_dirty(slot)
    integer, intent(in) :: slot
    solvers(slot)%phi_current = .false.
  end subroutine mg_gpu_mark_phi_dirty

  subroutine mg_gpu_set_download_tmp(slot, flag)
    integer, intent(in) :: slot
    logical, intent(in) :: flag
    solvers(slot)%download_tmp = flag
  end subroutine mg_gpu_set_download_tmp

  !> 64-bit mix without overflow (shifts and xor only): one step of the topology hash
  pure subroutine hash_mix(h, v)
    integer(c_int64_t), intent(inout) :: h
    integer, intent(in)               :: v
    h = ieor(h, int(v, c_int64_t) + 40503_c_int64_t)
    h = ieor(h, ishft(h, 13))
    h = ieor(h, ishft(h, -7))
    h = ieor(h, ishft(h, 17))
  end subroutine hash_mix

  !> Hash of everything afmg_set_tree receives: af_adjust_refinement reuses the ids of removed boxes, so counts
  !> and highest_id do not identify a tree; the complete id lists with parents, children and neighbours do
  function topology_hash(tree) result(h)
    type(af_t), intent(in) :: tree
    integer(c_int64_t)     :: h
    integer                :: lvl, i, id, n
    h = 1469598103934665603_c_int64_t
    call hash_mix(h, tree%highest_lvl)
    call hash_mix(h, tree%highest_id)
    do lvl = 1, tree%highest_lvl
       call hash_mix(h, size(tree%lvls(lvl)%ids))
       do i = 1, size(tree%lvls(lvl)%ids)
          id = tree%lvls(lvl)%ids(i)
          call hash_mix(h, id)
          call hash_mix(h, tree%boxes(id)%parent)
          call hash_mix(h, tree%boxes(id)%children(1))
          do n = 1, af_num_neighbors
             call hash_mix(h, tree%boxes(id)%neighbors(n))
          end do
          do n = 1, NDIM
             call hash_mix(h, tree%boxes(id)%ix(n))
          end do
       end do
    end do
    if (h == -1) h = 0   ! -1 means "nothing on the device"
  end function topology_hash

  !> Steps 1-3 of the header: everything a solve needs on the device except cell data
  subroutine prepare_solve(tree, mg, slot)
    type(af_t), intent(inout) :: tree
    type(mg_t), intent(in)    :: mg
    integer, intent(in)       :: slot
    call mg_use(tree, mg)          ! tags + stencils of new boxes, tree%mg_current_operator_mask
    call sync_tree(tree, mg, slot) ! topology + stencils, when the tree changed
    call sync_bc(tree, mg, slot)   ! boundary values carry the voltage: every solve
    call check(afmg_set_lsf_boundary_value(solvers(slot)%h, mg%lsf_boundary_value), "afmg_set_lsf_boundary_value")
    call sync_lsf_boundary_values(tree, mg, slot)
  end subroutine prepare_solve

  !> done_with_mg of the reference (m_af_multigrid.f90:128-132; private there)
  subroutine finish_solve(tree)
    type(af_t), intent(inout) :: tree
    tree%mg_current_operator_mask = -1
  end subroutine finish_solve

  !> Forward the topology (and with it the explicit stencils) when the tree changed (after af_adjust_refinement)
  subroutine sync_tree(tree, mg, slot)
    type(af_t), intent(inout) :: tree
    type(mg_t), intent(in)    :: mg
    integer, intent(in)       :: slot
    integer :: lvl, n, i, id
    integer(c_int64_t) :: hash
    integer(c_int32_t), allocatable, target :: counts(:), ids(:), blvl(:), ix(:, :), parent(:), &
         children(:, :), neighbors(:, :), nmat(:, :)
    real(c_double), allocatable, target :: r_min(:, :)
    type(afmg_tree) :: t

    hash = topology_hash(tree)
    if (hash == solvers(slot)%tree_hash) return
    solvers(slot)%tree_hash = hash
    solvers(slot)%phi_current = .false.   ! new boxes hold prolonged values the device has not seen

    n = tree%highest_id
    allocate(counts(tree%highest_lvl), blvl(0:n), ix(NDIM, 0:n), parent(0:n))
    allocate(children(2**NDIM, 0:n), neighbors(2*NDIM, 0:n), nmat(3**NDIM, 0:n), r_min(NDIM, 0:n))
    blvl = 0; ix = 0; parent = 0; children = 0; neighbors = 0; nmat = 0; r_min = 0
    do lvl = 1, tree%highest_lvl
       counts(lvl) = size(tree%lvls(lvl)%ids)
    end do
    allocate(ids(sum(counts)))
    i = 0
    do lvl = 1, tree%highest_lvl
       ids(i+1:i+counts(lvl)) = tree%lvls(lvl)%ids
       i = i + counts(lvl)
    end do
    do i = 1, size(ids)
       id = ids(i)
       blvl(id) = tree%boxes(id)%lvl; ix(:, id) = tree%boxes(id)%ix
       parent(id) = tree%boxes(id)%parent; children(:, id) = tree%boxes(id)%children
       neighbors(:, id) = tree%boxes(id)%neighbors
       nmat(:, id) = reshape(tree%boxes(id)%neighbor_mat, [3**NDIM])
       r_min(:, id) = tree%boxes(id)%r_min
    end do
    t%highest_lvl = tree%highest_lvl; t%highest_id = n
    t%lvl_counts = c_loc(counts); t%lvl_ids = c_loc(ids); t%lvl = c_loc(blvl); t%ix = c_loc(ix)
    t%parent = c_loc(parent); t%children = c_loc(children); t%neighbors = c_loc(neighbors)
    t%neighbor_mat = c_loc(nmat); t%r_min = c_loc(r_min)
    call check(afmg_set_tree(solvers(slot)%h, t), "afmg_set_tree")
    call sync_stencils(tree, mg, slot)
  end subroutine sync_tree

  !> mg%sides_bc evaluated on the host for every physical face (it never depends on phi) -> afmg_set_bc.  Called
  !> before every solve: field_bc_homogeneous returns the current voltage (src/m_field.f90:590-610).  The library keeps
  !> its cycle graphs when only the values change; changed types rebuild the coarse solver.
  subroutine sync_bc(tree, mg, slot)
    type(af_t), intent(inout) :: tree
    type(mg_t), intent(in)    :: mg
    integer, intent(in)       :: slot
    integer :: lvl, i, id, nb, n_faces, nc2, bc_type
    integer(c_int32_t), allocatable :: f_ids(:), f_nbs(:), f_types(:)
    real(c_double), allocatable :: f_vals(:, :)
    real(dp), allocatable :: coords(:, :)

    nc2 = tree%n_cell**(NDIM-1)
    n_faces = 0
    do lvl = 1, tree%highest_lvl
       do i = 1, size(tree%lvls(lvl)%ids)
          n_faces = n_faces + count(tree%boxes(tree%lvls(lvl)%ids(i))%neighbors < af_no_box)
       end do
    end do
    allocate(f_ids(n_faces), f_nbs(n_faces), f_types(n_faces), f_vals(nc2, n_faces), coords(NDIM, nc2))
    n_faces = 0
    do lvl = 1, tree%highest_lvl
     do i = 1, size(tree%lvls(lvl)%ids)
       id = tree%lvls(lvl)%ids(i)
       do nb = 1, af_num_neighbors
          if (tree%boxes(id)%neighbors(nb) < af_no_box) then
             n_faces = n_faces + 1
             call af_get_face_coords(tree%boxes(id), nb, coords)
             call mg%sides_bc(tree%boxes(id), nb, mg%i_phi, coords, f_vals(:, n_faces), bc_type)
             f_ids(n_faces) = id; f_nbs(n_faces) = nb; f_types(n_faces) = bc_type
          end if
       end do
     end do
    end do
    call check(afmg_set_bc(solvers(slot)%h, n_faces, f_ids, f_nbs, f_types, f_vals), "afmg_set_bc")
  end subroutine sync_bc

  !> Ship the stencils mg_set_operators_lvl (m_af_multigrid.f90:1147-1185) stored in box%stencils for
  !> every box that is not a plain constant-Laplacian box: variable / constant eps operators
  !> (mg_box_lpld_stencil), level-set boxes (mg_box_lsf_stencil: v and f) and their prolongation
  !> stencils.  The builders themselves stay in afivo (they call mg%lsf and friends).  Also called from
  !> mg_gpu_update_operator_stencil after mg_update_operator_stencil changed eps / lsf stencils.
  subroutine sync_stencils(tree, mg, slot)
    type(af_t), intent(inout) :: tree
    type(mg_t), intent(in)    :: mg
    integer, intent(in)       :: slot
    type(afmg_stencil_desc), allocatable :: desc(:)
    real(c_double), allocatable :: blob(:)
    integer :: lvl, i, id, n, ix, ixp, pass
    integer(c_int64_t) :: off

    ! two passes: count (sizes), then fill
    do pass = 1, 2
       n = 0; off = 0
       do lvl = 1, tree%highest_lvl
          do i = 1, size(tree%lvls(lvl)%ids)
             id = tree%lvls(lvl)%ids(i)
             associate (box => tree%boxes(id))
               if (iand(box%tag, mg%operator_mask) == mg_normal_box) cycle
               ix = af_stencil_index(box, mg%operator_key)
               if (ix == af_stencil_none) cycle
               n = n + 1
               if (pass == 2) then
                  desc(n)%box_id = id; desc(n)%tag = box%tag
                  desc(n)%op_stype = box%stencils(ix)%stype
                  desc(n)%cylindrical_gradient = merge(1, 0, box%stencils(ix)%cylindrical_gradient)
                  desc(n)%op_offset = off; desc(n)%f_offset = -1; desc(n)%prolong_shape = 0
                  desc(n)%prolong_stype = 0; desc(n)%prolong_offset = 0
               end if
               if (box%stencils(ix)%stype == stencil_constant) then
                  if (pass == 2) blob(off+1:off+size(box%stencils(ix)%c)) = box%stencils(ix)%c
                  off = off + size(box%stencils(ix)%c)
               else
                  if (pass == 2) blob(off+1:off+size(box%stencils(ix)%v)) = &
                       reshape(box%stencils(ix)%v, [size(box%stencils(ix)%v)])
                  off = off + size(box%stencils(ix)%v)
               end if
               if (allocated(box%stencils(ix)%f)) then
                  if (pass == 2) then
                     desc(n)%f_offset = off
                     blob(off+1:off+size(box%stencils(ix)%f)) = reshape(box%stencils(ix)%f, [size(box%stencils(ix)%f)])
                  end if
                  off = off + size(box%stencils(ix)%f)
               end if
               ixp = af_stencil_index(box, mg%prolongation_key)
               if (lvl > 1 .and. ixp /= af_stencil_none) then
                  if (pass == 2) then
                     desc(n)%prolong_shape = box%stencils(ixp)%shape    ! af_stencil_p234 = 2, af_stencil_p248 = 3
                     desc(n)%prolong_stype = box%stencils(ixp)%stype
                     desc(n)%prolong_offset = off
                  end if
                  if (box%stencils(ixp)%stype == stencil_constant) then
                     if (pass == 2) blob(off+1:off+size(box%stencils(ixp)%c)) = box%stencils(ixp)%c
                     off = off + size(box%stencils(ixp)%c)
                  else
                     if (pass == 2) blob(off+1:off+size(box%stencils(ixp)%v)) = &
                          reshape(box%stencils(ixp)%v, [size(box%stencils(ixp)%v)])
                     off = off + size(box%stencils(ixp)%v)
                  end if
               end if
             end associate
          end do
       end do
       if (pass == 1) allocate(desc(max(n, 1)), blob(max(off, 1_c_int64_t)))
    end do
    call check(afmg_set_lsf_boundary_value(solvers(slot)%h, mg%lsf_boundary_value), "afmg_set_lsf_boundary_value")
    call check(afmg_set_stencils(solvers(slot)%h, n, desc, blob, off), "afmg_set_stencils")
  end subroutine sync_stencils

  !> mg%lsf_boundary_function (several electrodes at their own potentials, src/m_field.f90:294-374): evaluate
  !> mg_lsf_boundary_value (m_coarse_solver.f90:493-510) for every box that carries a level-set boundary and ship
  !> the values; call before each solve, like field_compute sets mg%lsf_boundary_value (src/m_field.f90:481-487)
  subroutine sync_lsf_boundary_values(tree, mg, slot)
    use m_coarse_solver, only: mg_lsf_boundary_value
    type(af_t), intent(in) :: tree
    type(mg_t), intent(in) :: mg
    integer, intent(in)    :: slot
    integer, allocatable   :: ids(:)
    real(c_double), allocatable :: vals(:, :)
    integer :: lvl, i, id, n, nc
    if (.not. associated(mg%lsf_boundary_function)) return
    nc = tree%n_cell
    n = 0
    do lvl = 1, tree%highest_lvl
       do i = 1, size(tree%lvls(lvl)%ids)
          if (iand(tree%boxes(tree%lvls(lvl)%ids(i))%tag, mg_lsf_box) > 0) n = n + 1
       end do
    end do
    allocate(ids(n), vals(nc**NDIM, n))
    n = 0
    do lvl = 1, tree%highest_lvl
       do i = 1, size(tree%lvls(lvl)%ids)
          id = tree%lvls(lvl)%ids(i)
          if (iand(tree%boxes(id)%tag, mg_lsf_box) > 0) then
             n = n + 1
             ids(n) = id
             vals(:, n) = reshape(mg_lsf_boundary_value(tree%boxes(id), mg), [nc**NDIM])
          end if
       end do
    end do
    call check(afmg_set_lsf_boundary_values(solvers(slot)%h, n, ids, vals), "afmg_set_lsf_boundary_values")
  end subroutine sync_lsf_boundary_values

  !> Page-locked packing buffer of at least n doubles, kept between calls (afmg_host_alloc: transfers from it are
  !> DMA copies at the PCIe rate; a pageable temporary is staged by the driver at a fraction of it)
  subroutine pinned_buffer(slot, n, buf_ptr)
    integer, intent(in)            :: slot
    integer(c_size_t), intent(in)  :: n
    type(c_ptr), intent(out)       :: buf_ptr
    if (n > solvers(slot)%pinned_len) then
       if (c_associated(solvers(slot)%pinned)) call afmg_host_free(solvers(slot)%pinned)
       solvers(slot)%pinned = afmg_host_alloc(8_c_size_t * n)
       if (.not. c_associated(solvers(slot)%pinned)) error stop "m_af_multigrid_gpu: afmg_host_alloc failed"
       solvers(slot)%pinned_len = n
    end if
    buf_ptr = solvers(slot)%pinned
  end subroutine pinned_buffer

  !> Pack box%cc(:, :, :, iv) of a list of boxes and upload / download
  subroutine transfer(tree, slot, iv, var, ids, up)
    type(af_t), intent(inout) :: tree
    integer, intent(in)       :: slot, iv, var, ids(:)
    logical, intent(in)       :: up
    real(c_double), pointer   :: buf(:, :)
    type(c_ptr)               :: buf_ptr
    integer :: i, n2
    n2 = (tree%n_cell + 2)**NDIM
    call pinned_buffer(slot, int(n2, c_size_t) * int(size(ids), c_size_t), buf_ptr)
    call c_f_pointer(buf_ptr, buf, [n2, size(ids)])
    if (up) then
       !$omp parallel do
       do i = 1, size(ids)
          buf(:, i) = reshape(tree%boxes(ids(i))%cc(DTIMES(:), iv), [n2])
       end do
       call check(afmg_upload(solvers(slot)%h, var, size(ids), ids, buf), "afmg_upload")
    else
       call check(afmg_download(solvers(slot)%h, var, size(ids), ids, buf), "afmg_download")
       !$omp parallel do
       do i = 1, size(ids)
          tree%boxes(ids(i))%cc(DTIMES(:), iv) = reshape(buf(:, i), shape(tree%boxes(ids(i))%cc(DTIMES(:), iv)))
       end do
    end if
  end subroutine transfer

  subroutine all_ids(tree, ids, leaves_only)
    type(af_t), intent(in) :: tree
    integer, allocatable, intent(out) :: ids(:)
    logical, intent(in) :: leaves_only
    integer :: lvl, n
    n = 0
    do lvl = 1, tree%highest_lvl
       n = n + merge(size(tree%lvls(lvl)%leaves), size(tree%lvls(lvl)%ids), leaves_only)
    end do
    allocate(ids(n))
    n = 0
    do lvl = 1, tree%highest_lvl
       if (leaves_only) then
          ids(n+1:n+size(tree%lvls(lvl)%leaves)) = tree%lvls(lvl)%leaves
          n = n + size(tree%lvls(lvl)%leaves)
       else
          ids(n+1:n+size(tree%lvls(lvl)%ids)) = tree%lvls(lvl)%ids
          n = n + size(tree%lvls(lvl)%ids)
       end if
    end do
  end subroutine all_ids

  !> mg_fas_fmg(tree, mg, set_residual, have_guess)  (m_af_multigrid.f90:137-180)
  subroutine mg_gpu_fas_fmg(tree, mg, set_residual, have_guess, slot)
    type(af_t), intent(inout) :: tree
    type(mg_t), intent(inout) :: mg
    logical, intent(in)       :: set_residual, have_guess
    integer, intent(in)       :: slot
    integer, allocatable      :: ids(:), leaves(:)
    call prepare_solve(tree, mg, slot)
    call all_ids(tree, leaves, .true.); call all_ids(tree, ids, .false.)
    call transfer(tree, slot, mg%i_rhs, afmg_rhs, leaves, .true.)      ! callers set rhs on leaves only
    if (have_guess .and. .not. solvers(slot)%phi_current) call transfer(tree, slot, mg%i_phi, afmg_phi, ids, .true.)
    call check(afmg_fas_fmg(solvers(slot)%h, merge(1, 0, set_residual), merge(1, 0, have_guess)), "afmg_fas_fmg")
    call transfer(tree, slot, mg%i_phi, afmg_phi, ids, .false.)
    solvers(slot)%phi_current = .true.
    if (set_residual .and. solvers(slot)%download_tmp) call transfer(tree, slot, mg%i_tmp, afmg_tmp, ids, .false.)
    call finish_solve(tree)
  end subroutine mg_gpu_fas_fmg

  !> mg_fas_vcycle(tree, mg, set_residual, highest_lvl, standalone)  (m_af_multigrid.f90:185-264)
  subroutine mg_gpu_fas_vcycle(tree, mg, set_residual, slot, highest_lvl, standalone)
    type(af_t), intent(inout)     :: tree
    type(mg_t), intent(in)        :: mg
    logical, intent(in)           :: set_residual
    integer, intent(in)           :: slot
    integer, intent(in), optional :: highest_lvl
    logical, intent(in), optional :: standalone
    integer, allocatable          :: ids(:), leaves(:)
    integer :: max_lvl, alone
    max_lvl = 0; if (present(highest_lvl)) max_lvl = highest_lvl
    alone = 1; if (present(standalone)) alone = merge(1, 0, standalone)
    call prepare_solve(tree, mg, slot)
    call all_ids(tree, leaves, .true.); call all_ids(tree, ids, .false.)
    call transfer(tree, slot, mg%i_rhs, afmg_rhs, leaves, .true.)
    if (.not. solvers(slot)%phi_current) call transfer(tree, slot, mg%i_phi, afmg_phi, ids, .true.)
    call check(afmg_fas_vcycle(solvers(slot)%h, merge(1, 0, set_residual), max_lvl, alone), "afmg_fas_vcycle")
    call transfer(tree, slot, mg%i_phi, afmg_phi, ids, .false.)
    solvers(slot)%phi_current = .true.
    if (set_residual .and. solvers(slot)%download_tmp) call transfer(tree, slot, mg%i_tmp, afmg_tmp, ids, .false.)
    call finish_solve(tree)
  end subroutine mg_gpu_fas_vcycle

  !> mg_compute_phi_gradient(tree, mg, i_fc, fac, i_norm) (m_af_multigrid.f90:1857-1898) on the device, from the
  !> potential the last solve left there.  The face-centred field (and the norm) are downloaded into box%fc /
  !> box%cc.  Variable-eps boxes need tree%mg_i_eps on the device, level-set boxes their distance stencils;
  !> both are (re-)sent when the tree changed.
  subroutine mg_gpu_compute_phi_gradient(tree, mg, i_fc, fac, slot, i_norm)
    type(af_t), intent(inout)     :: tree
    type(mg_t), intent(in)        :: mg
    integer, intent(in)           :: i_fc, slot
    real(dp), intent(in)          :: fac
    integer, intent(in), optional :: i_norm
    integer, allocatable          :: ids(:)
    real(c_double), allocatable   :: buf(:, :)
    integer :: i, n1
    call mg_use(tree, mg)            ! box tags / distance stencils of new boxes (m_af_multigrid.f90:1865)
    call sync_tree(tree, mg, slot)
    call check(afmg_set_lsf_boundary_value(solvers(slot)%h, mg%lsf_boundary_value), "afmg_set_lsf_boundary_value")
    call all_ids(tree, ids, .false.)
    if (tree%mg_i_eps > 0) call transfer(tree, slot, tree%mg_i_eps, afmg_eps, ids, .true.)
    if (tree%mg_i_lsf > 0) call sync_lsf_distances(tree, mg, slot, ids)
    call check(afmg_compute_phi_gradient(solvers(slot)%h, fac, merge(1, 0, present(i_norm))), "afmg_compute_phi_gradient")
    n1 = NDIM * (tree%n_cell + 1)**NDIM
    allocate(buf(n1, size(ids)))
    call check(afmg_download_fc(solvers(slot)%h, size(ids), ids, buf), "afmg_download_fc")
    !$omp parallel do
    do i = 1, size(ids)
       tree%boxes(ids(i))%fc(DTIMES(:), :, i_fc) = reshape(buf(:, i), shape(tree%boxes(ids(i))%fc(DTIMES(:), :, i_fc)))
    end do
    if (present(i_norm)) call transfer(tree, slot, i_norm, afmg_fld, ids, .false.)
    call finish_solve(tree)
  end subroutine mg_gpu_compute_phi_gradient

  !> mg_compute_field_norm (m_af_multigrid.f90:2002-2020) after the host changed box%fc (surface_correct_field_fc)
  subroutine mg_gpu_compute_field_norm(tree, i_fc, i_norm, slot)
    type(af_t), intent(inout)   :: tree
    integer, intent(in)         :: i_fc, i_norm, slot
    integer, allocatable        :: ids(:)
    real(c_double), allocatable :: buf(:, :)
    integer :: i, n1
    call all_ids(tree, ids, .false.)
    n1 = NDIM * (tree%n_cell + 1)**NDIM
    allocate(buf(n1, size(ids)))
    !$omp parallel do
    do i = 1, size(ids)
       buf(:, i) = reshape(tree%boxes(ids(i))%fc(DTIMES(:), :, i_fc), [n1])
    end do
    call check(afmg_upload_fc(solvers(slot)%h, size(ids), ids, buf), "afmg_upload_fc")
    call check(afmg_compute_field_norm(solvers(slot)%h), "afmg_compute_field_norm")
    call transfer(tree, slot, i_norm, afmg_fld, ids, .false.)
  end subroutine mg_gpu_compute_field_norm

  !> af_gc_tree(tree, [i_norm]) (m_af_ghostcell.f90:25-46) for the field norm with af_bc_neumann_zero / af_gc_interp
  !> (src/m_field.f90:392-393, :547); downloads the norm with its ghost cells
  subroutine mg_gpu_gc_tree_norm(tree, i_norm, slot)
    type(af_t), intent(inout) :: tree
    integer, intent(in)       :: i_norm, slot
    integer, allocatable      :: ids(:)
    call all_ids(tree, ids, .false.)
    call check(afmg_gc_tree(solvers(slot)%h, afmg_fld, 1), "afmg_gc_tree")
    call transfer(tree, slot, i_norm, afmg_fld, ids, .false.)
  end subroutine mg_gpu_gc_tree_norm

  !> Ship the sparse level-set distance stencils (mg_lsf_distance_key) and the lsf value of their cells
  subroutine sync_lsf_distances(tree, mg, slot, ids)
    type(af_t), intent(in) :: tree
    type(mg_t), intent(in) :: mg
    integer, intent(in)    :: slot, ids(:)
    integer, allocatable   :: n_entries(:), cells(:, :)
    real(c_double), allocatable :: dd(:, :), lsf(:)
    integer :: i, n, ix, ne, k, m
#if NDIM == 2
    integer :: ij(2)
#elif NDIM == 3
    integer :: ijk3(3)
#endif
    allocate(n_entries(size(ids)))
    ne = 0
    do i = 1, size(ids)
       ix = af_stencil_index(tree%boxes(ids(i)), mg_lsf_distance_key)
       n_entries(i) = 0
       if (ix /= af_stencil_none) n_entries(i) = size(tree%boxes(ids(i))%stencils(ix)%sparse_ix, 2)
       ne = ne + n_entries(i)
    end do
    allocate(cells(NDIM, max(ne, 1)), dd(2*NDIM, max(ne, 1)), lsf(max(ne, 1)))
    k = 0
    do i = 1, size(ids)
       if (n_entries(i) == 0) cycle
       ix = af_stencil_index(tree%boxes(ids(i)), mg_lsf_distance_key)
       associate (st => tree%boxes(ids(i))%stencils(ix), box => tree%boxes(ids(i)))
         do n = 1, n_entries(i)
            k = k + 1
            cells(:, k) = st%sparse_ix(:, n)
            dd(:, k) = st%sparse_v(:, n)
#if NDIM == 2
            ij = st%sparse_ix(:, n); lsf(k) = box%cc(ij(1), ij(2), mg%i_lsf)
#elif NDIM == 3
            ijk3 = st%sparse_ix(:, n); lsf(k) = box%cc(ijk3(1), ijk3(2), ijk3(3), mg%i_lsf)
#endif
         end do
       end associate
    end do
    m = size(ids)
    call check(afmg_set_lsf_distances(solvers(slot)%h, m, ids, n_entries, cells, dd, lsf), "afmg_set_lsf_distances")
  end subroutine sync_lsf_distances

  !> photoi_helmh_compute (src/m_photoi_helmh.f90:162-204) with all modes solved on the device: slots(n) is the
  !> solver of mg_helm(n); the shared rhs goes up once, the accumulated source i_photo comes back once.
  subroutine photoi_gpu_helmh_compute(tree, mg_helm, slots, coeffs, i_photo, max_fmg_cycles, max_rel_residual)
    type(af_t), intent(inout) :: tree
    type(mg_t), intent(inout) :: mg_helm(:)
    integer, intent(in)       :: slots(:), i_photo, max_fmg_cycles
    real(dp), intent(in)      :: coeffs(:), max_rel_residual
    type(c_ptr)               :: hs(size(slots))
    integer(c_int32_t)        :: n_cycles(size(slots))
    real(c_double)            :: residuals(size(slots))
    integer, allocatable      :: ids(:), leaves(:)
    integer :: n
    do n = 1, size(slots)
       call prepare_solve(tree, mg_helm(n), slots(n))
       hs(n) = solvers(slots(n))%h
    end do
    call all_ids(tree, leaves, .true.); call all_ids(tree, ids, .false.)
    call transfer(tree, slots(1), mg_helm(1)%i_rhs, afmg_rhs, leaves, .true.)
    call check(afmg_helmholtz_compute(hs, size(slots), coeffs, max_fmg_cycles, max_rel_residual, n_cycles, residuals), &
         "afmg_helmholtz_compute")
    call transfer(tree, slots(1), i_photo, afmg_photo, leaves, .false.)
    call finish_solve(tree)
  end subroutine photoi_gpu_helmh_compute

  !> The solve loop of field_compute (src/m_field.f90:491-524) in one call: FMG cycles until the residual criterion
  !> is met (only without a guess), then up to n_vcycles V-cycles; only the residual max-norms cross PCIe per
  !> cycle.  rhs goes up before, phi (and the residual in i_tmp) come back after.
  subroutine mg_gpu_field_solve(tree, mg, have_guess, residual_threshold, max_residual, max_fmg, n_vcycles, slot, &
       residuals, n_fmg, n_vc)
    type(af_t), intent(inout) :: tree
    type(mg_t), intent(inout) :: mg
    logical, intent(in)       :: have_guess
    real(dp), intent(in)      :: residual_threshold, max_residual
    integer, intent(in)       :: max_fmg, n_vcycles, slot
    real(dp), intent(out)     :: residuals(max_fmg + n_vcycles)
    integer, intent(out)      :: n_fmg, n_vc
    integer, allocatable      :: ids(:), leaves(:)
    integer(c_int)            :: rc
    call prepare_solve(tree, mg, slot)
    call all_ids(tree, leaves, .true.); call all_ids(tree, ids, .false.)
    call transfer(tree, slot, mg%i_rhs, afmg_rhs, leaves, .true.)
    if (have_guess .and. .not. solvers(slot)%phi_current) call transfer(tree, slot, mg%i_phi, afmg_phi, ids, .true.)
    rc = afmg_field_solve(solvers(slot)%h, merge(1, 0, have_guess), residual_threshold, max_residual, max_fmg, &
         n_vcycles, residuals, n_fmg, n_vc)
    if (rc == -7) error stop "No convergence in initial field computation"  ! AFMG_ERR_NOT_CONVERGED
    call check(rc, "afmg_field_solve")
    call transfer(tree, slot, mg%i_phi, afmg_phi, ids, .false.)
    solvers(slot)%phi_current = .true.
    if (solvers(slot)%download_tmp) call transfer(tree, slot, mg%i_tmp, afmg_tmp, ids, .false.)
    call finish_solve(tree)
  end subroutine mg_gpu_field_solve

  !> max |residual| over leaves without downloading i_tmp (af_tree_maxabs_cc, m_af_utils.f90:773)
  subroutine mg_gpu_tree_maxabs_tmp(slot, val)
    integer, intent(in)   :: slot
    real(dp), intent(out) :: val
    call check(afmg_max_abs(solvers(slot)%h, afmg_tmp, val), "afmg_max_abs")
  end subroutine mg_gpu_tree_maxabs_tmp

  !> mg_update_operator_stencil (m_af_multigrid.f90:1188-1214): call after the reference routine rebuilt
  !> the host stencils (new_lsf / new_eps); lambda and the explicit stencils are forwarded
  subroutine mg_gpu_update_operator_stencil(tree, mg, slot)
    type(af_t), intent(inout) :: tree
    type(mg_t), intent(in)    :: mg
    integer, intent(in)       :: slot
    call check(afmg_set_helmholtz_lambda(solvers(slot)%h, mg%helmholtz_lambda), "afmg_set_helmholtz_lambda")
    call check(afmg_update_operator_stencil(solvers(slot)%h), "afmg_update_operator_stencil")
    if (solvers(slot)%tree_hash /= topology_hash(tree)) then
       call sync_tree(tree, mg, slot)      ! ships the stencils as well
    else
       call sync_stencils(tree, mg, slot)
    end if
  end subroutine mg_gpu_update_operator_stencil

  !> mg_destroy (m_af_multigrid.f90:111-115)
  subroutine mg_gpu_destroy(mg, slot)
    type(mg_t), intent(inout) :: mg
    integer, intent(in)       :: slot
    call check(afmg_destroy(solvers(slot)%h), "afmg_destroy")
    solvers(slot)%h = c_null_ptr; solvers(slot)%tree_hash = -1; solvers(slot)%phi_current = .false.
    if (c_associated(solvers(slot)%pinned)) call afmg_host_free(solvers(slot)%pinned)
    solvers(slot)%pinned = c_null_ptr; solvers(slot)%pinned_len = 0
    mg%initialized = .false.
  end subroutine mg_gpu_destroy

end module m_af_multigrid_gpu

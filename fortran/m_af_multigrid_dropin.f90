!> The solver entry points of afivo's m_af_multigrid under their own names and with their own argument lists
!> (afivo/src/m_af_multigrid.f90:43 mg_init, :111 mg_destroy, :137 mg_fas_fmg, :185 mg_fas_vcycle,
!> :1188 mg_update_operator_stencil, :1857 mg_compute_phi_gradient, :1997 mg_compute_field_norm), forwarding to
!> libafmg.so through m_af_multigrid_gpu.  A caller switches by importing these names from here instead of from
!> m_af_multigrid, e.g. in src/m_field.f90 and src/m_photoi_helmh.f90
!>
!>   use m_af_all, ref_mg_init => mg_init, ref_mg_destroy => mg_destroy, ref_mg_fas_fmg => mg_fas_fmg, &
!>        ref_mg_fas_vcycle => mg_fas_vcycle, ref_mg_update_operator_stencil => mg_update_operator_stencil, &
!>        ref_mg_compute_phi_gradient => mg_compute_phi_gradient, ref_mg_compute_field_norm => mg_compute_field_norm
!>   use m_af_multigrid_dropin
!>
!> and nothing else changes: no slot argument.  The solver instance behind an mg_t is found from mg%i_phi, which is
!> distinct for every solver of a run (the field solver uses i_phi, the Helmholtz modes i_modes(n),
!> src/m_photoi_helmh.f90:110-124, 176-189).
!>
!> The reference's own mg_init / mg_update_operator_stencil still run first: they build the host-side stencils
!> (mg_set_operators_lvl) that other afivo routines read and that the shim ships for eps / electrode boxes.
!>
!> NOTE: like m_af_multigrid_gpu.f90 this file has not been compiled here (no Fortran compiler in this image).
module m_af_multigrid_dropin
  use m_af_types
  use m_af_multigrid, only: ref_mg_init => mg_init, ref_mg_destroy => mg_destroy, &
       ref_mg_update_operator_stencil => mg_update_operator_stencil
  use m_af_multigrid_gpu
  implicit none
  private

  integer, parameter :: max_solvers = 16
  integer, save      :: slot_key(max_solvers) = -1   ! mg%i_phi of the solver in each slot
  integer, save      :: field_slot = 1               ! slot of the last mg_compute_phi_gradient (for the norm)

  public :: mg_init, mg_destroy, mg_fas_fmg, mg_fas_vcycle, mg_update_operator_stencil
  public :: mg_compute_phi_gradient, mg_compute_field_norm

contains

  !> Slot of the solver that belongs to this mg_t; a free one is claimed when `claim` is set
  integer function slot_of(mg, claim)
    type(mg_t), intent(in) :: mg
    logical, intent(in)    :: claim
    integer                :: n

    do n = 1, max_solvers
       if (slot_key(n) == mg%i_phi) then
          slot_of = n
          return
       end if
    end do

    slot_of = -1
    if (claim) then
       do n = 1, max_solvers
          if (slot_key(n) == -1) then
             slot_key(n) = mg%i_phi
             slot_of = n
             return
          end if
       end do
    end if
    if (slot_of == -1) error stop "m_af_multigrid_dropin: mg_init has not been called for this mg_t"
  end function slot_of

  subroutine mg_init(tree, mg)
    type(af_t), intent(inout) :: tree ! the af_t the solver works on
    type(mg_t), intent(inout) :: mg   ! the caller's mg_t
    call ref_mg_init(tree, mg)
    call mg_gpu_init(tree, mg, slot_of(mg, .true.))
  end subroutine mg_init

  subroutine mg_destroy(mg)
    type(mg_t), intent(inout) :: mg   ! the caller's mg_t
    integer                   :: slot
    slot = slot_of(mg, .false.)
    call mg_gpu_destroy(mg, slot)
    slot_key(slot) = -1
    mg%initialized = .true.           ! the reference's mg_destroy checks it before freeing the coarse solver
    call ref_mg_destroy(mg)
    mg%initialized = .false.
  end subroutine mg_destroy

  subroutine mg_fas_fmg(tree, mg, set_residual, have_guess)
    type(af_t), intent(inout) :: tree         ! the af_t the solver works on
    type(mg_t), intent(inout) :: mg           ! the caller's mg_t
    logical, intent(in)       :: set_residual ! leave rhs - L(phi) in mg%i_tmp on return
    logical, intent(in)       :: have_guess   ! .false.: phi is cleared first
    call mg_gpu_fas_fmg(tree, mg, set_residual, have_guess, slot_of(mg, .false.))
  end subroutine mg_fas_fmg

  subroutine mg_fas_vcycle(tree, mg, set_residual, highest_lvl, standalone)
    type(af_t), intent(inout)     :: tree         ! the af_t the solver works on
    type(mg_t), intent(in)        :: mg           ! the caller's mg_t
    logical, intent(in)           :: set_residual ! leave rhs - L(phi) in mg%i_tmp on return
    integer, intent(in), optional :: highest_lvl  ! cycle only up to this level
    logical, intent(in), optional :: standalone   ! .false. when nested inside an FMG cycle
    call mg_gpu_fas_vcycle(tree, mg, set_residual, slot_of(mg, .false.), highest_lvl, standalone)
  end subroutine mg_fas_vcycle

  subroutine mg_update_operator_stencil(tree, mg, new_lsf, new_eps)
    type(af_t), intent(inout) :: tree
    type(mg_t), intent(inout) :: mg
    logical, intent(in)       :: new_lsf ! the level-set function changed
    logical, intent(in)       :: new_eps ! the permittivity changed
    call ref_mg_update_operator_stencil(tree, mg, new_lsf, new_eps)
    call mg_gpu_update_operator_stencil(tree, mg, slot_of(mg, .false.))
  end subroutine mg_update_operator_stencil

  subroutine mg_compute_phi_gradient(tree, mg, i_fc, fac, i_norm)
    type(af_t), intent(inout)     :: tree
    type(mg_t), intent(in)        :: mg
    integer, intent(in)           :: i_fc ! face-centred variable that receives the field
    real(dp), intent(in)          :: fac  ! field = fac * grad(phi)
    integer, intent(in), optional :: i_norm
    field_slot = slot_of(mg, .false.)
    call mg_gpu_compute_phi_gradient(tree, mg, i_fc, fac, field_slot, i_norm)
  end subroutine mg_compute_phi_gradient

  subroutine mg_compute_field_norm(tree, i_fc, i_norm)
    type(af_t), intent(inout) :: tree
    integer, intent(in)       :: i_fc   ! face-centred field variable
    integer, intent(in)       :: i_norm ! cell-centred variable that receives the norm
    call mg_gpu_compute_field_norm(tree, i_fc, i_norm, field_slot)
  end subroutine mg_compute_field_norm

end module m_af_multigrid_dropin

!!! lbm_gpu_binding.F90 -- ISO_C_BINDING interface to libtaxila_gpu.so (include/taxila_gpu.h).
!!!
!!! Drop this file into src/lbm/ of Taxila-LBM and `use LBM_GPU_Binding_module` from lbm_flow.F90,
!!! lbm_distribution_function.F90 and lbm.F90; INTEGRATION.md lists, procedure by procedure, which
!!! body forwards to which function below.  The image this repository is built in has no Fortran
!!! compiler, PETSc or MPI, so this file has not been compiled here; it is written against the C
!!! header, one interface per export, in the header's order.  shim/replay_bubble_2d.c makes the same
!!! calls from C and is run by the test-suite.
!!!
!!! Conventions (taxila_gpu.h): every function returns a PETSc-style error code (0 = ok); arrays are the
!!! reference's own local ghosted arrays passed as-is; nothing is retained after a call returns.

module LBM_GPU_Binding_module
  use, intrinsic :: iso_c_binding
  implicit none
  private

  integer(c_int), parameter, public :: TXG_NMAX_COMPONENTS = 5    ! lbm_definitions.h:71
  integer(c_int), parameter, public :: TXG_MAX_MINERALS = 100     ! WALL_MAX_MINERALS
  integer(c_int32_t), parameter, public :: TXG_D3Q19_DISCRETIZATION = 1, TXG_D2Q9_DISCRETIZATION = 2
  integer(c_int32_t), parameter, public :: TXG_RELAXATION_MODE_SRT = 0, TXG_RELAXATION_MODE_MRT = 1
  integer(c_int32_t), parameter, public :: TXG_EOS_NULL = 0, TXG_EOS_DENSITY = 1, TXG_EOS_SC = 2, &
       TXG_EOS_PR = 3, TXG_EOS_THERMO = 4

  ! struct txg_config -- field for field.  C arrays a[i][j] are Fortran arrays a(j, i).
  type, bind(C), public :: txg_config
     integer(c_int32_t) :: struct_bytes
     integer(c_int32_t) :: ndims
     integer(c_int32_t) :: discretization
     integer(c_int32_t) :: ncomponents
     integer(c_int32_t) :: NX, NY, NZ
     integer(c_int32_t) :: zs, zl
     integer(c_int32_t) :: periodic(3)
     integer(c_int32_t) :: stencil_size_rho
     integer(c_int32_t) :: relaxation_mode
     integer(c_int32_t) :: isotropy_order
     integer(c_int32_t) :: nminerals
     integer(c_int32_t) :: fluidfluid_forces
     integer(c_int32_t) :: fluidsolid_forces
     integer(c_int32_t) :: body_forces
     integer(c_int32_t) :: use_nonideal_eos
     integer(c_int32_t) :: eos_type(TXG_NMAX_COMPONENTS)
     integer(c_int32_t) :: rank, nranks
     integer(c_int32_t) :: bc_flags(6)
     real(c_double) :: tau(TXG_NMAX_COMPONENTS)
     real(c_double) :: s_c(TXG_NMAX_COMPONENTS)
     real(c_double) :: s_e(TXG_NMAX_COMPONENTS)
     real(c_double) :: s_e2(TXG_NMAX_COMPONENTS)
     real(c_double) :: s_q(TXG_NMAX_COMPONENTS)
     real(c_double) :: s_nu(TXG_NMAX_COMPONENTS)
     real(c_double) :: s_pi(TXG_NMAX_COMPONENTS)
     real(c_double) :: s_m(TXG_NMAX_COMPONENTS)
     real(c_double) :: mm(TXG_NMAX_COMPONENTS)
     real(c_double) :: gf(TXG_NMAX_COMPONENTS, TXG_NMAX_COMPONENTS)   ! gf(mprime, m) = C gf[m][mprime]
     real(c_double) :: eos_rho0(TXG_NMAX_COMPONENTS)
     real(c_double) :: gw(TXG_NMAX_COMPONENTS, TXG_MAX_MINERALS)      ! gw(m, mineral) = C gw[mineral][m]
     real(c_double) :: gvt(3)
     real(c_double) :: null_pressure
     real(c_double) :: reserved_d(8)
     real(c_double) :: eos_psi0(TXG_NMAX_COMPONENTS)
     real(c_double) :: eos_pr_a(TXG_NMAX_COMPONENTS)
     real(c_double) :: eos_pr_b(TXG_NMAX_COMPONENTS)
     real(c_double) :: eos_pr_R(TXG_NMAX_COMPONENTS)
     real(c_double) :: eos_pr_T(TXG_NMAX_COMPONENTS)
     real(c_double) :: eos_pr_Tc(TXG_NMAX_COMPONENTS)
     real(c_double) :: eos_pr_omega(TXG_NMAX_COMPONENTS)
  end type txg_config

  public :: txg_config_defaults, txg_create, txg_destroy, txg_last_error, txg_nccl_unique_id, txg_comm_init
  public :: txg_set_bc_values, txg_set_bc_pressure_outlet
  public :: txg_set_walls, txg_set_rho_u, txg_set_fi, txg_fi_init, txg_update_moments, txg_step
  public :: txg_collision, txg_communicate_fi, txg_stream, txg_bounceback, txg_apply_bcs, txg_update_flux
  public :: txg_get_fi, txg_get_state, txg_get_diagnostics, txg_get_node_class, txg_delta_norm, txg_synchronize
  public :: txg_last_step_ms, txg_enable_kernel_timing, txg_kernel_times, txg_reset_kernel_times
  public :: TxgErrorMessage

  interface
     ! ---- set-up -------------------------------------------------------------------------------
     integer(c_int) function txg_config_defaults(cfg) bind(C, name="txg_config_defaults")
       import :: c_int, txg_config
       type(txg_config), intent(out) :: cfg
     end function txg_config_defaults

     ! FlowCreate / FlowSetUp (lbm_flow.F90:104-154, 378-428)
     integer(c_int) function txg_create(h, cfg, device) bind(C, name="txg_create")
       import :: c_int, c_ptr, txg_config
       type(c_ptr), intent(out) :: h
       type(txg_config), intent(in) :: cfg
       integer(c_int), value :: device
     end function txg_create

     ! FlowDestroy (lbm_flow.F90:156-184)
     integer(c_int) function txg_destroy(h) bind(C, name="txg_destroy")
       import :: c_int, c_ptr
       type(c_ptr), value :: h
     end function txg_destroy

     ! message of the last failing call (h may be c_null_ptr for txg_create itself)
     type(c_ptr) function txg_last_error(h) bind(C, name="txg_last_error")
       import :: c_ptr
       type(c_ptr), value :: h
     end function txg_last_error

     ! replaces the DMDA communicator (lbm_grid.F90:159-212): rank 0 makes the id, MPI_Bcast, all init
     integer(c_int) function txg_nccl_unique_id(id_out) bind(C, name="txg_nccl_unique_id")
       import :: c_int, c_char
       character(kind=c_char), intent(out) :: id_out(128)
     end function txg_nccl_unique_id

     integer(c_int) function txg_comm_init(h, id) bind(C, name="txg_comm_init")
       import :: c_int, c_ptr, c_char
       type(c_ptr), value :: h
       character(kind=c_char), intent(in) :: id(128)
     end function txg_comm_init

     ! ---- state in -----------------------------------------------------------------------------
     ! walls(rgxs:rgxe, rgys:rgye[, rgzs:rgze]) after WallsSetGhostNodes + WallsCommunicate (lbm.F90:162-163,189)
     integer(c_int) function txg_set_walls(h, walls_rg) bind(C, name="txg_set_walls")
       import :: c_int, c_ptr, c_double
       type(c_ptr), value :: h
       real(c_double), intent(in) :: walls_rg(*)
     end function txg_set_walls

     ! BCSetValues result (lbm_bc.F90:215-228): bc%xm_a .. bc%zp_a of one boundary (0-based: BOUNDARY_XM-1 ..)
     integer(c_int) function txg_set_bc_values(h, boundary, vals) bind(C, name="txg_set_bc_values")
       import :: c_int, c_ptr, c_double
       type(c_ptr), value :: h
       integer(c_int), value :: boundary
       real(c_double), intent(in) :: vals(*)
     end function txg_set_bc_values

     ! flow%bc_flags(boundary) = BC_PRESSURE_OUTLET, flow%bc_data(1,boundary) = pressure (lbm_flow.F90:1170-1189)
     integer(c_int) function txg_set_bc_pressure_outlet(h, boundary, pressure) bind(C, name="txg_set_bc_pressure_outlet")
       import :: c_int, c_ptr, c_double
       type(c_ptr), value :: h
       integer(c_int), value :: boundary
       real(c_double), value :: pressure
     end function txg_set_bc_pressure_outlet

     ! LBMInitializeState result (lbm.F90:444-453); u_g may be c_null_ptr (= 0)
     integer(c_int) function txg_set_rho_u(h, rho_rg, u_g) bind(C, name="txg_set_rho_u")
       import :: c_int, c_ptr, c_double
       type(c_ptr), value :: h
       real(c_double), intent(in) :: rho_rg(*)
       type(c_ptr), value :: u_g
     end function txg_set_rho_u

     ! restart / IC from file (lbm.F90:482-544)
     integer(c_int) function txg_set_fi(h, fi_g) bind(C, name="txg_set_fi")
       import :: c_int, c_ptr, c_double
       type(c_ptr), value :: h
       real(c_double), intent(in) :: fi_g(*)
     end function txg_set_fi

     ! FlowFiInit (lbm_flow.F90:923-934)
     integer(c_int) function txg_fi_init(h) bind(C, name="txg_fi_init")
       import :: c_int, c_ptr
       type(c_ptr), value :: h
     end function txg_fi_init

     ! FlowUpdateMoments (lbm_flow.F90:466-478)
     integer(c_int) function txg_update_moments(h) bind(C, name="txg_update_moments")
       import :: c_int, c_ptr
       type(c_ptr), value :: h
     end function txg_update_moments

     ! ---- time stepping ------------------------------------------------------------------------
     ! LBMRun2's loop body nsteps times (lbm.F90:286-361)
     integer(c_int) function txg_step(h, nsteps) bind(C, name="txg_step")
       import :: c_int, c_ptr
       type(c_ptr), value :: h
       integer(c_int), value :: nsteps
     end function txg_step

     integer(c_int) function txg_collision(h) bind(C, name="txg_collision")            ! FlowCollision
       import :: c_int, c_ptr
       type(c_ptr), value :: h
     end function txg_collision
     integer(c_int) function txg_communicate_fi(h) bind(C, name="txg_communicate_fi")  ! DistributionCommunicateFi
       import :: c_int, c_ptr
       type(c_ptr), value :: h
     end function txg_communicate_fi
     integer(c_int) function txg_stream(h) bind(C, name="txg_stream")                  ! FlowStream
       import :: c_int, c_ptr
       type(c_ptr), value :: h
     end function txg_stream
     integer(c_int) function txg_bounceback(h) bind(C, name="txg_bounceback")          ! FlowBounceback
       import :: c_int, c_ptr
       type(c_ptr), value :: h
     end function txg_bounceback
     integer(c_int) function txg_apply_bcs(h) bind(C, name="txg_apply_bcs")            ! FlowApplyBCs
       import :: c_int, c_ptr
       type(c_ptr), value :: h
     end function txg_apply_bcs
     integer(c_int) function txg_update_flux(h) bind(C, name="txg_update_flux")        ! FlowUpdateFlux (device step runs here)
       import :: c_int, c_ptr
       type(c_ptr), value :: h
     end function txg_update_flux

     ! ---- state out ----------------------------------------------------------------------------
     integer(c_int) function txg_get_fi(h, fi_g) bind(C, name="txg_get_fi")
       import :: c_int, c_ptr, c_double
       type(c_ptr), value :: h
       real(c_double), intent(inout) :: fi_g(*)
     end function txg_get_fi

     ! any of rho_rg, u_g, forces_g may be c_null_ptr (pass c_loc(array) otherwise)
     integer(c_int) function txg_get_state(h, rho_rg, u_g, forces_g) bind(C, name="txg_get_state")
       import :: c_int, c_ptr
       type(c_ptr), value :: h, rho_rg, u_g, forces_g
     end function txg_get_state

     ! FlowUpdateDiagnostics (lbm_flow.F90:603-758): owned-only rhot, prs, velt in natural layout
     integer(c_int) function txg_get_diagnostics(h, rhot, prs, velt) bind(C, name="txg_get_diagnostics")
       import :: c_int, c_ptr
       type(c_ptr), value :: h, rhot, prs, velt
     end function txg_get_diagnostics

     integer(c_int) function txg_get_node_class(h, class_rg) bind(C, name="txg_get_node_class")
       import :: c_int, c_ptr, c_int8_t
       type(c_ptr), value :: h
       integer(c_int8_t), intent(out) :: class_rg(*)
     end function txg_get_node_class

     ! DistributionCalcDeltaNorm (lbm_distribution_function.F90:809-833)
     integer(c_int) function txg_delta_norm(h, norm) bind(C, name="txg_delta_norm")
       import :: c_int, c_ptr, c_double
       type(c_ptr), value :: h
       real(c_double), intent(out) :: norm
     end function txg_delta_norm

     integer(c_int) function txg_synchronize(h) bind(C, name="txg_synchronize")
       import :: c_int, c_ptr
       type(c_ptr), value :: h
     end function txg_synchronize

     ! ---- measurement hooks (not reference procedures) -------------------------------------------
     integer(c_int) function txg_last_step_ms(h, ms, launches) bind(C, name="txg_last_step_ms")
       import :: c_int, c_ptr, c_float, c_int64_t
       type(c_ptr), value :: h
       real(c_float), intent(out) :: ms
       integer(c_int64_t), intent(out) :: launches
     end function txg_last_step_ms
     integer(c_int) function txg_enable_kernel_timing(h, on) bind(C, name="txg_enable_kernel_timing")
       import :: c_int, c_ptr
       type(c_ptr), value :: h
       integer(c_int), value :: on
     end function txg_enable_kernel_timing
     integer(c_int) function txg_kernel_times(h, cap, names, ms, launches, n) bind(C, name="txg_kernel_times")
       import :: c_int, c_ptr, c_double, c_int64_t
       type(c_ptr), value :: h
       integer(c_int), value :: cap
       type(c_ptr), intent(out) :: names(*)
       real(c_double), intent(out) :: ms(*)
       integer(c_int64_t), intent(out) :: launches(*)
       integer(c_int), intent(out) :: n
     end function txg_kernel_times
     integer(c_int) function txg_reset_kernel_times(h) bind(C, name="txg_reset_kernel_times")
       import :: c_int, c_ptr
       type(c_ptr), value :: h
     end function txg_reset_kernel_times
  end interface

contains

  ! The library's message for the last failing call as a Fortran string, for
  !   if (ierr_c /= 0) call LBMError(comm, ierr_c, TxgErrorMessage(flow%gpu), ierr)     (lbm_error.F90:30-45)
  function TxgErrorMessage(h) result(msg)
    type(c_ptr), intent(in) :: h
    character(len=:), allocatable :: msg
    type(c_ptr) :: p
    character(kind=c_char), pointer :: s(:)
    integer :: n, i
    p = txg_last_error(h)
    if (.not. c_associated(p)) then
       msg = ''
       return
    end if
    call c_f_pointer(p, s, [512])
    n = 0
    do while (n < 512)
       if (s(n + 1) == c_null_char) exit
       n = n + 1
    end do
    allocate(character(len=n) :: msg)
    do i = 1, n
       msg(i:i) = s(i)
    end do
  end function TxgErrorMessage

end module LBM_GPU_Binding_module

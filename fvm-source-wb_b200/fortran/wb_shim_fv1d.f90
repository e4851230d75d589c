!-----------------------------------------------------------------------------------------------------------
! wb_shim_fv1d.f90 -- ISO_C_BINDING layer that lets the reference's benchmark_1d.f90 (BASELINE config 1, module parameters:
! solver 'FVM' | 'EQL' | 'WB1') call libwbeuler.so.
!
!   evolve(u,u_eq,x)                  benchmark_1d.f90:200-261   -> wb_fv1d_evolve            (scheme = module parameter `solver`)
!   compute_update(u,w_eq,dudt)       :263-377  ('EQL')          -> wb_fv1d_compute_update
!   compute_update_fvm(u,w_eq,dudt)   :454-549  ('FVM')          -> wb_fv1d_compute_update_fvm
!   compute_update_sr(u,w_eq,dudt)    :553-747  ('WB1')          -> wb_fv1d_compute_update_sr
!   compute_max_speed(u,cmax)         :157-170                   -> wb_fv1d_compute_max_speed
!
!   python tools/split_reference.py $REF/benchmark_1d.f90 build/benchmark_1d_driver.f90 157-170 200-261 263-377 454-549 553-747
!   gfortran -O3 -fallow-argument-mismatch $REF/parameters.f90 wb_shim_fv1d.f90 build/benchmark_1d_driver.f90 \
!            -L<repo>/fvm-source-wb_b200/wbeuler -lwbeuler -Wl,-rpath,<repo>/fvm-source-wb_b200/wbeuler -o benchmark_1d_gpu
!
! `program main`, get_x, get_initial_conditions, get_equilibrium_solution and output_file are compiled unchanged (the
! movie snapshots evolve writes when make_movie is set, :252-258, are an output feature the GPU loop does not reproduce).
! (No Fortran compiler in this image: integration recipe, checked textually by tests/test_abi.py.)
!-----------------------------------------------------------------------------------------------------------
module wb_fv1d_binding
  use iso_c_binding
  implicit none

  type, bind(C) :: wb_fv1d_params          ! include/wbeuler.h: wb_fv1d_params (same member order)
     integer(c_int) :: nx, nvar, bc, nequilibrium, solver
     real(c_double) :: gamma, boxlen
     integer(c_int) :: device
  end type wb_fv1d_params

  interface
     integer(c_int) function wb_fv1d_create(h, p) bind(C, name="wb_fv1d_create")
       import :: c_ptr, c_int, wb_fv1d_params
       type(c_ptr), intent(out) :: h
       type(wb_fv1d_params), intent(in) :: p
     end function
     integer(c_int) function wb_fv1d_destroy(h) bind(C, name="wb_fv1d_destroy")
       import :: c_ptr, c_int
       type(c_ptr), value :: h
     end function
     integer(c_int) function wb_fv1d_compute_update(h, u, w_eq, dudt) bind(C, name="wb_fv1d_compute_update")
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       real(c_double), intent(in) :: u(*), w_eq(*)
       real(c_double), intent(out) :: dudt(*)
     end function
     integer(c_int) function wb_fv1d_compute_update_fvm(h, u, w_eq, dudt) bind(C, name="wb_fv1d_compute_update_fvm")
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       real(c_double), intent(in) :: u(*), w_eq(*)
       real(c_double), intent(out) :: dudt(*)
     end function
     integer(c_int) function wb_fv1d_compute_update_sr(h, u, w_eq, dudt) bind(C, name="wb_fv1d_compute_update_sr")
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       real(c_double), intent(in) :: u(*), w_eq(*)
       real(c_double), intent(out) :: dudt(*)
     end function
     integer(c_int) function wb_fv1d_compute_max_speed(h, u, cmax) bind(C, name="wb_fv1d_compute_max_speed")
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       real(c_double), intent(in) :: u(*)
       real(c_double), intent(out) :: cmax
     end function
     integer(c_int) function wb_fv1d_evolve(h, u, w_eq, tend, max_iter, iters, t, last_dt) bind(C, name="wb_fv1d_evolve")
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       real(c_double), intent(inout) :: u(*)
       real(c_double), intent(in) :: w_eq(*)
       real(c_double), value :: tend
       integer(c_int), value :: max_iter
       integer(c_int), intent(out) :: iters
       real(c_double), intent(out) :: t, last_dt
     end function
     function wb_last_error() bind(C, name="wb_last_error") result(msg)
       import :: c_ptr
       type(c_ptr) :: msg
     end function
  end interface

  type(c_ptr), save :: handle = c_null_ptr

contains

  subroutine wb_check(status)
    integer(c_int), intent(in) :: status
    character(kind=c_char), pointer :: msg(:)
    if (status /= 0) then
       call c_f_pointer(wb_last_error(), msg, [256])
       write(*,*) 'wbeuler error', status, ': ', msg(1:index(transfer(msg, repeat(' ',256)), c_null_char)-1)
       stop 1
    end if
  end subroutine wb_check

  subroutine wb_get_handle()
    use parameters, only: c_nx => nx, c_nvar => nvar, c_bc => bc, c_neq => nequilibrium, c_solver => solver, gamma, boxlen
    type(wb_fv1d_params) :: p
    if (c_associated(handle)) return
    p%nx = c_nx; p%nvar = c_nvar; p%bc = c_bc; p%nequilibrium = c_neq
    select case (c_solver)                 ! parameters.f90:8
    case ('FVM'); p%solver = 1
    case ('EQL'); p%solver = 2
    case ('WB1'); p%solver = 3
    case default
       write(*,*) 'wbeuler: unknown solver ', c_solver
       stop 1
    end select
    p%gamma = gamma; p%boxlen = boxlen; p%device = -1
    call wb_check(wb_fv1d_create(handle, p))
  end subroutine wb_get_handle

end module wb_fv1d_binding

! replaces benchmark_1d.f90:157-170
subroutine compute_max_speed(u, cmax)
  use parameters
  use wb_fv1d_binding
  implicit none
  real(kind=8),dimension(1:nvar,1:nx)::u
  real(kind=8)::cmax
  call wb_get_handle()
  call wb_check(wb_fv1d_compute_max_speed(handle, u, cmax))
end subroutine compute_max_speed

! replaces benchmark_1d.f90:263-377 ('EQL')
subroutine compute_update(u, w_eq, dudt)
  use parameters
  use wb_fv1d_binding
  implicit none
  real(kind=8),dimension(1:nvar,1:nx)::u, w_eq, dudt
  call wb_get_handle()
  call wb_check(wb_fv1d_compute_update(handle, u, w_eq, dudt))
end subroutine compute_update

! replaces benchmark_1d.f90:454-549 ('FVM')
subroutine compute_update_fvm(u, w_eq, dudt)
  use parameters
  use wb_fv1d_binding
  implicit none
  real(kind=8),dimension(1:nvar,1:nx)::u, w_eq, dudt
  call wb_get_handle()
  call wb_check(wb_fv1d_compute_update_fvm(handle, u, w_eq, dudt))
end subroutine compute_update_fvm

! replaces benchmark_1d.f90:553-747 ('WB1')
subroutine compute_update_sr(u, w_eq, dudt)
  use parameters
  use wb_fv1d_binding
  implicit none
  real(kind=8),dimension(1:nvar,1:nx)::u, w_eq, dudt
  call wb_get_handle()
  call wb_check(wb_fv1d_compute_update_sr(handle, u, w_eq, dudt))
end subroutine compute_update_sr

! replaces benchmark_1d.f90:200-261: the whole `do while (t < tend)` loop runs on the GPU
subroutine evolve(u, u_eq, x)
  use parameters
  use wb_fv1d_binding
  implicit none
  real(kind=8),dimension(1:nvar,1:nx)::u, u_eq
  real(kind=8),dimension(1:nx)::x
  integer(c_int)::iters
  real(c_double)::t, dt
  call wb_get_handle()
  call wb_check(wb_fv1d_evolve(handle, u, u_eq, tend, -1_c_int, iters, t, dt))
  write(*,*)'time=',iters,t,dt
end subroutine evolve

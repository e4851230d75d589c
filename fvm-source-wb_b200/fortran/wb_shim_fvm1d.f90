!-----------------------------------------------------------------------------------------------------------
! wb_shim_fvm1d.f90 -- ISO_C_BINDING layer that lets the reference's fvm.f90 (BASELINE config 1, module fvm_commons)
! call libwbeuler.so.
!
!   compute_update(u,dudt)       fvm.f90:188-251   -> wb_fvm1d_compute_update
!   compute_max_speed(u,cmax)    fvm.f90:320-336   -> wb_fvm1d_compute_max_speed
!   the RK2 time loop            fvm.f90:56-76     -> wb_fvm1d_evolve  (through `call wb_fvm1d_time_loop(u,t,dt,iter)`)
!
! The time loop of fvm.f90 lives in `program fvm` itself, so the splitter replaces its line range by one call:
!
!   python tools/split_reference.py $REF/fvm.f90 build/fvm_driver.f90 188-251 320-336 \
!          "56-76=  call wb_fvm1d_time_loop(u,t,dt,iter)"
!   gfortran -O3 -fallow-argument-mismatch $REF/fvm_commons.f90 wb_shim_fvm1d.f90 build/fvm_driver.f90 \
!            -L<repo>/fvm-source-wb_b200/wbeuler -lwbeuler -Wl,-rpath,<repo>/fvm-source-wb_b200/wbeuler -o fvm_gpu
!
! condinit, compute_primitive and the output code of `program fvm` are compiled unchanged.
! (This image has no Fortran compiler: the file is the integration recipe; tests/test_abi.py checks every interface block
!  against the C prototypes of include/wbeuler.h and the splitter ranges against the reference text.)
!-----------------------------------------------------------------------------------------------------------
module wb_fvm1d_binding
  use iso_c_binding
  implicit none

  type, bind(C) :: wb_fvm1d_params         ! include/wbeuler.h: wb_fvm1d_params (same member order)
     integer(c_int) :: nx, nvar, bc, source, n
     real(c_double) :: gamma, boxlen
     integer(c_int) :: device
  end type wb_fvm1d_params

  interface
     integer(c_int) function wb_fvm1d_create(h, p) bind(C, name="wb_fvm1d_create")
       import :: c_ptr, c_int, wb_fvm1d_params
       type(c_ptr), intent(out) :: h
       type(wb_fvm1d_params), intent(in) :: p
     end function
     integer(c_int) function wb_fvm1d_destroy(h) bind(C, name="wb_fvm1d_destroy")
       import :: c_ptr, c_int
       type(c_ptr), value :: h
     end function
     integer(c_int) function wb_fvm1d_compute_update(h, u, dudt) bind(C, name="wb_fvm1d_compute_update")
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       real(c_double), intent(in) :: u(*)
       real(c_double), intent(out) :: dudt(*)
     end function
     integer(c_int) function wb_fvm1d_compute_max_speed(h, u, cmax) bind(C, name="wb_fvm1d_compute_max_speed")
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       real(c_double), intent(in) :: u(*)
       real(c_double), intent(out) :: cmax
     end function
     integer(c_int) function wb_fvm1d_evolve(h, u, tend, max_iter, iters, t, last_dt) bind(C, name="wb_fvm1d_evolve")
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       real(c_double), intent(inout) :: u(*)
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
    use fvm_commons, only: c_nx => nx, c_nvar => nvar, c_bc => bc, c_source => source, c_n => n, gamma, boxlen
    type(wb_fvm1d_params) :: p
    if (c_associated(handle)) return
    p%nx = c_nx; p%nvar = c_nvar; p%bc = c_bc; p%source = c_source; p%n = c_n      ! integer,parameter values are not linker symbols
    p%gamma = gamma; p%boxlen = boxlen; p%device = -1
    call wb_check(wb_fvm1d_create(handle, p))
  end subroutine wb_get_handle

end module wb_fvm1d_binding

! replaces fvm.f90:188-251
subroutine compute_update(u, dudt)
  use fvm_commons
  use wb_fvm1d_binding
  implicit none
  real(kind=8),dimension(1:nvar,1:nx)::u, dudt
  call wb_get_handle()
  call wb_check(wb_fvm1d_compute_update(handle, u, dudt))
end subroutine compute_update

! replaces fvm.f90:320-336
subroutine compute_max_speed(u, cmax)
  use fvm_commons
  use wb_fvm1d_binding
  implicit none
  real(kind=8),dimension(1:nvar,1:nx)::u
  real(kind=8)::cmax
  call wb_get_handle()
  call wb_check(wb_fvm1d_compute_max_speed(handle, u, cmax))
end subroutine compute_max_speed

! replaces the `do while(t < tend)` loop of program fvm (fvm.f90:56-76): compute_max_speed, dt = 0.8*dx/cmax/(2n+1), SSP-RK2
subroutine wb_fvm1d_time_loop(u, t, dt, iter)
  use fvm_commons
  use wb_fvm1d_binding
  implicit none
  real(kind=8),dimension(1:nvar,1:nx)::u
  real(kind=8)::t, dt
  integer::iter
  integer(c_int)::iters
  call wb_get_handle()
  call wb_check(wb_fvm1d_evolve(handle, u, tend, -1_c_int, iters, t, dt))
  iter = iters
  write(*,*)'time=',iter,t,dt
end subroutine wb_fvm1d_time_loop

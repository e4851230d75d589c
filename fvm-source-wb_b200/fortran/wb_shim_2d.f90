!-----------------------------------------------------------------------------------------------------------
! wb_shim_2d.f90 -- ISO_C_BINDING layer that lets the reference's benchmark_2d.f90 driver call libwbeuler.so.
!
! It provides replacement external subroutines with the reference's own names and argument lists
! (compute_update_exact, compute_max_speed, evolve), so `program main` (benchmark_2d.f90:4-22), get_coords,
! get_initial_conditions, get_equilibrium_solution and output_file are compiled unchanged.  Because the reference
! keeps program and subroutines in ONE file, build with tools/split_reference.py, which copies the file omitting the
! line ranges of the replaced routines (benchmark_2d.f90:221-260, :264-279, :465-618) -- no other edit:
!
!   python tools/split_reference.py /path/to/reference/benchmark_2d.f90 build/benchmark_2d_driver.f90 221-260 264-279 465-618
!   gfortran -O3 -fallow-argument-mismatch parameters_2d.f90 wb_shim_2d.f90 build/benchmark_2d_driver.f90 \
!            -L<repo>/fvm-source-wb_b200/wbeuler -lwbeuler -Wl,-rpath,<repo>/fvm-source-wb_b200/wbeuler -o benchmark_2d_gpu
!
! (This image has no Fortran compiler, so this file is provided as the integration recipe; the C-ABI it binds is
!  exercised by the C/ctypes tests.)
!-----------------------------------------------------------------------------------------------------------
module wb_fv2d_binding
  use iso_c_binding
  implicit none

  type, bind(C) :: wb_fv2d_params          ! include/wbeuler.h: wb_fv2d_params
     integer(c_int) :: nx, ny, nvar, nequilibrium
     real(c_double) :: gamma, boxlen_x, boxlen_y, cfl
     integer(c_int) :: arith, device, rank, nranks
  end type wb_fv2d_params

  interface
     integer(c_int) function wb_fv2d_create(h, p) bind(C, name="wb_fv2d_create")
       import :: c_ptr, c_int, wb_fv2d_params
       type(c_ptr), intent(out) :: h
       type(wb_fv2d_params), intent(in) :: p
     end function
     integer(c_int) function wb_fv2d_destroy(h) bind(C, name="wb_fv2d_destroy")
       import :: c_ptr, c_int
       type(c_ptr), value :: h
     end function
     integer(c_int) function wb_fv2d_compute_update_exact(h, u, w_eq, dudt) bind(C, name="wb_fv2d_compute_update_exact")
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       real(c_double), intent(in) :: u(*), w_eq(*)
       real(c_double), intent(out) :: dudt(*)
     end function
     integer(c_int) function wb_fv2d_compute_max_speed(h, u, cmax) bind(C, name="wb_fv2d_compute_max_speed")
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       real(c_double), intent(in) :: u(*)
       real(c_double), intent(out) :: cmax
     end function
     integer(c_int) function wb_fv2d_evolve(h, u, w_eq, tend, max_iter, iters, t, last_dt) bind(C, name="wb_fv2d_evolve")
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
    use parameters_2d
    type(wb_fv2d_params) :: p
    if (c_associated(handle)) return
    p%nx = nx; p%ny = ny; p%nvar = nvar; p%nequilibrium = nequilibrium      ! integer,parameter values are not linker symbols
    p%gamma = gamma; p%boxlen_x = boxlen_x; p%boxlen_y = boxlen_y; p%cfl = cfl
    p%arith = 0; p%device = -1; p%rank = 0; p%nranks = 1
    call wb_check(wb_fv2d_create(handle, p))
  end subroutine wb_get_handle

end module wb_fv2d_binding

! replaces benchmark_2d.f90:465-618
subroutine compute_update_exact(u, w_eq, dudt)
  use parameters_2d
  use wb_fv2d_binding
  implicit none
  real(kind=8),dimension(1:nvar,1:nx,1:ny)::u, w_eq, dudt
  call wb_get_handle()
  call wb_check(wb_fv2d_compute_update_exact(handle, u, w_eq, dudt))
end subroutine compute_update_exact

! replaces benchmark_2d.f90:264-279
subroutine compute_max_speed(u, cmax)
  use parameters_2d
  use wb_fv2d_binding
  implicit none
  real(kind=8),dimension(1:nvar,1:nx,1:ny)::u
  real(kind=8)::cmax
  call wb_get_handle()
  call wb_check(wb_fv2d_compute_max_speed(handle, u, cmax))
end subroutine compute_max_speed

! replaces benchmark_2d.f90:221-260: the whole `do while (t < tend)` loop runs on the GPU (state resident in HBM)
subroutine evolve(u, u_eq)
  use parameters_2d
  use wb_fv2d_binding
  implicit none
  real(kind=8),dimension(1:nvar,1:nx,1:ny)::u, u_eq
  integer(c_int)::iters
  real(c_double)::t, dt
  call wb_get_handle()
  call wb_check(wb_fv2d_evolve(handle, u, u_eq, tend, -1_c_int, iters, t, dt))
  write(*,*)'time=',iters,t,dt
end subroutine evolve

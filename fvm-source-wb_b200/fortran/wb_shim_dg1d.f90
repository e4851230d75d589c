!-----------------------------------------------------------------------------------------------------------
! wb_shim_dg1d.f90 -- ISO_C_BINDING layer that lets the reference's dg_with_source.f90 (BASELINE config 2, module dg_commons,
! basis: root legendre.f90) call libwbeuler.so.
!
!   the time loop of program dg          dg_with_source.f90:173-336  -> wb_dg1d_evolve ('RKi', default) | wb_dg1d_evolve_rk
!                                                                       ('RK1'..'RK4') | wb_dg1d_evolve_w ('RKw', 'RKe')
!   compute_update_exact_delta(delta_u,u_eq,dudt)   :1749-2031       -> wb_dg1d_compute_update_exact_delta
!   compute_update_exact(u,u_eq,dudt)               :1380-1744       -> wb_dg1d_compute_update_exact
!   compute_update(u,dudt)                          :807-1028        -> wb_dg1d_compute_update
!   limiter(u) / limiter_TDV(u) / limiter_cons(u)   :414-519 / :523-606 / :610-734 -> wb_dg1d_limiter / _limiter_tdv / _limiter_cons
!   compute_max_speed(u,cmax)                       :1136-1152       -> wb_dg1d_compute_max_speed
!
! The time loop lives in `program dg` itself, so the splitter replaces its line range by one call:
!
!   python tools/split_reference.py $REF/dg_with_source.f90 build/dg_driver.f90 414-519 523-606 610-734 807-1028 1136-1152 \
!          1380-1744 1749-2031 "173-336=  call wb_dg1d_time_loop(u,delta_u,u_eq,u_eq_modes,uinit,t,dt,iter)"
!   gfortran -O3 -fallow-argument-mismatch $REF/dg_commons.f90 $REF/legendre.f90 wb_shim_dg1d.f90 build/dg_driver.f90 \
!            -L<repo>/fvm-source-wb_b200/wbeuler -lwbeuler -Wl,-rpath,<repo>/fvm-source-wb_b200/wbeuler -o dg_gpu
!
! The set-up of program dg (projection of the initial condition and of the equilibrium, :24-171), condinit, get_eq_solution,
! modes_to_nodes / nodes_to_modes and the output code stay Fortran.
! (No Fortran compiler in this image: integration recipe, checked textually by tests/test_abi.py.)
!-----------------------------------------------------------------------------------------------------------
module wb_dg1d_binding
  use iso_c_binding
  implicit none

  type, bind(C) :: wb_dg1d_params          ! include/wbeuler.h: wb_dg1d_params (same member order)
     integer(c_int) :: n, nx, nvar, riemann, source
     real(c_double) :: gamma, boxlen
     integer(c_int) :: device, bc, use_limiter
  end type wb_dg1d_params

  interface
     integer(c_int) function wb_dg1d_create(h, p) bind(C, name="wb_dg1d_create")
       import :: c_ptr, c_int, wb_dg1d_params
       type(c_ptr), intent(out) :: h
       type(wb_dg1d_params), intent(in) :: p
     end function
     integer(c_int) function wb_dg1d_destroy(h) bind(C, name="wb_dg1d_destroy")
       import :: c_ptr, c_int
       type(c_ptr), value :: h
     end function
     integer(c_int) function wb_dg1d_compute_update_exact_delta(h, delta_u, u_eq, dudt) &
          bind(C, name="wb_dg1d_compute_update_exact_delta")
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       real(c_double), intent(in) :: delta_u(*), u_eq(*)
       real(c_double), intent(out) :: dudt(*)
     end function
     integer(c_int) function wb_dg1d_compute_max_speed(h, u_nodes, cmax) bind(C, name="wb_dg1d_compute_max_speed")
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       real(c_double), intent(in) :: u_nodes(*)
       real(c_double), intent(out) :: cmax
     end function
     integer(c_int) function wb_dg1d_evolve(h, delta_u, u_eq, uinit, tend, max_iter, iters, t, last_dt) &
          bind(C, name="wb_dg1d_evolve")
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       real(c_double), intent(inout) :: delta_u(*)
       real(c_double), intent(in) :: u_eq(*)
       real(c_double), intent(inout) :: uinit(*)
       real(c_double), value :: tend
       integer(c_int), value :: max_iter
       integer(c_int), intent(out) :: iters
       real(c_double), intent(out) :: t, last_dt
     end function
     integer(c_int) function wb_dg1d_compute_update(h, u, dudt) bind(C, name="wb_dg1d_compute_update")
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       real(c_double), intent(in) :: u(*)
       real(c_double), intent(out) :: dudt(*)
     end function
     integer(c_int) function wb_dg1d_limiter(h, u) bind(C, name="wb_dg1d_limiter")
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       real(c_double), intent(inout) :: u(*)
     end function
     integer(c_int) function wb_dg1d_evolve_rk(h, integrator_id, u, delta_u, u_eq, uinit, tend, max_iter, iters, t, last_dt) &
          bind(C, name="wb_dg1d_evolve_rk")
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       integer(c_int), value :: integrator_id
       real(c_double), intent(inout) :: u(*)
       real(c_double), intent(in) :: delta_u(*), u_eq(*)
       real(c_double), intent(inout) :: uinit(*)
       real(c_double), value :: tend
       integer(c_int), value :: max_iter
       integer(c_int), intent(out) :: iters
       real(c_double), intent(out) :: t, last_dt
     end function
     integer(c_int) function wb_dg1d_compute_update_exact(h, u, u_eq_modes, dudt) bind(C, name="wb_dg1d_compute_update_exact")
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       real(c_double), intent(in) :: u(*), u_eq_modes(*)
       real(c_double), intent(out) :: dudt(*)
     end function
     integer(c_int) function wb_dg1d_limiter_tdv(h, u) bind(C, name="wb_dg1d_limiter_tdv")
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       real(c_double), intent(inout) :: u(*)
     end function
     integer(c_int) function wb_dg1d_limiter_cons(h, u) bind(C, name="wb_dg1d_limiter_cons")
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       real(c_double), intent(inout) :: u(*)
     end function
     integer(c_int) function wb_dg1d_evolve_w(h, integrator_id, u, delta_u, u_eq_nodes, u_eq_modes, uinit, tend, max_iter, &
          iters, t, last_dt) bind(C, name="wb_dg1d_evolve_w")
       import :: c_ptr, c_int, c_double
       type(c_ptr), value :: h
       integer(c_int), value :: integrator_id
       real(c_double), intent(inout) :: u(*), delta_u(*)
       real(c_double), intent(in) :: u_eq_nodes(*), u_eq_modes(*)
       real(c_double), intent(inout) :: uinit(*)
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
    use dg_commons, only: c_n => n, c_nx => nx, c_nvar => nvar, c_riemann => riemann, c_source => source, c_bc => bc, &
         c_use_limiter => use_limiter, gamma, boxlen
    type(wb_dg1d_params) :: p
    if (c_associated(handle)) return
    p%n = c_n; p%nx = c_nx; p%nvar = c_nvar; p%riemann = c_riemann; p%source = c_source      ! integer,parameter values
    p%gamma = gamma; p%boxlen = boxlen; p%device = -1; p%bc = c_bc
    p%use_limiter = merge(1, 0, c_use_limiter)
    call wb_check(wb_dg1d_create(handle, p))
  end subroutine wb_get_handle

end module wb_dg1d_binding

! replaces dg_with_source.f90:1749-2031
subroutine compute_update_exact_delta(delta_u, u_eq, dudt)
  use dg_commons
  use wb_dg1d_binding
  implicit none
  real(kind=8),dimension(1:nvar,1:n,1:nx)::delta_u, dudt, u_eq
  call wb_get_handle()
  call wb_check(wb_dg1d_compute_update_exact_delta(handle, delta_u, u_eq, dudt))
end subroutine compute_update_exact_delta

! replaces dg_with_source.f90:1380-1744 (u_eq = equilibrium MODES)
subroutine compute_update_exact(u, u_eq, dudt)
  use dg_commons
  use wb_dg1d_binding
  implicit none
  real(kind=8),dimension(1:nvar,1:n,1:nx)::u, dudt, u_eq
  call wb_get_handle()
  call wb_check(wb_dg1d_compute_update_exact(handle, u, u_eq, dudt))
end subroutine compute_update_exact

! replaces dg_with_source.f90:807-1028
subroutine compute_update(u, dudt)
  use dg_commons
  use wb_dg1d_binding
  implicit none
  real(kind=8),dimension(1:nvar,1:n,1:nx)::u, dudt
  call wb_get_handle()
  call wb_check(wb_dg1d_compute_update(handle, u, dudt))
end subroutine compute_update

! replaces dg_with_source.f90:414-519
subroutine limiter(u)
  use dg_commons
  use wb_dg1d_binding
  implicit none
  real(kind=8),dimension(1:nvar,1:n,1:nx)::u
  call wb_get_handle()
  call wb_check(wb_dg1d_limiter(handle, u))
end subroutine limiter

! replaces dg_with_source.f90:523-606
subroutine limiter_TDV(u)
  use dg_commons
  use wb_dg1d_binding
  implicit none
  real(kind=8),dimension(1:nvar,1:n,1:nx)::u
  call wb_get_handle()
  call wb_check(wb_dg1d_limiter_tdv(handle, u))
end subroutine limiter_TDV

! replaces dg_with_source.f90:610-734
subroutine limiter_cons(u)
  use dg_commons
  use wb_dg1d_binding
  implicit none
  real(kind=8),dimension(1:nvar,1:n,1:nx)::u
  call wb_get_handle()
  call wb_check(wb_dg1d_limiter_cons(handle, u))
end subroutine limiter_cons

! replaces dg_with_source.f90:1136-1152 (first node of every cell of a NODAL field)
subroutine compute_max_speed(u, cmax)
  use dg_commons
  use wb_dg1d_binding
  implicit none
  real(kind=8),dimension(1:nvar,1:n,1:nx)::u
  real(kind=8)::cmax
  call wb_get_handle()
  call wb_check(wb_dg1d_compute_max_speed(handle, u, cmax))
end subroutine compute_max_speed

! replaces the `do while(t < tend)` loop of program dg (dg_with_source.f90:173-336) for every integrator of the module
subroutine wb_dg1d_time_loop(u, delta_u, u_eq, u_eq_modes, uinit, t, dt, iter)
  use dg_commons
  use wb_dg1d_binding
  implicit none
  real(kind=8),dimension(1:nvar,1:n,1:nx)::u, delta_u, u_eq, u_eq_modes, uinit
  real(kind=8)::t, dt
  integer::iter
  integer(c_int)::iters
  call wb_get_handle()
  select case (integrator)
  case ('RKi')
     call wb_check(wb_dg1d_evolve(handle, delta_u, u_eq, uinit, tend, -1_c_int, iters, t, dt))
  case ('RK1')
     call wb_check(wb_dg1d_evolve_rk(handle, 1_c_int, u, delta_u, u_eq, uinit, tend, -1_c_int, iters, t, dt))
  case ('RK2')
     call wb_check(wb_dg1d_evolve_rk(handle, 2_c_int, u, delta_u, u_eq, uinit, tend, -1_c_int, iters, t, dt))
  case ('RK3')
     call wb_check(wb_dg1d_evolve_rk(handle, 3_c_int, u, delta_u, u_eq, uinit, tend, -1_c_int, iters, t, dt))
  case ('RK4')
     call wb_check(wb_dg1d_evolve_rk(handle, 4_c_int, u, delta_u, u_eq, uinit, tend, -1_c_int, iters, t, dt))
  case ('RKw')
     call wb_check(wb_dg1d_evolve_w(handle, 5_c_int, u, delta_u, u_eq, u_eq_modes, uinit, tend, -1_c_int, iters, t, dt))
  case ('RKe')
     call wb_check(wb_dg1d_evolve_w(handle, 6_c_int, u, delta_u, u_eq, u_eq_modes, uinit, tend, -1_c_int, iters, t, dt))
  case default
     write(*,*) 'wbeuler: unknown integrator ', integrator
     stop 1
  end select
  iter = iters
  write(*,*)'time=',iter,t,dt
end subroutine wb_dg1d_time_loop

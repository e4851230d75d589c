/* wbeuler.h -- C-ABI of the B200-native explicit time-step hot path of hanveiga/fvm-source-wb.
 *
 * The reference has no FFI: its de-facto ABI is gfortran's external-procedure convention
 * (every argument by reference, column-major real(8) arrays).  Each entry point below names the
 * reference subroutine it replaces (file:line relative to the reference checkout); the
 * ISO_C_BINDING interface blocks that bind them are in fvm-source-wb_b200/fortran/ and
 * INTEGRATION.md.
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error; wb_last_error() gives the message
 *     (thread-local).  The library never prints and never aborts.
 *   - all arithmetic is FP64.  Host arrays keep the reference's Fortran layout byte for byte:
 *       FV 2D  u(nvar,nx,ny)            == C double[ny][nx][4]
 *       DG 2D  u(nvar,nx,ny,mx,my)      == C double[my][mx][ny][nx][4]
 *       1D     u(nvar,nx) / u(nvar,n,nx)== C double[nx][nvar] / double[nx][n][nvar]
 *     the library transposes to structure-of-arrays planes on the device.
 *   - the library never retains host pointers; handles own all device memory; a handle is used
 *     from one host thread at a time.
 *   - there is no CPU fallback: every entry point fails with WB_ERR_CUDA when no sm_100 device
 *     (or no CUDA driver) is present.
 */
#ifndef WBEULER_H
#define WBEULER_H

#ifdef __cplusplus
extern "C" {
#endif

#define WB_OK            0
#define WB_ERR_ARG      -1   /* bad argument / unsupported parameter value            */
#define WB_ERR_CUDA     -2   /* CUDA runtime error, or no usable GPU                  */
#define WB_ERR_NCCL     -3   /* NCCL error, or libnccl could not be loaded            */
#define WB_ERR_STATE    -4   /* call sequence error (e.g. step before upload)         */

const char* wb_last_error(void);
/* library version string, e.g. "wbeuler-b200 0.1 (sm_100a)" */
const char* wb_version(void);
/* number of kernel launches issued by this process' library calls so far (for bench.py gpu_launches) */
long long wb_kernel_launch_count(void);
/* one number in the reference's output format 1PE12.5 (benchmark_2d.f90:139, 2d/benchmark_2d_dg.f90:489): 12 characters + NUL */
void wb_format_1pe12_5(double v, char* out13);

/* ------------------------------------------------------------------------------------------
 * Multi-GPU plumbing: one process per GPU.  Rank 0 calls wb_nccl_get_unique_id(), ships the
 * 128 bytes to the other ranks by whatever the host program has (torch.distributed / MPI),
 * then every rank passes them to the *_comm_init of its handle.
 * ------------------------------------------------------------------------------------------ */
#define WB_NCCL_UNIQUE_ID_BYTES 128
int wb_nccl_get_unique_id(void* id128);

/* ==========================================================================================
 * 2D well-balanced finite volumes -- benchmark_2d.f90 (module parameters_2d.f90)
 * ========================================================================================== */
typedef struct wb_fv2d wb_fv2d;

typedef struct {
  int nx, ny;            /* GLOBAL grid (parameters_2d.f90:3-4)                                  */
  int nvar;              /* must be 4 (parameters_2d.f90:6)                                      */
  int nequilibrium;      /* 1..4 (parameters_2d.f90:14; benchmark_2d.f90:189-216)                */
  double gamma;          /* parameters_2d.f90:19 (the reference value is 1.4 as real(4))         */
  double boxlen_x;       /* parameters_2d.f90:17                                                 */
  double boxlen_y;       /* parameters_2d.f90:18                                                 */
  double cfl;            /* parameters_2d.f90:20                                                 */
  int arith;             /* 0 = fused/fast arithmetic (<=1e-12 of the reference, default)
                            1 = reference operation order (no FMA, IEEE div/sqrt, libm-style exp) */
  int device;            /* CUDA device ordinal; -1 = current device                             */
  int rank, nranks;      /* y-slab decomposition: this handle owns global rows
                            [ny*rank/nranks, ny*(rank+1)/nranks); 0,1 for a single GPU           */
} wb_fv2d_params;

int wb_fv2d_create(wb_fv2d** h, const wb_fv2d_params* p);
int wb_fv2d_destroy(wb_fv2d* h);
/* rows of the global grid owned by this handle: j0 (0-based) and count */
int wb_fv2d_local_rows(const wb_fv2d* h, int* j0, int* nrows);
/* enqueue all work of this handle on a caller-owned cudaStream_t (e.g. torch's current stream) */
int wb_fv2d_set_stream(wb_fv2d* h, void* cuda_stream);
/* nranks > 1 only: create the NCCL communicator used for the per-stage ghost-row send/recv and
 * the per-step max all-reduce */
int wb_fv2d_comm_init(wb_fv2d* h, const void* nccl_unique_id128);
/* how the per-stage ghost rows travel: "p2p" = the boundary-row launch of the stage kernel stores them straight into the
 * neighbours' ghost rows (buffers mapped with CUDA IPC at comm_init, one flag word per direction), "nccl" = ncclSend/ncclRecv
 * (fallback when IPC / peer access is unavailable or WB_FV2D_P2P=0), "none" = single rank.  No reference counterpart. */
const char* wb_fv2d_exchange_kind(const wb_fv2d* h);

/* --- stateless entries: same contract as the Fortran routines (host arrays in, host arrays out;
 *     H2D + kernel + D2H inside the call).  In slab mode the arrays are the LOCAL rows. -------- */
/* replaces compute_update_exact(u,w_eq,dudt)   benchmark_2d.f90:465-618 */
int wb_fv2d_compute_update_exact(wb_fv2d* h, const double* u, const double* w_eq, double* dudt);
/* replaces compute_update(u,w_eq,dudt) (plain, non well-balanced)   benchmark_2d.f90:370-463 */
int wb_fv2d_compute_update(wb_fv2d* h, const double* u, const double* w_eq, double* dudt);
/* replaces compute_max_speed(u,cmax)           benchmark_2d.f90:264-279 (global max over ranks) */
int wb_fv2d_compute_max_speed(wb_fv2d* h, const double* u, double* cmax);
/* replaces evolve(u,u_eq)                      benchmark_2d.f90:221-260
 * loops `do while (t < tend)` (no clamp of the last dt, as the reference); max_iter < 0 = no cap */
int wb_fv2d_evolve(wb_fv2d* h, double* u_inout, const double* w_eq, double tend, int max_iter,
                   int* iters_out, double* t_out, double* last_dt_out);
/* replaces get_equilibrium_solution at cell centres + get_initial_conditions
 * (benchmark_2d.f90:174-218, :45-113) for the local rows; host arrays out (either may be NULL) */
int wb_fv2d_get_initial_conditions(wb_fv2d* h, int ninit, double eta, double* u_out, double* w_eq_out);

/* --- resident path (state stays in HBM between calls) ---------------------------------------- */
int wb_fv2d_upload(wb_fv2d* h, const double* u, const double* w_eq);
/* fill u (ninit 1..4, benchmark_2d.f90:56-109) and the centre equilibrium on the device */
int wb_fv2d_init_device(wb_fv2d* h, int ninit, double eta);
/* enqueue nsteps RK2 steps (2 fused stage kernels each) WITHOUT synchronising; t/iter bookkeeping
 * and the CFL reduction stay on the device.  tend caps as in the reference (steps after t>=tend
 * are no-ops). */
int wb_fv2d_step_async(wb_fv2d* h, int nsteps, double tend);
/* wait for the stream and read the bookkeeping */
int wb_fv2d_sync(wb_fv2d* h, int* iters_out, double* t_out, double* last_dt_out, double* last_cmax_out);
int wb_fv2d_download(wb_fv2d* h, double* u_out);
/* replaces output_file(x,y,u,filen)   benchmark_2d.f90:115-143 for the RESIDENT state: one line '(7(1PE12.5,1X))' per cell,
 * icell outer / jcell inner: x, y, p - p_eq (compute_primitive of the cell, get_equilibrium_solution at its centre).
 * Asynchronous: a packing kernel on the handle's stream, the D2H copy on a private stream and the formatting + file I/O on
 * a host thread, so stepping continues meanwhile (movie snapshots); wb_fv2d_output_wait joins and reports. */
int wb_fv2d_output_file(wb_fv2d* h, const char* path);
int wb_fv2d_output_wait(wb_fv2d* h);
/* reset t = 0, iter = 0 and recompute the max wave speed of the resident state */
int wb_fv2d_reset_clock(wb_fv2d* h);


/* ==========================================================================================
 * 2D modal discontinuous Galerkin -- 2d/benchmark_2d_dg.f90, 2d/legendre.f90, 2d/limiters.f90
 * (module 2d/parameters_dg_2d.f90).  Host layout u(nvar,nx,ny,mx,my) == double[my][mx][ny][nx][4],
 * x,y(nx,ny,mx,my) == double[my][mx][ny][nx].
 * ========================================================================================== */
typedef struct wb_dg2d wb_dg2d;

typedef struct {
  int nx, ny;            /* 2d/parameters_dg_2d.f90:3-4 (nx == ny required: the reference wraps x-face
                            neighbours with ny, 2d/benchmark_2d_dg.f90:1338-1339)                      */
  int mx, my;            /* :5-6, 1..4, mx == my required (the reference mixes the x and y rules)       */
  int nvar;              /* must be 4                                                                    */
  int bc;                /* :18   1 periodic, 2|3 index clamp                                            */
  int source;            /* :20   1 none, 2 gravity (get_source + grad_phi), 3 advection sink            */
  int grad_phi_case;     /* :21   1: g=(x,y) [sic], 2: softened Keplerian centred at (3,3)               */
  int flux_id;           /* :15   0 = as shipped ('llf' matches no branch: numerical flux stays 0),
                                  1 = 'llf1' local Lax-Friedrichs, 2 = 'hll2' (compute_hllflux :1008-1026),
                                  3 = 'hllc' (compute_hllcflux :1030-1134, as shipped incl. its typos)   */
  int limiter_id;        /* :14   0 = use_limiter .false., 1 'ONP', 2 'HIO', 3 '1OR', 4 'LOW', 5 'POS', 6 'PO3' */
  int solver_id;         /* :13   1 'RK4' SSPRK(5,4), 2 'SS4' (same after real(4) rounding), 3 'EQL' RK2,
                                  4 'DEB' forward Euler                                                  */
  int ninit;             /* :17   only used for special_boundary_conditions (ninit == 12)                */
  double gamma, boxlen_x, boxlen_y, cfl, eps, M;   /* :23-35                                             */
  int device;
  int arith;             /* 0 = fused sum-factorised stage kernel (<= 1e-12 of the reference, default; used by evolve /
                            step_async when the limiter is element-local: 'ONP' or none), 1 = reference operation order
                            (bit-for-bit with the CPU restatement; also what the stateless entries always run)        */
  int rank, nranks;      /* y-slab decomposition, one process per GPU: rank r owns global rows [ny*r/R, ny*(r+1)/R);
                            host arrays of a handle hold its own rows only.  nranks <= 1: whole grid (default).
                            Built for the fused stage kernel (arith 0, limiter 'ONP' or none).                        */
} wb_dg2d_params;

int wb_dg2d_create(wb_dg2d** h, const wb_dg2d_params* p);
int wb_dg2d_destroy(wb_dg2d* h);
int wb_dg2d_set_stream(wb_dg2d* h, void* cuda_stream);
/* slab mode: NCCL communicator from the 128-byte id of wb_nccl_get_unique_id (same id on every rank); the ghost rows of
 * modes travel once per RK stage (periodic box: ring, bc 2|3: chain), the order-dependent max-speed scan of
 * compute_max_speed is all-reduced in its two-phase form */
int wb_dg2d_comm_init(wb_dg2d* h, const void* nccl_unique_id_128);
/* like wb_fv2d_exchange_kind: "p2p" = the boundary-row launches of the fused stage kernel store their rows straight into the
 * neighbours' ghost rows (element-local limiter flows; buffers mapped with CUDA IPC at comm_init), "nccl" = pack, ncclSend/Recv,
 * unpack (always used by the neighbour-reading limiter flows and the stateless entries), "none" = single rank. */
const char* wb_dg2d_exchange_kind(const wb_dg2d* h);
int wb_dg2d_local_rows(const wb_dg2d* h, int* j0, int* nrows);
/* which RK-stage kernel evolve / step_async launch on this handle: "split" (k_dg_stage_split: element split over four
 * threads, faces once, rows staged by TMA; nx % 32 == 0), "tma" / "march" / "fast" (one thread per element: earlier
 * data paths, same bits), "reference" (reference operation order: arith 1 or a neighbour-reading limiter) */
const char* wb_dg2d_stage_kernel(const wb_dg2d* h);
/* Gauss-Legendre nodes/weights exactly as gl_quadrature computes them (2d/legendre.f90:77-108) */
int wb_dg2d_quadrature(wb_dg2d* h, double* x_quad, double* w_quad);
/* replaces get_modes_from_nodes / get_nodes_from_modes   2d/benchmark_2d_dg.f90:497-542 / :544-592
 * (also 2d/commons.f90, the routines 2d/test2d.f90 exercises) */
int wb_dg2d_get_modes_from_nodes(wb_dg2d* h, const double* nodes, double* modes);
int wb_dg2d_get_nodes_from_modes(wb_dg2d* h, const double* modes, double* nodes);
/* replaces the arithmetic of compute_error(u,x,y,t,u_anal)   2d/benchmark_2d_dg.f90:23-89: the reference compares the nodal
 * state with get_initial_conditions(x,y) (its `t` and `u_anal` are never used) and prints max |u - u_init|, the L1 sums and
 * sqrt of the L2 sums per variable.  The caller passes the initial nodes (the Fortran initialiser stays on the Fortran side);
 * returned: lmax[4], l1[4] and l2[4] = the accumulators BEFORE the sqrt.  Sums over elements are a fixed-shape tree on the
 * device (deterministic; equal to the reference's sequential sums to a few ulp). */
int wb_dg2d_compute_error(wb_dg2d* h, const double* u_nodes, const double* u_init_nodes, double* lmax4, double* l1_4, double* l2_4);
/* the same norms between the RESIDENT state (reconstructed at the nodes) and the initial condition `ninit` translated by
 * (shift_x, shift_y) in the periodic box, both evaluated on the device: u_anal of an advected profile at time t is the
 * initial state shifted by v*t -- convergence studies at grids whose nodal arrays the host does not hold */
int wb_dg2d_compute_error_resident(wb_dg2d* h, int ninit, double eta, double shift_x, double shift_y, double* lmax4, double* l1_4,
                                   double* l2_4);
/* replaces compute_update(delta_u,x,y,u_eq,dudt)   2d/benchmark_2d_dg.f90:1137-1479 (u_eq is never read there) */
int wb_dg2d_compute_update(wb_dg2d* h, const double* modes, const double* x, const double* y, double* dudt);
/* replaces apply_limiter(u)   2d/benchmark_2d_dg.f90:1516-1555 -> 2d/limiters.f90 */
int wb_dg2d_apply_limiter(wb_dg2d* h, double* modes_inout);
/* replaces compute_max_speed(u(:,:,:,1,1),cs_max,v_xmax,v_ymax,speed_max)   2d/benchmark_2d_dg.f90:826-870;
 * mean_mode is u(nvar,nx,ny) == double[ny][nx][4] */
int wb_dg2d_compute_max_speed(wb_dg2d* h, const double* mean_mode, double* cs_max, double* v_xmax, double* v_ymax,
                              double* speed_max);
/* replaces evolve(u,x,y,u_eq)   2d/benchmark_2d_dg.f90:624-775: nodal values in, nodal values out */
int wb_dg2d_evolve(wb_dg2d* h, double* u_nodes_inout, const double* x, const double* y, double tend, int max_iter,
                   int* iters_out, double* t_out, double* last_dt_out);
/* resident path: upload nodal values (projected to modes and limited on the device, :644,:659), step, download nodes */
int wb_dg2d_upload(wb_dg2d* h, const double* u_nodes, const double* x, const double* y);
/* get_coords + get_initial_conditions (2d/benchmark_2d_dg.f90:93-120, :122-466; every ninit 1..12 of the reference: pulse,
 * hydrostatic + bump, the Riemann problems, isentropic vortex, rotating disks, advection tests, Keplerian disk) on the
 * device, then projection + initial limiter as in evolve -- for grids whose nodal arrays the host cannot hold */
int wb_dg2d_init_device(wb_dg2d* h, int ninit, double eta);
/* the same nodal initial state returned to the host (owned rows), without touching the clock: u(nvar,nx,ny,mx,my) */
int wb_dg2d_get_initial_conditions(wb_dg2d* h, int ninit, double eta, double* u_nodes_out);
int wb_dg2d_step_async(wb_dg2d* h, int nsteps, double tend);
int wb_dg2d_sync(wb_dg2d* h, int* iters_out, double* t_out, double* last_dt_out);
int wb_dg2d_download(wb_dg2d* h, double* u_nodes_out);
int wb_dg2d_download_modes(wb_dg2d* h, double* modes_out);
/* replaces output_file(x,y,nodes,var,filen)   2d/benchmark_2d_dg.f90:468-495 for the RESIDENT state: per element (icell
 * outer) x, y of node (1,1) and w - w_eq of the variables var..nvar there; `nequilibrium` is the module parameter that
 * selects get_equilibrium_solution (:594-622; the shipped 3 = zero).  The movie frames of evolve (:759-766) are this call
 * every `interval` steps between wb_dg2d_step_async calls.  Asynchronous like wb_fv2d_output_file. */
int wb_dg2d_output_file(wb_dg2d* h, int var, int nequilibrium, const char* path);
int wb_dg2d_output_wait(wb_dg2d* h);

/* ==========================================================================================
 * 1D finite volumes.  Host layout u(nvar,nx) == C double[nx][3].
 * ========================================================================================== */
/* ---- fvm.f90 (module fvm_commons.f90): plain first-order FV, LLF, centred gravity source, SSP-RK2 ---- */
typedef struct wb_fvm1d wb_fvm1d;
typedef struct {
  int nx;          /* fvm_commons.f90:6                                   */
  int nvar;        /* must be 3 (:7)                                      */
  int bc;          /* :13   1 periodic, 2 zero gradient                   */
  int source;      /* :14   1 none, 2 gravity (compute_source :253-264)   */
  int n;           /* :4    only enters dt = 0.8*dx/cmax/(2n+1) (fvm.f90:59) */
  double gamma;    /* :17                                                 */
  double boxlen;   /* :16                                                 */
  int device;
} wb_fvm1d_params;
int wb_fvm1d_create(wb_fvm1d** h, const wb_fvm1d_params* p);
int wb_fvm1d_destroy(wb_fvm1d* h);
/* replaces compute_update(u,dudt)      fvm.f90:188-251 */
int wb_fvm1d_compute_update(wb_fvm1d* h, const double* u, double* dudt);
/* replaces compute_max_speed(u,cmax)   fvm.f90:320-336 */
int wb_fvm1d_compute_max_speed(wb_fvm1d* h, const double* u, double* cmax);
/* replaces the main time loop          fvm.f90:56-76 */
int wb_fvm1d_evolve(wb_fvm1d* h, double* u_inout, double tend, int max_iter, int* iters_out, double* t_out,
                    double* last_dt_out);

/* ---- benchmark_1d.f90 (module parameters.f90): 'FVM' | 'EQL' | 'WB1' -------------------------------- */
typedef struct wb_fv1d wb_fv1d;
typedef struct {
  int nx;            /* parameters.f90:3                                                      */
  int nvar;          /* must be 3 (:4)                                                        */
  int bc;            /* :12   1 periodic, 2 zero gradient, 3 reflexive                        */
  int nequilibrium;  /* :13   1|2 isothermal exp(-x), 3 isentropic                            */
  int solver;        /* :8    1 'FVM', 2 'EQL', 3 'WB1' (default)                             */
  double gamma;      /* :17                                                                   */
  double boxlen;     /* :16                                                                   */
  int device;
} wb_fv1d_params;
int wb_fv1d_create(wb_fv1d** h, const wb_fv1d_params* p);
int wb_fv1d_destroy(wb_fv1d* h);
/* replaces compute_update(u,w_eq,dudt) ('EQL')        benchmark_1d.f90:263-377 */
int wb_fv1d_compute_update(wb_fv1d* h, const double* u, const double* w_eq, double* dudt);
/* replaces compute_update_fvm(u,w_eq,dudt) ('FVM')    benchmark_1d.f90:454-549 */
int wb_fv1d_compute_update_fvm(wb_fv1d* h, const double* u, const double* w_eq, double* dudt);
/* replaces compute_update_sr(u,w_eq,dudt) ('WB1')     benchmark_1d.f90:553-747 (w_eq is not read, may be NULL) */
int wb_fv1d_compute_update_sr(wb_fv1d* h, const double* u, const double* w_eq, double* dudt);
/* replaces compute_max_speed(u,cmax)                  benchmark_1d.f90:157-170 */
int wb_fv1d_compute_max_speed(wb_fv1d* h, const double* u, double* cmax);
/* replaces evolve(u,u_eq,x)                           benchmark_1d.f90:200-261 (scheme = params.solver) */
int wb_fv1d_evolve(wb_fv1d* h, double* u_inout, const double* w_eq, double tend, int max_iter, int* iters_out,
                   double* t_out, double* last_dt_out);

/* ==========================================================================================
 * 1D modal DG on the perturbation -- dg_with_source.f90 (module dg_commons.f90, basis: root legendre.f90),
 * default integrator 'RKi'.  Host layout u(nvar,n,nx) == C double[nx][n][3].
 * ========================================================================================== */
typedef struct wb_dg1d wb_dg1d;
typedef struct {
  int n;           /* dg_commons.f90:4 (nquad = n), 1..3                          */
  int nx;          /* :6                                                          */
  int nvar;        /* must be 3 (:8)                                              */
  int riemann;     /* :9    1 riemann_llf, 2 riemann_hllc (default)               */
  int source;      /* :15   1 none, 2 gravity                                     */
  double gamma;    /* :18                                                         */
  double boxlen;   /* :17                                                         */
  int device;
  int bc;          /* :14   1 periodic, 2 zero gradient, 3|4 reflective, 5 none (default; the delta-form
                            update freezes the end cells, the plain update needs 1..4)            */
  int use_limiter; /* :10   moment-limiter part of limiter() on/off (default off)                  */
} wb_dg1d_params;
int wb_dg1d_create(wb_dg1d** h, const wb_dg1d_params* p);
int wb_dg1d_destroy(wb_dg1d* h);
/* gl_quadrature(chsi_quad, w_quad, nquad) of the root legendre.f90:77-128 */
int wb_dg1d_quadrature(wb_dg1d* h, double* chsi_quad, double* w_quad);
/* replaces compute_update_exact_delta(delta_u,u_eq,dudt)   dg_with_source.f90:1749-2031 (u_eq = nodal equilibrium) */
int wb_dg1d_compute_update_exact_delta(wb_dg1d* h, const double* delta_u, const double* u_eq, double* dudt);
/* replaces compute_max_speed(u,cmax)   dg_with_source.f90:1136-1152 (first node of every cell of a NODAL field) */
int wb_dg1d_compute_max_speed(wb_dg1d* h, const double* u_nodes, double* cmax);
/* replaces the main time loop with integrator 'RKi'   dg_with_source.f90:173-336 (:282-305): delta_u (modes) and
 * uinit (nodal state used for the time step, = u_eq + reconstructed delta) are updated in place */
int wb_dg1d_evolve(wb_dg1d* h, double* delta_u_inout, const double* u_eq, double* uinit_inout, double tend, int max_iter,
                   int* iters_out, double* t_out, double* last_dt_out);
/* replaces compute_update(u,dudt) on the full state   dg_with_source.f90:807-1028 (bc 1..4) */
int wb_dg1d_compute_update(wb_dg1d* h, const double* u, double* dudt);
/* replaces limiter(u)   dg_with_source.f90:414-519 (Krivodonova moment limiter in characteristic variables when
 * use_limiter, then the positivity fallback) */
int wb_dg1d_limiter(wb_dg1d* h, double* u_inout);
/* replaces the main time loop with integrator 'RK1' (1) .. 'RK4' (4)   dg_with_source.f90:173-227, :313-336.
 * delta_u is constant on these paths; it only feeds the nodal state `uinit` the time step is computed from. */
int wb_dg1d_evolve_rk(wb_dg1d* h, int integrator, double* u_inout, const double* delta_u, const double* u_eq,
                      double* uinit_inout, double tend, int max_iter, int* iters_out, double* t_out, double* last_dt_out);
/* replaces compute_update_exact(u,u_eq_modes,dudt)   dg_with_source.f90:1380-1744: full-state modes against the
 * equilibrium MODES; bc 4 | 5 (with any other bc the reference uses out-of-bounds face states) */
int wb_dg1d_compute_update_exact(wb_dg1d* h, const double* u, const double* u_eq_modes, double* dudt);
/* replaces limiter_TDV(u)   :523-606 -- only with use_limiter = .false. (the shipped value): its moment-limiting block
 * indexes the neighbours with a stale loop variable (:550-552), what remains is the positivity fallback on the traces */
int wb_dg1d_limiter_tdv(wb_dg1d* h, double* u_inout);
/* replaces limiter_cons(u)   :610-734 */
int wb_dg1d_limiter_cons(wb_dg1d* h, double* u_inout);
/* replaces the main time loop with integrator 'RKw' (5, :229-270) or 'RKe' (6, :273-280).  u (full-state modes, 'RKw'),
 * delta_u (perturbation modes; input of 'RKe', output of both) and uinit are updated in place */
int wb_dg1d_evolve_w(wb_dg1d* h, int integrator, double* u_inout, double* delta_u_inout, const double* u_eq_nodes,
                     const double* u_eq_modes, double* uinit_inout, double tend, int max_iter, int* iters_out, double* t_out,
                     double* last_dt_out);

#ifdef __cplusplus
}
#endif
#endif /* WBEULER_H */

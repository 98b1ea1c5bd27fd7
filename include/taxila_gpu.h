/* taxila_gpu.h -- C ABI of the B200 flow hot path for Taxila-LBM.
 *
 * This is the drop-in boundary.  The reference (Fortran-90 + PETSc, no FFI of its
 * own) drives its flow update through a fixed set of module procedures called
 * from src/lbm/lbm.F90; each export below names the procedure(s) whose body an
 * ISO_C_BINDING shim forwards here (see INTEGRATION.md for the shim).  All
 * citations are relative to the reference tree.
 *
 * Conventions
 *  - plain C: pointers and sizes only, no C++/torch types.
 *  - every call returns 0 on success or a non-zero PETSc-style error code
 *    (the reference's convention is `PetscErrorCode ierr` + LBMError/SETERRQ,
 *    src/lbm/lbm_error.F90:30-45); the message is kept per handle and is read
 *    with txg_last_error().  CUDA and NCCL failures are never swallowed.
 *  - host arrays are the reference's own LOCAL GHOSTED Fortran arrays, passed
 *    as-is (column-major, first index fastest):
 *        walls(rgxs:rgxe, rgys:rgye[, rgzs:rgze])            ghost width R
 *        rho  (S, rgxs:rgxe, rgys:rgye[, rgzs:rgze])         ghost width R
 *        fi   (S, 0:b, gxs:gxe, gys:gye[, gzs:gze])          ghost width 1
 *        u, forces (S, ndims, gxs:gxe, gys:gye[, gzs:gze])   ghost width 1
 *    (src/lbm/lbm_flow.F90:963-970), R = stencil_size_rho (lbm_grid.F90:107-120).
 *    Only owned entries are read/written unless stated; nothing handed across
 *    the ABI is retained after the call returns.
 *  - one handle per rank == one GPU.  Calls on a handle come from one thread.
 *    With nranks > 1 every rank makes the same sequence of calls (NCCL inside).
 *  - decomposition: x and y are never split; rank r owns the contiguous z-slab
 *    [zs, zs+zl) of a D3Q19 box (run the reference with -da_processors_x 1
 *    -da_processors_y 1).  D2Q9 runs on one rank.
 */
#ifndef TAXILA_GPU_H
#define TAXILA_GPU_H

#include <stdint.h>

#if defined(__GNUC__)
#define TXG_API __attribute__((visibility("default")))
#else
#define TXG_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* ---- constants mirrored from include/lbm_definitions.h ---------------------- */
#define TXG_D3Q19_DISCRETIZATION 1 /* lbm_definitions.h:10 */
#define TXG_D2Q9_DISCRETIZATION 2  /* lbm_definitions.h:11 */
#define TXG_RELAXATION_MODE_SRT 0  /* lbm_definitions.h:65 */
#define TXG_RELAXATION_MODE_MRT 1  /* lbm_definitions.h:66 */
#define TXG_EOS_NULL 0             /* lbm_definitions.h:79-83 */
#define TXG_EOS_DENSITY 1
#define TXG_EOS_SC 2
#define TXG_EOS_PR 3
#define TXG_EOS_THERMO 4
#define TXG_NMAX_COMPONENTS 5      /* lbm_definitions.h:71 */
#define TXG_MAX_MINERALS 100       /* WALL_MAX_MINERALS, lbm_definitions.h:105 */
/* wall codes, stored by the reference as doubles (lbm_definitions.h:104-111) */
#define TXG_WALL_PORESPACE 0.0
#define TXG_WALL_NONREACTIVE 800.0
#define TXG_WALL_NORMAL_X 900.0
#define TXG_WALL_NORMAL_Y 901.0
#define TXG_WALL_NORMAL_Z 902.0
#define TXG_WALL_GHOST 999.0
/* bc%flags values (lbm_definitions.h:29-35) and boundary numbers (:45-50, here 0-based) */
#define TXG_BC_NULL 0
#define TXG_BC_PERIODIC 1
#define TXG_BC_REFLECTING 2
#define TXG_BC_DIRICHLET 3
#define TXG_BC_NEUMANN 4
#define TXG_BC_VELOCITY 5
#define TXG_BOUNDARY_XM 0
#define TXG_BOUNDARY_XP 1
#define TXG_BOUNDARY_YM 2
#define TXG_BOUNDARY_YP 3
#define TXG_BOUNDARY_ZM 4
#define TXG_BOUNDARY_ZP 5
/* device node classes (u8) the wall codes are mapped to, one-to-one */
#define TXG_CLASS_PORE 0u          /* 1..100 = mineral id, unchanged */
#define TXG_CLASS_NORMAL_X 250u
#define TXG_CLASS_NORMAL_Y 251u
#define TXG_CLASS_NORMAL_Z 252u
#define TXG_CLASS_NONREACTIVE 253u
#define TXG_CLASS_OTHER_SOLID 254u /* any other positive code: plain bounce-back */
#define TXG_CLASS_GHOST 255u

/* ---- flat POD configuration --------------------------------------------------
 * Everything FlowSetFromOptions / ComponentSetFromOptions / RelaxationSetFromOptions
 * / MineralSetFromOptions / InfoSetFromOptions leave in flow_type, component_type,
 * relaxation_type, mineral_type and info_type that the hot path reads
 * (lbm_flow.F90:37-82,186-290; lbm_component.F90:131-176; lbm_relaxation.F90:117-151;
 * lbm_mineral.F90:103-134; lbm_info.F90:25-46). */
typedef struct txg_config {
  int32_t struct_bytes;     /* = sizeof(txg_config); guards against ABI drift      */
  int32_t ndims;            /* 2 | 3                                                */
  int32_t discretization;   /* TXG_D2Q9_DISCRETIZATION | TXG_D3Q19_DISCRETIZATION   */
  int32_t ncomponents;      /* S, 1..TXG_NMAX_COMPONENTS                            */
  int32_t NX, NY, NZ;       /* global box (NZ = 1 for ndims == 2)                   */
  int32_t zs, zl;           /* owned z-slab of this rank: 0-based start, length     */
  int32_t periodic[3];      /* info%periodic(X,Y,Z)                                 */
  int32_t stencil_size_rho; /* R: ghost width of rho/walls host arrays (1,2,3)      */
  int32_t relaxation_mode;  /* TXG_RELAXATION_MODE_*                                */
  int32_t isotropy_order;   /* 4 | 8 | 10 (10: D2Q9 only), lbm_discretization.F90:78-89 */
  int32_t nminerals;        /* options%nminerals, lbm_options.F90:136               */
  int32_t fluidfluid_forces;/* flow%fluidfluid_forces, lbm_flow.F90:233-239         */
  int32_t fluidsolid_forces;/* options%flow_fluidsolid_forces, lbm_walls.F90:117-125*/
  int32_t body_forces;      /* flow%body_forces (-gvt given), lbm_flow.F90:226-231  */
  int32_t use_nonideal_eos; /* flow%use_nonideal_eos, lbm_options.F90:142           */
  int32_t eos_type[TXG_NMAX_COMPONENTS]; /* TXG_EOS_DENSITY | SC | PR | THERMO (lbm_eos.F90:104-141)      */
  int32_t rank, nranks;     /* position in the z-slab ring                          */
  int32_t bc_flags[6];      /* bc%flags(BOUNDARY_XM..ZP), lbm_bc.F90:36: TXG_BC_*; 0 = none  */
  /* per component m (relaxation_type, component_type) */
  double tau[TXG_NMAX_COMPONENTS];   /* SRT; s_c = 1/tau (lbm_relaxation.F90:133)   */
  double s_c[TXG_NMAX_COMPONENTS];   /* MRT rates, lbm_relaxation.F90:134-149       */
  double s_e[TXG_NMAX_COMPONENTS];
  double s_e2[TXG_NMAX_COMPONENTS];
  double s_q[TXG_NMAX_COMPONENTS];
  double s_nu[TXG_NMAX_COMPONENTS];
  double s_pi[TXG_NMAX_COMPONENTS];
  double s_m[TXG_NMAX_COMPONENTS];
  double mm[TXG_NMAX_COMPONENTS];    /* molecular mass; d_k = 1 - 2/(3 mm)          */
  double gf[TXG_NMAX_COMPONENTS][TXG_NMAX_COMPONENTS]; /* gf[m][m'] = option -g_<m+1><m'+1> */
  double eos_rho0[TXG_NMAX_COMPONENTS];                /* EOS_SC / EOS_THERMO rho0 (lbm_eos.F90:212,247) */
  double gw[TXG_MAX_MINERALS][TXG_NMAX_COMPONENTS];    /* gw[mineral-1][m]          */
  double gvt[3];                     /* body acceleration, lbm_flow.F90:226-231     */
  double null_pressure;              /* flow%null_pressure, written to prs on walls */
  double reserved_d[8];
  /* eos_type fields of EOSSetFromOptions_Thermo / _PR (lbm_eos.F90:236-320), per component */
  double eos_psi0[TXG_NMAX_COMPONENTS];     /* EOS_THERMO psi0                                   */
  double eos_pr_a[TXG_NMAX_COMPONENTS];     /* EOS_PR a (default 2/49), b (2/21), R (1)          */
  double eos_pr_b[TXG_NMAX_COMPONENTS];
  double eos_pr_R[TXG_NMAX_COMPONENTS];
  double eos_pr_T[TXG_NMAX_COMPONENTS];     /* temperature and critical temperature as set up by */
  double eos_pr_Tc[TXG_NMAX_COMPONENTS];    /* EOSSetFromOptions_PR (:287-309)                   */
  double eos_pr_omega[TXG_NMAX_COMPONENTS]; /* acentric factor (default: the default-real 0.344) */
} txg_config;

typedef struct txg_flow *txg_handle;

/* Fill *cfg with the reference's defaults (OptionsCreate lbm_options.F90:93-152,
 * RelaxationCreate lbm_relaxation.F90:60-83: tau = s_* = 1, mm = 1, isotropy 4,
 * nminerals 1, everything periodic = 0).  Not a reference procedure. */
TXG_API int txg_config_defaults(txg_config *cfg);

/* FlowCreate/FlowSetUp (lbm_flow.F90:104-154,378-428): validates cfg, selects
 * cuda device `device`, allocates the device-resident SoA lattice for the slab. */
TXG_API int txg_create(txg_handle *h, const txg_config *cfg, int device);

/* FlowDestroy (lbm_flow.F90:156-184). */
TXG_API int txg_destroy(txg_handle h);

/* Message of the last failing call on this handle ("" if none).  h may be NULL
 * for failures of txg_create itself. */
TXG_API const char *txg_last_error(txg_handle h);

/* Multi-GPU wiring (replaces the DMDA communicator set up in lbm_grid.F90:159-212).
 * Rank 0 calls txg_nccl_unique_id and broadcasts the 128 bytes with whatever
 * the host program has (MPI_Bcast in the Fortran driver, torch.distributed in
 * bench.py); then every rank calls txg_comm_init.  Not needed for nranks == 1. */
TXG_API int txg_nccl_unique_id(unsigned char id_out[128]);
TXG_API int txg_comm_init(txg_handle h, const unsigned char id[128]);

/* WallsSetValues + WallsSetGhostNodes + WallsCommunicate result
 * (lbm_walls.F90:151-244, called lbm.F90:162-163,189): the local ghosted
 * walls(rg..) array of doubles.  Classified into u8 node classes on the device. */
TXG_API int txg_set_walls(txg_handle h, const double *walls_rg);

/* BCSetValues result (lbm_bc.F90:215-228; filled by the user's initialize_bcs or by
 * FlowSetUpBCsD2/D3, lbm_flow.F90:1311-1956): the face array of boundary
 * `boundary` (TXG_BOUNDARY_*), this rank's part, in the reference's own layout
 *   xm/xp_vals(nbcs, ys:ye[, zs:ze])  ym/yp_vals(nbcs, xs:xe[, zs:ze])  zm/zp_vals(nbcs, xs:xe, ys:ye)
 * with nbcs = ndims * ncomponents (lbm_flow.F90:245) viewed by the node routines
 * as (S, ndims): densities in (m,1) for BC_DIRICHLET, momentum (m,d) for BC_NEUMANN,
 * velocity (1,d) for BC_VELOCITY (lbm_bc.F90:1273-1333,1533-1593,1793-1865).
 * Only faces whose bc_flags entry is DIRICHLET / NEUMANN / VELOCITY read values; a BC_REFLECTING face
 * (BCApplyReflectingD3/D2, lbm_bc.F90:809-1073) needs none.
 * May be called again at any time (the outlet updates of FlowApplyBCs,
 * lbm_flow.F90:1958-1991, stay on the host and re-upload). */
TXG_API int txg_set_bc_values(txg_handle h, int boundary, const double *vals);

/* flow%bc_flags(boundary) = BC_PRESSURE_OUTLET with flow%bc_data(1,boundary) = pressure
 * (FlowParseBC, lbm_flow.F90:1170-1189).  The face must be BC_DIRICHLET in bc_flags and its array
 * uploaded with txg_set_bc_values.  With two components every step then starts FlowApplyBCs the way the
 * reference does (FlowUpdateBCPressureOutlet + FlowUpdateDensityFromPressure, lbm_flow.F90:1993-2263):
 * the face densities are re-derived on the device from the pressure and the phase fraction of the node one
 * step inside.  Fails like the reference for a non-ideal EOS or g_11 /= 0 (:2000-2005) and for more than
 * two components (:2260).  The flux-outlet update (FlowUpdateBCFluxOutlet, :2265-2497) tests for
 * BC_PRESSURE_OUTLET faces only (:2301 ... :2480), so a BC_FLUX_OUTLET face keeps its constant BC_NEUMANN values:
 * nothing to call.  (A run with both kinds of outlet would have that routine overwrite entries of the
 * pressure-outlet faces; that interplay is not reproduced.) */
TXG_API int txg_set_bc_pressure_outlet(txg_handle h, int boundary, double pressure);

/* LBMInitializeState result (lbm.F90:444-453): host rho(S,rg..) and u(S,ndims,g..)
 * as filled by the user's initialize_state.  u may be NULL (= 0, what every
 * shipped initialize_state sets). */
TXG_API int txg_set_rho_u(txg_handle h, const double *rho_rg, const double *u_g);

/* Restart / IC-from-file (lbm.F90:482-544): host fi(S,0:b,g..). */
TXG_API int txg_set_fi(txg_handle h, const double *fi_g);

/* FlowFiInit (lbm_flow.F90:923-934): forces from rho0, fi = (1 - prefactor/2) feq(rho0, u0). */
TXG_API int txg_fi_init(txg_handle h);

/* FlowUpdateMoments (lbm_flow.F90:466-478): rho, flux, forces, common velocity
 * from the current fi.  On the device these are recomputed inside every step, so
 * this only refreshes the exported copies (txg_get_*). */
TXG_API int txg_update_moments(txg_handle h);

/* LBMRun2 inner body (lbm.F90:286-361), nsteps times: FlowCollision,
 * DistributionCommunicateFi, FlowStream (BCPreStream + stream), FlowBounceback
 * (plain and free-slip 900-902 walls), FlowApplyBCs (FlowCalcRhoForces, BCApply,
 * BCUpdateRho for the faces flagged in bc_flags), FlowUpdateFlux.  Asynchronous on
 * the handle's streams. */
TXG_API int txg_step(txg_handle h, int nsteps);

/* The six reference procedures individually, for a shim that keeps LBMRun2's
 * loop body unchanged.  The fused device step runs when txg_update_flux is
 * reached; the other five only check the call order (a call out of the
 * reference's order is an error, not a silent no-op). */
TXG_API int txg_collision(txg_handle h);      /* FlowCollision             lbm_flow.F90:936  */
TXG_API int txg_communicate_fi(txg_handle h); /* DistributionCommunicateFi lbm_distribution_function.F90:309 */
TXG_API int txg_stream(txg_handle h);         /* FlowStream                lbm_flow.F90:810  */
TXG_API int txg_bounceback(txg_handle h);     /* FlowBounceback            lbm_flow.F90:816  */
TXG_API int txg_apply_bcs(txg_handle h);      /* FlowApplyBCs              lbm_flow.F90:1958 */
TXG_API int txg_update_flux(txg_handle h);    /* FlowUpdateFlux            lbm_flow.F90:458  */

/* FlowGetArrays view of the state (lbm_flow.F90:431-436): copy the device state
 * back into the reference's host arrays.  Any pointer may be NULL.  Synchronises.
 *   fi_g      fi(S,0:b,g..)     post-stream/bounce-back populations (what -output_flow_fi writes)
 *   rho_rg    rho(S,rg..)       DistributionCalcDensity (owned entries; ghosts untouched)
 *   u_g       flux(S,ndims,g..) common velocity u' after FlowUpdateUE (lbm_flow.F90:494-574)
 *   forces_g  forces(S,ndims,g..) after FlowCalcForces (lbm_flow.F90:760-808)          */
TXG_API int txg_get_fi(txg_handle h, double *fi_g);
TXG_API int txg_get_state(txg_handle h, double *rho_rg, double *u_g, double *forces_g);

/* FlowUpdateDiagnostics (lbm_flow.F90:603-758): owned-only arrays in the
 * reference's global (natural) layout: rhot(x,y,z), prs(x,y,z), velt(ndims,x,y,z). */
TXG_API int txg_get_diagnostics(txg_handle h, double *rhot, double *prs, double *velt);

/* The device node-class array of the slab with its ghost layers, as u8 in the
 * layout of walls(rg..) -- for the bit-exact classification check. */
TXG_API int txg_get_node_class(txg_handle h, uint8_t *class_rg);

/* DistributionCalcDeltaNorm (lbm_distribution_function.F90:809-833), fi variant:
 * max |(fi_old - fi)/fi| over owned entries, then fi_old = fi.  First call
 * returns 1e99 like the reference's initial value. */
TXG_API int txg_delta_norm(txg_handle h, double *norm);

/* Block until all device work queued on the handle is done. */
TXG_API int txg_synchronize(txg_handle h);

/* ---- measurement hooks (not reference procedures) ---------------------------- */
/* Device-side time of the last txg_step call in milliseconds (CUDA events on the
 * handle's compute stream), kernel launches it issued, and accumulated per-kernel
 * time: names/ms/launches for up to `cap` kernels; returns the count in *n.      */
TXG_API int txg_last_step_ms(txg_handle h, float *ms, int64_t *launches);
TXG_API int txg_enable_kernel_timing(txg_handle h, int on);
TXG_API int txg_kernel_times(txg_handle h, int cap, const char **names, double *ms, int64_t *launches, int *n);
TXG_API int txg_reset_kernel_times(txg_handle h);

#ifdef __cplusplus
}
#endif
#endif /* TAXILA_GPU_H */

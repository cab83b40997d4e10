/*
 * pbx.h -- C ABI of poissbox-b200: the B200 (sm_100a) implementation of the compact-scheme
 * gradient / divergence / Laplacian hot path of 3decomp/poissbox, its batched tridiagonal solves,
 * and the conjugate-gradient loop that drives the Laplacian.
 *
 * This is the drop-in boundary.  Each entry point names the reference interface it replaces
 * (file:line relative to the reference tree); INTEGRATION.md shows the ISO_C_BINDING module and
 * the PETSc MATSHELL glue a maintainer adds on the reference side.
 *
 * Conventions
 *   - All arrays are IEEE fp64 in Fortran column-major order: f(i,j,k) <-> f[i + nx*(j + ny*k)],
 *     df(i,j,k,c) adds c*nx*ny*nz (src/compact_schemes.f90:19-23,46).
 *   - Every function returns 0 on success or a PBX_ERR_* code; PBX_ERR_SIZE is 7 on purpose: it
 *     is what the reference reports with `stop 7` (src/compact_schemes.f90:177-180, 292-295).
 *   - Caller owns every array; outputs are fully overwritten.  Scratch and coefficient tables
 *     belong to a handle (pbx_create) -- there is no hidden global state except the handle
 *     cache behind the *_host convenience calls: ONE per process (at most four (box, device)
 *     entries), under one mutex, so *_host calls from several threads serialise (their copies and
 *     the stream synchronisation included), and pbx_host_set_mode applies to every thread.
 *   - `*_device` calls take device pointers and are asynchronous on the handle's CUDA stream.
 *     `*_host` calls take host pointers, stage H2D/D2H themselves and return when the result is
 *     in the caller's buffer.
 *   - There is no CPU fallback: without a CUDA device every compute call returns PBX_ERR_CUDA.
 *   - One handle per rank/GPU; calls on one handle must not be issued concurrently.
 */
#ifndef PBX_H
#define PBX_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PBX_VERSION 100

#define PBX_OK 0
#define PBX_ERR_ARG 1         /* bad argument (null pointer, n too small, unknown enum) */
#define PBX_ERR_CUDA 2        /* CUDA runtime error or no device; see pbx_last_error() */
#define PBX_ERR_NCCL 3        /* NCCL error */
#define PBX_ERR_UNSUPPORTED 4 /* valid request this build cannot serve */
#define PBX_ERR_NOMEM 5
#define PBX_ERR_SIZE 7        /* array-length mismatch: the reference's `stop 7` */

/* Laplacian schedules.  Both compute src/compact_schemes.f90:17-37 (lapl = div(grad)).
 *   REFERENCE  the reference's own order of operations (16 line-operator sweeps, sequential
 *              Thomas + Sherman-Morrison per line, true division, no FMA contraction):
 *              bit-identical to the CPU oracle.
 *   FAST       one sweep per axis (the 1-D operators of different axes commute), each sweep a
 *              chunked constant-coefficient recursion held in registers; agrees with REFERENCE to
 *              a few 1e-16 of max|result| (not bit-identical; tests/test_parity_gpu.py). */
#define PBX_MODE_FAST 0
#define PBX_MODE_REFERENCE 1

/* stagger argument of the 1-D/3-D operators, as in the reference's opt_stagger:
 * -1 = cell -> vertex (default of grad_1d / interp_1d), +1 = vertex -> cell (div_1d, *_div). */
#define PBX_STAGGER_BACKWARD (-1)
#define PBX_STAGGER_FORWARD (+1)

/* KSPConvergedReason values used by pbx_cg_solve (PETSc numbering) */
#define PBX_CONVERGED_RTOL 2
#define PBX_CONVERGED_ATOL 3
#define PBX_DIVERGED_ITS (-3)
#define PBX_DIVERGED_DTOL (-4)
#define PBX_DIVERGED_INDEFINITE_MAT (-10)
#define PBX_DIVERGED_NANORINF (-9)

typedef struct pbx_handle_s *pbx_handle;

int pbx_version(void);
const char *pbx_error_string(int code);
/* text of the last failure on the calling thread ("" if none) */
const char *pbx_last_error(void);
/* number of CUDA devices visible (0 when there is none or the driver is missing) */
int pbx_device_count(void);

/* ---------------------------------------------------------------------------------------------
 * Lifecycle.  A handle binds a local brick nx x ny x nz_local, the grid spacings, a device, a
 * stream and (optionally) an NCCL communicator for the z-slab decomposition.
 * Replaces: the per-call allocate/deallocate of the reference (src/compact_schemes.f90:30,59,69,
 * 183-185,225,235) and the shell-matrix context `mat_ctx` (src/poissbox.f90:17-20).
 *   nccl_comm  NULL for a single GPU; otherwise an ncclComm_t whose rank r owns global planes
 *              [r*nz, (r+1)*nz) of a periodic box of nz*nranks planes.
 * ------------------------------------------------------------------------------------------- */
int pbx_create(int nx, int ny, int nz, const double dx[3], int device, void *nccl_comm,
               pbx_handle *h);
int pbx_destroy(pbx_handle h);
int pbx_set_mode(pbx_handle h, int mode);
int pbx_get_mode(pbx_handle h, int *mode);
/* cudaStream_t; NULL selects the legacy default stream */
int pbx_set_stream(pbx_handle h, void *stream);
int pbx_synchronize(pbx_handle h);
/* sizes the handle was created with */
int pbx_get_dims(pbx_handle h, int *nx, int *ny, int *nz);
/* number of kernels this handle has launched since creation (bench.py's gpu_launches) */
long long pbx_launch_count(pbx_handle h);

/* NCCL bootstrap helpers so that a host language without NCCL bindings can build the
 * communicator: rank 0 fills `id` (128 bytes), ships it to the other ranks by any means, and
 * every rank calls pbx_comm_init_rank.  The returned pointer is an ncclComm_t. */
int pbx_comm_unique_id(void *id128);
int pbx_comm_init_rank(const void *id128, int nranks, int rank, int device, void **comm);
int pbx_comm_destroy(void *comm);

/* ---------------------------------------------------------------------------------------------
 * z-slab decomposition driven phase by phase, without NCCL: for emulating P ranks in one process
 * (tests on a single GPU) or for a host that owns the exchange itself.  A handle made by
 * pbx_create with an ncclComm_t runs the same three steps inside pbx_lapl_device.
 *   pbx_slab_phase1   x and y sweeps + the R <= 8 boundary "moments" per z-line for each neighbour
 *   (exchange)        rank r's send-up array goes to rank r+1, its send-down array to rank r-1
 *   pbx_slab_phase2   z sweep on the open slab + low-rank boundary corrections -> d2f
 * pbx_dist_tables_host exposes the correction tables (host, no GPU needed): for each boundary
 * s = 0 (bottom) / 1 (top), U[s][48][8], VnbM/VnbD[s][48][8] (neighbour planes) and
 * VsM/VsD[s][48][8] (own planes, first *ncs rows), numerical rank R[s].
 * ------------------------------------------------------------------------------------------- */
int pbx_create_slab(int nx, int ny, int nz_local, const double dx[3], int device, int rank,
                    int nranks, pbx_handle *h);
int pbx_slab_phase1(pbx_handle h, const double *f);
int pbx_slab_phase2(pbx_handle h, double *d2f);
int pbx_slab_exchange_local(pbx_handle *hs, int n);
/* grad / div / interp on slabs, phase-driven in the same way: phase 1 = the stages before the z
 * operators plus the boundary sweeps of their inputs (three numbers per z line, operator and
 * neighbour); exchange; phase 2 = the z operators and the remaining stages.  `in` must be the same
 * array in both phases.  With an ncclComm_t, pbx_grad_device / pbx_div_device / pbx_interp_device
 * run the three steps themselves. */
#define PBX_OP_GRAD 1
#define PBX_OP_DIV 2
#define PBX_OP_INTERP 3      /* stagger -1 */
#define PBX_OP_INTERP_DIV 4  /* stagger +1 */
#define PBX_OP_STAR 5        /* the 2nd-order star: one plane travels each way */
int pbx_slab_op_phase1(pbx_handle h, int op, const double *in);
int pbx_slab_op_phase2(pbx_handle h, int op, const double *in, double *out);
/* For a host that owns the exchange itself (MPI, sockets, ...): after phase 1 copy the two outgoing
 * messages out, move them to the neighbours by any means, install what arrived, run phase 2.
 * Every message is pbx_slab_message_count doubles (nine planes of nx*ny); all pointers are device
 * pointers; `up` goes to rank+1 (which installs it as from_lo), `dn` to rank-1 (as from_up). */
int pbx_slab_message_count(pbx_handle h, long long *count);
int pbx_slab_get_messages(pbx_handle h, double *up, double *dn);
int pbx_slab_put_messages(pbx_handle h, const double *from_lo, const double *from_up);
/* Peer boards: ranks that can store into one another's memory (slab handles of one process on
 * one device or on peer-enabled devices; NVLink peer mappings the host has opened itself; in the
 * tests, processes sharing a mapping) need neither NCCL nor a host-owned exchange.  Every rank owns
 * a receive buffer of pbx_slab_recv_bytes bytes (pbx_slab_recv_buffer: the one the handle allocated;
 * 128-byte aligned, zeroed), and pbx_slab_link_peers hands the handle all n of them, bufs[r] being
 * rank r's buffer as addressable from this handle's device (bufs[rank]: its own; a pointer other
 * than the handle's buffer replaces it and stays owned by the caller, who has zeroed it BEFORE any
 * rank links).  From then on the boundary sweeps store their messages straight into the neighbours'
 * buffers, the exchange is a flag barrier between neighbours, the CG's sums are all-reduced inside
 * its reduction kernels, and pbx_lapl_device / pbx_grad_device / ... / pbx_cg_solve_device work on
 * the handle as they do on one created with an NCCL communicator.  Each rank needs its own stream
 * and host thread: the kernels wait on the device for their peers.  (With a communicator,
 * pbx_create sets up the same thing over cudaIpc by default; PBX_PEER_SYNC=0 keeps the NCCL collectives.) */
int pbx_slab_recv_bytes(pbx_handle h, size_t *bytes);
int pbx_slab_recv_buffer(pbx_handle h, void **buf);
int pbx_slab_link_peers(pbx_handle h, void *const *bufs, int n);
/* 1 when the ranks of this handle synchronise and reduce over the peer boards (no NCCL call per apply / iteration),
 * 0 when they go through NCCL (PBX_PEER_SYNC=0, PBX_NO_PEER=1, or peer mappings that could not be opened), negative
 * error code for a handle without a decomposition.  Diagnostic: so that a host can report which path it timed. */
int pbx_peer_sync_active(pbx_handle h);
/* the exchange step alone (NCCL communicator or peer boards), and the CG's scalar all-reduce
 * (exposed for profiling the communication steps on their own) */
int pbx_slab_exchange(pbx_handle h);
int pbx_allreduce_sum(pbx_handle h, double *dev, int count);
int pbx_dist_tables_host(int nzl, double dz, int *ncs, int *nrow, int R[2], double *U,
                         double *VnbM, double *VsM, double *VnbD, double *VsD);

/* ---------------------------------------------------------------------------------------------
 * 3-D compact operators on device-resident fields (asynchronous on the handle's stream).
 * ------------------------------------------------------------------------------------------- */
/* compact_schemes::lapl  src/compact_schemes.f90:17-37;  also the body of the MATSHELL MatMult
 * callback `mfmult` src/poissbox.f90:300-322 once it is re-pointed at the compact operator. */
int pbx_lapl_device(pbx_handle h, const double *f, double *d2f);
/* as above, and additionally *dot_dev (device double) <- sum f * d2f (the CG p.Ap, fused) */
int pbx_lapl_dot_device(pbx_handle h, const double *f, double *d2f, double *dot_dev);
/* Measurement aid (bench.py): runs the FAST Laplacian `reps` times and returns the average
 * duration in milliseconds of each of its three kernels (x, y, z pass), taken with CUDA events
 * recorded between the launches on the handle's stream.  Synchronises the stream. */
int pbx_lapl_profile_device(pbx_handle h, const double *f, double *d2f, int reps, double ms[3]);
/* grad / div / interp follow the handle's mode: REFERENCE = thread-per-line Thomas sweeps, FAST =
 * chunked-recursion line operators in the same stage order (pbx_fast_lineop.cu). */
/* compact_schemes::grad  src/compact_schemes.f90:42-88;  df has 3 components */
int pbx_grad_device(pbx_handle h, const double *f, double *df);
/* compact_schemes::div   src/compact_schemes.f90:207-257;  f has 3 components */
int pbx_div_device(pbx_handle h, const double *f, double *df);
/* compact_schemes::interp / interp_div  src/compact_schemes.f90:93-152 */
int pbx_interp_device(pbx_handle h, const double *f, double *fi, int stagger);

/* ---------------------------------------------------------------------------------------------
 * The 2nd-order 7-point star on the periodic box: what the reference's MatMult callback applies
 * today (mfmult -> compute_lapl_pointwise, src/poissbox.f90:84-148 and :300-322, coefficients
 * src/coefficients.f90:22-48) and the matrix P it preconditions with (:294).  Bit-identical to the
 * reference's dot_product over the 3x3x3 box.
 *   pbx_set_operator   which operator the shell matrix is: PBX_OPERATOR_COMPACT (default, the
 *                      compact Laplacian the drop-in re-points mfmult at) or PBX_OPERATOR_STAR
 *   pbx_matmult_device y = A x for the handle's operator: the body of the MATSHELL callback; the
 *                      CG (pbx_cg_solve_*) runs on the same operator
 * ------------------------------------------------------------------------------------------- */
#define PBX_OPERATOR_COMPACT 0
#define PBX_OPERATOR_STAR 1
int pbx_set_operator(pbx_handle h, int op);
int pbx_get_operator(pbx_handle h, int *op);
int pbx_matmult_device(pbx_handle h, const double *x, double *y);
int pbx_star_device(pbx_handle h, const double *x, double *y);
int pbx_star_host(int nx, int ny, int nz, const double *x, const double dx[3], double *y);

/* ---------------------------------------------------------------------------------------------
 * Batched 1-D compact operators: one call applies grad_1d / interp_1d to `nlines` periodic lines
 * of n points.  Point i of line l lives at base[l*line_stride + i*elem_stride] (in doubles), for
 * both input and output, which must not overlap.
 * Replaces: grad_1d :155-204, div_1d :260-268, interp_1d :271-319, interp_1d_div :322-329 and
 * eval_1d_rhs :332-372 of src/compact_schemes.f90, looped over lines at :60-86 and :226-253.
 * ------------------------------------------------------------------------------------------- */
int pbx_grad_1d_batch_device(int n, long long nlines, long long elem_stride, long long line_stride,
                             const double *f, double dx, double *df, int stagger, void *stream);
int pbx_interp_1d_batch_device(int n, long long nlines, long long elem_stride,
                               long long line_stride, const double *f, double *fi, int stagger,
                               void *stream);

/* ---------------------------------------------------------------------------------------------
 * Batched general-coefficient tridiagonal solves (module tridsol, src/tridsol.f90:16-18).
 * a = sub-diagonal, b = DIAGONAL, c = super-diagonal, d = rhs/solution -- the reference's dummy
 * names (its in-file comments swap b and c; tests/tridiag/test_tdma.f90:42-44 has it right).
 * All four arrays share the layout base[l*line_stride + i*elem_stride].
 *   tdma           :22-32   non-periodic; b is overwritten with the pivots exactly as :92 does
 *   tdma_periodic  :34-74   Sherman-Morrison closure; b is left untouched
 *   fwd_sweep      :76-96   b and d overwritten
 *   bwd_sweep      :98-115  d overwritten
 * ------------------------------------------------------------------------------------------- */
int pbx_tdma_batch_device(int n, long long nlines, long long elem_stride, long long line_stride,
                          const double *a, double *b, const double *c, double *d, void *stream);
int pbx_tdma_periodic_batch_device(int n, long long nlines, long long elem_stride,
                                   long long line_stride, const double *a, const double *b,
                                   const double *c, double *d, void *stream);
int pbx_fwd_sweep_batch_device(int n, long long nlines, long long elem_stride,
                               long long line_stride, const double *a, double *b, const double *c,
                               double *d, void *stream);
int pbx_bwd_sweep_batch_device(int n, long long nlines, long long elem_stride,
                               long long line_stride, const double *b, const double *c, double *d,
                               void *stream);

/* ---------------------------------------------------------------------------------------------
 * Conjugate gradients on the compact Laplacian, entirely on the device, with the semantics of
 * PETSc's `-ksp_type cg -pc_type none` as the reference's solve() sets it up
 * (src/poissbox.f90:269-298: constant MatNullSpace on A, KSPSetFromOptions, KSPSolve):
 * x0 = 0, z = r - mean(r), preconditioned norm ||z||_2, converged when
 * ||z|| <= max(rtol*||z_0||, abstol), diverged at 1e4*||z_0|| or max_it.
 *   hist  may be NULL; else receives ||z|| for iterations 0..its (at most nhist values).
 * The call returns when the solve has finished.  While it runs the host thread stays one iteration ahead of the
 * device: every kernel that changes solver state checks the device status word, and the host learns the status from a
 * single word the device posts into pinned, mapped host memory each iteration (no copy or event in the stream); the
 * wait for that word is bounded by the stream's own state, so a failed kernel surfaces as PBX_ERR_CUDA, not as a hang.
 * ------------------------------------------------------------------------------------------- */
int pbx_cg_solve_device(pbx_handle h, const double *b, double *x, double rtol, double abstol,
                        int maxit, int *its, double *rnorm, int *reason, double *hist, int nhist);

/* Preconditioner of the CG.  The reference gives KSP the assembled 2nd-order star
 * as the preconditioning matrix (KSPSetOperators(ksp, A, P), src/poissbox.f90:294) and its README
 * runs `-pc_type gamg`: a multigrid preconditioner built on P.  PBX_PC_MG fills that role with a
 * geometric V(nu, nu) cycle on the same star (damped Jacobi, cell-centred trilinear transfer,
 * symmetric positive definite; pbx_mg.cu).  It is not PETSc's GAMG: iteration counts are its own.
 * KSPCG semantics with a preconditioner: z = M^-1 r with the constant removed, preconditioned norm
 * ||z||, beta = z.r, KSP_DIVERGED_INDEFINITE_PC (-8) if beta < 0.  On z slabs the same cycle runs
 * with halo exchanges on the distributed levels and gathered coarse levels.
 *   pbx_set_pc(h, PBX_PC_NONE, 0)  (default)   |   pbx_set_pc(h, PBX_PC_MG, nu)  nu = 0 -> 2
 *   pbx_pc_apply_device            z = M^-1 (r - mean r), mean-free: one application, for tests */
#define PBX_PC_NONE 0
#define PBX_PC_MG 1
#define PBX_DIVERGED_INDEFINITE_PC (-8)
int pbx_set_pc(pbx_handle h, int pc, int nu);
int pbx_pc_apply_device(pbx_handle h, const double *r, double *z);

/* The solve configured the way the reference configures it: by PETSc option names.  solve() calls
 * KSPSetFromOptions (src/poissbox.f90:295) and the README runs
 *     -ksp_type cg -pc_type gamg -ksp_rtol ... -ksp_monitor -ksp_converged_reason   (README.md:43-49);
 * a host without PETSc passes the same string here.  Understood (PETSc's defaults where absent):
 *     -ksp_type cg                      anything else: PBX_ERR_UNSUPPORTED
 *     -ksp_rtol r  -ksp_atol a  -ksp_max_it n          (1e-5, 1e-50, 10000)
 *     -pc_type none | mg | gamg         gamg selects the geometric multigrid stand-in (PBX_PC_MG)
 *     -pc_mg_smoothup n / -pc_mg_smoothdown n / -mg_levels_ksp_max_it n    sweeps of the V(n, n) cycle
 *     -ksp_monitor                      "  k KSP Residual norm x" per iteration on stdout, after the solve
 *     -ksp_converged_reason             "Linear solve converged due to CONVERGED_RTOL iterations k"
 * Other options are ignored, as PETSc ignores options nobody queries.  The handle's preconditioner
 * setting is changed only if -pc_type is given. */
int pbx_ksp_solve_device(pbx_handle h, const char *options, const double *b, double *x, int *its,
                         double *rnorm, int *reason);

/* ---------------------------------------------------------------------------------------------
 * Host-pointer convenience variants (what the Fortran module bodies call).
 * They use a cached handle for (nx,ny,nz,dx) on the current device, copy in, run, copy out.
 * PBX_MODE_FAST needs extents that are multiples of 16; for any other box these calls run the
 * REFERENCE schedule instead (same result to rounding, slower) -- they never fail on the shape.
 * ------------------------------------------------------------------------------------------- */
int pbx_lapl_host(int nx, int ny, int nz, const double *f, const double dx[3], double *d2f,
                  int mode);
/* `count` independent fields of the same box through one call, double-buffered on three streams
 * (copy-in of the next field and copy-out of the previous one overlap the compute of the current
 * one): f[k] -> d2f[k].  Host buffers should be pinned for the overlap to take place. */
int pbx_lapl_host_batch(int nx, int ny, int nz, int count, const double *const *f, const double dx[3],
                        double *const *d2f, int mode);
int pbx_grad_host(int nx, int ny, int nz, const double *f, const double dx[3], double *df);
int pbx_div_host(int nx, int ny, int nz, const double *f, const double dx[3], double *df);
int pbx_interp_host(int nx, int ny, int nz, const double *f, double *fi, int stagger);
/* single-line forms with the reference's size check (nf != ndf -> PBX_ERR_SIZE) */
int pbx_grad_1d_host(int nf, const double *f, double dx, int ndf, double *df, int stagger);
int pbx_interp_1d_host(int nf, const double *f, int nfi, double *fi, int stagger);
int pbx_tdma_host(int n, const double *a, double *b, const double *c, double *d);
int pbx_tdma_periodic_host(int n, const double *a, const double *b, const double *c, double *d);
int pbx_fwd_sweep_host(int n, const double *a, double *b, const double *c, double *d);
int pbx_bwd_sweep_host(int n, const double *b, const double *c, double *d);
int pbx_cg_solve_host(int nx, int ny, int nz, const double dx[3], const double *b, double *x,
                      double rtol, double abstol, int maxit, int mode, int *its, double *rnorm,
                      int *reason, double *hist, int nhist);
/* schedule used by the grad / div / interp host variants (they carry no mode argument, as the
 * reference's routines do not): PBX_MODE_REFERENCE (default, bit-identical to the reference order
 * of operations) or PBX_MODE_FAST (chunked-recursion line operators, agrees to rounding) */
int pbx_host_set_mode(int mode);
/* drop the cached handles of the *_host calls (frees their device memory) */
int pbx_host_cache_clear(void);

#ifdef __cplusplus
}
#endif
#endif /* PBX_H */

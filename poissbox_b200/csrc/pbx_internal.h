// pbx_internal.h -- shared declarations of the poissbox-b200 CUDA library (not part of the ABI).
#pragma once

#include <cstdlib>
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/pbx.h"

struct CUtensorMap_st;

namespace pbx {

// ------------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------------
// A/B switches of kernel variants (read per call): `NAME=0` / `NAME=1` override the built-in default.
// They exist so that a variant can be timed against the one it replaced; results are the same bits.
inline bool env_switch(const char *name, bool dflt)
{
    const char *e = getenv(name);
    if (!e || !e[0]) return dflt;
    return e[0] != '0';
}

void set_last_error(const std::string &msg);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);

#define PBX_CUDA(call)                                                          \
    do {                                                                        \
        cudaError_t e__ = (call);                                               \
        if (e__ != cudaSuccess) return ::pbx::cuda_fail(e__, #call, __FILE__, __LINE__); \
    } while (0)

#define PBX_TRY(call)                   \
    do {                                \
        int rc__ = (call);              \
        if (rc__ != PBX_OK) return rc__; \
    } while (0)

// ------------------------------------------------------------------------------------------------
// Scheme constants (src/compact_schemes.f90:188-190 and :303-305)
// ------------------------------------------------------------------------------------------------
enum OpKind { OP_INTERP = 0, OP_DERIV = 1 };

inline double scheme_alpha(OpKind k) { return k == OP_DERIV ? 9.0 / 62.0 : 3.0 / 10.0; }
inline void scheme_ab(OpKind k, double dx, double *a, double *b)
{
    if (k == OP_DERIV) {
        *a = 63.0 / 62.0 / dx;          // :188
        *b = 17.0 / 62.0 / (3.0 * dx);  // :189
    } else {
        *a = 0.75;                      // :303
        *b = 1.0 / 20.0;                // :304
    }
}

// ------------------------------------------------------------------------------------------------
// FAST schedule coefficients.
//
// A 1-D composite operator  O = P^+ P^-  (P = derivative or interpolation; ^- cell->vertex,
// ^+ vertex->cell) is, because every factor is a circulant,  O = A^-2 S  with A = circ[al,1,al]
// and S = B^+ B^- the symmetric 7-point product of the two 4-point right-hand-side stencils.
// A = (1 - r E^-1)(1 - r E)/(1 + r^2) with r the root of al r^2 + r + al = 0 inside the unit
// circle (E = shift), so A^-2 is two causal first-order recursions followed by two anti-causal
// ones; the factor (1 + r^2)^2 is folded into S.
// ------------------------------------------------------------------------------------------------
constexpr int LC = 16;        // points per thread chunk, held in registers
constexpr int MAXLOOK = 4;    // chunk-state look-back depth (r^(16*4) * 64 < 1e-29 for r = 1/3)

struct CompositeCoef {
    double c0, c1, c2, c3;    // S, scaled by (1 + r^2)^2
    double r;                 // recursion pole (negative)
    double pw[LC];            // r^(k+1), k = 0..LC-1
    double look[MAXLOOK];     // r^(LC*m), m = 0..MAXLOOK-1  (look[0] = 1)
    int nlook;                // look-back levels actually needed for < 1e-17 truncation
    int pad_;
};

void make_composite_coef(OpKind kind, double dx, CompositeCoef *out);

// ------------------------------------------------------------------------------------------------
// REFERENCE schedule tables: the data-independent part of tdma_periodic on [al,1,al]
// (src/tridsol.f90:51-70), computed on the host with the same IEEE operations the reference
// performs, so that the device sweep reproduces the oracle bit for bit.
// ------------------------------------------------------------------------------------------------
struct RefLineTables {
    int n = 0;
    double alpha = 0;
    double *w = nullptr;     // device: w(i) = a(i)/bmod(i-1), i = 1..n-1 (0-based; w[0] unused)
    double *piv = nullptr;   // device: bmod after the forward sweep
    double *u = nullptr;     // device: solution of the auxiliary system (:62-66)
    double a1g = 0;          // a(1)/gamma
    double den = 0;          // 1 + (u(1) + (a(1)/gamma) u(n))
};
int make_ref_tables(int n, double alpha, RefLineTables *t);
void free_ref_tables(RefLineTables *t);

// ------------------------------------------------------------------------------------------------
// kernels' host-side launchers
// ------------------------------------------------------------------------------------------------
struct Brick {
    int nx, ny, nz;
    size_t N() const { return (size_t)nx * ny * nz; }
};

// REFERENCE-order 1-D line operator over every line of a brick along `dir`, or over a strided batch
int ref_line_op(cudaStream_t s, int n, long long nlines1, long long nlines2, long long elem_stride,
                long long line_stride1, long long line_stride2, OpKind kind, int stagger, double dx,
                const RefLineTables &tab, const double *in, double *out, long long *launches);
int ref_add(cudaStream_t s, size_t N, const double *a, const double *b, double *out,
            long long *launches);

// general-coefficient batched tridiagonal kernels (src/tridsol.f90)
int tdma_fwd_batch(cudaStream_t s, int n, long long nl, long long es, long long ls, const double *a,
                   double *b, const double *c, double *d);
int tdma_bwd_batch(cudaStream_t s, int n, long long nl, long long es, long long ls, const double *b,
                   const double *c, double *d);
void tdma_trim_workspace();
int tdma_periodic_batch(cudaStream_t s, int n, long long nl, long long es, long long ls,
                        const double *a, const double *b, const double *c, double *d);

// z-slab decomposition: low-rank boundary corrections (pbx_dist_tables.cu)
constexpr int DIST_NB = 48;     // neighbour planes that matter (r^48 * 48 < 1e-21)
constexpr int DIST_RMAX = 8;    // storage stride of the moment index (numerical rank is 7 / 5)
struct DistSide {
    int R = 0;
    std::vector<double> U, VnbM, VsM, VnbD, VsD;   // row-major, DIST_RMAX columns
};
struct DistTables {
    int nzl = 0, ncs = 0, nrow = 0;
    DistSide side[2];   // 0 = bottom boundary ("A"), 1 = top boundary ("B")
};
int build_dist_tables(int nzl, const CompositeCoef &cm, const CompositeCoef &cd, DistTables *T);

// What the z pass needs to know when the brick is one slab of a z-decomposed box (pbx_dist.cu).
// The slab is an OPEN line; everything the rest of the periodic line contributes enters through
// the pass's native inputs at the two slab boundaries -- recursion states and stencil halos --
// reconstructed per z-line from DIST_MSG numbers received from each neighbour:
//   from the lower rank (its "up" message):  0 yM, 1 zM, 2 zM', 3 zM''   causal state of the
//       interpolation recursion at its top plane and the two previous z values,
//       4 yD, 5 zD   causal state of the derivative recursion on its zero-halo stencil,
//       6 d(-1), 7 d(-2), 8 d(-3)   its top three planes of the derivative input;
//   from the upper rank (its "down" message):  0 A0M, 1 A1M, 2 A0D, 3 A1D   the moments
//       sum r^j u_j and sum j r^j u_j of its bottom planes (u = c, resp. its zero-halo stencil of d),
//       4 d(0), 5 d(1), 6 d(2),  7 c(0), 8 c(1).
constexpr int DIST_MSG = 9;
struct ZOpen {
    int open = 0;                                   // 0: periodic line (single rank); 1: slab
    const double *from_lo = nullptr, *from_up = nullptr;   // device, [DIST_MSG][nlines]
    long long nlines = 0;
    // response of the anti-causal state (w, x) at the first plane above the slab, per operator
    // (0 = interpolation, 1 = derivative):
    //   to the neighbour's moments:        w = gw A0,            x = gx0 A0 + gw A1
    //   to my own outgoing causal state:   w = kwz Z + kwy Y,    x = kxz Z + kxy Y
    double gw[2] = {0, 0}, gx0[2] = {0, 0};
    double kwy[2] = {0, 0}, kwz[2] = {0, 0}, kxy[2] = {0, 0}, kxz[2] = {0, 0};
    double rinv = 0;                                // 1 / r of the interpolation recursion
};

// Peer boards (pbx_dist.cu): a small block of flags and records at the end of every rank's receive
// buffer, stored into by the other ranks over NVLink peer mappings (or, for slab handles linked by
// the host with pbx_slab_link_peers, through whatever shared mapping the host provides).  With the
// boards in place the ranks synchronise and reduce among themselves without NCCL:
//   * neighbour barrier of the slab exchange: each rank stores a sequence number into its two
//     neighbours' boards after its boundary sweep and waits for theirs;
//   * all-reduce of up to PEER_VALS doubles: every rank stores its partial into record
//     [seq % PEER_RING][rank] of EVERY board, waits for all records of its own board and sums them
//     in rank order -- the same bits on every rank.  A rank can be at most one reduction ahead of
//     the slowest one, so a ring of PEER_RING records is never overrun.
constexpr int PEER_MAXR = 16;
constexpr int PEER_RING = 4;
constexpr int PEER_VALS = 4;
struct PeerRec {
    double v[PEER_VALS];
    unsigned long long seq;
    unsigned long long pad_[3];
};
struct PeerBoard {
    unsigned long long bar_from_lo, pad0_[15];
    unsigned long long bar_from_up, pad1_[15];
    PeerRec rec[PEER_RING][PEER_MAXR];
};
struct PeerLinks {
    PeerBoard *board[PEER_MAXR];   // board[r]: rank r's board as addressable from this device
    int n, rank, lower, upper;
};

// Reduction tail (pbx_cg_dev.cuh): what the kernel that writes per-CTA partial sums needs in order to
// reduce them itself, all-reduce the result over the peer boards and run the scalar step of the CG.
struct RedTail {
    int on = 0;                    // 0: the kernel leaves the reduction to later launches
    unsigned *ticket = nullptr;    // device counter, zero between launches
    double *sc = nullptr;          // the CG's scalar block
    double *dst = nullptr;         // where the sums go (inside the scalar block)
    double *hist = nullptr;
    int nhist = 0;
    const double *part = nullptr;  // the partial sums: narr arrays of cnt values, `stride` apart
    int cnt = 0, stride = 0, narr = 0;
    int phase = -1, guarded = 0;
    PeerLinks L{};                 // L.n <= 1: single rank
    unsigned long long seq = 0;
};

// Long lines.  A CTA holds at most SEG_T chunks (512 points) of a y or z line.  A longer line is cut
// into segments of `iseg` chunks that are computed independently as OPEN lines of SEG_T chunks:
// `hlo` halo chunks below and the rest above the segment are loaded (wrapping around the periodic
// line) and computed but not stored.  A chunk's recursion state depends on the nlook <= 3
// preceding chunks only and its solved stencil halo on one chunk more, so SEG_HALO = 4 halo
// chunks reproduce the whole-line result to the truncation the look-back already makes (< 1e-19).
constexpr int SEG_T = 32;
constexpr int SEG_HALO = 4;
struct SegGeom {
    int nseg;   // 1: the whole periodic line in one CTA (T = NC, hlo = 0, iseg = NC)
    int iseg;   // interior chunks per segment (a multiple of 4 when nseg > 1)
    int hlo;    // halo chunks in front of the interior
    int T;      // chunks per CTA
    int NC;     // chunks per line
};
inline SegGeom seg_geometry(int nchunks)
{
    SegGeom s;
    s.NC = nchunks;
    if (nchunks <= SEG_T) {
        s.nseg = 1;
        s.iseg = s.T = nchunks;
        s.hlo = 0;
        return s;
    }
    const int imax = SEG_T - 2 * SEG_HALO;
    s.nseg = (nchunks + imax - 1) / imax;
    s.iseg = ((nchunks + s.nseg - 1) / s.nseg + 3) & ~3;
    s.hlo = SEG_HALO;
    s.T = SEG_T;
    return s;
}

// FAST schedule passes
struct FastCoefs {
    CompositeCoef D[3];   // derivative composite, per direction (depends on dx)
    CompositeCoef M;      // interpolation composite
};
bool fast_supported(int nx, int ny, int nz);
// xpass: f -> A = Dxx f, B = Mxx f
int fast_xpass(cudaStream_t s, const Brick &g, const FastCoefs &fc, const double *f, double *A,
               double *B, long long *launches);
// ypass: A,B -> C = Myy A + Dyy B, D = Myy B
int fast_ypass(cudaStream_t s, const Brick &g, const FastCoefs &fc, const double *A,
               const double *B, double *C, double *D, long long *launches);
// zpass: C,D -> out = Mzz C + Dzz D ; optional partial sums of p.out into dot_partials
int fast_zpass(cudaStream_t s, const Brick &g, const FastCoefs &fc, const double *C,
               const double *D, double *out, const double *p, double *dot_partials,
               int *n_partials, const ZOpen &zo, long long *launches);
int fast_zpass_max_partials(const Brick &g);
// FAST formulation of a single 1-D compact operator along dir (pbx_fast_lineop.cu)
// from_lo / from_up (dir == 2 only): the brick is one slab of a z-decomposed box and these are the
// neighbours' messages produced by fast_line_boundary ([3][nx*ny] each)
int fast_line_op(cudaStream_t s, const Brick &g, int dir, OpKind kind, int stagger, double dx,
                 const double *in, double *out, long long *launches, const double *from_lo = nullptr,
                 const double *from_up = nullptr, const double *addend = nullptr);   // out = op(in) + addend (y, z)
int fast_line_boundary(cudaStream_t s, const Brick &g, OpKind kind, int stagger, double dx,
                       const double *in, double *msg_dn, double *msg_up, long long *launches);
// TMA-pipelined persistent variants (pbx_fast_tma.cu); PBX_ERR_UNSUPPORTED = use the generic kernel
bool fast_tma_available();
int fast_line_op_tma(cudaStream_t s, const Brick &g, int dir, OpKind kind, int stagger, double dx,
                     const double *in, double *out, const double *addend, long long *launches,
                     double *out2 = nullptr);
int fast_line_op_sum_tma(cudaStream_t s, const Brick &g, int dir, OpKind kindA, OpKind kindB, int stagger,
                         double dx, const double *inA, const double *inB, double *out, long long *launches);
bool tma_make_map_2d(CUtensorMap_st *m, const double *base, unsigned long long dim0, unsigned long long dim1,
                     unsigned long long stride1_bytes, unsigned box0, unsigned box1, bool swizzle128);
// line-major batches by TMA tiles (pbx_tdma_tma.cu); PBX_ERR_UNSUPPORTED = use the generic kernel
int tdma_fwd_batch_lm(cudaStream_t s, int n, long long nl, long long es, long long ls, const double *a,
                      double *b, const double *c, double *d);
int tdma_bwd_batch_lm(cudaStream_t s, int n, long long nl, long long es, long long ls, const double *b,
                      const double *c, double *d);
int tdma_periodic_batch_lm(cudaStream_t s, int n, long long nl, long long es, long long ls, const double *a,
                           const double *b, const double *c, double *d, double *ws);
int fast_xpass_tma(cudaStream_t s, const Brick &g, const FastCoefs &fc, const double *f, double *A,
                   double *B, int rev, long long *launches);
// tail (z pass with a fused dot only): reduce the partial sums inside the kernel (RedTail); *tail_used
// tells whether the launched kernel took it
int fast_yzpass_tma(cudaStream_t s, const Brick &g, const FastCoefs &fc, int dir, const double *in0,
                    const double *in1, double *out0, double *out1, const double *pvec,
                    double *partials, const ZOpen &zo, int rev, long long *launches,
                    const RedTail *tail = nullptr, bool *tail_used = nullptr);

}  // namespace pbx

// ------------------------------------------------------------------------------------------------
// the handle
// ------------------------------------------------------------------------------------------------
struct pbx_handle_s {
    int nx = 0, ny = 0, nz = 0;
    double dx[3] = {0, 0, 0};
    int device = 0;
    int mode = PBX_MODE_FAST;
    int op = PBX_OPERATOR_COMPACT;   // what pbx_matmult_device and the CG apply
    cudaStream_t stream = nullptr;
    void *comm = nullptr;   // ncclComm_t
    int rank = 0, nranks = 1;
    long long launches = 0;

    pbx::FastCoefs fc;
    bool fast_ok = false;
    bool use_tma = true;     // PBX_NO_TMA=1 in the environment selects the generic kernels
    bool use_tma_yz = true;  // PBX_TMA_YZ=0: generic y/z passes, TMA x pass

    // REFERENCE-schedule tables: [dir][kind]
    pbx::RefLineTables ref[3][2];

    // scratch fields (device), allocated on first use
    double *scratch[10] = {nullptr};
    int nscratch = 0;

    // CG workspace (pbx_cg.cu)
    double *cg_r = nullptr, *cg_p = nullptr, *cg_w = nullptr;
    double *cg_partials = nullptr;   // per-block partial sums, 2 x cg_npartials
    int cg_npartials = 0;
    double *cg_scal = nullptr;       // device scalar block
    double *cg_host = nullptr;       // pinned host mirror of the scalar block
    double *cg_hist = nullptr;       // device residual history
    int cg_hist_cap = 0;
    unsigned *cg_ticket = nullptr;   // ticket counter of the reduction tails (zero between launches)
    const pbx::RedTail *pending_tail = nullptr;   // set by the CG around a MatMult: the z pass may take it

    // z-slab decomposition (pbx_dist.cu)
    void *dist = nullptr;

    // preconditioner of the CG (pbx_mg.cu): PBX_PC_NONE or PBX_PC_MG
    int pc = PBX_PC_NONE;
    void *mg = nullptr;
    double *cg_z = nullptr;          // preconditioned residual (allocated when a PC is set)
};

namespace pbx {
int ensure_scratch(pbx_handle_s *h, int count);
int lapl_reference(pbx_handle_s *h, const double *f, double *out);
int lapl_fast(pbx_handle_s *h, const double *f, double *out, const double *p, double *dot_dev);
int fast_pass(pbx_handle_s *h, int dir, const double *in0, const double *in1, double *out0,
              double *out1, const double *p, double *partials, const ZOpen *zo = nullptr,
              int rev = 0);
// grad / div / interp in the reference's stage order; fast = true uses the FAST line operators
int grad_stages_run(pbx_handle_s *h, const double *f, double *df, bool fast);
int div_stages_run(pbx_handle_s *h, const double *f, double *out, bool fast);
int interp_stages_run(pbx_handle_s *h, const double *f, double *fi, int stagger, bool fast);
// multigrid V-cycle on the star (pbx_mg.cu): z = M^-1 (r - *mean)
int mg_setup(pbx_handle_s *h, int nu);
int mg_vcycle(pbx_handle_s *h, const double *r, const double *mean, double *z);
void mg_free(pbx_handle_s *h);
// z = M^-1 (r - mean(r)) for the handle's preconditioner, mean removed from z (pbx_cg.cu)
int pc_apply(pbx_handle_s *h, const double *r, double *z);
// 2nd-order star (pbx_star.cu); lo / up: neighbour planes of a slab (nullptr: periodic in z)
int star_apply(pbx_handle_s *h, const double *x, double *y, const double *lo, const double *up);
// y = A x for the handle's operator, whatever the decomposition (pbx_api.cu)
int matmult(pbx_handle_s *h, const double *x, double *y);
int cg_solve(pbx_handle_s *h, const double *b, double *x, double rtol, double abstol, int maxit,
             int *its, double *rnorm, int *reason, double *hist, int nhist);
int cg_lapl_dot(pbx_handle_s *h, const double *f, double *out, double *dot_dev);
void cg_free(pbx_handle_s *h);
// z-slab decomposition over an NCCL communicator (or driven phase by phase by the caller)
int dist_setup(pbx_handle_s *h, int rank, int nranks);
int dist_phase1(pbx_handle_s *h, const double *f, int in_cg = 0);
int dist_phase2(pbx_handle_s *h, double *out, const double *p, double *partials);
int dist_attach(pbx_handle_s *h);
void dist_free(pbx_handle_s *h);
int dist_lapl(pbx_handle_s *h, const double *f, double *out, const double *p, double *partials);
// grad / div / interp on slabs (pbx_api.cu): message slots of the exchange buffers.  A slot holds
// the three planes a z line operator needs from each neighbour; DIST_MSG / 3 slots per exchange.
int dist_begin_epoch(pbx_handle_s *h);                                         // next exchange round
int dist_line_dst(pbx_handle_s *h, int slot, double **msg_dn, double **msg_up);   // where phase 1 writes
int dist_line_msgs(pbx_handle_s *h, int slot, const double **from_lo, const double **from_up);
int dist_exchange(pbx_handle_s *h);                                            // over the communicator
// sum `count` doubles in place over the handle's communicator (no-op for a single rank)
int dist_allreduce_sum(pbx_handle_s *h, double *dev, int count);
// multigrid on slabs (pbx_mg.cu): one plane each way, and an all-gather of `count` doubles per rank
int dist_halo_planes(pbx_handle_s *h, const double *field, size_t plane, int nz, const double **lo,
                     const double **hi);
int dist_halo_begin(pbx_handle_s *h, double **dn, double **up);
int dist_halo_end(pbx_handle_s *h, const double **lo, const double **hi);
int dist_allgather(pbx_handle_s *h, const double *mine, size_t count, double **full);
int mg_slab_plan(int nx, int ny, int nzl, int nranks, size_t *gather_doubles);
// true when the handle can talk to the other ranks (NCCL communicator or linked peer boards)
bool dist_connected(const pbx_handle_s *h);
// peer boards in place: fills *L and the sequence number of the next all-reduce, returns true
bool dist_peer_next(pbx_handle_s *h, PeerLinks *L, unsigned long long *seq);
void dist_peer_unget(pbx_handle_s *h, unsigned long long seq);
}  // namespace pbx

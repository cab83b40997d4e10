// pbx_api.cu -- the C ABI (include/pbx.h): handle lifecycle, operator drivers for both schedules,
// host-pointer convenience variants.
#include <algorithm>
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>

#include "pbx_internal.h"

namespace pbx {

static thread_local std::string g_last_error;

void set_last_error(const std::string &msg) { g_last_error = msg; }

int cuda_fail(cudaError_t e, const char *what, const char *file, int line)
{
    char buf[512];
    snprintf(buf, sizeof buf, "CUDA error %d (%s) in %s at %s:%d", (int)e, cudaGetErrorString(e),
             what, file, line);
    g_last_error = buf;
    return PBX_ERR_CUDA;
}

int ensure_scratch(pbx_handle_s *h, int count)
{
    const size_t bytes = sizeof(double) * (size_t)h->nx * h->ny * h->nz;
    while (h->nscratch < count) {
        if (h->nscratch >= 10) return PBX_ERR_NOMEM;
        PBX_CUDA(cudaMalloc(&h->scratch[h->nscratch], bytes));
        ++h->nscratch;
    }
    return PBX_OK;
}

// ------------------------------------------------------------------------------------------------
// REFERENCE schedule drivers: the reference's own stage order.
// ------------------------------------------------------------------------------------------------
namespace {

struct Lines {
    int n;
    long long nl1, nl2, es, ls1, ls2;
};

Lines lines_of(const pbx_handle_s *h, int dir)
{
    const long long nx = h->nx, ny = h->ny, nz = h->nz;
    switch (dir) {
    case 0: return {h->nx, ny, nz, 1, nx, nx * ny};        // x lines: (j,k)
    case 1: return {h->ny, nx, nz, nx, 1, nx * ny};        // y lines: (i,k)
    default: return {h->nz, nx, ny, nx * ny, 1, nx};       // z lines: (i,j)
    }
}

// zslot: on a slab of a z-decomposed box a z operator takes the neighbours' messages that
// line_boundary() put in slot `zslot` of the exchange buffers before the exchange
// addend (FAST only): out = op(in) + addend
int line_op(pbx_handle_s *h, int dir, OpKind kind, int stagger, const double *in, double *out,
            bool fast = false, int zslot = 0, const double *addend = nullptr)
{
    if (h->nranks > 1 && dir == 2) {
        const double *lo = nullptr, *up = nullptr;
        PBX_TRY(dist_line_msgs(h, zslot, &lo, &up));
        return fast_line_op(h->stream, Brick{h->nx, h->ny, h->nz}, dir, kind, stagger, h->dx[dir], in,
                            out, &h->launches, lo, up, addend);
    }
    if (fast)
        return fast_line_op(h->stream, Brick{h->nx, h->ny, h->nz}, dir, kind, stagger, h->dx[dir], in,
                            out, &h->launches, nullptr, nullptr, addend);
    if (addend) return PBX_ERR_ARG;
    Lines L = lines_of(h, dir);
    return ref_line_op(h->stream, L.n, L.nl1, L.nl2, L.es, L.ls1, L.ls2, kind, stagger, h->dx[dir],
                       h->ref[dir][kind], in, out, &h->launches);
}

// interpolation and derivative of the SAME field along dir (grad's z and y stages): one launch that
// reads the input once where the TMA line operators are in use, else two line operators
int line_op_pair(pbx_handle_s *h, int dir, int stagger, const double *in, double *out_interp,
                 double *out_deriv, bool fast, int zslot_interp, int zslot_deriv)
{
    if (fast && dir != 0 && !(h->nranks > 1 && dir == 2) && in != out_interp && in != out_deriv &&
        !((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out_interp) |
           reinterpret_cast<uintptr_t>(out_deriv)) & 15)) {
        const int rc = fast_line_op_tma(h->stream, Brick{h->nx, h->ny, h->nz}, dir, OP_INTERP, stagger, h->dx[dir],
                                        in, out_interp, nullptr, &h->launches, out_deriv);
        if (rc != PBX_ERR_UNSUPPORTED) return rc;
    }
    PBX_TRY(line_op(h, dir, OP_INTERP, stagger, in, out_interp, fast, zslot_interp));
    return line_op(h, dir, OP_DERIV, stagger, in, out_deriv, fast, zslot_deriv);
}

}  // namespace

// src/compact_schemes.f90:42-88 (Z -> Y -> X, backward stagger).  S = 5 scratch fields.
static int grad_stages(pbx_handle_s *h, const double *f, double *o1, double *o2, double *o3,
                       bool fast = false)
{
    PBX_TRY(ensure_scratch(h, 5));
    double **S = h->scratch;
    const int B = PBX_STAGGER_BACKWARD;
    PBX_TRY(line_op_pair(h, 2, B, f, S[0], S[1], fast, 0, 1));      // dff1 (= dff2, :63), dff3
    PBX_TRY(line_op_pair(h, 1, B, S[0], S[2], S[3], fast, 0, 0));   // dfe1, dfe2
    PBX_TRY(line_op(h, 1, OP_INTERP, B, S[1], S[4], fast));   // dfe3
    PBX_TRY(line_op(h, 0, OP_DERIV, B, S[2], o1, fast));      // df1
    PBX_TRY(line_op(h, 0, OP_INTERP, B, S[3], o2, fast));     // df2
    PBX_TRY(line_op(h, 0, OP_INTERP, B, S[4], o3, fast));     // df3
    return PBX_OK;
}

// src/compact_schemes.f90:207-257 (X -> Y -> Z, forward stagger).  i1..i3 may be scratch 0..2.
// div_stages_xy: the X and Y stages, leaving the inputs of the two Z operators in S[zi]
// (interpolation; zi = 4, or 2 on the FAST schedule, which folds the sums into the operators'
// stores) and S[3] (derivative); div_stages_z: the Z stage.
static int div_zi(bool fast) { return fast ? 2 : 4; }
static int div_stages_xy(pbx_handle_s *h, const double *i1, const double *i2, const double *i3,
                         bool fast = false)
{
    PBX_TRY(ensure_scratch(h, 5));
    double **S = h->scratch;
    const int F = PBX_STAGGER_FORWARD;
    const size_t N = (size_t)h->nx * h->ny * h->nz;
    // order chosen so that an input living in S[0..2] is consumed before its slot is reused
    PBX_TRY(line_op(h, 0, OP_DERIV, F, i1, S[3], fast));      // dfe1
    PBX_TRY(line_op(h, 0, OP_INTERP, F, i2, S[4], fast));     // dfe2
    double *e3 = S[0];                                  // i1 (possibly S[0]) is consumed by now
    PBX_TRY(line_op(h, 0, OP_INTERP, F, i3, e3, fast));       // dfe3
    if (fast) {
        // dff1 + dff2 (:249): in one launch where the TMA line operators are in use (the first summand stays
        // in registers), else the interpolation to S[1] and the derivative with S[1] as its addend
        int rc = fast_line_op_sum_tma(h->stream, Brick{h->nx, h->ny, h->nz}, 1, OP_INTERP, OP_DERIV, F, h->dx[1],
                                      S[3], S[4], S[2], &h->launches);
        if (rc == PBX_ERR_UNSUPPORTED) {
            PBX_TRY(line_op(h, 1, OP_INTERP, F, S[3], S[1], fast));            // dff1
            rc = line_op(h, 1, OP_DERIV, F, S[4], S[2], fast, 0, S[1]);
        }
        PBX_TRY(rc);
        return line_op(h, 1, OP_INTERP, F, e3, S[3], fast);              // dff3
    }
    PBX_TRY(line_op(h, 1, OP_INTERP, F, S[3], S[1], fast));   // dff1
    PBX_TRY(line_op(h, 1, OP_DERIV, F, S[4], S[2], fast));    // dff2
    PBX_TRY(line_op(h, 1, OP_INTERP, F, e3, S[3], fast));     // dff3
    return ref_add(h->stream, N, S[1], S[2], S[4], &h->launches);   // :249
}

static int div_stages_z(pbx_handle_s *h, double *out, bool fast = false)
{
    double **S = h->scratch;
    const int F = PBX_STAGGER_FORWARD;
    const size_t N = (size_t)h->nx * h->ny * h->nz;
    if (fast && h->nranks == 1) {   // dfc + df (:251) in one launch where the TMA line operators are in use
        const int rc = fast_line_op_sum_tma(h->stream, Brick{h->nx, h->ny, h->nz}, 2, OP_INTERP, OP_DERIV, F, h->dx[2],
                                            S[div_zi(fast)], S[3], out, &h->launches);
        if (rc != PBX_ERR_UNSUPPORTED) return rc;
    }
    PBX_TRY(line_op(h, 2, OP_INTERP, F, S[div_zi(fast)], S[0], fast, 0));   // dfc
    if (fast) return line_op(h, 2, OP_DERIV, F, S[3], out, fast, 1, S[0]);    // df + dfc (:251)
    // the z derivative goes to a scratch field first: the sum is a separate, reference-order step
    PBX_TRY(line_op(h, 2, OP_DERIV, F, S[3], S[1], fast, 1));    // df
    return ref_add(h->stream, N, S[1], S[0], out, &h->launches);    // :251
}

static int div_stages(pbx_handle_s *h, const double *i1, const double *i2, const double *i3,
                      double *out, bool fast = false)
{
    PBX_TRY(div_stages_xy(h, i1, i2, i3, fast));
    return div_stages_z(h, out, fast);
}

int lapl_reference(pbx_handle_s *h, const double *f, double *out)
{
    PBX_TRY(ensure_scratch(h, 5));
    double **S = h->scratch;
    PBX_TRY(grad_stages(h, f, S[0], S[1], S[2]));
    return div_stages(h, S[0], S[1], S[2], out);
}

int grad_stages_run(pbx_handle_s *h, const double *f, double *df, bool fast)
{
    const size_t N = (size_t)h->nx * h->ny * h->nz;
    return grad_stages(h, f, df, df + N, df + 2 * N, fast && h->fast_ok);
}

int div_stages_run(pbx_handle_s *h, const double *f, double *out, bool fast)
{
    const size_t N = (size_t)h->nx * h->ny * h->nz;
    return div_stages(h, f, f + N, f + 2 * N, out, fast && h->fast_ok);
}

// ---- grad / div / interp on one slab of a z-decomposed box: two phases around ONE exchange ---------
// phase 1 runs whatever precedes the Z stage and the boundary sweeps of the Z operators' inputs
// (three planes of nx*ny numbers per operator and neighbour); phase 2 the Z operators, which take
// the neighbours' messages as true stencil halos and recursion states, and whatever follows.
static int line_boundary(pbx_handle_s *h, int zslot, OpKind kind, int stagger, const double *in)
{
    double *dn = nullptr, *up = nullptr;
    PBX_TRY(dist_line_dst(h, zslot, &dn, &up));
    return fast_line_boundary(h->stream, Brick{h->nx, h->ny, h->nz}, kind, stagger, h->dx[2], in, dn, up,
                              &h->launches);
}

int slab_op_phase1(pbx_handle_s *h, int op, const double *in)
{
    const size_t N = (size_t)h->nx * h->ny * h->nz;
    PBX_TRY(ensure_scratch(h, 5));
    PBX_TRY(dist_begin_epoch(h));
    switch (op) {
    case PBX_OP_GRAD:
        PBX_TRY(line_boundary(h, 0, OP_INTERP, PBX_STAGGER_BACKWARD, in));
        return line_boundary(h, 1, OP_DERIV, PBX_STAGGER_BACKWARD, in);
    case PBX_OP_DIV:
        PBX_TRY(div_stages_xy(h, in, in + N, in + 2 * N, true));
        PBX_TRY(line_boundary(h, 0, OP_INTERP, PBX_STAGGER_FORWARD, h->scratch[div_zi(true)]));
        return line_boundary(h, 1, OP_DERIV, PBX_STAGGER_FORWARD, h->scratch[3]);
    case PBX_OP_INTERP:
        return line_boundary(h, 0, OP_INTERP, PBX_STAGGER_BACKWARD, in);
    case PBX_OP_INTERP_DIV:
        return line_boundary(h, 0, OP_INTERP, PBX_STAGGER_FORWARD, in);
    case PBX_OP_STAR: {
        // my bottom plane is the lower rank's plane above its top, my top plane the upper rank's
        // plane below its bottom
        double *dn = nullptr, *up = nullptr;
        PBX_TRY(dist_line_dst(h, 0, &dn, &up));
        const size_t plane = (size_t)h->nx * h->ny;
        PBX_CUDA(cudaMemcpyAsync(dn, in, plane * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
        PBX_CUDA(cudaMemcpyAsync(up, in + plane * (h->nz - 1), plane * sizeof(double),
                                 cudaMemcpyDeviceToDevice, h->stream));
        return PBX_OK;
    }
    default:
        return PBX_ERR_ARG;
    }
}

int slab_op_phase2(pbx_handle_s *h, int op, const double *in, double *out)
{
    const size_t N = (size_t)h->nx * h->ny * h->nz;
    switch (op) {
    case PBX_OP_GRAD:
        return grad_stages(h, in, out, out + N, out + 2 * N, true);
    case PBX_OP_DIV:
        return div_stages_z(h, out, true);
    case PBX_OP_INTERP:
        return interp_stages_run(h, in, out, PBX_STAGGER_BACKWARD, true);
    case PBX_OP_INTERP_DIV:
        return interp_stages_run(h, in, out, PBX_STAGGER_FORWARD, true);
    case PBX_OP_STAR: {
        const double *lo = nullptr, *up = nullptr;
        PBX_TRY(dist_line_msgs(h, 0, &lo, &up));
        return star_apply(h, in, out, lo, up);
    }
    default:
        return PBX_ERR_ARG;
    }
}

// src/compact_schemes.f90:93-142
int interp_stages_run(pbx_handle_s *h, const double *f, double *fi, int stagger, bool fast)
{
    PBX_TRY(ensure_scratch(h, 2));
    double **S = h->scratch;
    fast = fast && h->fast_ok;
    PBX_TRY(line_op(h, 2, OP_INTERP, stagger, f, S[0], fast));
    PBX_TRY(line_op(h, 1, OP_INTERP, stagger, S[0], S[1], fast));
    PBX_TRY(line_op(h, 0, OP_INTERP, stagger, S[1], fi, fast));
    return PBX_OK;
}

// ------------------------------------------------------------------------------------------------
// FAST schedule driver
// ------------------------------------------------------------------------------------------------
// one pass (0 = x, 1 = y, 2 = z): the TMA-pipelined kernel when the shape fits, else the generic one
int fast_pass(pbx_handle_s *h, int dir, const double *in0, const double *in1, double *out0,
              double *out1, const double *p, double *partials, const ZOpen *zop, int rev)
{
    Brick g{h->nx, h->ny, h->nz};
    const ZOpen zo = zop ? *zop : ZOpen();
    // measured on B200 at 512^3 (profiles/README.md): TMA-pipelined x / y / z passes run at
    // 90 / 85 / 76 % of the measured HBM peak against 45 / 74 / 66 % for the generic kernels.
    if (h->use_tma && (dir == 0 || h->use_tma_yz)) {
        bool used = false;
        int rc = dir == 0 ? fast_xpass_tma(h->stream, g, h->fc, in0, out0, out1, rev, &h->launches)
                          : fast_yzpass_tma(h->stream, g, h->fc, dir, in0, in1, out0, out1, p,
                                            partials, zo, rev, &h->launches,
                                            (dir == 2 && p && partials) ? h->pending_tail : nullptr, &used);
        if (used) h->pending_tail = nullptr;   // the kernel reduces its partial sums itself
        if (rc != PBX_ERR_UNSUPPORTED) return rc;
    }
    if (dir == 0) return fast_xpass(h->stream, g, h->fc, in0, out0, out1, &h->launches);
    if (dir == 1) return fast_ypass(h->stream, g, h->fc, in0, in1, out0, out1, &h->launches);
    return fast_zpass(h->stream, g, h->fc, in0, in1, out0, p, partials, nullptr, zo, &h->launches);
}

int lapl_fast(pbx_handle_s *h, const double *f, double *out, const double *p, double *partials)
{
    if (!h->fast_ok) {
        set_last_error("FAST schedule needs nx, ny, nz multiples of 16 (>= 16; nx <= 4096)");
        return PBX_ERR_UNSUPPORTED;
    }
    if ((reinterpret_cast<uintptr_t>(f) | reinterpret_cast<uintptr_t>(out)) & 15) {
        set_last_error("FAST schedule needs 16-byte aligned fields");
        return PBX_ERR_ARG;
    }
    // a y line longer than one CTA holds is cut into overlapping segments whose halos other CTAs
    // read: that y pass cannot run in place and writes to a second pair of scratch fields
    const bool yseg = seg_geometry(h->ny / LC).nseg > 1;
    PBX_TRY(ensure_scratch(h, yseg ? 4 : 2));
    double **S = h->scratch;
    double *C = yseg ? S[2] : S[0], *D = yseg ? S[3] : S[1];
    // Tile order and the 126 MB L2: a pass that starts where its producer has just finished finds
    // the last ~50 MB of each input still resident.  Inside the CG the input p was written front to
    // back by the p-update, so the x pass walks back to front and the y pass front to back; for a
    // stand-alone apply the x pass walks forward and the y pass backward.
    const int xrev = p ? 1 : 0;
    PBX_TRY(fast_pass(h, 0, f, nullptr, S[0], S[1], nullptr, nullptr, nullptr, xrev));
    // the y pass runs in place: a tile is read completely before any of it is written
    PBX_TRY(fast_pass(h, 1, S[0], S[1], C, D, nullptr, nullptr, nullptr, 1 - xrev));
    PBX_TRY(fast_pass(h, 2, C, D, out, nullptr, p, partials));
    return PBX_OK;
}

}  // namespace pbx

// ================================================================================================
// C ABI
// ================================================================================================
using namespace pbx;

extern "C" {

int pbx_version(void) { return PBX_VERSION; }

const char *pbx_error_string(int code)
{
    switch (code) {
    case PBX_OK: return "success";
    case PBX_ERR_ARG: return "invalid argument";
    case PBX_ERR_CUDA: return "CUDA error or no CUDA device";
    case PBX_ERR_NCCL: return "NCCL error";
    case PBX_ERR_UNSUPPORTED: return "unsupported configuration";
    case PBX_ERR_NOMEM: return "out of memory";
    case PBX_ERR_SIZE: return "array size mismatch";
    default: return "unknown error";
    }
}

const char *pbx_last_error(void) { return g_last_error.c_str(); }

int pbx_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int pbx_create(int nx, int ny, int nz, const double dx[3], int device, void *nccl_comm,
               pbx_handle *out)
{
    if (!out || !dx) return PBX_ERR_ARG;
    *out = nullptr;
    if (nx < 3 || ny < 3 || nz < 3 || !(dx[0] > 0) || !(dx[1] > 0) || !(dx[2] > 0)) {
        set_last_error("pbx_create: need nx, ny, nz >= 3 and positive spacings");
        return PBX_ERR_ARG;
    }
    if (pbx_device_count() <= 0) {
        set_last_error("pbx_create: no CUDA device (there is no CPU fallback)");
        return PBX_ERR_CUDA;
    }
    PBX_CUDA(cudaSetDevice(device));
    pbx_handle_s *h = new pbx_handle_s();
    h->nx = nx;
    h->ny = ny;
    h->nz = nz;
    for (int d = 0; d < 3; ++d) h->dx[d] = dx[d];
    h->device = device;
    h->comm = nccl_comm;
    const int nn[3] = {nx, ny, nz};
    for (int d = 0; d < 3; ++d) {
        make_composite_coef(OP_DERIV, dx[d], &h->fc.D[d]);
        for (int k = 0; k < 2; ++k) {
            int rc = make_ref_tables(nn[d], scheme_alpha((OpKind)k), &h->ref[d][k]);
            if (rc != PBX_OK) {
                pbx_destroy(h);
                return rc;
            }
        }
    }
    make_composite_coef(OP_INTERP, 1.0, &h->fc.M);
    h->fast_ok = fast_supported(nx, ny, nz);
    {
        const char *e = getenv("PBX_NO_TMA");
        h->use_tma = !(e && e[0] == '1');
        e = getenv("PBX_TMA_YZ");
        h->use_tma_yz = !(e && e[0] == '0');
    }
    h->mode = h->fast_ok ? PBX_MODE_FAST : PBX_MODE_REFERENCE;
    if (nccl_comm) {
        int rc = dist_attach(h);
        if (rc != PBX_OK) {
            pbx_destroy(h);
            return rc;
        }
    }
    *out = h;
    return PBX_OK;
}

int pbx_destroy(pbx_handle h)
{
    if (!h) return PBX_OK;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    for (int d = 0; d < 3; ++d)
        for (int k = 0; k < 2; ++k) free_ref_tables(&h->ref[d][k]);
    for (int i = 0; i < h->nscratch; ++i) cudaFree(h->scratch[i]);
    cg_free(h);
    mg_free(h);
    dist_free(h);
    delete h;
    return PBX_OK;
}

int pbx_set_mode(pbx_handle h, int mode)
{
    if (!h || (mode != PBX_MODE_FAST && mode != PBX_MODE_REFERENCE)) return PBX_ERR_ARG;
    if (mode == PBX_MODE_FAST && !h->fast_ok) {
        set_last_error("FAST schedule needs nx, ny, nz multiples of 16 (>= 16; nx <= 4096)");
        return PBX_ERR_UNSUPPORTED;
    }
    if (mode == PBX_MODE_REFERENCE && h->nranks > 1) {
        set_last_error("REFERENCE schedule is single-rank (as the reference itself is)");
        return PBX_ERR_UNSUPPORTED;
    }
    h->mode = mode;
    return PBX_OK;
}

int pbx_get_mode(pbx_handle h, int *mode)
{
    if (!h || !mode) return PBX_ERR_ARG;
    *mode = h->mode;
    return PBX_OK;
}

int pbx_set_stream(pbx_handle h, void *stream)
{
    if (!h) return PBX_ERR_ARG;
    h->stream = (cudaStream_t)stream;
    return PBX_OK;
}

int pbx_synchronize(pbx_handle h)
{
    if (!h) return PBX_ERR_ARG;
    PBX_CUDA(cudaStreamSynchronize(h->stream));
    return PBX_OK;
}

int pbx_get_dims(pbx_handle h, int *nx, int *ny, int *nz)
{
    if (!h) return PBX_ERR_ARG;
    if (nx) *nx = h->nx;
    if (ny) *ny = h->ny;
    if (nz) *nz = h->nz;
    return PBX_OK;
}

long long pbx_launch_count(pbx_handle h) { return h ? h->launches : 0; }

// ---- 3-D operators -----------------------------------------------------------------------------
int pbx_lapl_device(pbx_handle h, const double *f, double *d2f)
{
    if (!h || !f || !d2f || f == d2f) return PBX_ERR_ARG;
    PBX_CUDA(cudaSetDevice(h->device));
    if (h->nranks > 1) return dist_lapl(h, f, d2f, nullptr, nullptr);
    return h->mode == PBX_MODE_FAST ? lapl_fast(h, f, d2f, nullptr, nullptr)
                                    : lapl_reference(h, f, d2f);
}

int pbx_lapl_dot_device(pbx_handle h, const double *f, double *d2f, double *dot_dev)
{
    if (!h || !f || !d2f || !dot_dev || f == d2f) return PBX_ERR_ARG;
    PBX_CUDA(cudaSetDevice(h->device));
    return cg_lapl_dot(h, f, d2f, dot_dev);
}

int pbx_lapl_profile_device(pbx_handle h, const double *f, double *d2f, int reps, double ms[3])
{
    if (!h || !f || !d2f || !ms || reps < 1 || f == d2f) return PBX_ERR_ARG;
    if (!h->fast_ok || h->nranks > 1) return PBX_ERR_UNSUPPORTED;
    PBX_CUDA(cudaSetDevice(h->device));
    const bool yseg = seg_geometry(h->ny / LC).nseg > 1;
    PBX_TRY(ensure_scratch(h, yseg ? 4 : 2));
    double **S = h->scratch;
    double *C = yseg ? S[2] : S[0], *D = yseg ? S[3] : S[1];
    cudaEvent_t ev[4];
    for (auto &e : ev) PBX_CUDA(cudaEventCreate(&e));
    ms[0] = ms[1] = ms[2] = 0.0;
    int rc = PBX_OK;
    for (int r = 0; r < reps && rc == PBX_OK; ++r) {
        cudaEventRecord(ev[0], h->stream);
        rc = fast_pass(h, 0, f, nullptr, S[0], S[1], nullptr, nullptr);
        cudaEventRecord(ev[1], h->stream);
        if (rc == PBX_OK) rc = fast_pass(h, 1, S[0], S[1], C, D, nullptr, nullptr, nullptr, 1);
        cudaEventRecord(ev[2], h->stream);
        if (rc == PBX_OK) rc = fast_pass(h, 2, C, D, d2f, nullptr, nullptr, nullptr);
        cudaEventRecord(ev[3], h->stream);
        if (cudaEventSynchronize(ev[3]) != cudaSuccess) rc = PBX_ERR_CUDA;
        for (int k = 0; k < 3 && rc == PBX_OK; ++k) {
            float t = 0;
            cudaEventElapsedTime(&t, ev[k], ev[k + 1]);
            ms[k] += t;
        }
    }
    for (auto &e : ev) cudaEventDestroy(e);
    PBX_TRY(rc);
    for (int k = 0; k < 3; ++k) ms[k] /= reps;
    PBX_CUDA(cudaGetLastError());
    return PBX_OK;
}

// on a z-decomposed box: phase 1, the exchange over the communicator, phase 2
static int slab_op(pbx_handle h, int op, const double *in, double *out)
{
    if (!dist_connected(h)) {
        set_last_error("slab handle without a communicator: drive it with pbx_slab_op_phase1/2");
        return PBX_ERR_UNSUPPORTED;
    }
    if ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) {
        set_last_error("FAST schedule needs 16-byte aligned fields");
        return PBX_ERR_ARG;
    }
    PBX_TRY(slab_op_phase1(h, op, in));
    PBX_TRY(dist_exchange(h));
    return slab_op_phase2(h, op, in, out);
}

}  // extern "C"

namespace pbx {
int matmult(pbx_handle_s *h, const double *x, double *y)
{
    if (h->op == PBX_OPERATOR_STAR) {
        if (h->nranks > 1) return slab_op(h, PBX_OP_STAR, x, y);
        return star_apply(h, x, y, nullptr, nullptr);
    }
    if (h->nranks > 1) return dist_lapl(h, x, y, nullptr, nullptr);
    return h->mode == PBX_MODE_FAST ? lapl_fast(h, x, y, nullptr, nullptr) : lapl_reference(h, x, y);
}
}  // namespace pbx

extern "C" {

int pbx_set_operator(pbx_handle h, int op)
{
    if (!h || (op != PBX_OPERATOR_COMPACT && op != PBX_OPERATOR_STAR)) return PBX_ERR_ARG;
    h->op = op;
    return PBX_OK;
}

int pbx_get_operator(pbx_handle h, int *op)
{
    if (!h || !op) return PBX_ERR_ARG;
    *op = h->op;
    return PBX_OK;
}

int pbx_matmult_device(pbx_handle h, const double *x, double *y)
{
    if (!h || !x || !y || x == y) return PBX_ERR_ARG;
    PBX_CUDA(cudaSetDevice(h->device));
    return matmult(h, x, y);
}

int pbx_set_pc(pbx_handle h, int pc, int nu)
{
    if (!h || (pc != PBX_PC_NONE && pc != PBX_PC_MG) || nu < 0) return PBX_ERR_ARG;
    if (pc == PBX_PC_MG) {
        if (h->nranks > 1 && !dist_connected(h)) {
            set_last_error("multigrid on slabs needs a communicator or linked peer boards");
            return PBX_ERR_UNSUPPORTED;
        }
        PBX_CUDA(cudaSetDevice(h->device));
        PBX_TRY(mg_setup(h, nu == 0 ? 2 : nu));
    }
    h->pc = pc;
    return PBX_OK;
}

int pbx_pc_apply_device(pbx_handle h, const double *r, double *z)
{
    if (!h || !r || !z || r == z) return PBX_ERR_ARG;
    if ((reinterpret_cast<uintptr_t>(r) | reinterpret_cast<uintptr_t>(z)) & 15) {
        set_last_error("pbx_pc_apply_device: r and z must be 16-byte aligned");
        return PBX_ERR_ARG;
    }
    PBX_CUDA(cudaSetDevice(h->device));
    return pc_apply(h, r, z);
}

int pbx_star_device(pbx_handle h, const double *x, double *y)
{
    if (!h || !x || !y || x == y) return PBX_ERR_ARG;
    PBX_CUDA(cudaSetDevice(h->device));
    if (h->nranks > 1) return slab_op(h, PBX_OP_STAR, x, y);
    return star_apply(h, x, y, nullptr, nullptr);
}

int pbx_grad_device(pbx_handle h, const double *f, double *df)
{
    if (!h || !f || !df) return PBX_ERR_ARG;
    PBX_CUDA(cudaSetDevice(h->device));
    if (h->nranks > 1) return slab_op(h, PBX_OP_GRAD, f, df);
    return grad_stages_run(h, f, df, h->mode == PBX_MODE_FAST);
}

int pbx_div_device(pbx_handle h, const double *f, double *df)
{
    if (!h || !f || !df) return PBX_ERR_ARG;
    PBX_CUDA(cudaSetDevice(h->device));
    if (h->nranks > 1) return slab_op(h, PBX_OP_DIV, f, df);
    return div_stages_run(h, f, df, h->mode == PBX_MODE_FAST);
}

int pbx_interp_device(pbx_handle h, const double *f, double *fi, int stagger)
{
    if (!h || !f || !fi || (stagger != -1 && stagger != 1)) return PBX_ERR_ARG;
    PBX_CUDA(cudaSetDevice(h->device));
    if (h->nranks > 1)
        return slab_op(h, stagger == PBX_STAGGER_BACKWARD ? PBX_OP_INTERP : PBX_OP_INTERP_DIV, f, fi);
    return interp_stages_run(h, f, fi, stagger, h->mode == PBX_MODE_FAST);
}

int pbx_slab_op_phase1(pbx_handle h, int op, const double *in)
{
    if (!h || !in || !h->dist) return PBX_ERR_ARG;
    PBX_CUDA(cudaSetDevice(h->device));
    return slab_op_phase1(h, op, in);
}

int pbx_slab_op_phase2(pbx_handle h, int op, const double *in, double *out)
{
    if (!h || !in || !out || !h->dist) return PBX_ERR_ARG;
    PBX_CUDA(cudaSetDevice(h->device));
    return slab_op_phase2(h, op, in, out);
}

// ---- batched 1-D operators ---------------------------------------------------------------------
namespace {

std::mutex g_tab_mutex;
std::map<std::tuple<int, int, int>, RefLineTables> g_tabs;   // (device, n, kind)

int cached_tables(int n, OpKind kind, const RefLineTables **out)
{
    int dev = 0;
    PBX_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_tab_mutex);
    auto key = std::make_tuple(dev, n, (int)kind);
    auto it = g_tabs.find(key);
    if (it == g_tabs.end()) {
        RefLineTables t;
        PBX_TRY(make_ref_tables(n, scheme_alpha(kind), &t));
        it = g_tabs.emplace(key, t).first;
    }
    *out = &it->second;
    return PBX_OK;
}

int line_batch(int n, long long nl, long long es, long long ls, OpKind kind, const double *f,
               double dx, double *o, int stagger, void *stream)
{
    if (!f || !o || nl < 0 || (stagger != -1 && stagger != 1)) return PBX_ERR_ARG;
    if (n < 3) {
        set_last_error("compact line operators need n >= 3");
        return PBX_ERR_ARG;
    }
    if (pbx_device_count() <= 0) {
        set_last_error("no CUDA device (there is no CPU fallback)");
        return PBX_ERR_CUDA;
    }
    if (nl == 0) return PBX_OK;
    const RefLineTables *t = nullptr;
    PBX_TRY(cached_tables(n, kind, &t));
    return ref_line_op((cudaStream_t)stream, n, nl, 1, es, ls, 0, kind, stagger, dx, *t, f, o,
                       nullptr);
}

}  // namespace

int pbx_grad_1d_batch_device(int n, long long nlines, long long es, long long ls, const double *f,
                             double dx, double *df, int stagger, void *stream)
{
    if (!(dx > 0)) return PBX_ERR_ARG;
    return line_batch(n, nlines, es, ls, OP_DERIV, f, dx, df, stagger, stream);
}

int pbx_interp_1d_batch_device(int n, long long nlines, long long es, long long ls,
                               const double *f, double *fi, int stagger, void *stream)
{
    return line_batch(n, nlines, es, ls, OP_INTERP, f, 1.0, fi, stagger, stream);
}

// ---- batched tridiagonal solves ----------------------------------------------------------------
#define PBX_NEED_DEVICE()                                                   \
    do {                                                                    \
        if (pbx_device_count() <= 0) {                                      \
            set_last_error("no CUDA device (there is no CPU fallback)");    \
            return PBX_ERR_CUDA;                                            \
        }                                                                   \
    } while (0)

int pbx_tdma_batch_device(int n, long long nl, long long es, long long ls, const double *a,
                          double *b, const double *c, double *d, void *stream)
{
    if (!a || !b || !c || !d) return PBX_ERR_ARG;
    PBX_NEED_DEVICE();
    PBX_TRY(tdma_fwd_batch((cudaStream_t)stream, n, nl, es, ls, a, b, c, d));
    return tdma_bwd_batch((cudaStream_t)stream, n, nl, es, ls, b, c, d);
}

int pbx_tdma_periodic_batch_device(int n, long long nl, long long es, long long ls,
                                   const double *a, const double *b, const double *c, double *d,
                                   void *stream)
{
    if (!a || !b || !c || !d) return PBX_ERR_ARG;
    PBX_NEED_DEVICE();
    return tdma_periodic_batch((cudaStream_t)stream, n, nl, es, ls, a, b, c, d);
}

int pbx_fwd_sweep_batch_device(int n, long long nl, long long es, long long ls, const double *a,
                               double *b, const double *c, double *d, void *stream)
{
    if (!a || !b || !c || !d) return PBX_ERR_ARG;
    PBX_NEED_DEVICE();
    return tdma_fwd_batch((cudaStream_t)stream, n, nl, es, ls, a, b, c, d);
}

int pbx_bwd_sweep_batch_device(int n, long long nl, long long es, long long ls, const double *b,
                               const double *c, double *d, void *stream)
{
    if (!b || !c || !d) return PBX_ERR_ARG;
    PBX_NEED_DEVICE();
    return tdma_bwd_batch((cudaStream_t)stream, n, nl, es, ls, b, c, d);
}

// ---- CG ----------------------------------------------------------------------------------------
int pbx_cg_solve_device(pbx_handle h, const double *b, double *x, double rtol, double abstol,
                        int maxit, int *its, double *rnorm, int *reason, double *hist, int nhist)
{
    if (!h || !b || !x || maxit < 0) return PBX_ERR_ARG;
    PBX_CUDA(cudaSetDevice(h->device));
    return cg_solve(h, b, x, rtol, abstol, maxit, its, rnorm, reason, hist, nhist);
}

static const char *reason_name(int r)
{
    switch (r) {
    case PBX_CONVERGED_RTOL: return "CONVERGED_RTOL";
    case PBX_CONVERGED_ATOL: return "CONVERGED_ATOL";
    case PBX_DIVERGED_ITS: return "DIVERGED_ITS";
    case PBX_DIVERGED_DTOL: return "DIVERGED_DTOL";
    case PBX_DIVERGED_INDEFINITE_MAT: return "DIVERGED_INDEFINITE_MAT";
    case PBX_DIVERGED_INDEFINITE_PC: return "DIVERGED_INDEFINITE_PC";
    case PBX_DIVERGED_NANORINF: return "DIVERGED_NANORINF";
    default: return "UNKNOWN";
    }
}

// PETSc option names of the reference's solve (README.md:43-49, src/poissbox.f90:295)
int pbx_ksp_solve_device(pbx_handle h, const char *options, const double *b, double *x, int *its,
                         double *rnorm, int *reason)
{
    if (!h || !b || !x) return PBX_ERR_ARG;
    double rtol = 1e-5, atol = 1e-50;
    int maxit = 10000, nu = 0, pc = -1;
    bool monitor = false, why = false;
    std::vector<std::string> tok;
    if (options) {
        std::string cur;
        for (const char *c = options;; ++c) {
            if (*c == 0 || *c == ' ' || *c == '\t' || *c == '\n') {
                if (!cur.empty()) tok.push_back(cur);
                cur.clear();
                if (*c == 0) break;
            } else {
                cur += *c;
            }
        }
    }
    auto value = [&](size_t i, std::string *v) {
        if (i + 1 >= tok.size() || (tok[i + 1].size() > 1 && tok[i + 1][0] == '-' && !isdigit((unsigned char)tok[i + 1][1]) && tok[i + 1][1] != '.')) {
            set_last_error("option " + tok[i] + " needs a value");
            return false;
        }
        *v = tok[i + 1];
        return true;
    };
    for (size_t i = 0; i < tok.size(); ++i) {
        const std::string &o = tok[i];
        std::string v;
        char *end = nullptr;
        if (o == "-ksp_monitor") {
            monitor = true;
        } else if (o == "-ksp_converged_reason") {
            why = true;
        } else if (o == "-ksp_type") {
            if (!value(i, &v)) return PBX_ERR_ARG;
            if (v != "cg") {
                set_last_error("-ksp_type " + v + ": only cg is built");
                return PBX_ERR_UNSUPPORTED;
            }
            ++i;
        } else if (o == "-pc_type") {
            if (!value(i, &v)) return PBX_ERR_ARG;
            if (v == "none")
                pc = PBX_PC_NONE;
            else if (v == "mg" || v == "gamg")
                pc = PBX_PC_MG;
            else {
                set_last_error("-pc_type " + v + ": none, mg (and gamg as its stand-in) are built");
                return PBX_ERR_UNSUPPORTED;
            }
            ++i;
        } else if (o == "-ksp_rtol" || o == "-ksp_atol") {
            if (!value(i, &v)) return PBX_ERR_ARG;
            const double d = strtod(v.c_str(), &end);
            if (end == v.c_str() || *end || !(d >= 0)) {
                set_last_error(o + " " + v + ": not a tolerance");
                return PBX_ERR_ARG;
            }
            (o == "-ksp_rtol" ? rtol : atol) = d;
            ++i;
        } else if (o == "-ksp_max_it" || o == "-pc_mg_smoothup" || o == "-pc_mg_smoothdown" ||
                   o == "-mg_levels_ksp_max_it") {
            if (!value(i, &v)) return PBX_ERR_ARG;
            const long n = strtol(v.c_str(), &end, 10);
            if (end == v.c_str() || *end || n < 0 || n > 1000000000L) {
                set_last_error(o + " " + v + ": not a count");
                return PBX_ERR_ARG;
            }
            if (o == "-ksp_max_it")
                maxit = (int)n;
            else
                nu = (int)n;
            ++i;
        }
    }
    if (pc >= 0) PBX_TRY(pbx_set_pc(h, pc, pc == PBX_PC_MG ? nu : 0));
    else if (nu > 0 && h->pc == PBX_PC_MG) PBX_TRY(pbx_set_pc(h, PBX_PC_MG, nu));
    // -ksp_monitor keeps at most the first 2^22 norms (32 MB); a solve that runs longer stops printing
    const size_t hist_cap = (size_t)1 << 22;
    std::vector<double> hist(monitor ? std::min((size_t)maxit + 1, hist_cap) : 0);
    int k = 0, r = 0;
    double rn = 0.0;
    PBX_TRY(pbx_cg_solve_device(h, b, x, rtol, atol, maxit, &k, &rn, &r, monitor ? hist.data() : nullptr,
                                (int)hist.size()));
    if (monitor && h->rank == 0)
        for (int i = 0; i <= k && i < (int)hist.size(); ++i) printf("%3d KSP Residual norm %14.12e\n", i, hist[i]);
    if (why && h->rank == 0)
        printf("Linear solve %s due to %s iterations %d\n", r > 0 ? "converged" : "did not converge", reason_name(r), k);
    if (monitor || why) fflush(stdout);
    if (its) *its = k;
    if (rnorm) *rnorm = rn;
    if (reason) *reason = r;
    return PBX_OK;
}

}  // extern "C"

// ---- host-pointer convenience variants ---------------------------------------------------------
namespace {

struct HostEntry {
    pbx_handle h = nullptr;
    double *din = nullptr, *dout = nullptr;   // 3 fields each
    size_t N = 0;
};
std::mutex g_host_mutex;
std::vector<HostEntry> g_host;
int g_host_mode = PBX_MODE_REFERENCE;   // schedule of the grad / div / interp host variants

int host_entry(int nx, int ny, int nz, const double dx[3], HostEntry **out)
{
    int dev = 0;
    if (pbx_device_count() <= 0) {
        set_last_error("no CUDA device (there is no CPU fallback)");
        return PBX_ERR_CUDA;
    }
    PBX_CUDA(cudaGetDevice(&dev));
    for (auto &e : g_host) {
        pbx_handle_s *h = e.h;
        if (h->nx == nx && h->ny == ny && h->nz == nz && h->device == dev && h->dx[0] == dx[0] &&
            h->dx[1] == dx[1] && h->dx[2] == dx[2]) {
            *out = &e;
            return PBX_OK;
        }
    }
    if (g_host.size() >= 4) {   // small LRU-less cache: drop the oldest
        HostEntry &o = g_host.front();
        cudaFree(o.din);
        cudaFree(o.dout);
        pbx_destroy(o.h);
        g_host.erase(g_host.begin());
    }
    HostEntry e;
    PBX_TRY(pbx_create(nx, ny, nz, dx, dev, nullptr, &e.h));
    e.N = (size_t)nx * ny * nz;
    if (cudaMalloc(&e.din, 3 * e.N * sizeof(double)) != cudaSuccess ||
        cudaMalloc(&e.dout, 3 * e.N * sizeof(double)) != cudaSuccess) {
        cudaGetLastError();
        cudaFree(e.din);
        pbx_destroy(e.h);
        set_last_error("host staging allocation failed");
        return PBX_ERR_NOMEM;
    }
    g_host.push_back(e);
    *out = &g_host.back();
    return PBX_OK;
}

template <class Fn>
int host_run(int nx, int ny, int nz, const double dx[3], const double *in, int nin, double *out,
             int nout, Fn fn)
{
    if (!in || !out || !dx) return PBX_ERR_ARG;
    std::lock_guard<std::mutex> lk(g_host_mutex);
    HostEntry *e = nullptr;
    PBX_TRY(host_entry(nx, ny, nz, dx, &e));
    cudaStream_t s = e->h->stream;
    PBX_CUDA(cudaMemcpyAsync(e->din, in, nin * e->N * sizeof(double), cudaMemcpyHostToDevice, s));
    PBX_TRY(fn(e));
    PBX_CUDA(cudaMemcpyAsync(out, e->dout, nout * e->N * sizeof(double), cudaMemcpyDeviceToHost, s));
    PBX_CUDA(cudaStreamSynchronize(s));
    return PBX_OK;
}

// single-line / small-batch host staging
struct DevBuf {
    double *p = nullptr;
    ~DevBuf()
    {
        if (p) cudaFree(p);
    }
    int alloc(size_t n)
    {
        PBX_CUDA(cudaMalloc(&p, n * sizeof(double)));
        return PBX_OK;
    }
};

}  // namespace

extern "C" {

int pbx_lapl_host(int nx, int ny, int nz, const double *f, const double dx[3], double *d2f,
                  int mode)
{
    return host_run(nx, ny, nz, dx, f, 1, d2f, 1, [&](HostEntry *e) {
        int m = mode;
        if (m == PBX_MODE_FAST && !e->h->fast_ok) m = PBX_MODE_REFERENCE;
        PBX_TRY(pbx_set_mode(e->h, m));
        return pbx_lapl_device(e->h, e->din, e->dout);
    });
}

// `count` independent fields through one handle, double-buffered: the copy-in of field k + 1 and
// the copy-out of field k - 1 overlap the compute of field k on three streams, so that a batch
// moves at the rate of ONE direction of the host link instead of the sum of both.  The host buffers
// should be pinned (pageable memory makes the copies synchronous and the overlap disappears).
int pbx_lapl_host_batch(int nx, int ny, int nz, int count, const double *const *f, const double dx[3],
                        double *const *d2f, int mode)
{
    if (!f || !d2f || !dx || count < 1) return PBX_ERR_ARG;
    for (int k = 0; k < count; ++k)
        if (!f[k] || !d2f[k]) return PBX_ERR_ARG;
    std::lock_guard<std::mutex> lk(g_host_mutex);
    HostEntry *e = nullptr;
    PBX_TRY(host_entry(nx, ny, nz, dx, &e));
    int m = mode;
    if (m == PBX_MODE_FAST && !e->h->fast_ok) m = PBX_MODE_REFERENCE;
    PBX_TRY(pbx_set_mode(e->h, m));
    const size_t bytes = e->N * sizeof(double);
    cudaStream_t st[3] = {nullptr, nullptr, nullptr};   // copy in, compute, copy out
    cudaEvent_t in_done[2] = {nullptr, nullptr}, comp_done[2] = {nullptr, nullptr}, out_done[2] = {nullptr, nullptr};
    cudaStream_t saved = e->h->stream;
    int rc = PBX_OK;
    auto ok = [&](cudaError_t ce) {
        if (ce != cudaSuccess && rc == PBX_OK) rc = cuda_fail(ce, "pbx_lapl_host_batch", __FILE__, __LINE__);
        return ce == cudaSuccess;
    };
    for (int i = 0; i < 3 && rc == PBX_OK; ++i) ok(cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking));
    for (int i = 0; i < 2 && rc == PBX_OK; ++i) {
        ok(cudaEventCreateWithFlags(&in_done[i], cudaEventDisableTiming));
        ok(cudaEventCreateWithFlags(&comp_done[i], cudaEventDisableTiming));
        ok(cudaEventCreateWithFlags(&out_done[i], cudaEventDisableTiming));
    }
    if (rc == PBX_OK) {
        // the handle's earlier work (on its own stream) is complete before the batch starts
        ok(cudaStreamSynchronize(saved));
        e->h->stream = st[1];
        for (int k = 0; k < count && rc == PBX_OK; ++k) {
            const int slot = k & 1;
            double *din = e->din + (size_t)slot * e->N, *dout = e->dout + (size_t)slot * e->N;
            // copy in: the compute of field k - 2 has read this input slot
            if (k >= 2) ok(cudaStreamWaitEvent(st[0], comp_done[slot], 0));
            ok(cudaMemcpyAsync(din, f[k], bytes, cudaMemcpyHostToDevice, st[0]));
            ok(cudaEventRecord(in_done[slot], st[0]));
            // compute: the input has arrived, and the copy-out of field k - 2 has read this output slot
            ok(cudaStreamWaitEvent(st[1], in_done[slot], 0));
            if (k >= 2) ok(cudaStreamWaitEvent(st[1], out_done[slot], 0));
            if (rc == PBX_OK) rc = pbx_lapl_device(e->h, din, dout);
            ok(cudaEventRecord(comp_done[slot], st[1]));
            // copy out
            ok(cudaStreamWaitEvent(st[2], comp_done[slot], 0));
            ok(cudaMemcpyAsync(d2f[k], dout, bytes, cudaMemcpyDeviceToHost, st[2]));
            ok(cudaEventRecord(out_done[slot], st[2]));
        }
        for (int i = 0; i < 3; ++i) ok(cudaStreamSynchronize(st[i]));
        e->h->stream = saved;
    }
    for (int i = 0; i < 2; ++i) {
        if (in_done[i]) cudaEventDestroy(in_done[i]);
        if (comp_done[i]) cudaEventDestroy(comp_done[i]);
        if (out_done[i]) cudaEventDestroy(out_done[i]);
    }
    for (int i = 0; i < 3; ++i)
        if (st[i]) cudaStreamDestroy(st[i]);
    return rc;
}

int pbx_star_host(int nx, int ny, int nz, const double *x, const double dx[3], double *y)
{
    return host_run(nx, ny, nz, dx, x, 1, y, 1,
                    [&](HostEntry *e) { return pbx_star_device(e->h, e->din, e->dout); });
}

static int host_mode_for(HostEntry *e)
{
    int m = g_host_mode;
    if (m == PBX_MODE_FAST && !e->h->fast_ok) m = PBX_MODE_REFERENCE;
    return pbx_set_mode(e->h, m);
}

int pbx_host_set_mode(int mode)
{
    if (mode != PBX_MODE_FAST && mode != PBX_MODE_REFERENCE) return PBX_ERR_ARG;
    std::lock_guard<std::mutex> lk(g_host_mutex);
    g_host_mode = mode;
    return PBX_OK;
}

int pbx_grad_host(int nx, int ny, int nz, const double *f, const double dx[3], double *df)
{
    return host_run(nx, ny, nz, dx, f, 1, df, 3, [&](HostEntry *e) {
        PBX_TRY(host_mode_for(e));
        return pbx_grad_device(e->h, e->din, e->dout);
    });
}

int pbx_div_host(int nx, int ny, int nz, const double *f, const double dx[3], double *df)
{
    return host_run(nx, ny, nz, dx, f, 3, df, 1, [&](HostEntry *e) {
        PBX_TRY(host_mode_for(e));
        return pbx_div_device(e->h, e->din, e->dout);
    });
}

int pbx_interp_host(int nx, int ny, int nz, const double *f, double *fi, int stagger)
{
    const double dx[3] = {1.0, 1.0, 1.0};
    return host_run(nx, ny, nz, dx, f, 1, fi, 1, [&](HostEntry *e) {
        PBX_TRY(host_mode_for(e));
        return pbx_interp_device(e->h, e->din, e->dout, stagger);
    });
}

static int line_host(int nf, const double *f, double dx, int nout, double *o, int stagger,
                     bool deriv)
{
    if (!f || !o) return PBX_ERR_ARG;
    if (nf != nout) {   // src/compact_schemes.f90:177-180, 292-295
        set_last_error("ERROR: periodic gradient is same length as field!");
        return PBX_ERR_SIZE;
    }
    PBX_NEED_DEVICE();
    DevBuf in, out;
    PBX_TRY(in.alloc(nf));
    PBX_TRY(out.alloc(nf));
    PBX_CUDA(cudaMemcpy(in.p, f, nf * sizeof(double), cudaMemcpyHostToDevice));
    if (deriv)
        PBX_TRY(pbx_grad_1d_batch_device(nf, 1, 1, nf, in.p, dx, out.p, stagger, nullptr));
    else
        PBX_TRY(pbx_interp_1d_batch_device(nf, 1, 1, nf, in.p, out.p, stagger, nullptr));
    PBX_CUDA(cudaMemcpy(o, out.p, nf * sizeof(double), cudaMemcpyDeviceToHost));
    return PBX_OK;
}

int pbx_grad_1d_host(int nf, const double *f, double dx, int ndf, double *df, int stagger)
{
    return line_host(nf, f, dx, ndf, df, stagger, true);
}

int pbx_interp_1d_host(int nf, const double *f, int nfi, double *fi, int stagger)
{
    return line_host(nf, f, 1.0, nfi, fi, stagger, false);
}

namespace {
// which: 0 tdma, 1 periodic, 2 fwd, 3 bwd
int tri_host(int which, int n, const double *a, double *b, const double *c, double *d)
{
    if (!b || !c || !d || (which != 3 && !a) || n < 1) return PBX_ERR_ARG;
    PBX_NEED_DEVICE();
    DevBuf buf;
    PBX_TRY(buf.alloc(4 * (size_t)n));
    double *da = buf.p, *db = da + n, *dc = db + n, *dd = dc + n;
    const size_t by = n * sizeof(double);
    if (a) PBX_CUDA(cudaMemcpy(da, a, by, cudaMemcpyHostToDevice));
    PBX_CUDA(cudaMemcpy(db, b, by, cudaMemcpyHostToDevice));
    PBX_CUDA(cudaMemcpy(dc, c, by, cudaMemcpyHostToDevice));
    PBX_CUDA(cudaMemcpy(dd, d, by, cudaMemcpyHostToDevice));
    int rc = PBX_OK;
    switch (which) {
    case 0: rc = pbx_tdma_batch_device(n, 1, 1, n, da, db, dc, dd, nullptr); break;
    case 1: rc = pbx_tdma_periodic_batch_device(n, 1, 1, n, da, db, dc, dd, nullptr); break;
    case 2: rc = pbx_fwd_sweep_batch_device(n, 1, 1, n, da, db, dc, dd, nullptr); break;
    default: rc = pbx_bwd_sweep_batch_device(n, 1, 1, n, db, dc, dd, nullptr); break;
    }
    PBX_TRY(rc);
    PBX_CUDA(cudaMemcpy(d, dd, by, cudaMemcpyDeviceToHost));
    if (which == 0 || which == 2) PBX_CUDA(cudaMemcpy(b, db, by, cudaMemcpyDeviceToHost));
    return PBX_OK;
}
}  // namespace

int pbx_tdma_host(int n, const double *a, double *b, const double *c, double *d)
{
    return tri_host(0, n, a, b, c, d);
}
int pbx_tdma_periodic_host(int n, const double *a, const double *b, const double *c, double *d)
{
    return tri_host(1, n, a, const_cast<double *>(b), c, d);
}
int pbx_fwd_sweep_host(int n, const double *a, double *b, const double *c, double *d)
{
    return tri_host(2, n, a, b, c, d);
}
int pbx_bwd_sweep_host(int n, const double *b, const double *c, double *d)
{
    return tri_host(3, n, nullptr, const_cast<double *>(b), c, d);
}

int pbx_cg_solve_host(int nx, int ny, int nz, const double dx[3], const double *b, double *x,
                      double rtol, double abstol, int maxit, int mode, int *its, double *rnorm,
                      int *reason, double *hist, int nhist)
{
    return host_run(nx, ny, nz, dx, b, 1, x, 1, [&](HostEntry *e) {
        int m = mode;
        if (m == PBX_MODE_FAST && !e->h->fast_ok) m = PBX_MODE_REFERENCE;
        PBX_TRY(pbx_set_mode(e->h, m));
        return pbx_cg_solve_device(e->h, e->din, e->dout, rtol, abstol, maxit, its, rnorm, reason,
                                   hist, nhist);
    });
}

int pbx_host_cache_clear(void)
{
    std::lock_guard<std::mutex> lk(g_host_mutex);
    for (auto &e : g_host) {
        cudaFree(e.din);
        cudaFree(e.dout);
        pbx_destroy(e.h);
    }
    g_host.clear();
    tdma_trim_workspace();
    return PBX_OK;
}

}  // extern "C"

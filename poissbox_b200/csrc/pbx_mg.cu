// pbx_mg.cu -- geometric multigrid V-cycle on the 2nd-order 7-point star, as the preconditioner
// of the conjugate-gradient solve.
//
// What it stands in for: the reference hands KSP two matrices, the shell operator A and the
// assembled 2nd-order star P (KSPSetOperators(ksp, A, P), src/poissbox.f90:294; P from
// assemble_laplacian, src/coefficients.f90:52-123), and its README runs the solve with
// `-pc_type gamg`, i.e. a multigrid preconditioner built on P.  PETSc's GAMG is third-party
// (algebraic, version unpinned, absent here): this is NOT a restatement of it -- parity is
// unpinned by construction -- but the same role filled with a geometric V-cycle on the same P:
//   S = -P (symmetric positive semi-definite, constant null space), periodic, cell-centred;
//   smoother: nu damped-Jacobi sweeps before and after (omega = 6/7, the optimum for the 3-D star);
//   transfer: cell-centred trilinear prolongation Pr (weights 3/4, 1/4 per direction) and
//             restriction R = Pr^T / 8 (weights 1/8, 3/8, 3/8, 1/8), so the cycle is a symmetric
//             positive definite operator on the zero-mean subspace, as CG requires;
//   coarse operators by re-discretisation (spacing doubles); the coarsest grid (a side <= 4 or
//   odd) gets a fixed number of Jacobi sweeps from a zero guess (again a symmetric polynomial in S).
// The CG removes the constant from the result (MatNullSpace semantics); the mean of the right-hand
// side is removed on entry (it is the invariant mean of b).
//
// Every kernel is a bandwidth-bound stencil / transfer sweep in the structure of pbx_star.cu
// (a thread owns an (i,j) column of a block of planes, z neighbours in registers, in-plane
// neighbours out of L1).  Per V(2,2) cycle the finest level moves about 110 B/DoF, all levels 8/7
// of that.
#include <vector>

#include "pbx_internal.h"

namespace pbx {

namespace {

constexpr int MBX = 32, MBY = 8, MKZ = 4;
constexpr double MG_OMEGA = 6.0 / 7.0;
constexpr int MG_COARSE_SWEEPS = 30;

struct Lv {
    int nx, ny, nz;
    double cx, cy, cz;   // 1 / h^2 per direction
    double wd;           // omega / diag(S)
};

// MODE 0: out = z + wd ((r - m) - S z)      damped Jacobi
// MODE 1: out = (r - m) - S z               residual
// MODE 2: the first TWO Jacobi sweeps from a zero guess in one pass over r (z = r on entry):
//         z1 = wd (r - m),  out = z1 + wd ((r - m) - S z1) = wd ((r - m) + ((r - m) - wd S r))
// A thread owns MKZ consecutive planes of one (i,j) column; all its loads (MKZ + 2 values of its own
// column, 4 MKZ in-plane neighbours, MKZ right-hand sides) are issued before the arithmetic, so
// that a warp keeps ~30 requests in flight instead of walking a dependent chain of planes.
template <int MODE>
__global__ void __launch_bounds__(MBX * MBY)
mg_sweep_kernel(const __grid_constant__ Lv lv, const double *__restrict__ z,
                const double *__restrict__ r, const double *__restrict__ mean,
                double *__restrict__ out)
{
    const int nx = lv.nx, ny = lv.ny, nz = lv.nz;
    const int i = blockIdx.x * MBX + threadIdx.x, j = blockIdx.y * MBY + threadIdx.y;
    if (i >= nx || j >= ny) return;
    const double m = mean ? *mean : 0.0;
    const int k0 = blockIdx.z * MKZ;
    const size_t plane = (size_t)nx * ny;
    const size_t col = i + (size_t)nx * j;
    const size_t im = (i == 0 ? nx - 1 : i - 1) + (size_t)nx * j, ip = (i == nx - 1 ? 0 : i + 1) + (size_t)nx * j;
    const size_t jm = i + (size_t)nx * (j == 0 ? ny - 1 : j - 1), jp = i + (size_t)nx * (j == ny - 1 ? 0 : j + 1);
    double zc[MKZ + 2], zim[MKZ], zip[MKZ], zjm[MKZ], zjp[MKZ], rr[MKZ];
#pragma unroll
    for (int u = 0; u < MKZ + 2; ++u) {
        int k = k0 - 1 + u;
        k = k < 0 ? nz - 1 : (k >= nz ? k - nz : k);
        zc[u] = (k0 + u - 1 <= nz) ? z[col + plane * k] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < MKZ; ++u) {
        const bool in = k0 + u < nz;
        const size_t pk = plane * (in ? k0 + u : 0);
        zim[u] = in ? __ldg(z + pk + im) : 0.0;
        zip[u] = in ? __ldg(z + pk + ip) : 0.0;
        zjm[u] = in ? __ldg(z + pk + jm) : 0.0;
        zjp[u] = in ? __ldg(z + pk + jp) : 0.0;
        rr[u] = (in && MODE != 2) ? r[col + pk] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < MKZ; ++u) {
        if (k0 + u < nz) {
            const double centre = zc[u + 1], c2 = 2.0 * centre;
            double sz = lv.cx * (c2 - zim[u] - zip[u]);
            sz = fma(lv.cy, c2 - zjm[u] - zjp[u], sz);
            sz = fma(lv.cz, c2 - zc[u] - zc[u + 2], sz);
            if (MODE == 2) {
                const double rb = centre - m;
                out[col + plane * (k0 + u)] = lv.wd * (rb + fma(-lv.wd, sz, rb));
            } else {
                const double res = (rr[u] - m) - sz;
                out[col + plane * (k0 + u)] = MODE == 0 ? fma(lv.wd, res, centre) : res;
            }
        }
    }
}

// out = wd (r - m): the first Jacobi sweep from a zero guess
__global__ void __launch_bounds__(256)
mg_scale_kernel(size_t n, double wd, const double *__restrict__ r, const double *__restrict__ mean,
                double *__restrict__ out)
{
    const double m = mean ? *mean : 0.0;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t st = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += st) out[i] = wd * (r[i] - m);
}

// coarse(I,J,K) = sum over the 4x4x4 fine cells 2I-1 .. 2I+2 (periodic) with weights
// (1/8, 3/8, 3/8, 1/8) per direction.  lv = the FINE level; one thread per coarse cell, CTA = 32 x 8
// coarse cells of plane blockIdx.z.
__global__ void __launch_bounds__(MBX * MBY)
mg_restrict_kernel(const __grid_constant__ Lv lv, const double *__restrict__ fine,
                   double *__restrict__ coarse)
{
    const int cnx = lv.nx / 2, cny = lv.ny / 2;
    const int I = blockIdx.x * MBX + threadIdx.x, J = blockIdx.y * MBY + threadIdx.y, K = blockIdx.z;
    if (I >= cnx || J >= cny) return;
    const double w[4] = {0.125, 0.375, 0.375, 0.125};
    const int i0 = 2 * I;
    const int im = i0 == 0 ? lv.nx - 1 : i0 - 1, ip = i0 + 2 >= lv.nx ? 0 : i0 + 2;
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        int k = 2 * K - 1 + c;
        k = k < 0 ? k + lv.nz : (k >= lv.nz ? k - lv.nz : k);
        double sk = 0.0;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            int j = 2 * J - 1 + b;
            j = j < 0 ? j + lv.ny : (j >= lv.ny ? j - lv.ny : j);
            const double *row = fine + (size_t)lv.nx * (j + (size_t)lv.ny * k);
            const double2 mid = *reinterpret_cast<const double2 *>(row + i0);   // i0 is even
            sk = fma(w[b], fma(w[0], __ldg(row + im) + __ldg(row + ip), w[1] * (mid.x + mid.y)), sk);
        }
        s = fma(w[c], sk, s);
    }
    coarse[I + (size_t)cnx * (J + (size_t)cny * K)] = s;
}

// fine += Pr coarse: fine cell 2I takes 3/4 c(I) + 1/4 c(I-1), cell 2I+1 takes 3/4 c(I) + 1/4 c(I+1),
// per direction.  lv = the FINE level.  A thread owns the fine column (i,j) over four fine planes
// 4 B .. 4 B + 3 (B = blockIdx.z): it interpolates the four coarse planes 2 B - 1 .. 2 B + 2 in x and y
// first (16 loads issued together), then along z.
__global__ void __launch_bounds__(MBX * MBY)
mg_prolong_kernel(const __grid_constant__ Lv lv, const double *__restrict__ coarse,
                  double *__restrict__ fine)
{
    const int i = blockIdx.x * MBX + threadIdx.x, j = blockIdx.y * MBY + threadIdx.y;
    if (i >= lv.nx || j >= lv.ny) return;
    const int cnx = lv.nx / 2, cny = lv.ny / 2, cnz = lv.nz / 2;
    auto nb = [](int f, int cn, int &a, int &b) {   // a: the parent (3/4), b: the other one (1/4)
        a = f >> 1;
        b = (f & 1) ? (a + 1 == cn ? 0 : a + 1) : (a == 0 ? cn - 1 : a - 1);
    };
    int ia, ib, ja, jb;
    nb(i, cnx, ia, ib);
    nb(j, cny, ja, jb);
    const int K0 = 2 * blockIdx.z;                    // first of the two parent planes
    double q[4];                                      // in-plane interpolants of planes K0 - 1 .. K0 + 2
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        int K = K0 - 1 + u;
        K = K < 0 ? cnz - 1 : (K >= cnz ? K - cnz : K);
        const double *pl = coarse + (size_t)cnx * cny * K;
        const double a0 = __ldg(pl + ia + (size_t)cnx * ja), b0 = __ldg(pl + ib + (size_t)cnx * ja);
        const double a1 = __ldg(pl + ia + (size_t)cnx * jb), b1 = __ldg(pl + ib + (size_t)cnx * jb);
        q[u] = fma(0.25, fma(0.25, b1, 0.75 * a1), 0.75 * fma(0.25, b0, 0.75 * a0));
    }
    const size_t plane = (size_t)lv.nx * lv.ny, col = i + (size_t)lv.nx * j;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int k = 2 * K0 + u;                     // fine plane; parent K0 + u / 2
        if (k < lv.nz) {
            const double par = q[1 + u / 2], oth = (u & 1) ? q[2 + u / 2] : q[u / 2];
            fine[col + plane * k] += fma(0.25, oth, 0.75 * par);
        }
    }
}

}  // namespace

struct MgLevel {
    Lv lv;
    double *r = nullptr, *z = nullptr, *t = nullptr;   // right-hand side, iterate, spare
    size_t n = 0;
};

struct MgState {
    std::vector<MgLevel> lev;   // lev[0] = the finest grid (r and z are the caller's arrays)
    int nu = 2;
};

namespace {

int sweep(pbx_handle_s *h, int mode, const Lv &lv, const double *z, const double *r, const double *mean,
          double *out)
{
    dim3 block(MBX, MBY), grid((lv.nx + MBX - 1) / MBX, (lv.ny + MBY - 1) / MBY, (lv.nz + MKZ - 1) / MKZ);
    if (mode == 0)
        mg_sweep_kernel<0><<<grid, block, 0, h->stream>>>(lv, z, r, mean, out);
    else if (mode == 1)
        mg_sweep_kernel<1><<<grid, block, 0, h->stream>>>(lv, z, r, mean, out);
    else
        mg_sweep_kernel<2><<<grid, block, 0, h->stream>>>(lv, z, r, mean, out);
    ++h->launches;
    return PBX_OK;
}

unsigned blocks_for(size_t n) { return (unsigned)((n + 255) / 256); }

// `count` Jacobi sweeps on S z = r - mean from a zero guess; the result ends in *cur (the spare in *alt)
int smooth_from_zero(pbx_handle_s *h, const MgLevel &L, const double *r, const double *mean, int count,
                     double **cur, double **alt)
{
    int done = 1;
    if (count >= 2) {
        // two sweeps in one pass; the caller's buffer parity counts sweeps, so the result goes where
        // the second sweep would have put it
        PBX_TRY(sweep(h, 2, L.lv, r, r, mean, *alt));
        std::swap(*cur, *alt);
        done = 2;
    } else {
        unsigned nb = blocks_for(L.n);
        if (nb > 148 * 16) nb = 148 * 16;
        mg_scale_kernel<<<nb, 256, 0, h->stream>>>(L.n, L.lv.wd, r, mean, *cur);
        ++h->launches;
    }
    for (int s = done; s < count; ++s) {
        PBX_TRY(sweep(h, 0, L.lv, *cur, r, mean, *alt));
        std::swap(*cur, *alt);
    }
    return PBX_OK;
}

}  // namespace

void mg_free(pbx_handle_s *h)
{
    MgState *m = (MgState *)h->mg;
    if (!m) return;
    for (size_t l = 0; l < m->lev.size(); ++l) {
        if (l > 0) {
            cudaFree(m->lev[l].r);
            cudaFree(m->lev[l].z);
        }
        cudaFree(m->lev[l].t);
    }
    delete m;
    h->mg = nullptr;
}

int mg_setup(pbx_handle_s *h, int nu)
{
    if (nu < 1 || nu > 8) return PBX_ERR_ARG;
    if (h->mg) {
        ((MgState *)h->mg)->nu = nu;
        return PBX_OK;
    }
    MgState *m = new MgState();
    m->nu = nu;
    h->mg = m;
    int n[3] = {h->nx, h->ny, h->nz};
    double hh[3] = {h->dx[0], h->dx[1], h->dx[2]};
    for (int l = 0;; ++l) {
        MgLevel L;
        L.lv.nx = n[0];
        L.lv.ny = n[1];
        L.lv.nz = n[2];
        L.lv.cx = 1.0 / (hh[0] * hh[0]);
        L.lv.cy = 1.0 / (hh[1] * hh[1]);
        L.lv.cz = 1.0 / (hh[2] * hh[2]);
        L.lv.wd = MG_OMEGA / (2.0 * (L.lv.cx + L.lv.cy + L.lv.cz));
        L.n = (size_t)n[0] * n[1] * n[2];
        m->lev.push_back(L);
        MgLevel &R = m->lev.back();
        if (cudaMalloc(&R.t, L.n * sizeof(double)) != cudaSuccess ||
            (l > 0 && (cudaMalloc(&R.r, L.n * sizeof(double)) != cudaSuccess ||
                       cudaMalloc(&R.z, L.n * sizeof(double)) != cudaSuccess))) {
            cudaGetLastError();
            mg_free(h);
            set_last_error("multigrid hierarchy allocation failed");
            return PBX_ERR_NOMEM;
        }
        const bool coarsest = n[0] <= 4 || n[1] <= 4 || n[2] <= 4 || (n[0] & 1) || (n[1] & 1) || (n[2] & 1);
        if (coarsest) break;
        for (int d = 0; d < 3; ++d) {
            n[d] /= 2;
            hh[d] *= 2.0;
        }
    }
    return PBX_OK;
}

// z = M^-1 (r - mean): one V(nu, nu) cycle on S = -P.  mean: device scalar (may be nullptr = 0).
int mg_vcycle(pbx_handle_s *h, const double *r, const double *mean, double *z)
{
    MgState *m = (MgState *)h->mg;
    if (!m) return PBX_ERR_ARG;
    const int nl = (int)m->lev.size(), nu = m->nu;
    m->lev[0].r = const_cast<double *>(r);
    m->lev[0].z = z;
    std::vector<double *> cur(nl), alt(nl);
    // downward leg
    for (int l = 0; l < nl; ++l) {
        MgLevel &L = m->lev[l];
        const double *mp = l == 0 ? mean : nullptr;
        if (l == nl - 1) {
            // coarsest grid: a fixed number of sweeps, ending in L.z
            cur[l] = (MG_COARSE_SWEEPS & 1) ? L.z : L.t;
            alt[l] = (MG_COARSE_SWEEPS & 1) ? L.t : L.z;
            PBX_TRY(smooth_from_zero(h, L, L.r, mp, MG_COARSE_SWEEPS, &cur[l], &alt[l]));
            break;
        }
        // nu pre-sweeps (the first from the zero guess) + nu post-sweeps = 2 nu - 1 buffer swaps: start
        // in the spare so that the result ends in L.z
        cur[l] = L.t;
        alt[l] = L.z;
        PBX_TRY(smooth_from_zero(h, L, L.r, mp, nu, &cur[l], &alt[l]));
        PBX_TRY(sweep(h, 1, L.lv, cur[l], L.r, mp, alt[l]));          // residual into the spare
        MgLevel &C = m->lev[l + 1];
        mg_restrict_kernel<<<dim3((C.lv.nx + MBX - 1) / MBX, (C.lv.ny + MBY - 1) / MBY, C.lv.nz), dim3(MBX, MBY), 0,
                             h->stream>>>(L.lv, alt[l], C.r);
        ++h->launches;
    }
    // upward leg
    for (int l = nl - 2; l >= 0; --l) {
        MgLevel &L = m->lev[l];
        const double *mp = l == 0 ? mean : nullptr;
        mg_prolong_kernel<<<dim3((L.lv.nx + MBX - 1) / MBX, (L.lv.ny + MBY - 1) / MBY, (L.lv.nz + 3) / 4),
                            dim3(MBX, MBY), 0, h->stream>>>(L.lv, m->lev[l + 1].z, cur[l]);
        ++h->launches;
        for (int s = 0; s < nu; ++s) {
            PBX_TRY(sweep(h, 0, L.lv, cur[l], L.r, mp, alt[l]));
            std::swap(cur[l], alt[l]);
        }
        if (cur[l] != L.z) {
            set_last_error("multigrid buffer parity broken");
            return PBX_ERR_ARG;
        }
    }
    m->lev[0].r = m->lev[0].z = nullptr;
    PBX_CUDA(cudaGetLastError());
    return PBX_OK;
}

}  // namespace pbx

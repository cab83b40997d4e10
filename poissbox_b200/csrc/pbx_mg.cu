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
                double *__restrict__ out, const double *__restrict__ zlo, const double *__restrict__ zhi,
                double *__restrict__ pdn, double *__restrict__ pup)
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
        // plane k0 - 1 + u of the column; below / above the brick: the periodic image, or on a slab
        // (zlo / zhi given) the neighbour rank's plane
        const int k = k0 - 1 + u;
        const double *src = k < 0 ? (zlo ? zlo + col : z + col + plane * (nz - 1))
                                  : (k >= nz ? (zhi ? zhi + col : z + col + plane * (k - nz)) : z + col + plane * k);
        zc[u] = (k <= nz) ? *src : 0.0;
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
            double val;
            if (MODE == 2) {
                const double rb = centre - m;
                val = lv.wd * (rb + fma(-lv.wd, sz, rb));
            } else {
                const double res = (rr[u] - m) - sz;
                val = MODE == 0 ? fma(lv.wd, res, centre) : res;
            }
            out[col + plane * (k0 + u)] = val;
            // slabs: my bottom / top plane goes straight into the neighbours' halo buffers as well
            if (pdn && k0 + u == 0) pdn[col] = val;
            if (pup && k0 + u == nz - 1) pup[col] = val;
        }
    }
}

// out = wd (r - m): the first Jacobi sweep from a zero guess
__global__ void __launch_bounds__(256)
mg_scale_kernel(size_t n, double wd, const double *__restrict__ r, const double *__restrict__ mean,
                double *__restrict__ out, size_t plane, double *__restrict__ pdn, double *__restrict__ pup)
{
    const double m = mean ? *mean : 0.0;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t st = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += st) {
        const double val = wd * (r[i] - m);
        out[i] = val;
        if (pdn && i < plane) pdn[i] = val;
        if (pup && i >= n - plane) pup[i - (n - plane)] = val;
    }
}

// coarse(I,J,K) = sum over the 4x4x4 fine cells 2I-1 .. 2I+2 (periodic) with weights
// (1/8, 3/8, 3/8, 1/8) per direction.  lv = the FINE level; one thread per coarse cell, CTA = 32 x 8
// coarse cells of plane blockIdx.z.
__global__ void __launch_bounds__(MBX * MBY)
mg_restrict_kernel(const __grid_constant__ Lv lv, const double *__restrict__ fine,
                   double *__restrict__ coarse, const double *__restrict__ flo, const double *__restrict__ fhi,
                   double *__restrict__ pdn, double *__restrict__ pup)
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
        const int k = 2 * K - 1 + c;
        // fine plane k; below / above the brick: periodic image or the neighbour rank's plane
        const double *pl = k < 0 ? (flo ? flo : fine + (size_t)lv.nx * lv.ny * (k + lv.nz))
                                 : (k >= lv.nz ? (fhi ? fhi : fine + (size_t)lv.nx * lv.ny * (k - lv.nz))
                                               : fine + (size_t)lv.nx * lv.ny * k);
        double sk = 0.0;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            int j = 2 * J - 1 + b;
            j = j < 0 ? j + lv.ny : (j >= lv.ny ? j - lv.ny : j);
            const double *row = pl + (size_t)lv.nx * j;
            const double2 mid = *reinterpret_cast<const double2 *>(row + i0);   // i0 is even
            sk = fma(w[b], fma(w[0], __ldg(row + im) + __ldg(row + ip), w[1] * (mid.x + mid.y)), sk);
        }
        s = fma(w[c], sk, s);
    }
    coarse[I + (size_t)cnx * (J + (size_t)cny * K)] = s;
    if (pdn && K == 0) pdn[I + (size_t)cnx * J] = s;
    if (pup && K == (int)gridDim.z - 1) pup[I + (size_t)cnx * J] = s;
}

// fine += Pr coarse: fine cell 2I takes 3/4 c(I) + 1/4 c(I-1), cell 2I+1 takes 3/4 c(I) + 1/4 c(I+1),
// per direction.  lv = the FINE level.  A thread owns the fine column (i,j) over four fine planes
// 4 B .. 4 B + 3 (B = blockIdx.z): it interpolates the four coarse planes 2 B - 1 .. 2 B + 2 in x and y
// first (16 loads issued together), then along z.
__global__ void __launch_bounds__(MBX * MBY)
mg_prolong_kernel(const __grid_constant__ Lv lv, const double *__restrict__ coarse,
                  double *__restrict__ fine, const double *__restrict__ clo, const double *__restrict__ chi,
                  double *__restrict__ pdn, double *__restrict__ pup)
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
        const int K = K0 - 1 + u;
        const size_t cplane = (size_t)cnx * cny;
        const double *pl = K < 0 ? (clo ? clo : coarse + cplane * (cnz - 1))
                                 : (K >= cnz ? (chi ? chi : coarse + cplane * (K - cnz)) : coarse + cplane * K);
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
            const double val = fine[col + plane * k] + fma(0.25, oth, 0.75 * par);
            fine[col + plane * k] = val;
            if (pdn && k == 0) pdn[col] = val;
            if (pup && k == lv.nz - 1) pup[col] = val;
        }
    }
}

}  // namespace

struct MgLevel {
    Lv lv;
    double *r = nullptr, *z = nullptr, *t = nullptr;   // right-hand side, iterate, spare
    size_t n = 0;
};

// One hierarchy: lev[0] is its finest grid.  `own0`: r and z of lev[0] belong to the hierarchy (false
// for the handle's finest grid, whose r and z are the caller's vectors).
struct MgHier {
    std::vector<MgLevel> lev;
    bool own0 = false;
};

// On z slabs (nranks > 1) the cycle is the same cycle, split at level g (mg_slab_plan):
//   levels 0 .. g-1   distributed like the fine grid: every rank owns its planes; a kernel that needs
//                     the planes next to the slab gets them from the neighbours first (dist_halo_planes:
//                     one plane each way, the exchange of the star operator);
//   levels g ..       too thin to cut (<= 4 planes per rank): the right-hand side of level g is
//                     all-gathered (dist_allgather) and every rank runs the rest of the cycle on the
//                     whole coarse grid, redundantly, then prolongs its own part.
// Point for point the arithmetic is that of the single-rank cycle, so the result carries the same bits.
struct MgState {
    MgHier top;        // single rank: the whole hierarchy; slabs: the distributed levels 0 .. g-1
    MgHier rep;        // slabs: levels g .. on the whole (global) coarse grid, replicated
    double *gsrc = nullptr;   // slabs: my planes of level g's right-hand side
    size_t gcount = 0;        //        their number of doubles
    int nu = 2;
};

namespace {

// where a producer kernel publishes its bottom / top plane (the neighbours' halo buffers), and where a
// consumer finds the planes next to its slab
struct Pub {
    double *dn = nullptr, *up = nullptr;
};
struct Nb {
    const double *lo = nullptr, *hi = nullptr;
};

int sweep(pbx_handle_s *h, int mode, const Lv &lv, const double *z, const double *r, const double *mean,
          double *out, Nb nb, Pub pub)
{
    dim3 block(MBX, MBY), grid((lv.nx + MBX - 1) / MBX, (lv.ny + MBY - 1) / MBY, (lv.nz + MKZ - 1) / MKZ);
    if (mode == 0)
        mg_sweep_kernel<0><<<grid, block, 0, h->stream>>>(lv, z, r, mean, out, nb.lo, nb.hi, pub.dn, pub.up);
    else if (mode == 1)
        mg_sweep_kernel<1><<<grid, block, 0, h->stream>>>(lv, z, r, mean, out, nb.lo, nb.hi, pub.dn, pub.up);
    else
        mg_sweep_kernel<2><<<grid, block, 0, h->stream>>>(lv, z, r, mean, out, nb.lo, nb.hi, pub.dn, pub.up);
    ++h->launches;
    return PBX_OK;
}

int restrict_to(pbx_handle_s *h, const Lv &fine, const double *res, double *coarse_r, Nb nb, Pub pub)
{
    const int cnx = fine.nx / 2, cny = fine.ny / 2, cnz = fine.nz / 2;
    mg_restrict_kernel<<<dim3((cnx + MBX - 1) / MBX, (cny + MBY - 1) / MBY, cnz), dim3(MBX, MBY), 0, h->stream>>>(
        fine, res, coarse_r, nb.lo, nb.hi, pub.dn, pub.up);
    ++h->launches;
    return PBX_OK;
}

int prolong_add(pbx_handle_s *h, const Lv &fine, const double *coarse_z, double *fine_z, Nb nb, Pub pub)
{
    mg_prolong_kernel<<<dim3((fine.nx + MBX - 1) / MBX, (fine.ny + MBY - 1) / MBY, (fine.nz + 3) / 4),
                        dim3(MBX, MBY), 0, h->stream>>>(fine, coarse_z, fine_z, nb.lo, nb.hi, pub.dn, pub.up);
    ++h->launches;
    return PBX_OK;
}

unsigned blocks_for(size_t n) { return (unsigned)((n + 255) / 256); }

// The halo protocol of a slab cycle.  A kernel whose output the next kernel needs WITH its neighbour
// planes publishes its own bottom and top plane into the neighbours' buffers while it writes them
// (begin: the buffers of the next exchange round); the consumer is launched after the barrier of that
// round (end) with the planes that arrived.  One round is in flight at a time; the receive buffers
// alternate, so a kernel may read the planes of round e while it publishes for round e + 1.
struct Halo {
    pbx_handle_s *h;
    bool slab;
    int begin(Pub *p) const
    {
        *p = Pub();
        return slab ? dist_halo_begin(h, &p->dn, &p->up) : PBX_OK;
    }
    int end(Nb *n) const
    {
        *n = Nb();
        return slab ? dist_halo_end(h, &n->lo, &n->hi) : PBX_OK;
    }
    // a field no kernel of the cycle produced (the right-hand side of the finest level)
    int planes(const Lv &lv, const double *field, Nb *n) const
    {
        *n = Nb();
        return slab ? dist_halo_planes(h, field, (size_t)lv.nx * lv.ny, lv.nz, &n->lo, &n->hi) : PBX_OK;
    }
};

// `count` Jacobi sweeps on S z = r - mean from a zero guess; the result ends in *cur (the spare in
// *alt) and has been published (Halo::begin) if `publish`.  r_nb: the planes next to r (count >= 2).
int smooth_from_zero(const Halo &H, const MgLevel &L, const double *r, Nb r_nb, const double *mean, int count,
                     double **cur, double **alt, bool publish)
{
    pbx_handle_s *h = H.h;
    Pub pub;
    int done = 1;
    if (count >= 2) {
        // two sweeps in one pass; the caller's buffer parity counts sweeps, so the result goes where
        // the second sweep would have put it
        if (count > 2 || publish) PBX_TRY(H.begin(&pub));
        PBX_TRY(sweep(h, 2, L.lv, r, r, mean, *alt, r_nb, pub));
        std::swap(*cur, *alt);
        done = 2;
    } else {
        unsigned nb = blocks_for(L.n);
        if (nb > 148 * 16) nb = 148 * 16;
        if (publish) PBX_TRY(H.begin(&pub));
        mg_scale_kernel<<<nb, 256, 0, h->stream>>>(L.n, L.lv.wd, r, mean, *cur, (size_t)L.lv.nx * L.lv.ny, pub.dn,
                                                   pub.up);
        ++h->launches;
    }
    for (int s = done; s < count; ++s) {
        Nb nbp;
        PBX_TRY(H.end(&nbp));
        pub = Pub();
        if (s + 1 < count || publish) PBX_TRY(H.begin(&pub));
        PBX_TRY(sweep(h, 0, L.lv, *cur, r, mean, *alt, nbp, pub));
        std::swap(*cur, *alt);
    }
    return PBX_OK;
}

void free_hier(MgHier &H)
{
    for (size_t l = 0; l < H.lev.size(); ++l) {
        if (l > 0 || H.own0) {
            cudaFree(H.lev[l].r);
            cudaFree(H.lev[l].z);
        }
        cudaFree(H.lev[l].t);
    }
    H.lev.clear();
}

bool is_coarsest(const int n[3])
{
    return n[0] <= 4 || n[1] <= 4 || n[2] <= 4 || (n[0] & 1) || (n[1] & 1) || (n[2] & 1);
}

// levels from grid n (spacing hh) down: at most `maxlev` of them (0: until the coarsest grid); the z
// extent of a level's arrays is n[2] / zdiv (zdiv > 1: my planes of a slab decomposition)
int build_hier(MgHier &H, const int n0[3], const double hh0[3], int zdiv, int maxlev, bool own0)
{
    int n[3] = {n0[0], n0[1], n0[2]};
    double hh[3] = {hh0[0], hh0[1], hh0[2]};
    H.own0 = own0;
    for (int l = 0;; ++l) {
        MgLevel L;
        L.lv.nx = n[0];
        L.lv.ny = n[1];
        L.lv.nz = n[2] / zdiv;
        L.lv.cx = 1.0 / (hh[0] * hh[0]);
        L.lv.cy = 1.0 / (hh[1] * hh[1]);
        L.lv.cz = 1.0 / (hh[2] * hh[2]);
        L.lv.wd = MG_OMEGA / (2.0 * (L.lv.cx + L.lv.cy + L.lv.cz));
        L.n = (size_t)L.lv.nx * L.lv.ny * L.lv.nz;
        H.lev.push_back(L);
        MgLevel &R = H.lev.back();
        if (cudaMalloc(&R.t, L.n * sizeof(double)) != cudaSuccess ||
            ((l > 0 || own0) && (cudaMalloc(&R.r, L.n * sizeof(double)) != cudaSuccess ||
                                 cudaMalloc(&R.z, L.n * sizeof(double)) != cudaSuccess))) {
            cudaGetLastError();
            set_last_error("multigrid hierarchy allocation failed");
            return PBX_ERR_NOMEM;
        }
        if (is_coarsest(n) || (maxlev > 0 && l + 1 == maxlev)) break;
        for (int d = 0; d < 3; ++d) {
            n[d] /= 2;
            hh[d] *= 2.0;
        }
    }
    return PBX_OK;
}

// One V(nu, nu) cycle on hierarchy H (lev[0].r and lev[0].z set by the caller).  slab: the levels are
// my planes of a z-decomposed box.  `below`: called when the downward leg has produced the right-hand
// side `next_r` of the level under H's last one (nullptr: H ends with the coarsest grid); it returns
// that level's solution as (coarse, planes next to my part of it).
template <class Below>
int cycle(pbx_handle_s *h, MgHier &Hi, bool slab, int nu, const double *mean, double *next_r, Below below)
{
    const Halo H{h, slab};
    const int nl = (int)Hi.lev.size();
    std::vector<double *> cur(nl), alt(nl);
    const int ndown = next_r ? nl : nl - 1;        // levels that smooth, form a residual and restrict it
    for (int l = 0; l < nl; ++l) {
        MgLevel &L = Hi.lev[l];
        const double *mp = l == 0 ? mean : nullptr;
        const int count = l == ndown ? MG_COARSE_SWEEPS : nu;
        // the planes next to r: published by the restriction that produced it, or (finest level) fetched
        Nb r_nb;
        if (count >= 2) PBX_TRY(l == 0 ? H.planes(L.lv, L.r, &r_nb) : H.end(&r_nb));
        if (l == ndown) {
            // coarsest grid: a fixed number of sweeps, ending in L.z
            cur[l] = (MG_COARSE_SWEEPS & 1) ? L.z : L.t;
            alt[l] = (MG_COARSE_SWEEPS & 1) ? L.t : L.z;
            PBX_TRY(smooth_from_zero(H, L, L.r, r_nb, mp, count, &cur[l], &alt[l], l > 0));
            break;
        }
        // nu pre-sweeps (the first from the zero guess) + nu post-sweeps = 2 nu - 1 buffer swaps: start
        // in the spare so that the result ends in L.z
        cur[l] = L.t;
        alt[l] = L.z;
        PBX_TRY(smooth_from_zero(H, L, L.r, r_nb, mp, count, &cur[l], &alt[l], true));
        Nb nb;
        Pub pub;
        PBX_TRY(H.end(&nb));
        PBX_TRY(H.begin(&pub));
        PBX_TRY(sweep(h, 1, L.lv, cur[l], L.r, mp, alt[l], nb, pub));          // residual into the spare
        PBX_TRY(H.end(&nb));
        pub = Pub();
        // the next level's first sweeps read the planes next to its right-hand side (two or more sweeps)
        const bool next_in_h = l + 1 < nl;
        const int next_count = (l + 1 == ndown) ? MG_COARSE_SWEEPS : nu;
        if (next_in_h && next_count >= 2) PBX_TRY(H.begin(&pub));
        PBX_TRY(restrict_to(h, L.lv, alt[l], next_in_h ? Hi.lev[l + 1].r : next_r, nb, pub));
    }
    for (int l = ndown - 1; l >= 0; --l) {
        MgLevel &L = Hi.lev[l];
        const double *mp = l == 0 ? mean : nullptr;
        const double *coarse = nullptr;
        Nb nb;
        if (l + 1 < nl) {
            coarse = Hi.lev[l + 1].z;       // published by the last sweep of the level below
            PBX_TRY(H.end(&nb));
        } else {
            PBX_TRY(below(&coarse, &nb.lo, &nb.hi));
        }
        Pub pub;
        PBX_TRY(H.begin(&pub));
        PBX_TRY(prolong_add(h, L.lv, coarse, cur[l], nb, pub));
        for (int s = 0; s < nu; ++s) {
            PBX_TRY(H.end(&nb));
            pub = Pub();
            if (s + 1 < nu || l > 0) PBX_TRY(H.begin(&pub));   // the finest level's last sweep feeds the CG
            PBX_TRY(sweep(h, 0, L.lv, cur[l], L.r, mp, alt[l], nb, pub));
            std::swap(cur[l], alt[l]);
        }
        if (cur[l] != L.z) {
            set_last_error("multigrid buffer parity broken");
            return PBX_ERR_ARG;
        }
    }
    return PBX_OK;
}

int no_below(const double **, const double **, const double **) { return PBX_ERR_ARG; }

}  // namespace

// Slab decomposition: the number g of distributed levels for bricks of nzl planes on each of P ranks
// (0: the hierarchy cannot be cut) and the size of level g's global grid, which is all-gathered.
int mg_slab_plan(int nx, int ny, int nzl, int nranks, size_t *gather_doubles)
{
    int n[3] = {nx, ny, nzl * nranks};
    int g = 0, loc = nzl;
    while (!is_coarsest(n) && loc > 4 && !(loc & 1)) {
        for (int d = 0; d < 3; ++d) n[d] /= 2;
        loc /= 2;
        ++g;
    }
    if (gather_doubles) *gather_doubles = (size_t)n[0] * n[1] * n[2];
    return g;
}

void mg_free(pbx_handle_s *h)
{
    MgState *m = (MgState *)h->mg;
    if (!m) return;
    free_hier(m->top);
    free_hier(m->rep);
    if (m->gsrc) cudaFree(m->gsrc);
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
    const int P = h->nranks;
    const int n[3] = {h->nx, h->ny, h->nz * P};
    int rc;
    if (P == 1) {
        rc = build_hier(m->top, n, h->dx, 1, 0, false);
    } else {
        size_t gd = 0;
        const int g = mg_slab_plan(h->nx, h->ny, h->nz, P, &gd);
        if (g < 1) {
            mg_free(h);
            set_last_error("multigrid on slabs: the brick cannot be coarsened");
            return PBX_ERR_UNSUPPORTED;
        }
        rc = build_hier(m->top, n, h->dx, P, g, false);
        if (rc == PBX_OK) {
            int ng[3] = {n[0] >> g, n[1] >> g, n[2] >> g};
            double hg[3] = {h->dx[0] * (1 << g), h->dx[1] * (1 << g), h->dx[2] * (1 << g)};
            rc = build_hier(m->rep, ng, hg, 1, 0, true);
            m->gcount = gd / P;
            if (rc == PBX_OK && cudaMalloc(&m->gsrc, m->gcount * sizeof(double)) != cudaSuccess) {
                cudaGetLastError();
                set_last_error("multigrid hierarchy allocation failed");
                rc = PBX_ERR_NOMEM;
            }
        }
    }
    if (rc != PBX_OK) mg_free(h);
    return rc;
}

// z = M^-1 (r - mean): one V(nu, nu) cycle on S = -P.  mean: device scalar (may be nullptr = 0).
int mg_vcycle(pbx_handle_s *h, const double *r, const double *mean, double *z)
{
    MgState *m = (MgState *)h->mg;
    if (!m) return PBX_ERR_ARG;
    m->top.lev[0].r = const_cast<double *>(r);
    m->top.lev[0].z = z;
    int rc;
    if (h->nranks == 1) {
        rc = cycle(h, m->top, false, m->nu, mean, nullptr, no_below);
    } else {
        rc = cycle(h, m->top, true, m->nu, mean, m->gsrc,
                   [&](const double **coarse, const double **lo, const double **hi) -> int {
                       // level g: gather its right-hand side, solve the rest of the hierarchy on the
                       // whole coarse grid (every rank the same), hand back my planes and their neighbours
                       double *full = nullptr;
                       PBX_TRY(dist_allgather(h, m->gsrc, m->gcount, &full));
                       MgLevel &G = m->rep.lev[0];
                       PBX_CUDA(cudaMemcpyAsync(G.r, full, G.n * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
                       PBX_TRY(cycle(h, m->rep, false, m->nu, nullptr, nullptr, no_below));
                       const size_t plane = (size_t)G.lv.nx * G.lv.ny;
                       const int nzg = G.lv.nz, mine = nzg / h->nranks, k0 = h->rank * mine;
                       *coarse = G.z + plane * k0;
                       *lo = G.z + plane * ((k0 + nzg - 1) % nzg);
                       *hi = G.z + plane * ((k0 + mine) % nzg);
                       return PBX_OK;
                   });
    }
    m->top.lev[0].r = m->top.lev[0].z = nullptr;
    PBX_TRY(rc);
    PBX_CUDA(cudaGetLastError());
    return PBX_OK;
}

}  // namespace pbx

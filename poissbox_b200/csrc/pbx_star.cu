// pbx_star.cu -- the 2nd-order 7-point star Laplacian on a periodic box: the operator the
// reference's MATSHELL callback applies TODAY (mfmult -> compute_lapl_pointwise ->
// evaluate_laplacian_pointwise, src/poissbox.f90:84-148, 300-322, with the coefficients of
// src/coefficients.f90:22-48), and the operator P the reference hands to KSPSetOperators as the
// preconditioning matrix (src/poissbox.f90:294).
//
// Bit-identical to the CPU oracle: the reference evaluates dot_product(f(27), coeffs(27)) in
// array element order; twenty coefficients are exactly zero, and adding +-0 never changes the
// running sum (which starts at +0 and can never become -0), so the seven non-zero terms taken in
// the same order with round-to-nearest intrinsics give the same bits for finite inputs:
//   (i,j,k-1) cz, (i,j-1,k) cy, (i-1,j,k) cx, (i,j,k) c0, (i+1,j,k) cx, (i,j+1,k) cy, (i,j,k+1) cz.
//
// HBM-bound streaming kernel, 16 B/DoF: a thread owns one (i,j) column of a block of planes and
// marches along k with the three z values in registers; the four in-plane neighbours are other
// threads' centre values of the same plane and come out of L1.
#include "pbx_internal.h"

namespace pbx {

namespace {

struct StarCoef {
    double cx, cy, cz, c0;
};

StarCoef star_coef(const double dx[3])
{
    // src/coefficients.f90:29-33 and :45-48, same operations in the same order
    volatile double ix = 1.0 / (dx[0] * dx[0]), iy = 1.0 / (dx[1] * dx[1]), iz = 1.0 / (dx[2] * dx[2]);
    volatile double c = 0.0;
    c = c + -2.0 * ix;
    c = c + -2.0 * iy;
    c = c + -2.0 * iz;
    return {ix, iy, iz, c};
}

constexpr int SBX = 32, SBY = 8;   // CTA = 32 x 8 columns
constexpr int SKZ = 4;             // planes per thread

// lo / up: the plane below k = 0 / above k = nz - 1 when the brick is one slab of a z-decomposed
// box (nullptr: periodic in z within the brick).  All of a thread's loads (SKZ + 2 values of its own
// column, 4 SKZ in-plane neighbours) are issued before the arithmetic.
__global__ void __launch_bounds__(SBX * SBY)
star_kernel(int nx, int ny, int nz, const __grid_constant__ StarCoef c, const double *__restrict__ x,
            const double *__restrict__ lo, const double *__restrict__ up, double *__restrict__ y)
{
    const int i = blockIdx.x * SBX + threadIdx.x, j = blockIdx.y * SBY + threadIdx.y;
    if (i >= nx || j >= ny) return;
    const int k0 = blockIdx.z * SKZ;
    const size_t plane = (size_t)nx * ny;
    const size_t col = i + (size_t)nx * j;
    const size_t im = (i == 0 ? nx - 1 : i - 1) + (size_t)nx * j, ip = (i == nx - 1 ? 0 : i + 1) + (size_t)nx * j;
    const size_t jm = i + (size_t)nx * (j == 0 ? ny - 1 : j - 1), jp = i + (size_t)nx * (j == ny - 1 ? 0 : j + 1);
    double xc[SKZ + 2], xim[SKZ], xip[SKZ], xjm[SKZ], xjp[SKZ];
#pragma unroll
    for (int u = 0; u < SKZ + 2; ++u) {
        const int k = k0 - 1 + u;
        if (k > nz)
            xc[u] = 0.0;
        else if (k < 0)
            xc[u] = lo ? lo[col] : x[col + plane * (nz - 1)];
        else if (k == nz)
            xc[u] = up ? up[col] : x[col];
        else
            xc[u] = x[col + plane * k];
    }
#pragma unroll
    for (int u = 0; u < SKZ; ++u) {
        const bool in = k0 + u < nz;
        const size_t pk = plane * (in ? k0 + u : 0);
        xim[u] = in ? __ldg(x + pk + im) : 0.0;
        xip[u] = in ? __ldg(x + pk + ip) : 0.0;
        xjm[u] = in ? __ldg(x + pk + jm) : 0.0;
        xjp[u] = in ? __ldg(x + pk + jp) : 0.0;
    }
#pragma unroll
    for (int u = 0; u < SKZ; ++u) {
        if (k0 + u < nz) {
            double s = __dadd_rn(0.0, __dmul_rn(xc[u], c.cz));
            s = __dadd_rn(s, __dmul_rn(xjm[u], c.cy));
            s = __dadd_rn(s, __dmul_rn(xim[u], c.cx));
            s = __dadd_rn(s, __dmul_rn(xc[u + 1], c.c0));
            s = __dadd_rn(s, __dmul_rn(xip[u], c.cx));
            s = __dadd_rn(s, __dmul_rn(xjp[u], c.cy));
            s = __dadd_rn(s, __dmul_rn(xc[u + 2], c.cz));
            y[col + plane * (k0 + u)] = s;
        }
    }
}

}  // namespace

int star_apply(pbx_handle_s *h, const double *x, double *y, const double *lo, const double *up)
{
    const StarCoef c = star_coef(h->dx);
    dim3 block(SBX, SBY), grid((h->nx + SBX - 1) / SBX, (h->ny + SBY - 1) / SBY, (h->nz + SKZ - 1) / SKZ);
    if (grid.y > 65535 || grid.z > 65535) return PBX_ERR_UNSUPPORTED;
    star_kernel<<<grid, block, 0, h->stream>>>(h->nx, h->ny, h->nz, c, x, lo, up, y);
    ++h->launches;
    PBX_CUDA(cudaGetLastError());
    return PBX_OK;
}

}  // namespace pbx

// pbx_coeffs.cu -- host-side coefficient tables for both Laplacian schedules.
#include <cmath>
#include <cstdlib>
#include <vector>

#include "pbx_internal.h"

namespace pbx {

// FAST schedule: see the derivation in pbx_internal.h.  With s = opsign (+1 interpolation, -1
// derivative) the two right-hand-side stencils of src/compact_schemes.f90:332-372 are
//   B^- = a (1 + s E^-1) + b (E + s E^-2)        (stagger -1, :356-358,364,367)
//   B^+ = a (E + s)      + b (E^2 + s E^-1)      (stagger +1, :360,364,369-370)
// and their product is the symmetric stencil
//   S = s(2a^2 + 2b^2) + (a^2 + 2 s a b)(E + E^-1) + 2ab (E^2 + E^-2) + b^2 (E^3 + E^-3).
void make_composite_coef(OpKind kind, double dx, CompositeCoef *o)
{
    double a, b;
    scheme_ab(kind, dx, &a, &b);
    const double al = scheme_alpha(kind);
    const double s = (kind == OP_DERIV) ? -1.0 : 1.0;
    const double r = (-1.0 + std::sqrt(1.0 - 4.0 * al * al)) / (2.0 * al);
    const double sc = (1.0 + r * r) * (1.0 + r * r);
    o->c0 = sc * s * (2.0 * a * a + 2.0 * b * b);
    o->c1 = sc * (a * a + 2.0 * s * a * b);
    o->c2 = sc * (2.0 * a * b);
    o->c3 = sc * (b * b);
    o->r = r;
    double p = 1.0;
    for (int k = 0; k < LC; ++k) {
        p *= r;
        o->pw[k] = p;
    }
    // look[m] = r^(LC*m)
    const double rl = o->pw[LC - 1];
    double q = 1.0;
    o->nlook = MAXLOOK;
    for (int m = 0; m < MAXLOOK; ++m) {
        o->look[m] = q;
        q *= rl;
    }
    // number of look-back levels needed: the first neglected term is ~ (LC*m) * |r|^(LC*m)
    for (int m = 1; m <= MAXLOOK; ++m) {
        double tail = (double)(LC * m) * std::pow(std::fabs(r), LC * m);
        if (tail < 1e-19) {
            o->nlook = m;
            break;
        }
    }
    o->pad_ = 0;
}

// REFERENCE schedule: replay src/tridsol.f90:51-66 for a(:) = c(:) = alpha, b(:) = 1.
int make_ref_tables(int n, double alpha, RefLineTables *t)
{
    std::vector<double> a(n, alpha), c(n, alpha), bmod(n, 1.0), w(n, 0.0), u(n, 0.0);
    const double gamma = -1.0;                       // :51, b(1) = 1
    volatile double t0;
    bmod[0] = bmod[0] - gamma;                       // :55
    t0 = c[n - 1] * a[0];
    t0 = t0 / gamma;
    bmod[n - 1] = bmod[n - 1] - t0;                  // :56
    // forward sweep of the pivots (:90-92); the data part d(i) -= w d(i-1) runs on the device
    for (int i = 1; i < n; ++i) {
        volatile double wi = a[i] / bmod[i - 1];
        volatile double prod = wi * c[i - 1];
        w[i] = wi;
        bmod[i] = bmod[i] - prod;
    }
    // auxiliary system (:62-66): u = (gamma, 0, ..., 0, c(n)), same matrix
    u[0] = gamma;
    u[n - 1] = c[n - 1];
    for (int i = 1; i < n; ++i) {
        volatile double prod = w[i] * u[i - 1];
        u[i] = u[i] - prod;
    }
    u[n - 1] = u[n - 1] / bmod[n - 1];
    for (int i = n - 2; i >= 0; --i) {
        volatile double prod = c[i] * u[i + 1];
        volatile double diff = u[i] - prod;
        u[i] = diff / bmod[i];
    }
    t->n = n;
    t->alpha = alpha;
    t->a1g = a[0] / gamma;
    {
        volatile double prod = t->a1g * u[n - 1];
        volatile double sum = u[0] + prod;
        t->den = 1.0 + sum;                          // :70
    }
    size_t bytes = sizeof(double) * (size_t)n;
    PBX_CUDA(cudaMalloc(&t->w, bytes));
    PBX_CUDA(cudaMalloc(&t->piv, bytes));
    PBX_CUDA(cudaMalloc(&t->u, bytes));
    PBX_CUDA(cudaMemcpy(t->w, w.data(), bytes, cudaMemcpyHostToDevice));
    PBX_CUDA(cudaMemcpy(t->piv, bmod.data(), bytes, cudaMemcpyHostToDevice));
    PBX_CUDA(cudaMemcpy(t->u, u.data(), bytes, cudaMemcpyHostToDevice));
    return PBX_OK;
}

void free_ref_tables(RefLineTables *t)
{
    if (t->w) cudaFree(t->w);
    if (t->piv) cudaFree(t->piv);
    if (t->u) cudaFree(t->u);
    t->w = t->piv = t->u = nullptr;
}

}  // namespace pbx

// pbx_host.hpp -- C++ host-side mirror of the reference's Fortran module interfaces on top of the
// C ABI (pbx.h).  The reference is compiled code; this header gives a C++ caller the same names,
// argument meaning and error behaviour:
//
//   namespace compact_schemes  <->  module compact_schemes  (src/compact_schemes.f90:9-13)
//   namespace tridsol          <->  module tridsol          (src/tridsol.f90:16-18)
//
// Arrays are plain contiguous buffers in Fortran order f(i,j,k) = f[i + nx*(j + ny*k)].  A length
// mismatch terminates with status 7 exactly like the reference's `stop 7`
// (src/compact_schemes.f90:177-180); any other library failure terminates with status 2.
#pragma once

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "pbx.h"

namespace pbx_host {

inline void check(int rc)
{
    if (rc == PBX_OK) return;
    if (rc == PBX_ERR_SIZE) {
        std::puts(" ERROR: periodic gradient is same length as field!");
        std::exit(7);
    }
    std::fprintf(stderr, "libpbx error %d: %s\n", rc, pbx_last_error());
    std::exit(2);
}

// a dense Fortran-ordered array with up to four extents
struct Field {
    int n[4] = {1, 1, 1, 1};
    std::vector<double> v;
    Field() = default;
    Field(int n1, int n2 = 1, int n3 = 1, int n4 = 1) : n{n1, n2, n3, n4}, v((size_t)n1 * n2 * n3 * n4) {}
    double &operator()(int i, int j = 0, int k = 0, int c = 0)
    {
        return v[i + (size_t)n[0] * (j + (size_t)n[1] * (k + (size_t)n[2] * c))];
    }
    double operator()(int i, int j = 0, int k = 0, int c = 0) const
    {
        return v[i + (size_t)n[0] * (j + (size_t)n[1] * (k + (size_t)n[2] * c))];
    }
    void fill(double x) { v.assign(v.size(), x); }
    size_t size() const { return v.size(); }
};

}  // namespace pbx_host

namespace compact_schemes {

using pbx_host::Field;

inline int lapl_mode = PBX_MODE_FAST;

// src/compact_schemes.f90:17-37
inline void lapl(const Field &f, const double dx[3], Field &d2fdx2)
{
    pbx_host::check(pbx_lapl_host(f.n[0], f.n[1], f.n[2], f.v.data(), dx, d2fdx2.v.data(), lapl_mode));
}
// :42-88
inline void grad(const Field &f, const double dx[3], Field &df)
{
    pbx_host::check(pbx_grad_host(f.n[0], f.n[1], f.n[2], f.v.data(), dx, df.v.data()));
}
// :207-257
inline void div(const Field &f, const double dx[3], Field &df)
{
    pbx_host::check(pbx_div_host(f.n[0], f.n[1], f.n[2], f.v.data(), dx, df.v.data()));
}
// :93-142 (opt_stagger defaults to -1)
inline void interp(const Field &f, Field &fi, int opt_stagger = -1)
{
    pbx_host::check(pbx_interp_host(f.n[0], f.n[1], f.n[2], f.v.data(), fi.v.data(), opt_stagger));
}
// :144-152
inline void interp_div(const Field &f, Field &fi) { interp(f, fi, +1); }
// :155-204
inline void grad_1d(const std::vector<double> &f, double dx, std::vector<double> &df, int opt_stagger = -1)
{
    pbx_host::check(pbx_grad_1d_host((int)f.size(), f.data(), dx, (int)df.size(), df.data(), opt_stagger));
}
// :260-268
inline void div_1d(const std::vector<double> &f, double dx, std::vector<double> &df) { grad_1d(f, dx, df, +1); }
// :271-319
inline void interp_1d(const std::vector<double> &f, std::vector<double> &fi, int opt_stagger = -1)
{
    pbx_host::check(pbx_interp_1d_host((int)f.size(), f.data(), (int)fi.size(), fi.data(), opt_stagger));
}
// :322-329
inline void interp_1d_div(const std::vector<double> &f, std::vector<double> &fi) { interp_1d(f, fi, +1); }

}  // namespace compact_schemes

namespace tridsol {

// dummy names as in the reference: a sub-diagonal, b DIAGONAL, c super-diagonal, d rhs/solution
inline void tdma(const std::vector<double> &a, std::vector<double> &b, const std::vector<double> &c,
                 std::vector<double> &d)
{
    pbx_host::check(pbx_tdma_host((int)d.size(), a.data(), b.data(), c.data(), d.data()));
}
inline void tdma_periodic(const std::vector<double> &a, std::vector<double> &b, const std::vector<double> &c,
                          std::vector<double> &d)
{
    pbx_host::check(pbx_tdma_periodic_host((int)d.size(), a.data(), b.data(), c.data(), d.data()));
}
inline void fwd_sweep(const std::vector<double> &a, std::vector<double> &b, const std::vector<double> &c,
                      std::vector<double> &d)
{
    pbx_host::check(pbx_fwd_sweep_host((int)d.size(), a.data(), b.data(), c.data(), d.data()));
}
inline void bwd_sweep(const std::vector<double> &b, const std::vector<double> &c, std::vector<double> &d)
{
    pbx_host::check(pbx_bwd_sweep_host((int)d.size(), b.data(), c.data(), d.data()));
}

}  // namespace tridsol

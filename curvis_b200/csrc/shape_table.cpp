// shape_table.cpp — host-side generation of the table described in shape_table.h (x87 long double:
// 64-bit significand, glibc atanl / log1pl).
#include "shape_table.h"
#include <cfloat>
#include <cmath>

// The Chebyshev fits below need a 64-bit significand: with long double == double (e.g. -mlong-double-64) the coefficients
// would silently lose several bits and the fast kernels would drift with no signal.
static_assert(LDBL_MANT_DIG >= 64, "shape_table.cpp needs an extended-precision long double (x87 or IEEE quad)");

namespace curvis {

namespace {

const long double kPiL = 3.14159265358979323846264338327950288L;

long double shape_f(long double x) { return x * atanl(x) - 0.5L * log1pl(x * x); }
long double shape_g(long double x) { return (2.0L / kPiL) * atanl(x); }

// Monomial coefficients (in tau = t / w, |tau| <= 1) of the degree-(N-1) interpolant of f through
// the N Chebyshev nodes of [c - w, c + w]: Chebyshev coefficients by the discrete cosine sums,
// then the T_j -> monomial recurrence.  F: any callable long double -> long double.
template <int N, class F>
void fit(F f, long double c, long double w, long double mono[N]) {
    static long double node[N], basis[N][N];
    static bool ready = false;
    if (!ready) {      // (single-threaded callers: context creation / first launch of a metric)
        for (int k = 0; k < N; ++k) {
            node[k] = cosl(kPiL * (2 * k + 1) / (2 * N));
            for (int j = 0; j < N; ++j) basis[j][k] = cosl(j * kPiL * (2 * k + 1) / (2 * N));
        }
        ready = true;
    }
    long double fv[N], cj[N];
    for (int k = 0; k < N; ++k) fv[k] = f(c + w * node[k]);
    for (int j = 0; j < N; ++j) {
        long double s = 0.0L;
        for (int k = 0; k < N; ++k) s += fv[k] * basis[j][k];
        cj[j] = s * 2.0L / N;
    }
    cj[0] *= 0.5L;
    long double T[N][N] = {};
    T[0][0] = 1.0L;
    if (N > 1) T[1][1] = 1.0L;
    for (int j = 2; j < N; ++j)
        for (int k = 0; k < N; ++k) T[j][k] = (k > 0 ? 2.0L * T[j - 1][k - 1] : 0.0L) - T[j - 2][k];
    for (int k = 0; k < N; ++k) {
        long double s = 0.0L;
        for (int j = 0; j < N; ++j) s += cj[j] * T[j][k];
        mono[k] = s;
    }
}

// One table: `per_binade_log2` intervals per binade, degree N-1, N coefficients of F then N of G per interval.
template <int N, class Real>
void build(Real* out, int per_binade_log2) {
    const int per_binade = 1 << per_binade_log2;
    size_t idx = 0;
    for (int e = kShapeTabEmin; e < kShapeTabEmax; ++e) {
        const long double x0 = ldexpl(1.0L, e);
        const int wexp = e - per_binade_log2 - 1;
        const long double w = ldexpl(1.0L, wexp);                      // half width: a power of two
        for (int j = 0; j < per_binade; ++j, ++idx) {
            const long double c = x0 + (2 * j + 1) * w;
            long double mf[N], mg[N];
            fit<N>(shape_f, c, w, mf);
            fit<N>(shape_g, c, w, mg);
            Real* o = out + idx * 2 * N;
            for (int k = 0; k < N; ++k) {                              // tau^k = t^k / w^k, exact scaling
                o[k] = (Real)ldexpl(mf[k], -k * wexp);
                o[N + k] = (Real)ldexpl(mg[k], -k * wexp);
            }
        }
    }
}

}  // namespace

void build_interstellar_shape_table(double* out) { build<kShapeTabDegree + 1, double>(out, kShapeTabK); }

void build_interstellar_inverse_table(double rho, double m, double* out) {
    constexpr int N = kShapeTabDegree + 1;
    const int per_binade = 1 << kInvTabK;
    const long double rho_l = rho, m_l = m;
    const long double xscale = 2.0L / (kPiL * m_l);                 // x = 2 (|l| - a) / (pi m), metrics.rs:463
    auto inv_r2 = [=](long double z) { const long double r = rho_l + m_l * shape_f(z * xscale); return 1.0L / (r * r); };
    auto force = [=](long double z) { const long double r = rho_l + m_l * shape_f(z * xscale); return shape_g(z * xscale) / (r * r * r); };
    size_t idx = 0;
    for (int e = kInvTabEmin; e < kInvTabEmax; ++e) {
        const long double z0 = ldexpl(1.0L, e);
        const int wexp = e - kInvTabK - 1;
        const long double w = ldexpl(1.0L, wexp);
        for (int j = 0; j < per_binade; ++j, ++idx) {
            const long double c = z0 + (2 * j + 1) * w;
            long double mu[N], mh[N];
            fit<N>(inv_r2, c, w, mu);
            fit<N>(force, c, w, mh);
            // The kernel measures t from the interval's LOWER edge (z with its low mantissa bits cleared: one mask instead of
            // mask + midpoint bit), so the polynomials are re-expanded about tau = -1: tau = sigma - 1, sigma in [0, 2].
            long double su[N], sh[N];
            for (int jj = 0; jj < N; ++jj) {
                long double au = 0.0L, ah = 0.0L, binom = 1.0L;          // binom = C(k, jj)
                for (int k = jj; k < N; ++k) {
                    const long double sign = ((k - jj) & 1) ? -1.0L : 1.0L;
                    au += mu[k] * binom * sign;
                    ah += mh[k] * binom * sign;
                    binom = binom * (k + 1) / (k + 1 - jj);
                }
                su[jj] = au; sh[jj] = ah;
            }
            double* o = out + idx * 2 * N;
            for (int k = 0; k < N; ++k) {
                o[k] = (double)ldexpl(su[k], -k * wexp);
                o[N + k] = (double)ldexpl(sh[k], -k * wexp);
            }
        }
    }
    double* o = out + kInvTabConstRow * 2 * N;     // the plateau: r = rho, r' = 0 (metrics.rs:470, :482)
    for (int k = 0; k < 2 * N; ++k) o[k] = 0.0;
    o[0] = (double)(1.0L / (rho_l * rho_l));
    for (int k = 0; k < 2 * N; ++k) out[kInvTabSelfRow * 2 * N + k] = 0.0;   // the device address goes here after the upload
}

double interstellar_table_l_limit(double, double a) {
    return a + ldexp(1.0, kInvTabEmax) * (1.0 - 0x1p-20);
}

// One function on 2^kShapeTabK intervals per binade of [2^emin, 2^emax): degree-5 coefficients in t = x - midpoint.
template <class F>
static void build_one(F f, int emin, int emax, double* out) {
    constexpr int N = kShapeTabDegree + 1;
    size_t idx = 0;
    for (int e = emin; e < emax; ++e) {
        const long double x0 = ldexpl(1.0L, e);
        const int wexp = e - kShapeTabK - 1;
        const long double w = ldexpl(1.0L, wexp);
        for (int j = 0; j < (1 << kShapeTabK); ++j, ++idx) {
            long double mono[N];
            fit<N>(f, x0 + (2 * j + 1) * w, w, mono);
            for (int k = 0; k < N; ++k) out[idx * N + k] = (double)ldexpl(mono[k], -k * wexp);
        }
    }
}

void build_atan_table(double* out) { build_one([](long double x) { return atanl(x); }, kShapeTabEmin, kShapeTabEmax, out); }
void build_log_table(double* out) { build_one([](long double y) { return logl(y); }, 0, kLogTabEmax, out); }

void build_interstellar_shape_table_f32(float* out) { build<kShapeTab32Degree + 1, float>(out, kShapeTab32K); }

}  // namespace curvis

// Test hook: the per-metric table evaluated on the host exactly as FastInterstellar::prepare evaluates it (z = |l| - a is the
// caller's business: this takes z).  y[i] = 1/(rho + m F(x))^2, g[i] = (2/pi) atan x / (rho + m F(x))^3 at x = 2 z / (pi m);
// z below 2^kInvTabEmin (or <= 0) reads the constant row.  Returns 1 when every z was below 2^kInvTabEmax.
extern "C" int curvis_debug_inverse_table_host(double rho, double m, const double* z, double* y, double* g, size_t n) {
    using namespace curvis;
    double* table = new double[kInvTabIntervals * kShapeTabDoubles];
    build_interstellar_inverse_table(rho, m, table);
    int all = 1;
    for (size_t i = 0; i < n; ++i) {
        unsigned long long bits;
        __builtin_memcpy(&bits, &z[i], 8);
        const unsigned hi = (unsigned)(bits >> 32);
        unsigned idx = (hi >> kInvTabShift) - kInvTabBase;
        if (z[i] >= ldexp(1.0, kInvTabEmax)) { y[i] = g[i] = NAN; all = 0; continue; }
        if (idx > (unsigned)kInvTabConstRow) idx = (unsigned)kInvTabConstRow;
        const unsigned long long cbits = (unsigned long long)(hi & ~((1u << kInvTabShift) - 1u)) << 32;     // the interval's lower edge
        double c;
        __builtin_memcpy(&c, &cbits, 8);
        const double t = z[i] - c;
        const double* a = table + (size_t)idx * kShapeTabDoubles;
        double U = a[kShapeTabDegree], H = a[kShapeTabDoubles - 1];
        for (int k = kShapeTabDegree - 1; k >= 0; --k) {
            U = fma(t, U, a[k]);
            H = fma(t, H, a[kShapeTabDegree + 1 + k]);
        }
        y[i] = U; g[i] = H;
    }
    delete[] table;
    return all;
}

// Test hook: the atan (which = 0) / ln (which = 1) tables of the operation-for-operation kernel evaluated on the host exactly as
// the kernel evaluates them.  Returns 1 when every argument was inside the table's range (NaN entries otherwise).
extern "C" int curvis_debug_fn_table_host(int which, const double* x, double* out, size_t n) {
    using namespace curvis;
    static double* tabs[2] = {nullptr, nullptr};
    if (!tabs[0]) {
        double* a = new double[kAtanTabIntervals * kFnTabDoubles];
        double* l = new double[kLogTabIntervals * kFnTabDoubles];
        build_atan_table(a); build_log_table(l);
        tabs[1] = l; tabs[0] = a;
    }
    const unsigned base = which ? kLogTabBase : kShapeTabBase;
    const size_t count = which ? kLogTabIntervals : kAtanTabIntervals;
    int all = 1;
    for (size_t i = 0; i < n; ++i) {
        unsigned long long bits;
        __builtin_memcpy(&bits, &x[i], 8);
        const unsigned hi = (unsigned)(bits >> 32);
        const unsigned idx = (hi >> kShapeTabShift) - base;
        if (idx >= (unsigned)count) { out[i] = NAN; all = 0; continue; }
        const unsigned long long cbits = (unsigned long long)((hi & ~((1u << kShapeTabShift) - 1u)) | (1u << (kShapeTabShift - 1))) << 32;
        double c;
        __builtin_memcpy(&c, &cbits, 8);
        const double t = x[i] - c;
        const double* a = tabs[which ? 1 : 0] + (size_t)idx * kFnTabDoubles;
        double v = a[kShapeTabDegree];
        for (int k = kShapeTabDegree - 1; k >= 0; --k) v = fma(t, v, a[k]);
        out[i] = v;
    }
    return all;
}

// Test hook (include/curvis_gpu.h): the host-built table evaluated on the host exactly as the kernel
// evaluates it (index from the high word, exact t, two fma Horner chains).  No GPU involved; it lets
// the CPU test-suite check the generator.  Returns 0 when x is outside the table's range.
extern "C" int curvis_debug_shape_table_host(const double* x, double* f, double* g, size_t n) {
    using namespace curvis;
    static double* table = nullptr;
    if (!table) {
        double* t = new double[kShapeTabIntervals * kShapeTabDoubles];
        build_interstellar_shape_table(t);
        table = t;
    }
    int all = 1;
    for (size_t i = 0; i < n; ++i) {
        unsigned long long bits;
        __builtin_memcpy(&bits, &x[i], 8);
        const unsigned hi = (unsigned)(bits >> 32);
        const unsigned idx = (hi >> kShapeTabShift) - kShapeTabBase;
        if (idx >= (unsigned)kShapeTabIntervals) { f[i] = g[i] = NAN; all = 0; continue; }
        const unsigned long long cbits = (unsigned long long)((hi & ~((1u << kShapeTabShift) - 1u)) | (1u << (kShapeTabShift - 1))) << 32;
        double c;
        __builtin_memcpy(&c, &cbits, 8);
        const double t = x[i] - c;
        const double* a = table + (size_t)idx * kShapeTabDoubles;
        double F = a[kShapeTabDegree], G = a[kShapeTabDoubles - 1];
        for (int k = kShapeTabDegree - 1; k >= 0; --k) {
            F = fma(t, F, a[k]);
            G = fma(t, G, a[kShapeTabDegree + 1 + k]);
        }
        f[i] = F; g[i] = G;
    }
    return all;
}

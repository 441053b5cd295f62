// Inverse of VisualSFM's one-parameter radial model, as Image::undistort applies it (/root/reference/src/hpmvs/Image.cpp:68-149): for
// a target pixel at normalised position (x, y) return the normalised position (mx, my) of its SOURCE in the distorted image.
// The source radius solves a depressed cubic; the reference writes out Cardano's formula in double (k1 > 0: one real root) and in
// complex<double> (k1 < 0: the root of the three that continues the identity).  The operations and their order below are the
// reference's (so the host build, with the same libm, returns the same bits); the quantities are named for what they are.
// One header for both sides: the host entry point hpmvs_undistort_rgb (host_io.cpp, std::complex + libm) and the device kernel
// hp::undistort_kernel (patch_kernels.cuh), which uses the small complex helpers below with CUDA's double-precision libm.
#pragma once

#include <math.h>

#if defined(__CUDACC__)
#define UD_HD __host__ __device__ __forceinline__
#else
#define UD_HD inline
#include <complex>
#endif

namespace ud {

struct Source { float mx, my; };

// k1 > 0.  yy = y^2, kr = k1 * r^2 (r^2 = x^2 + y^2).  With p = yy / kr the cubic in the source ordinate m is m^3 + p*m - p*y = 0;
// Cardano: m = c - p / (3c), c = cbrt(sqrt(p^3/4 * ... ) + p*y/2) - the reference keeps 1/kr as a factor instead of forming p.
UD_HD Source source_positive_k1(float x, float y, double kr) {
    const double yy = y * y;
    const double y6 = yy * yy * yy;
    const double inv_kr = 1.0 / kr;
    const double q = y6 / (kr * kr);
    const double disc_root = sqrt(q * (0.25 + inv_kr / 27.0));
    const double half_py = yy * inv_kr * y * 0.5;
    const double c = pow(disc_root + half_py, 1.0 / 3.0);
    const double m = c - yy * inv_kr / (c * 3.0);
    Source s;
    s.mx = m * x / y;
    s.my = m;
    return s;
}

#if !defined(__CUDACC__)
// k1 < 0 on the host: the discriminant may be negative, so the reference switches to complex arithmetic and takes the conjugate-pair
// combination -(c + w)/2 + p/(6c) with w = (c + p/(3c)) * i*sqrt(3)  (the root that tends to y for k1 -> 0).
inline Source source_negative_k1_host(float x, float y, double kr) {
    typedef std::complex<double> cd;
    const double yy = y * y;
    const double y6 = yy * yy * yy;
    const double quarter = y6 / (kr * kr * 4.0);
    const double cube = y6 / (kr * kr * kr * 27.0);
    const cd disc = quarter + cube;
    const cd disc_root = sqrt(disc);
    const double p = yy / kr;
    const double half_py = p * y * 0.5;
    const cd radicand = disc_root + half_py;
    const cd c = pow(radicand, 1.0 / 3.0);
    const cd w = (c + p / (c * 3.0)) * cd(0.0, sqrt(3.0));
    const cd m = -0.5 * (c + w) + p / (c * 6.0);
    Source s;
    s.mx = m.real() * x / y;
    s.my = m.real();
    return s;
}
#endif

// the same on the device: complex sqrt / cube root in polar form with CUDA's libm (agrees with the host to a few ulp of double, i.e.
// the f32 source position is the same except for rare last-place roundings; tests/test_next_rows.py states the bar)
struct C2 { double re, im; };
UD_HD C2 c_add(C2 a, C2 b) { return C2{a.re + b.re, a.im + b.im}; }
UD_HD C2 c_mul(C2 a, C2 b) { return C2{a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
UD_HD C2 c_div_real_by(double p, C2 z) { const double n = z.re * z.re + z.im * z.im; return C2{p * z.re / n, -p * z.im / n}; }
UD_HD C2 c_sqrt_real(double v) { return v >= 0.0 ? C2{sqrt(v), 0.0} : C2{0.0, sqrt(-v)}; }
UD_HD C2 c_cbrt(C2 z) {
    const double r = hypot(z.re, z.im), th = atan2(z.im, z.re);
    const double rr = pow(r, 1.0 / 3.0);
    double sn, cs;
    sincos(th / 3.0, &sn, &cs);
    return C2{rr * cs, rr * sn};
}
UD_HD Source source_negative_k1_polar(float x, float y, double kr) {
    const double yy = y * y;
    const double y6 = yy * yy * yy;
    const double quarter = y6 / (kr * kr * 4.0);
    const double cube = y6 / (kr * kr * kr * 27.0);
    const C2 disc_root = c_sqrt_real(quarter + cube);
    const double p = yy / kr;
    const double half_py = p * y * 0.5;
    const C2 c = c_cbrt(C2{disc_root.re + half_py, disc_root.im});
    const C2 p3c = c_div_real_by(p, C2{c.re * 3.0, c.im * 3.0});
    const C2 w = c_mul(c_add(c, p3c), C2{0.0, sqrt(3.0)});
    const C2 p6c = c_div_real_by(p, C2{c.re * 6.0, c.im * 6.0});
    const double m_re = -0.5 * (c.re + w.re) + p6c.re;
    Source s;
    s.mx = m_re * x / y;
    s.my = m_re;
    return s;
}

}  // namespace ud

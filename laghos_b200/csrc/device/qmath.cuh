// Device 2x2 / 3x3 dense helpers for the quadrature-point physics (QUpdate).
//
// Same algorithms as MFEM's linalg/kernels.hpp routines that the reference's
// QUpdateBody calls (reference laghos_solver.cpp:1078-1158): determinant and
// adjugate inverse, scaled closed-form symmetric eigen-solver (trigonometric root
// of the characteristic polynomial that is best separated, deflation by its
// eigenvector, Parlett's 2x2 rotation), smallest singular value through the
// eigenvalues of J^t J.  Divisions by 3, 6 and by the power-of-two scaling factor are
// written as multiplications (<= 1 ulp from MFEM's quotients, far inside the parity
// tolerance): an fp64 division costs ~20 instructions on the GPU and QUpdate is
// instruction-bound.  QUpdateBody only consumes the smallest eigenvalue and its
// eigenvector (compr_dir, laghos_solver.cpp:1113-1124), so only that pair is
// returned.  Written with scalars and selects (no dynamically indexed arrays) so
// everything stays in registers.  Column-major like MFEM.
#pragma once
#include <cuda_runtime.h>
#include <cfloat>
#include <cmath>
#include <cstring>

namespace lagb {
namespace qm {

#define LAGB_HD __host__ __device__ __forceinline__

LAGB_HD double scaling_factor(double d_max)
{
   if (d_max > 0.0)
   {
      int e; double m = frexp(d_max, &e);
      if (e == DBL_MAX_EXP) { m *= 2.0; }
      return d_max/m;
   }
   return 1.0;
}

// The same power of two 2^e (d_max = m 2^e, m in [0.5,1)) built from the exponent bits,
// together with its exact inverse: x/mult == x*inv bit for bit, and a multiplication costs a
// tenth of an fp64 division on the GPU.  Zero / subnormal / huge d_max (never reached by
// physical Jacobians) fall back to the generic routine.
LAGB_HD void scaling_factor_inv(double d_max, double &mult, double &inv)
{
#ifdef __CUDA_ARCH__
   const long long bits = __double_as_longlong(d_max);
   const long long be = (bits >> 52) & 0x7ff;            // biased exponent of d_max (d_max >= 0)
   if (be >= 2 && be <= 2040)
   {
      mult = __longlong_as_double((be + 1) << 52);       // 2^(be+1-1023) = 2^e
      inv = __longlong_as_double((2046 - (be + 1)) << 52);
      return;
   }
#endif
   mult = scaling_factor(d_max);
   inv = 1.0/mult;
}

// Parlett: rotation diagonalising [d1 d12; d12 d2]; eigenvectors (c,-s), (s,c)
LAGB_HD void eigensystem2s(const double d12, double &d1, double &d2, double &c, double &s)
{
   const double sqrt_1_eps = 67108864.0; // sqrt(1/DBL_EPSILON) = 2^26
   if (d12 == 0.0) { c = 1.0; s = 0.0; return; }
   double t;
   const double zeta = (d2 - d1)/(2*d12);
   const double azeta = fabs(zeta);
   if (azeta < sqrt_1_eps) { t = copysign(1.0/(azeta + sqrt(1.0 + zeta*zeta)), zeta); }
   else { t = copysign(0.5/azeta, zeta); }
#ifdef __CUDA_ARCH__
   c = rsqrt(1.0 + t*t);
#else
   c = sqrt(1.0/(1.0 + t*t));
#endif
   s = c*t;
   t *= d12;
   d1 -= t;
   d2 += t;
}

// ---------------------------------------------------------------------------------------------
// Cheap exact-enough primitives for the instruction-bound point physics (device only; the host
// versions are the plain operators).  ncu of the first QUpdate kernel: 28 % of all executed
// instructions were the libm acos/cos pairs of the cubic root, 7 % fmax(fabs()) chains, ~10 %
// IEEE divisions and square roots with their slow-path calls.
// ---------------------------------------------------------------------------------------------
// 1/x, 1/sqrt(x), sqrt(x), a/b: hardware seed (MUFU.RCP64H / RSQ64H, 2^-22) + two Newton steps,
// <= 1 ulp, no subnormal / special-case branches (arguments are O(1) physical quantities)
LAGB_HD double frcp(double x)
{
#ifdef __CUDA_ARCH__
   double r;
   asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
   double e = fma(-x, r, 1.0); r = fma(r, e, r);
   e = fma(-x, r, 1.0); r = fma(r, e, r);
   return r;
#else
   return 1.0/x;
#endif
}
LAGB_HD double fdiv(double a, double b)
{
#ifdef __CUDA_ARCH__
   const double r = frcp(b);
   const double q = a*r;
   return fma(fma(-b, q, a), r, q);
#else
   return a/b;
#endif
}
LAGB_HD double frsqrt(double x)   // x > 0
{
#ifdef __CUDA_ARCH__
   double y;
   asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
   const double hx = 0.5*x;
   double e = fma(-hx*y, y, 0.5); y = fma(y, e, y);
   e = fma(-hx*y, y, 0.5); y = fma(y, e, y);
   return y;
#else
   return 1.0/sqrt(x);
#endif
}
LAGB_HD double fsqrt(double x)    // x >= 0
{
#ifdef __CUDA_ARCH__
   const double y = frsqrt(x);
   double s = x*y;
   s = fma(fma(-s, s, x), 0.5*y, s);
   return (x > 0.0) ? s : 0.0;
#else
   return sqrt(x);
#endif
}

// high word of |x|: orders like |x| itself up to the low 32 mantissa bits, which is all the
// power-of-two scaling needs (max over the entries of the biased exponent)
LAGB_HD int imax(int a, int b)
{
#ifdef __CUDA_ARCH__
   return max(a, b);      // one VIMNMX instead of ISETP + SEL
#else
   return a > b ? a : b;
#endif
}
LAGB_HD int abs_hi(double x)
{
#ifdef __CUDA_ARCH__
   return __double2hiint(x) & 0x7fffffff;
#else
   long long b; memcpy(&b, &x, 8);
   return (int)((b >> 32) & 0x7fffffff);
#endif
}
// mult = 2^e with max|x| = m 2^e, m in [0.5,1), from the largest high word; false when the exponent
// is outside the range the bit construction covers (zero / subnormal / huge: caller falls back)
LAGB_HD bool scaling_from_hi(int hmax, double &mult, double &inv)
{
   const long long be = hmax >> 20;
   if (be < 2 || be > 2040) { return false; }
#ifdef __CUDA_ARCH__
   mult = __longlong_as_double((be + 1) << 52);
   inv = __longlong_as_double((2046 - (be + 1)) << 52);
#else
   mult = ldexp(1.0, (int)be + 1 - 1023);
   inv = ldexp(1.0, 1023 - (int)be - 1);
#endif
   return true;
}
// max(|a|,|b|) of doubles through their bit patterns (non-negative doubles order like integers):
// 64-bit integer compare/select instead of the DSETP/FSEL/NaN sequence of fmax(fabs(), fabs())
LAGB_HD double max_abs2(double a, double b)
{
#ifdef __CUDA_ARCH__
   const unsigned long long m = 0x7fffffffffffffffull;
   const unsigned long long ua = (unsigned long long)__double_as_longlong(a) & m;
   const unsigned long long ub = (unsigned long long)__double_as_longlong(b) & m;
   return __longlong_as_double((long long)(ua > ub ? ua : ub));
#else
   return fmax(fabs(a), fabs(b));
#endif
}

// cos(acos(x)/3) for x in [-0.9, 1]: the root t in [0.62, 1] of 4 t^3 - 3 t = x.  Degree-6 seed
// (2e-3) and three Newton steps with a single-precision reciprocal of the slope (the slope error
// only enters the convergence factor): 0.94 ulp against 1.09 ulp of cos(acos(x)/3) in libm, at
// ~30 instructions instead of ~200.
LAGB_HD double cos_acos_third(double x)
{
   double t = -0.02875468921234147;
   t = fma(t, x, 0.037923394343120996);
   t = fma(t, x, -0.005179232483301509);
   t = fma(t, x, 0.010631209065749747);
   t = fma(t, x, -0.04939436622952864);
   t = fma(t, x, 0.16822898759893845);
   t = fma(t, x, 0.8660421155284468);
   for (int it = 0; it < 3; it++)   // constant trip count: unrolled
   {
      const double t2 = t*t;
      const double g = fma(t, fma(4.0, t2, -3.0), -x);
      const double gp = fma(12.0, t2, -3.0);
#ifdef __CUDA_ARCH__
      float rf;
      asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rf) : "f"((float)gp));
      t = fma(-g, (double)rf, t);
#else
      t = fma(-g, (double)(1.0f/(float)gp), t);
#endif
   }
   return t;
}

// unit vector in the near-kernel of the symmetric [c1 d12 d13; d12 c2 d23; d13 d23 c3]
LAGB_HD bool kernel_vector3s(double c1, double c2, double c3, double d12, double d13, double d23,
                             double &z0, double &z1, double &z2)
{
   // cross products of the row pairs (r0,r1), (r0,r2), (r1,r2)
   const double a0 = d12*d23 - d13*c2,  a1 = d13*d12 - c1*d23,  a2 = c1*c2 - d12*d12;
   const double b0 = d12*c3 - d13*d23,  b1 = d13*d13 - c1*c3,   b2 = c1*d23 - d12*d13;
   const double e0 = c2*c3 - d23*d23,   e1 = d23*d13 - d12*c3,  e2 = d12*d23 - c2*d13;
   const double na = a0*a0 + a1*a1 + a2*a2;
   const double nb = b0*b0 + b1*b1 + b2*b2;
   const double ne = e0*e0 + e1*e1 + e2*e2;
   double n = na; z0 = a0; z1 = a1; z2 = a2;
   if (nb > n) { n = nb; z0 = b0; z1 = b1; z2 = b2; }
   if (ne > n) { n = ne; z0 = e0; z1 = e1; z2 = e2; }
   const double amax = max_abs2(max_abs2(max_abs2(c1, c2), max_abs2(c3, d12)), max_abs2(d13, d23));
   if (!(n > 1e-28*amax*amax*amax*amax) || n == 0.0) { return false; }
   const double inv = frsqrt(n);
   z0 *= inv; z1 *= inv; z2 *= inv;
   return true;
}

// Householder completion: u, w orthonormal and orthogonal to the unit vector z
LAGB_HD void complete_basis(double z0, double z1, double z2,
                            double &u0, double &u1, double &u2, double &w0, double &w1, double &w2)
{
   int k = 0; double zk = z0;
   if (fabs(z1) < fabs(zk)) { k = 1; zk = z1; }
   if (fabs(z2) < fabs(zk)) { k = 2; zk = z2; }
   const double sgn = (zk >= 0.) ? 1.0 : -1.0;
   const double v0 = z0 + (k == 0 ? sgn : 0.0);
   const double v1 = z1 + (k == 1 ? sgn : 0.0);
   const double v2 = z2 + (k == 2 ? sgn : 0.0);
   const double vn2 = v0*v0 + v1*v1 + v2*v2;
   // i1 = (k+1)%3, i2 = (k+2)%3
   const double vi1 = (k == 0) ? v1 : (k == 1) ? v2 : v0;
   const double vi2 = (k == 0) ? v2 : (k == 1) ? v0 : v1;
   const int i1 = (k + 1) % 3, i2 = (k + 2) % 3;
   const double tw = 2.0*frcp(vn2);           // one reciprocal for both columns
   const double f1 = vi1*tw, f2 = vi2*tw;
   u0 = ((i1 == 0) ? 1.0 : 0.0) - v0*f1;
   u1 = ((i1 == 1) ? 1.0 : 0.0) - v1*f1;
   u2 = ((i1 == 2) ? 1.0 : 0.0) - v2*f1;
   w0 = ((i2 == 0) ? 1.0 : 0.0) - v0*f2;
   w1 = ((i2 == 1) ? 1.0 : 0.0) - v1*f2;
   w2 = ((i2 == 2) ? 1.0 : 0.0) - v2*f2;
}

// Parlett's rotation with one reciprocal and one square root: with delta = (d2 - d1)/2,
// h = |delta| + sqrt(d12^2 + delta^2):  t = sign(zeta) |d12| / h  (= 1/(|zeta| + sqrt(1 + zeta^2))),
// t d12 = sign(delta) d12^2 / h.  Same values as eigensystem2s up to round-off.
LAGB_HD void eigensystem2s_fast(const double d12, double &d1, double &d2, double &c, double &s)
{
   if (d12 == 0.0) { c = 1.0; s = 0.0; return; }
   const double delta = 0.5*(d2 - d1);
   const double ih = frcp(fabs(delta) + fsqrt(fma(d12, d12, delta*delta)));
   const double sz = ((delta < 0.0) != (d12 < 0.0)) ? -1.0 : 1.0;      // sign(zeta), zeta = delta/d12
   const double t = sz*fabs(d12)*ih;
   c = frsqrt(fma(t, t, 1.0));
   s = c*t;
   const double td = t*d12;
   d1 -= td;
   d2 += td;
}

// smallest eigenvalue and its unit eigenvector of the symmetric 2x2 (a d; d b)
LAGB_HD void min_eig2(double a, double d, double b, double &lmin, double &x0, double &x1)
{
   double c, s;
   eigensystem2s(d, a, b, c, s);
   if (a <= b) { lmin = a; x0 = c; x1 = -s; }
   else        { lmin = b; x0 = s; x1 = c; }
}

// smallest eigenvalue and its unit eigenvector of the symmetric 3x3 with entries
// (d11 d12 d13; . d22 d23; . . d33).
// Same algorithm as before (scaled trigonometric root that is best separated, deflation by its
// eigenvector, Parlett rotation), restructured for the GPU: branch-free root (cos_acos_third for
// both signs of R: cos((acos(R) + 2 pi)/3) = -cos(acos(-R)/3)), and when the best separated root IS the
// smallest one (R >= 0) its Rayleigh quotient and kernel vector are the answer, so the deflation is
// skipped; the 2x2 block could only win the ordering if the three eigenvalues agreed to round-off
// (guarded by Q > 1e-20 in the scaled matrix).
LAGB_HD void min_eig3(double d11, double d12, double d13, double d22, double d23, double d33,
                      double &lmin, double &x0, double &x1, double &x2)
{
   const int hmax = imax(imax(imax(abs_hi(d11), abs_hi(d22)), imax(abs_hi(d33), abs_hi(d12))), imax(abs_hi(d13), abs_hi(d23)));
   double mult, imult;
   if (!scaling_from_hi(hmax, mult, imult))
   {
      const double d_max = fmax(fmax(fmax(fabs(d11), fabs(d22)), fmax(fabs(d33), fabs(d12))), fmax(fabs(d13), fabs(d23)));
      if (d_max == 0.0) { lmin = 0.0; x0 = 1.0; x1 = 0.0; x2 = 0.0; return; }   // zero matrix: what the general path returns
      mult = scaling_factor(d_max); imult = 1.0/mult;
   }
   d11 *= imult; d22 *= imult; d33 *= imult;
   d12 *= imult; d13 *= imult; d23 *= imult;
   double aa = (d11 + d22 + d33)*(1.0/3.0);
   double c1 = d11 - aa, c2 = d22 - aa, c3 = d33 - aa;
   const double Q = (2*(d12*d12 + d13*d13 + d23*d23) + c1*c1 + c2*c2 + c3*c3)*(1.0/6.0);
   const double R = (c1*(d23*d23 - c2*c3) + d12*(d12*c3 - 2*d13*d23) + d13*d13*c2)/2;
   bool ident = true;
   if (Q > 0.)
   {
      const double sqrtQ = fsqrt(Q);
      const double sqrtQ3 = Q*sqrtQ;
      const double aR = fabs(R);
      const double t = (aR >= sqrtQ3) ? 1.0 : cos_acos_third(fdiv(aR, sqrtQ3));
      const double r = (R < 0.) ? 2*sqrtQ*t : -2*sqrtQ*t;     // R < 0: largest root, else smallest
      aa += r;
      c1 = d11 - aa; c2 = d22 - aa; c3 = d33 - aa;
      double z0, z1, z2;
      if (kernel_vector3s(c1, c2, c3, d12, d13, d23, z0, z1, z2))
      {
         ident = false;
         const double Az0 = d11*z0 + d12*z1 + d13*z2, Az1 = d12*z0 + d22*z1 + d23*z2, Az2 = d13*z0 + d23*z1 + d33*z2;
         const double l1 = z0*Az0 + z1*Az1 + z2*Az2;
         lmin = l1; x0 = z0; x1 = z1; x2 = z2;
         if (R < 0. || !(Q > 1e-20))
         {
            double u0, u1, u2, w0, w1, w2;
            complete_basis(z0, z1, z2, u0, u1, u2, w0, w1, w2);
            const double Au0 = d11*u0 + d12*u1 + d13*u2, Au1 = d12*u0 + d22*u1 + d23*u2, Au2 = d13*u0 + d23*u1 + d33*u2;
            const double Aw0 = d11*w0 + d12*w1 + d13*w2, Aw1 = d12*w0 + d22*w1 + d23*w2, Aw2 = d13*w0 + d23*w1 + d33*w2;
            double b22 = u0*Au0 + u1*Au1 + u2*Au2;
            double b33 = w0*Aw0 + w1*Aw1 + w2*Aw2;
            const double b23 = u0*Aw0 + u1*Aw1 + u2*Aw2;
            double c, s;
            eigensystem2s_fast(b23, b22, b33, c, s);
            // stable ascending selection among (l1, b22, b33)
            if (b22 < lmin) { lmin = b22; x0 = c*u0 - s*w0; x1 = c*u1 - s*w1; x2 = c*u2 - s*w2; }
            if (b33 < lmin) { lmin = b33; x0 = s*u0 + c*w0; x1 = s*u1 + c*w1; x2 = s*u2 + c*w2; }
         }
      }
   }
   if (ident) { lmin = aa; x0 = 1.0; x1 = 0.0; x2 = 0.0; }
   lmin *= mult;
}

// smallest singular value of the 2x2 column-major (d0 d2; d1 d3)
LAGB_HD double min_sv2(double d0, double d1, double d2, double d3)
{
   const double d_max = fmax(fmax(fabs(d0), fabs(d1)), fmax(fabs(d2), fabs(d3)));
   const double mult = scaling_factor(d_max);
   d0 /= mult; d1 /= mult; d2 /= mult; d3 /= mult;
   double t = 0.5*((d0 + d2)*(d0 - d2) + (d1 - d3)*(d1 + d3));
   double s = d0*d2 + d1*d3;
   s = sqrt(0.5*(d0*d0 + d1*d1 + d2*d2 + d3*d3) + sqrt(t*t + s*s));
   if (s == 0.0) { return 0.0; }
   t = fabs(d0*d3 - d1*d2)/s;
   if (t > s) { return s*mult; }
   return t*mult;
}

// smallest singular value of the 3x3 column-major J (d0..d8): sqrt of the smallest eigenvalue of
// J^t J.  One branch-free evaluation of the cubic root serves the three cases of the reference
// algorithm (|R| <= 0.9: the smallest root directly; R > 0.9: the smallest root is the best separated
// one; R < -0.9: the largest root is, and the two small ones come from the deflated 2x2 problem,
// which is the only case that still takes the long path).
LAGB_HD double min_sv3(double d0, double d1, double d2, double d3, double d4,
                       double d5, double d6, double d7, double d8)
{
   const int hmax = imax(imax(imax(abs_hi(d0), abs_hi(d1)), imax(abs_hi(d2), abs_hi(d3))),
                         imax(imax(abs_hi(d4), abs_hi(d5)), imax(abs_hi(d6), imax(abs_hi(d7), abs_hi(d8)))));
   double mult, imult;
   if (!scaling_from_hi(hmax, mult, imult))
   {
      const double d_max = fmax(fmax(fmax(fabs(d0), fabs(d1)), fmax(fabs(d2), fabs(d3))),
                                fmax(fmax(fabs(d4), fabs(d5)), fmax(fabs(d6), fmax(fabs(d7), fabs(d8)))));
      mult = scaling_factor(d_max); imult = 1.0/mult;
   }
   d0 *= imult; d1 *= imult; d2 *= imult; d3 *= imult; d4 *= imult;
   d5 *= imult; d6 *= imult; d7 *= imult; d8 *= imult;
   const double b11 = d0*d0 + d1*d1 + d2*d2;
   const double b12 = d0*d3 + d1*d4 + d2*d5;
   const double b13 = d0*d6 + d1*d7 + d2*d8;
   const double b22 = d3*d3 + d4*d4 + d5*d5;
   const double b23 = d3*d6 + d4*d7 + d5*d8;
   const double b33 = d6*d6 + d7*d7 + d8*d8;
   double aa = (b11 + b22 + b33)*(1.0/3.0);
   const double b11_b22 = ((d0 - d3)*(d0 + d3) + (d1 - d4)*(d1 + d4) + (d2 - d5)*(d2 + d5));
   const double b22_b33 = ((d3 - d6)*(d3 + d6) + (d4 - d7)*(d4 + d7) + (d5 - d8)*(d5 + d8));
   const double b33_b11 = ((d6 - d0)*(d6 + d0) + (d7 - d1)*(d7 + d1) + (d8 - d2)*(d8 + d2));
   const double c1 = (b11_b22 - b33_b11)*(1.0/3.0);
   const double c2 = (b22_b33 - b11_b22)*(1.0/3.0);
   const double c3 = (b33_b11 - b22_b33)*(1.0/3.0);
   const double Q = (2*(b12*b12 + b13*b13 + b23*b23) + c1*c1 + c2*c2 + c3*c3)*(1.0/6.0);
   const double R = (c1*(b23*b23 - c2*c3) + b12*(b12*c3 - 2*b13*b23) + b13*b13*c2)/2;
   if (Q > 0.)
   {
      const double sqrtQ = fsqrt(Q);
      const double sqrtQ3 = Q*sqrtQ;
      const bool big = fabs(R) >= sqrtQ3;
      const double Rn = big ? ((R < 0.) ? -1.0 : 1.0) : fdiv(R, sqrtQ3);
      const bool defl = big || Rn < -0.9;                      // the small pair is resolved by deflation
      const double t = big ? 1.0 : cos_acos_third((Rn < -0.9) ? -Rn : Rn);
      if (!defl) { aa -= 2*sqrtQ*t; }
      else
      {
         const double l1 = aa + ((R < 0.) ? 2*sqrtQ*t : -2*sqrtQ*t);
         double z0, z1, z2;
         if (!kernel_vector3s(b11 - l1, b22 - l1, b33 - l1, b12, b13, b23, z0, z1, z2)) { aa = l1; }
         else
         {
            double u0, u1, u2, w0, w1, w2;
            complete_basis(z0, z1, z2, u0, u1, u2, w0, w1, w2);
            const double Bz0 = b11*z0 + b12*z1 + b13*z2, Bz1 = b12*z0 + b22*z1 + b23*z2, Bz2 = b13*z0 + b23*z1 + b33*z2;
            const double Bu0 = b11*u0 + b12*u1 + b13*u2, Bu1 = b12*u0 + b22*u1 + b23*u2, Bu2 = b13*u0 + b23*u1 + b33*u2;
            const double Bw0 = b11*w0 + b12*w1 + b13*w2, Bw1 = b12*w0 + b22*w1 + b23*w2, Bw2 = b13*w0 + b23*w1 + b33*w2;
            const double e1 = z0*Bz0 + z1*Bz1 + z2*Bz2;
            double e2 = u0*Bu0 + u1*Bu1 + u2*Bu2;
            double e3 = w0*Bw0 + w1*Bw1 + w2*Bw2;
            const double e23 = u0*Bw0 + u1*Bw1 + u2*Bw2;
            double c, s;
            eigensystem2s_fast(e23, e2, e3, c, s);
            aa = fmin(fmin(e1, e2), e3);
         }
      }
   }
   return fsqrt(fabs(aa))*mult;
}

// MFEM kernels::Norml2 (scaled 2-norm), sizes 2 and 3
LAGB_HD double norml2_accum(double &scale, double &sum, double v)
{
   if (v != 0.0)
   {
      const double a = fabs(v);
      if (scale <= a) { const double q = scale/a; sum = 1.0 + sum*(q*q); scale = a; }
      else { const double q = a/scale; sum += q*q; }
   }
   return 0.0;
}
// On the device the 2-norm is taken directly: the arguments are element lengths and unit
// vectors (|x| in [1e-150, 1e150] cannot over/underflow when squared), and the scaled
// recurrence above costs three fp64 divisions per call.  Agrees with it to ~1 ulp.
LAGB_HD double norml2_2(double a, double b)
{
#ifdef __CUDA_ARCH__
   return sqrt(a*a + b*b);
#else
   double scale = 0.0, sum = 0.0;
   norml2_accum(scale, sum, a); norml2_accum(scale, sum, b);
   return scale*sqrt(sum);
#endif
}
LAGB_HD double norml2_3(double a, double b, double c)
{
#ifdef __CUDA_ARCH__
   return fsqrt(a*a + b*b + c*c);
#else
   double scale = 0.0, sum = 0.0;
   norml2_accum(scale, sum, a); norml2_accum(scale, sum, b); norml2_accum(scale, sum, c);
   return scale*sqrt(sum);
#endif
}

LAGB_HD double smooth_step_01(double x, double eps)
{
   const double y = (x + eps)*(0.5/eps);   // eps is a literal at the call site: folded at compile time
   if (y < 0.0) { return 0.0; }
   if (y > 1.0) { return 1.0; }
   return (3.0 - 2.0*y)*y*y;
}

} // namespace qm
} // namespace lagb

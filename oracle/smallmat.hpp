// TEST INFRASTRUCTURE — CPU oracle, not a product path (see oracle/README.md).
//
// Restatement of the small dense-matrix helpers that the reference's QUpdateBody
// calls from MFEM's linalg/kernels.hpp (reference laghos_solver.cpp:1078-1158:
// kernels::Det, CalcInverse, Mult, Symmetrize, CalcEigenvalues, Add, Norml2,
// CalcSingularvalue, MultABt).  MFEM is NOT in /root/reference (un-vendored
// dependency, GitHub master, unpinned: README.md:137-138), so these follow MFEM's
// published algorithm (scaled closed-form trigonometric eigen-solver with
// deflation by a near-kernel vector, Parlett's 2x2 symmetric eigensystem) and are
// pinned end-to-end by the reference's --checks table (laghos.cpp:1441-1463),
// see tests/test_oracle_golden.py.
// All matrices are column-major, as in MFEM.
#pragma once
#include <cmath>
#include <cfloat>
#include <algorithm>

namespace oracle {
namespace sm {

template<int DIM> inline double Det(const double *d);
template<> inline double Det<2>(const double *d) { return d[0]*d[3] - d[1]*d[2]; }
template<> inline double Det<3>(const double *d)
{
   return d[0]*(d[4]*d[8] - d[5]*d[7]) +
          d[3]*(d[2]*d[7] - d[1]*d[8]) +
          d[6]*(d[1]*d[5] - d[2]*d[4]);
}

template<int DIM> inline void CalcInverse(const double *d, double *inv);
template<> inline void CalcInverse<2>(const double *d, double *inv)
{
   const double t = 1.0/Det<2>(d);
   inv[0] =  d[3]*t;
   inv[1] = -d[1]*t;
   inv[2] = -d[2]*t;
   inv[3] =  d[0]*t;
}
template<> inline void CalcInverse<3>(const double *d, double *inv)
{
   const double t = 1.0/Det<3>(d);
   inv[0] = (d[4]*d[8] - d[5]*d[7])*t;
   inv[1] = (d[7]*d[2] - d[8]*d[1])*t;
   inv[2] = (d[1]*d[5] - d[2]*d[4])*t;
   inv[3] = (d[5]*d[6] - d[3]*d[8])*t;
   inv[4] = (d[8]*d[0] - d[6]*d[2])*t;
   inv[5] = (d[2]*d[3] - d[0]*d[5])*t;
   inv[6] = (d[3]*d[7] - d[4]*d[6])*t;
   inv[7] = (d[6]*d[1] - d[7]*d[0])*t;
   inv[8] = (d[0]*d[4] - d[1]*d[3])*t;
}

// AB = A*B, A is ah x aw, B is aw x bw.
inline void Mult(int ah, int aw, int bw, const double *A, const double *B, double *AB)
{
   for (int i = 0; i < ah*bw; i++) { AB[i] = 0.0; }
   for (int j = 0; j < bw; j++)
      for (int k = 0; k < aw; k++)
         for (int i = 0; i < ah; i++) { AB[i + j*ah] += A[i + k*ah]*B[k + j*aw]; }
}
// y = A*x
inline void MultV(int h, int w, const double *A, const double *x, double *y)
{
   for (int i = 0; i < h; i++) { y[i] = 0.0; }
   for (int j = 0; j < w; j++)
      for (int i = 0; i < h; i++) { y[i] += A[i + j*h]*x[j]; }
}
// ABt = A*B^t, A ah x aw, B bh x aw
inline void MultABt(int ah, int aw, int bh, const double *A, const double *B, double *ABt)
{
   for (int i = 0; i < ah*bh; i++) { ABt[i] = 0.0; }
   for (int k = 0; k < aw; k++)
      for (int j = 0; j < bh; j++)
         for (int i = 0; i < ah; i++) { ABt[i + j*ah] += A[i + k*ah]*B[j + k*bh]; }
}
inline void Symmetrize(int n, double *d)
{
   for (int i = 0; i < n; i++)
      for (int j = 0; j < i; j++)
      {
         const double a = 0.5*(d[i*n + j] + d[j*n + i]);
         d[j*n + i] = d[i*n + j] = a;
      }
}
// C = A + c*B
inline void Add(int h, int w, double c, const double *A, const double *B, double *C)
{
   for (int i = 0; i < h*w; i++) { C[i] = A[i] + c*B[i]; }
}
inline double Norml2(int size, const double *data)
{
   if (size == 0) { return 0.0; }
   if (size == 1) { return std::fabs(data[0]); }
   double scale = 0.0, sum = 0.0;
   for (int i = 0; i < size; i++)
   {
      if (data[i] != 0.0)
      {
         const double absdata = std::fabs(data[i]);
         if (scale <= absdata)
         {
            const double sqr_arg = scale/absdata;
            sum = 1.0 + sum*(sqr_arg*sqr_arg);
            scale = absdata;
            continue;
         }
         const double sqr_arg = absdata/scale;
         sum += sqr_arg*sqr_arg;
      }
   }
   return scale*std::sqrt(sum);
}

// power-of-two scale so that d_max/mult is in [0.5,1)
inline double ScalingFactor(double d_max)
{
   if (d_max > 0.0)
   {
      int e; double m = std::frexp(d_max, &e);
      if (e == DBL_MAX_EXP) { m *= 2.0; }
      return d_max/m;
   }
   return 1.0;
}

// Parlett, "The Symmetric Eigenvalue Problem", pp.189-190: rotation (c,s) that
// diagonalises [d1 d12; d12 d2]; on return d1, d2 are the eigenvalues with
// eigenvectors (c,-s) and (s,c).
inline void Eigensystem2S(const double d12, double &d1, double &d2, double &c, double &s)
{
   const double sqrt_1_eps = std::sqrt(1.0/DBL_EPSILON);
   if (d12 == 0.0) { c = 1.0; s = 0.0; return; }
   double t;
   const double zeta = (d2 - d1)/(2*d12);
   const double azeta = std::fabs(zeta);
   if (azeta < sqrt_1_eps)
   {
      t = std::copysign(1.0/(azeta + std::sqrt(1.0 + zeta*zeta)), zeta);
   }
   else
   {
      t = std::copysign(0.5/azeta, zeta);
   }
   c = std::sqrt(1.0/(1.0 + t*t));
   s = c*t;
   t *= d12;
   d1 -= t;
   d2 += t;
}

template<int DIM> inline void CalcEigenvalues(const double *data, double *lambda, double *vec);

// symmetric 2x2: eigenvalues ascending, eigenvectors as columns of vec
template<> inline void CalcEigenvalues<2>(const double *data, double *lambda, double *vec)
{
   double d0 = data[0];
   const double d2 = data[2];
   double d3 = data[3];
   double c, s;
   Eigensystem2S(d2, d0, d3, c, s);
   if (d0 <= d3)
   {
      lambda[0] = d0; lambda[1] = d3;
      vec[0] =  c; vec[1] = -s;
      vec[2] =  s; vec[3] =  c;
   }
   else
   {
      lambda[0] = d3; lambda[1] = d0;
      vec[0] =  s; vec[1] =  c;
      vec[2] =  c; vec[3] = -s;
   }
}

// Unit vector in the (near-)kernel of the symmetric matrix
// [c1 d12 d13; d12 c2 d23; d13 d23 c3], which is singular to round-off because
// one eigenvalue has been subtracted from its diagonal.  The kernel direction is
// the cross product of two rows; the pair with the largest cross product is the
// best conditioned.  Returns false if the matrix has rank <= 1 to round-off.
inline bool KernelVector3S(double c1, double c2, double c3,
                           double d12, double d13, double d23, double *z)
{
   const double r0[3] = {c1, d12, d13}, r1[3] = {d12, c2, d23}, r2[3] = {d13, d23, c3};
   double cr[3][3];
   auto cross = [](const double *a, const double *b, double *o)
   {
      o[0] = a[1]*b[2] - a[2]*b[1];
      o[1] = a[2]*b[0] - a[0]*b[2];
      o[2] = a[0]*b[1] - a[1]*b[0];
   };
   cross(r0, r1, cr[0]); cross(r0, r2, cr[1]); cross(r1, r2, cr[2]);
   int best = 0; double nb = -1.0;
   for (int i = 0; i < 3; i++)
   {
      const double n = cr[i][0]*cr[i][0] + cr[i][1]*cr[i][1] + cr[i][2]*cr[i][2];
      if (n > nb) { nb = n; best = i; }
   }
   const double amax = std::max({std::fabs(c1), std::fabs(c2), std::fabs(c3),
                                 std::fabs(d12), std::fabs(d13), std::fabs(d23)});
   // |cross| ~ amax^2 * sin(angle); rank <= 1 when it is at round-off level
   if (!(nb > 1e-28*amax*amax*amax*amax) || nb == 0.0) { return false; }
   const double inv = 1.0/std::sqrt(nb);
   z[0] = cr[best][0]*inv; z[1] = cr[best][1]*inv; z[2] = cr[best][2]*inv;
   return true;
}

// symmetric 3x3 (upper triangle of column-major data is used): eigenvalues
// ascending in lambda[0..2], eigenvectors as columns vec[0..2], vec[3..5], vec[6..8].
template<> inline void CalcEigenvalues<3>(const double *data, double *lambda, double *vec)
{
   double d11 = data[0];
   double d12 = data[3];
   double d22 = data[4];
   double d13 = data[6];
   double d23 = data[7];
   double d33 = data[8];

   const double d_max = std::max({std::fabs(d11), std::fabs(d22), std::fabs(d33),
                                  std::fabs(d12), std::fabs(d13), std::fabs(d23)});
   const double mult = ScalingFactor(d_max);
   d11 /= mult; d22 /= mult; d33 /= mult;
   d12 /= mult; d13 /= mult; d23 /= mult;

   double aa = (d11 + d22 + d33)/3;
   double c1 = d11 - aa, c2 = d22 - aa, c3 = d33 - aa;

   const double Q = (2*(d12*d12 + d13*d13 + d23*d23) + c1*c1 + c2*c2 + c3*c3)/6;
   double R = (c1*(d23*d23 - c2*c3) + d12*(d12*c3 - 2*d13*d23) + d13*d13*c2)/2;

   auto identity = [&]()
   {
      lambda[0] = lambda[1] = lambda[2] = aa;
      vec[0] = 1.; vec[3] = 0.; vec[6] = 0.;
      vec[1] = 0.; vec[4] = 1.; vec[7] = 0.;
      vec[2] = 0.; vec[5] = 0.; vec[8] = 1.;
   };

   if (Q <= 0.) { identity(); }
   else
   {
      const double sqrtQ = std::sqrt(Q);
      const double sqrtQ3 = Q*sqrtQ;
      double r;
      // the root of the characteristic polynomial that is best separated from
      // the other two
      if (std::fabs(R) >= sqrtQ3)
      {
         r = (R < 0.) ? 2*sqrtQ : -2*sqrtQ;
      }
      else
      {
         R = R/sqrtQ3;
         if (R < 0.) { r = -2*sqrtQ*std::cos((std::acos(R) + 2.0*M_PI)/3); } // max
         else        { r = -2*sqrtQ*std::cos(std::acos(R)/3); }              // min
      }
      aa += r;
      c1 = d11 - aa; c2 = d22 - aa; c3 = d33 - aa;

      double z[3];
      if (!KernelVector3S(c1, c2, c3, d12, d13, d23, z))
      {
         // A - aa*I vanishes to round-off: triple eigenvalue
         identity();
      }
      else
      {
         // Orthonormal completion {z,u,w}: Householder reflector H = I - 2 v v^t
         // with H e_k = +-z, k = index of the smallest |z_k|; columns of H other
         // than k span the orthogonal complement of z.
         int k = 0;
         if (std::fabs(z[1]) < std::fabs(z[k])) { k = 1; }
         if (std::fabs(z[2]) < std::fabs(z[k])) { k = 2; }
         double v[3] = {z[0], z[1], z[2]};
         const double sgn = (z[k] >= 0.) ? 1.0 : -1.0;
         v[k] += sgn;                       // v = z + sgn*e_k
         const double vn2 = v[0]*v[0] + v[1]*v[1] + v[2]*v[2];
         const int i1 = (k + 1) % 3, i2 = (k + 2) % 3;
         double u[3], w[3];
         for (int i = 0; i < 3; i++)
         {
            u[i] = ((i == i1) ? 1.0 : 0.0) - 2.0*v[i]*v[i1]/vn2;
            w[i] = ((i == i2) ? 1.0 : 0.0) - 2.0*v[i]*v[i2]/vn2;
         }
         // A*u, A*w with the scaled A
         auto Amul = [&](const double *x, double *y)
         {
            y[0] = d11*x[0] + d12*x[1] + d13*x[2];
            y[1] = d12*x[0] + d22*x[1] + d23*x[2];
            y[2] = d13*x[0] + d23*x[1] + d33*x[2];
         };
         double Az[3], Au[3], Aw[3];
         Amul(z, Az); Amul(u, Au); Amul(w, Aw);
         const double l1 = z[0]*Az[0] + z[1]*Az[1] + z[2]*Az[2]; // Rayleigh quotient
         double b22 = u[0]*Au[0] + u[1]*Au[1] + u[2]*Au[2];
         double b33 = w[0]*Aw[0] + w[1]*Aw[1] + w[2]*Aw[2];
         const double b23 = u[0]*Aw[0] + u[1]*Aw[1] + u[2]*Aw[2];
         double c, s;
         Eigensystem2S(b23, b22, b33, c, s);
         // eigenvectors of the 2x2 block: (c,-s) for b22, (s,c) for b33
         double e2[3], e3[3];
         for (int i = 0; i < 3; i++)
         {
            e2[i] = c*u[i] - s*w[i];
            e3[i] = s*u[i] + c*w[i];
         }
         const double lam[3] = {l1, b22, b33};
         const double *ev[3] = {z, e2, e3};
         int ord[3] = {0, 1, 2};
         // ascending, stable
         if (lam[ord[1]] < lam[ord[0]]) { std::swap(ord[0], ord[1]); }
         if (lam[ord[2]] < lam[ord[1]]) { std::swap(ord[1], ord[2]); }
         if (lam[ord[1]] < lam[ord[0]]) { std::swap(ord[0], ord[1]); }
         for (int j = 0; j < 3; j++)
         {
            lambda[j] = lam[ord[j]];
            for (int i = 0; i < 3; i++) { vec[i + 3*j] = ev[ord[j]][i]; }
         }
      }
   }
   lambda[0] *= mult; lambda[1] *= mult; lambda[2] *= mult;
}

template<int DIM> inline double CalcSingularvalue(const double *data, int i);

// i-th singular value (descending order: i = DIM-1 is the smallest)
template<> inline double CalcSingularvalue<2>(const double *data, int i)
{
   double d0 = data[0], d1 = data[1], d2 = data[2], d3 = data[3];
   const double d_max = std::max({std::fabs(d0), std::fabs(d1), std::fabs(d2), std::fabs(d3)});
   const double mult = ScalingFactor(d_max);
   d0 /= mult; d1 /= mult; d2 /= mult; d3 /= mult;
   double t = 0.5*((d0 + d2)*(d0 - d2) + (d1 - d3)*(d1 + d3));
   double s = d0*d2 + d1*d3;
   s = std::sqrt(0.5*(d0*d0 + d1*d1 + d2*d2 + d3*d3) + std::sqrt(t*t + s*s));
   if (s == 0.0) { return 0.0; }
   t = std::fabs(d0*d3 - d1*d2)/s;
   if (t > s)
   {
      if (i == 0) { return t*mult; }
      return s*mult;
   }
   if (i == 0) { return s*mult; }
   return t*mult;
}

template<> inline double CalcSingularvalue<3>(const double *data, int i)
{
   double d0 = data[0], d1 = data[1], d2 = data[2];
   double d3 = data[3], d4 = data[4], d5 = data[5];
   double d6 = data[6], d7 = data[7], d8 = data[8];
   double d_max = 0.0;
   for (int k = 0; k < 9; k++) { d_max = std::max(d_max, std::fabs(data[k])); }
   const double mult = ScalingFactor(d_max);
   d0 /= mult; d1 /= mult; d2 /= mult; d3 /= mult; d4 /= mult;
   d5 /= mult; d6 /= mult; d7 /= mult; d8 /= mult;

   // B = J^t J
   double b11 = d0*d0 + d1*d1 + d2*d2;
   double b12 = d0*d3 + d1*d4 + d2*d5;
   double b13 = d0*d6 + d1*d7 + d2*d8;
   double b22 = d3*d3 + d4*d4 + d5*d5;
   double b23 = d3*d6 + d4*d7 + d5*d8;
   double b33 = d6*d6 + d7*d7 + d8*d8;

   double aa = (b11 + b22 + b33)/3;
   double c1, c2, c3;
   {
      // differences of the diagonal entries computed without cancellation
      const double b11_b22 = ((d0 - d3)*(d0 + d3) + (d1 - d4)*(d1 + d4) + (d2 - d5)*(d2 + d5));
      const double b22_b33 = ((d3 - d6)*(d3 + d6) + (d4 - d7)*(d4 + d7) + (d5 - d8)*(d5 + d8));
      const double b33_b11 = ((d6 - d0)*(d6 + d0) + (d7 - d1)*(d7 + d1) + (d8 - d2)*(d8 + d2));
      c1 = (b11_b22 - b33_b11)/3;
      c2 = (b22_b33 - b11_b22)/3;
      c3 = (b33_b11 - b22_b33)/3;
   }
   const double Q = (2*(b12*b12 + b13*b13 + b23*b23) + c1*c1 + c2*c2 + c3*c3)/6;
   double R = (c1*(b23*b23 - c2*c3) + b12*(b12*c3 - 2*b13*b23) + b13*b13*c2)/2;

   if (Q > 0.)
   {
      const double sqrtQ = std::sqrt(Q);
      const double sqrtQ3 = Q*sqrtQ;
      double r;
      bool have_aa = false;
      if (std::fabs(R) >= sqrtQ3)
      {
         r = (R < 0.) ? 2*sqrtQ : -2*sqrtQ;
      }
      else
      {
         R = R/sqrtQ3;
         if (std::fabs(R) <= 0.9)
         {
            if (i == 2)      { aa -= 2*sqrtQ*std::cos(std::acos(R)/3); }                // min
            else if (i == 0) { aa -= 2*sqrtQ*std::cos((std::acos(R) + 2.0*M_PI)/3); }   // max
            else             { aa -= 2*sqrtQ*std::cos((std::acos(R) - 2.0*M_PI)/3); }   // mid
            have_aa = true;
         }
         else if (R < 0.)
         {
            r = -2*sqrtQ*std::cos((std::acos(R) + 2.0*M_PI)/3); // max
            if (i == 0) { aa += r; have_aa = true; }
         }
         else
         {
            r = -2*sqrtQ*std::cos(std::acos(R)/3); // min
            if (i == 2) { aa += r; have_aa = true; }
         }
      }
      if (!have_aa)
      {
         // (aa + r) is the well separated root; the other two are close to each
         // other: deflate with its eigenvector and solve the 2x2 problem.
         const double l1 = aa + r;
         double z[3];
         if (!KernelVector3S(b11 - l1, b22 - l1, b33 - l1, b12, b13, b23, z))
         {
            aa = l1;
         }
         else
         {
            int k = 0;
            if (std::fabs(z[1]) < std::fabs(z[k])) { k = 1; }
            if (std::fabs(z[2]) < std::fabs(z[k])) { k = 2; }
            double v[3] = {z[0], z[1], z[2]};
            v[k] += (z[k] >= 0.) ? 1.0 : -1.0;
            const double vn2 = v[0]*v[0] + v[1]*v[1] + v[2]*v[2];
            const int i1 = (k + 1) % 3, i2 = (k + 2) % 3;
            double u[3], w[3];
            for (int j = 0; j < 3; j++)
            {
               u[j] = ((j == i1) ? 1.0 : 0.0) - 2.0*v[j]*v[i1]/vn2;
               w[j] = ((j == i2) ? 1.0 : 0.0) - 2.0*v[j]*v[i2]/vn2;
            }
            auto Bmul = [&](const double *x, double *y)
            {
               y[0] = b11*x[0] + b12*x[1] + b13*x[2];
               y[1] = b12*x[0] + b22*x[1] + b23*x[2];
               y[2] = b13*x[0] + b23*x[1] + b33*x[2];
            };
            double Bz[3], Bu[3], Bw[3];
            Bmul(z, Bz); Bmul(u, Bu); Bmul(w, Bw);
            const double e1 = z[0]*Bz[0] + z[1]*Bz[1] + z[2]*Bz[2];
            double e2 = u[0]*Bu[0] + u[1]*Bu[1] + u[2]*Bu[2];
            double e3 = w[0]*Bw[0] + w[1]*Bw[1] + w[2]*Bw[2];
            const double e23 = u[0]*Bw[0] + u[1]*Bw[1] + u[2]*Bw[2];
            double c, s;
            Eigensystem2S(e23, e2, e3, c, s);
            double lo = std::min(std::min(e1, e2), e3);
            double hi = std::max(std::max(e1, e2), e3);
            if (i == 2) { aa = lo; }
            else if (i == 0) { aa = hi; }
            else { aa = e1 + e2 + e3 - lo - hi; }
         }
      }
   }
   return std::sqrt(std::fabs(aa))*mult;
}

} // namespace sm
} // namespace oracle

// TEST INFRASTRUCTURE — CPU oracle, not a product path (see oracle/README.md).
//
// CPU restatement of the reference's serial partial-assembly path ("-pa -d cpu"):
//   * force operator F / F^T       : reference laghos_assembly.cpp:145-294 (2D),
//                                    296-514 (3D), 567-713, 715-924, 557-565, 965-973
//                                    (= serial/laghos_assembly.cpp:141-963)
//   * quadrature-data update       : laghos_solver.cpp:1042-1168 (QUpdateBody),
//                                    1263-1411 (QKernel, UpdateQuadratureData)
//   * t=0 setup Rho0DetJ0Vol, h0   : laghos_solver.cpp:1170-1261, 251-262
//   * mass apply / diagonal        : MFEM MassIntegrator PA (not in tree), arithmetic as
//                                    restated in amr/laghos_assembly.cpp:878-963
//   * PCG / Jacobi                 : MFEM CGSolver, OperatorJacobiSmoother
//   * hydro operator + RK4/RK2Avg  : laghos_solver.cpp:308-540, 1436-1487; MFEM RK4Solver
//   * time loop with dt control    : laghos.cpp:706-778, 792-795
// One element at a time, stack arrays, sum factorisation — the same algorithm the
// reference runs on a CPU.  Mesh and initial conditions come from
// laghos_b200/csrc/host/problem.hpp; the 1D tables, quadrature weights and the gather map are the oracle's own
// (own_tables.hpp: independent algorithms, cross-checked against the host set-up, then used instead of it).
#pragma once
#include "../laghos_b200/csrc/host/problem.hpp"
#include "smallmat.hpp"
#include "own_tables.hpp"
#include <chrono>
#include <cstring>
#include <limits>
#include <functional>
#include <thread>
#include <mutex>
#include <condition_variable>

namespace oracle {

using lagb::Problem;

// out = M applied along AXIS of in.  in has extents (N0,N1,N2), x fastest; the
// contracted axis has extent NA and becomes NB.  M(b,a) = M[b + NB*a].
template<int NA, int NB, int N0, int N1, int N2, int AXIS>
static inline void contract(const double *M, const double *in, double *out)
{
   constexpr int O0 = (AXIS == 0) ? NB : N0;
   constexpr int O1 = (AXIS == 1) ? NB : N1;
   constexpr int O2 = (AXIS == 2) ? NB : N2;
   for (int k = 0; k < O2; k++)
      for (int j = 0; j < O1; j++)
         for (int i = 0; i < O0; i++)
         {
            double u = 0.0;
            for (int a = 0; a < NA; a++)
            {
               const int b = (AXIS == 0) ? i : (AXIS == 1) ? j : k;
               const int s0 = (AXIS == 0) ? a : i;
               const int s1 = (AXIS == 1) ? a : j;
               const int s2 = (AXIS == 2) ? a : k;
               u += M[b + NB*a]*in[s0 + N0*(s1 + N1*s2)];
            }
            out[i + O0*(j + O1*k)] = u;
         }
}

struct QuadratureData   // reference laghos_assembly.hpp:31-62
{
   std::vector<double> Jac0inv;      // [i + dim*(j + dim*(e*NQ+q))]
   std::vector<double> stressJinvT;  // [(e*NQ+q) + NE*NQ*(g + dim*c)]
   std::vector<double> rho0DetJ0w;   // [e*NQ+q]
   double h0 = 0.0, dt_est = 0.0;
};

struct TimingData       // reference laghos_solver.hpp:39-56
{
   double sw_cgH1 = 0, sw_cgL2 = 0, sw_force = 0, sw_qdata = 0;
   long long H1iter = 0, L2iter = 0, quad_tstep = 0;
};

static inline double now_s()
{
   return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// ---------------------------------------------------------------------------
// element kernels, templated like the reference (id = DIM<<8 | D1D<<4 | Q1D)
// ---------------------------------------------------------------------------
template<int DIM, int D1D, int Q1D>
struct Kernels
{
   static constexpr int L1D = D1D - 1;
   static constexpr int DZ = (DIM == 3) ? D1D : 1;
   static constexpr int QZ = (DIM == 3) ? Q1D : 1;
   static constexpr int LZ = (DIM == 3) ? L1D : 1;
   static constexpr int ND = D1D*D1D*DZ, NQ = Q1D*Q1D*QZ, NL = L1D*L1D*LZ;

   // values at quadrature points of a scalar field with 1D table T (Q1D x N1D)
   template<int N1D>
   static inline void interp(const double *T, const double *in, double *out)
   {
      constexpr int NZ = (DIM == 3) ? N1D : 1;
      double t0[Q1D*N1D*NZ], t1[Q1D*Q1D*NZ];
      contract<N1D, Q1D, N1D, N1D, NZ, 0>(T, in, t0);
      contract<N1D, Q1D, Q1D, N1D, NZ, 1>(T, t0, t1);
      if (DIM == 3) { contract<N1D, Q1D, Q1D, Q1D, NZ, 2>(T, t1, out); }
      else { for (int i = 0; i < NQ; i++) { out[i] = t1[i]; } }
   }
   // transpose of interp: quadrature -> dofs with Tt (N1D x Q1D)
   template<int N1D>
   static inline void interp_t(const double *Tt, const double *in, double *out)
   {
      constexpr int NZ = (DIM == 3) ? N1D : 1;
      double t0[N1D*Q1D*QZ], t1[N1D*N1D*QZ];
      contract<Q1D, N1D, Q1D, Q1D, QZ, 0>(Tt, in, t0);
      contract<Q1D, N1D, N1D, Q1D, QZ, 1>(Tt, t0, t1);
      if (DIM == 3) { contract<Q1D, N1D, N1D, N1D, QZ, 2>(Tt, t1, out); }
      else { for (int i = 0; i < N1D*N1D*NZ; i++) { out[i] = t1[i]; } }
   }
   // reference-space derivative d/dxi_g of an H1 field at the quadrature points
   static inline void grad(const double *B, const double *G, int g, const double *in, double *out)
   {
      double t0[Q1D*D1D*DZ], t1[Q1D*Q1D*DZ];
      contract<D1D, Q1D, D1D, D1D, DZ, 0>(g == 0 ? G : B, in, t0);
      contract<D1D, Q1D, Q1D, D1D, DZ, 1>(g == 1 ? G : B, t0, t1);
      if (DIM == 3) { contract<D1D, Q1D, Q1D, Q1D, DZ, 2>(g == 2 ? G : B, t1, out); }
      else { for (int i = 0; i < NQ; i++) { out[i] = t1[i]; } }
   }
   // transpose of grad
   static inline void grad_t(const double *Bt, const double *Gt, int g, const double *in, double *out)
   {
      double t0[D1D*Q1D*QZ], t1[D1D*D1D*QZ];
      contract<Q1D, D1D, Q1D, Q1D, QZ, 0>(g == 0 ? Gt : Bt, in, t0);
      contract<Q1D, D1D, D1D, Q1D, QZ, 1>(g == 1 ? Gt : Bt, t0, t1);
      if (DIM == 3) { contract<Q1D, D1D, D1D, D1D, QZ, 2>(g == 2 ? Gt : Bt, t1, out); }
      else { for (int i = 0; i < ND; i++) { out[i] = t1[i]; } }
   }

   // y += G^t B^t D B G x  over elements [e0,e1)  (scalar H1 mass)
   static void MassH1(const Problem &P, const double *D, const double *x, double *y, int e0, int e1)
   {
      const double *B = P.tab.B.data(), *Bt = P.tab.Bt.data();
      for (int e = e0; e < e1; e++)
      {
         const int *map = P.h1_map.data() + (size_t)e*ND;
         double X[ND], QQ[NQ], Y[ND];
         for (int i = 0; i < ND; i++) { X[i] = x[map[i]]; }
         interp<D1D>(B, X, QQ);
         const double *d = D + (size_t)e*NQ;
         for (int q = 0; q < NQ; q++) { QQ[q] *= d[q]; }
         interp_t<D1D>(Bt, QQ, Y);
         for (int i = 0; i < ND; i++) { y[map[i]] += Y[i]; }
      }
   }
   static void MassH1Diag(const Problem &P, const double *D, double *diag)
   {
      const double *B = P.tab.B.data();
      double B2[Q1D*D1D], Bt2[Q1D*D1D];
      for (int q = 0; q < Q1D; q++)
         for (int d = 0; d < D1D; d++)
         {
            B2[q + Q1D*d] = B[q + Q1D*d]*B[q + Q1D*d];
            Bt2[d + D1D*q] = B2[q + Q1D*d];
         }
      for (int e = 0; e < P.NE; e++)
      {
         const int *map = P.h1_map.data() + (size_t)e*ND;
         double Y[ND];
         interp_t<D1D>(Bt2, D + (size_t)e*NQ, Y);
         for (int i = 0; i < ND; i++) { diag[map[i]] += Y[i]; }
      }
   }
   // reference InternalEnergy / KineticEnergy (laghos_solver.cpp:639-697) through
   // ComputeVolumeIntegral (:565-637): sum_q rho0DetJ0w(q) * sum_k f_k(q)^norm, f = e (norm 1)
   // or v (norm 2), interpolated at the quadrature points
   static double EnergyIntegralL2(const Problem &P, const double *rho0DetJ0w, const double *e_gf)
   {
      const double *BL = P.tab.BL.data();
      double sum = 0.0;
      for (int e = 0; e < P.NE; e++)
      {
         double QQ[NQ];
         interp<L1D>(BL, e_gf + (size_t)e*NL, QQ);
         const double *d = rho0DetJ0w + (size_t)e*NQ;
         for (int q = 0; q < NQ; q++) { sum += QQ[q]*d[q]; }
      }
      return sum;
   }
   // reference LagrangianHydroOperator::ComputeDensity (laghos_solver.cpp:542-563): per element
   // rho_z = Mrho^-1 rhs with Mrho = MassIntegrator on the CURRENT mesh (sum_q w detJ phi_i phi_j) and
   // rhs_i = sum_q rho0DetJ0w phi_i (DensityIntegrator, laghos_assembly.cpp:26-41); dense LU with partial
   // pivoting like MFEM's DenseMatrixInverse.
   static void ComputeDensity(const Problem &P, const double *x, const double *rho0DetJ0w, double *rho)
   {
      const double *BL = P.tab.BL.data();
      std::vector<double> phi((size_t)NL*NQ);
      for (int i = 0; i < NL; i++)
      {
         double unit[NL];
         for (int k = 0; k < NL; k++) { unit[k] = (k == i) ? 1.0 : 0.0; }
         interp<L1D>(BL, unit, phi.data() + (size_t)i*NQ);
      }
      for (int e = 0; e < P.NE; e++)
      {
         double J[DIM*DIM*NQ], wd[NQ];
         VectorGrad(P, x, e, J);
         for (int q = 0; q < NQ; q++) { wd[q] = P.qweights[q]*sm::Det<DIM>(J + DIM*DIM*q); }
         std::vector<double> M((size_t)NL*NL), b(NL);
         for (int i = 0; i < NL; i++)
         {
            double r = 0.0;
            for (int q = 0; q < NQ; q++) { r += rho0DetJ0w[(size_t)e*NQ + q]*phi[(size_t)i*NQ + q]; }
            b[i] = r;
            for (int j = 0; j < NL; j++)
            {
               double a = 0.0;
               for (int q = 0; q < NQ; q++) { a += wd[q]*phi[(size_t)i*NQ + q]*phi[(size_t)j*NQ + q]; }
               M[(size_t)i*NL + j] = a;
            }
         }
         // LU with partial pivoting, in place; then forward / back substitution
         for (int k = 0; k < NL; k++)
         {
            int piv = k;
            for (int i = k + 1; i < NL; i++) { if (std::fabs(M[(size_t)i*NL + k]) > std::fabs(M[(size_t)piv*NL + k])) { piv = i; } }
            if (piv != k)
            {
               for (int j = 0; j < NL; j++) { std::swap(M[(size_t)k*NL + j], M[(size_t)piv*NL + j]); }
               std::swap(b[k], b[piv]);
            }
            for (int i = k + 1; i < NL; i++)
            {
               const double f = M[(size_t)i*NL + k]/M[(size_t)k*NL + k];
               for (int j = k + 1; j < NL; j++) { M[(size_t)i*NL + j] -= f*M[(size_t)k*NL + j]; }
               b[i] -= f*b[k];
            }
         }
         for (int i = NL - 1; i >= 0; i--)
         {
            double r = b[i];
            for (int j = i + 1; j < NL; j++) { r -= M[(size_t)i*NL + j]*rho[(size_t)e*NL + j]; }
            rho[(size_t)e*NL + i] = r/M[(size_t)i*NL + i];
         }
      }
   }
   static double EnergyIntegralH1(const Problem &P, const double *rho0DetJ0w, const double *v)
   {
      const double *B = P.tab.B.data();
      double sum = 0.0;
      for (int e = 0; e < P.NE; e++)
      {
         const int *map = P.h1_map.data() + (size_t)e*ND;
         const double *d = rho0DetJ0w + (size_t)e*NQ;
         double vmag[NQ];
         for (int q = 0; q < NQ; q++) { vmag[q] = 0.0; }
         for (int c = 0; c < DIM; c++)
         {
            double X[ND], QQ[NQ];
            for (int i = 0; i < ND; i++) { X[i] = v[(size_t)c*P.ndofs_h1 + map[i]]; }
            interp<D1D>(B, X, QQ);
            for (int q = 0; q < NQ; q++) { vmag[q] += QQ[q]*QQ[q]; }
         }
         for (int q = 0; q < NQ; q++) { sum += vmag[q]*d[q]; }
      }
      return sum;
   }
   // L2 (Bernstein) mass, block diagonal: y = BL^t D BL x
   static void MassL2(const Problem &P, const double *D, const double *x, double *y, int e0, int e1)
   {
      const double *BL = P.tab.BL.data(), *BLt = P.tab.BLt.data();
      for (int e = e0; e < e1; e++)
      {
         double QQ[NQ];
         interp<L1D>(BL, x + (size_t)e*NL, QQ);
         const double *d = D + (size_t)e*NQ;
         for (int q = 0; q < NQ; q++) { QQ[q] *= d[q]; }
         interp_t<L1D>(BLt, QQ, y + (size_t)e*NL);
      }
   }

   // reference ForceMult2D/3D + H1R->MultTranspose: y (H1 L-vector, byNODES) += ...
   static void ForceMult(const Problem &P, const double *sJit, const double *x, double *y, int e0, int e1)
   {
      const double *BL = P.tab.BL.data(), *Bt = P.tab.Bt.data(), *Gt = P.tab.Gt.data();
      const double eps1 = std::numeric_limits<double>::epsilon();
      const double eps2 = eps1*eps1;
      const size_t NEQ = (size_t)P.NE*NQ;
      for (int e = e0; e < e1; e++)
      {
         const int *map = P.h1_map.data() + (size_t)e*ND;
         double QQQ[NQ], QQg[NQ], part[DIM][ND];
         interp<L1D>(BL, x + (size_t)e*NL, QQQ);
         for (int c = 0; c < DIM; c++)
         {
            for (int g = 0; g < DIM; g++)
            {
               const double *s = sJit + (size_t)e*NQ + NEQ*(g + DIM*c);
               for (int q = 0; q < NQ; q++) { QQg[q] = QQQ[q]*s[q]; }
               grad_t(Bt, Gt, g, QQg, part[g]);
            }
            for (int i = 0; i < ND; i++)
            {
               double v = part[0][i] + part[1][i];
               if (DIM == 3) { v += part[2][i]; }
               if (std::fabs(v) < eps2) { v = 0.0; }
               y[(size_t)c*P.ndofs_h1 + map[i]] += v;
            }
         }
      }
   }
   // reference H1R->Mult + ForceMultTranspose2D/3D: e_out (L2 layout) overwritten
   static void ForceMultTranspose(const Problem &P, const double *sJit, const double *v, double *eout, int e0, int e1)
   {
      const double *BLt = P.tab.BLt.data(), *B = P.tab.B.data(), *G = P.tab.G.data();
      const size_t NEQ = (size_t)P.NE*NQ;
      for (int e = e0; e < e1; e++)
      {
         const int *map = P.h1_map.data() + (size_t)e*ND;
         double QQQ[NQ], V[ND], dq[DIM][NQ];
         for (int q = 0; q < NQ; q++) { QQQ[q] = 0.0; }
         for (int c = 0; c < DIM; c++)
         {
            for (int i = 0; i < ND; i++) { V[i] = v[(size_t)c*P.ndofs_h1 + map[i]]; }
            for (int g = 0; g < DIM; g++) { grad(B, G, g, V, dq[g]); }
            for (int q = 0; q < NQ; q++)
            {
               double s = dq[0][q]*sJit[(size_t)e*NQ + q + NEQ*(0 + DIM*c)] +
                          dq[1][q]*sJit[(size_t)e*NQ + q + NEQ*(1 + DIM*c)];
               if (DIM == 3) { s += dq[2][q]*sJit[(size_t)e*NQ + q + NEQ*(2 + DIM*c)]; }
               QQQ[q] += s;
            }
         }
         interp_t<L1D>(BLt, QQQ, eout + (size_t)e*NL);
      }
   }

   // J(c,d) = d x_c / d xi_d at the quadrature points, layout [c + DIM*(d + DIM*q)]
   // (QuadratureInterpolator::Derivatives, QVectorLayout::byVDIM)
   static inline void VectorGrad(const Problem &P, const double *field, int e, double *out)
   {
      const double *B = P.tab.B.data(), *G = P.tab.G.data();
      const int *map = P.h1_map.data() + (size_t)e*ND;
      double V[ND], dq[NQ];
      for (int c = 0; c < DIM; c++)
      {
         for (int i = 0; i < ND; i++) { V[i] = field[(size_t)c*P.ndofs_h1 + map[i]]; }
         for (int d = 0; d < DIM; d++)
         {
            grad(B, G, d, V, dq);
            for (int q = 0; q < NQ; q++) { out[c + DIM*(d + DIM*q)] = dq[q]; }
         }
      }
   }

   // reference Rho0DetJ0Vol (laghos_solver.cpp:1170-1261)
   static void Rho0DetJ0Vol(const Problem &P, const double *x0, const double *rho0_gf,
                            QuadratureData &qd, std::vector<double> &massD, double &volume)
   {
      const double *BL = P.tab.BL.data();
      double vol = 0.0;
      for (int e = 0; e < P.NE; e++)
      {
         double J[DIM*DIM*NQ], R[NQ];
         VectorGrad(P, x0, e, J);
         interp<L1D>(BL, rho0_gf + (size_t)e*NL, R);
         for (int q = 0; q < NQ; q++)
         {
            const double *Jq = J + DIM*DIM*q;
            const double det = sm::Det<DIM>(Jq);
            const double W = P.qweights[q];
            qd.rho0DetJ0w[(size_t)e*NQ + q] = W*R[q]*det;
            massD[(size_t)e*NQ + q] = W*P.rho0_q[(size_t)e*NQ + q]*det;
            double *inv = qd.Jac0inv.data() + (size_t)DIM*DIM*((size_t)e*NQ + q);
            const double r_idetJ = 1.0/det;
            if (DIM == 2)
            {
               const double J11 = Jq[0], J21 = Jq[1], J12 = Jq[2], J22 = Jq[3];
               inv[0] =  J22*r_idetJ;
               inv[1] = -J21*r_idetJ;
               inv[2] = -J12*r_idetJ;
               inv[3] =  J11*r_idetJ;
            }
            else
            {
               // J(i,j) column-major: Jq[i + 3*j]
               const double J11 = Jq[0], J21 = Jq[1], J31 = Jq[2];
               const double J12 = Jq[3], J22 = Jq[4], J32 = Jq[5];
               const double J13 = Jq[6], J23 = Jq[7], J33 = Jq[8];
               // Index placement exactly as the reference writes invJ(i,j,q,e) at
               // [i + 3*j] (laghos_solver.cpp:1243-1251).  Note that this is the
               // TRANSPOSE of the textbook inverse; it is harmless for the
               // reference's Cartesian initial meshes (diagonal J0) and is kept
               // for parity.
               inv[0] = r_idetJ*((J22*J33) - (J23*J32));
               inv[1] = r_idetJ*((J32*J13) - (J33*J12));
               inv[2] = r_idetJ*((J12*J23) - (J13*J22));
               inv[3] = r_idetJ*((J23*J31) - (J21*J33));
               inv[4] = r_idetJ*((J33*J11) - (J31*J13));
               inv[5] = r_idetJ*((J13*J21) - (J11*J23));
               inv[6] = r_idetJ*((J21*J32) - (J22*J31));
               inv[7] = r_idetJ*((J31*J12) - (J32*J11));
               inv[8] = r_idetJ*((J11*J22) - (J12*J21));
            }
            vol += W*det;
         }
      }
      volume = vol;
   }

   static inline double smooth_step_01(double x, double eps)
   {
      const double y = (x + eps)/(2.0*eps);
      if (y < 0.0) { return 0.0; }
      if (y > 1.0) { return 1.0; }
      return (3.0 - 2.0*y)*y*y;
   }

   // reference QUpdate::UpdateQuadratureData + QKernel + QUpdateBody; returns the
   // minimum of dt_est over elements [e0,e1), starting from dt_in at every point.
   static double QUpdate(const Problem &P, const double *S, bool use_viscosity, bool use_vorticity,
                         double cfl, double dt_in, QuadratureData &qd, int e0, int e1)
   {
      constexpr int DIM2 = DIM*DIM;
      const double *x = S, *v = S + P.h1_vsize(), *en = S + 2*P.h1_vsize();
      const double *BL = P.tab.BL.data();
      const double h0 = qd.h0, h1order = (double)(D1D - 1);
      const double infinity = std::numeric_limits<double>::infinity();
      const size_t NEQ = (size_t)P.NE*NQ;
      double dt_min = dt_in;
      for (int e = e0; e < e1; e++)
      {
         double Jall[DIM2*NQ], dVall[DIM2*NQ], eq[NQ];
         VectorGrad(P, x, e, Jall);
         if (use_viscosity) { VectorGrad(P, v, e, dVall); }
         interp<L1D>(BL, en + (size_t)e*NL, eq);
         const double gamma = P.gamma[e];
         for (int q = 0; q < NQ; q++)
         {
            double Jinv[DIM2], stress[DIM2], sgrad_v[DIM2], eig_val_data[3], eig_vec_data[9];
            double compr_dir[DIM], Jpi[DIM2], ph_dir[DIM], stressJiT[DIM2];
            double min_detJ = infinity;
            const size_t eqi = (size_t)e*NQ + q;
            const double weight = P.qweights[q];
            const double inv_weight = 1./weight;
            const double *J = Jall + DIM2*q;
            const double detJ = sm::Det<DIM>(J);
            min_detJ = std::fmin(min_detJ, detJ);
            sm::CalcInverse<DIM>(J, Jinv);
            const double R = inv_weight*qd.rho0DetJ0w[eqi]/detJ;
            const double E = std::fmax(0.0, eq[q]);
            const double Pr = (gamma - 1.0)*R*E;
            const double Sd = std::sqrt(gamma*(gamma - 1.0)*E);
            for (int k = 0; k < DIM2; k++) { stress[k] = 0.0; }
            for (int d = 0; d < DIM; d++) { stress[d*DIM + d] = -Pr; }
            double visc_coeff = 0.0;
            if (use_viscosity)
            {
               const double *dV = dVall + DIM2*q;
               sm::Mult(DIM, DIM, DIM, dV, Jinv, sgrad_v);
               double vorticity_coeff = 1.0;
               if (use_vorticity)
               {
                  // FNorm / Trace: reference laghos_solver.cpp:988-1040
                  double max_norm = 0.0;
                  for (int i = 0; i < DIM2; i++) { max_norm = std::fmax(max_norm, std::fabs(sgrad_v[i])); }
                  double grad_norm = 0.0;
                  if (max_norm != 0.0)
                  {
                     double fnorm2 = 0.0;
                     for (int i = 0; i < DIM2; i++) { const double en_ = sgrad_v[i]/max_norm; fnorm2 += en_*en_; }
                     grad_norm = max_norm*std::sqrt(fnorm2);
                  }
                  double tr = 0.0;
                  for (int i = 0; i < DIM; i++) { tr += sgrad_v[i + i*DIM]; }
                  const double div_v = std::fabs(tr);
                  vorticity_coeff = (grad_norm > 0.0) ? div_v/grad_norm : 1.0;
               }
               sm::Symmetrize(DIM, sgrad_v);
               sm::CalcEigenvalues<DIM>(sgrad_v, eig_val_data, eig_vec_data);
               for (int k = 0; k < DIM; k++) { compr_dir[k] = eig_vec_data[k]; }
               sm::Mult(DIM, DIM, DIM, J, qd.Jac0inv.data() + eqi*DIM2, Jpi);
               sm::MultV(DIM, DIM, Jpi, compr_dir, ph_dir);
               const double ph_dir_nl2 = sm::Norml2(DIM, ph_dir);
               const double compr_dir_nl2 = sm::Norml2(DIM, compr_dir);
               const double H = h0*ph_dir_nl2/compr_dir_nl2;
               const double mu = eig_val_data[0];
               visc_coeff = 2.0*R*H*H*std::fabs(mu);
               const double eps = 1e-12;
               visc_coeff += 0.5*R*H*Sd*vorticity_coeff*(1.0 - smooth_step_01(mu - 2.0*eps, eps));
               sm::Add(DIM, DIM, visc_coeff, stress, sgrad_v, stress);
            }
            const double sv = sm::CalcSingularvalue<DIM>(J, DIM - 1);
            const double h_min = sv/h1order;
            const double ih_min = 1./h_min;
            const double irho_ih_min_sq = ih_min*ih_min/R;
            const double idt = Sd*ih_min + 2.5*visc_coeff*irho_ih_min_sq;
            double dt_q = dt_in;
            if (min_detJ < 0.0) { dt_q = 0.0; }
            else if (idt > 0.0) { dt_q = std::fmin(dt_q, cfl/idt); }
            dt_min = std::fmin(dt_min, dt_q);
            sm::MultABt(DIM, DIM, DIM, stress, Jinv, stressJiT);
            for (int k = 0; k < DIM2; k++) { stressJiT[k] *= weight*detJ; }
            for (int vd = 0; vd < DIM; vd++)
               for (int gd = 0; gd < DIM; gd++)
               {
                  qd.stressJinvT[eqi + NEQ*(gd + vd*DIM)] = stressJiT[vd + gd*DIM];
               }
         }
      }
      return dt_min;
   }

   // 2D Taylor-Green energy source (reference laghos_solver.cpp:455-465,
   // laghos_solver.hpp:208-218): e_src_i = sum_q w detJ f(x_q) phi_i(q), current mesh.
   static void TaylorSource(const Problem &P, const double *x, double *esrc)
   {
      const double *B = P.tab.B.data(), *BLt = P.tab.BLt.data();
      for (int e = 0; e < P.NE; e++)
      {
         const int *map = P.h1_map.data() + (size_t)e*ND;
         double J[DIM*DIM*NQ], xq[DIM][NQ], X[ND], f[NQ];
         VectorGrad(P, x, e, J);
         for (int c = 0; c < DIM; c++)
         {
            for (int i = 0; i < ND; i++) { X[i] = x[(size_t)c*P.ndofs_h1 + map[i]]; }
            interp<D1D>(B, X, xq[c]);
         }
         for (int q = 0; q < NQ; q++)
         {
            const double det = sm::Det<DIM>(J + DIM*DIM*q);
            const double fx = 3.0/8.0*M_PI*(std::cos(3.0*M_PI*xq[0][q])*std::cos(M_PI*xq[1][q]) -
                                           std::cos(M_PI*xq[0][q])*std::cos(3.0*M_PI*xq[1][q]));
            f[q] = P.qweights[q]*det*fx;
         }
         interp_t<L1D>(BLt, f, esrc + (size_t)e*NL);
      }
   }
};

// type-erased kernel table
struct KernelTable
{
   void (*MassH1)(const Problem&, const double*, const double*, double*, int, int) = nullptr;
   void (*MassH1Diag)(const Problem&, const double*, double*) = nullptr;
   void (*MassL2)(const Problem&, const double*, const double*, double*, int, int) = nullptr;
   double (*EnergyIntegralL2)(const Problem&, const double*, const double*) = nullptr;
   double (*EnergyIntegralH1)(const Problem&, const double*, const double*) = nullptr;
   void (*ForceMult)(const Problem&, const double*, const double*, double*, int, int) = nullptr;
   void (*ForceMultTranspose)(const Problem&, const double*, const double*, double*, int, int) = nullptr;
   void (*Rho0DetJ0Vol)(const Problem&, const double*, const double*, QuadratureData&, std::vector<double>&, double&) = nullptr;
   double (*QUpdate)(const Problem&, const double*, bool, bool, double, double, QuadratureData&, int, int) = nullptr;
   void (*TaylorSource)(const Problem&, const double*, double*) = nullptr;
   void (*ComputeDensity)(const Problem&, const double*, const double*, double*) = nullptr;
};

template<int DIM, int D1D, int Q1D>
static KernelTable make_table()
{
   using K = Kernels<DIM, D1D, Q1D>;
   KernelTable t;
   t.MassH1 = &K::MassH1; t.MassH1Diag = &K::MassH1Diag; t.MassL2 = &K::MassL2;
   t.EnergyIntegralL2 = &K::EnergyIntegralL2; t.EnergyIntegralH1 = &K::EnergyIntegralH1;
   t.ForceMult = &K::ForceMult; t.ForceMultTranspose = &K::ForceMultTranspose;
   t.Rho0DetJ0Vol = &K::Rho0DetJ0Vol; t.QUpdate = &K::QUpdate; t.TaylorSource = &K::TaylorSource;
   t.ComputeDensity = &K::ComputeDensity;
   return t;
}

// (D1D,Q1D) pairs of reference laghos_assembly.cpp:536-548 plus the 3D (6,10)
// the BASELINE order sweep asks for and the reference lacks.
static inline KernelTable get_table(int dim, int D1D, int Q1D)
{
   const int id = (dim << 8) | (D1D << 4) | Q1D;
   switch (id)
   {
      case 0x222: return make_table<2,2,2>();
      case 0x234: return make_table<2,3,4>();
      case 0x246: return make_table<2,4,6>();
      case 0x258: return make_table<2,5,8>();
      case 0x26A: return make_table<2,6,10>();
      case 0x322: return make_table<3,2,2>();
      case 0x334: return make_table<3,3,4>();
      case 0x346: return make_table<3,4,6>();
      case 0x358: return make_table<3,5,8>();
      case 0x36A: return make_table<3,6,10>();
   }
   char msg[64]; snprintf(msg, sizeof msg, "Unknown kernel 0x%x", id);
   throw std::runtime_error(msg);
}

// element-parallel helper (stand-in for `mpirun -np <cores>`); threads = 1 is the
// serial reference path.  Scatter kernels are run serially unless coloured.
// Persistent worker pool: a fork-join per call would cost more than the kernels at high
// thread counts (128-core GPU hosts).
struct WorkerPool
{
   std::vector<std::thread> workers;
   std::mutex mtx; std::condition_variable cv_start, cv_done;
   const std::function<void(int,int)> *job = nullptr;
   int job_n = 0, job_parts = 0, generation = 0, pending = 0;
   bool stop = false;
   void worker(int id)
   {
      int seen = 0;
      for (;;)
      {
         const std::function<void(int,int)> *f; int n, parts;
         {
            std::unique_lock<std::mutex> lk(mtx);
            cv_start.wait(lk, [&] { return stop || generation != seen; });
            if (stop) { return; }
            seen = generation; f = job; n = job_n; parts = job_parts;
         }
         if (id < parts - 1)
         {
            const int a = (int)((long long)n*id/parts), b = (int)((long long)n*(id + 1)/parts);
            (*f)(a, b);
         }
         {
            std::lock_guard<std::mutex> lk(mtx);
            if (--pending == 0) { cv_done.notify_one(); }
         }
      }
   }
   void ensure(int n)
   {
      while ((int)workers.size() < n) { const int id = (int)workers.size(); workers.emplace_back([this, id] { worker(id); }); }
   }
   // run f over [0,n) split into `parts` contiguous chunks (chunk 0 on the caller)
   void run(int n, int parts, const std::function<void(int,int)> &f)
   {
      ensure(parts - 1);
      {
         std::lock_guard<std::mutex> lk(mtx);
         job = &f; job_n = n; job_parts = parts; pending = (int)workers.size(); generation++;
      }
      cv_start.notify_all();
      // the caller takes the LAST chunk (workers take chunks 0..parts-2 by id)
      { const int id = parts - 1; const int a = (int)((long long)n*id/parts), b = (int)((long long)n*(id + 1)/parts); f(a, b); }
      std::unique_lock<std::mutex> lk(mtx);
      cv_done.wait(lk, [&] { return pending == 0; });
   }
   ~WorkerPool()
   {
      { std::lock_guard<std::mutex> lk(mtx); stop = true; }
      cv_start.notify_all();
      for (auto &w : workers) { w.join(); }
   }
};
inline WorkerPool &worker_pool() { static WorkerPool p; return p; }

inline void parallel_elements(int NE, int nthreads, const std::function<void(int,int)> &f)
{
   if (nthreads <= 1 || NE < 2) { f(0, NE); return; }
   const int nt = std::min(nthreads, NE);
   // workers with id >= parts - 1 idle; chunk parts-1 runs on the caller
   WorkerPool &p = worker_pool();
   p.run(NE, nt, [&](int a, int b) { f(a, b); });
}

// ---------------------------------------------------------------------------
// Hydro operator (reference LagrangianHydroOperator, laghos_solver.cpp:104-540)
// ---------------------------------------------------------------------------
struct Hydro
{
   const Problem &P;
   KernelTable K;
   QuadratureData qd;
   std::vector<double> massD;      // w * rho0(x_q) * detJ0 : MassIntegrator(rho0_coeff)
   std::vector<double> diag, dinv; // H1 scalar mass diagonal (Jacobi)
   TimingData timer;
   double cfl = 0.5, cg_rel_tol = 1e-8;
   int cg_max_iter = 300;
   int nthreads = 1;
   mutable bool qdata_is_current = false;
   std::vector<double> one, rhs, e_rhs, Bv, Xv, cg_r, cg_d, cg_z, l2_r, l2_d, l2_z;
   std::vector<std::vector<int>> colors; // element colouring for threaded scatter

   Hydro(const Problem &P_, double cfl_, double cgt, int cgm, int nthreads_ = 1)
      : P(P_), cfl(cfl_), cg_rel_tol(cgt), cg_max_iter(cgm), nthreads(nthreads_)
   {
      K = get_table(P.dim, P.D1D, P.Q1D);
      const size_t NEQ = (size_t)P.NE*P.NQ;
      qd.Jac0inv.assign(NEQ*P.dim*P.dim, 0.0);
      qd.stressJinvT.assign(NEQ*P.dim*P.dim, 0.0);
      qd.rho0DetJ0w.assign(NEQ, 0.0);
      massD.assign(NEQ, 0.0);
      double vol = 0.0;
      K.Rho0DetJ0Vol(P, P.S0.data(), P.rho0_gf.data(), qd, massD, vol);
      // reference laghos_solver.cpp:253-262 (SQUARE / CUBE only)
      qd.h0 = (P.dim == 2) ? std::sqrt(vol/P.NE) : std::pow(vol/P.NE, 1./3.);
      qd.h0 /= (double)(P.D1D - 1);
      qd.dt_est = std::numeric_limits<double>::infinity();
      diag.assign(P.ndofs_h1, 0.0); dinv.assign(P.ndofs_h1, 0.0);
      K.MassH1Diag(P, massD.data(), diag.data());
      for (int64_t i = 0; i < P.ndofs_h1; i++) { dinv[i] = 1.0/diag[i]; }
      one.assign(P.ndofs_l2, 1.0);
      rhs.assign(P.h1_vsize(), 0.0);
      e_rhs.assign(P.ndofs_l2, 0.0);
      Bv.assign(P.ndofs_h1, 0.0); Xv.assign(P.ndofs_h1, 0.0);
      cg_r.assign(P.ndofs_h1, 0.0); cg_d = cg_r; cg_z = cg_r;
      l2_r.assign(P.ndofs_l2, 0.0); l2_d = l2_r; l2_z = l2_r;
      build_colors();
   }

   void build_colors()
   {
      const int nc = 1 << P.dim;
      colors.assign(nc, {});
      for (int e = 0; e < P.NE; e++)
      {
         const int ix = e % P.nloc[0], iy = (e/P.nloc[0]) % P.nloc[1];
         const int iz = e/(P.nloc[0]*P.nloc[1]);
         colors[(ix & 1) | ((iy & 1) << 1) | ((iz & 1) << 2)].push_back(e);
      }
   }

   // run a scatter-type element kernel: serial in element order (the reference's
   // order) when nthreads == 1, otherwise colour by colour across threads.
   template<typename F> void for_elements_scatter(F f) const
   {
      if (nthreads <= 1) { f(0, P.NE); return; }
      for (const auto &col : colors)
      {
         parallel_elements((int)col.size(), nthreads, [&](int a, int b)
         {
            for (int i = a; i < b; i++) { f(col[i], col[i] + 1); }
         });
      }
   }
   template<typename F> void for_elements(F f) const
   {
      if (nthreads <= 1) { f(0, P.NE); return; }
      parallel_elements(P.NE, nthreads, f);
   }

   // MassPAOperator::Mult (reference laghos_assembly.cpp:117-121), scalar H1
   void VMassMult(const std::vector<int> *ess, const double *x, double *y) const
   {
      std::fill(y, y + P.ndofs_h1, 0.0);
      for_elements_scatter([&](int a, int b) { K.MassH1(P, massD.data(), x, y, a, b); });
      if (ess) { for (int i : *ess) { y[i] = 0.0; } }
   }
   // reference LagrangianHydroOperator::InternalEnergy / KineticEnergy (laghos_solver.cpp:639-697)
   double InternalEnergy(const double *e_gf) const { return K.EnergyIntegralL2(P, qd.rho0DetJ0w.data(), e_gf); }
   double KineticEnergy(const double *v) const { return 0.5*K.EnergyIntegralH1(P, qd.rho0DetJ0w.data(), v); }
   void ComputeDensity(const double *x, double *rho) const { K.ComputeDensity(P, x, qd.rho0DetJ0w.data(), rho); }
   void EMassMult(const double *x, double *y) const
   {
      for_elements([&](int a, int b) { K.MassL2(P, massD.data(), x, y, a, b); });
   }
   // ForcePAOperator::Mult / MultTranspose (reference laghos_assembly.cpp:557-565, 965-973)
   void ForceMult(const double *x, double *y) const
   {
      std::fill(y, y + P.h1_vsize(), 0.0);
      for_elements_scatter([&](int a, int b) { K.ForceMult(P, qd.stressJinvT.data(), x, y, a, b); });
   }
   void ForceMultTranspose(const double *v, double *e) const
   {
      for_elements([&](int a, int b) { K.ForceMultTranspose(P, qd.stressJinvT.data(), v, e, a, b); });
   }

   static double dot(const double *a, const double *b, int64_t n)
   {
      double s = 0.0;
      for (int64_t i = 0; i < n; i++) { s += a[i]*b[i]; }
      return s;
   }

   // MFEM CGSolver::Mult restated.  Returns final_iter.
   // prec == nullptr: unpreconditioned.  iterative_mode: r = b - A x.
   // vector loops of the CG: serial (the reference's order) for nthreads == 1, chunked over
   // the worker pool otherwise (what `mpirun -np <cores>` does to MFEM's vector kernels)
   template<typename F> void vec_for(int64_t n, F f) const
   {
      if (nthreads <= 1 || n < 65536) { f((int64_t)0, n); return; }
      parallel_elements((int)n, nthreads, [&](int a, int b) { f((int64_t)a, (int64_t)b); });
   }
   double pdot(const double *a, const double *b, int64_t n) const
   {
      if (nthreads <= 1 || n < 65536) { return dot(a, b, n); }
      const int parts = (int)std::min<int64_t>(nthreads, n);
      std::vector<double> part(parts, 0.0);
      parallel_elements((int)n, nthreads, [&](int lo, int hi)
      {
         const int id = (int)(((long long)lo*parts + n - 1)/n);
         part[id] = dot(a + lo, b + lo, hi - lo);
      });
      double s = 0.0;
      for (double v : part) { s += v; }   // fixed chunk order
      return s;
   }

   int CG(const std::function<void(const double*, double*)> &A, const double *prec,
          bool iterative_mode, const double *b, double *x, int64_t n,
          double *r, double *d, double *z) const
   {
      const double rel_tol = cg_rel_tol, abs_tol = 0.0;
      if (iterative_mode)
      {
         A(x, r);
         vec_for(n, [&](int64_t lo, int64_t hi) { for (int64_t i = lo; i < hi; i++) { r[i] = b[i] - r[i]; } });
      }
      else
      {
         vec_for(n, [&](int64_t lo, int64_t hi) { for (int64_t i = lo; i < hi; i++) { r[i] = b[i]; x[i] = 0.0; } });
      }
      if (prec) { vec_for(n, [&](int64_t lo, int64_t hi) { for (int64_t i = lo; i < hi; i++) { z[i] = prec[i]*r[i]; d[i] = z[i]; } }); }
      else { vec_for(n, [&](int64_t lo, int64_t hi) { for (int64_t i = lo; i < hi; i++) { d[i] = r[i]; } }); }
      double nom = pdot(d, r, n);
      if (nom < 0.0) { return 0; }
      const double r0 = std::max(nom*rel_tol*rel_tol, abs_tol*abs_tol);
      if (nom <= r0) { return 0; }
      A(d, z);
      double den = pdot(z, d, n);
      if (den <= 0.0) { if (den == 0.0) { return 0; } }
      int final_iter = cg_max_iter;
      for (int i = 1; true; )
      {
         const double alpha = nom/den;
         vec_for(n, [&](int64_t lo, int64_t hi)
         {
            for (int64_t k = lo; k < hi; k++) { x[k] = x[k] + alpha*d[k]; }
            for (int64_t k = lo; k < hi; k++) { r[k] = r[k] - alpha*z[k]; }
            if (prec) { for (int64_t k = lo; k < hi; k++) { z[k] = prec[k]*r[k]; } }
         });
         const double betanom = prec ? pdot(r, z, n) : pdot(r, r, n);
         if (betanom < 0.0) { final_iter = i; break; }
         if (betanom <= r0) { final_iter = i; break; }
         if (++i > cg_max_iter) { break; }
         const double beta = betanom/nom;
         if (prec) { vec_for(n, [&](int64_t lo, int64_t hi) { for (int64_t k = lo; k < hi; k++) { d[k] = z[k] + beta*d[k]; } }); }
         else { vec_for(n, [&](int64_t lo, int64_t hi) { for (int64_t k = lo; k < hi; k++) { d[k] = r[k] + beta*d[k]; } }); }
         A(d, z);
         den = pdot(d, z, n);
         if (den <= 0.0) { if (den == 0.0) { final_iter = i; break; } }
         nom = betanom;
      }
      return final_iter;
   }

   // reference laghos_solver.cpp:807-814 + QUpdate::UpdateQuadratureData
   void UpdateQuadratureData(const double *S)
   {
      if (qdata_is_current) { return; }
      qdata_is_current = true;
      const double t0 = now_s();
      double dt = qd.dt_est;
      if (nthreads <= 1)
      {
         dt = K.QUpdate(P, S, P.use_visc, P.use_vort, cfl, qd.dt_est, qd, 0, P.NE);
      }
      else
      {
         std::vector<double> part(nthreads, qd.dt_est);
         const double dt_in = qd.dt_est;
         int chunk = 0;
         parallel_elements(P.NE, nthreads, [&](int a, int b)
         {
            const int id = __atomic_fetch_add(&chunk, 1, __ATOMIC_RELAXED);
            part[id] = K.QUpdate(P, S, P.use_visc, P.use_vort, cfl, dt_in, qd, a, b);
         });
         for (double p : part) { dt = std::fmin(dt, p); }
      }
      qd.dt_est = dt;
      timer.sw_qdata += now_s() - t0;
      timer.quad_tstep += P.NE;
   }

   // reference laghos_solver.cpp:329-399
   void SolveVelocity(const double *S, double *dS_dt)
   {
      UpdateQuadratureData(S);
      double *dv = dS_dt + P.h1_vsize();
      std::fill(dv, dv + P.h1_vsize(), 0.0);
      double t0 = now_s();
      ForceMult(one.data(), rhs.data());
      timer.sw_force += now_s() - t0;
      for (auto &r : rhs) { r = -r; }
      const int64_t size = P.ndofs_h1;
      for (int c = 0; c < P.dim; c++)
      {
         double *dvc = dv + c*size;
         for (int64_t i = 0; i < size; i++) { Bv[i] = rhs[c*size + i]; }
         if (P.source == 2)
         {
            // Rayleigh-Taylor acceleration source (laghos_solver.cpp:340-380):
            // B += M * accel_c, accel = (0,-1), full operator (no essential rows)
            std::vector<double> AC(size, (c == 1) ? -1.0 : 0.0), BA(size, 0.0);
            VMassMult(nullptr, AC.data(), BA.data());
            for (int64_t i = 0; i < size; i++) { Bv[i] += BA[i]; }
         }
         for (int64_t i = 0; i < size; i++) { Xv[i] = dvc[i]; }
         const std::vector<int> &ess = P.ess[c];
         for (int i : ess) { Bv[i] = 0.0; }   // EliminateRHS
         t0 = now_s();
         const int it = CG([&](const double *x, double *y) { VMassMult(&ess, x, y); },
                           dinv.data(), true, Bv.data(), Xv.data(), size,
                           cg_r.data(), cg_d.data(), cg_z.data());
         timer.sw_cgH1 += now_s() - t0;
         timer.H1iter += it;
         for (int64_t i = 0; i < size; i++) { dvc[i] = Xv[i]; }
      }
   }

   // reference laghos_solver.cpp:442-490
   void SolveEnergy(const double *S, const double *v, double *dS_dt)
   {
      UpdateQuadratureData(S);
      double *de = dS_dt + 2*P.h1_vsize();
      std::fill(de, de + P.ndofs_l2, 0.0);
      std::vector<double> e_source;
      if (P.source == 1)
      {
         e_source.assign(P.ndofs_l2, 0.0);
         K.TaylorSource(P, S, e_source.data());
      }
      double t0 = now_s();
      ForceMultTranspose(v, e_rhs.data());
      timer.sw_force += now_s() - t0;
      if (P.source == 1) { for (int64_t i = 0; i < P.ndofs_l2; i++) { e_rhs[i] += e_source[i]; } }
      t0 = now_s();
      const int it = CG([&](const double *x, double *y) { EMassMult(x, y); }, nullptr, false,
                        e_rhs.data(), de, P.ndofs_l2, l2_r.data(), l2_d.data(), l2_z.data());
      timer.sw_cgL2 += now_s() - t0;
      timer.L2iter += (it == 0) ? 1 : it;
   }

   // reference laghos_solver.cpp:308-327
   void Mult(const double *S, double *dS_dt)
   {
      const double *v = S + P.h1_vsize();
      for (int64_t i = 0; i < P.h1_vsize(); i++) { dS_dt[i] = v[i]; }
      SolveVelocity(S, dS_dt);
      SolveEnergy(S, v, dS_dt);
      qdata_is_current = false;
   }
   double GetTimeStepEstimate(const double *S)
   {
      UpdateQuadratureData(S);
      return qd.dt_est;
   }
   void ResetTimeStepEstimate() { qd.dt_est = std::numeric_limits<double>::infinity(); }
   void ResetQuadratureData() { qdata_is_current = false; }

   // FOM of reference laghos_solver.cpp:699-747 (steps already multiplied by stages)
   void FOM(long long steps_x_stages, double fom[5]) const
   {
      const double H1GTVSize = (double)P.h1_vsize(), L2GTVSize = (double)P.ndofs_l2;
      const long long H1it = timer.H1iter/P.dim;
      const double T0 = timer.sw_cgH1, T2 = timer.sw_force, T3 = timer.sw_qdata;
      const double T4 = T0 + T2 + T3;
      fom[1] = 1e-6*H1GTVSize*H1it/T0;
      fom[2] = 1e-6*steps_x_stages*(H1GTVSize + L2GTVSize)/T2;
      fom[3] = 1e-6*timer.quad_tstep*P.NQ/T3;
      fom[0] = (fom[1]*T0 + fom[2]*T2 + fom[3]*T3)/T4;
      fom[4] = T4;
   }
};

struct RunOptions
{
   int ode_solver_type = 4;   // 4: RK4 (default, laghos.cpp:142), 7: RK2Avg, 1: Euler, 2: RK2(0.5), 3: RK3SSP
   double t_final = 0.6;
   int max_tsteps = -1;
   double cfl = 0.5, cg_tol = 1e-8;
   int cg_max_iter = 300;
   int nthreads = 1;
   bool verbose = false;
   int vis_steps = 5;
};
struct RunResult
{
   int steps = 0;         // including repeated ones (laghos.cpp:760)
   int ti_last = 0;       // accepted step index of the last line printed
   double t = 0, dt = 0, e_norm = 0;
   std::vector<std::pair<int,double>> e_norm_history; // (ti, |e|) after every accepted step
   double fom[5] = {0, 0, 0, 0, 0};
   TimingData timer;
   int stages = 4;
   double energy_init = 0, energy_final = 0;   // IE + KE (laghos.cpp:664-665, 956-962)
};

// time loop: reference laghos.cpp:706-778 + ODE solvers (MFEM RK4Solver et al.,
// in-tree RK2AvgSolver laghos_solver.cpp:1436-1487)
static inline RunResult run(const Problem &P, const RunOptions &opt, std::vector<double> *S_out = nullptr)
{
   Hydro hydro(P, opt.cfl, opt.cg_tol, opt.cg_max_iter, opt.nthreads);
   const int64_t N = P.s_size(), NV = P.h1_vsize();
   std::vector<double> S(P.S0), S_old(N), k(N), y(N), z(N), V(NV), S0(N);
   RunResult res;
   res.energy_init = hydro.InternalEnergy(S.data() + 2*NV) + hydro.KineticEnergy(S.data() + NV);
   hydro.ResetTimeStepEstimate();
   double t = 0.0, dt = hydro.GetTimeStepEstimate(S.data()), t_old;
   bool last_step = false;
   int steps = 0;
   auto add = [&](const std::vector<double> &a, double c, const std::vector<double> &b, std::vector<double> &o, int64_t n)
   { for (int64_t i = 0; i < n; i++) { o[i] = a[i] + c*b[i]; } };

   std::vector<std::vector<double>> k6;
   auto Step = [&](std::vector<double> &x, double &tt, double dtt)
   {
      switch (opt.ode_solver_type)
      {
         case 1:
            hydro.Mult(x.data(), k.data());
            add(x, dtt, k, x, N); break;
         case 2: // RK2Solver(0.5): midpoint
         {
            const double a = 0.5, b = 0.5/a;
            hydro.Mult(x.data(), k.data());
            add(x, (1. - b)*dtt, k, y, N);
            add(x, a*dtt, k, x, N);
            hydro.Mult(x.data(), k.data());
            add(y, b*dtt, k, x, N);
            break;
         }
         case 3: // RK3SSPSolver
            hydro.Mult(x.data(), k.data());
            add(x, dtt, k, y, N);
            hydro.Mult(y.data(), k.data());
            for (int64_t i = 0; i < N; i++) { y[i] = y[i] + dtt*k[i]; }
            for (int64_t i = 0; i < N; i++) { y[i] = (3./4)*x[i] + (1./4)*y[i]; }
            hydro.Mult(y.data(), k.data());
            for (int64_t i = 0; i < N; i++) { y[i] = y[i] + dtt*k[i]; }
            for (int64_t i = 0; i < N; i++) { x[i] = (1./3)*x[i] + (2./3)*y[i]; }
            break;
         case 4:
            hydro.Mult(x.data(), k.data());           // k1
            add(x, dtt/2, k, y, N);
            add(x, dtt/6, k, z, N);
            hydro.Mult(y.data(), k.data());           // k2
            add(x, dtt/2, k, y, N);
            for (int64_t i = 0; i < N; i++) { z[i] += dtt/3*k[i]; }
            hydro.Mult(y.data(), k.data());           // k3
            add(x, dtt, k, y, N);
            for (int64_t i = 0; i < N; i++) { z[i] += dtt/3*k[i]; }
            hydro.Mult(y.data(), k.data());           // k4
            add(z, dtt/6, k, x, N);
            break;
         case 6: // RK6Solver: MFEM's 8-stage 6th-order Verner scheme through ExplicitRKSolver::Step (laghos.cpp:529)
         {
            static const double a6[28] = {.6e-1,
   .1923996296296296296296296296296296296296e-1, .7669337037037037037037037037037037037037e-1,
   .35975e-1, 0., .107925,
   1.318683415233148260919747276431735612861, 0., -5.042058063628562225427761634715637693344, 4.220674648395413964508014358283902080483,
   -41.87259166432751461803757780644346812905, 0., 159.4325621631374917700365669070346830453, -122.1192135650100309202516203389242140663, 5.531743066200053768252631238332999150076,
   -54.43015693531650433250642051294142461271, 0., 207.0672513650184644273657173866509835987, -158.6108137845899991828742424365058599469, 6.991816585950242321992597280791793907096, -.1859723106220323397765171799549294623692e-1,
   -54.66374178728197680241215648050386959351, 0., 207.9528062553893734515824816699834244238, -159.2889574744995071508959805871426654216, 7.018743740796944434698170760964252490817, -.1833878590504572306472782005141738268361e-1, -.5119484997882099077875432497245168395840e-3};
            static const double b6[8] = {.3438957868357036009278820124728322386520e-1, 0., 0., .2582624555633503404659558098586120858767, .4209371189673537150642551514069801967032,
   4.405396469669310170148836816197095664891, -176.4831190242986576151740942499002125029, 172.3641334014150730294022582711902413315};
            if (k6.size() != 8) { k6.assign(8, std::vector<double>(N, 0.0)); }
            hydro.Mult(x.data(), k6[0].data());
            for (int l = 0, i = 1; i < 8; i++)
            {
               add(x, a6[l++]*dtt, k6[0], y, N);
               for (int j = 1; j < i; j++) { const double c = a6[l++]*dtt; for (int64_t q = 0; q < N; q++) { y[q] += c*k6[j][q]; } }
               hydro.Mult(y.data(), k6[i].data());
            }
            for (int i = 0; i < 8; i++) { const double c = b6[i]*dtt; for (int64_t q = 0; q < N; q++) { x[q] += c*k6[i][q]; } }
            break;
         }
         case 7: // RK2AvgSolver::Step
         {
            S0 = x;
            const double *v0 = S0.data() + NV;
            double *dv_dt = k.data() + NV;
            hydro.SolveVelocity(x.data(), k.data());
            for (int64_t i = 0; i < NV; i++) { V[i] = v0[i] + 0.5*dtt*dv_dt[i]; }
            hydro.SolveEnergy(x.data(), V.data(), k.data());
            for (int64_t i = 0; i < NV; i++) { k[i] = V[i]; }
            add(S0, 0.5*dtt, k, x, N);
            hydro.ResetQuadratureData();
            hydro.SolveVelocity(x.data(), k.data());
            for (int64_t i = 0; i < NV; i++) { V[i] = v0[i] + 0.5*dtt*dv_dt[i]; }
            hydro.SolveEnergy(x.data(), V.data(), k.data());
            for (int64_t i = 0; i < NV; i++) { k[i] = V[i]; }
            add(S0, dtt, k, x, N);
            hydro.ResetQuadratureData();
            break;
         }
         default: throw std::runtime_error("Unknown ODE solver type");
      }
      tt += dtt;
   };

   for (int ti = 1; !last_step; ti++)
   {
      if (t + dt >= opt.t_final) { dt = opt.t_final - t; last_step = true; }
      if (steps == opt.max_tsteps) { last_step = true; }
      S_old = S; t_old = t;
      hydro.ResetTimeStepEstimate();
      Step(S, t, dt);
      steps++;
      const double dt_est = hydro.GetTimeStepEstimate(S.data());
      if (dt_est < dt)
      {
         dt *= 0.85;
         if (dt < std::numeric_limits<double>::epsilon()) { throw std::runtime_error("The time step crashed!"); }
         t = t_old; S = S_old;
         hydro.ResetQuadratureData();
         if (opt.verbose) { printf("Repeating step %d\n", ti); }
         if (steps < opt.max_tsteps) { last_step = false; }
         ti--; continue;
      }
      else if (dt_est > 1.25*dt) { dt *= 1.02; }
      const double *e = S.data() + 2*NV;
      const double nrm = std::sqrt(Hydro::dot(e, e, P.ndofs_l2));
      res.e_norm_history.push_back({ti, nrm});
      res.ti_last = ti; res.e_norm = nrm;
      if (opt.verbose && (last_step || (ti % opt.vis_steps) == 0))
      {
         printf("step %5d,\tt = %5.4f,\tdt = %5.6f,\t|e| = %.10e\n", ti, t, dt, nrm);
      }
   }
   res.steps = steps; res.t = t; res.dt = dt;
   res.energy_final = hydro.InternalEnergy(S.data() + 2*NV) + hydro.KineticEnergy(S.data() + NV);
   int stages = 1;
   switch (opt.ode_solver_type) { case 2: stages = 2; break; case 3: stages = 3; break; case 4: stages = 4; break; case 6: stages = 8; break; case 7: stages = 2; break; }
   res.stages = stages;
   hydro.FOM((long long)steps*stages, res.fom);
   res.timer = hydro.timer;
   if (S_out) { *S_out = S; }
   return res;
}

} // namespace oracle

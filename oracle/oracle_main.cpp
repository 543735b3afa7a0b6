// TEST INFRASTRUCTURE — CPU oracle, not a product path (see oracle/README.md).
//
// Command-line twin of the reference driver for the restated CPU path:
// same flags as reference laghos.cpp:181-285 where they matter for the hot path
// (-p -m -rs -ok -ot -oq -s -tf -cfl -cgt -cgm -ms -E0 -iv -vs), plus
//   -serial        use the serial driver's blast energy 0.25 (serial/laghos.cpp:101)
//   -nt N          element-parallel over N threads (stand-in for mpirun -np N)
//   --checks       compare against the known-answer table laghos.cpp:1441-1463
// Mesh names are the stems of the reference's data/*.mesh files.
#include "laghos_oracle.hpp"
#include <cstdlib>

int main(int argc, char **argv)
{
   lagb::ProblemSpec sp;
   oracle::RunOptions opt;
   std::string mesh = "cube01_hex";
   int rs = 2; double E0 = 1.0; bool serial = false, check = false, fom = false;
   opt.verbose = true;
   for (int i = 1; i < argc; i++)
   {
      std::string a = argv[i];
      auto nxt = [&]() { if (i + 1 >= argc) { fprintf(stderr, "missing value for %s\n", a.c_str()); exit(1); } return std::string(argv[++i]); };
      if (a == "-p") { sp.problem = atoi(nxt().c_str()); }
      else if (a == "-m")
      {
         mesh = nxt();
         size_t s = mesh.find_last_of('/'); if (s != std::string::npos) { mesh = mesh.substr(s + 1); }
         s = mesh.find(".mesh"); if (s != std::string::npos) { mesh = mesh.substr(0, s); }
      }
      else if (a == "-rs") { rs = atoi(nxt().c_str()); }
      else if (a == "-ok") { sp.ok = atoi(nxt().c_str()); }
      else if (a == "-ot") { sp.ot = atoi(nxt().c_str()); }
      else if (a == "-oq") { sp.oq = atoi(nxt().c_str()); }
      else if (a == "-s") { opt.ode_solver_type = atoi(nxt().c_str()); }
      else if (a == "-tf") { opt.t_final = atof(nxt().c_str()); }
      else if (a == "-cfl") { opt.cfl = atof(nxt().c_str()); }
      else if (a == "-cgt") { opt.cg_tol = atof(nxt().c_str()); }
      else if (a == "-cgm") { opt.cg_max_iter = atoi(nxt().c_str()); }
      else if (a == "-ms") { opt.max_tsteps = atoi(nxt().c_str()); }
      else if (a == "-vs") { opt.vis_steps = atoi(nxt().c_str()); }
      else if (a == "-E0") { E0 = atof(nxt().c_str()); }
      else if (a == "-iv") { sp.impose_visc = true; }
      else if (a == "-nt") { opt.nthreads = atoi(nxt().c_str()); }
      else if (a == "-serial") { serial = true; }
      else if (a == "-pa") { }
      else if (a == "-f" || a == "--fom") { fom = true; }
      else if (a == "--checks" || a == "-chk") { check = true; }
      else if (a == "-q") { opt.verbose = false; }
      else { fprintf(stderr, "unknown option %s\n", a.c_str()); return 1; }
   }
   std::vector<double> coarse[3]; int dim;
   if (!lagb::named_coarse_mesh(mesh, dim, coarse)) { fprintf(stderr, "unknown mesh %s\n", mesh.c_str()); return 1; }
   sp.dim = dim;
   sp.blast_scale = serial ? 0.25 : E0/pow(2, dim);
   lagb::RectMesh rm; rm.build(dim, coarse, rs);
   lagb::Problem P; P.build(sp, rm); oracle::install_own_tables(P);
   printf("Zones: %d, H1 vdofs: %lld, L2 dofs: %lld, Q1D %d\n", P.NE, (long long)P.h1_vsize(), (long long)P.ndofs_l2, P.Q1D);
   oracle::RunResult r = oracle::run(P, opt);
   printf("final: steps %d ti %d t %.6f dt %.6f |e| %.15e\n", r.steps, r.ti_last, r.t, r.dt, r.e_norm);
   printf("CG (H1) total time: %g\nCG (H1) rate (megadofs x cg_iterations / second): %g\n", r.timer.sw_cgH1, r.fom[1]);
   printf("CG (L2) total time: %g\n", r.timer.sw_cgL2);
   printf("Forces total time: %g\nForces rate (megadofs x timesteps / second): %g\n", r.timer.sw_force, r.fom[2]);
   printf("UpdateQuadData total time: %g\nUpdateQuadData rate (megaquads x timesteps / second): %g\n", r.timer.sw_qdata, r.fom[3]);
   printf("Major kernels total time (seconds): %g\nMajor kernels total rate (megadofs x time steps / second): %g\n", r.fom[4], r.fom[0]);
   (void)fom;
   if (check)
   {
      // reference laghos.cpp:1441-1463 (parallel driver) / serial/laghos.cpp:803-869
      static const double it_norms[2][8][2][2] =
      {
         {
            {{5, 6.546538624534384e+00}, { 27, 7.588576357792927e+00}},
            {{5, 3.508254945225794e+00}, { 15, 2.756444596823211e+00}},
            {{5, 1.020745795651244e+01}, { 59, 1.721590205901898e+01}},
            {{5, 8.000000000000000e+00}, { 16, 8.000000000000000e+00}},
            {{5, 3.446324942352448e+01}, { 18, 3.446844033767240e+01}},
            {{5, 1.030899557252528e+01}, { 36, 1.057362418574309e+01}},
            {{5, 8.039707010835693e+00}, { 36, 8.316970976817373e+00}},
            {{5, 1.514929259650760e+01}, { 25, 1.514931278155159e+01}},
         },
         {
            {{5, 1.198510951452527e+03}, {188, 1.199384410059154e+03}},
            {{5, 6.695818592962833e+00}, { 20, 4.267902387082487e+00}},
            {{5, 2.041491591302486e+01}, { 59, 3.443180411803796e+01}},
            {{5, 1.600000000000000e+01}, { 16, 1.600000000000000e+01}},
            {{5, 6.892649884704898e+01}, { 18, 6.893688067534482e+01}},
            {{5, 2.061984481890964e+01}, { 36, 2.114519664792607e+01}},
            {{5, 1.607988713996459e+01}, { 36, 1.662736010353023e+01}},
            {{5, 3.029858112572883e+01}, { 24, 3.029858832743707e+01}}
         }
      };
      int ok = 0;
      for (int i = 0; i < 2; i++)
      {
         const int it = (int)it_norms[dim-2][sp.problem][i][0];
         const double ref = it_norms[dim-2][sp.problem][i][1];
         for (auto &h : r.e_norm_history)
         {
            if (h.first == it)
            {
               const double rel = fabs(h.second - ref)/fabs(ref);
               printf("check p%d dim%d it %d: %.15e ref %.15e rel %.3e %s\n", sp.problem, dim, it, h.second, ref, rel, rel < 1e-13 ? "OK" : "FAIL");
               ok += (rel < 1e-13);
            }
         }
      }
      return ok == 2 ? 0 : 2;
   }
   return 0;
}

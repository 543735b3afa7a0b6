// The reference driver's self-test (`--checks` / `-chk`, laghos.cpp:904-926, 1403-1474): with
//   -rs 0 -rp 0 -ok 2 -ot 1 -s 4 -tf 0.6 -cfl 0.5 on data/square01_quad.mesh or data/cube01_hex.mesh (-cgt 1e-14 in the
// reference's makefile:199) the norm |e| after two given steps of every problem must match a table to 1e-13 relative
// (both ways), and exactly two checks must have fired by the end of the run.  The table is the reference's data
// (it_norms, laghos.cpp:1441-1463); tests/test_checks_mode.py verifies this transcription against
// tests/golden/checks_table.json and the logic against CPU runs.
#pragma once
#include <cmath>
#include <string>

namespace lagb {

struct CheckEntry { int it; double norm; };
// [dim - 2][problem][k]
static const CheckEntry checks_table[16][2] =
{
   { {5, 6.546538624534384}, {27, 7.588576357792927} },   // dim 2, problem 0
   { {5, 3.508254945225794}, {15, 2.756444596823211} },   // dim 2, problem 1
   { {5, 10.20745795651244}, {59, 17.21590205901898} },   // dim 2, problem 2
   { {5, 8.0}, {16, 8.0} },   // dim 2, problem 3
   { {5, 34.46324942352448}, {18, 34.4684403376724} },   // dim 2, problem 4
   { {5, 10.30899557252528}, {36, 10.57362418574309} },   // dim 2, problem 5
   { {5, 8.039707010835693}, {36, 8.316970976817373} },   // dim 2, problem 6
   { {5, 15.1492925965076}, {25, 15.14931278155159} },   // dim 2, problem 7
   { {5, 1198.510951452527}, {188, 1199.384410059154} },   // dim 3, problem 0
   { {5, 6.695818592962833}, {20, 4.267902387082487} },   // dim 3, problem 1
   { {5, 20.41491591302486}, {59, 34.43180411803796} },   // dim 3, problem 2
   { {5, 16.0}, {16, 16.0} },   // dim 3, problem 3
   { {5, 68.92649884704898}, {18, 68.93688067534482} },   // dim 3, problem 4
   { {5, 20.61984481890964}, {36, 21.14519664792607} },   // dim 3, problem 5
   { {5, 16.07988713996459}, {36, 16.62736010353023} },   // dim 3, problem 6
   { {5, 30.29858112572883}, {24, 30.29858832743707} },   // dim 3, problem 7
};

// preconditions of laghos.cpp:909-920; returns an empty string when they hold
inline std::string checks_preconditions(const std::string &mesh, int dim, int rs, int ok, int ot, int ode_solver_type,
                                        double t_final, double cfl)
{
   if (rs != 0) { return "check: rs, rp"; }
   if (ok != 2) { return "check: order_v"; }
   if (ot != 1) { return "check: order_e"; }
   if (ode_solver_type != 4) { return "check: ode_solver_type"; }
   if (t_final != 0.6) { return "check: t_final"; }
   if (cfl != 0.5) { return "check: cfl"; }
   if (dim != 2 && dim != 3) { return "check: dimension"; }
   if (!(mesh == "square01_quad" || mesh == "cube01_hex" || mesh == "default" || mesh == "default_2d")) { return "check: mesh_file"; }
   return "";
}

// Checks(ti, nrm, chk) of laghos.cpp:1403-1474: 0 = no entry for this step, 1 = entry matched (chk incremented),
// -1 = mismatch ("P<problem>, #<it>"); eps = 1e-13 in the reference
inline int checks_step(int dim, int problem, int ti, double nrm, double eps, int &chk, std::string &msg)
{
   if (dim < 2 || dim > 3 || problem < 0 || problem > 7) { return 0; }
   int rc = 0;
   for (int k = 0; k < 2; k++)
   {
      const CheckEntry &c = checks_table[(dim - 2)*8 + problem][k];
      if (c.it != ti) { continue; }
      chk++;
      if (!(std::fabs(nrm) > eps && std::fabs(c.norm) > eps)) { msg = "One value is near zero!"; return -1; }
      const double err_a = std::fabs((nrm - c.norm)/nrm), err_v = std::fabs((nrm - c.norm)/c.norm);
      if (!(std::fmax(err_a, err_v) < eps)) { msg = "P" + std::to_string(problem) + ", #" + std::to_string(ti); return -1; }
      rc = 1;
   }
   return rc;
}

} // namespace lagb

// TEST INFRASTRUCTURE ONLY -- CPU oracle (see fv3_oracle.hpp header).
// Restates model/a2b_edge.F90 of the reference: a2b_ord4 (:47-327), a2b_ord2 (:329-450),
// extrap_corner (:452-462); and great_circle_dist, model/fv_grid_utils.F90:1974-1995.
#include "fv3_oracle.hpp"

namespace fv3o {

static const double r3 = 1. / 3.;
static const double a1 = 0.5625, a2 = -0.0625;   // a2b_edge.F90:34-35
static const double b1 = 7. / 12., b2 = -1. / 12.;  // a2b_edge.F90:39-40

// fv_grid_utils.F90:1974-1995 (haversine form)
double great_circle_dist(const double q1[2], const double q2[2], double radius) {
  double p1 = (q1[1] - q2[1]) / 2.;
  double p2 = (q1[0] - q2[0]) / 2.;
  double s1 = std::sin(p1), s2 = std::sin(p2);
  double beta = std::asin(std::sqrt(s1 * s1 + std::cos(q1[1]) * std::cos(q2[1]) * (s2 * s2))) * 2.;
  return radius * beta;
}

static double extrap_corner(const double p0[2], const Grid& g, int i1, int j1, int i2, int j2, double q1, double q2) {
  double p1[2] = {g.agrid(i1, j1, 1), g.agrid(i1, j1, 2)};
  double p2[2] = {g.agrid(i2, j2, 1), g.agrid(i2, j2, 2)};
  double x1 = great_circle_dist(p1, p0, 1.0);
  double x2 = great_circle_dist(p2, p0, 1.0);
  return q1 + x1 / (x2 - x1) * (q1 - q2);
}

// a2b_edge.F90:329-450 a2b_ord2 (replace absent / false): A-grid -> cell corners (is:ie+1, js:je+1), second order
void a2b_ord2(V2 qin, V2 qout, const Grid& g, const Bd& bd) {
  const int is = bd.is, ie = bd.ie, js = bd.js, je = bd.je, npx = bd.npx, npy = bd.npy;
  const double r3 = 1. / 3.;
  if (bd.grid_type < 3 && !bd.bounded_domain) {
    const int is1 = std::max(1, is - 1), js1 = std::max(1, js - 1), is2 = std::max(2, is), js2 = std::max(2, js);
    const int ie1 = std::min(npx - 1, ie + 1), je1 = std::min(npy - 1, je + 1);
    L1 q1(is - 1, ie + 1), q2(js - 1, je + 1);
    for (int j = js2; j <= je1; j++)
      for (int i = is2; i <= ie1; i++) qout(i, j) = 0.25 * (qin(i - 1, j - 1) + qin(i, j - 1) + qin(i - 1, j) + qin(i, j));
    if (bd.sw_corner) qout(1, 1) = r3 * (qin(1, 1) + qin(1, 0) + qin(0, 1));
    if (bd.se_corner) qout(npx, 1) = r3 * (qin(npx - 1, 1) + qin(npx - 1, 0) + qin(npx, 1));
    if (bd.ne_corner) qout(npx, npy) = r3 * (qin(npx - 1, npy - 1) + qin(npx, npy - 1) + qin(npx - 1, npy));
    if (bd.nw_corner) qout(1, npy) = r3 * (qin(1, npy - 1) + qin(0, npy - 1) + qin(1, npy));
    if (is == 1) {
      for (int j = js1; j <= je1; j++) q2(j) = 0.5 * (qin(0, j) + qin(1, j));
      for (int j = js2; j <= je1; j++) qout(1, j) = g.edge_w[j - 1] * q2(j - 1) + (1. - g.edge_w[j - 1]) * q2(j);
    }
    if (ie + 1 == npx) {
      for (int j = js1; j <= je1; j++) q2(j) = 0.5 * (qin(npx - 1, j) + qin(npx, j));
      for (int j = js2; j <= je1; j++) qout(npx, j) = g.edge_e[j - 1] * q2(j - 1) + (1. - g.edge_e[j - 1]) * q2(j);
    }
    if (js == 1) {
      for (int i = is1; i <= ie1; i++) q1(i) = 0.5 * (qin(i, 0) + qin(i, 1));
      for (int i = is2; i <= ie1; i++) qout(i, 1) = g.edge_s[i - 1] * q1(i - 1) + (1. - g.edge_s[i - 1]) * q1(i);
    }
    if (je + 1 == npy) {
      for (int i = is1; i <= ie1; i++) q1(i) = 0.5 * (qin(i, npy - 1) + qin(i, npy));
      for (int i = is2; i <= ie1; i++) qout(i, npy) = g.edge_n[i - 1] * q1(i - 1) + (1. - g.edge_n[i - 1]) * q1(i);
    }
  } else {
    for (int j = js; j <= je + 1; j++)
      for (int i = is; i <= ie + 1; i++) qout(i, j) = 0.25 * (qin(i - 1, j - 1) + qin(i, j - 1) + qin(i - 1, j) + qin(i, j));
  }
}

void a2b_ord4(V2 qin, V2 qout, const Grid& g, const Bd& bd, bool replace) {
  const int is = bd.is, ie = bd.ie, js = bd.js, je = bd.je, ng = bd.ng, npx = bd.npx, npy = bd.npy;
  const double c1 = 2. / 3., c2 = -1. / 6.;
  L2 qx(is, ie + 1, js - ng, je + ng), qy(is - ng, ie + ng, js, je + 1);
  L2 qxx(is - ng, ie + ng, js - ng, je + ng), qyy(is - ng, ie + ng, js - ng, je + ng);
  L1 q1(is - 1, ie + 1), q2(js - 1, je + 1);
  const V2& dxa = g.dxa; const V2& dya = g.dya;

  if (bd.grid_type < 3) {
    const int is1 = std::max(1, is - 1), js1 = std::max(1, js - 1), is2 = std::max(2, is), js2 = std::max(2, js);
    const int ie1 = std::min(npx - 1, ie + 1), je1 = std::min(npy - 1, je + 1);
    if (bd.bounded_domain) {
      for (int j = js - 2; j <= je + 2; j++)
        for (int i = is; i <= ie + 1; i++) qx(i, j) = b2 * (qin(i - 2, j) + qin(i + 1, j)) + b1 * (qin(i - 1, j) + qin(i, j));
    } else {
      if (bd.sw_corner) {
        double p0[2] = {g.grid(1, 1, 1), g.grid(1, 1, 2)};
        qout(1, 1) = (extrap_corner(p0, g, 1, 1, 2, 2, qin(1, 1), qin(2, 2)) +
                      extrap_corner(p0, g, 0, 1, -1, 2, qin(0, 1), qin(-1, 2)) +
                      extrap_corner(p0, g, 1, 0, 2, -1, qin(1, 0), qin(2, -1))) * r3;
      }
      if (bd.se_corner) {
        double p0[2] = {g.grid(npx, 1, 1), g.grid(npx, 1, 2)};
        qout(npx, 1) = (extrap_corner(p0, g, npx - 1, 1, npx - 2, 2, qin(npx - 1, 1), qin(npx - 2, 2)) +
                        extrap_corner(p0, g, npx - 1, 0, npx - 2, -1, qin(npx - 1, 0), qin(npx - 2, -1)) +
                        extrap_corner(p0, g, npx, 1, npx + 1, 2, qin(npx, 1), qin(npx + 1, 2))) * r3;
      }
      if (bd.ne_corner) {
        double p0[2] = {g.grid(npx, npy, 1), g.grid(npx, npy, 2)};
        qout(npx, npy) = (extrap_corner(p0, g, npx - 1, npy - 1, npx - 2, npy - 2, qin(npx - 1, npy - 1), qin(npx - 2, npy - 2)) +
                          extrap_corner(p0, g, npx, npy - 1, npx + 1, npy - 2, qin(npx, npy - 1), qin(npx + 1, npy - 2)) +
                          extrap_corner(p0, g, npx - 1, npy, npx - 2, npy + 1, qin(npx - 1, npy), qin(npx - 2, npy + 1))) * r3;
      }
      if (bd.nw_corner) {
        double p0[2] = {g.grid(1, npy, 1), g.grid(1, npy, 2)};
        qout(1, npy) = (extrap_corner(p0, g, 1, npy - 1, 2, npy - 2, qin(1, npy - 1), qin(2, npy - 2)) +
                        extrap_corner(p0, g, 0, npy - 1, -1, npy - 2, qin(0, npy - 1), qin(-1, npy - 2)) +
                        extrap_corner(p0, g, 1, npy, 2, npy + 1, qin(1, npy), qin(2, npy + 1))) * r3;
      }
      // X-interior
      for (int j = std::max(1, js - 2); j <= std::min(npy - 1, je + 2); j++)
        for (int i = std::max(3, is); i <= std::min(npx - 2, ie + 1); i++)
          qx(i, j) = b2 * (qin(i - 2, j) + qin(i + 1, j)) + b1 * (qin(i - 1, j) + qin(i, j));
      if (is == 1) {  // West edges
        for (int j = js1; j <= je1; j++) q2(j) = (qin(0, j) * dxa(1, j) + qin(1, j) * dxa(0, j)) / (dxa(0, j) + dxa(1, j));
        for (int j = js2; j <= je1; j++) qout(1, j) = g.edge_w[j - 1] * q2(j - 1) + (1. - g.edge_w[j - 1]) * q2(j);
        for (int j = std::max(1, js - 2); j <= std::min(npy - 1, je + 2); j++) {
          double g_in = dxa(2, j) / dxa(1, j);
          double g_ou = dxa(-1, j) / dxa(0, j);
          qx(1, j) = 0.5 * (((2. + g_in) * qin(1, j) - qin(2, j)) / (1. + g_in) + ((2. + g_ou) * qin(0, j) - qin(-1, j)) / (1. + g_ou));
          qx(2, j) = (3. * (g_in * qin(1, j) + qin(2, j)) - (g_in * qx(1, j) + qx(3, j))) / (2. + 2. * g_in);
        }
      }
      if ((ie + 1) == npx) {  // East edges
        for (int j = js1; j <= je1; j++)
          q2(j) = (qin(npx - 1, j) * dxa(npx, j) + qin(npx, j) * dxa(npx - 1, j)) / (dxa(npx - 1, j) + dxa(npx, j));
        for (int j = js2; j <= je1; j++) qout(npx, j) = g.edge_e[j - 1] * q2(j - 1) + (1. - g.edge_e[j - 1]) * q2(j);
        for (int j = std::max(1, js - 2); j <= std::min(npy - 1, je + 2); j++) {
          double g_in = dxa(npx - 2, j) / dxa(npx - 1, j);
          double g_ou = dxa(npx + 1, j) / dxa(npx, j);
          qx(npx, j) = 0.5 * (((2. + g_in) * qin(npx - 1, j) - qin(npx - 2, j)) / (1. + g_in) +
                              ((2. + g_ou) * qin(npx, j) - qin(npx + 1, j)) / (1. + g_ou));
          qx(npx - 1, j) = (3. * (qin(npx - 2, j) + g_in * qin(npx - 1, j)) - (g_in * qx(npx, j) + qx(npx - 2, j))) / (2. + 2. * g_in);
        }
      }
    }
    // Y-interior
    if (bd.bounded_domain) {
      for (int j = js; j <= je + 1; j++)
        for (int i = is - 2; i <= ie + 2; i++) qy(i, j) = b2 * (qin(i, j - 2) + qin(i, j + 1)) + b1 * (qin(i, j - 1) + qin(i, j));
    } else {
      for (int j = std::max(3, js); j <= std::min(npy - 2, je + 1); j++)
        for (int i = std::max(1, is - 2); i <= std::min(npx - 1, ie + 2); i++)
          qy(i, j) = b2 * (qin(i, j - 2) + qin(i, j + 1)) + b1 * (qin(i, j - 1) + qin(i, j));
      if (js == 1) {  // South edges
        for (int i = is1; i <= ie1; i++) q1(i) = (qin(i, 0) * dya(i, 1) + qin(i, 1) * dya(i, 0)) / (dya(i, 0) + dya(i, 1));
        for (int i = is2; i <= ie1; i++) qout(i, 1) = g.edge_s[i - 1] * q1(i - 1) + (1. - g.edge_s[i - 1]) * q1(i);
        for (int i = std::max(1, is - 2); i <= std::min(npx - 1, ie + 2); i++) {
          double g_in = dya(i, 2) / dya(i, 1);
          double g_ou = dya(i, -1) / dya(i, 0);
          qy(i, 1) = 0.5 * (((2. + g_in) * qin(i, 1) - qin(i, 2)) / (1. + g_in) + ((2. + g_ou) * qin(i, 0) - qin(i, -1)) / (1. + g_ou));
          qy(i, 2) = (3. * (g_in * qin(i, 1) + qin(i, 2)) - (g_in * qy(i, 1) + qy(i, 3))) / (2. + 2. * g_in);
        }
      }
      if ((je + 1) == npy) {  // North edges
        for (int i = is1; i <= ie1; i++)
          q1(i) = (qin(i, npy - 1) * dya(i, npy) + qin(i, npy) * dya(i, npy - 1)) / (dya(i, npy - 1) + dya(i, npy));
        for (int i = is2; i <= ie1; i++) qout(i, npy) = g.edge_n[i - 1] * q1(i - 1) + (1. - g.edge_n[i - 1]) * q1(i);
        for (int i = std::max(1, is - 2); i <= std::min(npx - 1, ie + 2); i++) {
          double g_in = dya(i, npy - 2) / dya(i, npy - 1);
          double g_ou = dya(i, npy + 1) / dya(i, npy);
          qy(i, npy) = 0.5 * (((2. + g_in) * qin(i, npy - 1) - qin(i, npy - 2)) / (1. + g_in) +
                              ((2. + g_ou) * qin(i, npy) - qin(i, npy + 1)) / (1. + g_ou));
          qy(i, npy - 1) = (3. * (qin(i, npy - 2) + g_in * qin(i, npy - 1)) - (g_in * qy(i, npy) + qy(i, npy - 2))) / (2. + 2. * g_in);
        }
      }
    }
    if (bd.bounded_domain) {
      for (int j = js; j <= je + 1; j++)
        for (int i = is; i <= ie + 1; i++) qxx(i, j) = a2 * (qx(i, j - 2) + qx(i, j + 1)) + a1 * (qx(i, j - 1) + qx(i, j));
      for (int j = js; j <= je + 1; j++) {
        for (int i = is; i <= ie + 1; i++) qyy(i, j) = a2 * (qy(i - 2, j) + qy(i + 1, j)) + a1 * (qy(i - 1, j) + qy(i, j));
        for (int i = is; i <= ie + 1; i++) qout(i, j) = 0.5 * (qxx(i, j) + qyy(i, j));
      }
    } else {
      for (int j = std::max(3, js); j <= std::min(npy - 2, je + 1); j++)
        for (int i = std::max(2, is); i <= std::min(npx - 1, ie + 1); i++)
          qxx(i, j) = a2 * (qx(i, j - 2) + qx(i, j + 1)) + a1 * (qx(i, j - 1) + qx(i, j));
      if (js == 1)
        for (int i = std::max(2, is); i <= std::min(npx - 1, ie + 1); i++)
          qxx(i, 2) = c1 * (qx(i, 1) + qx(i, 2)) + c2 * (qout(i, 1) + qxx(i, 3));
      if ((je + 1) == npy)
        for (int i = std::max(2, is); i <= std::min(npx - 1, ie + 1); i++)
          qxx(i, npy - 1) = c1 * (qx(i, npy - 2) + qx(i, npy - 1)) + c2 * (qout(i, npy) + qxx(i, npy - 2));
      for (int j = std::max(2, js); j <= std::min(npy - 1, je + 1); j++) {
        for (int i = std::max(3, is); i <= std::min(npx - 2, ie + 1); i++)
          qyy(i, j) = a2 * (qy(i - 2, j) + qy(i + 1, j)) + a1 * (qy(i - 1, j) + qy(i, j));
        if (is == 1) qyy(2, j) = c1 * (qy(1, j) + qy(2, j)) + c2 * (qout(1, j) + qyy(3, j));
        if ((ie + 1) == npx) qyy(npx - 1, j) = c1 * (qy(npx - 2, j) + qy(npx - 1, j)) + c2 * (qout(npx, j) + qyy(npx - 2, j));
        for (int i = std::max(2, is); i <= std::min(npx - 1, ie + 1); i++) qout(i, j) = 0.5 * (qxx(i, j) + qyy(i, j));
      }
    }
  } else {
    for (int j = js - 2; j <= je + 2; j++)
      for (int i = is; i <= ie + 1; i++) qx(i, j) = b1 * (qin(i - 1, j) + qin(i, j)) + b2 * (qin(i - 2, j) + qin(i + 1, j));
    for (int j = js; j <= je + 1; j++)
      for (int i = is - 2; i <= ie + 2; i++) qy(i, j) = b1 * (qin(i, j - 1) + qin(i, j)) + b2 * (qin(i, j - 2) + qin(i, j + 1));
    for (int j = js; j <= je + 1; j++)
      for (int i = is; i <= ie + 1; i++)
        qout(i, j) = 0.5 * (a1 * (qx(i, j - 1) + qx(i, j) + qy(i - 1, j) + qy(i, j)) +
                            a2 * (qx(i, j - 2) + qx(i, j + 1) + qy(i - 2, j) + qy(i + 1, j)));
  }
  if (replace)
    for (int j = js; j <= je + 1; j++) for (int i = is; i <= ie + 1; i++) qin(i, j) = qout(i, j);
}

void a2b_ord2(V2 qin, V2 qout, const Grid& g, const Bd& bd, bool replace) {
  const int is = bd.is, ie = bd.ie, js = bd.js, je = bd.je, npx = bd.npx, npy = bd.npy;
  if (bd.grid_type < 3 && !bd.bounded_domain) {
    const int is1 = std::max(1, is - 1), js1 = std::max(1, js - 1), is2 = std::max(2, is), js2 = std::max(2, js);
    const int ie1 = std::min(npx - 1, ie + 1), je1 = std::min(npy - 1, je + 1);
    L1 q1(1, npx), q2(1, npy);
    for (int j = js2; j <= je1; j++)
      for (int i = is2; i <= ie1; i++) qout(i, j) = 0.25 * (qin(i - 1, j - 1) + qin(i, j - 1) + qin(i - 1, j) + qin(i, j));
    if (bd.sw_corner) qout(1, 1) = r3 * (qin(1, 1) + qin(1, 0) + qin(0, 1));
    if (bd.se_corner) qout(npx, 1) = r3 * (qin(npx - 1, 1) + qin(npx - 1, 0) + qin(npx, 1));
    if (bd.ne_corner) qout(npx, npy) = r3 * (qin(npx - 1, npy - 1) + qin(npx, npy - 1) + qin(npx - 1, npy));
    if (bd.nw_corner) qout(1, npy) = r3 * (qin(1, npy - 1) + qin(0, npy - 1) + qin(1, npy));
    if (is == 1) {
      for (int j = js1; j <= je1; j++) q2(j) = 0.5 * (qin(0, j) + qin(1, j));
      for (int j = js2; j <= je1; j++) qout(1, j) = g.edge_w[j - 1] * q2(j - 1) + (1. - g.edge_w[j - 1]) * q2(j);
    }
    if ((ie + 1) == npx) {
      for (int j = js1; j <= je1; j++) q2(j) = 0.5 * (qin(npx - 1, j) + qin(npx, j));
      for (int j = js2; j <= je1; j++) qout(npx, j) = g.edge_e[j - 1] * q2(j - 1) + (1. - g.edge_e[j - 1]) * q2(j);
    }
    if (js == 1) {
      for (int i = is1; i <= ie1; i++) q1(i) = 0.5 * (qin(i, 0) + qin(i, 1));
      for (int i = is2; i <= ie1; i++) qout(i, 1) = g.edge_s[i - 1] * q1(i - 1) + (1. - g.edge_s[i - 1]) * q1(i);
    }
    if ((je + 1) == npy) {
      for (int i = is1; i <= ie1; i++) q1(i) = 0.5 * (qin(i, npy - 1) + qin(i, npy));
      for (int i = is2; i <= ie1; i++) qout(i, npy) = g.edge_n[i - 1] * q1(i - 1) + (1. - g.edge_n[i - 1]) * q1(i);
    }
  } else {
    for (int j = js; j <= je + 1; j++)
      for (int i = is; i <= ie + 1; i++) qout(i, j) = 0.25 * (qin(i - 1, j - 1) + qin(i, j - 1) + qin(i - 1, j) + qin(i, j));
  }
  if (replace)
    for (int j = js; j <= je + 1; j++) for (int i = is; i <= ie + 1; i++) qin(i, j) = qout(i, j);
}

}  // namespace fv3o

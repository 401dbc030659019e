// TEST INFRASTRUCTURE ONLY -- CPU oracle (see fv3_oracle.hpp header).
// Restates model/tp_core.F90 of the reference: fv_tp_2d (:85-241),
// copy_corners (:245-322), xppm (:324-712), yppm (:715-1152),
// pert_ppm (:1206-1264), deln_flux (:1267-1447).  Same loop bounds and
// floating-point operation order as the Fortran.
#include "fv3_oracle.hpp"

namespace fv3o {

// tp_core.F90:35-70
static const double ppm_fac = 1.5;
static const double r3 = 1. / 3.;
static const double near_zero = 1.E-25;
static const double r12 = 1. / 12.;
static const double s11 = 11. / 14., s14 = 4. / 7., s15 = 3. / 14.;
static const double c1 = -2. / 14.;
static const double c2 = 11. / 14.;
static const double c3 = 5. / 14.;
static const double p1 = 7. / 12.;
static const double p2 = -1. / 12.;

// tp_core.F90:245-322
void copy_corners(V2 q, int npx, int npy, int dir, const Bd& bd) {
  const int ng = bd.ng;
  if (bd.bounded_domain) return;
  if (dir == 1) {
    if (bd.sw_corner)
      for (int j = 1 - ng; j <= 0; j++)
        for (int i = 1 - ng; i <= 0; i++) q(i, j) = q(j, 1 - i);
    if (bd.se_corner)
      for (int j = 1 - ng; j <= 0; j++)
        for (int i = npx; i <= npx + ng - 1; i++) q(i, j) = q(npy - j, i - npx + 1);
    if (bd.ne_corner)
      for (int j = npy; j <= npy + ng - 1; j++)
        for (int i = npx; i <= npx + ng - 1; i++) q(i, j) = q(j, 2 * npx - 1 - i);
    if (bd.nw_corner)
      for (int j = npy; j <= npy + ng - 1; j++)
        for (int i = 1 - ng; i <= 0; i++) q(i, j) = q(npy - j, i - 1 + npx);
  } else if (dir == 2) {
    if (bd.sw_corner)
      for (int j = 1 - ng; j <= 0; j++)
        for (int i = 1 - ng; i <= 0; i++) q(i, j) = q(1 - j, i);
    if (bd.se_corner)
      for (int j = 1 - ng; j <= 0; j++)
        for (int i = npx; i <= npx + ng - 1; i++) q(i, j) = q(npy + j - 1, npx - i);
    if (bd.ne_corner)
      for (int j = npy; j <= npy + ng - 1; j++)
        for (int i = npx; i <= npx + ng - 1; i++) q(i, j) = q(2 * npy - 1 - j, i);
    if (bd.nw_corner)
      for (int j = npy; j <= npy + ng - 1; j++)
        for (int i = 1 - ng; i <= 0; i++) q(i, j) = q(j + 1 - npx, npy - i);
  }
}

// tp_core.F90:1206-1264
void pert_ppm(int im, const double* a0, double* al, double* ar, int iv) {
  if (iv == 0) {
    for (int i = 0; i < im; i++) {
      if (a0[i] <= 0.) {
        al[i] = 0.; ar[i] = 0.;
      } else {
        double a4 = -3. * (ar[i] + al[i]);
        double da1 = ar[i] - al[i];
        if (std::fabs(da1) < -a4) {
          double fmin = a0[i] + 0.25 / a4 * (da1 * da1) + a4 * r12;
          if (fmin < 0.) {
            if (ar[i] > 0. && al[i] > 0.) { ar[i] = 0.; al[i] = 0.; }
            else if (da1 > 0.) ar[i] = -2. * al[i];
            else al[i] = -2. * ar[i];
          }
        }
      }
    }
  } else {
    for (int i = 0; i < im; i++) {
      if (al[i] * ar[i] < 0.) {
        double da1 = al[i] - ar[i];
        double da2 = da1 * da1;
        double a6da = 3. * (al[i] + ar[i]) * da1;
        if (a6da < -da2) ar[i] = -2. * al[i];
        else if (a6da > da2) al[i] = -2. * ar[i];
      } else {
        al[i] = 0.; ar[i] = 0.;
      }
    }
  }
}

// Final flux formula shared by the monotone family, tp_core.F90:701-707
#define PPM_FLUX_MONO(qm1, q0, cc, blm, brm, bl0, br0)                       \
  ((cc) > 0. ? (qm1) + (1. - (cc)) * ((brm) - (cc) * ((blm) + (brm)))        \
             : (q0) + (1. + (cc)) * ((bl0) + (cc) * ((bl0) + (br0))))

// tp_core.F90:324-712
void xppm(V2 flux, V2 q, V2 c, int iord, int is, int ie, int isd, int ied, int jfirst, int jlast,
          int jsd, int jed, int npx, int npy, V2 dxa, bool bounded_domain, int grid_type, double lim_fac) {
  (void)jsd; (void)jed; (void)npy;
  L1 bl(is - 1, ie + 1), br(is - 1, ie + 1), b0(is - 1, ie + 1), a4(is - 1, ie + 1), da1(is - 1, ie + 1);
  L1 q1(isd, ied);
  L1 fx1(is, ie + 1), xt1(is, ie + 1);
  LB1 ext5(is - 1, ie + 1), ext6(is - 1, ie + 1), smt5(is - 1, ie + 1), smt6(is - 1, ie + 1);
  LB1 hi5(is, ie + 1), hi6(is, ie + 1);
  L1 al(is - 1, ie + 2), dm(is - 2, ie + 2), dq(is - 3, ie + 2);
  int is1, ie3, ie1;
  const bool cube = (!bounded_domain && grid_type < 3);
  if (cube) {
    is1 = std::max(3, is - 1); ie3 = std::min(npx - 2, ie + 2); ie1 = std::min(npx - 3, ie + 1);
  } else {
    is1 = is - 1; ie3 = ie + 2; ie1 = ie + 1;
  }
  const int mord = std::abs(iord);

  for (int j = jfirst; j <= jlast; j++) {
    for (int i = isd; i <= ied; i++) q1(i) = q(i, j);

    if (iord < 7) {
      for (int i = is1; i <= ie3; i++) al(i) = p1 * (q1(i - 1) + q1(i)) + p2 * (q1(i - 2) + q1(i + 1));
      if (cube) {
        if (is == 1) {
          al(0) = c1 * q1(-2) + c2 * q1(-1) + c3 * q1(0);
          al(1) = 0.5 * (((2. * dxa(0, j) + dxa(-1, j)) * q1(0) - dxa(0, j) * q1(-1)) / (dxa(-1, j) + dxa(0, j)) +
                         ((2. * dxa(1, j) + dxa(2, j)) * q1(1) - dxa(1, j) * q1(2)) / (dxa(1, j) + dxa(2, j)));
          al(2) = c3 * q1(1) + c2 * q1(2) + c1 * q1(3);
        }
        if ((ie + 1) == npx) {
          al(npx - 1) = c1 * q1(npx - 3) + c2 * q1(npx - 2) + c3 * q1(npx - 1);
          al(npx) = 0.5 * (((2. * dxa(npx - 1, j) + dxa(npx - 2, j)) * q1(npx - 1) - dxa(npx - 1, j) * q1(npx - 2)) /
                               (dxa(npx - 2, j) + dxa(npx - 1, j)) +
                           ((2. * dxa(npx, j) + dxa(npx + 1, j)) * q1(npx) - dxa(npx, j) * q1(npx + 1)) /
                               (dxa(npx, j) + dxa(npx + 1, j)));
          al(npx + 1) = c3 * q1(npx) + c2 * q1(npx + 1) + c1 * q1(npx + 2);
        }
      }
      if (iord < 0)
        for (int i = is - 1; i <= ie + 2; i++) al(i) = std::max(0., al(i));

      if (mord == 1) {
        for (int i = is - 1; i <= ie + 1; i++) {
          bl(i) = al(i) - q1(i); br(i) = al(i + 1) - q1(i); b0(i) = bl(i) + br(i);
          smt5(i) = std::fabs(lim_fac * b0(i)) < std::fabs(bl(i) - br(i));
        }
        for (int i = is; i <= ie + 1; i++) {
          if (c(i, j) > 0.) { fx1(i) = (1. - c(i, j)) * (br(i - 1) - c(i, j) * b0(i - 1)); flux(i, j) = q1(i - 1); }
          else { fx1(i) = (1. + c(i, j)) * (bl(i) + c(i, j) * b0(i)); flux(i, j) = q1(i); }
          if (smt5(i - 1) || smt5(i)) flux(i, j) = flux(i, j) + fx1(i);
        }
      } else if (mord == 2) {
        for (int i = is; i <= ie + 1; i++) {
          double xt = c(i, j), qtmp;
          if (xt > 0.) {
            qtmp = q1(i - 1);
            flux(i, j) = qtmp + (1. - xt) * (al(i) - qtmp - xt * (al(i - 1) + al(i) - (qtmp + qtmp)));
          } else {
            qtmp = q1(i);
            flux(i, j) = qtmp + (1. + xt) * (al(i) - qtmp + xt * (al(i) + al(i + 1) - (qtmp + qtmp)));
          }
        }
      } else if (mord == 3) {
        for (int i = is - 1; i <= ie + 1; i++) {
          bl(i) = al(i) - q1(i); br(i) = al(i + 1) - q1(i); b0(i) = bl(i) + br(i);
          double x0 = std::fabs(b0(i)), xt = std::fabs(bl(i) - br(i));
          smt5(i) = x0 < xt; smt6(i) = 3. * x0 < xt;
        }
        for (int i = is; i <= ie + 1; i++) {
          xt1(i) = c(i, j);
          if (xt1(i) > 0.) {
            if (smt5(i - 1) || smt6(i)) flux(i, j) = q1(i - 1) + (1. - xt1(i)) * (br(i - 1) - xt1(i) * b0(i - 1));
            else flux(i, j) = q1(i - 1);
          } else {
            if (smt6(i - 1) || smt5(i)) flux(i, j) = q1(i) + (1. + xt1(i)) * (bl(i) + xt1(i) * b0(i));
            else flux(i, j) = q1(i);
          }
        }
      } else if (mord == 4) {
        for (int i = is - 1; i <= ie + 1; i++) {
          bl(i) = al(i) - q1(i); br(i) = al(i + 1) - q1(i); b0(i) = bl(i) + br(i);
          double x0 = std::fabs(b0(i)), xt = std::fabs(bl(i) - br(i));
          smt5(i) = x0 < xt; smt6(i) = 3. * x0 < xt;
        }
        for (int i = is; i <= ie + 1; i++) {
          xt1(i) = c(i, j);
          hi5(i) = smt5(i - 1) && smt5(i);
          hi6(i) = smt6(i - 1) || smt6(i);
          hi5(i) = hi5(i) || hi6(i);
        }
        for (int i = is; i <= ie + 1; i++) {
          if (xt1(i) > 0.) { fx1(i) = (1. - xt1(i)) * (br(i - 1) - xt1(i) * b0(i - 1)); flux(i, j) = q1(i - 1); }
          else { fx1(i) = (1. + xt1(i)) * (bl(i) + xt1(i) * b0(i)); flux(i, j) = q1(i); }
          if (hi5(i)) flux(i, j) = flux(i, j) + fx1(i);
        }
      } else {
        if (iord == 5) {
          for (int i = is - 1; i <= ie + 1; i++) {
            bl(i) = al(i) - q1(i); br(i) = al(i + 1) - q1(i); b0(i) = bl(i) + br(i);
            smt5(i) = bl(i) * br(i) < 0.;
          }
        } else {
          if (iord == -5) {
            for (int i = is - 1; i <= ie + 1; i++) {
              bl(i) = al(i) - q1(i); br(i) = al(i + 1) - q1(i); b0(i) = bl(i) + br(i);
              smt5(i) = bl(i) * br(i) < 0.;
              da1(i) = br(i) - bl(i);
              a4(i) = -3. * b0(i);
            }
            for (int i = is - 1; i <= ie + 1; i++) {
              if (std::fabs(da1(i)) < -a4(i)) {
                if (q1(i) + 0.25 / a4(i) * (da1(i) * da1(i)) + a4(i) * r12 < 0.) {
                  if (!smt5(i)) { br(i) = 0.; bl(i) = 0.; b0(i) = 0.; }
                  else if (da1(i) > 0.) { br(i) = -2. * bl(i); b0(i) = -bl(i); }
                  else { bl(i) = -2. * br(i); b0(i) = -br(i); }
                }
              }
            }
          } else {
            for (int i = is - 1; i <= ie + 1; i++) {
              bl(i) = al(i) - q1(i); br(i) = al(i + 1) - q1(i); b0(i) = bl(i) + br(i);
              smt5(i) = 3. * std::fabs(b0(i)) < std::fabs(bl(i) - br(i));
            }
          }
          if (cube) {
            if (is == 1) { smt5(0) = bl(0) * br(0) < 0.; smt5(1) = bl(1) * br(1) < 0.; }
            if ((ie + 1) == npx) { smt5(npx - 1) = bl(npx - 1) * br(npx - 1) < 0.; smt5(npx) = bl(npx) * br(npx) < 0.; }
          }
        }
        for (int i = is; i <= ie + 1; i++) {
          if (c(i, j) > 0.) { fx1(i) = (1. - c(i, j)) * (br(i - 1) - c(i, j) * b0(i - 1)); flux(i, j) = q1(i - 1); }
          else { fx1(i) = (1. + c(i, j)) * (bl(i) + c(i, j) * b0(i)); flux(i, j) = q1(i); }
          if (smt5(i - 1) || smt5(i)) flux(i, j) = flux(i, j) + fx1(i);
        }
      }
      continue;  // goto 666
    }

    // Monotonic constraints (iord >= 7), tp_core.F90:563-708
    for (int i = is - 2; i <= ie + 2; i++) {
      double xt = 0.25 * (q1(i + 1) - q1(i - 1));
      dm(i) = fsign(std::min(std::min(std::fabs(xt), max3(q1(i - 1), q1(i), q1(i + 1)) - q1(i)),
                             q1(i) - min3(q1(i - 1), q1(i), q1(i + 1))), xt);
    }
    for (int i = is1; i <= ie1 + 1; i++) al(i) = 0.5 * (q1(i - 1) + q1(i)) + r3 * (dm(i - 1) - dm(i));

    if (iord == 8) {
      for (int i = is1; i <= ie1; i++) {
        double xt = 2. * dm(i);
        bl(i) = -fsign(std::min(std::fabs(xt), std::fabs(al(i) - q1(i))), xt);
        br(i) = fsign(std::min(std::fabs(xt), std::fabs(al(i + 1) - q1(i))), xt);
      }
    } else if (iord == 10) {
      for (int i = is1 - 2; i <= ie1 + 1; i++) dq(i) = 2. * (q1(i + 1) - q1(i));
      for (int i = is1; i <= ie1; i++) {
        bl(i) = al(i) - q1(i);
        br(i) = al(i + 1) - q1(i);
        if (std::fabs(dm(i - 1)) + std::fabs(dm(i)) + std::fabs(dm(i + 1)) < near_zero) {
          bl(i) = 0.; br(i) = 0.;
        } else if (std::fabs(3. * (bl(i) + br(i))) > std::fabs(bl(i) - br(i))) {
          double pmp_2 = dq(i - 1);
          double lac_2 = pmp_2 - 0.75 * dq(i - 2);
          br(i) = std::min(max3(0., pmp_2, lac_2), std::max(br(i), min3(0., pmp_2, lac_2)));
          double pmp_1 = -dq(i);
          double lac_1 = pmp_1 + 0.75 * dq(i + 1);
          bl(i) = std::min(max3(0., pmp_1, lac_1), std::max(bl(i), min3(0., pmp_1, lac_1)));
        }
      }
    } else if (iord == 11) {
      for (int i = is1; i <= ie1; i++) {
        double xt = ppm_fac * dm(i);
        bl(i) = -fsign(std::min(std::fabs(xt), std::fabs(al(i) - q1(i))), xt);
        br(i) = fsign(std::min(std::fabs(xt), std::fabs(al(i + 1) - q1(i))), xt);
      }
    } else if (iord == 7 || iord == 12) {
      for (int i = is1; i <= ie1; i++) {
        bl(i) = al(i) - q1(i); br(i) = al(i + 1) - q1(i);
        a4(i) = -3. * (bl(i) + br(i));
        da1(i) = br(i) - bl(i);
        ext5(i) = br(i) * bl(i) > 0.;
        ext6(i) = std::fabs(da1(i)) < -a4(i);
      }
      for (int i = is1; i <= ie1; i++) {
        if (ext6(i)) {
          if (q1(i) + 0.25 / a4(i) * (da1(i) * da1(i)) + a4(i) * r12 < 0.) {
            if (ext5(i)) { br(i) = 0.; bl(i) = 0.; }
            else if (da1(i) > 0.) br(i) = -2. * bl(i);
            else bl(i) = -2. * br(i);
          }
        }
      }
    } else {
      for (int i = is1; i <= ie1; i++) { bl(i) = al(i) - q1(i); br(i) = al(i + 1) - q1(i); }
    }
    if (iord == 9 || iord == 13) pert_ppm(ie1 - is1 + 1, &q1(is1), &bl(is1), &br(is1), 0);

    if (cube) {
      if (is == 1) {
        bl(0) = s14 * dm(-1) + s11 * (q1(-1) - q1(0));
        double xt = 0.5 * (((2. * dxa(0, j) + dxa(-1, j)) * q1(0) - dxa(0, j) * q1(-1)) / (dxa(-1, j) + dxa(0, j)) +
                           ((2. * dxa(1, j) + dxa(2, j)) * q1(1) - dxa(1, j) * q1(2)) / (dxa(1, j) + dxa(2, j)));
        xt = std::max(xt, min4(q1(-1), q1(0), q1(1), q1(2)));
        xt = std::min(xt, max4(q1(-1), q1(0), q1(1), q1(2)));
        br(0) = xt - q1(0);
        bl(1) = xt - q1(1);
        xt = s15 * q1(1) + s11 * q1(2) - s14 * dm(2);
        br(1) = xt - q1(1);
        bl(2) = xt - q1(2);
        br(2) = al(3) - q1(2);
        pert_ppm(3, &q1(0), &bl(0), &br(0), 1);
      }
      if ((ie + 1) == npx) {
        bl(npx - 2) = al(npx - 2) - q1(npx - 2);
        double xt = s15 * q1(npx - 1) + s11 * q1(npx - 2) + s14 * dm(npx - 2);
        br(npx - 2) = xt - q1(npx - 2);
        bl(npx - 1) = xt - q1(npx - 1);
        xt = 0.5 * (((2. * dxa(npx - 1, j) + dxa(npx - 2, j)) * q1(npx - 1) - dxa(npx - 1, j) * q1(npx - 2)) /
                        (dxa(npx - 2, j) + dxa(npx - 1, j)) +
                    ((2. * dxa(npx, j) + dxa(npx + 1, j)) * q1(npx) - dxa(npx, j) * q1(npx + 1)) /
                        (dxa(npx, j) + dxa(npx + 1, j)));
        xt = std::max(xt, min4(q1(npx - 2), q1(npx - 1), q1(npx), q1(npx + 1)));
        xt = std::min(xt, max4(q1(npx - 2), q1(npx - 1), q1(npx), q1(npx + 1)));
        br(npx - 1) = xt - q1(npx - 1);
        bl(npx) = xt - q1(npx);
        br(npx) = s11 * (q1(npx + 1) - q1(npx)) - s14 * dm(npx + 1);
        pert_ppm(3, &q1(npx - 2), &bl(npx - 2), &br(npx - 2), 1);
      }
    }

    if (iord == 7) {
      for (int i = is - 1; i <= ie + 1; i++) { b0(i) = bl(i) + br(i); smt5(i) = bl(i) * br(i) < 0.; }
      for (int i = is; i <= ie + 1; i++) {
        if (c(i, j) > 0.) { fx1(i) = (1. - c(i, j)) * (br(i - 1) - c(i, j) * b0(i - 1)); flux(i, j) = q1(i - 1); }
        else { fx1(i) = (1. + c(i, j)) * (bl(i) + c(i, j) * b0(i)); flux(i, j) = q1(i); }
        if (smt5(i - 1) || smt5(i)) flux(i, j) = flux(i, j) + fx1(i);
      }
    } else {
      for (int i = is; i <= ie + 1; i++)
        flux(i, j) = PPM_FLUX_MONO(q1(i - 1), q1(i), c(i, j), bl(i - 1), br(i - 1), bl(i), br(i));
    }
  }
}

// tp_core.F90:715-1152
void yppm(V2 flux, V2 q, V2 c, int jord, int ifirst, int ilast, int isd, int ied, int js, int je,
          int jsd, int jed, int npx, int npy, V2 dya, bool bounded_domain, int grid_type, double lim_fac) {
  (void)isd; (void)ied; (void)jsd; (void)jed; (void)npx;
  L2 dm(ifirst, ilast, js - 2, je + 2), al(ifirst, ilast, js - 1, je + 2);
  L2 bl(ifirst, ilast, js - 1, je + 1), br(ifirst, ilast, js - 1, je + 1), b0(ifirst, ilast, js - 1, je + 1);
  L2 dq(ifirst, ilast, js - 3, je + 2);
  L1 fx1(ifirst, ilast), xt1(ifirst, ilast), a4(ifirst, ilast);
  LB2 smt5(ifirst, ilast, js - 1, je + 1), smt6(ifirst, ilast, js - 1, je + 1);
  LB1 hi5(ifirst, ilast), hi6(ifirst, ilast);
  int js1, je3, je1;
  const bool cube = (!bounded_domain && grid_type < 3);
  if (cube) {
    js1 = std::max(3, js - 1); je3 = std::min(npy - 2, je + 2); je1 = std::min(npy - 3, je + 1);
  } else {
    js1 = js - 1; je3 = je + 2; je1 = je + 1;
  }
  const int mord = std::abs(jord);

  if (jord < 7) {
    for (int j = js1; j <= je3; j++)
      for (int i = ifirst; i <= ilast; i++)
        al(i, j) = p1 * (q(i, j - 1) + q(i, j)) + p2 * (q(i, j - 2) + q(i, j + 1));
    if (cube) {
      if (js == 1) {
        for (int i = ifirst; i <= ilast; i++) {
          al(i, 0) = c1 * q(i, -2) + c2 * q(i, -1) + c3 * q(i, 0);
          al(i, 1) = 0.5 * (((2. * dya(i, 0) + dya(i, -1)) * q(i, 0) - dya(i, 0) * q(i, -1)) / (dya(i, -1) + dya(i, 0)) +
                            ((2. * dya(i, 1) + dya(i, 2)) * q(i, 1) - dya(i, 1) * q(i, 2)) / (dya(i, 1) + dya(i, 2)));
          al(i, 2) = c3 * q(i, 1) + c2 * q(i, 2) + c1 * q(i, 3);
        }
      }
      if ((je + 1) == npy) {
        for (int i = ifirst; i <= ilast; i++) {
          al(i, npy - 1) = c1 * q(i, npy - 3) + c2 * q(i, npy - 2) + c3 * q(i, npy - 1);
          al(i, npy) = 0.5 * (((2. * dya(i, npy - 1) + dya(i, npy - 2)) * q(i, npy - 1) - dya(i, npy - 1) * q(i, npy - 2)) /
                                  (dya(i, npy - 2) + dya(i, npy - 1)) +
                              ((2. * dya(i, npy) + dya(i, npy + 1)) * q(i, npy) - dya(i, npy) * q(i, npy + 1)) /
                                  (dya(i, npy) + dya(i, npy + 1)));
          al(i, npy + 1) = c3 * q(i, npy) + c2 * q(i, npy + 1) + c1 * q(i, npy + 2);
        }
      }
    }
    if (jord < 0)
      for (int j = js - 1; j <= je + 2; j++)
        for (int i = ifirst; i <= ilast; i++) al(i, j) = std::max(0., al(i, j));

    if (mord == 1) {
      for (int j = js - 1; j <= je + 1; j++)
        for (int i = ifirst; i <= ilast; i++) {
          bl(i, j) = al(i, j) - q(i, j); br(i, j) = al(i, j + 1) - q(i, j); b0(i, j) = bl(i, j) + br(i, j);
          smt5(i, j) = std::fabs(lim_fac * b0(i, j)) < std::fabs(bl(i, j) - br(i, j));
        }
      for (int j = js; j <= je + 1; j++)
        for (int i = ifirst; i <= ilast; i++) {
          if (c(i, j) > 0.) { fx1(i) = (1. - c(i, j)) * (br(i, j - 1) - c(i, j) * b0(i, j - 1)); flux(i, j) = q(i, j - 1); }
          else { fx1(i) = (1. + c(i, j)) * (bl(i, j) + c(i, j) * b0(i, j)); flux(i, j) = q(i, j); }
          if (smt5(i, j - 1) || smt5(i, j)) flux(i, j) = flux(i, j) + fx1(i);
        }
    } else if (mord == 2) {
      for (int j = js; j <= je + 1; j++)
        for (int i = ifirst; i <= ilast; i++) {
          double xt = c(i, j), qtmp;
          if (xt > 0.) {
            qtmp = q(i, j - 1);
            flux(i, j) = qtmp + (1. - xt) * (al(i, j) - qtmp - xt * (al(i, j - 1) + al(i, j) - (qtmp + qtmp)));
          } else {
            qtmp = q(i, j);
            flux(i, j) = qtmp + (1. + xt) * (al(i, j) - qtmp + xt * (al(i, j) + al(i, j + 1) - (qtmp + qtmp)));
          }
        }
    } else if (mord == 3) {
      for (int j = js - 1; j <= je + 1; j++)
        for (int i = ifirst; i <= ilast; i++) {
          bl(i, j) = al(i, j) - q(i, j); br(i, j) = al(i, j + 1) - q(i, j); b0(i, j) = bl(i, j) + br(i, j);
          double x0 = std::fabs(b0(i, j)), xt = std::fabs(bl(i, j) - br(i, j));
          smt5(i, j) = x0 < xt; smt6(i, j) = 3. * x0 < xt;
        }
      for (int j = js; j <= je + 1; j++) {
        for (int i = ifirst; i <= ilast; i++) xt1(i) = c(i, j);
        for (int i = ifirst; i <= ilast; i++) {
          if (xt1(i) > 0.) {
            if (smt5(i, j - 1) || smt6(i, j)) flux(i, j) = q(i, j - 1) + (1. - xt1(i)) * (br(i, j - 1) - xt1(i) * b0(i, j - 1));
            else flux(i, j) = q(i, j - 1);
          } else {
            if (smt6(i, j - 1) || smt5(i, j)) flux(i, j) = q(i, j) + (1. + xt1(i)) * (bl(i, j) + xt1(i) * b0(i, j));
            else flux(i, j) = q(i, j);
          }
        }
      }
    } else if (mord == 4) {
      for (int j = js - 1; j <= je + 1; j++)
        for (int i = ifirst; i <= ilast; i++) {
          bl(i, j) = al(i, j) - q(i, j); br(i, j) = al(i, j + 1) - q(i, j); b0(i, j) = bl(i, j) + br(i, j);
          double x0 = std::fabs(b0(i, j)), xt = std::fabs(bl(i, j) - br(i, j));
          smt5(i, j) = x0 < xt; smt6(i, j) = 3. * x0 < xt;
        }
      for (int j = js; j <= je + 1; j++) {
        for (int i = ifirst; i <= ilast; i++) {
          xt1(i) = c(i, j);
          hi5(i) = smt5(i, j - 1) && smt5(i, j);
          hi6(i) = smt6(i, j - 1) || smt6(i, j);
          hi5(i) = hi5(i) || hi6(i);
        }
        for (int i = ifirst; i <= ilast; i++) {
          if (xt1(i) > 0.) { fx1(i) = (1. - xt1(i)) * (br(i, j - 1) - xt1(i) * b0(i, j - 1)); flux(i, j) = q(i, j - 1); }
          else { fx1(i) = (1. + xt1(i)) * (bl(i, j) + xt1(i) * b0(i, j)); flux(i, j) = q(i, j); }
          if (hi5(i)) flux(i, j) = flux(i, j) + fx1(i);
        }
      }
    } else {
      if (jord == 5) {
        for (int j = js - 1; j <= je + 1; j++)
          for (int i = ifirst; i <= ilast; i++) {
            bl(i, j) = al(i, j) - q(i, j); br(i, j) = al(i, j + 1) - q(i, j); b0(i, j) = bl(i, j) + br(i, j);
            smt5(i, j) = bl(i, j) * br(i, j) < 0.;
          }
      } else {
        if (jord == -5) {
          for (int j = js - 1; j <= je + 1; j++) {
            for (int i = ifirst; i <= ilast; i++) {
              bl(i, j) = al(i, j) - q(i, j); br(i, j) = al(i, j + 1) - q(i, j); b0(i, j) = bl(i, j) + br(i, j);
              xt1(i) = br(i, j) - bl(i, j);
              a4(i) = -3. * b0(i, j);
              smt5(i, j) = bl(i, j) * br(i, j) < 0.;
            }
            for (int i = ifirst; i <= ilast; i++) {
              if (std::fabs(xt1(i)) < -a4(i)) {
                if (q(i, j) + 0.25 / a4(i) * (xt1(i) * xt1(i)) + a4(i) * r12 < 0.) {
                  if (!smt5(i, j)) { br(i, j) = 0.; bl(i, j) = 0.; b0(i, j) = 0.; }
                  else if (xt1(i) > 0.) { br(i, j) = -2. * bl(i, j); b0(i, j) = -bl(i, j); }
                  else { bl(i, j) = -2. * br(i, j); b0(i, j) = -br(i, j); }
                }
              }
            }
          }
        } else {
          for (int j = js - 1; j <= je + 1; j++)
            for (int i = ifirst; i <= ilast; i++) {
              bl(i, j) = al(i, j) - q(i, j); br(i, j) = al(i, j + 1) - q(i, j); b0(i, j) = bl(i, j) + br(i, j);
              smt5(i, j) = 3. * std::fabs(b0(i, j)) < std::fabs(bl(i, j) - br(i, j));
            }
        }
        if (cube) {
          if (js == 1)
            for (int i = ifirst; i <= ilast; i++) {
              smt5(i, 0) = bl(i, 0) * br(i, 0) < 0.;
              smt5(i, 1) = bl(i, 1) * br(i, 1) < 0.;
            }
          if ((je + 1) == npy)
            for (int i = ifirst; i <= ilast; i++) {
              smt5(i, npy - 1) = bl(i, npy - 1) * br(i, npy - 1) < 0.;
              smt5(i, npy) = bl(i, npy) * br(i, npy) < 0.;
            }
        }
      }
      for (int j = js; j <= je + 1; j++)
        for (int i = ifirst; i <= ilast; i++) {
          if (c(i, j) > 0.) { fx1(i) = (1. - c(i, j)) * (br(i, j - 1) - c(i, j) * b0(i, j - 1)); flux(i, j) = q(i, j - 1); }
          else { fx1(i) = (1. + c(i, j)) * (bl(i, j) + c(i, j) * b0(i, j)); flux(i, j) = q(i, j); }
          if (smt5(i, j - 1) || smt5(i, j)) flux(i, j) = flux(i, j) + fx1(i);
        }
    }
    return;
  }

  // Monotonic constraints, tp_core.F90:977-1150
  for (int j = js - 2; j <= je + 2; j++)
    for (int i = ifirst; i <= ilast; i++) {
      double xt = 0.25 * (q(i, j + 1) - q(i, j - 1));
      dm(i, j) = fsign(std::min(std::min(std::fabs(xt), max3(q(i, j - 1), q(i, j), q(i, j + 1)) - q(i, j)),
                                q(i, j) - min3(q(i, j - 1), q(i, j), q(i, j + 1))), xt);
    }
  for (int j = js1; j <= je1 + 1; j++)
    for (int i = ifirst; i <= ilast; i++)
      al(i, j) = 0.5 * (q(i, j - 1) + q(i, j)) + r3 * (dm(i, j - 1) - dm(i, j));

  if (jord == 8) {
    for (int j = js1; j <= je1; j++)
      for (int i = ifirst; i <= ilast; i++) {
        double xt = 2. * dm(i, j);
        bl(i, j) = -fsign(std::min(std::fabs(xt), std::fabs(al(i, j) - q(i, j))), xt);
        br(i, j) = fsign(std::min(std::fabs(xt), std::fabs(al(i, j + 1) - q(i, j))), xt);
      }
  } else if (jord == 10) {
    for (int j = js1 - 2; j <= je1 + 1; j++)
      for (int i = ifirst; i <= ilast; i++) dq(i, j) = 2. * (q(i, j + 1) - q(i, j));
    for (int j = js1; j <= je1; j++)
      for (int i = ifirst; i <= ilast; i++) {
        bl(i, j) = al(i, j) - q(i, j);
        br(i, j) = al(i, j + 1) - q(i, j);
        if (std::fabs(dm(i, j - 1)) + std::fabs(dm(i, j)) + std::fabs(dm(i, j + 1)) < near_zero) {
          bl(i, j) = 0.; br(i, j) = 0.;
        } else if (std::fabs(3. * (bl(i, j) + br(i, j))) > std::fabs(bl(i, j) - br(i, j))) {
          double pmp_2 = dq(i, j - 1);
          double lac_2 = pmp_2 - 0.75 * dq(i, j - 2);
          br(i, j) = std::min(max3(0., pmp_2, lac_2), std::max(br(i, j), min3(0., pmp_2, lac_2)));
          double pmp_1 = -dq(i, j);
          double lac_1 = pmp_1 + 0.75 * dq(i, j + 1);
          bl(i, j) = std::min(max3(0., pmp_1, lac_1), std::max(bl(i, j), min3(0., pmp_1, lac_1)));
        }
      }
  } else if (jord == 11) {
    for (int j = js1; j <= je1; j++)
      for (int i = ifirst; i <= ilast; i++) {
        double xt = ppm_fac * dm(i, j);
        bl(i, j) = -fsign(std::min(std::fabs(xt), std::fabs(al(i, j) - q(i, j))), xt);
        br(i, j) = fsign(std::min(std::fabs(xt), std::fabs(al(i, j + 1) - q(i, j))), xt);
      }
  } else if (jord == 7 || jord == 12) {
    for (int j = js1; j <= je1; j++) {
      for (int i = ifirst; i <= ilast; i++) {
        bl(i, j) = al(i, j) - q(i, j); br(i, j) = al(i, j + 1) - q(i, j);
        xt1(i) = br(i, j) - bl(i, j);
        a4(i) = -3. * (br(i, j) + bl(i, j));
        hi5(i) = bl(i, j) * br(i, j) > 0.;
        hi6(i) = std::fabs(xt1(i)) < -a4(i);
      }
      for (int i = ifirst; i <= ilast; i++) {
        if (hi6(i)) {
          if (q(i, j) + 0.25 / a4(i) * (xt1(i) * xt1(i)) + a4(i) * r12 < 0.) {
            if (hi5(i)) { br(i, j) = 0.; bl(i, j) = 0.; }
            else if (xt1(i) > 0.) br(i, j) = -2. * bl(i, j);
            else bl(i, j) = -2. * br(i, j);
          }
        }
      }
    }
  } else {
    for (int j = js1; j <= je1; j++)
      for (int i = ifirst; i <= ilast; i++) { bl(i, j) = al(i, j) - q(i, j); br(i, j) = al(i, j + 1) - q(i, j); }
  }
  if (jord == 9 || jord == 13)
    for (int j = js1; j <= je1; j++)
      for (int i = ifirst; i <= ilast; i++) pert_ppm(1, &q(i, j), &bl(i, j), &br(i, j), 0);

  if (cube) {
    if (js == 1) {
      for (int i = ifirst; i <= ilast; i++) {
        bl(i, 0) = s14 * dm(i, -1) + s11 * (q(i, -1) - q(i, 0));
        double xt = 0.5 * (((2. * dya(i, 0) + dya(i, -1)) * q(i, 0) - dya(i, 0) * q(i, -1)) / (dya(i, -1) + dya(i, 0)) +
                           ((2. * dya(i, 1) + dya(i, 2)) * q(i, 1) - dya(i, 1) * q(i, 2)) / (dya(i, 1) + dya(i, 2)));
        xt = std::max(xt, min4(q(i, -1), q(i, 0), q(i, 1), q(i, 2)));
        xt = std::min(xt, max4(q(i, -1), q(i, 0), q(i, 1), q(i, 2)));
        br(i, 0) = xt - q(i, 0);
        bl(i, 1) = xt - q(i, 1);
        xt = s15 * q(i, 1) + s11 * q(i, 2) - s14 * dm(i, 2);
        br(i, 1) = xt - q(i, 1);
        bl(i, 2) = xt - q(i, 2);
        br(i, 2) = al(i, 3) - q(i, 2);
      }
      // tp_core.F90:1094 -- sequence-associated 3 rows; pointwise so per element
      for (int j = 0; j <= 2; j++)
        for (int i = ifirst; i <= ilast; i++) pert_ppm(1, &q(i, j), &bl(i, j), &br(i, j), 1);
    }
    if ((je + 1) == npy) {
      for (int i = ifirst; i <= ilast; i++) {
        bl(i, npy - 2) = al(i, npy - 2) - q(i, npy - 2);
        double xt = s15 * q(i, npy - 1) + s11 * q(i, npy - 2) + s14 * dm(i, npy - 2);
        br(i, npy - 2) = xt - q(i, npy - 2);
        bl(i, npy - 1) = xt - q(i, npy - 1);
        xt = 0.5 * (((2. * dya(i, npy - 1) + dya(i, npy - 2)) * q(i, npy - 1) - dya(i, npy - 1) * q(i, npy - 2)) /
                        (dya(i, npy - 2) + dya(i, npy - 1)) +
                    ((2. * dya(i, npy) + dya(i, npy + 1)) * q(i, npy) - dya(i, npy) * q(i, npy + 1)) /
                        (dya(i, npy) + dya(i, npy + 1)));
        xt = std::max(xt, min4(q(i, npy - 2), q(i, npy - 1), q(i, npy), q(i, npy + 1)));
        xt = std::min(xt, max4(q(i, npy - 2), q(i, npy - 1), q(i, npy), q(i, npy + 1)));
        br(i, npy - 1) = xt - q(i, npy - 1);
        bl(i, npy) = xt - q(i, npy);
        br(i, npy) = s11 * (q(i, npy + 1) - q(i, npy)) - s14 * dm(i, npy + 1);
      }
      for (int j = npy - 2; j <= npy; j++)
        for (int i = ifirst; i <= ilast; i++) pert_ppm(1, &q(i, j), &bl(i, j), &br(i, j), 1);
    }
  }

  if (jord == 7) {
    for (int j = js - 1; j <= je + 1; j++)
      for (int i = ifirst; i <= ilast; i++) { b0(i, j) = bl(i, j) + br(i, j); smt5(i, j) = bl(i, j) * br(i, j) < 0.; }
    for (int j = js; j <= je + 1; j++)
      for (int i = ifirst; i <= ilast; i++) {
        if (c(i, j) > 0.) { fx1(i) = (1. - c(i, j)) * (br(i, j - 1) - c(i, j) * b0(i, j - 1)); flux(i, j) = q(i, j - 1); }
        else { fx1(i) = (1. + c(i, j)) * (bl(i, j) + c(i, j) * b0(i, j)); flux(i, j) = q(i, j); }
        if (smt5(i, j - 1) || smt5(i, j)) flux(i, j) = flux(i, j) + fx1(i);
      }
  } else {
    for (int j = js; j <= je + 1; j++)
      for (int i = ifirst; i <= ilast; i++)
        flux(i, j) = PPM_FLUX_MONO(q(i, j - 1), q(i, j), c(i, j), bl(i, j - 1), br(i, j - 1), bl(i, j), br(i, j));
  }
}

// tp_core.F90:1267-1447 (non-USE_SG branch; damp_Km not supported: SHiELD 2-D Smagorinsky option)
void deln_flux(int nord, int is, int ie, int js, int je, int npx, int npy, double damp, V2 q, V2 fx, V2 fy,
               const Grid& g, const Bd& bd, const V2* mass) {
  L2 fx2(bd.isd, bd.ied + 1, bd.jsd, bd.jed), fy2(bd.isd, bd.ied, bd.jsd, bd.jed + 1);
  L2 d2(bd.isd, bd.ied, bd.jsd, bd.jed);
  const int i1 = is - 1 - nord, i2 = ie + 1 + nord, j1 = js - 1 - nord, j2 = je + 1 + nord;
  if (!mass) {
    for (int j = j1; j <= j2; j++) for (int i = i1; i <= i2; i++) d2(i, j) = damp * q(i, j);
  } else {
    for (int j = j1; j <= j2; j++) for (int i = i1; i <= i2; i++) d2(i, j) = q(i, j);
  }
  if (nord > 0) copy_corners(d2, npx, npy, 1, bd);
  for (int j = js - nord; j <= je + nord; j++)
    for (int i = is - nord; i <= ie + nord + 1; i++) fx2(i, j) = g.del6_v(i, j) * (d2(i - 1, j) - d2(i, j));
  if (nord > 0) copy_corners(d2, npx, npy, 2, bd);
  for (int j = js - nord; j <= je + nord + 1; j++)
    for (int i = is - nord; i <= ie + nord; i++) fy2(i, j) = g.del6_u(i, j) * (d2(i, j - 1) - d2(i, j));
  if (nord > 0) {
    for (int n = 1; n <= nord; n++) {
      const int nt = nord - n;
      for (int j = js - nt - 1; j <= je + nt + 1; j++)
        for (int i = is - nt - 1; i <= ie + nt + 1; i++)
          d2(i, j) = (fx2(i, j) - fx2(i + 1, j) + fy2(i, j) - fy2(i, j + 1)) * g.rarea(i, j);
      copy_corners(d2, npx, npy, 1, bd);
      for (int j = js - nt; j <= je + nt; j++)
        for (int i = is - nt; i <= ie + nt + 1; i++) fx2(i, j) = g.del6_v(i, j) * (d2(i, j) - d2(i - 1, j));
      copy_corners(d2, npx, npy, 2, bd);
      for (int j = js - nt; j <= je + nt + 1; j++)
        for (int i = is - nt; i <= ie + nt; i++) fy2(i, j) = g.del6_u(i, j) * (d2(i, j) - d2(i, j - 1));
    }
  }
  if (mass) {
    const V2& m = *mass;
    const double damp2 = 0.5 * damp;
    for (int j = js; j <= je; j++)
      for (int i = is; i <= ie + 1; i++) fx(i, j) = fx(i, j) + damp2 * (m(i - 1, j) + m(i, j)) * fx2(i, j);
    for (int j = js; j <= je + 1; j++)
      for (int i = is; i <= ie; i++) fy(i, j) = fy(i, j) + damp2 * (m(i, j - 1) + m(i, j)) * fy2(i, j);
  } else {
    for (int j = js; j <= je; j++) for (int i = is; i <= ie + 1; i++) fx(i, j) = fx(i, j) + fx2(i, j);
    for (int j = js; j <= je + 1; j++) for (int i = is; i <= ie; i++) fy(i, j) = fy(i, j) + fy2(i, j);
  }
}

// tp_core.F90:85-241
void fv_tp_2d(V2 q, V2 crx, V2 cry, int npx, int npy, int hord, V2 fx, V2 fy, V2 xfx, V2 yfx,
              const Grid& g, const Bd& bd, V2 ra_x, V2 ra_y, double lim_fac, const V2* mfx, const V2* mfy,
              const V2* mass, bool use_damp, int nord, double damp_c) {
  const int is = bd.is, ie = bd.ie, js = bd.js, je = bd.je, isd = bd.isd, ied = bd.ied, jsd = bd.jsd, jed = bd.jed;
  L2 q_i(isd, ied, js, je), q_j(is, ie, jsd, jed);
  L2 fx2(is, ie + 1, jsd, jed), fy2(isd, ied, js, je + 1), fyy(isd, ied, js, je + 1);
  L1 fx1(is, ie + 1);
  const int ord_in = (hord == 10) ? 8 : hord;
  const int ord_ou = hord;

  if (!bd.bounded_domain) copy_corners(q, npx, npy, 2, bd);
  yppm(fy2, q, cry, ord_in, isd, ied, isd, ied, js, je, jsd, jed, npx, npy, g.dya, bd.bounded_domain, bd.grid_type, lim_fac);
  for (int j = js; j <= je + 1; j++)
    for (int i = isd; i <= ied; i++) fyy(i, j) = yfx(i, j) * fy2(i, j);
  for (int j = js; j <= je; j++)
    for (int i = isd; i <= ied; i++) q_i(i, j) = (q(i, j) * g.area(i, j) + fyy(i, j) - fyy(i, j + 1)) / ra_y(i, j);
  xppm(fx, q_i, crx, ord_ou, is, ie, isd, ied, js, je, jsd, jed, npx, npy, g.dxa, bd.bounded_domain, bd.grid_type, lim_fac);

  if (!bd.bounded_domain) copy_corners(q, npx, npy, 1, bd);
  xppm(fx2, q, crx, ord_in, is, ie, isd, ied, jsd, jed, jsd, jed, npx, npy, g.dxa, bd.bounded_domain, bd.grid_type, lim_fac);
  for (int j = jsd; j <= jed; j++) {
    for (int i = is; i <= ie + 1; i++) fx1(i) = xfx(i, j) * fx2(i, j);
    for (int i = is; i <= ie; i++) q_j(i, j) = (q(i, j) * g.area(i, j) + fx1(i) - fx1(i + 1)) / ra_x(i, j);
  }
  yppm(fy, q_j, cry, ord_ou, is, ie, isd, ied, js, je, jsd, jed, npx, npy, g.dya, bd.bounded_domain, bd.grid_type, lim_fac);

  if (mfx && mfy) {
    for (int j = js; j <= je; j++)
      for (int i = is; i <= ie + 1; i++) fx(i, j) = 0.5 * (fx(i, j) + fx2(i, j)) * (*mfx)(i, j);
    for (int j = js; j <= je + 1; j++)
      for (int i = is; i <= ie; i++) fy(i, j) = 0.5 * (fy(i, j) + fy2(i, j)) * (*mfy)(i, j);
    if (use_damp && mass) {
      if (damp_c > 1.e-4) {
        double damp = std::pow(damp_c * g.da_min, (double)(nord + 1));
        deln_flux(nord, is, ie, js, je, npx, npy, damp, q, fx, fy, g, bd, mass);
      }
    }
  } else {
    for (int j = js; j <= je; j++)
      for (int i = is; i <= ie + 1; i++) fx(i, j) = 0.5 * (fx(i, j) + fx2(i, j)) * xfx(i, j);
    for (int j = js; j <= je + 1; j++)
      for (int i = is; i <= ie; i++) fy(i, j) = 0.5 * (fy(i, j) + fy2(i, j)) * yfx(i, j);
    if (use_damp) {
      if (damp_c > 1.E-4) {
        double damp = std::pow(damp_c * g.da_min, (double)(nord + 1));
        deln_flux(nord, is, ie, js, je, npx, npy, damp, q, fx, fy, g, bd, nullptr);
      }
    }
  }
}

}  // namespace fv3o

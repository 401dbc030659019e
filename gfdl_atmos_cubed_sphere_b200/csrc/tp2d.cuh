// fv_tp_2d / deln_flux / del6_vt_flux device pipeline (internal header).
// Reference: model/tp_core.F90:85-241 (fv_tp_2d), :1267-1447 (deln_flux),
// model/sw_core.F90:1608-1737 (del6_vt_flux).
#pragma once
#include "fv3_ctx.hpp"

// per-level parameter tables living in c->d_kint / c->d_kdbl (FV3_KSLOTS slots x (npz+1))
enum { KI_NORD = 0, KI_NORD_V = 1, KI_NORD_W = 2, KI_NORD_T = 3 };
enum { KD_D2BG = 0, KD_DAMP_V = 1, KD_DAMP_W = 2, KD_DAMP_T = 3, KD_DCON = 4, KD_DAMP4_W = 5, KD_DAMP4_V = 6, KD_DELN = 7, KD_DELN_T = 8, KD_DZ = 9, KD_DD8 = 10 };
#define FV3_KSLOTS 12

struct Tp2d {
  const double* q;            // (isd:ied, jsd:jed, nk)  transported scalar
  const double *crx, *cry, *xfx, *yfx;
  const double *ra_x, *ra_y;  // nullable: then area + xfx(i)-xfx(i+1) is formed on the fly
  double *fx, *fy;            // out
  const double *mfx, *mfy;    // nullable: weight by xfx,yfx instead (tp_core.F90:213-226)
  int hord;
  int nk;
  // optional: instead of storing fx, fy apply update_dz_d's height update (nh_utils.F90:282-299) in the epilogue:
  // zn = (q*area + div(fx, fy)) / (ra_x + ra_y - area) + del-n flux divergence (zn_dfx, zn_dfy, coefficient slot zn_slot)
  double* zn = nullptr; const double *zn_dfx = nullptr, *zn_dfy = nullptr; int zn_slot = 0;
};

// Lin-Rood 2-D transport fluxes for nk levels (all arrays indexed from their level 0).
int launch_tp2d(fv3_ctx* c, const Tp2d& a);

// del-n diffusive fluxes of q (nord(k) from d_kint[slot_nord], coefficient from d_kdbl[slot_damp];
// levels whose coefficient <= thresh are skipped).  premul: d2 = damp*q (else d2 = q).
// Results in fx2 (is:ie+1, js:je), fy2 (is:ie, js:je+1).  d2, t1, t2: scratch planes.
struct Deln {
  const double* q; double *fx2, *fy2, *d2;
  int slot_nord, slot_damp; double thresh; int premul; int nk;
  int nord_const; double damp_const;  // used when slot_nord < 0
  // levels [k_lo, k_hi] are launched (the others have a zero coefficient: in flag-set A only the two sponge levels
  // damp w, and launching all npz levels cost 0.16 ms of blocks that exit at once); nord_max = largest order among them
  int k_lo = 0, k_hi = -1, nord_max = -1;
};
int launch_deln(fv3_ctx* c, const Deln& a);

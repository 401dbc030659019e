// Host-side execution of the kernels' own flux arithmetic (ppm.cuh / tp_tile.cuh, compiled __host__ __device__).
// Driven by tests/test_host_device_math.py, which compares the results with the oracle.  No GPU involved: only host code runs.
//   stdin : int mode, n, iord, p0, p1, then mode-specific doubles (binary);  stdout: double flux[n+1] (binary)
//   mode 0: periodic line through tpt::line_flux_na (interior fast path of the tile kernels)
//           p0 = stride, p1 = rare;            doubles: q[n], c[n+1]
//   mode 1: one full cube-face line through ppm::flux_scalar (the cube-edge operator of the frame tiles, edge_flux)
//           p0 = stride, p1 = rare;            doubles: q[n+6] (index -2..n+3), c[n+1] (faces 1..n+1), dxa[n+6]
//   mode 2: one full cube-face line of the wind operator ppm::flux_wind (xtp_u / ytp_v of k_dsw_ke)
//           p0 = zero (edge line), p1 unused;  doubles: u[n+6], c[n+1], dx[n+6], rdx[n+6]
//   mode 3: ppm::flux_wind_fast_g (interior fast path of k_dsw_ke<true>) on the faces 4..n-2 of the same line; same input as mode 2
#include <cstdio>
#include <vector>
#include "../gfdl_atmos_cubed_sphere_b200/csrc/tp_tile.cuh"

static bool rd(std::vector<double>& v) { return fread(v.data(), sizeof(double), v.size(), stdin) == v.size(); }

int main() {
  int hdr[5];
  if (fread(hdr, sizeof(int), 5, stdin) != 5) return 2;
  const int mode = hdr[0], n = hdr[1], iord = hdr[2], p0 = hdr[3], p1 = hdr[4];
  std::vector<double> flux(n + 1);
  if (mode == 0) {
    const int stride = p0, rare = p1;
    std::vector<double> q(n), c(n + 1);
    if (!rd(q) || !rd(c)) return 2;
    // periodic line with 3 halo cells on each side, stored with the requested stride (stride > 1 = a y line of a tile array)
    std::vector<double> qh((size_t)(n + 6) * stride, -1.0e300);
    for (int i = -3; i < n + 3; i++) qh[(size_t)(i + 3) * stride] = q[((i % n) + n) % n];
    const bool mono = iord >= 7;   // dm family (tp_core.F90:364, 563), as tp_compute decides for the general instantiation
    for (int i = 0; i <= n; i++) {
      const double* p = &qh[(size_t)(i + 3) * stride];
      flux[i] = rare ? tpt::line_flux_na<true>(mono, p, stride, c[i], iord) : tpt::line_flux_na<false>(mono, p, stride, c[i], iord);
    }
  } else if (mode == 1) {
    const int stride = p0, rare = p1;
    std::vector<double> q(n + 6), c(n + 1), dxa(n + 6);
    if (!rd(q) || !rd(c) || !rd(dxa)) return 2;
    std::vector<double> qs((size_t)(n + 6) * stride, -1.0e300);
    for (int m = 0; m < n + 6; m++) qs[(size_t)m * stride] = q[m];
    const tpt::SAcc qa{qs.data(), stride, -2};                 // sweep index -2 is element 0
    const ppm::Acc da{dxa.data(), 2, 1};                        // dxa(s) = dxa[2 + s]
    for (int i = 1; i <= n + 1; i++)
      flux[i - 1] = rare ? ppm::flux_scalar<true>(qa, da, i, c[i - 1], iord, n + 1, true)
                         : ppm::flux_scalar<false>(qa, da, i, c[i - 1], iord, n + 1, true);
  } else if (mode == 2) {
    const bool zero = p0 != 0;
    std::vector<double> u(n + 6), c(n + 1), dx(n + 6), rdx(n + 6);
    if (!rd(u) || !rd(c) || !rd(dx) || !rd(rdx)) return 2;
    const ppm::Acc ua{u.data(), 2, 1}, da{dx.data(), 2, 1}, ra{rdx.data(), 2, 1};
    for (int i = 1; i <= n + 1; i++) flux[i - 1] = ppm::flux_wind(ua, da, ra, i, c[i - 1], iord, n + 1, true, zero);
  } else if (mode == 3) {   // interior fast path of k_dsw_ke<true>: faces 4..n-2 only (the others are left 0)
    std::vector<double> u(n + 6), c(n + 1), dx(n + 6), rdx(n + 6);
    if (!rd(u) || !rd(c) || !rd(dx) || !rd(rdx)) return 2;
    for (int i = 4; i <= n - 2; i++)
      flux[i - 1] = ppm::flux_wind_fast_g(&u[i + 2], 1, c[i - 1], rdx[i - 1 + 2], rdx[i + 2], iord);
  } else return 3;
  fwrite(flux.data(), sizeof(double), n + 1, stdout);
  return 0;
}

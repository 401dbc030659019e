// Host-side execution of the kernels' own flux arithmetic (ppm.cuh / tp_tile.cuh line_flux_na, compiled __host__ __device__):
// 1-D periodic line, every scheme of tp_valid_schemes.  Driven by tests/test_host_device_math.py, which compares the result with
// the oracle's xppm / yppm (fv3o_ppm_periodic).  No GPU involved: only host code runs.
//   stdin : int n, int iord, int stride, int rare, double q[n], double c[n+1]      (binary)
//   stdout: double flux[n+1]                                                        (binary)
#include <cstdio>
#include <vector>
#include "../gfdl_atmos_cubed_sphere_b200/csrc/tp_tile.cuh"

int main() {
  int hdr[4];
  if (fread(hdr, sizeof(int), 4, stdin) != 4) return 2;
  const int n = hdr[0], iord = hdr[1], stride = hdr[2], rare = hdr[3];
  std::vector<double> q(n), c(n + 1), flux(n + 1);
  if (fread(q.data(), sizeof(double), n, stdin) != (size_t)n) return 2;
  if (fread(c.data(), sizeof(double), n + 1, stdin) != (size_t)(n + 1)) return 2;
  // periodic line with 3 halo cells on each side, stored with the requested stride (stride > 1 = a y line of a tile array)
  std::vector<double> qh((size_t)(n + 6) * stride, -1.0e300);
  for (int i = -3; i < n + 3; i++) qh[(size_t)(i + 3) * stride] = q[((i % n) + n) % n];
  const bool mono = iord >= 7;   // dm family (tp_core.F90:364, 563), as tp_compute decides for the general instantiation
  for (int i = 0; i <= n; i++) {
    const double* p = &qh[(size_t)(i + 3) * stride];
    flux[i] = rare ? tpt::line_flux_na<true>(mono, p, stride, c[i], iord) : tpt::line_flux_na<false>(mono, p, stride, c[i], iord);
  }
  fwrite(flux.data(), sizeof(double), n + 1, stdout);
  return 0;
}

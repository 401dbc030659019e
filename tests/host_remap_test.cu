// Host-side execution of the vertical-remap column operators the CUDA kernels run (csrc/remap_col.cuh, compiled __host__ __device__).
// Driven by tests/test_host_remap.py, which compares the result with the oracle (oracle/remap.cpp).  No GPU involved.
//   stdin : int km, ncol, iv, kord, scalar, mapn; double qmin; then doubles pe1[km+1][ncol], pe2[km+1][ncol], q[km][ncol], qs[ncol]
//           (level-major like the device planes: level stride = ncol);  stdout: double q[km][ncol] remapped in place
//   iv = 99: fillz instead (fillz_column), with the layer thicknesses in the first km levels of the pe2 block
#include <cstdio>
#include <vector>
#include "../gfdl_atmos_cubed_sphere_b200/csrc/remap_col.cuh"

static bool rd(std::vector<double>& v) { return fread(v.data(), sizeof(double), v.size(), stdin) == v.size(); }

int main() {
  int h[6]; double qmin;
  if (fread(h, sizeof(int), 6, stdin) != 6 || fread(&qmin, sizeof(double), 1, stdin) != 1) return 2;
  const int km = h[0], ncol = h[1], iv = h[2], kord = h[3], scalar = h[4], mapn = h[5];
  std::vector<double> pe1((size_t)(km + 1) * ncol), pe2((size_t)(km + 1) * ncol), q((size_t)km * ncol), qs(ncol);
  if (!rd(pe1) || !rd(pe2) || !rd(q) || !rd(qs)) return 2;
  const size_t n = (size_t)(km + 1) * ncol;
  std::vector<double> a1(n), a2(n), a3(n), a4(n), qi(n), gam(n);
  std::vector<int> fl(n);
  for (int c = 0; c < ncol; c++) {
    const rmp::Col C{a1.data() + c, a2.data() + c, a3.data() + c, a4.data() + c, qi.data() + c, gam.data() + c, fl.data() + c, (long long)ncol};
    const double *p1 = pe1.data() + c, *p2 = pe2.data() + c;
    auto P1 = [&](int k) { return p1[(size_t)(k - 1) * ncol]; };
    auto P2 = [&](int k) { return p2[(size_t)(k - 1) * ncol]; };
    if (iv == 99) { rmp::fillz_column(km, q.data() + c, p2, (long long)ncol); continue; }
    rmp::remap_field<true>(C, km, P1, P2, q.data() + c, qs[c], iv, kord, qmin, scalar != 0, mapn != 0);
  }
  fwrite(q.data(), sizeof(double), q.size(), stdout);
  return 0;
}

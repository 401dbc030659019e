// Host-side check of the compact launch enumerations (fv3_ctx.hpp FrameGrid / FramePts, tp_tile.cuh tile_maps):
// every tile / point of the frame is produced exactly once, nothing of the interior, nothing outside.
// Built and run by tests/test_launch_enumerations.py with nvcc (no GPU needed: only host code runs).
#include <cstdio>
#include <set>
#include <utility>
#include "../gfdl_atmos_cubed_sphere_b200/csrc/tp_tile.cuh"

static Lay make_lay(int n, int cube) {
  Lay L{};
  L.npx = L.npy = n + 1; L.npz = 3; L.ng = 3;
  L.is = L.js = 1; L.ie = L.je = n; L.isd = L.jsd = -2; L.ied = L.jed = n + 3;
  L.NI = ((5 + n + 7) + 7) / 8 * 8; L.NJ = n + 7; L.plane = ((long long)L.NI * L.NJ + 15) / 16 * 16;
  L.cube = cube; L.grid_type = cube ? 0 : 4;
  return L;
}

static int check_frame_grid(const Lay& L, int ilo, int ihi, int jlo, int jhi) {
  const FrameGrid f = frame_grid(L, ilo, ihi, jlo, jhi);
  std::set<std::pair<int, int>> seen;
  for (int t = 0; t < f.count(); t++) {
    int bx, by; f.map(t, bx, by);
    if (bx < 0 || bx >= f.nbx || by < 0 || by >= f.nby) { printf("frame_grid: tile out of range\n"); return 1; }
    if (!seen.insert({bx, by}).second) { printf("frame_grid: duplicate tile\n"); return 1; }
  }
  // every point outside the interior box must lie in an enumerated tile
  for (int j = L.jsd; j <= L.jed + 1; j++)
    for (int i = L.isd; i <= L.ied + 1; i++) {
      const bool inside = i >= ilo && i <= ihi && j >= jlo && j <= jhi;
      const int bx = (i - (L.isd - FV3_IOFF)) / 32, by = (j - L.jsd) / 8;
      if (!inside && !seen.count({bx, by})) { printf("frame_grid: point (%d,%d) not covered\n", i, j); return 1; }
    }
  return 0;
}

static int check_frame_pts(int o0, int o1, int i0, int i1) {
  const FramePts f = frame_pts(o0, o1, o0, o1, i0, i1, i0, i1);
  std::set<std::pair<int, int>> seen;
  for (int t = 0; t < f.count() + 5; t++) {
    int i, j;
    const bool ok = f.map(t, i, j);
    if (ok != (t < f.count())) { printf("frame_pts: count mismatch\n"); return 1; }
    if (!ok) continue;
    const bool inside = i1 >= i0 && i >= i0 && i <= i1 && j >= i0 && j <= i1;
    if (i < o0 || i > o1 || j < o0 || j > o1 || inside) { printf("frame_pts: bad point (%d,%d)\n", i, j); return 1; }
    if (!seen.insert({i, j}).second) { printf("frame_pts: duplicate point\n"); return 1; }
  }
  const long long all = (long long)(o1 - o0 + 1) * (o1 - o0 + 1), in = i1 >= i0 ? (long long)(i1 - i0 + 1) * (i1 - i0 + 1) : 0;
  if ((long long)seen.size() != all - in) { printf("frame_pts: %zu points, expected %lld\n", seen.size(), all - in); return 1; }
  return 0;
}

static int check_tile_maps(const Lay& L) {
  tpt::TileMap in, fr; int n_in, n_fr;
  tpt::tile_maps(L, in, fr, n_in, n_fr);
  const FrameGrid& f = fr.fg;
  std::set<std::pair<int, int>> seen;
  const int nin = f.nbx - f.cl - f.cr;
  for (int t = 0; t < n_in; t++) {
    const int by = f.a + t / nin, bx = f.cl + t % nin;
    const int i0 = L.is + bx * tpt::TX, j0 = L.js + by * tpt::TY;
    if (L.cube && (i0 < 4 || i0 + tpt::TX > L.npx - 3 || j0 < 4 || j0 + tpt::TY > L.npy - 3)) { printf("tile_maps: edge tile in the interior set\n"); return 1; }
    if (i0 + tpt::TX - 1 > L.ie || j0 + tpt::TY - 1 > L.je) { printf("tile_maps: overhanging tile in the interior set\n"); return 1; }
    if (!seen.insert({bx, by}).second) { printf("tile_maps: duplicate interior tile\n"); return 1; }
  }
  for (int t = 0; t < n_fr; t++) {
    int bx, by; f.map(t, bx, by);
    if (!seen.insert({bx, by}).second) { printf("tile_maps: duplicate / overlapping frame tile\n"); return 1; }
  }
  if ((int)seen.size() != f.nbx * f.nby) { printf("tile_maps: %zu of %d tiles\n", seen.size(), f.nbx * f.nby); return 1; }
  return 0;
}

int main() {
  int bad = 0;
  for (int n : {8, 16, 24, 48, 56, 96, 192, 384}) {
    for (int cube = 0; cube < 2; cube++) {
      const Lay L = make_lay(n, cube);
      bad |= check_tile_maps(L);
      bad |= check_frame_grid(L, 6, L.npx - 5, 6, L.npy - 5);
      bad |= check_frame_grid(L, L.is, L.ie, L.js, L.je);
      bad |= check_frame_grid(L, 1, 0, 1, 0);
    }
    bad |= check_frame_pts(-1, n + 2, 6, n - 4);
    bad |= check_frame_pts(1, n + 1, 3, n - 1);
    bad |= check_frame_pts(1, n + 1, 1, 0);
  }
  printf(bad ? "FAILED\n" : "OK\n");
  return bad;
}

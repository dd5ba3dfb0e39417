// Host emulation of the lane arithmetic of symm_panel16_kernel's FMA columns (nbed_b200/csrc/jk_panel16.cuh):
// fragment indexing (tile_swz / xoff / yoff), the reduce-scatter over the four tq lanes and the row a lane finally
// owns (mi_own).  32 lanes are stepped in lock step; __shfl_xor_sync becomes an array lookup.  g++ -std=c++17.
#include <cmath>
#include <cstdio>
#include <random>
#include <vector>

static int tile_swz(int r, int c) { return r * 32 + (c ^ ((r & 3) << 2)); }

int main() {
  std::mt19937 rng(3);
  std::normal_distribution<double> N(0, 1);
  int fails = 0;
  for (int row_task = 0; row_task < 2; ++row_task) {
    std::vector<double> tile(1024), C(32 * 2);  // tile(r, c) swizzled; C[contraction][fi]
    std::vector<double> plain(1024);
    for (int r = 0; r < 32; ++r)
      for (int c = 0; c < 32; ++c) tile[tile_swz(r, c)] = plain[r * 32 + c] = N(rng);
    for (auto& x : C) x = N(rng);
    // per-lane partial sums, exactly the kernel's loops
    double pf[32][2][4] = {};
    for (int lane = 0; lane < 32; ++lane) {
      const int gq = lane >> 2, tq = lane & 3;
      int xoff[4], yoff[4];
      for (int q4 = 0; q4 < 4; ++q4) xoff[q4] = 4 * (q4 ^ (gq & 3)) + tq;
      for (int mi = 0; mi < 4; ++mi) yoff[mi] = (8 * mi + gq) ^ (4 * tq);
      for (int ks = 0; ks < 8; ++ks)
        for (int mi = 0; mi < 4; ++mi) {
          const double a = row_task ? tile[(8 * mi + gq) * 32 + ((ks >> 2) << 4) + xoff[ks & 3]] : tile[(4 * ks + tq) * 32 + yoff[mi]];
          for (int fi = 0; fi < 2; ++fi) pf[lane][fi][mi] = std::fma(a, C[(4 * ks + tq) * 2 + fi], pf[lane][fi][mi]);
        }
    }
    // reduce-scatter (xk16_reduce_scatter) with emulated shuffles
    double out[32][2] = {};  // out[row in panel][fi]
    for (int fi = 0; fi < 2; ++fi) {
      double k0[32], k1[32], s0[32], s1[32], keep[32], send[32];
      for (int lane = 0; lane < 32; ++lane) {
        const int tq = lane & 3;
        const bool odd = tq & 1;
        k0[lane] = odd ? pf[lane][fi][2] : pf[lane][fi][0];
        s0[lane] = odd ? pf[lane][fi][0] : pf[lane][fi][2];
        k1[lane] = odd ? pf[lane][fi][3] : pf[lane][fi][1];
        s1[lane] = odd ? pf[lane][fi][1] : pf[lane][fi][3];
      }
      double k0n[32], k1n[32];
      for (int lane = 0; lane < 32; ++lane) {
        k0n[lane] = k0[lane] + s0[lane ^ 1];
        k1n[lane] = k1[lane] + s1[lane ^ 1];
      }
      for (int lane = 0; lane < 32; ++lane) {
        const bool hi = (lane & 3) & 2;
        keep[lane] = hi ? k1n[lane] : k0n[lane];
        send[lane] = hi ? k0n[lane] : k1n[lane];
      }
      for (int lane = 0; lane < 32; ++lane) {
        const int gq = lane >> 2, tq = lane & 3;
        const int mi_own = ((tq & 1) << 1) | (tq >> 1);
        out[8 * mi_own + gq][fi] += keep[lane] + send[lane ^ 2];
      }
    }
    double err = 0;
    for (int x = 0; x < 32; ++x)
      for (int fi = 0; fi < 2; ++fi) {
        double ref = 0;
        for (int y = 0; y < 32; ++y) ref += (row_task ? plain[x * 32 + y] : plain[y * 32 + x]) * C[y * 2 + fi];
        err = std::fmax(err, std::fabs(ref - out[x][fi]));
      }
    if (!(err < 1e-12)) {
      ++fails;
      printf("FAIL %s task: max error %.3e\n", row_task ? "row" : "column", err);
    }
  }
  printf("panel16_emul fails=%d\n", fails);
  return fails != 0;
}

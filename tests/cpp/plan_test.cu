#include <cstdio>
#include <cstdint>
#include "../../nbed_b200/csrc/jk_panel16.cuh"
int main() {
  int bad = 0;
  for (int nb = 1; nb <= 96; ++nb)
    for (int S = 2; S <= 16; ++S) {
      auto seq = nbd::build_tile_sequence(nb);
      auto pl = nbd::build_panel_plan(nb, S, seq);
      if (pl.S != S) { printf("FAIL nb=%d S=%d\n", nb, S); ++bad; }
    }
  // experimental 16-warp kernel: its own task lists on the 16 x 16 super-block order, the 8-warp lists on that order,
  // and the sequence itself (a permutation of the lower-triangle tiles; nw = 8 reproduces the default order)
  for (int nb = 1; nb <= 48; ++nb) {
    auto seq16 = nbd::build_tile_sequence_w(nb, 16);
    if (nbd::build_tile_sequence_w(nb, 8) != nbd::build_tile_sequence(nb)) { printf("FAIL seq_w(8) nb=%d\n", nb); ++bad; }
    std::vector<int> seen(nb * nb, 0);
    for (int t : seq16) {
      const int I = t >> 16, J = t & 0xffff;
      if (I >= nb || J > I || seen[I * nb + J]++) { printf("FAIL seq16 nb=%d\n", nb); ++bad; break; }
    }
    if ((int)seq16.size() != nb * (nb + 1) / 2) { printf("FAIL seq16 size nb=%d\n", nb); ++bad; }
    for (int S = 2; S <= 16; ++S) {
      if (nbd::build_panel_plan_w(nb, S, seq16, 16).S != S) { printf("FAIL plan16 nb=%d S=%d\n", nb, S); ++bad; }
      if (nbd::build_panel_plan(nb, S, seq16).S != S) { printf("FAIL plan8-on-seq16 nb=%d S=%d\n", nb, S); ++bad; }
    }
  }
  auto pl = nbd::build_panel_plan(43, 13, nbd::build_tile_sequence(43));
  printf("nb=43 S=13: %zu events for 946 tiles; bad=%d\n", pl.events.size(), bad);
  return bad != 0;
}

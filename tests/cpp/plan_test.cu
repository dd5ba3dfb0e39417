#include <cstdio>
#include <cstdint>
#include "../../nbed_b200/csrc/jk.cuh"
int main() {
  int bad = 0;
  for (int nb = 1; nb <= 96; ++nb)
    for (int S = 2; S <= 16; ++S) {
      auto seq = nbd::build_tile_sequence(nb);
      auto pl = nbd::build_panel_plan(nb, S, seq);
      if (pl.S != S) { printf("FAIL nb=%d S=%d\n", nb, S); ++bad; }
    }
  auto pl = nbd::build_panel_plan(43, 13, nbd::build_tile_sequence(43));
  printf("nb=43 S=13: %zu events for 946 tiles; bad=%d\n", pl.events.size(), bad);
  return bad != 0;
}

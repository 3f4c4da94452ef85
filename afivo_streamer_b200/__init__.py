"""afmg-b200: B200-native FAS multigrid behind afivo's mg_t / mg_fas_fmg interface."""

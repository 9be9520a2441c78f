// k_tree_fast instantiations for padded q = 16, mode "g", s = 3 (see ghm_tree_kernel.cuh / ghm_tree_fast.cuh)
#include "ghm_tree_kernel.cuh"

GHM_TREE_FAST_DEFINE(16, g, MODE_GIVEN, true, 3)

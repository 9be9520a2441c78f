// k_tree_fast instantiations for padded q = 8, mode "g", s = 4 (see ghm_tree_kernel.cuh / ghm_tree_fast.cuh)
#include "ghm_tree_kernel.cuh"

GHM_TREE_FAST_DEFINE(8, g, MODE_GIVEN, true, 4)

// k_tree_fast instantiations for padded q = 10, mode "g", s = 2 (see ghm_tree_kernel.cuh / ghm_tree_fast.cuh)
#include "ghm_tree_kernel.cuh"

GHM_TREE_FAST_DEFINE(10, g, MODE_GIVEN, true, 2)

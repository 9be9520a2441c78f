// k_tree_fast instantiations for padded q = 8, mode "sb", s = 3 (see ghm_tree_kernel.cuh / ghm_tree_fast.cuh)
#include "ghm_tree_kernel.cuh"

GHM_TREE_FAST_DEFINE(8, sb, MODE_PHILOX, true, 3)

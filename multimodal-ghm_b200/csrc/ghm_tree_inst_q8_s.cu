// k_tree2 instantiations for padded q = 8, mode "s" (see ghm_tree_kernel.cuh)
#include "ghm_tree_kernel.cuh"

GHM_TREE_DEFINE(8, s, MODE_PHILOX, false)

// k_tree2 instantiations for padded q = 10, mode "s" (see ghm_tree_kernel.cuh)
#include "ghm_tree_kernel.cuh"

GHM_TREE_DEFINE(10, s, MODE_PHILOX, false)

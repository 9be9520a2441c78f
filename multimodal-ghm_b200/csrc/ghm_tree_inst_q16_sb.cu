// k_tree2 instantiations for padded q = 16, mode "sb" (see ghm_tree_kernel.cuh)
#include "ghm_tree_kernel.cuh"

GHM_TREE_DEFINE_SPLIT(16, sb, MODE_PHILOX, true)

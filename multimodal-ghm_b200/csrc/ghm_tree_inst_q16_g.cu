// k_tree2 instantiations for padded q = 16, mode "g" (see ghm_tree_kernel.cuh)
#include "ghm_tree_kernel.cuh"

GHM_TREE_DEFINE_SPLIT(16, g, MODE_GIVEN, true)

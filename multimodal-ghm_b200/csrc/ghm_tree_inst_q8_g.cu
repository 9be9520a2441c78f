// k_tree2 instantiations for padded q = 8, mode "g" (see ghm_tree_kernel.cuh)
#include "ghm_tree_kernel.cuh"

GHM_TREE_DEFINE_SPLIT(8, g, MODE_GIVEN, true)

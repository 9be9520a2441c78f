// k_dns2 instantiations for padded q = 8, s = 2, 3, 4 (see ghm_dns2_kernel.cuh)
#include "ghm_dns2_kernel.cuh"

GHM_DNS2_DEFINE(8)

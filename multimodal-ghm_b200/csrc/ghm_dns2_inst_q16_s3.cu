// k_dns2 instantiation for padded q = 16, s = 3 (see ghm_dns2_kernel.cuh; one file per s: each takes minutes to compile)
#include "ghm_dns2_kernel.cuh"

GHM_DNS2_DEFINE_S(16, 3)

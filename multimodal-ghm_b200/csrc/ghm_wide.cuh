// ghm_wide.cuh -- internal interface of the wide (16 < q <= 256) belief-propagation path.
#pragma once
#include "ghm_common.cuh"

int64_t ghm_wide_cls_workspace_bytes(const ghm_model* m, int64_t B);
int ghm_wide_bp_cls(const ghm_model* m, int64_t B, const void* leaves, int leaf_dtype, float* post, float* root_hd,
                    void* workspace, cudaStream_t st);
int64_t ghm_wide_dns_workspace_bytes(const ghm_model* m, int64_t B);
int ghm_wide_bp_dns(const ghm_model* m, int64_t B, const float* z, float sigma, const float* ext, float* mean, float* root_bu,
                    void* workspace, cudaStream_t st);

// Y[node][b][:] = X[node][b][:] @ W[mat(level, node)]  for the n_nodes nodes of `level`;
// down = 0: W[k][n] = T[n][k] (child -> parent, `T @ m`), down = 1: W[k][n] = T[k][n] (parent -> child, `T.T @ m`)
int ghm_wide_gemm(const ghm_model* m, int64_t B, int level, int n_nodes, int down, const float* X, float* Y, cudaStream_t st);
// tcgen05 TF32 / BF16 variant (ghm_wide_tc.cu); returns GHM_EUNSUP for shapes it does not cover
int ghm_wide_gemm_tc(const ghm_model* m, int64_t B, int level, int n_nodes, int down, const float* X, float* Y, cudaStream_t st);

// BP_DNS leaf level with the likelihood / cavity / posterior-mean row kernels folded into the TF32 GEMM
// (ghm_wide_tc.cu); GHM_EUNSUP when the model is not in TF32 mode or the shape is not covered
int ghm_wide_leaf_up_fused(const ghm_model* m, int64_t B, const float* z, float c2, float* Uout, cudaStream_t st);
int ghm_wide_leaf_down_fused(const ghm_model* m, int64_t B, const float* z, float c2, const float* BUpar, const float* U,
                             float* mean, cudaStream_t st);
// internal level l of the downward pass: beliefs BU[l] = hd * (T^T (BU[l-1] / u)) / max in one kernel
int ghm_wide_down_fused(const ghm_model* m, int64_t B, int level, const float* BUpar, const float* U, const float* H, float* BU,
                        cudaStream_t st);
// BP_CLS, depth L-1 (s <= 3): the leaf-row products are formed inside the GEMM, U[j] = T_j h_j comes out directly
int ghm_wide_cls_leaf_fused(const ghm_model* m, int64_t B, const void* leaves, int leaf_dtype, float* U, cudaStream_t st);

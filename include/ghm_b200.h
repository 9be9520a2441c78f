/* ghm_b200.h -- C ABI of the B200-native JGHM sample + exact-BP library (libghm_b200.so).
 *
 * The reference (willcai7/Multimodal-GHM) has NO plugin / FFI interface for this
 * path: the boundary is the Python module surface
 *   src/ghmclip/data/data_random_GHM.py   (re-exported by src/ghmclip/data/__init__.py:5)
 * Each entry point below names the reference function it replaces (file:line,
 * relative to the reference root).  The Python facade in
 * multimodal-ghm_b200/ghm_b200/ binds these with ctypes and mirrors the reference
 * class/method surface on top of them.
 *
 * Conventions
 *   - every function returns 0 on success, a GHM_E* code otherwise;
 *     ghm_last_error() returns a thread-local message for the last failure.
 *   - device entry points take raw DEVICE pointers + sizes + a cudaStream_t passed
 *     as void*; they never allocate or free caller memory and never synchronise.
 *   - ghm_host_* entry points take HOST pointers, do their own H2D/D2H on the
 *     model's internal stream and synchronise before returning.
 *   - batch-major layouts: leaves [B, n_L]; posteriors [B, q]; z / mean [B, n_L];
 *     next-token posteriors [B, n_L-1, q]; guides [B, n_L(-1), C]  (row-major).
 *   - there is no CPU fallback anywhere: without a CUDA device every compute entry
 *     point fails with GHM_ECUDA.
 */
#ifndef GHM_B200_H
#define GHM_B200_H

#include <stdint.h>

#if defined(__GNUC__)
#define GHM_API __attribute__((visibility("default")))
#else
#define GHM_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define GHM_OK        0
#define GHM_EINVAL    1   /* bad argument (shape, dtype code, null pointer, unsupported q/L/s) */
#define GHM_ECUDA     2   /* CUDA runtime error (message has the cudaError string) */
#define GHM_ENOMEM    3
#define GHM_EUNSUP    4   /* configuration outside what this build instantiates */

/* leaf storage dtype codes */
#define GHM_LEAF_I64  0   /* torch.long, the reference API dtype (data_random_GHM.py:697) */
#define GHM_LEAF_U8   1   /* compact device format (q <= 256) */

/* arithmetic of the wide-q (16 < q <= 256) row-GEMMs, ghm_model_set_gemm_mode */
#define GHM_GEMM_F32   0   /* FP32 CUDA cores: <= 1e-5 relative against the float64 reference (default) */
#define GHM_GEMM_TF32  1   /* tcgen05 kind::tf32, FP32 accumulate in TMEM: <= 2e-3 relative */
#define GHM_GEMM_BF16  2   /* tcgen05 kind::f16 (BF16 operands), FP32 accumulate: <= 2e-2 relative */

/* root modes for ghm_sample */
#define GHM_ROOT_GIVEN    0   /* root_in supplied  (reference GHMTree(root=...), :153-156) */
#define GHM_ROOT_PRIOR    1   /* draw from the model prior p_y (:158) */
#define GHM_ROOT_UNIFORM  2   /* np.random.choice(q,size=B): uniform, ignores p_y (:674,758,858,906) */
#define GHM_ROOT_SHARED   3   /* (ghm_sample_paired) uniform roots re-drawn from the PARTNER modality's Philox key */

typedef struct ghm_model ghm_model_t;   /* opaque; owns only the device-side tables */

GHM_API const char* ghm_last_error(void);
GHM_API const char* ghm_version(void);
/* number of CUDA devices visible, or -1 on a CUDA error */
GHM_API int ghm_device_count(void);

/* ---- model tables -------------------------------------------------------------
 * Replaces the table side of SingleSampler/DoubleSampler.__init__ (:621-634,
 * :645-658): the caller generates the transition matrices (GenTransition, :43-89)
 * on the host and hands them over as float64.
 *   T_host : [n_mat, q, q] row-major, T[m][a][b] = P(child=b | parent=a)
 *            translation-invariant (ti=1): n_mat = L*s, m = (level-1)*s + child
 *            per-edge (ti=0):              n_mat = E,   m = BFS edge index
 *   p_y    : [q] root prior (NULL -> uniform)
 * The library derives and uploads: f32 T, f32 T^T, f32 log T^T, f64 CDF (sequential
 * cumsum, bit-identical to np.cumsum, :165) and Walker alias tables (Philox mode, O(1) draws). */
GHM_API int ghm_model_create(ghm_model_t** out, int n_layer, int n_child, int q, int ti,
                     const double* T_host, const double* p_y_host, int device);
GHM_API int ghm_model_destroy(ghm_model_t* m);
/* New tables for an existing model of the same (L, s, q, ti): what constructing a new sampler per p_flip does in
 * the reference's sweeps (figures/eval-clip-ood.py:76-79, eval-cdm-ood.py:104-109, eval-vlm-ood.py:104-109).
 * Tables are derived on the host and uploaded with ONE async copy from pinned memory on `stream`
 * (ghm_model_table_bytes() bytes), followed on the same stream by one small kernel that tabulates the message of a
 * depth-(L-1) node as a function of its s leaf states (:191-208; the leaf memo of the fused sampler + BP_CLS kernel);
 * kernels enqueued on `stream` afterwards use them. */
GHM_API int ghm_model_update(ghm_model_t* m, const double* T_host, const double* p_y_host, void* stream);
GHM_API int64_t ghm_model_table_bytes(const ghm_model_t* m);
/* wide-q models only (ignored for q <= 16): pick the GEMM arithmetic, see GHM_GEMM_* */
GHM_API int ghm_model_set_gemm_mode(ghm_model_t* m, int mode);
GHM_API int ghm_model_info(const ghm_model_t* m, int* n_layer, int* n_child, int* q, int* ti,
                   int64_t* n_leaves, int64_t* n_edges);
/* sticky device-side status word (bit0: a leaf/root value >= q was clamped).
 * Synchronises `stream`; for tests / debugging only. */
GHM_API int ghm_model_status(ghm_model_t* m, void* stream, int* status_out);

/* ---- K1: sampler  (GHMTree.gen_values, :145-165) -------------------------------
 *   U        : f64 [E, B] uniforms in BFS-edge-major order = the reference's E
 *              sequential np.random.rand(B,1) calls -> PARITY mode, f64 compare,
 *              leaves bit-identical to the reference.  NULL -> Philox4x32-10 keyed by
 *              (seed; global tree index = tree_offset + b, level, node).
 *   root_in  : i64 [B] when root_mode == GHM_ROOT_GIVEN, else ignored
 *   root_out : i64 [B] or NULL;  leaves_out : [B, n_L] of leaf_dtype or NULL
 *   post_out / root_hd_out : f32 [B, q] or NULL -> when given, BP_CLS (:185-221) is
 *              fused into the same pass (leaf states never leave the SM). */
GHM_API int ghm_sample(const ghm_model_t* m, int64_t B, int root_mode, const int64_t* root_in,
               const double* U, uint64_t seed, uint64_t tree_offset,
               int64_t* root_out, void* leaves_out, int leaf_dtype,
               float* post_out, float* root_hd_out, void* stream);

/* ClipSampler's image-side root layout in ONE launch (:759-760, `np.append(text_root[:2n], choice(q, n(K-1)))`):
 * trees [0, n_given) take root_in[b], trees [n_given, B) draw uniform roots.  Philox mode only. */
GHM_API int ghm_sample_mixed(const ghm_model_t* m, int64_t B, int64_t n_given, const int64_t* root_in,
                     uint64_t seed, uint64_t tree_offset, int64_t* root_out, void* leaves_out, int leaf_dtype,
                     float* post_out, float* root_hd_out, void* stream);

/* The image side of a text/image pair WITHOUT a dependency on the text launch (:858-861, :759-760): trees
 * [0, n_shared) re-draw the uniform root the partner modality drew (Philox key `root_seed` = the partner's seed,
 * same global tree index), trees [n_shared, B) draw their own uniform roots.  The two modalities can then be
 * launched on different streams and their CTAs fill each other's tail waves.  Philox mode only. */
/* A SHARD of a block-structured batch in one launch: local tree b has the global Philox index
 * tree_offset + (b / blk_len) * blk_stride + b % blk_len.  ClipSampler's layout [match1 | match2 | K-1 negatives] (:758-764)
 * sharded on the pair index (pairs [lo, hi) of n on this rank) is blk_len = hi - lo, blk_stride = n, tree_offset = base + lo:
 * the rank draws exactly the trees the unsharded launch would have drawn for those pairs.  root_mode as ghm_sample
 * (GHM_ROOT_GIVEN: every root from root_in) or GHM_ROOT_SHARED with n_given local trees re-drawing the partner's roots
 * (ghm_sample_paired).  Philox mode only. */
GHM_API int ghm_sample_blocked(const ghm_model_t* m, int64_t B, int64_t blk_len, int64_t blk_stride, int root_mode,
                       int64_t n_given, const int64_t* root_in, uint64_t root_seed, uint64_t seed, uint64_t tree_offset,
                       int64_t* root_out, void* leaves_out, int leaf_dtype, float* post_out, float* root_hd_out,
                       void* stream);

/* ClipSampler.get_Bayes (:786-817) as ONE device-side call: sample both modalities in the block layout (:758-760) with
 * the root-posterior BP fused (:767-768) -- the image launch on `side_stream`, forked from and joined to `stream` --
 * and accumulate the symmetric K-way contrastive loss (:794-817) of pairs [pair_lo, pair_hi) into `sums` (device
 * double[3]: sum, sum of squares, count).  All buffers are the caller's ((pair_hi - pair_lo) * (K + 1) trees per
 * modality; t_root / t_leaves / i_leaves may be null); nothing is copied to the host, nothing synchronises.  A rank's
 * shard draws exactly the trees the whole evaluation draws for those pairs (ghm_sample_blocked).  Philox mode. */
GHM_API int ghm_clip_bayes(const ghm_model_t* text, const ghm_model_t* image, int64_t n, int K, int64_t pair_lo,
                   int64_t pair_hi, uint64_t seed, uint64_t tree_offset, int64_t* t_root, void* t_leaves, void* i_leaves,
                   int leaf_dtype, float* t_pp, float* i_pp, double* sums, void* stream, void* side_stream);

GHM_API int ghm_sample_paired(const ghm_model_t* m, int64_t B, int64_t n_shared, uint64_t root_seed, uint64_t seed,
                      uint64_t tree_offset, int64_t* root_out, void* leaves_out, int leaf_dtype,
                      float* post_out, float* root_hd_out, void* stream);

/* ---- K2: root posterior  (GHMTree.BP_CLS, :185-221) ----------------------------
 *   post     : f32 [B, q]  p(root | leaves)               (posterior_probability_CLS^T)
 *   root_hd  : f32 [B, q]  max-shifted log-likelihood, NO prior (root_node.hd_message^T,
 *              the cross-modal "external" message, :871,919)   (either may be NULL) */
GHM_API int64_t ghm_bp_cls_workspace_bytes(const ghm_model_t* m, int64_t B);   /* 0 for q <= 16 */
GHM_API int ghm_bp_cls(const ghm_model_t* m, int64_t B, const void* leaves, int leaf_dtype,
               float* post, float* root_hd, void* workspace, void* stream);

/* ---- K3: Gaussian denoiser  (GHMTree.BP_DNS, :467-523) -------------------------
 *   z : f32 [B, n_L]; ext : f32 [B, q] external root log-message or NULL;
 *   mean : f32 [B, n_L] posterior mean of every leaf (posterior_mean_DNS^T)
 *   root_bu : f32 [B, q] or NULL -- root_node.hd_message after BP_DNS: the max-shifted root hd plus ext, NOT re-shifted
 *          (the reference's root bu_message aliases hd_message, :501-506)
 *   workspace : device scratch of ghm_bp_dns_workspace_bytes(m, B) bytes */
GHM_API int64_t ghm_bp_dns_workspace_bytes(const ghm_model_t* m, int64_t B);
GHM_API int ghm_bp_dns(const ghm_model_t* m, int64_t B, const float* z, float sigma, const float* ext,
               float* mean, float* root_bu, void* workspace, void* stream);

/* ---- K4: next-token posterior  (GHMTree.BP_NWP_autoregressive, :336-463) -------
 *   pp : f32 [B, n_L-1, q],  pp[b,t,:] = p(leaf_{t+1} | leaves_{<=t}, ext) */
GHM_API int64_t ghm_bp_nwp_workspace_bytes(const ghm_model_t* m, int64_t B);
GHM_API int ghm_bp_nwp(const ghm_model_t* m, int64_t B, const void* leaves, int leaf_dtype,
               const float* ext, float* pp, void* workspace, void* stream);

/* ---- K5: guide tensors  (GHMTree.guided_info, :526-592; NWP guides :357-459) ---
 * Log-domain BP with the reference's exact shift conventions, written straight in
 * the [B, n_L, C] f32 layout the guided losses index.  `guides` is an array of
 * DEVICE pointers living in HOST memory (read at enqueue time).
 *   cls : L tensors [B,n_L,q]           (depth L-1 .. 0 hd)
 *   dns : 2L+1 tensors [B,n_L,2q] x L, [B,n_L,2q], [B,n_L,3q] x L
 *   nwp : 2L+1 tensors [B,n_L-1,{q,2q..,2q,q..}]                                  */
GHM_API int ghm_guides_cls(const ghm_model_t* m, int64_t B, const void* leaves, int leaf_dtype,
                   float* const* guides, float* post, float* root_hd, void* stream);
GHM_API int64_t ghm_guides_dns_workspace_bytes(const ghm_model_t* m, int64_t B);
GHM_API int ghm_guides_dns(const ghm_model_t* m, int64_t B, const float* z, float sigma, const float* ext,
                   float* const* guides, float* mean, void* workspace, void* stream);
GHM_API int64_t ghm_guides_nwp_workspace_bytes(const ghm_model_t* m, int64_t B);
GHM_API int ghm_guides_nwp(const ghm_model_t* m, int64_t B, const void* leaves, int leaf_dtype,
                   const float* ext, float* const* guides, float* pp, void* workspace, void* stream);

/* ---- K6: risk reductions --------------------------------------------------------
 * Every risk kernel ACCUMULATES (atomically, f64) into sums[3] = {sum, sum of squares,
 * count}; the caller zeroes it, and may all-reduce it across ranks before mean / SE.
 *   clip : PPCLIPLoss == ClipSampler.get_Bayes == clip_loss_compute (:13-41,:794-817,
 *          :819-844).  t_pp, i_pp : f32 [n*(K+1), q] in the ClipSampler block layout
 *          (:758-760).  Pairs [pair_lo, pair_hi) only (multi-GPU: shard on the pair).
 *   cdm  : ConditionalDenoiseSampler.get_Bayes (:886-894): per tree sum_leaf (mean-x)^2
 *   ce   : token / root cross-entropy -log p[b, target[b]]  (NextWordPredictSampler.get_Bayes
 *          :931-942 with rows = B*(n_L-1); ClassificationSampler.get_Bayes :707-720)     */
GHM_API int ghm_risk_clip(const float* t_pp, const float* i_pp, int64_t n, int K, int q,
                  int64_t pair_lo, int64_t pair_hi, double* sums, void* stream);
GHM_API int ghm_risk_cdm(const float* mean, const void* leaves, int leaf_dtype, int64_t B, int64_t n_leaves,
                 double* sums, void* stream);
GHM_API int ghm_risk_ce(const float* pp, const void* target, int leaf_dtype, int64_t rows, int q,
                int64_t target_stride, int64_t target_offset, int64_t row_group,
                double* sums, void* stream);

/* zero-shot classification risk (figures/eval-zsc-risk.py:66-83): the image root posterior i_pp f32 [B, q] is pushed
 * down the leftmost path of the TEXT tree (x <- x @ T_l[0], every level) and scored against the first text leaf:
 * accumulates -log x[t_leaves[b, 0]] into sums[3]. */
GHM_API int ghm_risk_zsc(const ghm_model_t* text, int64_t B, const float* i_pp, const void* t_leaves, int leaf_dtype,
                 double* sums, void* stream);

/* ---- Gaussian observations  (image_tree_noise, :733,:867) -----------------------
 * z[b,i] = leaves[b,i] + sigma * N(0,1), Philox stream 1 keyed like ghm_sample. */
GHM_API int ghm_gauss_noise(const ghm_model_t* m, int64_t B, const void* leaves, int leaf_dtype, float sigma,
                    uint64_t seed, uint64_t tree_offset, float* z, void* stream);

/* ---- host-buffer entry points (end-to-end measurement / non-torch callers) ------
 * ClipSampler.get_Bayes (:786-817) in one call: host roots are NOT needed (roots are
 * drawn on the device, uniform as in :758-759).  sums_host[3] receives the reduced
 * {sum, sum sq, count}.  Optional host outputs (may be NULL): leaves as leaf_dtype
 * [n*(K+1), n_L] and posteriors f32 [n*(K+1), q] for each modality. */
GHM_API int ghm_host_clip_bayes(const ghm_model_t* text, const ghm_model_t* image, int64_t n, int K,
                        uint64_t seed, uint64_t tree_offset, double* sums_host,
                        void* t_leaves_host, void* i_leaves_host, int leaf_dtype,
                        float* t_pp_host, float* i_pp_host);

#ifdef __cplusplus
}
#endif
#endif /* GHM_B200_H */

// ghm_host.cu -- host-buffer entry points (end-to-end path: H2D/D2H inside the call).
//
// ghm_host_clip_bayes == ClipSampler.get_Bayes (reference src/ghmclip/data/data_random_GHM.py:786-817)
// as ONE call: sample both modalities in the ClipSampler block layout (:758-760), fused
// root-posterior BP (:767-768), contrastive reduction (:794-817), 24 bytes back to the host --
// plus, when asked, the leaves and posteriors ClipSampler.get_batch returns (:779-784).
// Not thread-safe per model (uses the text model's internal stream and scratch).
#include "ghm_common.cuh"

#define GHM_IMAGE_SEED_XOR 0x1234567887654321ull

static int ensure_dscratch(ghm_model* m, size_t bytes) {
    if (m->d_scratch_bytes >= bytes) return GHM_OK;
    if (m->d_scratch) { cudaFree(m->d_scratch); m->d_scratch = nullptr; m->d_scratch_bytes = 0; }
    cudaError_t e = cudaMalloc(&m->d_scratch, bytes);
    if (e != cudaSuccess) return ghm_fail(GHM_ENOMEM, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    m->d_scratch_bytes = bytes;
    return GHM_OK;
}

static size_t up256(size_t x) { return (x + 255) / 256 * 256; }

// ------------------------------------------------------------------------------------------------
// ghm_clip_bayes: the device-side form of the same evaluation -- every buffer is the caller's, nothing is copied to the
// host and nothing synchronises.  One call enqueues: image sampling + fused BP on `side_stream` (forked from `stream`
// through the image model's ordering event), text sampling + fused BP on `stream`, the join, and the contrastive
// reduction accumulated into `sums`.  Pairs [pair_lo, pair_hi) of the n-pair block layout are evaluated (a rank's
// shard; the whole evaluation is pair_lo = 0, pair_hi = n): ghm_sample_blocked gives them the global Philox indices
// of the unsharded launch.  Replaces five Python-level calls of the facade's get_Bayes with one.
// ------------------------------------------------------------------------------------------------
extern "C" int ghm_clip_bayes(const ghm_model_t* text_c, const ghm_model_t* image_c, int64_t n, int K, int64_t pair_lo,
                              int64_t pair_hi, uint64_t seed, uint64_t tree_offset, int64_t* t_root, void* t_leaves,
                              void* i_leaves, int leaf_dtype, float* t_pp, float* i_pp, double* sums, void* stream,
                              void* side_stream) {
    ghm_model* text = const_cast<ghm_model*>(text_c);
    ghm_model* image = const_cast<ghm_model*>(image_c);
    if (!text || !image || !t_pp || !i_pp || !sums) return ghm_fail(GHM_EINVAL, "ghm_clip_bayes: null argument");
    if (text->d.q != image->d.q) return ghm_fail(GHM_EINVAL, "text and image models disagree on q");
    if (text->device != image->device) return ghm_fail(GHM_EINVAL, "text and image models live on different devices");
    if (n <= 0 || K < 2 || pair_lo < 0 || pair_hi > n || pair_lo > pair_hi)
        return ghm_fail(GHM_EINVAL, "ghm_clip_bayes: bad n=%lld K=%d pairs [%lld, %lld)", (long long)n, K, (long long)pair_lo,
                        (long long)pair_hi);
    const int64_t nl = pair_hi - pair_lo;
    if (nl == 0) return GHM_OK;
    GhmDeviceGuard guard(text->device);
    cudaStream_t st = (cudaStream_t)stream, st2 = (cudaStream_t)side_stream;
    const int64_t Bl = nl * (K + 1);
    const uint64_t iseed = seed ^ GHM_IMAGE_SEED_XOR;
    int rc;
    if (st2 && st2 != st) {
        GHM_CUDA_TRY(cudaEventRecord(image->order_ev, st));            // fork: the side stream starts after what `stream` holds
        GHM_CUDA_TRY(cudaStreamWaitEvent(st2, image->order_ev, 0));
    } else {
        st2 = st;
    }
    rc = ghm_sample_blocked(image, Bl, nl, n, GHM_ROOT_SHARED, 2 * nl, nullptr, seed, iseed, tree_offset + (uint64_t)pair_lo,
                            nullptr, i_leaves, leaf_dtype, i_pp, nullptr, st2);
    if (rc) return rc;
    rc = ghm_sample_blocked(text, Bl, nl, n, GHM_ROOT_UNIFORM, 0, nullptr, 0, seed, tree_offset + (uint64_t)pair_lo, t_root,
                            t_leaves, leaf_dtype, t_pp, nullptr, st);
    if (rc) return rc;
    if (st2 != st) {
        GHM_CUDA_TRY(cudaEventRecord(text->order_ev, st2));             // join
        GHM_CUDA_TRY(cudaStreamWaitEvent(st, text->order_ev, 0));
    }
    return ghm_risk_clip(t_pp, i_pp, nl, K, text->d.q, 0, nl, sums, st);
}

extern "C" int ghm_host_clip_bayes(const ghm_model_t* text_c, const ghm_model_t* image_c, int64_t n, int K,
                                   uint64_t seed, uint64_t tree_offset, double* sums_host, void* t_leaves_host,
                                   void* i_leaves_host, int leaf_dtype, float* t_pp_host, float* i_pp_host) {
    ghm_model* text = const_cast<ghm_model*>(text_c);
    const ghm_model* image = image_c;
    if (!text || !image || !sums_host) return ghm_fail(GHM_EINVAL, "ghm_host_clip_bayes: null argument");
    if (text->d.q != image->d.q) return ghm_fail(GHM_EINVAL, "text and image models disagree on q");
    if (text->device != image->device) return ghm_fail(GHM_EINVAL, "text and image models live on different devices");
    if (n <= 0 || K < 2) return ghm_fail(GHM_EINVAL, "ghm_host_clip_bayes: bad n=%lld K=%d", (long long)n, K);
    if (leaf_dtype != GHM_LEAF_I64 && leaf_dtype != GHM_LEAF_U8) return ghm_fail(GHM_EINVAL, "bad leaf_dtype %d", leaf_dtype);
    const int q = text->d.q;
    const int64_t B = n * (K + 1);
    const size_t lsz = leaf_dtype == GHM_LEAF_I64 ? 8 : 1;
    int prev = 0;
    cudaGetDevice(&prev);
    GHM_CUDA_TRY(cudaSetDevice(text->device));
    struct Restore { int p; ~Restore() { cudaSetDevice(p); } } restore{prev};

    size_t off = 0;
    auto take = [&](size_t b) { size_t o = off; off = up256(off + b); return o; };
    const size_t o_sums = take(3 * sizeof(double));
    const size_t o_root = take((size_t)B * 8);
    const size_t o_tpp = take((size_t)B * q * 4), o_ipp = take((size_t)B * q * 4);
    const size_t o_tl = t_leaves_host ? take((size_t)B * text->d.n_leaves * lsz) : 0;
    const size_t o_il = i_leaves_host ? take((size_t)B * image->d.n_leaves * lsz) : 0;
    int rc = ensure_dscratch(text, off);
    if (rc) return rc;
    char* base = (char*)text->d_scratch;
    double* d_sums = (double*)(base + o_sums);
    int64_t* d_root = (int64_t*)(base + o_root);
    float* d_tpp = (float*)(base + o_tpp);
    float* d_ipp = (float*)(base + o_ipp);
    void* d_tl = t_leaves_host ? (void*)(base + o_tl) : nullptr;
    void* d_il = i_leaves_host ? (void*)(base + o_il) : nullptr;
    cudaStream_t st = text->stream;

    GHM_CUDA_TRY(cudaMemsetAsync(d_sums, 0, 3 * sizeof(double), st));
    // text: B trees, uniform roots (:758) on the text model's stream; image: the first 2n trees re-draw the text roots
    // from the text key, the other (K-1)n draw fresh uniform roots (:759-760) -- independent of the text launch, so it
    // runs concurrently on the image model's stream and the two kernels fill each other's tail waves
    const uint64_t iseed = seed ^ GHM_IMAGE_SEED_XOR;
    cudaStream_t st2 = image->stream;
    // a ghm_model_update enqueued on the caller's stream may still be copying the tables: both internal streams wait
    // for the upload fences (upload_done is only ever recorded by ghm_model_update), ordering uses order_ev
    GHM_CUDA_TRY(cudaStreamWaitEvent(st, text->upload_done, 0));
    GHM_CUDA_TRY(cudaStreamWaitEvent(st2, image->upload_done, 0));
    GHM_CUDA_TRY(cudaEventRecord(text->order_ev, st));
    GHM_CUDA_TRY(cudaStreamWaitEvent(st2, text->order_ev, 0));
    rc = ghm_sample_paired(image, B, 2 * n, seed, iseed, tree_offset, nullptr, d_il, leaf_dtype, d_ipp, nullptr, st2);
    if (rc) return rc;
    rc = ghm_sample(text, B, GHM_ROOT_UNIFORM, nullptr, nullptr, seed, tree_offset, d_root, d_tl, leaf_dtype, d_tpp,
                    nullptr, st);
    if (rc) return rc;
    GHM_CUDA_TRY(cudaEventRecord(image->order_ev, st2));
    GHM_CUDA_TRY(cudaStreamWaitEvent(st, image->order_ev, 0));
    rc = ghm_risk_clip(d_tpp, d_ipp, n, K, q, 0, n, d_sums, st);
    if (rc) return rc;
    GHM_CUDA_TRY(cudaMemcpyAsync(sums_host, d_sums, 3 * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (t_pp_host) GHM_CUDA_TRY(cudaMemcpyAsync(t_pp_host, d_tpp, (size_t)B * q * 4, cudaMemcpyDeviceToHost, st));
    if (i_pp_host) GHM_CUDA_TRY(cudaMemcpyAsync(i_pp_host, d_ipp, (size_t)B * q * 4, cudaMemcpyDeviceToHost, st));
    if (t_leaves_host)
        GHM_CUDA_TRY(cudaMemcpyAsync(t_leaves_host, d_tl, (size_t)B * text->d.n_leaves * lsz, cudaMemcpyDeviceToHost, st));
    if (i_leaves_host)
        GHM_CUDA_TRY(cudaMemcpyAsync(i_leaves_host, d_il, (size_t)B * image->d.n_leaves * lsz, cudaMemcpyDeviceToHost, st));
    GHM_CUDA_TRY(cudaStreamSynchronize(st));
    return GHM_OK;
}

// ghm_model.cu -- model tables, error plumbing, misc C-ABI entry points.
//
// Replaces the table side of the reference's SingleSampler/DoubleSampler constructors
// (src/ghmclip/data/data_random_GHM.py:621-634, :645-658): the host generates the
// transition matrices (GenTransition, :43-89) and this file derives every device table.
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "ghm_common.cuh"

static thread_local std::string g_last_error;

void ghm_set_error(const std::string& msg) { g_last_error = msg; }

int ghm_fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

extern "C" const char* ghm_last_error(void) { return g_last_error.c_str(); }
extern "C" const char* ghm_version(void) { return "ghm_b200 0.1 (sm_100a)"; }

extern "C" int ghm_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        ghm_fail(GHM_ECUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
        cudaGetLastError();
        return -1;
    }
    return n;
}

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ------------------------------------------------------------------------------------------------
// table derivation: host slab (pinned) with exactly the device slab's layout
// ------------------------------------------------------------------------------------------------
struct SlabLayout {
    size_t Tlin, TlinT, TlogT, TTp, alias, cdfd, Wup, Wdn, py, rcp, rcu, status, memo, host_bytes, bytes;
};

static SlabLayout slab_layout(const GhmDev& d) {
    SlabLayout o{};
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t r = off; off = align_up(off + bytes, 256); return r; };
    const size_t nm = (size_t)d.n_mat, QQ = (size_t)d.QP * d.QP, qq = (size_t)d.q * d.q;
    o.Tlin = take(nm * QQ * 4); o.TlinT = take(nm * QQ * 4); o.TlogT = take(nm * QQ * 4);
    o.TTp = take(nm * (size_t)d.QP * d.QS * 4); o.alias = take(nm * qq * 4); o.cdfd = take(nm * qq * 8);
    const size_t WW = (size_t)d.QW * d.QW;
    o.Wup = take(nm * WW * 4); o.Wdn = take(nm * WW * 4);
    const size_t PW = (size_t)(d.QW > d.QP ? d.QW : d.QP);
    o.py = take(PW * 4); o.rcp = take((size_t)d.QP * 4); o.rcu = take((size_t)d.QP * 4);
    o.status = take(sizeof(int));
    o.host_bytes = off;                                          // the pinned host image ends here: what follows is derived ON the device
    o.memo = take(ghm_leaf_memo_rows(d) * (size_t)d.QS * 4);
    o.bytes = off;
    return o;
}

// read-only validation pass: nothing is written before every input is known to be a probability, so a failed
// ghm_model_update leaves the host slab (the source of the constant-bank kernel parameters) and the device slab
// consistent with each other
static int validate_tables(const GhmDev& d, const double* T_host, const double* p_y_host) {
    const int q = d.q;
    for (size_t mi = 0; mi < (size_t)d.n_mat; ++mi)
        for (int a = 0; a < q; ++a)
            for (int b = 0; b < q; ++b) {
                const double t = T_host[(mi * q + a) * (size_t)q + b];
                if (!(t >= 0.0) || !isfinite(t))
                    return ghm_fail(GHM_EINVAL, "transition[%zu][%d][%d]=%g is not a probability", mi, a, b, t);
            }
    if (p_y_host)
        for (int a = 0; a < q; ++a)
            if (!(p_y_host[a] >= 0.0) || !isfinite(p_y_host[a]))
                return ghm_fail(GHM_EINVAL, "p_y[%d]=%g is not a probability", a, p_y_host[a]);
    return GHM_OK;
}

// fills the host slab from the caller's float64 matrices (validated by validate_tables)
static int derive_tables(const GhmDev& d, const double* T_host, const double* p_y_host, char* hs, bool validated = false) {
    if (!validated) {
        int vrc = validate_tables(d, T_host, p_y_host);
        if (vrc) return vrc;
    }
    const SlabLayout o = slab_layout(d);
    const int q = d.q, QP = d.QP, QS = d.QS;
    const size_t nm = (size_t)d.n_mat, QQ = (size_t)QP * QP;
    memset(hs, 0, o.host_bytes);
    float* Tlin = (float*)(hs + o.Tlin); float* TlinT = (float*)(hs + o.TlinT); float* TlogT = (float*)(hs + o.TlogT);
    float* Wup = (float*)(hs + o.Wup); float* Wdn = (float*)(hs + o.Wdn);
    float* TTp = (float*)(hs + o.TTp); uint32_t* alias = (uint32_t*)(hs + o.alias); double* cdfd = (double*)(hs + o.cdfd);
    for (size_t i = 0; i < nm * QQ; ++i) TlogT[i] = -INFINITY;
    std::vector<double> scaled(q);
    std::vector<int> small, large;
    for (size_t mi = 0; mi < nm; ++mi) {
        const double* T = T_host + mi * (size_t)q * q;
        for (int a = 0; a < q; ++a) {
            double run = 0.0;
            for (int b = 0; b < q; ++b) {
                double t = T[(size_t)a * q + b];
                if (!(t >= 0.0) || !isfinite(t))
                    return ghm_fail(GHM_EINVAL, "transition[%zu][%d][%d]=%g is not a probability", mi, a, b, t);
                Tlin[mi * QQ + (size_t)a * QP + b] = (float)t;
                TlinT[mi * QQ + (size_t)b * QP + a] = (float)t;
                TlogT[mi * QQ + (size_t)b * QP + a] = (float)log(t);
                TTp[mi * (size_t)QP * QS + (size_t)b * QS + a] = (float)t;
                if (d.QW) {
                    Wup[mi * (size_t)d.QW * d.QW + (size_t)a * d.QW + b] = (float)t;
                    Wdn[mi * (size_t)d.QW * d.QW + (size_t)b * d.QW + a] = (float)t;
                }
                run = (b == 0) ? t : run + t;                 // np.cumsum: sequential f64 adds
                cdfd[mi * (size_t)q * q + (size_t)a * q + b] = run;
            }
            // Walker/Vose alias table of row a (oracle/philox.py::alias_table performs the same f64 steps in
            // the same order, so Philox-mode samples are reproducible bit-for-bit on the CPU)
            small.clear(); large.clear();
            for (int b = 0; b < q; ++b) {
                scaled[b] = T[(size_t)a * q + b] * (double)q;
                (scaled[b] < 1.0 ? small : large).push_back(b);
            }
            uint32_t* arow = &alias[mi * (size_t)q * q + (size_t)a * q];
            for (int b = 0; b < q; ++b) arow[b] = 0xFFFFFF00u | (uint32_t)b;        // prob 1, alias = self
            while (!small.empty() && !large.empty()) {
                const int sm = small.back(); small.pop_back();
                const int lg = large.back(); large.pop_back();
                double thr = floor(scaled[sm] * 16777216.0);
                if (thr < 0.0) thr = 0.0;
                if (thr > 16777215.0) thr = 16777215.0;
                arow[sm] = ((uint32_t)thr << 8) | (uint32_t)lg;
                scaled[lg] = (scaled[lg] + scaled[sm]) - 1.0;
                (scaled[lg] < 1.0 ? small : large).push_back(lg);
            }
        }
    }
    float* py = (float*)(hs + o.py);
    uint32_t* rc_prior = (uint32_t*)(hs + o.rcp); uint32_t* rc_unif = (uint32_t*)(hs + o.rcu);
    for (int a = 0; a < QP; ++a) { rc_prior[a] = 0xFFFFFFFFu; rc_unif[a] = 0xFFFFFFFFu; }
    double run = 0.0, runu = 0.0;
    for (int a = 0; a < q; ++a) {
        double p = p_y_host ? p_y_host[a] : 1.0 / q;
        if (!(p >= 0.0)) return ghm_fail(GHM_EINVAL, "p_y[%d]=%g is not a probability", a, p);
        py[a] = (float)p;
        run += p; runu += 1.0 / q;
        double t1 = floor(run * 4294967296.0), t2 = floor(runu * 4294967296.0);
        rc_prior[a] = t1 >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)t1;
        rc_unif[a] = t2 >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)t2;
    }
    return GHM_OK;
}

// make slab `which` the active one: device table pointers of `d`, host sources of the constant-bank kernel parameters
static void point_tables(ghm_model* m, int which) {
    GhmDev& d = m->d;
    const SlabLayout o = slab_layout(d);
    char* base = (char*)m->slabs[which];
    d.Tlin = (const float*)(base + o.Tlin);
    d.TlinT = (const float*)(base + o.TlinT);
    d.TlogT = (const float*)(base + o.TlogT);
    d.TTp = (const float*)(base + o.TTp);
    d.Wup = (const float*)(base + o.Wup);
    d.Wdn = (const float*)(base + o.Wdn);
    d.alias = (const uint32_t*)(base + o.alias);
    d.cdfd = (const double*)(base + o.cdfd);
    d.py = (const float*)(base + o.py);
    d.root_cdfu_prior = (const uint32_t*)(base + o.rcp);
    d.root_cdfu_unif = (const uint32_t*)(base + o.rcu);
    d.leaf_memo = ghm_leaf_memo_rows(d) ? (const float*)(base + o.memo) : nullptr;
    char* hb = (char*)m->h_slabs[which];
    m->h_TTp = (float*)(hb + o.TTp);
    m->h_Tlin = (float*)(hb + o.Tlin);
    m->h_TlinT = (float*)(hb + o.TlinT);
    m->slab = m->slabs[which]; m->h_slab = m->h_slabs[which]; m->upload_done = m->upload_evs[which];
    m->active = which;
}

extern "C" int ghm_model_create(ghm_model_t** out, int L, int s, int q, int ti, const double* T_host,
                                const double* p_y_host, int device) {
    if (!out || !T_host) return ghm_fail(GHM_EINVAL, "ghm_model_create: null argument");
    if (L < 1 || L > GHM_MAX_LEVELS) return ghm_fail(GHM_EINVAL, "n_layer=%d outside [1,%d]", L, GHM_MAX_LEVELS);
    if (s < 1 || s > 64) return ghm_fail(GHM_EINVAL, "n_child=%d outside [1,64]", s);
    if (q < 2 || q > 256) return ghm_fail(GHM_EINVAL, "variable_type=%d outside [2,256]", q);
    double nl = pow((double)s, (double)L);
    if (nl > 65536.0) return ghm_fail(GHM_EINVAL, "s^L = %.0f leaves is too large (max 65536)", nl);

    ghm_model* m = new ghm_model();
    memset(m, 0, sizeof *m);
    GhmDev& d = m->d;
    d.L = L; d.s = s; d.q = q; d.ti = ti ? 1 : 0;
    d.QP = ghm_pad_q(q) ? ghm_pad_q(q) : (int)align_up(q, 4);
    d.QS = (int)align_up(d.QP, 4);
    d.QW = q > GHM_MAX_Q_REG ? (int)align_up(q, 32) : 0;
    d.spow[0] = 1;
    for (int l = 1; l <= L; ++l) d.spow[l] = d.spow[l - 1] * s;
    for (int l = L + 1; l <= GHM_MAX_LEVELS; ++l) d.spow[l] = 0;
    d.n_leaves = d.spow[L];
    int e = 0;
    for (int l = 1; l <= L; ++l) {
        d.edge_off[l] = e;
        d.mat_off[l] = d.ti ? (l - 1) * s : e;
        e += d.spow[l];
    }
    d.n_edges = e;
    d.n_mat = d.ti ? L * s : e;
    d.s_magic = s >= 2 ? (unsigned)((0x100000000ull + (unsigned)s - 1) / (unsigned)s) : 0u;
    for (int k = 0; k <= GHM_MAX_LEVELS; ++k)
        d.pow_magic[k] = (k >= 1 && k <= L && d.spow[k] >= 2)
                             ? (unsigned)((0x100000000ull + (unsigned)d.spow[k] - 1) / (unsigned)d.spow[k]) : 0u;
    m->device = device;
    const SlabLayout o = slab_layout(d);
    m->slab_bytes = o.bytes;

    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        delete m;
        return ghm_fail(GHM_ECUDA, "no CUDA device available (%s): libghm_b200 has no CPU fallback",
                        ce == cudaSuccess ? "0 devices" : cudaGetErrorString(ce));
    }
    int prev = 0;
    cudaGetDevice(&prev);
#define MC_TRY(expr)                                                                         \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) {                                                             \
            cudaSetDevice(prev);                                                             \
            ghm_model_destroy(m);                                                            \
            return ghm_fail(GHM_ECUDA, "%s failed: %s", #expr, cudaGetErrorString(_e));      \
        }                                                                                    \
    } while (0)
    MC_TRY(cudaSetDevice(device));
    MC_TRY(cudaMallocHost(&m->h_slab, o.host_bytes));
    int rc = derive_tables(d, T_host, p_y_host, (char*)m->h_slab);
    if (rc) { cudaSetDevice(prev); ghm_model_destroy(m); return rc; }
    MC_TRY(cudaMalloc(&m->slab, o.bytes));
    MC_TRY(cudaMemcpy(m->slab, m->h_slab, o.host_bytes, cudaMemcpyHostToDevice));
    MC_TRY(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking));
    MC_TRY(cudaEventCreateWithFlags(&m->upload_done, cudaEventDisableTiming));
    MC_TRY(cudaEventCreateWithFlags(&m->order_ev, cudaEventDisableTiming));
    MC_TRY(cudaSetDevice(prev));
#undef MC_TRY
    m->slabs[0] = m->slab; m->h_slabs[0] = m->h_slab; m->upload_evs[0] = m->upload_done;
    m->slabs[1] = nullptr; m->h_slabs[1] = nullptr; m->upload_evs[1] = nullptr;
    d.status = (int*)((char*)m->slab + o.status);              // the sticky status word stays in slab 0 for the model's lifetime
    point_tables(m, 0);
    {                                                            // device-derived tables of slab 0
        GhmDeviceGuard guard(device);
        rc = ghm_build_leaf_memo(m, m->stream);
        if (!rc && cudaStreamSynchronize(m->stream) != cudaSuccess) rc = ghm_fail(GHM_ECUDA, "leaf memo build failed");
        if (rc) { ghm_model_destroy(m); return rc; }
    }
    rc = ghm_guides_init(m);
    if (rc) { ghm_model_destroy(m); return rc; }
    *out = m;
    return GHM_OK;
}

// New tables for an existing model of the same shape (the p_flip sweeps of figures/eval-*-ood.py build a new
// sampler per p, :76-79): derive on the host, one H2D copy of the slab from pinned memory, enqueued on `stream`.
// Kernels launched AFTER this call (on `stream`, or on a stream that waits for it) see the new tables.  The tables are
// double buffered: kernels launched BEFORE this call keep reading the previous slab and may still be running on other
// streams; they must have finished before the update AFTER this one starts to overwrite that slab (same stream: by
// stream order; other streams: the caller makes `stream` wait for them).  The sticky status word is preserved.
extern "C" int ghm_model_update(ghm_model_t* m, const double* T_host, const double* p_y_host, void* stream) {
    if (!m || !T_host) return ghm_fail(GHM_EINVAL, "ghm_model_update: null argument");
    int prev = 0;
    cudaGetDevice(&prev);
    if (prev != m->device) cudaSetDevice(m->device);
    struct Restore { int p, dev; ~Restore() { if (p != dev) cudaSetDevice(p); } } restore{prev, m->device};
    int vrc = validate_tables(m->d, T_host, p_y_host);           // before anything is rewritten
    if (vrc) return vrc;
    const SlabLayout o = slab_layout(m->d);
    const int next = 1 - m->active;
    if (!m->slabs[next]) {                                       // first update: the second buffer pair
        GHM_CUDA_TRY(cudaMallocHost(&m->h_slabs[next], o.host_bytes));
        GHM_CUDA_TRY(cudaMalloc(&m->slabs[next], o.bytes));
        GHM_CUDA_TRY(cudaEventCreateWithFlags(&m->upload_evs[next], cudaEventDisableTiming));
    }
    GHM_CUDA_TRY(cudaEventSynchronize(m->upload_evs[next]));     // the upload that last used this pinned buffer has consumed it
    int rc = derive_tables(m->d, T_host, p_y_host, (char*)m->h_slabs[next], true);
    if (rc) return rc;
    GHM_CUDA_TRY(cudaMemcpyAsync(m->slabs[next], m->h_slabs[next], o.status, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    point_tables(m, next);
    rc = ghm_build_leaf_memo(m, (cudaStream_t)stream);           // device-derived tables of the new slab, behind its upload
    GHM_CUDA_TRY(cudaEventRecord(m->upload_evs[next], (cudaStream_t)stream));
    return rc;
}

extern "C" int ghm_model_set_gemm_mode(ghm_model_t* m, int mode) {
    if (!m) return ghm_fail(GHM_EINVAL, "ghm_model_set_gemm_mode: null model");
    if (mode != GHM_GEMM_F32 && mode != GHM_GEMM_TF32 && mode != GHM_GEMM_BF16)
        return ghm_fail(GHM_EINVAL, "ghm_model_set_gemm_mode: bad mode %d", mode);
    m->gemm_mode = mode;
    return GHM_OK;
}

extern "C" int64_t ghm_model_table_bytes(const ghm_model_t* m) { return m ? (int64_t)slab_layout(m->d).status : 0; }

extern "C" int ghm_model_destroy(ghm_model_t* m) {
    if (!m) return GHM_OK;
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(m->device);
    if (m->stream) cudaStreamDestroy(m->stream);
    if (!m->slabs[0]) { m->slabs[0] = m->slab; m->h_slabs[0] = m->h_slab; m->upload_evs[0] = m->upload_done; }   // create() failed early
    for (int i = 0; i < 2; ++i) {
        if (m->upload_evs[i]) cudaEventDestroy(m->upload_evs[i]);
        if (m->slabs[i]) cudaFree(m->slabs[i]);
        if (m->h_slabs[i]) cudaFreeHost(m->h_slabs[i]);
    }
    if (m->order_ev) cudaEventDestroy(m->order_ev);
    if (m->guide_tab) cudaFree(m->guide_tab);
    if (m->d_scratch) cudaFree(m->d_scratch);
    if (m->h_scratch) cudaFreeHost(m->h_scratch);
    cudaSetDevice(prev);
    delete m;
    return GHM_OK;
}

extern "C" int ghm_model_info(const ghm_model_t* m, int* L, int* s, int* q, int* ti, int64_t* n_leaves,
                              int64_t* n_edges) {
    if (!m) return ghm_fail(GHM_EINVAL, "ghm_model_info: null model");
    if (L) *L = m->d.L;
    if (s) *s = m->d.s;
    if (q) *q = m->d.q;
    if (ti) *ti = m->d.ti;
    if (n_leaves) *n_leaves = m->d.n_leaves;
    if (n_edges) *n_edges = m->d.n_edges;
    return GHM_OK;
}

extern "C" int ghm_model_status(ghm_model_t* m, void* stream, int* status_out) {
    if (!m || !status_out) return ghm_fail(GHM_EINVAL, "ghm_model_status: null argument");
    GHM_CUDA_TRY(cudaMemcpyAsync(status_out, m->d.status, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    GHM_CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    return GHM_OK;
}

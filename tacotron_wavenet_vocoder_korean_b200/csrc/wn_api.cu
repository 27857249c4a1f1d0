// wn_api.cu -- host side of libwn_b200.so: the C ABI declared in include/wn_b200.h.
//
// Responsibilities: hold the TF-named weights (generate.py:157-161), choose the CTA topology and the
// evaluation plan, pack every matrix into the per-CTA thread-major shared-memory images, own the
// mailboxes / dilation-queue rings, and launch the persistent kernel cooperatively.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <algorithm>
#include <map>
#include <string>
#include <vector>
#include "wn_params.h"

// single translation unit: the device code is compiled together with its launch sites
#include "wn_kernel.cu"
#include "wn_mel.cuh"
#include <cmath>
#include <mutex>

namespace {

constexpr int kManyRows = 10;
constexpr int kSharedRingRows = 16;    // cluster path: from this many rows on the layer kernel uses one shared ring per layer (wn_kernel_v2.cuh, SR); measured break-even
constexpr int kMaxDynSmem = 232448 - 2048;   // 227 KB opt-in limit minus static/reserved slack

std::string g_create_error;

struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
    cudaError_t ensure(size_t need)
    {
        if (need <= bytes) return cudaSuccess;
        if (p) { cudaError_t e = cudaFree(p); p = nullptr; bytes = 0; if (e != cudaSuccess) return e; }
        cudaError_t e = cudaMalloc(&p, need);
        if (e == cudaSuccess) bytes = need;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
};

}  // namespace

namespace {
// One-time, per-process die calibration: which die each SM id belongs to.
struct Calib {
    bool tried = false, ok = false;
    int n_sm = 0, ref_smid = -1;
    std::vector<unsigned char> sm_die;
    std::vector<unsigned> raw;
    void *d_sm_die = nullptr;
    std::string note;
};
Calib g_calib;
constexpr int kCalibSmem = 150 * 1024;     // forces one calibration CTA per SM

// two separated modes?  returns the midpoint between the 5th and 95th percentile
bool bimodal_threshold(std::vector<unsigned> v, unsigned *thr)
{
    if (v.size() < 4) return false;
    std::sort(v.begin(), v.end());
    unsigned lo = v[v.size() / 20], hi = v[v.size() - 1 - v.size() / 20];
    if (lo == 0 || hi < lo + lo / 4) return false;
    *thr = (lo + hi) / 2;
    return true;
}

// runs a job list on a one-CTA-per-SM grid; out must hold the offsets used by the jobs
cudaError_t run_calib_jobs(unsigned long long *grains, const std::vector<WnCalibJob> &jobs, int iters, std::vector<unsigned> &out,
                           int n_sm, bool *aborted)
{
    WnCalibJob *dj = nullptr; unsigned *dout = nullptr; int *dab = nullptr;
    cudaError_t e;
    if ((e = cudaMalloc(&dj, jobs.size() * sizeof(WnCalibJob))) != cudaSuccess) return e;
    if ((e = cudaMalloc(&dout, out.size() * 4)) != cudaSuccess) { cudaFree(dj); return e; }
    if ((e = cudaMalloc(&dab, 4)) != cudaSuccess) { cudaFree(dj); cudaFree(dout); return e; }
    cudaMemcpy(dj, jobs.data(), jobs.size() * sizeof(WnCalibJob), cudaMemcpyHostToDevice);
    cudaMemset(dout, 0, out.size() * 4);
    cudaMemset(dab, 0, 4);
    int n_jobs = (int)jobs.size();
    e = cudaFuncSetAttribute(wn_calib_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kCalibSmem);
    if (e == cudaSuccess) {
        void *args[] = {(void *)&grains, (void *)&dj, (void *)&n_jobs, (void *)&iters, (void *)&dout, (void *)&dab};
        e = cudaLaunchCooperativeKernel((const void *)wn_calib_kernel, dim3(n_sm), dim3(32), args, (size_t)kCalibSmem, 0);
    }
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    int ab = 0;
    if (e == cudaSuccess) {
        cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(&ab, dab, 4, cudaMemcpyDeviceToHost);
    }
    *aborted = ab != 0;
    cudaFree(dj); cudaFree(dout); cudaFree(dab);
    return e;
}

// SM -> die map: SM `ref` ping-pongs with every other SM over G probe grains.  A partner on ref's die shows
// two modes over the grains (fast = grain homed on that die); a partner on the other die is uniformly slow.
void calibrate_once(int n_sm)
{
    Calib &c = g_calib;
    if (c.tried) return;
    c.tried = true;
    c.n_sm = n_sm;
    if (getenv("WN_NO_DIE_AWARE")) { c.note = "disabled by WN_NO_DIE_AWARE"; return; }
    const int G = 16, ref = 0;
    unsigned long long *probe = nullptr;
    if (cudaMalloc(&probe, (size_t)G * 2048) != cudaSuccess) { c.note = "calibration alloc failed"; return; }
    cudaMemset(probe, 0, (size_t)G * 2048);
    std::vector<WnCalibJob> jobs;
    for (int s = 0; s < n_sm; ++s)
        if (s != ref) jobs.push_back(WnCalibJob{ref, s, 0, G, s * G});
    std::vector<unsigned> out((size_t)n_sm * G, 0);
    bool aborted = false;
    cudaError_t e = run_calib_jobs(probe, jobs, 12, out, n_sm, &aborted);
    cudaFree(probe);
    if (e != cudaSuccess || aborted) { c.note = std::string("calibration run failed: ") + (aborted ? "timeout" : cudaGetErrorString(e)); cudaGetLastError(); return; }
    c.raw = out;
    c.sm_die.assign(n_sm, 1);
    c.sm_die[ref] = 0;
    int n0 = 1, n1 = 0;
    for (int s = 0; s < n_sm; ++s) {
        if (s == ref) continue;
        std::vector<unsigned> row(out.begin() + (size_t)s * G, out.begin() + (size_t)(s + 1) * G);
        unsigned mn = *std::min_element(row.begin(), row.end()), mx = *std::max_element(row.begin(), row.end());
        if (mn == 0) { c.note = "no data for SM " + std::to_string(s); return; }
        if (mx > mn + mn / 2) { c.sm_die[s] = 0; ++n0; }     // two modes: same die as ref
        else ++n1;
    }
    if (n0 < 8 || n1 < 8) { c.note = "implausible die split " + std::to_string(n0) + "/" + std::to_string(n1); return; }
    if (cudaMalloc(&c.d_sm_die, n_sm) != cudaSuccess ||
        cudaMemcpy(c.d_sm_die, c.sm_die.data(), n_sm, cudaMemcpyHostToDevice) != cudaSuccess) { c.note = "upload failed"; return; }
    c.ref_smid = ref;
    c.ok = true;
    c.note = "dies " + std::to_string(n0) + "/" + std::to_string(n1) + " SMs";
}

// grain -> die map for a mailbox arena: pairs of SMs on ref's die bounce a word through every grain
bool classify_grains(unsigned long long *arena, size_t n_raw, std::vector<unsigned char> &is_far)
{
    const Calib &c = g_calib;
    std::vector<int> d0;
    for (int s = 0; s < c.n_sm; ++s) if (c.sm_die[s] == 0) d0.push_back(s);
    const int n_pairs = std::min<int>(24, (int)d0.size() / 2);
    std::vector<WnCalibJob> jobs;
    const int per = (int)((n_raw + n_pairs - 1) / n_pairs);
    for (int i = 0; i < n_pairs; ++i) {
        int g0 = i * per, ng = std::min<int>(per, (int)n_raw - g0);
        if (ng > 0) jobs.push_back(WnCalibJob{d0[2 * i], d0[2 * i + 1], g0, ng, g0});
    }
    std::vector<unsigned> out(n_raw, 0);
    bool aborted = false;
    if (run_calib_jobs(arena, jobs, 8, out, c.n_sm, &aborted) != cudaSuccess || aborted) { cudaGetLastError(); return false; }
    unsigned thr;
    if (!bimodal_threshold(out, &thr)) return false;
    is_far.resize(n_raw);
    for (size_t g = 0; g < n_raw; ++g) is_far[g] = out[g] > thr;
    return true;
}
}  // namespace

struct wn_handle {
    wn_config cfg;
    std::map<std::string, std::vector<float>> w;
    bool finalized = false;
    std::string err;
    int out_dim = 0;
    int sm_count = 0;
    wn_plan plan{};
    wn_info info{};
    WnParams base{};            // everything except the per-call fields
    int smem_layer = 0, smem_tail = 0, smem_samp = 0, smem_launch = 0;
    const void *kernel = nullptr;
    const void *kernel_many = nullptr;     // variant used when >= kManyRows rows are in flight (or null)
    // round-2 cluster path (wn_kernel_v2.cuh): layer chain in 8-CTA clusters + tail/sampler kernel on the SMs left over
    bool v2_planned = false, v2 = false;
    const void *kernel_v2_layers = nullptr, *kernel_v2_layers_prof = nullptr, *kernel_v2_tail = nullptr;
    const void *kernel_v2_layers16 = nullptr, *kernel_v2_layers16_prof = nullptr;
    const void *kernel_v2_layers_sr = nullptr, *kernel_v2_layers16_sr = nullptr;      // shared-ring builds (launches with >= kSharedRingRows rows)
    int v2_n16 = 0;                        // > 0: layers [0, 4*n16) run in n16 clusters of 16, the rest in clusters of 8 (second launch)
    cudaStream_t v2_sc = nullptr;
    cudaEvent_t v2_join_c = nullptr;
    bool v2_fast_act = false;
    int v2_grid_layers = 0, v2_smem_layers = 0, v2_smem_tail = 0;
    cudaStream_t v2_sa = nullptr, v2_sb = nullptr;
    cudaEvent_t v2_fork = nullptr, v2_join_a = nullptr, v2_join_b = nullptr;
    std::string v2_note;
    DevBuf layer_img, tail_img, samp_img, gc_table, wc_onehot, upk, mbox, ring, ring_off, status;
    DevBuf prof, mb_tab;
    bool prof_on = false;
    int mb_dual = 0;
    DevBuf up_tmp0, up_tmp1;                                  // upsample intermediates
    DevBuf h_forced, h_lc, h_mel, h_unif, h_out, h_logits;    // wn_generate_host staging
    DevBuf g_lc;
    DevBuf g_noise;                           // cluster path: transformed draw noise of the current launch (wn_noise_prep_kernel)
    DevBuf raw, raw_layers, step_ring_off;                    // TF-layout weights on the device + per-layer pointer table (wn_step)
    std::map<std::string, size_t> raw_off;
    long long step_ring_floats = 0;                                              // wn_generate with mel_dev on a path that materialises the upsampled condition
    size_t mbox_bytes = 0, ring_bytes = 0;
    std::vector<int> up_off;   // float offsets of the upsample kernels inside `upk`
    int64_t launches = 0;
};

namespace {

int fail(wn_handle *h, int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (h) h->err = buf; else g_create_error = buf;
    return code;
}

#define CUDA_TRY(h, expr)                                                                         \
    do {                                                                                          \
        cudaError_t e_ = (expr);                                                                  \
        if (e_ != cudaSuccess) return fail(h, WN_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e_)); \
    } while (0)

int align4(int v) { return (v + 3) & ~3; }
int align32(int v) { return (v + 31) & ~31; }

// evaluation-plan choice: as many k-chunks per column as there are idle threads (keeps the fma chains
// short), power of two, chunk length >= 4.
int choose_t(int ncols, int K, int cap)
{
    int t = 1;
    while (t * 2 <= cap && (long)ncols * (t * 2) <= WN_NT && K % (t * 2) == 0 && K / (t * 2) >= 4) t *= 2;
    return t;
}

WnMat make_mat(int K, int ncols, int t)
{
    WnMat m{};
    m.K = K; m.ncols = ncols; m.t = t; m.ch = K / t;
    m.V = (m.ch % 4 == 0) ? 4 : 1;
    m.gpp = WN_NT / t;
    m.npass = (ncols + m.gpp - 1) / m.gpp;
    if (m.V == 4) m.xstride = ((m.ch / 4) % 2 == 1) ? m.ch : m.ch + 4;
    else m.xstride = (m.ch % 2 == 1) ? m.ch : m.ch + 1;
    m.xlen = align4(t * m.xstride);
    m.in_smem = 1;
    m.u = (m.V == 4) ? wn_u_for_n4(m.ch / 4) : 1;      // mirrors dot_thread() in wn_kernel.cu
    if (t * m.u > 64 && t > 1) return make_mat(K, ncols, t / 2);   // canonical chunk count is capped at 64 per column
    return m;
}

int mat_floats(const WnMat &m) { return m.npass * m.ch * WN_NT; }

// pack W (accessed through get(k, col)) into the thread-major layout described in wn_params.h
template <class Get>
void pack_mat(const WnMat &m, float *dst, Get get)
{
    for (int pass = 0; pass < m.npass; ++pass)
        for (int tid = 0; tid < WN_NT; ++tid) {
            int col = pass * m.gpp + tid / m.t;
            int chunk = tid % m.t;
            for (int i = 0; i < m.ch; ++i) {
                float v = (col < m.ncols) ? get(chunk * m.ch + i, col) : 0.0f;
                size_t idx;
                if (m.V == 4) idx = (((size_t)pass * (m.ch / 4) + i / 4) * WN_NT + tid) * 4 + (i % 4);
                else idx = ((size_t)pass * m.ch + i) * WN_NT + tid;
                dst[idx] = v;
            }
        }
}

const std::vector<float> *find_w(wn_handle *h, const std::string &name)
{
    auto it = h->w.find(name);
    return it == h->w.end() ? nullptr : &it->second;
}

bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

}  // namespace

extern "C" {

const char *wn_last_error(const wn_handle *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int wn_receptive_field(int filter_width, const int32_t *dilations, int n, int scalar_input, int initial_filter_width)
{
    long s = 0;
    for (int i = 0; i < n; ++i) s += dilations[i];
    long rf = (long)(filter_width - 1) * s + 1;
    rf += scalar_input ? initial_filter_width - 1 : filter_width - 1;
    return (int)rf;
}

int wn_create(const wn_config *cfg, wn_handle **out)
{
    if (!cfg || !out) return fail(nullptr, WN_ERR_ARG, "null argument");
    *out = nullptr;
    if (cfg->filter_width != 2) return fail(nullptr, WN_ERR_ARG, "filter_width must be 2 (got %d)", cfg->filter_width);
    if (cfg->batch < 1 || cfg->batch > WN_MAX_BATCH) return fail(nullptr, WN_ERR_ARG, "batch must be in 1..%d", WN_MAX_BATCH);
    if (cfg->n_layers < 1 || cfg->n_layers > WN_MAX_LAYERS) return fail(nullptr, WN_ERR_ARG, "n_layers out of range");
    if (cfg->residual_channels < 1 || cfg->residual_channels > WN_NT) return fail(nullptr, WN_ERR_ARG, "residual_channels must be <= %d", WN_NT);
    if (cfg->dilation_channels < 1 || cfg->dilation_channels > WN_NT || cfg->skip_channels < 1)
        return fail(nullptr, WN_ERR_ARG, "dilation_channels must be in 1..%d, skip_channels >= 1", WN_NT);
    if (cfg->n_upsample < 0 || cfg->n_upsample > WN_MAX_UPSAMPLE) return fail(nullptr, WN_ERR_ARG, "too many upsample stages");
    for (int i = 0; i < cfg->n_layers; ++i)
        if (cfg->dilations[i] < 1) return fail(nullptr, WN_ERR_ARG, "dilation %d < 1", i);
    int out_dim = cfg->scalar_input ? cfg->out_channels : cfg->quantization_channels;
    if (cfg->scalar_input) {
        if (out_dim % 3 != 0 || out_dim / 3 > 32 || out_dim < 3) return fail(nullptr, WN_ERR_ARG, "out_channels must be 3*nr_mix, nr_mix <= 32");
        if (cfg->initial_filter_width < 1 || cfg->initial_filter_width > WN_NT) return fail(nullptr, WN_ERR_ARG, "initial_filter_width out of range");
    } else {
        if (!is_pow2(out_dim) || out_dim < 32 || out_dim > WN_NT) return fail(nullptr, WN_ERR_ARG, "quantization_channels must be a power of two in 32..%d", WN_NT);
    }
    if (cfg->gc_channels && cfg->gc_cardinality < 1) return fail(nullptr, WN_ERR_ARG, "gc_cardinality required with gc_channels");
    if (cfg->lc_channels > WN_NT || cfg->gc_channels > WN_NT) return fail(nullptr, WN_ERR_ARG, "conditioning channels must be <= %d", WN_NT);
    wn_handle *h = new wn_handle();
    h->cfg = *cfg;
    h->out_dim = out_dim;
    *out = h;
    return WN_OK;
}

void wn_destroy(wn_handle *h)
{
    if (!h) return;
    DevBuf *bufs[] = {&h->layer_img, &h->tail_img, &h->samp_img, &h->gc_table, &h->wc_onehot, &h->upk, &h->mbox, &h->ring,
                      &h->ring_off, &h->status, &h->prof, &h->mb_tab, &h->up_tmp0, &h->up_tmp1, &h->h_forced, &h->h_lc, &h->h_mel, &h->h_unif,
                      &h->h_out, &h->h_logits, &h->g_lc, &h->g_noise, &h->raw, &h->raw_layers, &h->step_ring_off};
    for (DevBuf *b : bufs) b->release();
    if (h->v2_sa) cudaStreamDestroy(h->v2_sa);
    if (h->v2_sb) cudaStreamDestroy(h->v2_sb);
    if (h->v2_sc) cudaStreamDestroy(h->v2_sc);
    if (h->v2_join_c) cudaEventDestroy(h->v2_join_c);
    if (h->v2_fork) cudaEventDestroy(h->v2_fork);
    if (h->v2_join_a) cudaEventDestroy(h->v2_join_a);
    if (h->v2_join_b) cudaEventDestroy(h->v2_join_b);
    delete h;
}

int wn_set_weight(wn_handle *h, const char *name, const float *data, int64_t n)
{
    if (!h || !name || !data || n <= 0) return fail(h, WN_ERR_ARG, "wn_set_weight: bad argument");
    h->w[name] = std::vector<float>(data, data + n);
    h->finalized = false;
    return WN_OK;
}

// Topology, evaluation plan and shared-memory layout: a pure function of (config, SM count), no CUDA calls.
static int plan_layout(wn_handle *h, int sm_count)
{
    const wn_config &c = h->cfg;
    const int N = c.batch, L = c.n_layers, R = c.residual_channels, D = c.dilation_channels, S = c.skip_channels;
    const int G = c.gc_channels, C = c.lc_channels, O = h->out_dim, Q = c.quantization_channels, ifw = c.initial_filter_width;
    h->info = wn_info{};
    h->sm_count = sm_count;

    // ---- topology --------------------------------------------------------------------------------
    int M = c.force_M ? c.force_M : std::min(4, std::max(1, D / 32));
    int Mt = c.force_Mt ? c.force_Mt : std::min(16, std::max(1, S / 32));
    if (!is_pow2(M) || M > 4 || D % M) return fail(h, WN_ERR_ARG, "layer split M=%d invalid for D=%d (power of two <= 4 dividing D)", M, D);
    if (!is_pow2(Mt) || Mt > 32 || S % Mt) return fail(h, WN_ERR_ARG, "tail split Mt=%d invalid for S=%d", Mt, S);
    if (!c.force_M) while (M > 1 && L * M + Mt + 1 > sm_count) M >>= 1;
    if (!c.force_Mt) while (Mt > 1 && L * M + Mt + 1 > sm_count) Mt >>= 1;
    const int grid = L * M + Mt + 1;
    if (grid > sm_count) return fail(h, WN_ERR_ARG, "%d layers x M=%d + %d tail + 1 sampler = %d CTAs exceed the %d SMs", L, M, Mt, grid, sm_count);
    const int Dm = D / M, Sm = S / M, St = S / Mt;
    if (2 * Dm > WN_NT) return fail(h, WN_ERR_ARG, "2*D/M = %d filter/gate columns exceed %d threads", 2 * Dm, WN_NT);
    if (St > WN_NT) return fail(h, WN_ERR_ARG, "S/Mt = %d exceeds %d threads", St, WN_NT);

    WnParams &p = h->base;
    memset(&p, 0, sizeof p);
    p.N = N; p.L = L; p.R = R; p.D = D; p.S = S; p.O = O; p.Q = Q; p.G = G; p.C = C; p.ifw = ifw;
    p.scalar_input = c.scalar_input; p.nr_mix = c.scalar_input ? O / 3 : 0;
    p.M = M; p.Mt = Mt; p.Dm = Dm; p.Sm = Sm; p.St = St; p.grid = grid;
    for (int i = 0; i < L; ++i) p.dil[i] = c.dilations[i];

    // ---- evaluation plan ----------------------------------------------------------------------------
    const int t_fg = choose_t(2 * Dm, R, 16);            // <= 16 so filter/gate partner columns share a warp
    p.cur = make_mat(R, 2 * Dm, t_fg);
    p.old = make_mat(R, 2 * Dm, t_fg);
    p.lc = make_mat(C ? C : 4, 2 * Dm, C ? choose_t(2 * Dm, C, 32) : 1);
    p.gc = make_mat(G ? G : 4, 2 * Dm, G ? choose_t(2 * Dm, G, 32) : 1);
    p.dense = make_mat(Dm, R, choose_t(R, Dm, 32));
    p.skip = make_mat(D, Sm, choose_t(Sm, D, 32));
    p.post1 = make_mat(S, St, choose_t(St, S, 32));
    p.post2 = make_mat(St, O, choose_t(O, St, 32));
    p.causal = make_mat(c.scalar_input ? ifw : 4, R, c.scalar_input ? choose_t(R, ifw, 32) : 1);
    if (p.cur.npass != 1 || p.old.npass != 1) return fail(h, WN_ERR_ARG, "internal: filter/gate must be single pass");
    h->plan = wn_plan{M, Mt, p.cur.t * p.cur.u, p.old.t * p.old.u, p.lc.t * p.lc.u, p.gc.t * p.gc.u, p.dense.t * p.dense.u,
                      p.skip.t * p.skip.u, p.post1.t * p.post1.u, p.post2.t * p.post2.u, p.causal.t * p.causal.u};

    // ---- layer image = [bfg | bd | bs | cur | dense | skip | old | lc | gc]; the resident prefix is what fits
    {
        int off = 0;
        p.off_bfg = off; off += align4(2 * Dm);
        p.off_bd = off; off += align4(R);
        p.off_bs = off; off += align4(Sm);
        WnMat *order[6] = {&p.cur, &p.dense, &p.skip, &p.old, &p.lc, &p.gc};
        bool present[6] = {true, true, true, true, C > 0, G > 0};
        int so = 0;
        p.ls.xs_cur = so; so += p.cur.xlen;
        p.ls.xs_old = so; so += p.old.xlen;
        p.ls.lcs = so; so += p.lc.xlen;
        p.ls.xraw = so; so += align4(R);
        p.ls.zs_dense = so; so += p.dense.xlen;
        p.ls.zs_skip = so; so += p.skip.xlen;
        p.ls.gvec = so; so += p.gc.xlen;
        p.ls.bfgN = so; so += align4(N * 2 * Dm);
        p.ls.pre = so; so += align4(N * 2 * Dm);
        p.ls.total_floats = so;
        so += align4(N * (R + Dm));            // rowbuf of the warp-specialised layer role (follows total_floats)
        const int cap = kMaxDynSmem / 4 - so;
        int resident = off;
        bool spill = false;
        for (int i = 0; i < 6; ++i) {
            WnMat &m = *order[i];
            m.off = off;
            int n = present[i] ? mat_floats(m) : 0;
            if (!present[i]) { m.in_smem = 1; continue; }
            if (!spill && off + n <= cap) { m.in_smem = 1; resident = off + n; }
            else { m.in_smem = 0; spill = true; h->info.weights_in_global += (int64_t)n * L * M; }
            off += n;
        }
        if (resident > cap) return fail(h, WN_ERR_ARG, "per-CTA vectors do not fit shared memory");
        p.layer_img_floats = align32(off);
        p.layer_smem_floats = align4(resident);
        h->smem_layer = (p.layer_smem_floats + so) * 4;
    }
    // ---- tail image = [b1 | post1 | post2]
    {
        int off = 0;
        p.off_b1 = off; off += align4(St);
        int so = 0;
        p.ts.as1 = so; so += p.post1.xlen;
        p.ts.c1s = so; so += p.post2.xlen;
        p.ts.total_floats = so;
        const int cap = kMaxDynSmem / 4 - so;
        p.post1.off = off; off += mat_floats(p.post1);
        p.post2.off = off; off += mat_floats(p.post2);
        int resident = off;
        if (off > cap) {   // spill conv2 first, then conv1
            p.post2.in_smem = 0; resident = p.post2.off;
            h->info.weights_in_global += (int64_t)mat_floats(p.post2) * Mt;
            if (resident > cap) { p.post1.in_smem = 0; resident = p.post1.off; h->info.weights_in_global += (int64_t)mat_floats(p.post1) * Mt; }
        }
        p.tail_img_floats = align32(off);
        p.tail_smem_floats = align4(resident);
        h->smem_tail = (p.tail_smem_floats + so) * 4;
    }
    // ---- sampler image = [b2 | causal]
    {
        int off = 0;
        p.off_b2 = off; off += align4(O);
        p.causal.off = off;
        if (c.scalar_input) off += mat_floats(p.causal);
        int so = 0;
        p.ss.c2s = so; so += align4(std::max(O, WN_NT));
        p.ss.cq = so; so += align4(N * std::max(ifw, 1));
        p.ss.cqx = so; so += p.causal.xlen;
        p.ss.ids = so; so += align4(2 * N);
        p.ss.qs = so; so += WN_NT;
        so = (so + 1) & ~1;
        p.ss.cdf = so; so += 2 * WN_NT;
        p.ss.red = so; so += 2 * 16;
        p.ss.misc = so; so += 32;
        p.ss.total_floats = so;
        if ((off + so) * 4 > kMaxDynSmem) return fail(h, WN_ERR_ARG, "sampler image does not fit shared memory");
        p.samp_img_floats = align32(off);
        p.samp_smem_floats = align4(off);
        h->smem_samp = (p.samp_smem_floats + so) * 4;
    }
    h->smem_launch = std::max(h->smem_layer, std::max(h->smem_tail, h->smem_samp));

    wn_info &inf = h->info;
    inf.grid = grid; inf.threads = WN_NT; inf.M = M; inf.Mt = Mt; inf.sm_count = sm_count;
    inf.smem_bytes_layer = h->smem_layer; inf.smem_bytes_tail = h->smem_tail; inf.smem_bytes_sampler = h->smem_samp;
    int64_t per_layer = 2LL * 2 * R * D + 2LL * D + 2LL * C * D + (int64_t)D * R + R + (int64_t)D * S + S;
    int64_t causal = c.scalar_input ? (int64_t)ifw * R : 2LL * R;
    inf.p_hot = per_layer * L + causal + (int64_t)S * S + S + (int64_t)S * O + O;
    inf.weights_in_smem = ((int64_t)p.layer_smem_floats * L * M + (int64_t)p.tail_smem_floats * Mt + p.samp_smem_floats);
    // compile-time specialised instantiation, when the planned shapes coincide with one
    h->kernel = (const void *)wn_persistent_kernel;
    h->kernel_many = nullptr;
    h->v2_planned = false;
    if (!(c.flags & WN_FLAG_GENERIC_KERNEL) && inf.weights_in_global == 0) {
        if (shape_matches<ShapeCfg2>(p)) {
            h->kernel = (const void *)wn_persistent_kernel_s<ShapeCfg2>;
            h->kernel_many = (const void *)wn_persistent_kernel_s<ShapeCfg2WS>;
            inf.static_shape = 1;
            // cluster path: same plan, same mailboxes; needs ceil(L/2) clusters of 8 CTAs + Mt + 1 further SMs
            h->v2_planned = !(c.flags & WN_FLAG_NO_CLUSTER) && !getenv("WN_NO_CLUSTER") &&
                            V2L<ShapeCfg2>::total_floats(N) * 4 <= kMaxDynSmem && ((L + 1) / 2) * V2_CS + Mt <= sm_count;
            h->v2_grid_layers = ((L + 1) / 2) * V2_CS;
            h->v2_smem_layers = V2L<ShapeCfg2>::total_floats(N) * 4;
            h->v2_smem_tail = (p.tail_smem_floats + 32 * 20 + 32) * 4;
            h->v2_fast_act = (c.flags & WN_FLAG_FAST_ACT) != 0;
        }
        else if (shape_matches<ShapeCfg1>(p)) { h->kernel = (const void *)wn_persistent_kernel_s<ShapeCfg1>; inf.static_shape = 2; }
        else if (shape_matches<ShapeHparams>(p)) { h->kernel = (const void *)wn_persistent_kernel_s<ShapeHparams>; inf.static_shape = 3; }
    }
    return WN_OK;
}

int wn_plan_config(const wn_config *cfg, int sm_count, wn_plan *plan, wn_info *info)
{
    wn_handle *h = nullptr;
    int rc = wn_create(cfg, &h);
    if (rc) return rc;
    rc = plan_layout(h, sm_count > 0 ? sm_count : 148);
    if (rc) g_create_error = h->err;
    if (!rc && plan) *plan = h->plan;
    if (!rc && info) *info = h->info;
    wn_destroy(h);
    return rc;
}

int wn_finalize(wn_handle *h)
{
    if (!h) return WN_ERR_ARG;
    const wn_config &c = h->cfg;
    const int N = c.batch, L = c.n_layers, R = c.residual_channels, D = c.dilation_channels, S = c.skip_channels;
    const int G = c.gc_channels, C = c.lc_channels, O = h->out_dim, Q = c.quantization_channels, ifw = c.initial_filter_width;

    h->raw.release();                          // wn_step's TF-layout copy follows the weights
    int dev = 0;
    cudaDeviceProp prop;
    CUDA_TRY(h, cudaGetDevice(&dev));
    CUDA_TRY(h, cudaGetDeviceProperties(&prop, dev));
    if (prop.major < 10) return fail(h, WN_ERR_CUDA, "device %s is sm_%d%d; this library contains sm_100a code only", prop.name, prop.major, prop.minor);
    if (!prop.cooperativeLaunch) return fail(h, WN_ERR_CUDA, "device does not support cooperative launch");
    {
        int rc = plan_layout(h, prop.multiProcessorCount);
        if (rc) return rc;
    }
    WnParams &p = h->base;
    const int M = p.M, Mt = p.Mt, Dm = p.Dm, Sm = p.Sm, St = p.St, grid = p.grid;

    // ---- required weights, in TF layout --------------------------------------------------------
    auto need = [&](const std::string &nm, size_t n, const std::vector<float> **out) -> int {
        const std::vector<float> *v = find_w(h, nm);
        if (!v) return fail(h, WN_ERR_STATE, "missing weight %s", nm.c_str());
        if (v->size() != n) return fail(h, WN_ERR_ARG, "weight %s has %zu elements, expected %zu", nm.c_str(), v->size(), n);
        *out = v;
        return WN_OK;
    };
    std::vector<float> zeros((size_t)std::max(std::max(S, D), std::max(R, O)) + 8, 0.0f);
    auto opt = [&](const std::string &nm, size_t n, const float **out) -> int {
        const std::vector<float> *v = find_w(h, nm);
        if (!v) { *out = zeros.data(); return WN_OK; }
        if (v->size() != n) return fail(h, WN_ERR_ARG, "weight %s has %zu elements, expected %zu", nm.c_str(), v->size(), n);
        *out = v->data();
        return WN_OK;
    };

    // ---- layer images -------------------------------------------------------------------------------
    std::vector<float> limg((size_t)p.layer_img_floats * L * M, 0.0f);
    for (int l = 0; l < L; ++l) {
        std::string pre = "wavenet/dilated_stack/layer" + std::to_string(l) + "/dilation_layer/";
        const std::vector<float> *wf, *wg, *wd, *ws, *gcf = nullptr, *gcg = nullptr, *lcf = nullptr, *lcg = nullptr;
        int rc;
        if ((rc = need(pre + "conv_filter/kernel", (size_t)2 * R * D, &wf))) return rc;
        if ((rc = need(pre + "conv_gate/kernel", (size_t)2 * R * D, &wg))) return rc;
        if ((rc = need(pre + "dense/kernel", (size_t)D * R, &wd))) return rc;
        if ((rc = need(pre + "skip/kernel", (size_t)D * S, &ws))) return rc;
        if (G) { if ((rc = need(pre + "gc_filter/kernel", (size_t)G * D, &gcf))) return rc; if ((rc = need(pre + "gc_gate/kernel", (size_t)G * D, &gcg))) return rc; }
        if (C) { if ((rc = need(pre + "lc_filter/kernel", (size_t)C * D, &lcf))) return rc; if ((rc = need(pre + "lc_gate/kernel", (size_t)C * D, &lcg))) return rc; }
        const float *bf, *bg, *bd, *bs;
        if ((rc = opt(pre + "conv_filter/bias", D, &bf))) return rc;
        if ((rc = opt(pre + "conv_gate/bias", D, &bg))) return rc;
        if ((rc = opt(pre + "dense/bias", R, &bd))) return rc;
        if ((rc = opt(pre + "skip/bias", S, &bs))) return rc;
        for (int m = 0; m < M; ++m) {
            float *img = limg.data() + (size_t)(l * M + m) * p.layer_img_floats;
            for (int j = 0; j < Dm; ++j) { img[p.off_bfg + 2 * j] = bf[m * Dm + j]; img[p.off_bfg + 2 * j + 1] = bg[m * Dm + j]; }
            for (int r = 0; r < R; ++r) img[p.off_bd + r] = bd[r];
            for (int s = 0; s < Sm; ++s) img[p.off_bs + s] = bs[m * Sm + s];
            // column c = 2*j + (0: filter, 1: gate) for gated channel m*Dm + j
            auto fgcol = [&](const std::vector<float> *f, const std::vector<float> *g, size_t base, int k, int col) {
                const std::vector<float> *src = (col & 1) ? g : f;
                return (*src)[base + (size_t)k * D + (m * Dm + (col >> 1))];
            };
            pack_mat(p.cur, img + p.cur.off, [&](int k, int col) { return fgcol(wf, wg, (size_t)R * D, k, col); });   // tap 1 = current
            pack_mat(p.old, img + p.old.off, [&](int k, int col) { return fgcol(wf, wg, 0, k, col); });               // tap 0 = dilated
            if (C) pack_mat(p.lc, img + p.lc.off, [&](int k, int col) { return fgcol(lcf, lcg, 0, k, col); });
            if (G) pack_mat(p.gc, img + p.gc.off, [&](int k, int col) { return fgcol(gcf, gcg, 0, k, col); });
            pack_mat(p.dense, img + p.dense.off, [&](int k, int r) { return (*wd)[(size_t)(m * Dm + k) * R + r]; });
            pack_mat(p.skip, img + p.skip.off, [&](int k, int s) { return (*ws)[(size_t)k * S + (m * Sm + s)]; });
        }
    }

    // ---- tail images -----------------------------------------------------------------------------------
    std::vector<float> timg((size_t)p.tail_img_floats * Mt, 0.0f);
    const float *b2v;
    {
        const std::vector<float> *w1, *w2;
        int rc;
        if ((rc = need("wavenet/conv1d_1/kernel", (size_t)S * S, &w1))) return rc;
        if ((rc = need("wavenet/conv1d_2/kernel", (size_t)S * O, &w2))) return rc;
        const float *b1;
        if ((rc = opt("wavenet/conv1d_1/bias", S, &b1))) return rc;
        if ((rc = opt("wavenet/conv1d_2/bias", O, &b2v))) return rc;
        for (int mt = 0; mt < Mt; ++mt) {
            float *img = timg.data() + (size_t)mt * p.tail_img_floats;
            for (int s = 0; s < St; ++s) img[p.off_b1 + s] = b1[mt * St + s];
            pack_mat(p.post1, img + p.post1.off, [&](int k, int col) { return (*w1)[(size_t)k * S + (mt * St + col)]; });
            pack_mat(p.post2, img + p.post2.off, [&](int k, int o) { return (*w2)[(size_t)(mt * St + k) * O + o]; });
        }
    }

    // ---- sampler image --------------------------------------------------------------------------------
    std::vector<float> simg((size_t)p.samp_img_floats, 0.0f);
    {
        for (int o = 0; o < O; ++o) simg[p.off_b2 + o] = b2v[o];
        const std::vector<float> *wc;
        int rc;
        if (c.scalar_input) {
            if ((rc = need("wavenet/conv1d/kernel", (size_t)ifw * R, &wc))) return rc;
            pack_mat(p.causal, simg.data() + p.causal.off, [&](int k, int r) { return (*wc)[(size_t)k * R + r]; });
        } else {
            if ((rc = need("wavenet/conv1d/kernel", (size_t)2 * Q * R, &wc))) return rc;
            CUDA_TRY(h, h->wc_onehot.ensure(wc->size() * 4));
            CUDA_TRY(h, cudaMemcpy(h->wc_onehot.p, wc->data(), wc->size() * 4, cudaMemcpyHostToDevice));
        }
    }

    // ---- conditioning tables / upsample kernels --------------------------------------------------------
    if (G) {
        const std::vector<float> *gt;
        int rc;
        if ((rc = need("wavenet/gc_embedding", (size_t)c.gc_cardinality * G, &gt))) return rc;
        CUDA_TRY(h, h->gc_table.ensure(gt->size() * 4));
        CUDA_TRY(h, cudaMemcpy(h->gc_table.p, gt->data(), gt->size() * 4, cudaMemcpyHostToDevice));
    }
    h->up_off.clear();
    if (C && c.n_upsample) {
        std::vector<float> all;
        for (int i = 0; i < c.n_upsample; ++i) {
            const std::vector<float> *k;
            int rc;
            if ((rc = need("wavenet/upsample" + std::to_string(i) + "/kernel", (size_t)c.upsample_factor[i] * 2, &k))) return rc;
            h->up_off.push_back((int)all.size());
            all.insert(all.end(), k->begin(), k->end());
        }
        CUDA_TRY(h, h->upk.ensure(all.size() * 4));
        CUDA_TRY(h, cudaMemcpy(h->upk.p, all.data(), all.size() * 4, cudaMemcpyHostToDevice));
    }

    // ---- upload images, allocate mailboxes and rings ---------------------------------------------------
    CUDA_TRY(h, h->layer_img.ensure(limg.size() * 4));
    CUDA_TRY(h, cudaMemcpy(h->layer_img.p, limg.data(), limg.size() * 4, cudaMemcpyHostToDevice));
    CUDA_TRY(h, h->tail_img.ensure(timg.size() * 4));
    CUDA_TRY(h, cudaMemcpy(h->tail_img.p, timg.data(), timg.size() * 4, cudaMemcpyHostToDevice));
    CUDA_TRY(h, h->samp_img.ensure(simg.size() * 4));
    CUDA_TRY(h, cudaMemcpy(h->samp_img.p, simg.data(), simg.size() * 4, cudaMemcpyHostToDevice));

    // ---- mailboxes: logical word space -> dual-homed 2 KB grains ----------------------------------------
    const size_t n_x = (size_t)N * L * M * R, n_z = (size_t)N * L * M * Dm, n_acc = (size_t)N * L * M * Sm, n_c2 = (size_t)N * Mt * O;
    const size_t n_lg = (n_x + n_z + n_acc + n_c2 + 255) / 256 + 3;            // logical grains (+ slack: readers may resolve the two grains after a word)
    calibrate_once(prop.multiProcessorCount);
    const bool want_dual = g_calib.ok && !(c.flags & WN_FLAG_NO_DIE_AWARE);
    const size_t n_raw = want_dual ? (n_lg * 5) / 2 + 64 : n_lg;
    h->mbox_bytes = n_raw * 2048;
    CUDA_TRY(h, h->mbox.ensure(h->mbox_bytes));
    CUDA_TRY(h, cudaMemset(h->mbox.p, 0, h->mbox_bytes));
    std::vector<unsigned long long> tabs(2 * n_lg);
    unsigned long long base = (unsigned long long)(uintptr_t)h->mbox.p;
    h->mb_dual = 0;
    for (size_t g = 0; g < n_lg; ++g) tabs[g] = tabs[n_lg + g] = base + g * 2048;
    h->info.die_aware = 0;
    if (want_dual) {
        std::vector<unsigned char> is_far;
        if (classify_grains((unsigned long long *)h->mbox.p, n_raw, is_far)) {
            std::vector<size_t> near_g, far_g;
            for (size_t g = 0; g < n_raw; ++g) (is_far[g] ? far_g : near_g).push_back(g);
            if (near_g.size() >= n_lg && far_g.size() >= n_lg) {
                // the classifying SM pairs are on die 0: near grains are homed on die 0
                for (size_t g = 0; g < n_lg; ++g) { tabs[g] = base + near_g[g] * 2048; tabs[n_lg + g] = base + far_g[g] * 2048; }
                h->mb_dual = 1;
                h->info.die_aware = 1;
            }
        }
        CUDA_TRY(h, cudaMemset(h->mbox.p, 0, h->mbox_bytes));
    }
    CUDA_TRY(h, h->mb_tab.ensure(tabs.size() * 8));
    CUDA_TRY(h, cudaMemcpy(h->mb_tab.p, tabs.data(), tabs.size() * 8, cudaMemcpyHostToDevice));
    std::vector<long long> roff(L);
    size_t rtot = 0;
    for (int l = 0; l < L; ++l) { roff[l] = (long long)rtot; rtot += (size_t)M * N * c.dilations[l] * R; }
    h->ring_bytes = rtot * 4;
    CUDA_TRY(h, h->ring.ensure(std::max<size_t>(h->ring_bytes, 16)));
    CUDA_TRY(h, h->ring_off.ensure(L * sizeof(long long)));
    CUDA_TRY(h, cudaMemcpy(h->ring_off.p, roff.data(), L * sizeof(long long), cudaMemcpyHostToDevice));
    CUDA_TRY(h, h->status.ensure(32));

    p.layer_img = (const float *)h->layer_img.p;
    p.tail_img = (const float *)h->tail_img.p;
    p.samp_img = (const float *)h->samp_img.p;
    p.gc_table = (const float *)h->gc_table.p;
    p.wc_onehot = (const float *)h->wc_onehot.p;
    p.mb_x = reinterpret_cast<unsigned long long *>((uintptr_t)0);                    // logical addresses
    p.mb_z = reinterpret_cast<unsigned long long *>((uintptr_t)(n_x * 8));
    p.mb_acc = reinterpret_cast<unsigned long long *>((uintptr_t)((n_x + n_z) * 8));
    p.mb_c2 = reinterpret_cast<unsigned long long *>((uintptr_t)((n_x + n_z + n_acc) * 8));
    p.mb_tab[0] = (unsigned long long *const *)h->mb_tab.p;
    p.mb_tab[1] = (unsigned long long *const *)h->mb_tab.p + n_lg;
    p.sm_die = h->mb_dual ? (const unsigned char *)g_calib.d_sm_die : nullptr;
    p.mb_dual = h->mb_dual;
    p.ring = (float *)h->ring.p;
    p.ring_off = (const long long *)h->ring_off.p;
    p.status = (int32_t *)h->status.p;

    CUDA_TRY(h, cudaFuncSetAttribute(h->kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem_launch));
    if (h->kernel_many) CUDA_TRY(h, cudaFuncSetAttribute(h->kernel_many, cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem_launch));
    int occ = 0;
    if (h->info.static_shape == 1) CUDA_TRY(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, wn_persistent_kernel_s<ShapeCfg2>, WN_NT, h->smem_launch));
    else if (h->info.static_shape == 2) CUDA_TRY(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, wn_persistent_kernel_s<ShapeCfg1>, WN_NT, h->smem_launch));
    else if (h->info.static_shape == 3) CUDA_TRY(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, wn_persistent_kernel_s<ShapeHparams>, WN_NT, h->smem_launch));
    else CUDA_TRY(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, wn_persistent_kernel, WN_NT, h->smem_launch));
    if ((long)occ * h->sm_count < grid) return fail(h, WN_ERR_CUDA, "cannot co-schedule %d CTAs (occupancy %d x %d SMs)", grid, occ, h->sm_count);
    h->v2 = false;
    h->info.cluster_path = 0;
    h->info.fast_act = 0;
    if (h->v2_planned) {
        const bool fa = h->v2_fast_act;
        h->kernel_v2_layers = fa ? (const void *)wn_layers_kernel_v2<ShapeCfg2, true, false, 8> : (const void *)wn_layers_kernel_v2<ShapeCfg2, false, false, 8>;
        h->kernel_v2_layers_prof = fa ? (const void *)wn_layers_kernel_v2<ShapeCfg2, true, true, 8> : (const void *)wn_layers_kernel_v2<ShapeCfg2, false, true, 8>;
        h->kernel_v2_layers16 = fa ? (const void *)wn_layers_kernel_v2<ShapeCfg2, true, false, 16> : (const void *)wn_layers_kernel_v2<ShapeCfg2, false, false, 16>;
        h->kernel_v2_layers16_prof = fa ? (const void *)wn_layers_kernel_v2<ShapeCfg2, true, true, 16> : (const void *)wn_layers_kernel_v2<ShapeCfg2, false, true, 16>;
        h->kernel_v2_layers_sr = fa ? (const void *)wn_layers_kernel_v2<ShapeCfg2, true, false, 8, true> : (const void *)wn_layers_kernel_v2<ShapeCfg2, false, false, 8, true>;
        h->kernel_v2_layers16_sr = fa ? (const void *)wn_layers_kernel_v2<ShapeCfg2, true, false, 16, true> : (const void *)wn_layers_kernel_v2<ShapeCfg2, false, false, 16, true>;
        h->kernel_v2_tail = (const void *)wn_tail_kernel_v2<ShapeCfg2>;
        cudaError_t e = cudaSuccess;
        const void *ks[6] = {h->kernel_v2_layers, h->kernel_v2_layers_prof, h->kernel_v2_layers_sr, h->kernel_v2_layers16, h->kernel_v2_layers16_prof,
                             h->kernel_v2_layers16_sr};
        for (int i = 0; i < 6 && e == cudaSuccess; ++i) {
            e = cudaFuncSetAttribute(ks[i], cudaFuncAttributeMaxDynamicSharedMemorySize, h->v2_smem_layers);
            if (e == cudaSuccess && i >= 3) e = cudaFuncSetAttribute(ks[i], cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        }
        if (e == cudaSuccess) e = cudaFuncSetAttribute(h->kernel_v2_tail, cudaFuncAttributeMaxDynamicSharedMemorySize, h->v2_smem_tail);
        auto max_clusters = [&](const void *k, int cs, int *n) -> cudaError_t {
            cudaLaunchConfig_t lc = {};
            lc.gridDim = dim3(cs * 32);
            lc.blockDim = dim3(V2_NT);
            lc.dynamicSmemBytes = (size_t)h->v2_smem_layers;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            lc.attrs = at; lc.numAttrs = 1;
            return cudaOccupancyMaxActiveClusters(n, k, &lc);
        };
        int n8 = 0, n16 = 0;
        if (e == cudaSuccess) e = max_clusters(h->kernel_v2_layers, 8, &n8);
        if (e == cudaSuccess && max_clusters(h->kernel_v2_layers16, 16, &n16) != cudaSuccess) { n16 = 0; cudaGetLastError(); }
        if (e != cudaSuccess) { h->v2_note = std::string("cluster path disabled: ") + cudaGetErrorString(e); cudaGetLastError(); }
        else if (n8 * V2_CS < h->v2_grid_layers) h->v2_note = "cluster path disabled: only " + std::to_string(n8) + " clusters of 8 are co-resident";
        else {
            int lo = 0, hi = 0;
            cudaDeviceGetStreamPriorityRange(&lo, &hi);          // lo = least priority (numerically greatest)
            if (!h->v2_sa) e = cudaStreamCreateWithPriority(&h->v2_sa, cudaStreamNonBlocking, hi);
            if (e == cudaSuccess && !h->v2_sc) e = cudaStreamCreateWithPriority(&h->v2_sc, cudaStreamNonBlocking, hi);
            if (e == cudaSuccess && !h->v2_sb) e = cudaStreamCreateWithPriority(&h->v2_sb, cudaStreamNonBlocking, lo);
            if (e == cudaSuccess && !h->v2_fork) e = cudaEventCreateWithFlags(&h->v2_fork, cudaEventDisableTiming);
            if (e == cudaSuccess && !h->v2_join_a) e = cudaEventCreateWithFlags(&h->v2_join_a, cudaEventDisableTiming);
            if (e == cudaSuccess && !h->v2_join_b) e = cudaEventCreateWithFlags(&h->v2_join_b, cudaEventDisableTiming);
            if (e == cudaSuccess && !h->v2_join_c) e = cudaEventCreateWithFlags(&h->v2_join_c, cudaEventDisableTiming);
            if (e != cudaSuccess) { h->v2_note = std::string("cluster path disabled: ") + cudaGetErrorString(e); cudaGetLastError(); }
            else {
                h->v2 = true; h->info.cluster_path = 1; h->info.fast_act = fa ? 1 : 0;
                // The dilation rings (scratch that never has to reach DRAM) are pinned in L2 for the generation streams: a slot is
                // rewritten every d steps, and without the hint the write-back of its dirty lines is most of the launch's DRAM writes.
                if (h->ring_bytes && !getenv("WN_NO_L2_PERSIST")) {
                    int maxwin = 0, l2size = 0;
                    cudaDeviceGetAttribute(&maxwin, cudaDevAttrMaxAccessPolicyWindowSize, dev);
                    cudaDeviceGetAttribute(&l2size, cudaDevAttrL2CacheSize, dev);
                    size_t want = std::min<size_t>(h->ring_bytes, (size_t)std::max(maxwin, 0));
                    size_t lim = std::min<size_t>(want, (size_t)l2size / 4 * 3);
                    if (lim > 0 && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, lim) == cudaSuccess) {
                        cudaStreamAttrValue av;
                        memset(&av, 0, sizeof av);
                        av.accessPolicyWindow.base_ptr = h->ring.p;
                        av.accessPolicyWindow.num_bytes = want;
                        av.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)lim / (double)want);
                        av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
                        av.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
                        cudaStreamSetAttribute(h->v2_sa, cudaStreamAttributeAccessPolicyWindow, &av);
                        cudaStreamSetAttribute(h->v2_sc, cudaStreamAttributeAccessPolicyWindow, &av);
                    }
                    cudaGetLastError();
                }
                h->v2_n16 = 0;
                h->v2_note = "cluster path: " + std::to_string(h->v2_grid_layers / V2_CS) + " clusters of 8 + " + std::to_string(Mt) + " tail CTAs";
                // 16-CTA clusters (4 layers each) halve the L2 hops of the chain; at most n16 of them are co-resident, the
                // remaining layers run in 8-CTA clusters of a second launch.  Whether the three kernels really share the GPU
                // depends on the part's GPC layout, so the shape is tried once on a two-step dummy job.
                const int want16 = L / 4, rem = L - 4 * (L / 4);
                if (!getenv("WN_NO_CLUSTER16") && want16 >= 1 && n16 >= want16 && want16 * 16 + ((rem + 1) / 2) * 8 + Mt <= h->sm_count) {
                    h->v2_n16 = want16;
                    h->finalized = true;
                    DevBuf dummy;
                    const size_t nu = (size_t)2 * (O / 3 + 1);
                    bool ok = dummy.ensure((nu + 8) * 4) == cudaSuccess && cudaMemset(dummy.p, 0, (nu + 8) * 4) == cudaSuccess;
                    if (ok) {
                        std::vector<float> hu(nu, 0.5f);
                        ok = cudaMemcpy((float *)dummy.p + 8, hu.data(), nu * 4, cudaMemcpyHostToDevice) == cudaSuccess;
                    }
                    if (ok) {
                        wn_generate_args ga;
                        memset(&ga, 0, sizeof ga);
                        int32_t gid = 0;
                        ga.rows = 1; ga.T = 2; ga.n_forced = 1; ga.forced_dev = (const float *)dummy.p;
                        ga.gc_ids = c.gc_channels ? &gid : nullptr;
                        ga.uniforms_dev = (const float *)dummy.p + 8; ga.temperature = 1.0f;
                        ga.out_samples_dev = (float *)dummy.p + 4;
                        ok = wn_generate(h, &ga, nullptr) == WN_OK && wn_sync_check(h, nullptr) == WN_OK;
                    }
                    dummy.release();
                    h->finalized = false;
                    if (ok) {
                        h->info.cluster_path = 2;
                        h->v2_note = "cluster path: " + std::to_string(want16) + " clusters of 16 + " + std::to_string((rem + 1) / 2) + " of 8 + " + std::to_string(Mt) + " tail CTAs";
                    } else {
                        h->v2_n16 = 0;
                        cudaGetLastError();
                        h->err.clear();
                    }
                }
            }
        }
    }
    (void)St; (void)Sm;
    h->finalized = true;
    return WN_OK;
}

int wn_get_plan(const wn_handle *h, wn_plan *plan)
{
    if (!h || !plan) return WN_ERR_ARG;
    if (!h->finalized) return WN_ERR_STATE;
    *plan = h->plan;
    return WN_OK;
}

int wn_get_info(const wn_handle *h, wn_info *info)
{
    if (!h || !info) return WN_ERR_ARG;
    if (!h->finalized) return WN_ERR_STATE;
    *info = h->info;
    info->kernel_launches = h->launches;
    return WN_OK;
}

int wn_upsample(wn_handle *h, const float *mel_dev, int rows, int t_mel, float *out_dev, void *stream)
{
    if (!h) return WN_ERR_ARG;
    if (!h->finalized) return fail(h, WN_ERR_STATE, "wn_upsample before wn_finalize");
    const wn_config &c = h->cfg;
    if (!c.lc_channels || !c.n_upsample) return fail(h, WN_ERR_STATE, "model has no local-condition upsampling network");
    if (!mel_dev || !out_dev || rows < 1 || t_mel < 1) return fail(h, WN_ERR_ARG, "wn_upsample: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    const int C = c.lc_channels;
    long long rows_in = (long long)rows * t_mel;
    size_t tmp_floats = 0;
    {
        long long r = rows_in;
        for (int i = 0; i + 1 < c.n_upsample; ++i) { r *= c.upsample_factor[i]; tmp_floats = std::max<size_t>(tmp_floats, (size_t)r * C); }
    }
    if (tmp_floats) { CUDA_TRY(h, h->up_tmp0.ensure(tmp_floats * 4)); CUDA_TRY(h, h->up_tmp1.ensure(tmp_floats * 4)); }
    const float *src = mel_dev;
    for (int i = 0; i < c.n_upsample; ++i) {
        const int F = c.upsample_factor[i];
        float *dst = (i + 1 == c.n_upsample) ? out_dev : (float *)((i & 1) ? h->up_tmp1.p : h->up_tmp0.p);
        long long total = rows_in * F * C;
        int blocks = (int)std::min<long long>((total + 255) / 256, (long long)h->sm_count * 16);
        wn_upsample_stage_kernel<<<blocks, 256, 0, st>>>(src, dst, (const float *)h->upk.p + h->up_off[i], rows_in, F, C);
        CUDA_TRY(h, cudaGetLastError());
        h->launches++;
        src = dst;
        rows_in *= F;
    }
    return WN_OK;
}

int wn_generate(wn_handle *h, const wn_generate_args *a, void *stream)
{
    if (!h || !a) return WN_ERR_ARG;
    if (!h->finalized) return fail(h, WN_ERR_STATE, "wn_generate before wn_finalize");
    const wn_config &c = h->cfg;
    if (a->rows < 1 || a->rows > c.batch) return fail(h, WN_ERR_ARG, "rows=%d outside 1..batch_size=%d", a->rows, c.batch);
    if (a->T < 0) return fail(h, WN_ERR_ARG, "T must be >= 0");
    if (a->T == 0) return WN_OK;
    if (a->n_forced < 1 || !a->forced_dev) return fail(h, WN_ERR_ARG, "n_forced must be >= 1 (forced[row][0] is the initial sample)");
    if (!a->uniforms_dev || !a->out_samples_dev) return fail(h, WN_ERR_ARG, "uniforms_dev / out_samples_dev required");
    if (c.gc_channels && !a->gc_ids) return fail(h, WN_ERR_ARG, "gc_ids required: the model is globally conditioned (generate.py:72-77)");
    if (!c.scalar_input && !(a->temperature > 0.0f)) return fail(h, WN_ERR_ARG, "temperature must be > 0");
    cudaStream_t st = (cudaStream_t)stream;

    WnParams p = h->base;
    p.N = a->rows;
    p.T = a->T;
    p.n_forced = a->n_forced;
    p.t_lc = a->lc_dev ? a->t_lc : 0;
    p.lc_shift = a->lc_shift;
    p.temperature = a->temperature;
    for (int b = 0; b < a->rows; ++b) {
        int tr = a->T_row ? a->T_row[b] : a->T;
        if (tr < 0 || tr > a->T) return fail(h, WN_ERR_ARG, "T_row[%d]=%d outside 0..T", b, tr);
        p.T_row[b] = tr;
        int gid = a->gc_ids ? a->gc_ids[b] : 0;
        if (c.gc_channels && (gid < 0 || gid >= c.gc_cardinality)) return fail(h, WN_ERR_ARG, "gc_ids[%d]=%d outside 0..%d", b, gid, c.gc_cardinality - 1);
        p.gc_id[b] = gid;
    }
    p.forced = a->forced_dev;
    p.lc_up = c.lc_channels ? a->lc_dev : nullptr;
    p.mel = nullptr;
    if (a->mel_dev && c.lc_channels) {
        if (!c.n_upsample) return fail(h, WN_ERR_STATE, "mel_dev given but the model has no upsampling network");
        if (a->t_mel < 1) return fail(h, WN_ERR_ARG, "t_mel must be >= 1");
        long hop = 1;
        for (int i = 0; i < c.n_upsample; ++i) hop *= c.upsample_factor[i];
        if (h->v2 && c.n_upsample == 3) {
            // folded into the layer CTAs: mel frames are staged by TMA, nothing is materialised
            p.mel = a->mel_dev; p.t_mel = a->t_mel; p.n_up = 3; p.hop = (int)hop;
            p.upk = (const float *)h->upk.p;
            for (int i = 0; i < 3; ++i) { p.up_f[i] = c.upsample_factor[i]; p.up_off[i] = h->up_off[i]; }
            p.lc_up = nullptr;
            p.t_lc = (int)(a->t_mel * hop);
        } else {
            const int t_up = (int)(a->t_mel * hop);
            CUDA_TRY(h, h->g_lc.ensure((size_t)a->rows * t_up * c.lc_channels * 4));
            int rc = wn_upsample(h, a->mel_dev, a->rows, a->t_mel, (float *)h->g_lc.p, st);
            if (rc) return rc;
            p.lc_up = (const float *)h->g_lc.p;
            p.t_lc = t_up;
        }
    }
    p.uniforms = a->uniforms_dev;
    p.noise = nullptr;
    if (h->v2) {
        // the draw's noise for every (row, step), transformed once before the launch (see wn_noise_prep_kernel)
        const int nr = c.out_channels / 3;
        const long long n = (long long)a->rows * a->T * (nr + 1);
        CUDA_TRY(h, h->g_noise.ensure((size_t)n * 4));
        const unsigned blocks = (unsigned)std::min<long long>((n + 255) / 256, (long long)8 * h->sm_count);
        wn_noise_prep_kernel<<<blocks, 256, 0, st>>>((const float *)a->uniforms_dev, (float *)h->g_noise.p, n, nr);
        CUDA_TRY(h, cudaGetLastError());
        p.noise = (const float *)h->g_noise.p;
        h->launches++;
    }
    p.out_samples = a->out_samples_dev;
    p.out_logits = a->out_logits_dev;
    // The mailbox / ring strides were laid out for cfg.batch rows; a smaller `rows` only uses a prefix
    // of each [N] dimension, so the kernel must index with the allocation's N.  Keep N = cfg.batch in
    // the strides by running the kernel with N = batch and T_row = 0 for the unused rows.
    p.N = c.batch;
    for (int b = a->rows; b < c.batch; ++b) { p.T_row[b] = 0; p.gc_id[b] = 0; }

    CUDA_TRY(h, cudaMemsetAsync(h->mbox.p, 0, h->mbox_bytes, st));
    // Shared-ring builds of the layer kernel (one tagged ring per layer instead of four private copies, wn_kernel_v2.cuh) from
    // kSharedRingRows rows on, where the private copies overflow the L2; WN_SHARED_RING=0 / 1 forces either build for A/B runs.
    const int sr_env = getenv("WN_SHARED_RING") ? atoi(getenv("WN_SHARED_RING")) : -1;
    const bool shared_ring = h->v2 && !h->prof_on && p.T < (1 << 30) && (sr_env >= 0 ? sr_env != 0 : a->rows >= kSharedRingRows);
    // the cluster path reads the private rings only once they hold data; the shared ring's tags must start at zero
    if (h->ring_bytes && (!h->v2 || shared_ring)) CUDA_TRY(h, cudaMemsetAsync(h->ring.p, 0, h->ring_bytes, st));
    CUDA_TRY(h, cudaMemsetAsync(h->status.p, 0, 32, st));
    p.prof = nullptr;
    if (h->prof_on) {
        CUDA_TRY(h, h->prof.ensure((size_t)p.grid * 16 * sizeof(long long)));
        CUDA_TRY(h, cudaMemsetAsync(h->prof.p, 0, (size_t)p.grid * 16 * sizeof(long long), st));
        p.prof = (long long *)h->prof.p;
    }
    void *args[] = {&p};
    if (h->v2) {
        // kernel A (layer clusters) and kernel B (tail + sampler) run concurrently on two internal streams,
        // forked from and joined back into the caller's stream
        CUDA_TRY(h, cudaEventRecord(h->v2_fork, st));
        CUDA_TRY(h, cudaStreamWaitEvent(h->v2_sa, h->v2_fork, 0));
        CUDA_TRY(h, cudaStreamWaitEvent(h->v2_sb, h->v2_fork, 0));
        p.claim_slot = 4;
        if (h->v2_n16 > 0) {
            // layers [0, 4*n16) in clusters of 16, then the rest in clusters of 8 on a second stream, then the tail
            const int split = 4 * h->v2_n16;
            p.layer_base = 0; p.layer_end = split;
            CUDA_TRY(h, cudaLaunchKernel(shared_ring ? h->kernel_v2_layers16_sr : h->prof_on ? h->kernel_v2_layers16_prof : h->kernel_v2_layers16,
                                         dim3(16 * h->v2_n16), dim3(V2_NT), args, (size_t)h->v2_smem_layers, h->v2_sa));
            if (split < p.L) {
                WnParams p2 = p;
                p2.layer_base = split; p2.layer_end = p.L; p2.claim_slot = 6;
                void *args2[] = {&p2};
                CUDA_TRY(h, cudaStreamWaitEvent(h->v2_sc, h->v2_fork, 0));
                CUDA_TRY(h, cudaLaunchKernel(shared_ring ? h->kernel_v2_layers_sr : h->prof_on ? h->kernel_v2_layers_prof : h->kernel_v2_layers,
                                             dim3(((p.L - split + 1) / 2) * V2_CS), dim3(V2_NT), args2, (size_t)h->v2_smem_layers, h->v2_sc));
                CUDA_TRY(h, cudaEventRecord(h->v2_join_c, h->v2_sc));
                CUDA_TRY(h, cudaStreamWaitEvent(st, h->v2_join_c, 0));
                h->launches++;
            }
        } else {
            p.layer_base = 0; p.layer_end = p.L;
            CUDA_TRY(h, cudaLaunchKernel(shared_ring ? h->kernel_v2_layers_sr : h->prof_on ? h->kernel_v2_layers_prof : h->kernel_v2_layers,
                                         dim3(h->v2_grid_layers), dim3(V2_NT), args, (size_t)h->v2_smem_layers, h->v2_sa));
        }
        CUDA_TRY(h, cudaLaunchKernel(h->kernel_v2_tail, dim3(p.Mt), dim3(WN_NT), args, (size_t)h->v2_smem_tail, h->v2_sb));
        CUDA_TRY(h, cudaEventRecord(h->v2_join_a, h->v2_sa));
        CUDA_TRY(h, cudaEventRecord(h->v2_join_b, h->v2_sb));
        CUDA_TRY(h, cudaStreamWaitEvent(st, h->v2_join_a, 0));
        CUDA_TRY(h, cudaStreamWaitEvent(st, h->v2_join_b, 0));
        h->launches += 2;
        return WN_OK;
    }
    const void *fn = (h->kernel_many && a->rows >= kManyRows) ? h->kernel_many : h->kernel;
    CUDA_TRY(h, cudaLaunchCooperativeKernel(fn, dim3(p.grid), dim3(WN_NT), args, (size_t)h->smem_launch, st));
    h->launches++;
    return WN_OK;
}

/* Diagnostics (not part of the reference-facing surface): per-CTA phase cycle counters of the next launches. */
int wn_debug_profile(wn_handle *h, int enable, long long *out, int n)
{
    if (!h) return WN_ERR_ARG;
    h->prof_on = enable != 0;
    if (out && h->prof.p) {
        CUDA_TRY(h, cudaDeviceSynchronize());
        size_t bytes = std::min((size_t)n * sizeof(long long), h->prof.bytes);
        CUDA_TRY(h, cudaMemcpy(out, h->prof.p, bytes, cudaMemcpyDeviceToHost));
    }
    return WN_OK;
}

/* Diagnostics: average LL-mailbox round trip between two CTAs, in SM clock cycles. */
long long wn_debug_pingpong(int iters)
{
    unsigned long long *box = nullptr;
    long long *out = nullptr, host = -1;
    if (cudaMalloc(&box, 256) != cudaSuccess || cudaMalloc(&out, 8) != cudaSuccess) return -1;
    cudaMemset(box, 0, 256);
    void *args[] = {&box, &iters, &out};
    if (cudaLaunchCooperativeKernel((const void *)wn_pingpong_kernel, dim3(2), dim3(32), args, 0, 0) != cudaSuccess) return -2;
    if (cudaDeviceSynchronize() != cudaSuccess) return -3;
    cudaMemcpy(&host, out, 8, cudaMemcpyDeviceToHost);
    cudaFree(box); cudaFree(out);
    return host;
}

/* Diagnostics: round trips of CTA 0 with each of grid-1 partners; out has 2*grid entries. */
int wn_debug_pingpong_all(int grid, int iters, int mode, long long *out_host)
{
    unsigned long long *box = nullptr;
    long long *out = nullptr;
    size_t bb = (size_t)grid * 2 * 32 * 8 + 512;
    if (cudaMalloc(&box, bb) != cudaSuccess || cudaMalloc(&out, (size_t)grid * 2 * 8) != cudaSuccess) return -1;
    cudaMemset(box, 0, bb);
    cudaMemset(out, 0, (size_t)grid * 2 * 8);
    void *args[] = {&box, &iters, &out, &mode};
    if (cudaLaunchCooperativeKernel((const void *)wn_pingpong_all_kernel, dim3(grid), dim3(32), args, 0, 0) != cudaSuccess) return -2;
    if (cudaDeviceSynchronize() != cudaSuccess) return -3;
    cudaMemcpy(out_host, out, (size_t)grid * 2 * 8, cudaMemcpyDeviceToHost);
    cudaFree(box); cudaFree(out);
    return 0;
}

/* Diagnostics: ping-pong RTT matrix over (partner, grain of partner inbox, grain of CTA0 inbox). */
int wn_debug_pingpong_grid(int grid, const int *part_host, int np, int ng, int iters, long long *out_host, unsigned *smids_host)
{
    unsigned long long *box = nullptr; int *part = nullptr; long long *out = nullptr; unsigned *smids = nullptr;
    size_t bb = (size_t)ng * 2048;
    if (cudaMalloc(&box, bb) != cudaSuccess || cudaMalloc(&part, np * 4) != cudaSuccess ||
        cudaMalloc(&out, (size_t)np * ng * ng * 8) != cudaSuccess || cudaMalloc(&smids, grid * 4) != cudaSuccess) return -1;
    cudaMemset(box, 0, bb);
    cudaMemcpy(part, part_host, np * 4, cudaMemcpyHostToDevice);
    void *args[] = {&box, &part, &np, &ng, &iters, &out, &smids};
    if (cudaLaunchCooperativeKernel((const void *)wn_pingpong_grid_kernel, dim3(grid), dim3(32), args, 0, 0) != cudaSuccess) return -2;
    if (cudaDeviceSynchronize() != cudaSuccess) return -3;
    cudaMemcpy(out_host, out, (size_t)np * ng * ng * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(smids_host, smids, grid * 4, cudaMemcpyDeviceToHost);
    cudaFree(box); cudaFree(part); cudaFree(out); cudaFree(smids);
    return 0;
}

/* Diagnostics: average cycles of one polling round (see wn_pollbench_kernel). */
long long wn_debug_pollbench(int ctas, int iters, int warps, int lanes, int K)
{
    unsigned long long *box = nullptr; long long *out = nullptr;
    size_t bb = (size_t)ctas * 4096 * 8 + 65536;
    if (cudaMalloc(&box, bb) != cudaSuccess || cudaMalloc(&out, ctas * 8) != cudaSuccess) return -1;
    cudaMemset(box, 0, bb);
    wn_pollbench_kernel<<<ctas, 256>>>(box, iters, warps, lanes, K, out);
    if (cudaDeviceSynchronize() != cudaSuccess) return -2;
    std::vector<long long> h(ctas);
    cudaMemcpy(h.data(), out, ctas * 8, cudaMemcpyDeviceToHost);
    cudaFree(box); cudaFree(out);
    long long s = 0;
    for (long long v : h) s += v;
    return s / ctas;
}

/* Diagnostics: outcome of the per-process die calibration. */
const char *wn_debug_calib_note(void) { return g_calib.note.c_str(); }

/* Diagnostics: raw SM-map calibration round trips, out[n_sm*16]; returns n_sm (0 if not calibrated). */
int wn_debug_calib_raw(unsigned *out, int max_sm)
{
    if (g_calib.raw.empty() || g_calib.n_sm > max_sm) return 0;
    memcpy(out, g_calib.raw.data(), g_calib.raw.size() * 4);
    return g_calib.n_sm;
}

int wn_sync_check(wn_handle *h, void *stream)
{
    if (!h) return WN_ERR_ARG;
    if (!h->finalized) return fail(h, WN_ERR_STATE, "not finalized");
    CUDA_TRY(h, cudaStreamSynchronize((cudaStream_t)stream));
    int32_t st[4] = {0, 0, 0, 0};
    CUDA_TRY(h, cudaMemcpy(st, h->status.p, 16, cudaMemcpyDeviceToHost));
    if (st[0] != 0) return fail(h, WN_ERR_TIMEOUT, "persistent kernel aborted on its watchdog (CTA %d, thread %d)", st[1], st[2]);
    return WN_OK;
}

int wn_generate_host(wn_handle *h, const wn_generate_args *a, const float *mel_host, int t_mel)
{
    if (!h || !a) return WN_ERR_ARG;
    if (!h->finalized) return fail(h, WN_ERR_STATE, "wn_generate_host before wn_finalize");
    const wn_config &c = h->cfg;
    const int rows = a->rows;
    if (rows < 1 || rows > c.batch) return fail(h, WN_ERR_ARG, "rows=%d outside 1..batch_size=%d", rows, c.batch);
    wn_generate_args d = *a;
    int T = a->T;
    cudaStream_t st = 0;
    if (mel_host) {
        if (!c.lc_channels || !c.n_upsample) return fail(h, WN_ERR_STATE, "mel given but the model has no local conditioning");
        long f = 1;
        for (int i = 0; i < c.n_upsample; ++i) f *= c.upsample_factor[i];
        const int t_up = (int)(t_mel * f);
        size_t mel_bytes = (size_t)rows * t_mel * c.lc_channels * 4;
        CUDA_TRY(h, h->h_mel.ensure(mel_bytes));
        CUDA_TRY(h, cudaMemcpyAsync(h->h_mel.p, mel_host, mel_bytes, cudaMemcpyHostToDevice, st));
        d.mel_dev = (const float *)h->h_mel.p;       // wn_generate folds or materialises the upsampling itself
        d.t_mel = t_mel;
        d.lc_dev = nullptr;
        d.t_lc = t_up;
    } else if (a->lc_dev) {
        size_t b = (size_t)rows * a->t_lc * c.lc_channels * 4;
        CUDA_TRY(h, h->h_lc.ensure(b));
        CUDA_TRY(h, cudaMemcpyAsync(h->h_lc.p, a->lc_dev, b, cudaMemcpyHostToDevice, st));
        d.lc_dev = (const float *)h->h_lc.p;
    }
    if (T < 1) return fail(h, WN_ERR_ARG, "T must be >= 1");
    size_t fb = (size_t)rows * a->n_forced * 4;
    size_t ub = c.scalar_input ? (size_t)rows * T * (h->out_dim / 3 + 1) * 4 : (size_t)rows * T * 8;
    size_t ob = (size_t)rows * T * 4;
    if (a->n_forced < 1 || !a->forced_dev || !a->uniforms_dev || !a->out_samples_dev) return fail(h, WN_ERR_ARG, "forced / uniforms / out_samples required");
    CUDA_TRY(h, h->h_forced.ensure(fb));
    CUDA_TRY(h, h->h_unif.ensure(ub));
    CUDA_TRY(h, h->h_out.ensure(ob));
    CUDA_TRY(h, cudaMemcpyAsync(h->h_forced.p, a->forced_dev, fb, cudaMemcpyHostToDevice, st));
    CUDA_TRY(h, cudaMemcpyAsync(h->h_unif.p, a->uniforms_dev, ub, cudaMemcpyHostToDevice, st));
    d.forced_dev = (const float *)h->h_forced.p;
    d.uniforms_dev = h->h_unif.p;
    d.out_samples_dev = (float *)h->h_out.p;
    if (a->out_logits_dev) {
        CUDA_TRY(h, h->h_logits.ensure(ob * h->out_dim));
        d.out_logits_dev = (float *)h->h_logits.p;
    }
    int rc = wn_generate(h, &d, st);
    if (rc) return rc;
    CUDA_TRY(h, cudaMemcpyAsync(a->out_samples_dev, h->h_out.p, ob, cudaMemcpyDeviceToHost, st));
    if (a->out_logits_dev) CUDA_TRY(h, cudaMemcpyAsync(a->out_logits_dev, h->h_logits.p, ob * h->out_dim, cudaMemcpyDeviceToHost, st));
    return wn_sync_check(h, st);
}

/* ---- single step with persistent queues (wn_step.cuh) --------------------------------------------------------- */
struct wn_state {
    wn_handle *h;
    int rows;
    DevBuf buf;
    long long stride;
    int off_cq, off_lc, off_ring0;
};

static int ensure_raw_weights(wn_handle *h)
{
    if (h->raw.p) return WN_OK;
    const wn_config &c = h->cfg;
    const int L = c.n_layers;
    size_t total = 0;
    h->raw_off.clear();
    for (auto &kv : h->w) { h->raw_off[kv.first] = total; total += (kv.second.size() + 3) & ~(size_t)3; }
    const size_t zeros_off = total;
    total += 1024;                                                   // shared zero vector for absent biases
    std::vector<float> host(total, 0.0f);
    for (auto &kv : h->w) memcpy(host.data() + h->raw_off[kv.first], kv.second.data(), kv.second.size() * 4);
    CUDA_TRY(h, h->raw.ensure(total * 4));
    CUDA_TRY(h, cudaMemcpy(h->raw.p, host.data(), total * 4, cudaMemcpyHostToDevice));
    const float *base = (const float *)h->raw.p;
    auto ptr = [&](const std::string &nm) -> const float * {
        auto it = h->raw_off.find(nm);
        return it == h->raw_off.end() ? base + zeros_off : base + it->second;
    };
    std::vector<WnStepLayer> lay(L);
    std::vector<long long> roff(L);
    long long rt = 0;
    for (int l = 0; l < L; ++l) {
        const std::string pre = "wavenet/dilated_stack/layer" + std::to_string(l) + "/dilation_layer/";
        lay[l] = WnStepLayer{ptr(pre + "conv_filter/kernel"), ptr(pre + "conv_gate/kernel"), ptr(pre + "conv_filter/bias"), ptr(pre + "conv_gate/bias"),
                             ptr(pre + "gc_filter/kernel"), ptr(pre + "gc_gate/kernel"), ptr(pre + "lc_filter/kernel"), ptr(pre + "lc_gate/kernel"),
                             ptr(pre + "dense/kernel"), ptr(pre + "dense/bias"), ptr(pre + "skip/kernel"), ptr(pre + "skip/bias")};
        roff[l] = rt;
        rt += (long long)c.dilations[l] * c.residual_channels;
    }
    if (std::max(std::max(c.skip_channels, c.dilation_channels), std::max(c.residual_channels, h->out_dim)) > 1024)
        return fail(h, WN_ERR_ARG, "wn_step: channel counts above 1024 are not supported");
    h->step_ring_floats = rt;
    CUDA_TRY(h, h->raw_layers.ensure(L * sizeof(WnStepLayer)));
    CUDA_TRY(h, cudaMemcpy(h->raw_layers.p, lay.data(), L * sizeof(WnStepLayer), cudaMemcpyHostToDevice));
    CUDA_TRY(h, h->step_ring_off.ensure(L * sizeof(long long)));
    CUDA_TRY(h, cudaMemcpy(h->step_ring_off.p, roff.data(), L * sizeof(long long), cudaMemcpyHostToDevice));
    return WN_OK;
}

int wn_state_create(wn_handle *h, int rows, wn_state **out)
{
    if (!h || !out) return WN_ERR_ARG;
    *out = nullptr;
    if (!h->finalized) return fail(h, WN_ERR_STATE, "wn_state_create before wn_finalize");
    if (rows < 1 || rows > WN_MAX_BATCH) return fail(h, WN_ERR_ARG, "rows must be in 1..%d", WN_MAX_BATCH);
    int rc = ensure_raw_weights(h);
    if (rc) return rc;
    const wn_config &c = h->cfg;
    wn_state *s = new wn_state();
    s->h = h; s->rows = rows;
    s->off_cq = 4;
    s->off_lc = s->off_cq + ((std::max(c.initial_filter_width, 1) + 3) & ~3);
    s->off_ring0 = s->off_lc + ((std::max(c.lc_channels, 1) + 3) & ~3);
    s->stride = s->off_ring0 + h->step_ring_floats;
    cudaError_t e = s->buf.ensure((size_t)rows * s->stride * 4);
    if (e != cudaSuccess) { delete s; return fail(h, WN_ERR_CUDA, "wn_state_create: %s", cudaGetErrorString(e)); }
    *out = s;
    return wn_state_reset(h, s, nullptr);
}

int wn_state_reset(wn_handle *h, wn_state *s, void *stream)
{
    if (!h || !s) return WN_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(h, cudaMemsetAsync(s->buf.p, 0, (size_t)s->rows * s->stride * 4, st));
    // one-hot models: the causal queue starts with all-zero one-hot rows (ids -1), queue_initializer
    std::vector<int> ids(3, -1);
    ids[0] = 0;
    for (int b = 0; b < s->rows; ++b)
        CUDA_TRY(h, cudaMemcpyAsync((char *)s->buf.p + (size_t)b * s->stride * 4, ids.data(), 12, cudaMemcpyHostToDevice, st));
    CUDA_TRY(h, cudaStreamSynchronize(st));
    return WN_OK;
}

void wn_state_destroy(wn_state *s)
{
    if (!s) return;
    s->buf.release();
    delete s;
}

int wn_step(wn_handle *h, wn_state *s, const wn_step_args *a, void *stream)
{
    if (!h || !s || !a) return WN_ERR_ARG;
    if (s->h != h) return fail(h, WN_ERR_ARG, "wn_step: the state belongs to another handle");
    if (a->rows != s->rows) return fail(h, WN_ERR_ARG, "wn_step: rows=%d but the state holds %d rows", a->rows, s->rows);
    if (!a->x_in_dev) return fail(h, WN_ERR_ARG, "wn_step: x_in_dev required");
    if (!h->finalized) return fail(h, WN_ERR_STATE, "wn_step before wn_finalize");
    { int rc = ensure_raw_weights(h); if (rc) return rc; }
    const wn_config &c = h->cfg;
    if (c.gc_channels && !a->gc_ids) return fail(h, WN_ERR_ARG, "gc_ids required: the model is globally conditioned (generate.py:72-77)");
    if (!c.scalar_input && a->uniforms_dev && !(a->temperature > 0.0f)) return fail(h, WN_ERR_ARG, "temperature must be > 0");
    WnStepParams p;
    memset(&p, 0, sizeof p);
    p.rows = a->rows; p.L = c.n_layers; p.R = c.residual_channels; p.D = c.dilation_channels; p.S = c.skip_channels;
    p.O = h->out_dim; p.Q = c.quantization_channels; p.G = c.gc_channels; p.C = c.lc_channels; p.ifw = c.initial_filter_width;
    p.scalar = c.scalar_input; p.nr_mix = c.scalar_input ? h->out_dim / 3 : 0;
    p.plan = h->plan;
    for (int i = 0; i < c.n_layers; ++i) p.dil[i] = c.dilations[i];
    const float *base = (const float *)h->raw.p;
    auto ptr = [&](const char *nm) -> const float * { auto it = h->raw_off.find(nm); return it == h->raw_off.end() ? nullptr : base + it->second; };
    p.layers = (const WnStepLayer *)h->raw_layers.p;
    p.wc = ptr("wavenet/conv1d/kernel"); p.w1 = ptr("wavenet/conv1d_1/kernel"); p.w2 = ptr("wavenet/conv1d_2/kernel");
    p.b1 = ptr("wavenet/conv1d_1/bias"); p.b2 = ptr("wavenet/conv1d_2/bias");
    if (!p.b1 || !p.b2) {                        // absent biases: the zero vector at the end of the arena
        size_t total = 0;
        for (auto &kv : h->w) total += (kv.second.size() + 3) & ~(size_t)3;
        if (!p.b1) p.b1 = base + total;
        if (!p.b2) p.b2 = base + total;
    }
    p.gc_table = ptr("wavenet/gc_embedding");
    p.state = (float *)s->buf.p; p.state_stride = s->stride;
    p.off_cq = s->off_cq; p.off_lc = s->off_lc; p.off_ring0 = s->off_ring0;
    p.ring_off = (const long long *)h->step_ring_off.p;
    p.x_in = a->x_in_dev; p.lc_row = c.lc_channels ? a->lc_row_dev : nullptr;
    for (int b = 0; b < a->rows; ++b) {
        int gid = a->gc_ids ? a->gc_ids[b] : 0;
        if (c.gc_channels && (gid < 0 || gid >= c.gc_cardinality)) return fail(h, WN_ERR_ARG, "gc_ids[%d]=%d outside 0..%d", b, gid, c.gc_cardinality - 1);
        p.gc_id[b] = gid;
    }
    p.uniforms = a->uniforms_dev; p.temperature = a->temperature;
    p.out_logits = a->out_logits_dev; p.out_probs = c.scalar_input ? nullptr : a->out_probs_dev; p.out_sample = a->out_sample_dev;
    const int maxc = std::max(std::max(c.skip_channels, c.dilation_channels), std::max(c.residual_channels, h->out_dim));
    p.smem_maxc = maxc;
    const int R = p.R, D = p.D, S = p.S, O = p.O;
    size_t floats = 2 * R + 3 * D + 2 * S + ((std::max(O, WN_NT) + 3) & ~3) + maxc + ((p.G + 3) & ~3) + ((p.C + 3) & ~3) + ((p.ifw + 3) & ~3) + 32 + 2 * 18 +
                    (size_t)64 * maxc + 16;
    const size_t smem = floats * 4;
    if (smem > (size_t)kMaxDynSmem) return fail(h, WN_ERR_ARG, "wn_step: shared-memory need %zu exceeds the device limit", smem);
    CUDA_TRY(h, cudaFuncSetAttribute(wn_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    wn_step_kernel<<<a->rows, WN_NT, smem, (cudaStream_t)stream>>>(p);
    CUDA_TRY(h, cudaGetLastError());
    h->launches++;
    return WN_OK;
}

int wn_mu_law_encode(const float *audio_dev, int64_t n, int quantization_channels, int32_t *out_dev, void *stream)
{
    if (!audio_dev || !out_dev || n < 0) return WN_ERR_ARG;
    if (n == 0) return WN_OK;
    int blocks = (int)std::min<int64_t>((n + 255) / 256, 148 * 8);
    wn_mu_law_encode_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(audio_dev, n, (float)(quantization_channels - 1), out_dev);
    return cudaGetLastError() == cudaSuccess ? WN_OK : WN_ERR_CUDA;
}

int wn_mu_law_decode(const float *in_dev, int64_t n, int quantization_channels, int quantization, float *out_dev, void *stream)
{
    if (!in_dev || !out_dev || n < 0) return WN_ERR_ARG;
    if (n == 0) return WN_OK;
    int blocks = (int)std::min<int64_t>((n + 255) / 256, 148 * 8);
    wn_mu_law_decode_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(in_dev, n, (float)(quantization_channels - 1), quantization, out_dev);
    return cudaGetLastError() == cudaSuccess ? WN_OK : WN_ERR_CUDA;
}

int wn_mol_sample(const float *y_dev, const float *uniforms_dev, int64_t rows, int nr_mix, float log_scale_min, float *out_dev, void *stream)
{
    if (!y_dev || !uniforms_dev || !out_dev || rows < 0 || nr_mix < 1) return WN_ERR_ARG;
    if (rows == 0) return WN_OK;
    int blocks = (int)std::min<int64_t>((rows + 127) / 128, 148 * 8);
    wn_mol_sample_kernel<<<blocks, 128, 0, (cudaStream_t)stream>>>(y_dev, uniforms_dev, rows, nr_mix, log_scale_min, out_dev);
    return cudaGetLastError() == cudaSuccess ? WN_OK : WN_ERR_CUDA;
}

int wn_mol_loss(const float *y_hat_dev, const float *y_dev, int64_t rows, int nr_mix, int num_class, float log_scale_min,
                float *loss_out_dev, double *sum_out_dev, void *stream)
{
    if (!y_hat_dev || !y_dev || rows < 0 || nr_mix < 1 || num_class < 2 || (!loss_out_dev && !sum_out_dev)) return WN_ERR_ARG;
    if (sum_out_dev && cudaMemsetAsync(sum_out_dev, 0, sizeof(double), (cudaStream_t)stream) != cudaSuccess) return WN_ERR_CUDA;
    if (rows == 0) return WN_OK;
    int blocks = (int)std::min<int64_t>((rows + 127) / 128, 148 * 8);
    const float half_bin = (float)(1.0 / (num_class - 1)), log_half = (float)std::log((num_class - 1) / 2.0);
    wn_mol_loss_kernel<<<blocks, 128, 0, (cudaStream_t)stream>>>(y_hat_dev, y_dev, rows, nr_mix, log_scale_min, half_bin, log_half,
                                                                 loss_out_dev, sum_out_dev);
    return cudaGetLastError() == cudaSuccess ? WN_OK : WN_ERR_CUDA;
}

}  // extern "C"

/* ---- STFT -> mel (utils/audio.py:69-75) ---------------------------------------------------------------- */
namespace {
struct MelTables {
    int sr, n_fft, win, n_mels;
    double *window = nullptr; double2 *twiddle = nullptr; int *start = nullptr, *len = nullptr, *off = nullptr; float *w = nullptr;
};
std::vector<MelTables> g_mel_tables;
std::mutex g_mel_mutex;

double hz_to_mel_slaney(double f)
{
    const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = std::log(6.4) / 27.0;
    return f >= min_log_hz ? min_log_mel + std::log(f / min_log_hz) / logstep : f / f_sp;
}
double mel_to_hz_slaney(double m)
{
    const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = std::log(6.4) / 27.0;
    return m >= min_log_mel ? min_log_hz * std::exp(logstep * (m - min_log_mel)) : f_sp * m;
}

// tables of librosa.stft's window / FFT twiddles and librosa.filters.mel (Slaney scale, area normalised)
const MelTables *mel_tables(int sr, int n_fft, int win, int n_mels)
{
    std::lock_guard<std::mutex> lock(g_mel_mutex);
    for (const MelTables &t : g_mel_tables)
        if (t.sr == sr && t.n_fft == n_fft && t.win == win && t.n_mels == n_mels) return &t;
    MelTables t; t.sr = sr; t.n_fft = n_fft; t.win = win; t.n_mels = n_mels;
    const double PI = 3.14159265358979323846;
    std::vector<double> window(win);
    for (int i = 0; i < win; ++i) window[i] = 0.5 - 0.5 * std::cos(2.0 * PI * i / win);      // periodic Hann
    std::vector<double2> tw(n_fft / 2);
    for (int k = 0; k < n_fft / 2; ++k) tw[k] = make_double2(std::cos(2.0 * PI * k / n_fft), -std::sin(2.0 * PI * k / n_fft));
    const int n_bins = n_fft / 2 + 1;
    std::vector<double> mel_f(n_mels + 2);
    const double m_lo = hz_to_mel_slaney(0.0), m_hi = hz_to_mel_slaney(sr / 2.0);
    for (int i = 0; i < n_mels + 2; ++i) mel_f[i] = mel_to_hz_slaney(m_lo + (m_hi - m_lo) * i / (n_mels + 1));
    std::vector<int> start(n_mels), len(n_mels), off(n_mels);
    std::vector<float> w;
    for (int i = 0; i < n_mels; ++i) {
        const double enorm = 2.0 / (mel_f[i + 2] - mel_f[i]);
        int first = -1, last = -2;
        std::vector<float> row(n_bins);
        for (int k = 0; k < n_bins; ++k) {
            const double fk = (sr / 2.0) * k / (n_bins - 1);
            const double lower = (fk - mel_f[i]) / (mel_f[i + 1] - mel_f[i]);
            const double upper = (mel_f[i + 2] - fk) / (mel_f[i + 2] - mel_f[i + 1]);
            const double v = std::max(0.0, std::min(lower, upper)) * enorm;
            row[k] = (float)v;
            if (row[k] != 0.0f) { if (first < 0) first = k; last = k; }
        }
        if (first < 0) { first = 0; last = -1; }
        start[i] = first; len[i] = last - first + 1; off[i] = (int)w.size();
        for (int k = first; k <= last; ++k) w.push_back(row[k]);
    }
    if (w.empty()) w.push_back(0.0f);
    bool ok = cudaMalloc(&t.window, win * 8) == cudaSuccess && cudaMalloc(&t.twiddle, (n_fft / 2) * 16) == cudaSuccess &&
              cudaMalloc(&t.start, n_mels * 4) == cudaSuccess && cudaMalloc(&t.len, n_mels * 4) == cudaSuccess &&
              cudaMalloc(&t.off, n_mels * 4) == cudaSuccess && cudaMalloc(&t.w, w.size() * 4) == cudaSuccess;
    ok = ok && cudaMemcpy(t.window, window.data(), win * 8, cudaMemcpyHostToDevice) == cudaSuccess &&
         cudaMemcpy(t.twiddle, tw.data(), (n_fft / 2) * 16, cudaMemcpyHostToDevice) == cudaSuccess &&
         cudaMemcpy(t.start, start.data(), n_mels * 4, cudaMemcpyHostToDevice) == cudaSuccess &&
         cudaMemcpy(t.len, len.data(), n_mels * 4, cudaMemcpyHostToDevice) == cudaSuccess &&
         cudaMemcpy(t.off, off.data(), n_mels * 4, cudaMemcpyHostToDevice) == cudaSuccess &&
         cudaMemcpy(t.w, w.data(), w.size() * 4, cudaMemcpyHostToDevice) == cudaSuccess;
    if (!ok) return nullptr;
    g_mel_tables.push_back(t);
    return &g_mel_tables.back();
}
}  // namespace

extern "C" {

int wn_melspectrogram(const float *wav_dev, int rows, int64_t n, const wn_mel_config *mc, float *out_dev, void *stream)
{
    if (!wav_dev || !out_dev || !mc || rows < 1 || n < 2) return fail(nullptr, WN_ERR_ARG, "wn_melspectrogram: bad argument");
    if (mc->fft_size < 64 || mc->fft_size > 8192 || (mc->fft_size & (mc->fft_size - 1)))
        return fail(nullptr, WN_ERR_ARG, "fft_size must be a power of two in 64..8192");
    if (mc->win_size < 1 || mc->win_size > mc->fft_size || mc->hop_size < 1 || mc->num_mels < 1 || mc->num_mels > 256)
        return fail(nullptr, WN_ERR_ARG, "bad win_size / hop_size / num_mels");
    if (n <= mc->fft_size / 2) return fail(nullptr, WN_ERR_ARG, "signal shorter than fft_size/2 cannot be reflect-padded");
    const MelTables *t = mel_tables(mc->sample_rate, mc->fft_size, mc->win_size, mc->num_mels);
    if (!t) { cudaGetLastError(); return fail(nullptr, WN_ERR_CUDA, "wn_melspectrogram: table upload failed (no CUDA device?)"); }
    WnMelParams p;
    p.wav = wav_dev; p.n = n; p.rows = rows; p.frames = (int)(1 + n / mc->hop_size);
    p.n_fft = mc->fft_size; p.log2n = 0;
    while ((1 << p.log2n) < p.n_fft) ++p.log2n;
    p.hop = mc->hop_size; p.win = mc->win_size; p.win_off = (mc->fft_size - mc->win_size) / 2;
    p.n_mels = mc->num_mels; p.n_bins = mc->fft_size / 2 + 1;
    p.window = t->window; p.twiddle = t->twiddle; p.mel_start = t->start; p.mel_len = t->len; p.mel_off = t->off; p.mel_w = t->w;
    p.preemph = mc->preemphasize ? (double)mc->preemphasis : 0.0;
    p.min_level = (float)std::exp(mc->min_level_db / 20.0 * std::log(10.0));
    p.ref_level_db = mc->ref_level_db; p.min_level_db = mc->min_level_db; p.max_abs = mc->max_abs_value;
    p.out = out_dev;
    const long long items = (long long)rows * ((p.frames + 1) / 2);        // one CTA iteration = one PAIR of frames
    static const bool radix2 = getenv("WN_MEL_RADIX2") != nullptr;         // A/B: the round-1 radix-2 kernel
    if (p.n_fft <= 4096 && !radix2) {
        // Stockham radix-8 passes through a padded ping-pong buffer (2 x 9/8 x n_fft double2); the magnitudes reuse the free half
        const size_t smem = (size_t)2 * (p.n_fft + (p.n_fft >> 3)) * 16;
        static const int want_ctas = getenv("WN_MEL_CTAS") ? atoi(getenv("WN_MEL_CTAS")) : 3;     // A/B: 2 = the 122-register build
        const int fit = std::max(1, (int)((size_t)(226 << 10) / (smem + 1024)));               // CTAs per SM that fit in shared memory
        const int per_sm = std::min(fit, want_ctas >= 3 ? 3 : 2);
        auto kern = per_sm >= 3 ? wn_mel_kernel_s8<3> : wn_mel_kernel_s8<2>;
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
            cudaGetLastError();
            return fail(nullptr, WN_ERR_CUDA, "wn_melspectrogram: shared memory request failed");
        }
        const int grid = (int)std::min<long long>(items, 148LL * per_sm);
        kern<<<grid, 256, smem, (cudaStream_t)stream>>>(p);
    } else {
        const size_t smem = (size_t)p.n_fft * 16 + (size_t)p.n_bins * 8 + 16;    // FFT buffer + two magnitude spectra
        if (cudaFuncSetAttribute(wn_mel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
            cudaGetLastError();
            return fail(nullptr, WN_ERR_CUDA, "wn_melspectrogram: shared memory request failed");
        }
        const int grid = (int)std::min<long long>(items, 148LL * 16);
        wn_mel_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(p);
    }
    if (cudaGetLastError() != cudaSuccess) return fail(nullptr, WN_ERR_CUDA, "wn_melspectrogram: launch failed");
    return WN_OK;
}

}  // extern "C"

// taco_api.cu -- C ABI (include/taco_b200.h) and host orchestration of the Tacotron text->mel path.
// Graph wiring follows tacotron/tacotron.py:36-235 of the reference; see taco_kernels.cuh for the kernels.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <string>
#include <thread>
#include <vector>

#include "../../include/taco_b200.h"
#include "taco_kernels.cuh"
#include "taco_gemm_tc.cuh"

using namespace taco;

namespace {

thread_local std::string g_create_error;

struct DevBuf {
    float *p = nullptr;
    size_t n = 0;
};

struct ConvLayer {           // tf.layers.conv1d / dense (+ optional batch norm)
    float *W = nullptr, *b = nullptr, *scale = nullptr, *shift = nullptr;
    int k = 1, ci = 0, co = 0;
};

struct GruDev {              // one bidirectional layer
    float *Wx = nullptr, *bx = nullptr;       // (n_in, 6U), (6U)
    float *Wgh[2] = {nullptr, nullptr}, *Wch[2] = {nullptr, nullptr};
    int n_in = 0, U = 0;
};

struct CbhgDev {
    std::vector<ConvLayer> bank, proj;
    ConvLayer dense;          // only when the widths differ
    bool has_dense = false;
    std::vector<ConvLayer> highway;   // interleaved (U, 2U)
    GruDev rnn;
    int n_in = 0, K = 0, C = 0, U = 0;
};

}  // namespace

struct taco_handle {
    taco_config cfg;
    std::string err;
    std::map<std::string, std::vector<float>> w;
    bool finalized = false;
    int device = 0, sm_count = 0;
    std::vector<void *> allocs;
    int64_t n_params = 0, launches = 0;

    // device weights
    float *embedding = nullptr, *spk_embedding = nullptr;
    std::vector<ConvLayer> spk_dense;         // before_highway, enc init, att init, dec init...
    std::vector<ConvLayer> enc_prenet;
    CbhgDev enc, post;
    ConvLayer memory_layer, final_dense;
    float *nv = nullptr, *ab = nullptr, *loc_conv_w = nullptr, *loc_conv_b = nullptr, *loc_w = nullptr;
    float score_bias = 0.f;

    // decoder plan
    DecParams dp_host;
    DecParams *dp_dev = nullptr;
    long long *prof_dev = nullptr;            // in-kernel phase profile of the last decoder launch (TACO_PROFILE=1)
    float *dec_img = nullptr;
    float *dec_p0_init = nullptr;             // relu(b1) of the first decoder prenet layer: its activation for the zero <GO> frame (fused plan)
    int dec_grid = 0, dec_smem = 0, dec_maxK = 0, dec_dyn_max = 0;
    unsigned *barrier_dev = nullptr;
    size_t smem_optin = 0;
    int rnn_w_in_smem = 0;

    // per-call workspace
    char *ws = nullptr;
    size_t ws_cap = 0;
    // taco_synthesize_host: device-side copies of the caller's host buffers (grow-only) and a pinned two-slot staging ring
    char *host_io = nullptr;
    size_t host_io_cap = 0;
    char *stage[2] = {nullptr, nullptr};
    size_t stage_bytes = (size_t)16 << 20;    // TACO_STAGE_BYTES overrides (tests force many chunks on a small problem)
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t stage_ev[2] = {nullptr, nullptr};
    GemmProb *probs_dev = nullptr;
    size_t probs_cap = 0;
    int32_t *ids_lens_dev = nullptr;          // lengths + speaker ids
    size_t ids_lens_cap = 0;
    std::map<std::string, DevBuf> taps;

    // tensor-core conv / dense path (taco_gemm_tc.cuh): split, transposed weights per launch group, keyed by the first W pointer
    struct TcWeights { float *hi = nullptr, *lo = nullptr; int rows = 0, Kp = 0, Cip = 0; };
    std::map<const void *, TcWeights> tc_w;
    void *tc_encode = nullptr;                // cuTensorMapEncodeTiled
    unsigned *tc_err = nullptr;
    float *op_times_dev = nullptr;            // TACO_TIME_OPS=1: per-op event timings of the last call
    size_t op_times_cap = 0;
    long long *tc_dbg = nullptr;              // TACO_TC_DEBUG=1: in-kernel timeline of CTA 0 of every tensor-core launch
    bool tc_on = false;
    int64_t tc_launches = 0;
};

namespace {

#define CK(call)                                                                                      \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess) {                                                                      \
            h->err = std::string(#call) + ": " + cudaGetErrorString(e_);                              \
            return TACO_ERR_CUDA;                                                                     \
        }                                                                                             \
    } while (0)

int fail(taco_handle *h, int code, const std::string &msg) {
    h->err = msg;
    return code;
}

const std::string P = "model/inference/";

const std::vector<float> *find_w(taco_handle *h, const std::string &name, size_t n) {
    auto it = h->w.find(P + name);
    if (it == h->w.end()) {
        h->err = "missing weight " + P + name;
        return nullptr;
    }
    if (it->second.size() != n) {
        h->err = "weight " + P + name + " has " + std::to_string(it->second.size()) + " floats, expected " + std::to_string(n);
        return nullptr;
    }
    return &it->second;
}

float *upload(taco_handle *h, const float *src, size_t n) {
    float *d = nullptr;
    if (cudaMalloc(&d, std::max<size_t>(n, 1) * sizeof(float)) != cudaSuccess) return nullptr;
    h->allocs.push_back(d);
    if (n && cudaMemcpy(d, src, n * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) return nullptr;
    h->n_params += (int64_t)n;
    return d;
}

bool load_conv(taco_handle *h, const std::string &scope, int k, int ci, int co, bool bn, bool bias, ConvLayer *L,
               const char *kernel_suffix = "/conv1d") {
    const std::string ks = scope + kernel_suffix;
    auto W = find_w(h, ks + "/kernel", (size_t)k * ci * co);
    if (!W) return false;
    L->k = k; L->ci = ci; L->co = co;
    L->W = upload(h, W->data(), W->size());
    if (bias) {
        auto b = find_w(h, ks + "/bias", co);
        if (!b) return false;
        L->b = upload(h, b->data(), co);
    }
    if (bn) {
        const std::string bs = scope + "/batch_normalization/";
        auto g = find_w(h, bs + "gamma", co), be = find_w(h, bs + "beta", co), mu = find_w(h, bs + "moving_mean", co),
             va = find_w(h, bs + "moving_variance", co);
        if (!g || !be || !mu || !va) return false;
        std::vector<float> sc(co), sh(co);
        for (int i = 0; i < co; ++i) {   // tf.layers.batch_normalization inference, epsilon 1e-3: inv = gamma*rsqrt(var+eps)
            const float inv = (*g)[i] / sqrtf((*va)[i] + 1e-3f);
            sc[i] = inv;
            sh[i] = (*be)[i] - (*mu)[i] * inv;
        }
        L->scale = upload(h, sc.data(), co);
        L->shift = upload(h, sh.data(), co);
    }
    return L->W != nullptr;
}

bool load_dense(taco_handle *h, const std::string &scope, int ci, int co, bool bias, ConvLayer *L) {
    return load_conv(h, scope, 1, ci, co, false, bias, L, "");
}

bool load_birnn(taco_handle *h, const std::string &scope, int n_in, int U, GruDev *G) {
    G->n_in = n_in; G->U = U;
    std::vector<float> Wx((size_t)n_in * 6 * U), bx(6 * U);
    const char *dirs[2] = {"fw", "bw"};
    for (int d = 0; d < 2; ++d) {
        const std::string s = scope + "/bidirectional_rnn/" + dirs[d] + "/gru_cell";
        auto gk = find_w(h, s + "/gates/kernel", (size_t)(n_in + U) * 2 * U), gb = find_w(h, s + "/gates/bias", 2 * U),
             ck = find_w(h, s + "/candidate/kernel", (size_t)(n_in + U) * U), cb = find_w(h, s + "/candidate/bias", U);
        if (!gk || !gb || !ck || !cb) return false;
        for (int i = 0; i < n_in; ++i) {
            for (int c = 0; c < 2 * U; ++c) Wx[(size_t)i * 6 * U + d * 3 * U + c] = (*gk)[(size_t)i * 2 * U + c];
            for (int c = 0; c < U; ++c) Wx[(size_t)i * 6 * U + d * 3 * U + 2 * U + c] = (*ck)[(size_t)i * U + c];
        }
        for (int c = 0; c < 2 * U; ++c) bx[d * 3 * U + c] = (*gb)[c];
        for (int c = 0; c < U; ++c) bx[d * 3 * U + 2 * U + c] = (*cb)[c];
        G->Wgh[d] = upload(h, gk->data() + (size_t)n_in * 2 * U, (size_t)U * 2 * U);
        G->Wch[d] = upload(h, ck->data() + (size_t)n_in * U, (size_t)U * U);
        h->n_params -= 0;
    }
    G->Wx = upload(h, Wx.data(), Wx.size());
    G->bx = upload(h, bx.data(), bx.size());
    return G->Wx && G->bx;
}

bool load_cbhg(taco_handle *h, const std::string &scope, int n_in, int K, int C, const int *proj, int n_proj, int proj_w,
               int depth, int U, CbhgDev *D) {
    D->n_in = n_in; D->K = K; D->C = C; D->U = U;
    D->bank.resize(K);
    for (int k = 1; k <= K; ++k)
        if (!load_conv(h, scope + "/conv_bank/conv1d_" + std::to_string(k), k, n_in, C, true, true, &D->bank[k - 1])) return false;
    D->proj.resize(n_proj);
    int ci = K * C;
    for (int i = 0; i < n_proj; ++i) {
        if (!load_conv(h, scope + "/proj_" + std::to_string(i + 1), proj_w, ci, proj[i], true, true, &D->proj[i])) return false;
        ci = proj[i];
    }
    if (ci != n_in) { h->err = scope + ": proj_sizes[-1] must equal the input width (residual, modules.py:47-53)"; return false; }
    D->has_dense = ci != U;
    if (D->has_dense && !load_dense(h, scope + "/dense", ci, U, true, &D->dense)) return false;
    D->highway.resize(depth);
    for (int i = 0; i < depth; ++i) {
        const std::string s = scope + "/highway_" + std::to_string(i + 1);
        auto HW = find_w(h, s + "/H/kernel", (size_t)U * U), Hb = find_w(h, s + "/H/bias", U), TW = find_w(h, s + "/T/kernel", (size_t)U * U),
             Tb = find_w(h, s + "/T/bias", U);
        if (!HW || !Hb || !TW || !Tb) return false;
        std::vector<float> W((size_t)U * 2 * U), b(2 * U);
        for (int r = 0; r < U; ++r)
            for (int c = 0; c < U; ++c) {
                W[(size_t)r * 2 * U + 2 * c] = (*HW)[(size_t)r * U + c];
                W[(size_t)r * 2 * U + 2 * c + 1] = (*TW)[(size_t)r * U + c];
            }
        for (int c = 0; c < U; ++c) { b[2 * c] = (*Hb)[c]; b[2 * c + 1] = (*Tb)[c]; }
        ConvLayer &L = D->highway[i];
        L.k = 1; L.ci = U; L.co = 2 * U;
        L.W = upload(h, W.data(), W.size());
        L.b = upload(h, b.data(), b.size());
        if (!L.W || !L.b) return false;
    }
    return load_birnn(h, scope, U, U, &D->rnn);
}

// ---- decoder plan ------------------------------------------------------------------------------------
struct DenseSrc {
    std::vector<float> W;   // (K, N)
    std::vector<float> b;   // (N)
};

int build_decoder(taco_handle *h) {
    const taco_config &c = h->cfg;
    DecParams &dp = h->dp_host;
    memset(&dp, 0, sizeof(dp));
    const int G = h->sm_count;
    const int mem = 2 * c.enc_rnn_size, H = c.attention_state_size, A = c.attention_size, R = c.dec_rnn_size;
    const int nm = c.num_mels, OD = c.num_mels * c.reduction_factor;
    const std::string D = "decoder/";
    std::vector<DenseSrc> srcs;
    int np = 0;
    auto add_dense = [&](const std::string &kname, const std::string &bname, int K, int N, int epi, std::vector<std::pair<int, int>> segs,
                         int out_buf, int h_buf, int res_in, int res_out, int U) -> bool {
        if (np >= DEC_MAX_PHASES) { h->err = "decoder: too many phases"; return false; }
        auto W = find_w(h, kname, (size_t)K * N);
        if (!W) return false;
        DenseSrc s;
        s.W = *W;
        if (!bname.empty()) {
            auto b = find_w(h, bname, N);
            if (!b) return false;
            s.b = *b;
        } else {
            s.b.assign(N, 0.f);
        }
        srcs.push_back(std::move(s));
        DecPhase &ph = dp.ph[np++];
        ph.kind = PH_DENSE; ph.K = K; ph.N = N; ph.epi = epi;
        ph.ncp = (N + G - 1) / G;
        ph.pad = ph.ncp <= 2 ? ph.ncp : ((ph.ncp + 3) & ~3);
        ph.nseg = (int)segs.size();
        int ks = 0;
        for (int i = 0; i < ph.nseg; ++i) {
            ph.seg_buf[i] = segs[i].first; ph.seg_K[i] = segs[i].second; ks += segs[i].second;
            // the candidate phase of a GRU re-reads the inputs its gates phase staged one barrier earlier
            ph.seg_keep[i] = (epi == DE_CAND && i + 1 < ph.nseg) ? 1 : 0;
        }
        if (ks != K) { h->err = "decoder: segment widths do not add up for " + kname; return false; }
        ph.out_buf = out_buf; ph.h_buf = h_buf; ph.res_in = res_in; ph.res_out = res_out; ph.U = U;
        ph.N1 = N; ph.epi2 = DE_LINEAR; ph.out_buf2 = -1;
        h->dec_maxK = std::max(h->dec_maxK, K);
        return true;
    };
    // Fused plan (default; TACO_NO_FUSE=1 keeps one phase per layer): the output projection is linear and the first prenet layer
    // reads its last frame (helpers.py:39-41), so  relu(W1 . (Wout_last . o + bout_last) + b1) = relu((Wout_last W1) . o + (bout_last W1 + b1))
    // is evaluated BY THE OUTPUT PHASE ITSELF as extra columns of its matrix (product formed once in fp64): one grid barrier
    // less per decoder step.  The step loop is rotated accordingly -- prenet layers 2.., attention, decoder cells, then
    // [output projection | first prenet layer of the NEXT step] -- and the first step's prenet activation relu(b1) (the <GO>
    // frame is zero) is written by the init kernel.
    const bool fuse_out = getenv("TACO_NO_FUSE") == nullptr;
    h->dec_p0_init = nullptr;
    int ci = nm, prev = DB_X;
    for (int i = 0; i < c.n_dec_prenet; ++i) {
        const std::string s = D + "decoder_prenet/dense_" + std::to_string(i + 1);
        if (i == 0 && fuse_out) {
            auto b1 = find_w(h, s + "/bias", c.dec_prenet_sizes[0]);
            if (!b1) return TACO_ERR_STATE;
            std::vector<float> r0(*b1);
            for (auto &v : r0) v = std::max(v, 0.f);
            h->dec_p0_init = upload(h, r0.data(), r0.size());
            if (!h->dec_p0_init) return fail(h, TACO_ERR_CUDA, "uploading the decoder prenet init failed");
            h->n_params -= (int64_t)r0.size();
        } else if (!add_dense(s + "/kernel", s + "/bias", ci, c.dec_prenet_sizes[i], DE_RELU, {{prev, ci}}, DB_P0 + i, -1, -1, -1, 0)) return TACO_ERR_STATE;
        prev = DB_P0 + i;
        ci = c.dec_prenet_sizes[i];
    }
    {
        const std::string s = D + "attention_cell/gru_cell";
        if (!add_dense(s + "/gates/kernel", s + "/gates/bias", ci + mem + H, 2 * H, DE_GATES, {{prev, ci}, {DB_CTX, mem}, {DB_HATT, H}}, -1, DB_HATT, -1, -1, H))
            return TACO_ERR_STATE;
        if (!add_dense(s + "/candidate/kernel", s + "/candidate/bias", ci + mem + H, H, DE_CAND, {{prev, ci}, {DB_CTX, mem}, {DB_RH, H}}, DB_HATT, -1, -1, -1, H))
            return TACO_ERR_STATE;
    }
    if (!add_dense(D + "attention/query_layer/kernel", "", H, A, DE_QUERY, {{DB_HATT, H}}, DB_Q, -1, -1, -1, 0)) return TACO_ERR_STATE;
    dp.ph[np++].kind = PH_ATT_SCORE;
    dp.ph[np++].kind = PH_ATT_CTX;
    srcs.emplace_back();
    srcs.emplace_back();
    if (!add_dense(D + "concat_projection/kernel", D + "concat_projection/bias", H + mem, R, DE_LINEAR, {{DB_HATT, H}, {DB_CTX, mem}}, DB_O0, -1, -1, -1, 0))
        return TACO_ERR_STATE;
    for (int i = 0; i < c.dec_layer_num; ++i) {
        const std::string s = D + "cell_" + std::to_string(i + 1) + "/gru_cell";
        if (!add_dense(s + "/gates/kernel", s + "/gates/bias", 2 * R, 2 * R, DE_GATES, {{DB_O0 + i, R}, {DB_H1 + i, R}}, -1, DB_H1 + i, -1, -1, R)) return TACO_ERR_STATE;
        if (!add_dense(s + "/candidate/kernel", s + "/candidate/bias", 2 * R, R, DE_CAND, {{DB_O0 + i, R}, {DB_RH, R}}, DB_H1 + i, -1, DB_O0 + i, DB_O0 + i + 1, R))
            return TACO_ERR_STATE;
    }
    if (!add_dense(D + "output_projection/kernel", D + "output_projection/bias", R, OD, DE_OUT, {{DB_O0 + c.dec_layer_num, R}}, DB_X, -1, -1, -1, 0))
        return TACO_ERR_STATE;
    if (fuse_out) {
        const int P0 = c.dec_prenet_sizes[0];
        const std::string s1 = D + "decoder_prenet/dense_1";
        auto W1 = find_w(h, s1 + "/kernel", (size_t)nm * P0), b1 = find_w(h, s1 + "/bias", P0);
        if (!W1 || !b1) return TACO_ERR_STATE;
        DenseSrc &so = srcs.back();                        // (R, OD) + (OD)
        DenseSrc f;
        f.W.assign((size_t)R * (OD + P0), 0.f);
        f.b.assign(OD + P0, 0.f);
        for (int k = 0; k < R; ++k) {
            for (int col = 0; col < OD; ++col) f.W[(size_t)k * (OD + P0) + col] = so.W[(size_t)k * OD + col];
            for (int j = 0; j < P0; ++j) {
                double acc = 0.0;
                for (int m = 0; m < nm; ++m) acc += (double)so.W[(size_t)k * OD + (OD - nm + m)] * (double)(*W1)[(size_t)m * P0 + j];
                f.W[(size_t)k * (OD + P0) + OD + j] = (float)acc;
            }
        }
        for (int col = 0; col < OD; ++col) f.b[col] = so.b[col];
        for (int j = 0; j < P0; ++j) {
            double acc = (double)(*b1)[j];
            for (int m = 0; m < nm; ++m) acc += (double)so.b[OD - nm + m] * (double)(*W1)[(size_t)m * P0 + j];
            f.b[OD + j] = (float)acc;
        }
        srcs.back() = std::move(f);
        DecPhase &ph = dp.ph[np - 1];
        ph.N = OD + P0; ph.N1 = OD; ph.epi2 = DE_RELU; ph.out_buf2 = DB_P0;
        ph.ncp = (ph.N + G - 1) / G;
        ph.pad = ph.ncp <= 2 ? ph.ncp : ((ph.ncp + 3) & ~3);
    }
    dp.n_phases = np;

    // which input slots were final before the previous phase started (staged while waiting at the barrier)
    {
        int writer[DB_COUNT];
        for (int b = 0; b < DB_COUNT; ++b) writer[b] = -100;
        for (int i = 0; i < np; ++i) {
            const DecPhase &ph = dp.ph[i];
            if (ph.kind == PH_ATT_CTX) writer[DB_CTX] = i;
            if (ph.kind != PH_DENSE) continue;
            if (ph.epi == DE_RELU || ph.epi == DE_LINEAR) writer[ph.out_buf] = i;
            if (ph.epi == DE_CAND) { writer[ph.out_buf] = i; if (ph.res_out >= 0) writer[ph.res_out] = i; }
            if (ph.epi == DE_OUT) writer[DB_X] = i;
            if (ph.N1 < ph.N && ph.out_buf2 >= 0) writer[ph.out_buf2] = i;
        }
        for (int i = 1; i < np; ++i) {
            DecPhase &ph = dp.ph[i];
            if (ph.kind != PH_DENSE) continue;
            for (int sidx = 0; sidx < ph.nseg; ++sidx) {
                const int b = ph.seg_buf[sidx];
                const bool stable = b != DB_RH && b != DB_U && writer[b] != i - 1 && writer[b] != i;
                ph.seg_pre[sidx] = (stable && !ph.seg_keep[sidx]) ? 1 : 0;
            }
        }
    }
    // per-CTA shared-memory image
    int off = 0;
    for (int i = 0; i < np; ++i) {
        DecPhase &ph = dp.ph[i];
        if (ph.kind != PH_DENSE) continue;
        off = (off + 3) & ~3;
        ph.w_off = off;
        off += ph.K * ph.pad;
        ph.b_off = off;
        off += ph.pad;
    }
    dp.img_floats = off;
    std::vector<float> img((size_t)G * off, 0.f);
    for (int i = 0; i < np; ++i) {
        const DecPhase &ph = dp.ph[i];
        if (ph.kind != PH_DENSE) continue;
        const DenseSrc &s = srcs[i];
        for (int cta = 0; cta < G; ++cta) {
            float *dst = img.data() + (size_t)cta * off;
            for (int cl = 0; cl < ph.ncp; ++cl) {
                const int col = cta * ph.ncp + cl;
                if (col >= ph.N) break;
                for (int k = 0; k < ph.K; ++k) dst[ph.w_off + k * ph.pad + cl] = s.W[(size_t)k * ph.N + col];
                dst[ph.b_off + cl] = s.b[col];
            }
        }
    }
    h->dec_img = upload(h, img.data(), img.size());
    h->n_params -= (int64_t)img.size();
    for (auto &s : srcs) h->n_params += (int64_t)(s.W.size() + s.b.size());
    if (!h->dec_img) return fail(h, TACO_ERR_CUDA, "uploading the decoder image failed");
    dp.img = h->dec_img;
    dp.att_type = c.attention_type; dp.A = A; dp.mem = mem; dp.H = H; dp.OD = OD; dp.nm = nm;

    // attention vectors
    if (c.attention_type == TACO_ATT_LOC_SEN) {
        auto v = find_w(h, D + "attention/attention_variable", A), b = find_w(h, D + "attention/attention_bias", A);
        auto cw = find_w(h, D + "attention/location_features_convolution/kernel", 31 * 32),
             cb = find_w(h, D + "attention/location_features_convolution/bias", 32),
             lw = find_w(h, D + "attention/location_features_layer/kernel", (size_t)32 * A);
        if (!v || !b || !cw || !cb || !lw) return TACO_ERR_STATE;
        h->nv = upload(h, v->data(), A);
        h->ab = upload(h, b->data(), A);
        h->loc_conv_w = upload(h, cw->data(), cw->size());
        h->loc_conv_b = upload(h, cb->data(), 32);
        h->loc_w = upload(h, lw->data(), lw->size());
        h->score_bias = 0.f;
    } else {
        auto v = find_w(h, D + "attention/attention_v", A);
        auto sb = find_w(h, D + "attention/attention_score_bias", 1);
        if (!v || !sb) return TACO_ERR_STATE;
        std::vector<float> nv(*v), ab(A, 0.f);
        if (c.attention_type == TACO_ATT_BAH_MON_NORM) {
            auto g = find_w(h, D + "attention/attention_g", 1), b = find_w(h, D + "attention/attention_b", A);
            if (!g || !b) return TACO_ERR_STATE;
            // _bahdanau_score(normalize=True): normed_v = g * v * rsqrt(sum(v^2)), evaluated in fp32 like TF
            float ss = 0.f;
            for (int i = 0; i < A; ++i) ss += (*v)[i] * (*v)[i];
            const float rs = 1.0f / sqrtf(ss);
            for (int i = 0; i < A; ++i) nv[i] = (*g)[0] * (*v)[i] * rs;
            ab = *b;
        }
        h->nv = upload(h, nv.data(), A);
        h->ab = upload(h, ab.data(), A);
        h->score_bias = (*sb)[0];
    }
    dp.nv = h->nv; dp.ab = h->ab; dp.score_bias = h->score_bias;
    dp.loc_conv_w = h->loc_conv_w; dp.loc_conv_b = h->loc_conv_b; dp.loc_w = h->loc_w;
    return TACO_OK;
}

struct DecSmem { int stage_floats, keys_floats, vals_floats; size_t bytes; };

DecSmem dec_smem_plan(const taco_handle *h, int N, int T_in, int chunks, int fslices, size_t limit) {
    const DecParams &dp = h->dp_host;
    const int fs = (dp.mem + fslices - 1) / fslices;
    const int cpos = (T_in + chunks - 1) / chunks;
    size_t stage = (size_t)h->dec_maxK * 32;
    stage = std::max(stage, (size_t)4 * T_in + (size_t)4 * fs + 16);
    stage = std::max(stage, (size_t)dp.A + T_in + 31 * 32 + 32 + 16);
    stage = (stage + 3) & ~(size_t)3;
    const size_t base = (size_t)((dp.img_floats + 3) & ~3) + DEC_WARPS * 4 * 32 + 2 * (size_t)dp.A + stage;
    DecSmem r{(int)stage, 0, 0, base * sizeof(float)};
    // keep this CTA's key chunk / value slice in shared memory when every item has its own CTA and it fits
    const size_t kf = (size_t)cpos * dp.A, vf = (size_t)T_in * fs;
    if (N * chunks <= h->dec_grid && (base + kf) * sizeof(float) <= limit) { r.keys_floats = (int)kf; r.bytes += kf * sizeof(float); }
    if (N * fslices <= h->dec_grid && r.bytes + vf * sizeof(float) <= limit) { r.vals_floats = (int)vf; r.bytes += vf * sizeof(float); }
    return r;
}

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace

extern "C" {

const char *taco_last_error(const taco_handle *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int taco_create(const taco_config *cfg, taco_handle **out) {
    if (!cfg || !out) { g_create_error = "taco_create: null argument"; return TACO_ERR_ARG; }
    const taco_config &c = *cfg;
    auto bad = [&](const char *m) { g_create_error = std::string("taco_create: ") + m; return TACO_ERR_ARG; };
    if (c.num_symbols < 2 || c.embedding_size < 1) return bad("bad embedding table shape");
    if (c.n_enc_prenet < 1 || c.n_enc_prenet > TACO_MAX_PRENET || c.n_dec_prenet < 1 || c.n_dec_prenet > TACO_MAX_PRENET) return bad("prenet depth out of range");
    if (c.n_enc_proj < 1 || c.n_enc_proj > TACO_MAX_PROJ || c.n_post_proj < 1 || c.n_post_proj > TACO_MAX_PROJ) return bad("projection depth out of range");
    if (c.dec_layer_num < 1 || c.dec_layer_num > TACO_MAX_DEC_LAYERS) return bad("dec_layer_num out of range");
    if (c.attention_type < 0 || c.attention_type > 2) return bad("unsupported attention_type (bah_mon, bah_mon_norm, loc_sen are built)");
    if (c.attention_size % 32 != 0) return bad("attention_size must be a multiple of 32");
    if (c.enc_rnn_size > 256 || c.post_rnn_size > 256 || c.enc_rnn_size % 4 || c.post_rnn_size % 4) return bad("rnn sizes must be multiples of 4 and <= 256");
    if (c.reduction_factor < 1 || c.max_iters < 1 || c.num_mels < 1 || c.num_freq < 1) return bad("bad output shape");
    if (c.num_speakers > 1 && c.speaker_embedding_size < 2) return bad("speaker_embedding_size == 1 (get_embed tables) is not built");
    if (c.enc_proj_sizes[c.n_enc_proj - 1] != c.enc_prenet_sizes[c.n_enc_prenet - 1]) return bad("enc_proj_sizes[-1] must equal enc_prenet_sizes[-1]");
    if (c.post_proj_sizes[c.n_post_proj - 1] != c.num_mels) return bad("post_proj_sizes[-1] must equal num_mels");
    taco_handle *h = new taco_handle();
    h->cfg = c;
    *out = h;
    return TACO_OK;
}

void taco_destroy(taco_handle *h) {
    if (!h) return;
    for (void *p : h->allocs) cudaFree(p);
    if (h->ws) cudaFree(h->ws);
    if (h->probs_dev) cudaFree(h->probs_dev);
    if (h->dp_dev) cudaFree(h->dp_dev);
    if (h->prof_dev) cudaFree(h->prof_dev);
    if (h->barrier_dev) cudaFree(h->barrier_dev);
    if (h->ids_lens_dev) cudaFree(h->ids_lens_dev);
    if (h->host_io) cudaFree(h->host_io);
    for (int i = 0; i < 2; ++i) {
        if (h->stage[i]) cudaFreeHost(h->stage[i]);
        if (h->stage_ev[i]) cudaEventDestroy(h->stage_ev[i]);
    }
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->tc_err) cudaFree(h->tc_err);
    if (h->tc_dbg) cudaFree(h->tc_dbg);
    if (h->op_times_dev) cudaFree(h->op_times_dev);
    for (auto &kv : h->tc_w) { cudaFree(kv.second.hi); cudaFree(kv.second.lo); }
    delete h;
}

int taco_set_weight(taco_handle *h, const char *name, const float *data, int64_t n) {
    if (!h || !name || (!data && n > 0) || n < 0) return h ? fail(h, TACO_ERR_ARG, "taco_set_weight: bad argument") : TACO_ERR_ARG;
    if (h->finalized) return fail(h, TACO_ERR_STATE, "taco_set_weight after taco_finalize");
    h->w[name].assign(data, data + n);
    return TACO_OK;
}

int taco_finalize(taco_handle *h) {
    if (!h) return TACO_ERR_ARG;
    if (h->finalized) return fail(h, TACO_ERR_STATE, "taco_finalize called twice");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(h, TACO_ERR_CUDA, "no CUDA device (there is no CPU fallback)");
    CK(cudaGetDevice(&h->device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, h->device));
    if (prop.major < 10) return fail(h, TACO_ERR_CUDA, "libtaco_b200 is built for sm_100a only");
    h->sm_count = prop.multiProcessorCount;
    h->smem_optin = prop.sharedMemPerBlockOptin;
    const taco_config &c = h->cfg;

    auto emb = find_w(h, "embedding", (size_t)c.num_symbols * c.embedding_size);
    if (!emb) return TACO_ERR_STATE;
    h->embedding = upload(h, emb->data(), emb->size());
    if (c.num_speakers > 1) {
        auto se = find_w(h, "speaker_embedding", (size_t)c.num_speakers * c.speaker_embedding_size);
        if (!se) return TACO_ERR_STATE;
        h->spk_embedding = upload(h, se->data(), se->size());
        std::vector<int> widths = {c.enc_prenet_sizes[c.n_enc_prenet - 1], 2 * c.enc_rnn_size, c.attention_state_size};
        for (int i = 0; i < c.dec_layer_num; ++i) widths.push_back(c.dec_rnn_size);
        h->spk_dense.resize(widths.size());
        for (size_t i = 0; i < widths.size(); ++i) {
            const std::string s = i == 0 ? "dense" : "dense_" + std::to_string(i);
            if (!load_dense(h, s, c.speaker_embedding_size, widths[i], true, &h->spk_dense[i])) return TACO_ERR_STATE;
        }
    }
    h->enc_prenet.resize(c.n_enc_prenet);
    int ci = c.embedding_size;
    for (int i = 0; i < c.n_enc_prenet; ++i) {
        if (!load_dense(h, "prenet/dense_" + std::to_string(i + 1), ci, c.enc_prenet_sizes[i], true, &h->enc_prenet[i])) return TACO_ERR_STATE;
        ci = c.enc_prenet_sizes[i];
    }
    if (!load_cbhg(h, "encoder_cbhg", ci, c.enc_bank_size, c.enc_bank_channel_size, c.enc_proj_sizes, c.n_enc_proj, c.enc_proj_width,
                   c.enc_highway_depth, c.enc_rnn_size, &h->enc))
        return TACO_ERR_STATE;
    if (!load_dense(h, "memory_layer", 2 * c.enc_rnn_size, c.attention_size, false, &h->memory_layer)) return TACO_ERR_STATE;
    if (!load_cbhg(h, "post_cbhg", c.num_mels, c.post_bank_size, c.post_bank_channel_size, c.post_proj_sizes, c.n_post_proj, c.post_proj_width,
                   c.post_highway_depth, c.post_rnn_size, &h->post))
        return TACO_ERR_STATE;
    {
        const int nd = c.num_speakers > 1 ? 3 + c.dec_layer_num : 0;
        const std::string s = nd == 0 ? "dense" : "dense_" + std::to_string(nd);
        if (!load_dense(h, s, 2 * c.post_rnn_size, c.num_freq, true, &h->final_dense)) return TACO_ERR_STATE;
    }
    int rc = build_decoder(h);
    if (rc != TACO_OK) return rc;

    // persistent decoder launch shape: one CTA per SM, co-resident (cooperative launch)
    h->dec_grid = h->sm_count;
    const int smem = (int)dec_smem_plan(h, 1, 1024, 1, 1, 0).bytes;
    if (smem > (int)prop.sharedMemPerBlockOptin)
        return fail(h, TACO_ERR_ARG, "decoder weight slices + staging (" + std::to_string(smem) + " B) exceed the shared memory of one SM");
    cudaFuncAttributes fa;
    CK(cudaFuncGetAttributes(&fa, taco_decoder_kernel));
    const int dec_dyn_max = (int)prop.sharedMemPerBlockOptin - (int)fa.sharedSizeBytes;   // static + dynamic share the opt-in limit
    if (smem > dec_dyn_max)
        return fail(h, TACO_ERR_ARG, "decoder weight slices + staging (" + std::to_string(smem) + " B) exceed the shared memory of one SM");
    CK(cudaFuncSetAttribute(taco_decoder_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dec_dyn_max));
    h->dec_dyn_max = dec_dyn_max;
    CK(cudaMalloc(&h->barrier_dev, 2 * sizeof(unsigned)));
    const int Umax = std::max(c.enc_rnn_size, c.post_rnn_size);
    const size_t rnn_smem = ((size_t)3 * Umax + (size_t)Umax * 3 * Umax) * sizeof(float);
    h->rnn_w_in_smem = rnn_smem <= prop.sharedMemPerBlockOptin ? 1 : 0;
    CK(cudaFuncGetAttributes(&fa, taco_bigru_kernel));
    CK(cudaFuncSetAttribute(taco_bigru_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prop.sharedMemPerBlockOptin - (int)fa.sharedSizeBytes));
    h->smem_optin = prop.sharedMemPerBlockOptin - fa.sharedSizeBytes;
    {
        int per_sm = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, taco_decoder_kernel, DEC_THREADS, (size_t)smem));
        if (per_sm < 1) return fail(h, TACO_ERR_CUDA, "the persistent decoder kernel does not fit one SM");
    }
    CK(cudaMalloc(&h->dp_dev, sizeof(DecParams)));
    // tensor-core path of the CBHG convolutions: needs the driver's tensor-map encoder; TACO_NO_TC=1 keeps the fp32 SIMT GEMM (A/B runs)
    {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn || q != cudaDriverEntryPointSuccess)
            return fail(h, TACO_ERR_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
        h->tc_encode = fn;
        CK(cudaMalloc(&h->tc_err, sizeof(unsigned)));
        CK(cudaMemset(h->tc_err, 0, sizeof(unsigned)));
        CK(cudaFuncSetAttribute(tc::gemm_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::Layout<128>::SMEM_BYTES));
        CK(cudaFuncSetAttribute(tc::gemm_tc_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::Layout<256>::SMEM_BYTES));
        h->tc_on = getenv("TACO_NO_TC") == nullptr;
    }
    h->w.clear();
    h->finalized = true;
    return TACO_OK;
}

int taco_get_info(const taco_handle *h, taco_info *info) {
    if (!h || !info) return TACO_ERR_ARG;
    memset(info, 0, sizeof(*info));
    info->sm_count = h->sm_count;
    info->dec_grid = h->dec_grid;
    info->dec_threads = DEC_THREADS;
    info->dec_smem_bytes = h->dec_smem;
    info->dec_phases_per_step = h->dp_host.n_phases;
    info->rnn_weights_in_smem = h->rnn_w_in_smem;
    info->n_params = h->n_params;
    info->kernel_launches = h->launches;
    info->workspace_bytes = (int64_t)h->ws_cap;
    info->tc_gemm_launches = h->tc_launches;
    return TACO_OK;
}

int taco_synthesize(taco_handle *h, const taco_synth_args *a, void *stream_) {
    if (!h || !a) return TACO_ERR_ARG;
    if (!h->finalized) return fail(h, TACO_ERR_STATE, "taco_synthesize before taco_finalize");
    const taco_config &c = h->cfg;
    cudaStream_t st = (cudaStream_t)stream_;
    const int N = a->N, T_in = a->T_in;
    const int S = a->n_steps > 0 ? a->n_steps : c.max_iters;
    if (N < 1 || T_in < 1 || T_in > 1024 || !a->ids_dev || !a->lengths || !a->mel_dev || !a->alignments_dev)
        return fail(h, TACO_ERR_ARG, "taco_synthesize: bad argument (N >= 1, 1 <= T_in <= 1024, ids/lengths/mel/alignments required)");
    for (int i = 0; i < N; ++i) {
        if (a->lengths[i] < 1 || a->lengths[i] > T_in) return fail(h, TACO_ERR_ARG, "taco_synthesize: lengths must be in [1, T_in]");
        if (a->speaker_ids && (a->speaker_ids[i] < 0 || a->speaker_ids[i] >= std::max(1, c.num_speakers)))
            return fail(h, TACO_ERR_ARG, "taco_synthesize: speaker id out of range");
    }
    const int r = c.reduction_factor, nm = c.num_mels, Tm = S * r;
    const int mem = 2 * c.enc_rnn_size, A = c.attention_size, H = c.attention_state_size, R = c.dec_rnn_size;
    const int tiles = (N + 31) / 32;
    const int G = h->dec_grid;
    const int chunks = std::max(1, std::min(std::min(G / N, 8), std::max(1, T_in / 8)));
    const int fslices = std::max(1, std::min(std::min(G / N, 8), std::max(1, mem / 32)));
    const DecSmem sp = dec_smem_plan(h, N, T_in, chunks, fslices, (size_t)h->dec_dyn_max);
    const int smem = (int)sp.bytes;
    if (smem > h->dec_dyn_max) return fail(h, TACO_ERR_ARG, "taco_synthesize: T_in too large for the decoder's shared-memory scratch");
    h->dec_smem = smem;

    // ---- workspace plan (bump allocator, two passes) ----
    h->taps.clear();
    char *base = nullptr;
    size_t off = 0;
    auto alloc = [&](size_t nfloats, const char *tap = nullptr) -> float * {
        float *p = base ? reinterpret_cast<float *>(base + off) : nullptr;
        if (base && tap) h->taps[tap] = DevBuf{p, nfloats};
        off += align_up(std::max<size_t>(nfloats, 1) * sizeof(float), 256);
        return p;
    };
    const size_t MT = (size_t)N * T_in, MP = (size_t)N * Tm;
    struct WS {
        float *emb, *pre[2], *spk, *spk_out[3 + TACO_MAX_DEC_LAYERS];
        float *e_bank, *e_proj[2], *e_hw[3], *e_xp, *memory, *keys;
        float *db[DB_COUNT], *score, *state[2], *q_row;
        float *p_bank, *p_proj[2], *p_hw[3], *p_xp, *p_out;
        float *tc_hi, *tc_lo;
    } w;
    int enc_pre_max = c.embedding_size;
    for (int i = 0; i < c.n_enc_prenet; ++i) enc_pre_max = std::max(enc_pre_max, c.enc_prenet_sizes[i]);
    int e_proj_max = 1, p_proj_max = 1;
    for (int i = 0; i < c.n_enc_proj; ++i) e_proj_max = std::max(e_proj_max, c.enc_proj_sizes[i]);
    for (int i = 0; i < c.n_post_proj; ++i) p_proj_max = std::max(p_proj_max, c.post_proj_sizes[i]);
    int db_width[DB_COUNT];
    memset(db_width, 0, sizeof(db_width));
    db_width[DB_X] = nm;
    for (int i = 0; i < c.n_dec_prenet; ++i) db_width[DB_P0 + i] = c.dec_prenet_sizes[i];
    db_width[DB_CTX] = mem; db_width[DB_HATT] = H; db_width[DB_RH] = std::max(H, R); db_width[DB_U] = std::max(H, R); db_width[DB_Q] = A;
    for (int i = 0; i <= c.dec_layer_num; ++i) db_width[DB_O0 + i] = R;
    for (int i = 0; i < c.dec_layer_num; ++i) db_width[DB_H1 + i] = R;
    auto rup = [](int x, int a) { return (x + a - 1) / a * a; };
    // largest split operand of the tensor-core path: rows x channels rounded up to 32 (bank input, pooled bank, projection inputs)
    size_t tc_elems = 0;
    if (h->tc_on) {
        auto upd = [&](size_t rows, int ci) { tc_elems = std::max(tc_elems, rows * (size_t)rup(ci, 32)); };
        upd(MT, h->enc.n_in); upd(MT, h->enc.K * h->enc.C);
        for (auto &L : h->enc.proj) upd(MT, L.co);
        upd(MT, h->enc.U);
        if (a->linear_dev) {
            upd(MP, h->post.n_in); upd(MP, h->post.K * h->post.C);
            for (auto &L : h->post.proj) upd(MP, L.co);
            upd(MP, h->post.U);
            upd(MP, 2 * c.post_rnn_size);
        }
    }
    auto plan = [&]() {
        off = 0;
        w.tc_hi = tc_elems ? alloc(tc_elems) : nullptr;
        w.tc_lo = tc_elems ? alloc(tc_elems) : nullptr;
        w.emb = alloc(MT * enc_pre_max);
        w.pre[0] = alloc(MT * enc_pre_max);
        w.pre[1] = alloc(MT * enc_pre_max, "enc_prenet");
        w.spk = alloc((size_t)N * std::max(1, c.speaker_embedding_size));
        for (size_t i = 0; i < h->spk_dense.size(); ++i) w.spk_out[i] = alloc((size_t)N * h->spk_dense[i].co);
        w.e_bank = alloc(MT * c.enc_bank_size * c.enc_bank_channel_size, "enc_bank");
        w.e_proj[0] = alloc(MT * e_proj_max);
        w.e_proj[1] = alloc(MT * e_proj_max);
        w.e_hw[0] = alloc(MT * c.enc_rnn_size);
        w.e_hw[1] = alloc(MT * c.enc_rnn_size);
        w.e_hw[2] = h->enc.has_dense ? alloc(MT * c.enc_rnn_size) : nullptr;
        w.e_xp = alloc(MT * 6 * c.enc_rnn_size);
        w.memory = alloc(MT * mem, "encoder_out");
        w.keys = alloc(MT * A, "keys");
        for (int b = 0; b < DB_COUNT; ++b) w.db[b] = db_width[b] ? alloc((size_t)tiles * db_width[b] * 32) : nullptr;
        w.q_row = alloc((size_t)tiles * 32 * A);
        w.score = alloc(MT);
        w.state[0] = alloc(MT);
        w.state[1] = alloc(MT);
        if (a->linear_dev) {
            w.p_bank = alloc(MP * c.post_bank_size * c.post_bank_channel_size, "post_bank");
            w.p_proj[0] = alloc(MP * p_proj_max);
            w.p_proj[1] = alloc(MP * p_proj_max);
            w.p_hw[0] = alloc(MP * c.post_rnn_size);
            w.p_hw[1] = alloc(MP * c.post_rnn_size);
            w.p_hw[2] = h->post.has_dense ? alloc(MP * c.post_rnn_size) : nullptr;
            w.p_xp = alloc(MP * 6 * c.post_rnn_size);
            w.p_out = alloc(MP * 2 * c.post_rnn_size, "post_out");
        }
    };
    plan();
    if (off > h->ws_cap) {
        CK(cudaStreamSynchronize(st));
        if (h->ws) cudaFree(h->ws);
        h->ws = nullptr;
        h->ws_cap = 0;
        CK(cudaMalloc(&h->ws, off));
        h->ws_cap = off;
    }
    base = h->ws;
    plan();

    // lengths / speaker ids on the device
    if ((size_t)2 * N > h->ids_lens_cap) {
        CK(cudaStreamSynchronize(st));
        if (h->ids_lens_dev) cudaFree(h->ids_lens_dev);
        CK(cudaMalloc(&h->ids_lens_dev, (size_t)2 * N * sizeof(int32_t)));
        h->ids_lens_cap = (size_t)2 * N;
    }
    std::vector<int32_t> il(2 * N, 0);
    for (int i = 0; i < N; ++i) { il[i] = a->lengths[i]; il[N + i] = a->speaker_ids ? a->speaker_ids[i] : 0; }
    CK(cudaMemcpyAsync(h->ids_lens_dev, il.data(), il.size() * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    const int32_t *lens_dev = h->ids_lens_dev, *spk_dev = h->ids_lens_dev + N;

    // ---- op list ----
    std::vector<GemmProb> probs;
    struct GemmLaunch { int first, count, B, T, maxN; };
    std::vector<std::function<int()>> ops;
    struct OpMeta { float kind, flops, M, N, K; };      // kind: 0 other, 1 fp32 SIMT GEMM, 2 tensor-core GEMM (split + MMA), 3 recurrent, 4 decoder
    std::vector<OpMeta> metas;
    auto gemm_group = [&](std::vector<GemmProb> group, int B, int T) {
        GemmLaunch L{(int)probs.size(), (int)group.size(), B, T, 0};
        double fl = 0, ksum = 0;
        for (auto &g : group) { L.maxN = std::max(L.maxN, g.N); probs.push_back(g); fl += 2.0 * B * T * (double)g.N * g.ktaps * g.Ci; ksum += (double)g.ktaps * g.Ci; }
        metas.push_back(OpMeta{1.f, (float)fl, (float)((double)B * T), (float)L.maxN, (float)ksum});
        ops.push_back([h, L, st]() -> int {
            dim3 grid((L.maxN + GBN - 1) / GBN, (unsigned)(((size_t)L.B * L.T + GBM - 1) / GBM), L.count);
            taco_gemm_kernel<<<grid, GTHREADS, 0, st>>>(h->probs_dev + L.first, L.B, L.T);
            h->launches++;
            return cudaGetLastError() == cudaSuccess ? 0 : -1;
        });
    };
    auto mk = [&](const ConvLayer &L, const float *Ain, int lda, float *Cout, int ldc, int act, bool pool = false) {
        GemmProb g;
        memset(&g, 0, sizeof(g));
        g.A = Ain; g.lda = lda; g.W = L.W; g.bias = L.b; g.bn_scale = L.scale; g.bn_shift = L.shift;
        g.C = Cout; g.ldc = ldc; g.Ci = L.ci; g.ktaps = L.k; g.pl = (L.k - 1) / 2; g.pool = pool ? 1 : 0;
        g.N = L.co; g.act = act; g.epi = EPI_LINEAR;
        return g;
    };
    // One launch group on the tensor cores: all problems read the same input (Ain, lda, channels Ci) and are at most 256 wide.
    // Plan time: split + transpose the weights once per handle, encode the four tensor maps; run time: split the input, one GEMM.
    typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                      const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    std::string tc_fail;
    int tc_ord = 0;
    if (getenv("TACO_TC_DEBUG") && !h->tc_dbg) {
        CK(cudaMalloc(&h->tc_dbg, 32 * 16 * sizeof(long long)));
        CK(cudaMemset(h->tc_dbg, 0, 32 * 16 * sizeof(long long)));
    }
    if (h->tc_dbg) h->taps["tc_dbg"] = DevBuf{reinterpret_cast<float *>(h->tc_dbg), 32 * 16 * 2};
    auto tc_group = [&](const std::vector<GemmProb> &group, int B, int T) -> bool {
        const int Ci = group[0].Ci, Cip = rup(Ci, 32), np_ = (int)group.size();
        int maxN = 0, maxK = 0;
        for (auto &g : group) { maxN = std::max(maxN, g.N); maxK = std::max(maxK, g.ktaps); }
        const int NT = maxN <= 128 ? 128 : 256;
        const int rows_per = rup(maxN, 16), rows_total = rows_per * np_ + NT, Kp = maxK * Cip;   // + NT: the last tile may overhang
        auto &tw = h->tc_w[group[0].W];
        if (!tw.hi) {
            const size_t n = (size_t)rows_total * Kp;
            if (cudaMalloc(&tw.hi, n * sizeof(float)) != cudaSuccess || cudaMalloc(&tw.lo, n * sizeof(float)) != cudaSuccess) { tc_fail = "cudaMalloc of split weights"; return false; }
            cudaMemsetAsync(tw.hi, 0, n * sizeof(float), st);
            cudaMemsetAsync(tw.lo, 0, n * sizeof(float), st);
            for (int i = 0; i < np_; ++i) {
                const long tot = (long)group[i].N * group[i].ktaps * Ci;
                tc::split_weight_kernel<<<(unsigned)std::min<long>((tot + 255) / 256, 2048), 256, 0, st>>>(group[i].W, group[i].ktaps, Ci, Cip, group[i].N, i * rows_per, Kp,
                                                                                                          tw.hi, tw.lo);
            }
            tw.rows = rows_total; tw.Kp = Kp; tw.Cip = Cip;
            h->launches += np_;
        }
        EncodeTiledFn enc = (EncodeTiledFn)h->tc_encode;
        CUtensorMap m_ahi, m_alo, m_bhi, m_blo;
        {
            cuuint64_t dims[3] = {(cuuint64_t)Cip, (cuuint64_t)T, (cuuint64_t)B};
            cuuint64_t strides[2] = {(cuuint64_t)Cip * 4, (cuuint64_t)T * Cip * 4};
            cuuint32_t box[3] = {(cuuint32_t)tc::TK, (cuuint32_t)tc::TM, 1}, es[3] = {1, 1, 1};
            CUresult r1 = enc(&m_ahi, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, w.tc_hi, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            CUresult r2 = enc(&m_alo, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, w.tc_lo, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            cuuint64_t bd[2] = {(cuuint64_t)Kp, (cuuint64_t)rows_total}, bs[1] = {(cuuint64_t)Kp * 4};
            cuuint32_t bb[2] = {(cuuint32_t)tc::TK, (cuuint32_t)NT}, be[2] = {1, 1};
            CUresult r3 = enc(&m_bhi, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, tw.hi, bd, bs, bb, be, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            CUresult r4 = enc(&m_blo, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, tw.lo, bd, bs, bb, be, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r1 != CUDA_SUCCESS || r2 != CUDA_SUCCESS || r3 != CUDA_SUCCESS || r4 != CUDA_SUCCESS) {
                tc_fail = "cuTensorMapEncodeTiled failed (" + std::to_string((int)r1) + "," + std::to_string((int)r2) + "," + std::to_string((int)r3) + "," + std::to_string((int)r4) + ")";
                return false;
            }
        }
        tc::Args ta;
        memset(&ta, 0, sizeof(ta));
        for (int i = 0; i < np_; ++i) {
            const GemmProb &g = group[i];
            tc::Prob &q = ta.p[i];
            q.bias = g.bias; q.bn_scale = g.bn_scale; q.bn_shift = g.bn_shift; q.R = g.R; q.rowvec = g.rowvec; q.C = g.C;
            q.ldc = g.ldc; q.ldr = g.ldr; q.ldrv = g.ldrv; q.ktaps = g.ktaps; q.pl = g.pl; q.N = g.N; q.act = g.act; q.epi = g.epi == EPI_HIGHWAY ? 1 : 0; q.b_row0 = i * rows_per;
        }
        ta.T = T; ta.B = B; ta.Cip = Cip; ta.err = h->tc_err; ta.dbg = nullptr;
        if (h->tc_dbg) { ta.dbg = h->tc_dbg + 16 * std::min(tc_ord, 31); ++tc_ord; }
        const GemmProb g0 = group[0];
        float *hi = w.tc_hi, *lo = w.tc_lo;
        {
            double fl = 0, ksum = 0;
            for (auto &g : group) { fl += 2.0 * B * T * (double)g.N * g.ktaps * g.Ci; ksum += (double)g.ktaps * g.Ci; }
            metas.push_back(OpMeta{2.f, (float)fl, (float)((double)B * T), (float)maxN, (float)ksum});
        }
        ops.push_back([=]() -> int {
            const long rows = (long)B * T, tot = rows * (Cip / 4);
            tc::split_act_kernel<<<(unsigned)std::min<long>((tot + 255) / 256, 8192), 256, 0, st>>>(g0.A, g0.lda, Ci, Cip, T, rows, g0.pool, hi, lo);
            dim3 grid((unsigned)(B * ((T + tc::TM - 1) / tc::TM)), (unsigned)((maxN + NT - 1) / NT), (unsigned)np_);
            if (NT == 128) tc::gemm_tc_kernel<128><<<grid, tc::THREADS, tc::Layout<128>::SMEM_BYTES, st>>>(m_ahi, m_alo, m_bhi, m_blo, ta);
            else tc::gemm_tc_kernel<256><<<grid, tc::THREADS, tc::Layout<256>::SMEM_BYTES, st>>>(m_ahi, m_alo, m_bhi, m_blo, ta);
            h->launches += 2;
            h->tc_launches += 1;
            return cudaGetLastError() == cudaSuccess ? 0 : -1;
        });
        return true;
    };
    // tensor cores when the group qualifies (plain conv / dense epilogue, 16 <= N, at most 16 problems), the fp32 SIMT GEMM otherwise
    auto conv_group = [&](std::vector<GemmProb> group, int B, int T) -> bool {
        // small problems stay on the SIMT kernel (TACO_TC_MIN_ROWS overrides the threshold: the parity tests push tiny cases through the tensor cores)
        const char *mr = getenv("TACO_TC_MIN_ROWS");
        bool ok = h->tc_on && group.size() <= (size_t)tc::MAX_PROBS && (size_t)B * T >= (size_t)(mr ? atoi(mr) : 256);
        for (auto &g : group) ok = ok && (g.epi == EPI_LINEAR || (g.epi == EPI_HIGHWAY && (g.N & 3) == 0 && (g.ldr & 1) == 0 && (g.ldc & 1) == 0)) && g.N >= 16 && g.Ci == group[0].Ci && g.A == group[0].A && g.lda == group[0].lda && g.pool == group[0].pool;
        if (!ok) { gemm_group(group, B, T); return true; }
        return tc_group(group, B, T);
    };
    auto cbhg = [&](const CbhgDev &D, const float *x, int B, int T, const float *before_highway, int ld_bh, const float *rnn_init,
                    const int32_t *lens, float *bank, float *proj[2], float *hw[3], float *xp, float *out, const char *tag) -> bool {
        const size_t MM = (size_t)B * T;
        std::vector<GemmProb> grp;
        const int ldb = D.K * D.C;
        for (int k = 0; k < D.K; ++k) grp.push_back(mk(D.bank[k], x, D.n_in, bank + (size_t)k * D.C, ldb, ACT_RELU));
        if (!conv_group(grp, B, T)) return false;
        const float *cur = bank;
        int ld = ldb;
        for (size_t i = 0; i < D.proj.size(); ++i) {
            const bool last = i + 1 == D.proj.size();
            GemmProb g = mk(D.proj[i], cur, ld, proj[i & 1], D.proj[i].co, last ? ACT_NONE : ACT_RELU, i == 0);
            if (last) {   // highway_input = proj_out + inputs (+ before_highway)   modules.py:47-53
                g.R = x; g.ldr = D.n_in;
                g.rowvec = before_highway; g.ldrv = ld_bh;
            }
            if (!conv_group({g}, B, T)) return false;
            cur = proj[i & 1];
            ld = D.proj[i].co;
        }
        float *hcur = hw[2];
        if (D.has_dense) {   // hw[2] is not reused by the highway ping-pong, so the debug tap stays valid
            if (!conv_group({mk(D.dense, cur, ld, hw[2], D.U, ACT_NONE)}, B, T)) return false;
        } else {
            // widths match: copy through a 1-tap identity is wasteful; read the projection output directly
            hcur = const_cast<float *>(cur);
        }
        h->taps[std::string(tag) + "_highway_in"] = DevBuf{hcur, MM * D.U};
        for (size_t i = 0; i < D.highway.size(); ++i) {
            float *dst = (hcur == hw[0]) ? hw[1] : hw[0];
            GemmProb g = mk(D.highway[i], hcur, D.U, dst, D.U, ACT_NONE);
            g.epi = EPI_HIGHWAY;
            g.R = hcur; g.ldr = D.U;
            if (!conv_group({g}, B, T)) return false;
            hcur = dst;
        }
        h->taps[std::string(tag) + "_rnn_in"] = DevBuf{hcur, MM * D.U};
        ConvLayer xl;
        xl.W = D.rnn.Wx; xl.b = D.rnn.bx; xl.k = 1; xl.ci = D.U; xl.co = 6 * D.U;
        if (!conv_group({mk(xl, hcur, D.U, xp, 6 * D.U, ACT_NONE)}, B, T)) return false;
        RnnParams rp;
        memset(&rp, 0, sizeof(rp));
        rp.XP = xp; rp.Wgh[0] = D.rnn.Wgh[0]; rp.Wgh[1] = D.rnn.Wgh[1]; rp.Wch[0] = D.rnn.Wch[0]; rp.Wch[1] = D.rnn.Wch[1];
        rp.init = rnn_init; rp.lengths = lens; rp.out = out; rp.N = B; rp.T = T; rp.U = D.U;
        const size_t need = ((size_t)3 * D.U + (size_t)D.U * 3 * D.U) * sizeof(float);
        rp.w_in_smem = need <= h->smem_optin ? 1 : 0;
        const size_t sm = rp.w_in_smem ? need : (size_t)3 * D.U * sizeof(float);
        const int grid = std::min(2 * B, 2 * h->sm_count);
        metas.push_back(OpMeta{3.f, 0.f, (float)((double)B * T), (float)D.U, 0.f});
        ops.push_back([h, rp, sm, grid, st]() -> int {
            if (rp.U == 128) taco_bigru_reg_kernel<128><<<grid, 256, 0, st>>>(rp);     // recurrent weights in registers
            else taco_bigru_kernel<<<grid, 256, sm, st>>>(rp);
            h->launches++;
            return cudaGetLastError() == cudaSuccess ? 0 : -1;
        });
        return true;
    };

    // embedding + speaker states
    metas.push_back(OpMeta{0.f, 0.f, 0.f, 0.f, 0.f});
    ops.push_back([=]() -> int {
        const size_t tot = MT * c.embedding_size;
        taco_embed_kernel<<<(unsigned)std::min<size_t>((tot + 255) / 256, 4096), 256, 0, st>>>(a->ids_dev, h->embedding, (int)MT, c.embedding_size,
                                                                                               c.num_symbols, 1, w.emb);
        h->launches++;
        if (c.num_speakers > 1) {
            taco_embed_kernel<<<1, 256, 0, st>>>(spk_dev, h->spk_embedding, N, c.speaker_embedding_size, c.num_speakers, 0, w.spk);
            h->launches++;
        }
        return cudaGetLastError() == cudaSuccess ? 0 : -1;
    });
    const float *before_highway = nullptr, *enc_init = nullptr, *att_init = nullptr, *dec_init[TACO_MAX_DEC_LAYERS] = {nullptr, nullptr, nullptr, nullptr};
    if (c.num_speakers > 1) {
        std::vector<GemmProb> grp;
        for (size_t i = 0; i < h->spk_dense.size(); ++i) grp.push_back(mk(h->spk_dense[i], w.spk, c.speaker_embedding_size, w.spk_out[i], h->spk_dense[i].co, ACT_SOFTSIGN));
        gemm_group(grp, N, 1);
        before_highway = w.spk_out[0]; enc_init = w.spk_out[1]; att_init = w.spk_out[2];
        for (int i = 0; i < c.dec_layer_num; ++i) dec_init[i] = w.spk_out[3 + i];
    }
    // encoder prenet
    const float *cur = w.emb;
    int ld = c.embedding_size;
    for (int i = 0; i < c.n_enc_prenet; ++i) {
        float *dst = (i == c.n_enc_prenet - 1) ? w.pre[1] : w.pre[0];
        if (dst == cur) dst = w.emb;
        gemm_group({mk(h->enc_prenet[i], cur, ld, dst, h->enc_prenet[i].co, ACT_RELU)}, N, T_in);
        cur = dst;
        ld = h->enc_prenet[i].co;
    }
    h->taps["enc_prenet"] = DevBuf{const_cast<float *>(cur), MT * ld};
    if (!cbhg(h->enc, cur, N, T_in, before_highway, before_highway ? h->spk_dense[0].co : 0, enc_init, lens_dev, w.e_bank, w.e_proj, w.e_hw, w.e_xp, w.memory, "enc"))
        return fail(h, TACO_ERR_CUDA, "taco_synthesize: tensor-core plan failed: " + tc_fail);
    gemm_group({mk(h->memory_layer, w.memory, mem, w.keys, A, ACT_NONE)}, N, T_in);

    // decoder
    DecParams dp = h->dp_host;
    dp.N = N; dp.tiles = tiles; dp.T_in = T_in; dp.n_steps = S; dp.chunks = chunks; dp.fslices = fslices;
    for (int b = 0; b < DB_COUNT; ++b) dp.buf[b] = w.db[b];
    dp.keys = w.keys; dp.values = w.memory; dp.lengths = lens_dev; dp.score = w.score; dp.state[0] = w.state[0]; dp.state[1] = w.state[1];
    dp.manual = a->manual_alignments_dev; dp.dec_out = a->mel_dev; dp.align = a->alignments_dev;
    dp.q_row = w.q_row; dp.barrier = h->barrier_dev;
    dp.stage_floats = sp.stage_floats; dp.keys_res = sp.keys_floats > 0; dp.vals_res = sp.vals_floats > 0;
    dp.prof = nullptr;
    if (getenv("TACO_PROFILE")) {
        if (!h->prof_dev) CK(cudaMalloc(&h->prof_dev, (size_t)h->dec_grid * 8 * sizeof(long long)));
        dp.prof = h->prof_dev;
        h->taps["dec_prof"] = DevBuf{reinterpret_cast<float *>(h->prof_dev), (size_t)h->dec_grid * 16};
    }
    DecInit di;
    memset(&di, 0, sizeof(di));
    di.n = 0;
    auto add_init = [&](int b, const float *src, const float *vec = nullptr) {
        di.dst[di.n] = w.db[b]; di.src[di.n] = src; di.vec[di.n] = vec; di.width[di.n] = db_width[b]; di.n++;
    };
    add_init(DB_X, nullptr);
    if (h->dec_p0_init) add_init(DB_P0, nullptr, h->dec_p0_init);
    add_init(DB_CTX, nullptr);
    add_init(DB_HATT, att_init);
    for (int i = 0; i < c.dec_layer_num; ++i) add_init(DB_H1 + i, dec_init[i]);
    di.state0 = w.state[0]; di.N = N; di.tiles = tiles; di.T_in = T_in; di.dirac = c.attention_type != TACO_ATT_LOC_SEN;
    metas.push_back(OpMeta{4.f, 0.f, (float)N, (float)S, 0.f});
    ops.push_back([h, dp, di, st, smem, G]() -> int {
        if (cudaMemcpyAsync(h->dp_dev, &dp, sizeof(dp), cudaMemcpyHostToDevice, st) != cudaSuccess) return -1;
        if (cudaMemsetAsync(h->barrier_dev, 0, 2 * sizeof(unsigned), st) != cudaSuccess) return -1;
        taco_dec_init_kernel<<<64, 256, 0, st>>>(di);
        h->launches++;
        const DecParams *arg = h->dp_dev;
        void *params[] = {(void *)&arg};
        if (cudaLaunchCooperativeKernel((const void *)taco_decoder_kernel, dim3(G), dim3(DEC_THREADS), params, (size_t)smem, st) != cudaSuccess) return -1;
        h->launches++;
        return 0;
    });

    // post-processing net
    if (a->linear_dev) {
        if (!cbhg(h->post, a->mel_dev, N, Tm, nullptr, 0, nullptr, nullptr, w.p_bank, w.p_proj, w.p_hw, w.p_xp, w.p_out, "post") ||
            !conv_group({mk(h->final_dense, w.p_out, 2 * c.post_rnn_size, a->linear_dev, c.num_freq, ACT_NONE)}, N, Tm))
            return fail(h, TACO_ERR_CUDA, "taco_synthesize: tensor-core plan failed: " + tc_fail);
    }

    // upload the problem table, then run
    if (probs.size() > h->probs_cap) {
        CK(cudaStreamSynchronize(st));
        if (h->probs_dev) cudaFree(h->probs_dev);
        CK(cudaMalloc(&h->probs_dev, probs.size() * sizeof(GemmProb)));
        h->probs_cap = probs.size();
    }
    CK(cudaMemcpyAsync(h->probs_dev, probs.data(), probs.size() * sizeof(GemmProb), cudaMemcpyHostToDevice, st));
    // TACO_TIME_OPS=1: CUDA events around every op of the list; taco_debug_get("op_times") returns rows of
    // (ms, kind, useful flops, M, N, sum of K) -- scripts/bench_taco.py builds its per-op table and roofline block from it
    const bool time_ops = getenv("TACO_TIME_OPS") != nullptr && metas.size() == ops.size();
    std::vector<cudaEvent_t> ev;
    if (time_ops) {
        ev.resize(ops.size() + 1);
        for (auto &e : ev) CK(cudaEventCreate(&e));
        CK(cudaEventRecord(ev[0], st));
    }
    for (size_t i = 0; i < ops.size(); ++i) {
        if (ops[i]() != 0) {
            h->err = std::string("taco_synthesize: launch failed: ") + cudaGetErrorString(cudaGetLastError());
            return TACO_ERR_CUDA;
        }
        if (time_ops) CK(cudaEventRecord(ev[i + 1], st));
    }
    if (time_ops) {
        CK(cudaStreamSynchronize(st));
        std::vector<float> rows(ops.size() * 6);
        for (size_t i = 0; i < ops.size(); ++i) {
            float ms = 0.f;
            CK(cudaEventElapsedTime(&ms, ev[i], ev[i + 1]));
            rows[i * 6 + 0] = ms; rows[i * 6 + 1] = metas[i].kind; rows[i * 6 + 2] = metas[i].flops; rows[i * 6 + 3] = metas[i].M; rows[i * 6 + 4] = metas[i].N; rows[i * 6 + 5] = metas[i].K;
        }
        for (auto &e : ev) cudaEventDestroy(e);
        if (h->op_times_cap < rows.size()) {
            if (h->op_times_dev) cudaFree(h->op_times_dev);
            CK(cudaMalloc(&h->op_times_dev, rows.size() * sizeof(float)));
            h->op_times_cap = rows.size();
        }
        CK(cudaMemcpy(h->op_times_dev, rows.data(), rows.size() * sizeof(float), cudaMemcpyHostToDevice));
        h->taps["op_times"] = DevBuf{h->op_times_dev, rows.size()};
    }
    return TACO_OK;
}

namespace {
constexpr int kCopyThreads = 4;

// Host copy of one staged chunk, spread over a few threads (one core moves ~10 GB/s, the link 50+).
void scatter_copy(char *dst, const char *src, size_t bytes) {
    if (bytes < ((size_t)1 << 20)) { std::memcpy(dst, src, bytes); return; }
    const size_t per = ((bytes + kCopyThreads - 1) / kCopyThreads + 63) & ~(size_t)63;
    std::thread th[kCopyThreads - 1];
    int n = 0;
    for (int i = 1; i < kCopyThreads; ++i) {
        const size_t o = per * i;
        if (o >= bytes) break;
        th[n++] = std::thread([=] { std::memcpy(dst + o, src + o, std::min(per, bytes - o)); });
    }
    std::memcpy(dst, src, std::min(per, bytes));
    for (int i = 0; i < n; ++i) th[i].join();
}

// Device -> pageable host through the pinned two-slot ring: the copy of chunk i + 1 runs while chunk i is moved to the
// caller's buffer.  (A plain cudaMemcpy into pageable memory staged 150 MB at 1.2-3.5 GB/s.)
cudaError_t d2h_pipelined(taco_handle *h, void *dst_, const void *src_, size_t bytes) {
    char *dst = (char *)dst_;
    const char *src = (const char *)src_;
    const size_t kStageBytes = h->stage_bytes;
    const size_t nchunks = (bytes + kStageBytes - 1) / kStageBytes;
    auto issue = [&](size_t i) {
        const size_t o = i * kStageBytes, n = std::min(kStageBytes, bytes - o);
        cudaError_t e = cudaMemcpyAsync(h->stage[i & 1], src + o, n, cudaMemcpyDeviceToHost, h->copy_stream);
        return e != cudaSuccess ? e : cudaEventRecord(h->stage_ev[i & 1], h->copy_stream);
    };
    cudaError_t e = nchunks ? issue(0) : cudaSuccess;
    for (size_t i = 0; i < nchunks && e == cudaSuccess; ++i) {
        if (i + 1 < nchunks) e = issue(i + 1);          // slot (i + 1) & 1 was emptied in iteration i - 1
        if (e == cudaSuccess) e = cudaEventSynchronize(h->stage_ev[i & 1]);
        if (e == cudaSuccess) scatter_copy(dst + i * kStageBytes, h->stage[i & 1], std::min(kStageBytes, bytes - i * kStageBytes));
    }
    if (e != cudaSuccess) cudaStreamSynchronize(h->copy_stream);
    return e;
}
}  // namespace

int taco_synthesize_host(taco_handle *h, const taco_synth_args *a) {
    if (!h || !a) return TACO_ERR_ARG;
    if (!h->finalized) return fail(h, TACO_ERR_STATE, "taco_synthesize_host before taco_finalize");
    const taco_config &c = h->cfg;
    const int S = a->n_steps > 0 ? a->n_steps : c.max_iters;
    const size_t n_ids = (size_t)a->N * a->T_in, n_mel = (size_t)a->N * S * c.reduction_factor * c.num_mels,
                 n_lin = a->linear_dev ? (size_t)a->N * S * c.reduction_factor * c.num_freq : 0, n_al = (size_t)a->N * a->T_in * S,
                 n_man = a->manual_alignments_dev ? n_al : 0;
    auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t o_ids = 0, o_mel = o_ids + up(n_ids * sizeof(int32_t)), o_lin = o_mel + up(n_mel * sizeof(float)),
                 o_al = o_lin + up(n_lin * sizeof(float)), o_man = o_al + up(n_al * sizeof(float)), total = o_man + up(n_man * sizeof(float));
    if (total > h->host_io_cap) {
        if (h->host_io) CK(cudaFree(h->host_io));
        h->host_io = nullptr;
        h->host_io_cap = 0;
        CK(cudaMalloc(&h->host_io, total));
        h->host_io_cap = total;
    }
    if (!h->copy_stream) {
        CK(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
        if (const char *e = getenv("TACO_STAGE_BYTES")) h->stage_bytes = std::max<size_t>((size_t)atoll(e) & ~(size_t)255, 4096);
        for (int i = 0; i < 2; ++i) {
            CK(cudaHostAlloc(&h->stage[i], h->stage_bytes, cudaHostAllocDefault));
            CK(cudaEventCreateWithFlags(&h->stage_ev[i], cudaEventDisableTiming));
        }
    }
    int32_t *ids = (int32_t *)(h->host_io + o_ids);
    float *mel = (float *)(h->host_io + o_mel), *lin = n_lin ? (float *)(h->host_io + o_lin) : nullptr, *al = (float *)(h->host_io + o_al),
          *man = n_man ? (float *)(h->host_io + o_man) : nullptr;
    CK(cudaMemcpy(ids, a->ids_dev, n_ids * sizeof(int32_t), cudaMemcpyHostToDevice));
    if (n_man) CK(cudaMemcpy(man, a->manual_alignments_dev, n_man * sizeof(float), cudaMemcpyHostToDevice));
    taco_synth_args d = *a;
    d.ids_dev = ids; d.mel_dev = mel; d.linear_dev = lin; d.alignments_dev = al; d.manual_alignments_dev = man;
    int rc = taco_synthesize(h, &d, nullptr);
    if (rc != TACO_OK) return rc;
    rc = taco_sync_check(h, nullptr);                  // synchronises: the copy stream below starts after the kernels
    if (rc != TACO_OK) return rc;
    CK(d2h_pipelined(h, a->mel_dev, mel, n_mel * sizeof(float)));
    if (n_lin) CK(d2h_pipelined(h, a->linear_dev, lin, n_lin * sizeof(float)));
    CK(d2h_pipelined(h, a->alignments_dev, al, n_al * sizeof(float)));
    return TACO_OK;
}

int taco_sync_check(taco_handle *h, void *stream) {
    if (!h) return TACO_ERR_ARG;
    CK(cudaStreamSynchronize((cudaStream_t)stream));
    if (!h->barrier_dev) return TACO_OK;
    unsigned flags[2] = {0, 0};
    CK(cudaMemcpy(flags, h->barrier_dev, sizeof(flags), cudaMemcpyDeviceToHost));
    if (flags[1]) return fail(h, TACO_ERR_TIMEOUT, "the persistent decoder kernel aborted on its barrier watchdog");
    if (h->tc_err) {
        unsigned e = 0;
        CK(cudaMemcpy(&e, h->tc_err, sizeof(e), cudaMemcpyDeviceToHost));
        if (e) return fail(h, TACO_ERR_TIMEOUT, "a tensor-core GEMM timed out on a pipeline barrier");
    }
    return TACO_OK;
}

int64_t taco_debug_get(taco_handle *h, const char *name, float *host_out, int64_t n) {
    if (!h || !name) return -1;
    auto it = h->taps.find(name);
    if (it == h->taps.end()) return -1;
    if (cudaDeviceSynchronize() != cudaSuccess) return -1;
    const size_t m = std::min<size_t>((size_t)std::max<int64_t>(n, 0), it->second.n);
    if (m && host_out && cudaMemcpy(host_out, it->second.p, m * sizeof(float), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return (int64_t)it->second.n;
}

}  // extern "C"

// wn_train.cu -- host side + C ABI (include/wn_train_b200.h) of the B200 WaveNet training step.
//
// One step = WaveNetModel.add_loss (wavenet/model.py:247-312) + the gradients of add_optimizer (:327), then
// apply_gradients + EMA (:333-346).  Contractions run as cuBLASLt GEMMs (bf16 inputs on the tensor cores, fp32
// accumulation; or all-fp32 for validation), the rest are the kernels of wn_train_kernels.cuh.  See that file for the
// absolute-time row layout that turns every dilated tap into a row offset of one matrix.
#include <cublasLt.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/wn_train_b200.h"
#include "wn_train_kernels.cuh"
#include "wn_train_fused.cuh"

using namespace wnt;

namespace {

std::string g_create_error;

struct View {             // a TF variable of shape (outer, rows, cols) inside the flat buffer
    int64_t base;
    int outer;
    int64_t outer_stride;
    int rows, cols, ld;
    int64_t size() const { return (int64_t)outer * rows * cols; }
};

struct Plan {
    cublasLtMatmulDesc_t op = nullptr;
    cublasLtMatrixLayout_t a = nullptr, b = nullptr, c = nullptr, d = nullptr;
    cublasLtMatmulAlgo_t algo;
};

typedef std::tuple<int, int, long, long, long, long, long, long, long, int, int, int, int, int> PlanKey;

}  // namespace

struct wnt_handle {
    wnt_config cfg;
    std::string err;
    // geometry
    int N, L, R, D, S, O, OP, K, C, G, card, ifw, T, T0, SL, OW, rf, mel_frames;
    bool scalar = true;               // scalar input + MoL head, or one-hot input + softmax head
    int Q = 0, cw = 0, cin = 1;       // classes; causal filter width (ifw or filter_width = 2) and its input channels (1 or Q)
    int32_t *ids = nullptr;           // (N, T) mu-law ids of the waveform (one-hot model)
    long M, Mo;
    std::vector<int> s, off;          // input start / output start of each layer
    bool bf;                          // bf16 storage
    size_t esz;
    // flat layout (floats)
    int64_t n_params = 0, n_weights = 0, n_trainable = 0;
    int64_t o_layer_w, layer_w_stride, o_wfg, o_wlc, o_wgc, o_wd;       // per-layer kernel block (offsets inside the block)
    int64_t o_ws, o_wc, o_w1, o_w2, o_e, o_up[WNT_MAX_UPSAMPLE];
    int64_t o_layer_b, layer_b_stride, o_bfg, o_bd;                     // per-layer bias block
    int64_t o_bs, o_b1, o_b2;
    std::map<std::string, View> views;
    std::vector<std::string> names;
    // bound buffers
    float *P = nullptr, *Gr = nullptr, *Am = nullptr, *Av = nullptr, *Ema = nullptr;
    void *Pc = nullptr;               // compute-dtype copy of the parameters (== P for fp32)
    // activations / workspaces
    std::vector<float *> U;           // upsample stage inputs (fp32); U[0] = mel copy
    std::vector<float *> dU;
    void *LC = nullptr;               // (M, C) storage
    std::vector<void *> X, TS;        // per layer (M, R), (M, 2D)
    void *Zs = nullptr, *Z = nullptr, *T1 = nullptr, *T2 = nullptr, *dY = nullptr, *dC1 = nullptr, *dTot = nullptr, *dZs = nullptr;
    void *dXb = nullptr, *dFG = nullptr;
    float *FG32 = nullptr, *TOT = nullptr, *Y = nullptr, *dT32 = nullptr, *dX32 = nullptr, *dZ32 = nullptr, *dLC32 = nullptr;
    float *GCB = nullptr, *SB = nullptr, *bsum = nullptr, *dbs = nullptr;
    double *acc = nullptr;            // [0] loss sum, [1] l2 term, [2] grad sumsq
    int64_t workspace_bytes = 0;
    // cuBLASLt
    cublasLtHandle_t lt = nullptr;
    void *lt_ws = nullptr;
    size_t lt_ws_bytes = 64ull << 20;
    std::map<PlanKey, Plan> plans;
    int64_t gemm_launches = 0, kernel_launches = 0;
    double flops = 0;
    bool count_flops = false;
    bool have_step = false;
    int sm_count = 148;
    // fused tcgen05 forward (R = D = 128, bf16)
    bool fused = false;
    void *Xall = nullptr;             // all layer inputs stacked: (L*M, R) -- one TMA tensor map
    bf16 *WfgT = nullptr, *WdT = nullptr, *WdP = nullptr, *WdxP = nullptr;
    bf16 *dXp[2] = {nullptr, nullptr};   // ping-pong gradient w.r.t. the layer outputs / inputs (fused backward, bf16)
    bool fused_bwd = false, fused_persistent = true;
    CUtensorMap map_x, map_lc, map_wfg, map_wd, map_dx[2], map_dfg, map_wdp, map_wdxp;
    CUtensorMap map_x64, map_lc64, map_z64, map_dfg64, map_dx64[2];   // 64-row x 64-channel boxes: MN-major operands of the weight-gradient kernel
    bool fused_wgrad = false;
    bf16 *Wcol = nullptr, *WcDup = nullptr;   // causal layer as a GEMM: hi|lo im2col of the waveform, duplicated kernel
    float *dWcTmp = nullptr;
    bool causal_gemm = false;
    unsigned *fused_err = nullptr;
    int64_t fused_launches = 0;
};

namespace {

int fail(wnt_handle *h, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (h) h->err = buf; else g_create_error = buf;
    return code;
}

#define CK(call)                                                                                          \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess) return fail(h, WNT_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)
#define CKB(call)                                                                                         \
    do {                                                                                                  \
        cublasStatus_t s_ = (call);                                                                       \
        if (s_ != CUBLAS_STATUS_SUCCESS) return fail(h, WNT_ERR_CUBLAS, "%s: cublas status %d (%s:%d)", #call, (int)s_, __FILE__, __LINE__); \
    } while (0)
#define CKR(call)                 \
    do {                          \
        int r_ = (call);          \
        if (r_ != WNT_OK) return r_; \
    } while (0)
#define KCHECK()                                      \
    do {                                              \
        h->kernel_launches++;                         \
        CK(cudaGetLastError());                       \
    } while (0)

bool pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }
int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }
int grid_for(long items, int per_block, int cap) {
    long g = (items + per_block - 1) / per_block;
    if (g < 1) g = 1;
    if (g > cap) g = cap;
    return (int)g;
}
int ptr_align(const void *p, long ld_bytes) {
    uintptr_t v = (uintptr_t)p | (uintptr_t)ld_bytes | 256u;
    return (int)(v & (~v + 1));
}

template <typename F>
int alloc(wnt_handle *h, F **p, size_t bytes) {
    if (bytes == 0) bytes = 16;
    CK(cudaMalloc((void **)p, bytes));
    CK(cudaMemset(*p, 0, bytes));
    h->workspace_bytes += (int64_t)bytes;
    return WNT_OK;
}

// Row-major D[m,n] = op(A)[m,k] * op(B)[k,n] + beta * C (+ bias[n]) (+ relu).  epi: 0 none, 1 bias, 2 relu+bias, 3 relu.
int gemm(wnt_handle *h, cudaStream_t st, bool tA, bool tB, long m, long n, long k, const void *A, long lda, const void *B, long ldb,
         cudaDataType tAB, float beta, const void *C, long ldc, void *Dp, long ldd, cudaDataType tCD, int epi = 0,
         const void *bias = nullptr) {
    if (m <= 0 || n <= 0 || k <= 0) return WNT_OK;
    const long eab = tAB == CUDA_R_32F ? 4 : 2, ecd = tCD == CUDA_R_32F ? 4 : 2;
    const int alA = ptr_align(A, lda * eab), alB = ptr_align(B, ldb * eab), alC = ptr_align(C, ldc * ecd), alD = ptr_align(Dp, ldd * ecd);
    const int al = std::min(std::min(alA, alB), std::min(alC, alD));
    PlanKey key(tA ? 1 : 0, tB ? 1 : 0, m, n, k, lda, ldb, ldc, ldd, (int)tAB, (int)tCD, epi, al, beta == 0.f ? 0 : 1);
    auto it = h->plans.find(key);
    if (it == h->plans.end()) {
        Plan p;
        CKB(cublasLtMatmulDescCreate(&p.op, CUBLAS_COMPUTE_32F, CUDA_R_32F));
        // column-major view: D^T (n x m) = op(B)^T * op(A)^T  ->  cublas A := B, cublas B := A
        cublasOperation_t ta = tB ? CUBLAS_OP_T : CUBLAS_OP_N, tb = tA ? CUBLAS_OP_T : CUBLAS_OP_N;
        CKB(cublasLtMatmulDescSetAttribute(p.op, CUBLASLT_MATMUL_DESC_TRANSA, &ta, sizeof ta));
        CKB(cublasLtMatmulDescSetAttribute(p.op, CUBLASLT_MATMUL_DESC_TRANSB, &tb, sizeof tb));
        if (epi) {
            cublasLtEpilogue_t e = epi == 1 ? CUBLASLT_EPILOGUE_BIAS : epi == 2 ? CUBLASLT_EPILOGUE_RELU_BIAS : CUBLASLT_EPILOGUE_RELU;
            CKB(cublasLtMatmulDescSetAttribute(p.op, CUBLASLT_MATMUL_DESC_EPILOGUE, &e, sizeof e));
            if (epi != 3) {
                cudaDataType bt = tCD;
                CKB(cublasLtMatmulDescSetAttribute(p.op, CUBLASLT_MATMUL_DESC_BIAS_DATA_TYPE, &bt, sizeof bt));
            }
        }
        CKB(cublasLtMatrixLayoutCreate(&p.a, tAB, tB ? k : n, tB ? n : k, ldb));
        CKB(cublasLtMatrixLayoutCreate(&p.b, tAB, tA ? m : k, tA ? k : m, lda));
        CKB(cublasLtMatrixLayoutCreate(&p.c, tCD, n, m, ldc));
        CKB(cublasLtMatrixLayoutCreate(&p.d, tCD, n, m, ldd));
        cublasLtMatmulPreference_t pref;
        CKB(cublasLtMatmulPreferenceCreate(&pref));
        CKB(cublasLtMatmulPreferenceSetAttribute(pref, CUBLASLT_MATMUL_PREF_MAX_WORKSPACE_BYTES, &h->lt_ws_bytes, sizeof h->lt_ws_bytes));
        uint32_t ua = (uint32_t)al, ub = (uint32_t)al, uc = (uint32_t)al, ud = (uint32_t)al;   // conservative: the plan is shared by every call with this key
        CKB(cublasLtMatmulPreferenceSetAttribute(pref, CUBLASLT_MATMUL_PREF_MIN_ALIGNMENT_A_BYTES, &ua, sizeof ua));
        CKB(cublasLtMatmulPreferenceSetAttribute(pref, CUBLASLT_MATMUL_PREF_MIN_ALIGNMENT_B_BYTES, &ub, sizeof ub));
        CKB(cublasLtMatmulPreferenceSetAttribute(pref, CUBLASLT_MATMUL_PREF_MIN_ALIGNMENT_C_BYTES, &uc, sizeof uc));
        CKB(cublasLtMatmulPreferenceSetAttribute(pref, CUBLASLT_MATMUL_PREF_MIN_ALIGNMENT_D_BYTES, &ud, sizeof ud));
        if (epi == 1 || epi == 2) {
            const void *bp = bias;
            CKB(cublasLtMatmulDescSetAttribute(p.op, CUBLASLT_MATMUL_DESC_BIAS_POINTER, &bp, sizeof bp));
        }
        cublasLtMatmulHeuristicResult_t heur;
        int found = 0;
        cublasStatus_t hs = cublasLtMatmulAlgoGetHeuristic(h->lt, p.op, p.a, p.b, p.c, p.d, pref, 1, &heur, &found);
        cublasLtMatmulPreferenceDestroy(pref);
        if (hs != CUBLAS_STATUS_SUCCESS || found == 0)
            return fail(h, WNT_ERR_CUBLAS, "no cuBLASLt algorithm for gemm m=%ld n=%ld k=%ld tA=%d tB=%d types %d/%d epi %d (status %d)", m, n, k,
                        (int)tA, (int)tB, (int)tAB, (int)tCD, epi, (int)hs);
        p.algo = heur.algo;
        it = h->plans.emplace(key, p).first;
    }
    Plan &p = it->second;
    if (epi == 1 || epi == 2) {
        const void *bp = bias;
        CKB(cublasLtMatmulDescSetAttribute(p.op, CUBLASLT_MATMUL_DESC_BIAS_POINTER, &bp, sizeof bp));
    }
    const float alpha = 1.f;
    CKB(cublasLtMatmul(h->lt, p.op, &alpha, B, p.a, A, p.b, &beta, C, p.c, Dp, p.d, &p.algo, h->lt_ws, h->lt_ws_bytes, st));
    h->gemm_launches++;
    if (h->count_flops) h->flops += 2.0 * (double)m * (double)n * (double)k;
    return WNT_OK;
}

void add_view(wnt_handle *h, const std::string &name, int64_t base, int outer, int64_t ostride, int rows, int cols, int ld) {
    View v{base, outer, ostride, rows, cols, ld};
    h->views[name] = v;
    h->names.push_back(name);
    h->n_trainable += v.size();
}

int build_layout(wnt_handle *h) {
    const int L = h->L, R = h->R, D = h->D, S = h->S, C = h->C, G = h->G, D2 = 2 * h->D;
    const int64_t A = 128;   // segment alignment in elements: 256 B in bf16
    // ---- kernels ("weights": L2-regularised) ----
    int64_t o = 0;
    h->o_wfg = 0;
    int64_t b = (int64_t)2 * R * D2;
    h->o_wlc = align_up(b, A);
    b = h->o_wlc + (int64_t)C * D2;
    h->o_wgc = align_up(b, A);
    b = h->o_wgc + (int64_t)G * D2;
    h->o_wd = align_up(b, A);
    b = h->o_wd + (int64_t)D * R;
    h->layer_w_stride = align_up(b, A);
    h->o_layer_w = o;
    o += h->layer_w_stride * L;
    h->o_ws = o; o = align_up(o + (int64_t)L * D * S, A);
    h->o_wc = o; o = align_up(o + (int64_t)h->cw * h->cin * R, A);
    h->o_w1 = o; o = align_up(o + (int64_t)S * S, A);
    h->o_w2 = o; o = align_up(o + (int64_t)S * h->OP, A);
    h->o_e = o; o = align_up(o + (int64_t)h->card * G, A);
    for (int i = 0; i < h->cfg.n_upsample; ++i) {
        h->o_up[i] = o;
        o = align_up(o + 2 * h->cfg.upsample_factor[i], A);
    }
    h->n_weights = o;
    // ---- biases ----
    h->o_bfg = 0;
    h->o_bd = align_up(D2, A);
    h->layer_b_stride = align_up(h->o_bd + R, A);
    h->o_layer_b = o;
    o += h->layer_b_stride * L;
    h->o_bs = o; o = align_up(o + (int64_t)L * S, A);
    h->o_b1 = o; o = align_up(o + S, A);
    h->o_b2 = o; o = align_up(o + h->OP, A);
    h->n_params = o;
    // ---- TF variables in tf.trainable_variables() (creation) order: model.py:194, 107, 44, 68-96, 159-165 ----
    const bool ub = h->cfg.use_biases != 0;
    if (G) add_view(h, "wavenet/gc_embedding", h->o_e, 1, 0, h->card, G, G);
    for (int i = 0; i < h->cfg.n_upsample; ++i)
        add_view(h, "wavenet/upsample" + std::to_string(i) + "/kernel", h->o_up[i], 1, 0, h->cfg.upsample_factor[i], 2, 2);
    add_view(h, "wavenet/conv1d/kernel", h->o_wc, 1, 0, h->cw * h->cin, R, R);   // (cw, cin, R)
    for (int l = 0; l < L; ++l) {
        const std::string pre = "wavenet/dilated_stack/layer" + std::to_string(l) + "/dilation_layer/";
        const int64_t w = h->o_layer_w + h->layer_w_stride * l, bb = h->o_layer_b + h->layer_b_stride * l;
        add_view(h, pre + "conv_filter/kernel", w + h->o_wfg, 2, (int64_t)R * D2, R, D, D2);
        if (ub) add_view(h, pre + "conv_filter/bias", bb + h->o_bfg, 1, 0, 1, D, D);
        add_view(h, pre + "conv_gate/kernel", w + h->o_wfg + D, 2, (int64_t)R * D2, R, D, D2);
        if (ub) add_view(h, pre + "conv_gate/bias", bb + h->o_bfg + D, 1, 0, 1, D, D);
        if (G) {
            add_view(h, pre + "gc_filter/kernel", w + h->o_wgc, 1, 0, G, D, D2);
            add_view(h, pre + "gc_gate/kernel", w + h->o_wgc + D, 1, 0, G, D, D2);
        }
        if (C) {
            add_view(h, pre + "lc_filter/kernel", w + h->o_wlc, 1, 0, C, D, D2);
            add_view(h, pre + "lc_gate/kernel", w + h->o_wlc + D, 1, 0, C, D, D2);
        }
        add_view(h, pre + "dense/kernel", w + h->o_wd, 1, 0, D, R, R);
        if (ub) add_view(h, pre + "dense/bias", bb + h->o_bd, 1, 0, 1, R, R);
        add_view(h, pre + "skip/kernel", h->o_ws + (int64_t)l * D * S, 1, 0, D, S, S);
        if (ub) add_view(h, pre + "skip/bias", h->o_bs + (int64_t)l * S, 1, 0, 1, S, S);
    }
    add_view(h, "wavenet/conv1d_1/kernel", h->o_w1, 1, 0, S, S, S);
    if (ub) add_view(h, "wavenet/conv1d_1/bias", h->o_b1, 1, 0, 1, S, S);
    add_view(h, "wavenet/conv1d_2/kernel", h->o_w2, 1, 0, S, h->O, h->OP);
    if (ub) add_view(h, "wavenet/conv1d_2/bias", h->o_b2, 1, 0, 1, h->O, h->O);
    return WNT_OK;
}

float *buf_of(wnt_handle *h, int which) {
    switch (which) {
        case 0: return h->P;
        case 1: return h->Gr;
        case 2: return h->Ema;
        case 3: return h->Am;
        case 4: return h->Av;
        default: return nullptr;
    }
}

template <typename T>
int refresh_copy_t(wnt_handle *h, cudaStream_t st) {
    cast_kernel<T><<<grid_for(h->n_params, EW_THREADS, 8 * h->sm_count), EW_THREADS, 0, st>>>(h->P, (T *)h->Pc, (size_t)h->n_params);
    KCHECK();
    return WNT_OK;
}
int refresh_copy(wnt_handle *h, cudaStream_t st) {
    if (!h->bf) return WNT_OK;
    CKR(refresh_copy_t<bf16>(h, st));
    if (h->fused) {
        const long tot = (long)h->L * (wntf::NFG * wntf::KTOT + 2 * wntf::ND * wntf::ND + wntf::ND * 512);
        wntf::transpose_weights_kernel<<<grid_for(tot, EW_THREADS, 8 * h->sm_count), EW_THREADS, 0, st>>>(
            h->P, h->o_layer_w, h->layer_w_stride, h->o_wfg, h->o_wlc, h->o_wd, h->L, h->C, h->WfgT, h->WdT, h->WdP, h->WdxP);
        KCHECK();
        if (h->causal_gemm) {
            wntf::causal_dup_kernel<<<(h->ifw * h->R + 255) / 256, 256, 0, st>>>(h->P + h->o_wc, h->WcDup, h->ifw, h->R);
            KCHECK();
        }
    }
    return WNT_OK;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// bf16 row-major (rows, cols) tensor, box = (box_rows, 64 columns), 128-byte swizzle, out-of-bounds elements read as zero
int make_map(wnt_handle *h, EncodeTiledFn enc, CUtensorMap *m, void *base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    cuuint64_t dims[2] = {cols, rows}, strides[1] = {cols * 2};
    cuuint32_t box[2] = {64, box_rows}, es[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(h, WNT_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for a %llu x %llu tensor", (int)r, (unsigned long long)rows, (unsigned long long)cols);
    return WNT_OK;
}

int setup_fused(wnt_handle *h) {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn || q != cudaDriverEntryPointSuccess) return fail(h, WNT_ERR_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
    EncodeTiledFn enc = (EncodeTiledFn)fn;
    CKR(make_map(h, enc, &h->map_x, h->Xall, (uint64_t)h->L * h->M, 128, wntf::TILE_M));
    if (h->C) CKR(make_map(h, enc, &h->map_lc, h->LC, (uint64_t)h->M, (uint64_t)h->C, wntf::TILE_M));
    else h->map_lc = h->map_x;
    CKR(make_map(h, enc, &h->map_wfg, h->WfgT, (uint64_t)h->L * wntf::NFG, wntf::KTOT, wntf::NFG));
    CKR(make_map(h, enc, &h->map_wd, h->WdT, (uint64_t)h->L * wntf::ND, wntf::ND, wntf::ND));
    CK(cudaFuncSetAttribute(wntf::layer_fwd_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, wntf::SMEM_BYTES));
    CK(cudaFuncSetAttribute(wntf::layer_fwd_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, wntf::PF_SMEM_BYTES));
    h->fused_persistent = getenv("WNT_NO_PERSISTENT") == nullptr;
    for (int i = 0; i < 2; ++i) CKR(make_map(h, enc, &h->map_dx[i], h->dXp[i], (uint64_t)h->M, 128, wntf::TILE_M));
    CKR(make_map(h, enc, &h->map_dfg, h->dFG, (uint64_t)h->M, 256, wntf::TILE_M));
    CKR(make_map(h, enc, &h->map_wdp, h->WdP, (uint64_t)h->L * wntf::ND, wntf::ND, wntf::ND));
    CKR(make_map(h, enc, &h->map_wdxp, h->WdxP, (uint64_t)h->L * wntf::ND, 512, wntf::ND));
    CK(cudaFuncSetAttribute(wntf::layer_bwd_kernel<wntf::MODE_GATE>, cudaFuncAttributeMaxDynamicSharedMemorySize, wntf::BW_SMEM_BYTES));
    CK(cudaFuncSetAttribute(wntf::layer_bwd_kernel<wntf::MODE_DX>, cudaFuncAttributeMaxDynamicSharedMemorySize, wntf::BW_SMEM_BYTES));
    CKR(make_map(h, enc, &h->map_x64, h->Xall, (uint64_t)h->L * h->M, 128, wntf::WG_KROWS));
    if (h->C) CKR(make_map(h, enc, &h->map_lc64, h->LC, (uint64_t)h->M, (uint64_t)h->C, wntf::WG_KROWS));
    else h->map_lc64 = h->map_x64;
    CKR(make_map(h, enc, &h->map_z64, h->Z, (uint64_t)h->M, 128, wntf::WG_KROWS));
    CKR(make_map(h, enc, &h->map_dfg64, h->dFG, (uint64_t)h->M, 256, wntf::WG_KROWS));
    for (int i = 0; i < 2; ++i) CKR(make_map(h, enc, &h->map_dx64[i], h->dXp[i], (uint64_t)h->M, 128, wntf::WG_KROWS));
    CK(cudaFuncSetAttribute(wntf::layer_wgrad_kernel<wntf::MODE_WFG>, cudaFuncAttributeMaxDynamicSharedMemorySize, wntf::WgLayout<wntf::MODE_WFG>::SMEM));
    CK(cudaFuncSetAttribute(wntf::layer_wgrad_kernel<wntf::MODE_WLD>, cudaFuncAttributeMaxDynamicSharedMemorySize, wntf::WgLayout<wntf::MODE_WLD>::SMEM));
    h->fused_wgrad = getenv("WNT_NO_FUSED_WGRAD") == nullptr;
    CK(cudaFuncSetAttribute(wntf::layer_bwd_persistent_kernel<wntf::MODE_GATE>, cudaFuncAttributeMaxDynamicSharedMemorySize, wntf::PbLayout<wntf::MODE_GATE>::SMEM));
    CK(cudaFuncSetAttribute(wntf::layer_bwd_persistent_kernel<wntf::MODE_DX>, cudaFuncAttributeMaxDynamicSharedMemorySize, wntf::PbLayout<wntf::MODE_DX>::SMEM));
    return WNT_OK;
}

// ---- the step ---------------------------------------------------------------------------------------------------------
template <typename T>
int step_t(wnt_handle *h, const float *wav, const float *mel, const int32_t *gc_ids, float l2, float *loss_dev, cudaStream_t st) {
    const int N = h->N, L = h->L, R = h->R, D = h->D, S = h->S, OP = h->OP, C = h->C, G = h->G, D2 = 2 * h->D, T0 = h->T0;
    const long M = h->M, Mo = h->Mo;
    const int LD = L * D;
    const cudaDataType ts = h->bf ? CUDA_R_16BF : CUDA_R_32F, f32 = CUDA_R_32F;
    const int cap = 8 * h->sm_count;
    const bool ub = h->cfg.use_biases != 0;
    const T *Pc = (const T *)h->Pc;
    const float *P = h->P;
    float *Gr = h->Gr;
    auto lw = [&](int l) { return h->o_layer_w + h->layer_w_stride * l; };
    auto lb = [&](int l) { return h->o_layer_b + h->layer_b_stride * l; };
    const int CH = 128;

    CK(cudaMemsetAsync(h->acc, 0, 3 * sizeof(double), st));
    // ---- conditioning ----
    if (C) {
        int Ti = h->mel_frames;
        const float *in = mel;
        for (int i = 0; i < h->cfg.n_upsample; ++i) {
            const int F = h->cfg.upsample_factor[i];
            const bool last = i == h->cfg.n_upsample - 1;
            const int rows_out = last ? T0 : Ti * F;
            const long total = (long)N * rows_out * C;
            if (last)
                ups_fwd_kernel<T><<<grid_for(total, EW_THREADS, cap), EW_THREADS, 0, st>>>(in, P + h->o_up[i], (T *)h->LC, N, Ti, F, C, rows_out);
            else
                ups_fwd_kernel<float><<<grid_for(total, EW_THREADS, cap), EW_THREADS, 0, st>>>(in, P + h->o_up[i], h->U[i + 1], N, Ti, F, C, rows_out);
            KCHECK();
            if (!last) in = h->U[i + 1];
            Ti *= F;
        }
    }
    if (G) {
        gc_bias_kernel<<<dim3(L, N), 256, 0, st>>>(P + h->o_e, gc_ids, P + h->o_layer_w + h->o_wgc, (size_t)h->layer_w_stride, h->GCB, N, G, D2);
        KCHECK();
    }
    // ---- causal layer ----
    if (!h->scalar) {
        mulaw_ids_kernel<<<grid_for((long)N * h->T, EW_THREADS, cap), EW_THREADS, 0, st>>>(wav, h->ids, (size_t)N * h->T, h->Q);
        KCHECK();
        onehot_causal_fwd_kernel<T><<<grid_for(M, EW_THREADS / (R / 2), cap), EW_THREADS, 0, st>>>(h->ids, P + h->o_wc, (T *)h->X[0], h->T, T0, M,
                                                                                                h->Q, R);
        KCHECK();
    } else {
        if (h->fused && h->causal_gemm) {
            wntf::wav_im2col_kernel<<<grid_for((long)M * h->ifw, EW_THREADS, cap), EW_THREADS, 0, st>>>(wav, h->Wcol, N, h->T, T0, h->ifw);
            KCHECK();
            CKR(gemm(h, st, false, false, M, R, 2 * h->ifw, h->Wcol, 2 * h->ifw, h->WcDup, R, ts, 0.f, h->X[0], R, h->X[0], R, ts));
        } else {
            const size_t sm = ((size_t)h->ifw * R + CH + h->ifw) * sizeof(float);
            causal_fwd_kernel<T><<<dim3((T0 + CH - 1) / CH, N), EW_THREADS, sm, st>>>(wav, P + h->o_wc, (T *)h->X[0], h->T, T0, h->ifw, R, CH);
            KCHECK();
        }
    }
    // ---- dilation stack ----
    for (int l = 0; l < L; ++l) {
        const int d = h->cfg.dilations[l];
        const long off = h->off[l], m = M - off;
        const T *Xl = (const T *)h->X[l];
        const T *W = Pc + lw(l);
        if (h->fused) {
            wntf::FusedArgs fa;
            fa.l = l; fa.d = d; fa.off = (int)off; fa.SL = h->SL; fa.OW = h->OW; fa.T0 = T0; fa.LD = LD; fa.zs_col0 = l * D;
            fa.do_dense = l + 1 < L ? 1 : 0; fa.has_lc = C ? 1 : 0; fa.N = N;
            fa.M = M; fa.x_row0 = (long)l * M;
            fa.bias = ub ? P + lb(l) + h->o_bfg : nullptr;
            fa.gcb = G ? h->GCB + (size_t)l * N * D2 : nullptr;
            fa.bd = ub ? P + lb(l) + h->o_bd : nullptr;
            fa.Xl = (const bf16 *)h->X[l];
            fa.Xn = l + 1 < L ? (bf16 *)h->X[l + 1] : nullptr;
            fa.TS = (bf16 *)h->TS[l];
            fa.Zs = (bf16 *)h->Zs;
            fa.err = h->fused_err;
            const unsigned tiles = (unsigned)((m + wntf::TILE_M - 1) / wntf::TILE_M);
            if (h->fused_persistent)
                wntf::layer_fwd_persistent_kernel<<<std::min<unsigned>(tiles, (unsigned)h->sm_count), wntf::PF_THREADS, wntf::PF_SMEM_BYTES, st>>>(
                    h->map_x, h->map_lc, h->map_wfg, h->map_wd, fa, (int)tiles);
            else
                wntf::layer_fwd_fused_kernel<<<tiles, wntf::THREADS, wntf::SMEM_BYTES, st>>>(h->map_x, h->map_lc, h->map_wfg, h->map_wd, fa);
            KCHECK();
            h->fused_launches++;
            if (h->count_flops) h->flops += 2.0 * (double)m * (D2 * (2.0 * R + C) + (l + 1 < L ? (double)D * R : 0.0));
            continue;
        }
        float *FG = h->FG32 + off * D2;
        CKR(gemm(h, st, false, false, m, D2, R, Xl + (off - d) * R, R, W + h->o_wfg, D2, ts, 0.f, FG, D2, FG, D2, f32));
        CKR(gemm(h, st, false, false, m, D2, R, Xl + off * R, R, W + h->o_wfg + (int64_t)R * D2, D2, ts, 1.f, FG, D2, FG, D2, f32));
        if (C) CKR(gemm(h, st, false, false, m, D2, C, h->LC, C, W + h->o_wlc, D2, ts, 1.f, FG, D2, FG, D2, f32));
        gate_fwd_kernel<T><<<grid_for(m, EW_THREADS / (D / 2), cap), EW_THREADS, 0, st>>>(
            h->FG32, ub ? P + lb(l) + h->o_bfg : nullptr, G ? h->GCB + (size_t)l * N * D2 : nullptr, (T *)h->TS[l], (T *)h->Z, (T *)h->Zs, off, M,
            T0, D, h->SL, h->OW, l * D, LD);
        KCHECK();
        if (l + 1 < L) {   // the last layer's residual output feeds nothing (model.py:147: only `outputs` is used)
            T *Xn = (T *)h->X[l + 1];
            CKR(gemm(h, st, false, false, m, R, D, (const T *)h->Z + off * D, D, W + h->o_wd, R, ts, 1.f, Xl + off * R, R, Xn + off * R, R, ts,
                     ub ? 1 : 0, ub ? (const void *)(Pc + lb(l) + h->o_bd) : nullptr));
        }
    }
    // ---- post-processing: sum of skips as ONE GEMM over the concatenated z (K = L*D) ----
    CKR(gemm(h, st, false, false, Mo, S, LD, h->Zs, LD, Pc + h->o_ws, S, ts, 0.f, h->TOT, S, h->TOT, S, f32));
    if (ub) {
        skip_bias_sum_kernel<<<(S + 255) / 256, 256, 0, st>>>(P + h->o_bs, h->bsum, L, S);
        KCHECK();
    }
    bias_relu_kernel<T><<<grid_for(Mo, EW_THREADS / (S / 2), cap), EW_THREADS, 0, st>>>(h->TOT, ub ? h->bsum : nullptr, (T *)h->T1, Mo, S);
    KCHECK();
    CKR(gemm(h, st, false, false, Mo, S, S, h->T1, S, Pc + h->o_w1, S, ts, 0.f, h->T2, S, h->T2, S, ts, ub ? 2 : 3,
             ub ? (const void *)(Pc + h->o_b1) : nullptr));
    CKR(gemm(h, st, false, false, Mo, OP, S, h->T2, S, Pc + h->o_w2, OP, ts, 0.f, h->Y, OP, h->Y, OP, f32, ub ? 1 : 0,
             ub ? (const void *)(P + h->o_b2) : nullptr));
    // ---- loss + d loss / d raw_output ----
    CK(cudaMemsetAsync(Gr + h->o_layer_b, 0, (size_t)(h->n_params - h->o_layer_b) * sizeof(float), st));   // all bias grads (atomics)
    if (h->scalar)
        mol_loss_kernel<T><<<(unsigned)((Mo + EW_THREADS - 1) / EW_THREADS), EW_THREADS, 0, st>>>(
            h->Y, wav, (T *)h->dY, ub ? Gr + h->o_b2 : nullptr, h->acc, Mo, h->OW, h->T, h->rf, h->K, OP, (float)log(1e-14), 1.0f / 65535.0f,
            (float)log(65535.0 / 2.0), (float)(1.0 / (double)Mo));
    else
        softmax_ce_kernel<T><<<grid_for(Mo, EW_THREADS / 32, cap), EW_THREADS, 0, st>>>(h->Y, h->ids, (T *)h->dY, ub ? Gr + h->o_b2 : nullptr, h->acc,
                                                                                     Mo, h->OW, h->T, h->rf, h->Q, OP, (float)(1.0 / (double)Mo));
    KCHECK();
    // ---- backward: post-processing ----
    CKR(gemm(h, st, true, false, S, OP, Mo, h->T2, S, h->dY, OP, ts, 0.f, Gr + h->o_w2, OP, Gr + h->o_w2, OP, f32));
    CKR(gemm(h, st, false, true, Mo, S, OP, h->dY, OP, Pc + h->o_w2, OP, ts, 0.f, h->dT32, S, h->dT32, S, f32));
    relu_bwd_colsum_kernel<T><<<grid_for(Mo, EW_THREADS / (S / 2), cap), EW_THREADS, 0, st>>>(h->dT32, (const T *)h->T2, (T *)h->dC1,
                                                                                            ub ? Gr + h->o_b1 : nullptr, Mo, S);
    KCHECK();
    CKR(gemm(h, st, true, false, S, S, Mo, h->T1, S, h->dC1, S, ts, 0.f, Gr + h->o_w1, S, Gr + h->o_w1, S, f32));
    CKR(gemm(h, st, false, true, Mo, S, S, h->dC1, S, Pc + h->o_w1, S, ts, 0.f, h->dT32, S, h->dT32, S, f32));
    CK(cudaMemsetAsync(h->dbs, 0, S * sizeof(float), st));
    relu_bwd_colsum_kernel<T><<<grid_for(Mo, EW_THREADS / (S / 2), cap), EW_THREADS, 0, st>>>(h->dT32, (const T *)h->T1, (T *)h->dTot,
                                                                                            ub ? h->dbs : nullptr, Mo, S);
    KCHECK();
    if (ub) {
        skip_bias_bcast_kernel<<<(L * S + 255) / 256, 256, 0, st>>>(h->dbs, Gr + h->o_bs, L, S);
        KCHECK();
    }
    CKR(gemm(h, st, true, false, LD, S, Mo, h->Zs, LD, h->dTot, S, ts, 0.f, Gr + h->o_ws, S, Gr + h->o_ws, S, f32));
    CKR(gemm(h, st, false, true, Mo, LD, S, h->dTot, S, Pc + h->o_ws, S, ts, 0.f, h->dZs, LD, h->dZs, LD, ts));
    // ---- backward: dilation stack ----
    if (h->dX32) CK(cudaMemsetAsync(h->dX32, 0, (size_t)M * R * sizeof(float), st));
    if (h->fused && h->fused_bwd)
        for (int i = 0; i < 2; ++i) CK(cudaMemsetAsync(h->dXp[i], 0, (size_t)M * R * 2, st));
    if (h->fused && h->fused_bwd && h->fused_wgrad)   // the split-K weight-gradient kernels accumulate with atomics
        CK(cudaMemsetAsync(Gr + h->o_layer_w, 0, (size_t)h->layer_w_stride * L * sizeof(float), st));
    if (C) CK(cudaMemsetAsync(h->dLC32, 0, (size_t)M * C * sizeof(float), st));
    CK(cudaMemsetAsync(h->SB, 0, (size_t)L * N * D2 * sizeof(float), st));
    for (int l = L - 1; l >= 0; --l) {
        const int d = h->cfg.dilations[l];
        const long off = h->off[l], m = M - off;
        const T *Xl = (const T *)h->X[l];
        const T *W = Pc + lw(l);
        float *GW = Gr + lw(l);
        const bool dense = l + 1 < L;
        T *dXb = (T *)h->dXb, *dFG = (T *)h->dFG, *Zb = (T *)h->Z;
        if (h->fused && h->fused_bwd) {
            // tcgen05 path: dx travels between layers in bf16 (ping-pong buffers, zeroed once per step so that the rows before
            // each layer's input start read as exact zeros); see wn_train_fused.cuh
            const int cur = (L - 1 - l) & 1;
            bf16 *dXn = h->dXp[cur], *dXo = h->dXp[cur ^ 1];
            const long s_l = off - d;
            const unsigned tiles = (unsigned)((M - s_l + wntf::TILE_M - 1) / wntf::TILE_M);
            if (dense && ub) {
                wntf::colsum_bf16_kernel<<<dim3((T0 + 1023) / 1024, N), EW_THREADS, 0, st>>>(dXn, Gr + lb(l) + h->o_bd, T0, R, 0, 0, 1024);
                KCHECK();
            }
            if (!dense) CK(cudaMemsetAsync(GW + h->o_wd, 0, (size_t)D * R * sizeof(float), st));
            wntf::BwdArgs ba;
            ba.l = l; ba.d = d; ba.off = (int)off; ba.s = (int)s_l; ba.SL = h->SL; ba.OW = h->OW; ba.T0 = T0; ba.LD = LD; ba.zs_col0 = l * D;
            ba.has_dense = dense ? 1 : 0; ba.M = M;
            ba.TS = (const bf16 *)h->TS[l]; ba.dZs = (const bf16 *)h->dZs; ba.dFG = (bf16 *)h->dFG; ba.Z = dense ? (bf16 *)h->Z : nullptr;
            ba.dXin = dXn; ba.dXout = dXo; ba.err = h->fused_err;
            const unsigned pgrid = std::min<unsigned>(tiles, (unsigned)h->sm_count);
            if (h->fused_persistent)
                wntf::layer_bwd_persistent_kernel<wntf::MODE_GATE><<<pgrid, wntf::PF_THREADS, wntf::PbLayout<wntf::MODE_GATE>::SMEM, st>>>(h->map_dx[cur], h->map_wdp, ba, (int)tiles);
            else
                wntf::layer_bwd_kernel<wntf::MODE_GATE><<<tiles, wntf::THREADS, wntf::BW_SMEM_BYTES, st>>>(h->map_dx[cur], h->map_wdp, ba);
            KCHECK();
            wntf::colsum_bf16_kernel<<<dim3((T0 + 511) / 512, N), EW_THREADS, 0, st>>>((const bf16 *)h->dFG, h->SB + (size_t)l * N * D2, T0, D2, (int)off, 1, 512);
            KCHECK();
            const bf16 *Xb = (const bf16 *)h->X[l], *dF = (const bf16 *)h->dFG, *Zq = (const bf16 *)h->Z;
            if (h->fused_wgrad) {
                // split-K tcgen05 weight gradients straight from the row-major activations (MN-major operands); dFG is read once per kernel
                wntf::WgArgs wa;
                wa.l = l; wa.d = d; wa.off = (int)off; wa.n_kblocks = (int)((m + wntf::WG_KROWS - 1) / wntf::WG_KROWS);
                wa.M = M; wa.x_row0 = (long)l * M; wa.err = h->fused_err;
                const unsigned wgrid = std::min<unsigned>((unsigned)wa.n_kblocks, (unsigned)h->sm_count);
                wa.dW0 = GW + h->o_wfg; wa.dW1 = GW + h->o_wfg + (int64_t)R * D2; wa.rows0 = 128;
                wntf::layer_wgrad_kernel<wntf::MODE_WFG><<<wgrid, wntf::PF_THREADS, wntf::WgLayout<wntf::MODE_WFG>::SMEM, st>>>(
                    h->map_x64, h->map_x64, h->map_dfg64, h->map_dfg64, wa);
                KCHECK();
                wa.dW0 = GW + h->o_wlc; wa.dW1 = GW + h->o_wd; wa.rows0 = C;
                wntf::layer_wgrad_kernel<wntf::MODE_WLD><<<wgrid, wntf::PF_THREADS, wntf::WgLayout<wntf::MODE_WLD>::SMEM, st>>>(
                    h->map_lc64, h->map_z64, h->map_dfg64, h->map_dx64[cur], wa);
                KCHECK();
                h->fused_launches += 2;
                if (h->count_flops) h->flops += 2.0 * (double)m * ((2.0 * R + C) * D2 + (dense ? (double)D * R : 0.0));
                if (C) CKR(gemm(h, st, false, true, m, C, D2, dF + off * D2, D2, W + h->o_wlc, D2, ts, 1.f, h->dLC32, C, h->dLC32, C, f32));
            } else {
            if (dense) CKR(gemm(h, st, true, false, D, R, m, Zq + off * D, D, dXn + off * R, R, ts, 0.f, GW + h->o_wd, R, GW + h->o_wd, R, f32));
            CKR(gemm(h, st, true, false, R, D2, m, Xb + (off - d) * R, R, dF + off * D2, D2, ts, 0.f, GW + h->o_wfg, D2, GW + h->o_wfg, D2, f32));
            CKR(gemm(h, st, true, false, R, D2, m, Xb + off * R, R, dF + off * D2, D2, ts, 0.f, GW + h->o_wfg + (int64_t)R * D2, D2,
                     GW + h->o_wfg + (int64_t)R * D2, D2, f32));
            if (C) {
                CKR(gemm(h, st, true, false, C, D2, m, h->LC, C, dF + off * D2, D2, ts, 0.f, GW + h->o_wlc, D2, GW + h->o_wlc, D2, f32));
                CKR(gemm(h, st, false, true, m, C, D2, dF + off * D2, D2, W + h->o_wlc, D2, ts, 1.f, h->dLC32, C, h->dLC32, C, f32));
            }
            }
            if (h->fused_persistent)
                wntf::layer_bwd_persistent_kernel<wntf::MODE_DX><<<pgrid, wntf::PF_THREADS, wntf::PbLayout<wntf::MODE_DX>::SMEM, st>>>(h->map_dfg, h->map_wdxp, ba, (int)tiles);
            else
                wntf::layer_bwd_kernel<wntf::MODE_DX><<<tiles, wntf::THREADS, wntf::BW_SMEM_BYTES, st>>>(h->map_dfg, h->map_wdxp, ba);
            KCHECK();
            h->fused_launches += 2;
            if (h->count_flops) h->flops += 2.0 * (double)(M - s_l) * ((dense ? (double)D * R : 0.0) + 2.0 * D2 * R);
            continue;
        }
        if (dense) {
            cast_colsum_kernel<T><<<grid_for(m, EW_THREADS / (R / 2), cap), EW_THREADS, 0, st>>>(h->dX32, dXb, ub ? Gr + lb(l) + h->o_bd : nullptr,
                                                                                              off, M, R);
            KCHECK();
            CKR(gemm(h, st, false, true, m, D, R, dXb + off * R, R, W + h->o_wd, R, ts, 0.f, h->dZ32 + off * D, D, h->dZ32 + off * D, D, f32));
        } else {
            CK(cudaMemsetAsync(GW + h->o_wd, 0, (size_t)D * R * sizeof(float), st));   // unconnected variable: zero gradient
        }
        gate_bwd_kernel<T><<<dim3((T0 + CH - 1) / CH, N), EW_THREADS, 0, st>>>(dense ? h->dZ32 : nullptr, (const T *)h->dZs, (const T *)h->TS[l], dFG,
                                                                              dense ? Zb : nullptr, h->SB + (size_t)l * N * D2, off, T0, D,
                                                                              h->SL, h->OW, (int)off, l * D, LD, CH);
        KCHECK();
        if (dense) CKR(gemm(h, st, true, false, D, R, m, Zb + off * D, D, dXb + off * R, R, ts, 0.f, GW + h->o_wd, R, GW + h->o_wd, R, f32));
        CKR(gemm(h, st, true, false, R, D2, m, Xl + (off - d) * R, R, dFG + off * D2, D2, ts, 0.f, GW + h->o_wfg, D2, GW + h->o_wfg, D2, f32));
        CKR(gemm(h, st, true, false, R, D2, m, Xl + off * R, R, dFG + off * D2, D2, ts, 0.f, GW + h->o_wfg + (int64_t)R * D2, D2,
                 GW + h->o_wfg + (int64_t)R * D2, D2, f32));
        if (C) {
            CKR(gemm(h, st, true, false, C, D2, m, h->LC, C, dFG + off * D2, D2, ts, 0.f, GW + h->o_wlc, D2, GW + h->o_wlc, D2, f32));
            CKR(gemm(h, st, false, true, m, C, D2, dFG + off * D2, D2, W + h->o_wlc, D2, ts, 1.f, h->dLC32, C, h->dLC32, C, f32));
        }
        // gradient w.r.t. this layer's input (x0's is consumed by the causal kernel): in place, dX32 already holds the residual path
        CKR(gemm(h, st, false, true, m, R, D2, dFG + off * D2, D2, W + h->o_wfg + (int64_t)R * D2, D2, ts, 1.f, h->dX32 + off * R, R,
                 h->dX32 + off * R, R, f32));
        CKR(gemm(h, st, false, true, m, R, D2, dFG + off * D2, D2, W + h->o_wfg, D2, ts, 1.f, h->dX32 + (off - d) * R, R, h->dX32 + (off - d) * R, R,
                 f32));
    }
    // ---- backward: causal layer, conditioning ----
    {
        CK(cudaMemsetAsync(Gr + h->o_wc, 0, (size_t)h->cw * h->cin * R * sizeof(float), st));
        const size_t sm = (size_t)(CH + h->ifw) * sizeof(float);
        const int og = grid_for(M, EW_THREADS / (R / 2), cap);
        if (!h->scalar && h->fused && h->fused_bwd)
            onehot_causal_bwd_kernel<bf16><<<og, EW_THREADS, 0, st>>>(h->ids, h->dXp[L & 1], Gr + h->o_wc, h->T, T0, M, h->Q, R);
        else if (!h->scalar)
            onehot_causal_bwd_kernel<float><<<og, EW_THREADS, 0, st>>>(h->ids, h->dX32, Gr + h->o_wc, h->T, T0, M, h->Q, R);
        else if (h->fused && h->fused_bwd && h->causal_gemm) {
            CKR(gemm(h, st, true, false, 2 * h->ifw, R, M, h->Wcol, 2 * h->ifw, h->dXp[L & 1], R, ts, 0.f, h->dWcTmp, R, h->dWcTmp, R, f32));
            wntf::causal_fold_kernel<<<(h->ifw * R + 255) / 256, 256, 0, st>>>(h->dWcTmp, Gr + h->o_wc, h->ifw, R);
        } else if (h->fused && h->fused_bwd)
            causal_bwd_kernel<bf16><<<dim3((T0 + CH - 1) / CH, N), EW_THREADS, sm, st>>>(wav, h->dXp[L & 1], Gr + h->o_wc, h->T, T0, h->ifw, R, CH);
        else
            causal_bwd_kernel<float><<<dim3((T0 + CH - 1) / CH, N), EW_THREADS, sm, st>>>(wav, h->dX32, Gr + h->o_wc, h->T, T0, h->ifw, R, CH);
        KCHECK();
    }
    if (ub) {
        sb_reduce_kernel<<<L, 256, 0, st>>>(h->SB, Gr + h->o_layer_b + h->o_bfg, (size_t)h->layer_b_stride, N, D2);
        KCHECK();
    }
    if (G) {
        gc_wgrad_kernel<<<dim3(L, G), 256, 0, st>>>(P + h->o_e, gc_ids, h->SB, Gr + h->o_layer_w + h->o_wgc, (size_t)h->layer_w_stride, N, G, D2);
        KCHECK();
        CK(cudaMemsetAsync(Gr + h->o_e, 0, (size_t)h->card * G * sizeof(float), st));
        gc_egrad_kernel<<<dim3(N, G), 256, 0, st>>>(gc_ids, h->SB, P + h->o_layer_w + h->o_wgc, (size_t)h->layer_w_stride, Gr + h->o_e, L, N, G, D2);
        KCHECK();
    }
    if (C) {
        const float *dout = h->dLC32;
        int rows_out = T0;
        for (int i = h->cfg.n_upsample - 1; i >= 0; --i) {
            const int F = h->cfg.upsample_factor[i];
            int Ti = h->mel_frames;
            for (int j = 0; j < i; ++j) Ti *= h->cfg.upsample_factor[j];
            const float *in = i == 0 ? mel : h->U[i];
            float *din = i == 0 ? nullptr : h->dU[i];
            CK(cudaMemsetAsync(Gr + h->o_up[i], 0, 2 * F * sizeof(float), st));
            ups_bwd_kernel<<<grid_for((long)N * Ti * C, EW_THREADS, cap), EW_THREADS, 0, st>>>(in, P + h->o_up[i], dout, Gr + h->o_up[i], din, N, Ti, F,
                                                                                            C, rows_out);
            KCHECK();
            dout = din;
            rows_out = Ti;
        }
    }
    if (l2 >= 0.f) {
        l2_kernel<<<grid_for(h->n_weights, EW_THREADS, cap), EW_THREADS, 0, st>>>(P, Gr, h->acc, (size_t)h->n_weights, l2);
        KCHECK();
    }
    loss_finish_kernel<<<1, 1, 0, st>>>(h->acc, loss_dev, 1.0 / (double)Mo);
    KCHECK();
    h->have_step = true;
    return WNT_OK;
}

template <typename T>
int apply_t(wnt_handle *h, const wnt_adam *a, cudaStream_t st) {
    const double lr_t = (double)a->learning_rate * std::sqrt(1.0 - std::pow((double)a->beta2, (double)a->t)) /
                        (1.0 - std::pow((double)a->beta1, (double)a->t));
    const double *ss = nullptr;
    if (a->clip_norm > 0.f) {
        CK(cudaMemsetAsync(h->acc + 2, 0, sizeof(double), st));
        sumsq_kernel<<<grid_for(h->n_params, EW_THREADS, 8 * h->sm_count), EW_THREADS, 0, st>>>(h->Gr, h->acc + 2, (size_t)h->n_params);
        KCHECK();
        ss = h->acc + 2;
    }
    adam_ema_kernel<T><<<grid_for(h->n_params, EW_THREADS, 8 * h->sm_count), EW_THREADS, 0, st>>>(
        h->P, h->Gr, h->Am, h->Av, h->Ema, h->bf ? (T *)h->Pc : nullptr, (size_t)h->n_params, (float)lr_t, a->beta1, a->beta2, a->epsilon,
        a->ema_decay, a->grad_scale, ss, a->clip_norm);
    KCHECK();
    return WNT_OK;
}

}  // namespace

// ---- C ABI ---------------------------------------------------------------------------------------------------------
extern "C" {

const char *wnt_last_error(const wnt_handle *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int wnt_create(const wnt_config *cfg, wnt_handle **out) {
    wnt_handle *h = nullptr;
    if (!cfg || !out) return fail(h, WNT_ERR_ARG, "null argument");
    *out = nullptr;
    if (cfg->n_layers < 1 || cfg->n_layers > WNT_MAX_LAYERS) return fail(h, WNT_ERR_ARG, "n_layers out of range");
    if (!pow2(cfg->residual_channels) || cfg->residual_channels < 8 || cfg->residual_channels > 256)
        return fail(h, WNT_ERR_UNSUPPORTED, "residual_channels must be a power of two in [8, 256]");
    if (!pow2(cfg->dilation_channels) || cfg->dilation_channels < 8 || cfg->dilation_channels > 512)
        return fail(h, WNT_ERR_UNSUPPORTED, "dilation_channels must be a power of two in [8, 512]");
    if (!pow2(cfg->skip_channels) || cfg->skip_channels < 8 || cfg->skip_channels > 512)
        return fail(h, WNT_ERR_UNSUPPORTED, "skip_channels must be a power of two in [8, 512]");
    if (cfg->scalar_input && (cfg->out_channels % 3 || cfg->out_channels < 3 || cfg->out_channels / 3 > MOL_MAX_K))
        return fail(h, WNT_ERR_ARG, "out_channels must be 3*nr_mix with nr_mix <= %d", MOL_MAX_K);
    if (!cfg->scalar_input && (cfg->quantization_channels < 32 || cfg->quantization_channels % 32 || cfg->quantization_channels > 32 * CE_PER_LANE))
        return fail(h, WNT_ERR_UNSUPPORTED, "quantization_channels must be a multiple of 32 in [32, %d]", 32 * CE_PER_LANE);
    if (cfg->lc_channels % 8) return fail(h, WNT_ERR_UNSUPPORTED, "lc_channels must be a multiple of 8");
    if (cfg->gc_channels < 0 || (cfg->gc_channels > 0 && cfg->gc_cardinality < 1)) return fail(h, WNT_ERR_ARG, "gc_cardinality missing");
    if (cfg->batch_size < 1 || cfg->initial_filter_width < 1 || cfg->initial_filter_width > 64) return fail(h, WNT_ERR_ARG, "bad batch_size / initial_filter_width");
    if (cfg->dtype != WNT_DTYPE_BF16 && cfg->dtype != WNT_DTYPE_FP32) return fail(h, WNT_ERR_ARG, "bad dtype");
    if (cfg->scalar_input) {
        const int KG = EW_THREADS / cfg->residual_channels;
        if ((cfg->initial_filter_width + KG - 1) / KG > CAUSAL_TAPS) return fail(h, WNT_ERR_UNSUPPORTED, "initial_filter_width too large for residual_channels");
    }
    int hop = 1;
    if (cfg->lc_channels) {
        if (cfg->n_upsample < 1 || cfg->n_upsample > WNT_MAX_UPSAMPLE) return fail(h, WNT_ERR_ARG, "upsample_factor required with local conditioning");
        for (int i = 0; i < cfg->n_upsample; ++i) {
            if (cfg->upsample_factor[i] < 1 || cfg->upsample_factor[i] > UPS_MAX_F) return fail(h, WNT_ERR_UNSUPPORTED, "upsample factor must be in [1, %d]", UPS_MAX_F);
            hop *= cfg->upsample_factor[i];
        }
        if (cfg->sample_size % hop) return fail(h, WNT_ERR_ARG, "sample_size %d is not a multiple of prod(upsample_factor) = %d", cfg->sample_size, hop);
    }
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess)
        return fail(h, WNT_ERR_CUDA, "no CUDA device (there is no CPU fallback)");
    if (major < 10) return fail(h, WNT_ERR_CUDA, "libwn_train_b200 is built for sm_100a only (device is sm_%d0)", major);
    h = new wnt_handle();
    h->cfg = *cfg;
    cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, dev);
    h->N = cfg->batch_size; h->L = cfg->n_layers; h->R = cfg->residual_channels; h->D = cfg->dilation_channels; h->S = cfg->skip_channels;
    h->scalar = cfg->scalar_input != 0;
    h->Q = cfg->quantization_channels;
    h->O = h->scalar ? cfg->out_channels : h->Q;            // model.py:163-165
    h->OP = (int)align_up(h->O, 8); h->K = h->scalar ? cfg->out_channels / 3 : 0;
    h->C = cfg->lc_channels; h->G = cfg->gc_channels; h->card = cfg->gc_channels ? cfg->gc_cardinality : 0;
    h->ifw = cfg->initial_filter_width; h->T = cfg->sample_size;
    h->cw = h->scalar ? h->ifw : 2;                         // model.py:41-46: initial_filter_width taps of 1 channel, or filter_width taps of Q
    h->cin = h->scalar ? 1 : h->Q;
    h->T0 = h->T - 1 - (h->cw - 1);
    h->bf = cfg->dtype == WNT_DTYPE_BF16;
    h->esz = h->bf ? 2 : 4;
    int sum = 0;
    for (int l = 0; l < h->L; ++l) {
        if (cfg->dilations[l] < 1) { delete h; return fail(nullptr, WNT_ERR_ARG, "bad dilation"); }
        h->s.push_back(sum);
        sum += cfg->dilations[l];
        h->off.push_back(sum);
    }
    h->SL = sum;
    h->rf = sum + 1 + (h->cw - 1);                       // model.py:31-39
    h->OW = h->T0 - h->SL;
    if (h->OW < 1) { delete h; return fail(nullptr, WNT_ERR_ARG, "sample_size %d does not exceed the receptive field %d", cfg->sample_size, h->rf); }
    h->M = (long)h->N * h->T0;
    h->Mo = (long)h->N * h->OW;
    h->mel_frames = h->C ? h->T / hop : 0;
    build_layout(h);
    wnt_handle *hh = h;
    auto bail = [&](int rc) { g_create_error = hh->err; wnt_destroy(hh); return rc; };
#define A_(p, bytes) do { int r_ = alloc(h, &(p), (bytes)); if (r_) return bail(r_); } while (0)
    const size_t e = h->esz;
    const long M = h->M, Mo = h->Mo;
    const int D2 = 2 * h->D, LD = h->L * h->D;
    if (h->bf) { A_(h->Pc, (size_t)h->n_params * 2); }
    if (!h->scalar) A_(h->ids, (size_t)h->N * h->T * sizeof(int32_t));
    if (h->C) {
        h->U.assign(h->cfg.n_upsample + 1, nullptr);
        h->dU.assign(h->cfg.n_upsample + 1, nullptr);
        long Ti = h->mel_frames;
        for (int i = 1; i < h->cfg.n_upsample; ++i) {
            Ti *= h->cfg.upsample_factor[i - 1];
            A_(h->U[i], (size_t)h->N * Ti * h->C * 4);
            A_(h->dU[i], (size_t)h->N * Ti * h->C * 4);
        }
        A_(h->LC, (size_t)M * h->C * e);
        A_(h->dLC32, (size_t)M * h->C * 4);
    }
    h->X.assign(h->L, nullptr);
    h->TS.assign(h->L, nullptr);
    h->fused = h->bf && h->R == 128 && h->D == 128 && h->C <= 128 && getenv("WNT_NO_FUSED") == nullptr;
    A_(h->Xall, (size_t)h->L * M * h->R * e);
    for (int l = 0; l < h->L; ++l) {
        h->X[l] = (char *)h->Xall + (size_t)l * M * h->R * e;
        A_(h->TS[l], (size_t)M * D2 * e);
    }
    if (h->fused) {
        A_(h->WfgT, (size_t)h->L * wntf::NFG * wntf::KTOT * 2);
        A_(h->WdT, (size_t)h->L * wntf::ND * wntf::ND * 2);
        A_(h->WdP, (size_t)h->L * wntf::ND * wntf::ND * 2);
        A_(h->WdxP, (size_t)h->L * wntf::ND * 512 * 2);
        A_(h->dXp[0], (size_t)M * h->R * 2);
        A_(h->dXp[1], (size_t)M * h->R * 2);
        h->fused_bwd = getenv("WNT_NO_FUSED_BWD") == nullptr;
        h->causal_gemm = h->scalar && h->ifw % 4 == 0 && getenv("WNT_NO_CAUSAL_GEMM") == nullptr;
        if (h->causal_gemm) {
            A_(h->Wcol, (size_t)M * 2 * h->ifw * 2);
            A_(h->WcDup, (size_t)2 * h->ifw * h->R * 2);
            A_(h->dWcTmp, (size_t)2 * h->ifw * h->R * 4);
        }
        A_(h->fused_err, 16);
    }
    A_(h->Zs, (size_t)Mo * LD * e);
    A_(h->dZs, (size_t)Mo * LD * e);
    A_(h->Z, (size_t)M * h->D * e);
    A_(h->T1, (size_t)Mo * h->S * e);
    A_(h->T2, (size_t)Mo * h->S * e);
    A_(h->dC1, (size_t)Mo * h->S * e);
    A_(h->dTot, (size_t)Mo * h->S * e);
    A_(h->dY, (size_t)Mo * h->OP * e);
    // the cuBLASLt path keeps fp32 pre-activations and an fp32 dx master; the tcgen05 path needs neither
    const bool lt_fwd = !h->fused, lt_bwd = !(h->fused && getenv("WNT_NO_FUSED_BWD") == nullptr);
    if (lt_bwd) A_(h->dXb, (size_t)M * h->R * e);
    A_(h->dFG, (size_t)M * D2 * e);
    if (lt_fwd) A_(h->FG32, (size_t)M * D2 * 4);
    A_(h->TOT, (size_t)Mo * h->S * 4);
    A_(h->dT32, (size_t)Mo * h->S * 4);
    A_(h->Y, (size_t)Mo * h->OP * 4);
    if (lt_bwd) {
        A_(h->dX32, (size_t)M * h->R * 4);
        A_(h->dZ32, (size_t)M * h->D * 4);
    }
    A_(h->GCB, (size_t)h->L * h->N * D2 * 4);
    A_(h->SB, (size_t)h->L * h->N * D2 * 4);
    A_(h->bsum, (size_t)h->S * 4);
    A_(h->dbs, (size_t)h->S * 4);
    A_(h->acc, 4 * sizeof(double));
    A_(h->lt_ws, h->lt_ws_bytes);
#undef A_
    if (cublasLtCreate(&h->lt) != CUBLAS_STATUS_SUCCESS) { h->err = "cublasLtCreate failed"; return bail(WNT_ERR_CUBLAS); }
    if (h->fused) { int r_ = setup_fused(h); if (r_) return bail(r_); }
    if (h->scalar) {
        const size_t sm = ((size_t)h->ifw * h->R + 128 + h->ifw) * sizeof(float);
        if (sm > 48 * 1024) {
            cudaFuncSetAttribute(causal_fwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
            cudaFuncSetAttribute(causal_fwd_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        }
    }
    *out = h;
    return WNT_OK;
}

void wnt_destroy(wnt_handle *h) {
    if (!h) return;
    for (auto &kv : h->plans) {
        Plan &p = kv.second;
        if (p.a) cublasLtMatrixLayoutDestroy(p.a);
        if (p.b) cublasLtMatrixLayoutDestroy(p.b);
        if (p.c) cublasLtMatrixLayoutDestroy(p.c);
        if (p.d) cublasLtMatrixLayoutDestroy(p.d);
        if (p.op) cublasLtMatmulDescDestroy(p.op);
    }
    if (h->lt) cublasLtDestroy(h->lt);
    auto fr = [](void *p) { if (p) cudaFree(p); };
    if (h->bf) fr(h->Pc);   // fp32: Pc aliases the caller's parameter buffer
    for (auto p : h->U) fr(p);
    for (auto p : h->dU) fr(p);
    fr(h->Wcol); fr(h->WcDup); fr(h->dWcTmp); fr(h->ids);
    fr(h->Xall); fr(h->WfgT); fr(h->WdT); fr(h->WdP); fr(h->WdxP); fr(h->dXp[0]); fr(h->dXp[1]); fr(h->fused_err);
    for (auto p : h->TS) fr(p);
    void *all[] = {h->LC, h->dLC32, h->Zs, h->dZs, h->Z, h->T1, h->T2, h->dC1, h->dTot, h->dY, h->dXb, h->dFG, h->FG32, h->TOT, h->dT32, h->Y,
                   h->dX32, h->dZ32, h->GCB, h->SB, h->bsum, h->dbs, h->acc, h->lt_ws};
    for (void *p : all) fr(p);
    delete h;
}

int wnt_get_info(const wnt_handle *h, wnt_info *info) {
    if (!h || !info) return WNT_ERR_ARG;
    info->n_params = h->n_params;
    info->n_weights = h->n_weights;
    info->n_trainable = h->n_trainable;
    info->workspace_bytes = h->workspace_bytes;
    info->receptive_field = h->rf;
    info->output_width = h->OW;
    info->rows_per_crop = h->T0;
    info->mel_frames = h->mel_frames;
    info->gemm_launches = h->gemm_launches;
    info->kernel_launches = h->kernel_launches;
    info->flops_per_step = h->flops;
    info->fused_launches = h->fused_launches;
    return WNT_OK;
}

int wnt_bind(wnt_handle *h, float *params_dev, float *grads_dev, float *adam_m_dev, float *adam_v_dev, float *ema_dev) {
    if (!h) return WNT_ERR_ARG;
    if (!params_dev || !grads_dev || !adam_m_dev || !adam_v_dev || !ema_dev) return fail(h, WNT_ERR_ARG, "null buffer");
    h->P = params_dev; h->Gr = grads_dev; h->Am = adam_m_dev; h->Av = adam_v_dev; h->Ema = ema_dev;
    if (!h->bf) h->Pc = params_dev;
    return refresh_copy(h, 0);
}

int wnt_params_changed(wnt_handle *h, void *stream) {
    if (!h || !h->P) return h ? fail(h, WNT_ERR_STATE, "wnt_bind first") : WNT_ERR_ARG;
    return refresh_copy(h, (cudaStream_t)stream);
}

int wnt_set_tensor(wnt_handle *h, int which, const char *name, const float *host, int64_t n) {
    if (!h || !name || !host) return h ? fail(h, WNT_ERR_ARG, "null argument") : WNT_ERR_ARG;
    float *base = buf_of(h, which);
    if (!base) return fail(h, WNT_ERR_STATE, "buffer %d is not bound (wnt_bind first)", which);
    auto it = h->views.find(name);
    if (it == h->views.end()) return fail(h, WNT_ERR_ARG, "unknown variable '%s'", name);
    const View &v = it->second;
    if (n != v.size()) return fail(h, WNT_ERR_ARG, "variable '%s' has %lld elements, got %lld", name, (long long)v.size(), (long long)n);
    for (int o = 0; o < v.outer; ++o)
        CK(cudaMemcpy2D(base + v.base + o * v.outer_stride, (size_t)v.ld * 4, host + (size_t)o * v.rows * v.cols, (size_t)v.cols * 4, (size_t)v.cols * 4,
                        v.rows, cudaMemcpyHostToDevice));
    return WNT_OK;
}

int64_t wnt_get_tensor(wnt_handle *h, int which, const char *name, float *host, int64_t n) {
    if (!h || !name) return WNT_ERR_ARG;
    float *base = buf_of(h, which);
    if (!base) return fail(h, WNT_ERR_STATE, "buffer %d is not bound (wnt_bind first)", which);
    auto it = h->views.find(name);
    if (it == h->views.end()) return fail(h, WNT_ERR_ARG, "unknown variable '%s'", name);
    const View &v = it->second;
    if (!host || n < v.size()) return v.size();
    CK(cudaDeviceSynchronize());
    for (int o = 0; o < v.outer; ++o)
        CK(cudaMemcpy2D(host + (size_t)o * v.rows * v.cols, (size_t)v.cols * 4, base + v.base + o * v.outer_stride, (size_t)v.ld * 4, (size_t)v.cols * 4,
                        v.rows, cudaMemcpyDeviceToHost));
    return v.size();
}

int64_t wnt_variable_names(const wnt_handle *h, char *out, int64_t n) {
    if (!h) return WNT_ERR_ARG;
    std::string s;
    for (size_t i = 0; i < h->names.size(); ++i) {
        if (i) s += '\n';
        s += h->names[i];
    }
    if (out && n > 0) {
        const size_t c = std::min((size_t)(n - 1), s.size());
        memcpy(out, s.data(), c);
        out[c] = 0;
    }
    return (int64_t)s.size();
}

int wnt_loss_and_grads(wnt_handle *h, const float *wav_dev, const float *mel_dev, const int32_t *gc_ids_dev, float l2_strength,
                       float *loss_dev, void *stream) {
    if (!h) return WNT_ERR_ARG;
    if (!h->P) return fail(h, WNT_ERR_STATE, "wnt_bind first");
    if (!wav_dev || !loss_dev) return fail(h, WNT_ERR_ARG, "null wav / loss pointer");
    if (h->C && !mel_dev) return fail(h, WNT_ERR_ARG, "local conditioning is configured: mel_dev is required");
    if (h->G && !gc_ids_dev) return fail(h, WNT_ERR_ARG, "global conditioning is configured: gc_ids_dev is required");
    h->count_flops = h->flops == 0;
    int rc = h->bf ? step_t<bf16>(h, wav_dev, mel_dev, gc_ids_dev, l2_strength, loss_dev, (cudaStream_t)stream)
                   : step_t<float>(h, wav_dev, mel_dev, gc_ids_dev, l2_strength, loss_dev, (cudaStream_t)stream);
    h->count_flops = false;
    return rc;
}

int wnt_apply(wnt_handle *h, const wnt_adam *a, void *stream) {
    if (!h || !a) return WNT_ERR_ARG;
    if (!h->P) return fail(h, WNT_ERR_STATE, "wnt_bind first");
    if (a->t < 1) return fail(h, WNT_ERR_ARG, "t is the 1-based update count");
    return h->bf ? apply_t<bf16>(h, a, (cudaStream_t)stream) : apply_t<float>(h, a, (cudaStream_t)stream);
}

int64_t wnt_debug_get(wnt_handle *h, const char *name, float *host, int64_t n) {
    if (!h || !name) return WNT_ERR_ARG;
    if (!h->have_step) return fail(h, WNT_ERR_STATE, "no step has run");
    CK(cudaDeviceSynchronize());
    const std::string s(name);
    if (s == "raw_output") {
        const int64_t tot = (int64_t)h->Mo * h->O;
        if (host && n >= tot) CK(cudaMemcpy2D(host, (size_t)h->O * 4, h->Y, (size_t)h->OP * 4, (size_t)h->O * 4, h->Mo, cudaMemcpyDeviceToHost));
        return tot;
    }
    const void *src = nullptr;
    int64_t tot = 0;
    if (s == "lc" && h->C) { src = h->LC; tot = (int64_t)h->M * h->C; }
    else if (s.size() > 1 && s[0] == 'x') {
        const int l = atoi(s.c_str() + 1);
        if (l < 0 || l >= h->L) return fail(h, WNT_ERR_ARG, "no such layer");
        src = h->X[l]; tot = (int64_t)h->M * h->R;
    } else return fail(h, WNT_ERR_ARG, "unknown intermediate '%s'", name);
    if (!host || n < tot) return tot;
    if (!h->bf) { CK(cudaMemcpy(host, src, (size_t)tot * 4, cudaMemcpyDeviceToHost)); return tot; }
    std::vector<uint16_t> tmp((size_t)tot);
    CK(cudaMemcpy(tmp.data(), src, (size_t)tot * 2, cudaMemcpyDeviceToHost));
    for (int64_t i = 0; i < tot; ++i) {
        const uint32_t u = (uint32_t)tmp[(size_t)i] << 16;
        memcpy(host + i, &u, 4);
    }
    return tot;
}

}  // extern "C"

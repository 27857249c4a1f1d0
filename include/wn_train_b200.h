/*
 * wn_train_b200.h -- C ABI of libwn_train_b200.so: one WaveNet vocoder TRAINING step on B200 (sm_100a)
 * (SURVEY.md section 8f "next-3", BASELINE configs[3]).
 *
 * The reference (hccho2/Tacotron-Wavenet-Vocoder-Korean) has no FFI: the step is `sess.run([global_step, loss,
 * optimize])` (train_vocoder.py:169) over the graph built by WaveNetModel.add_loss (wavenet/model.py:247-312) and
 * add_optimizer (wavenet/model.py:314-346).  Each entry point names what it stands in for; INTEGRATION.md shows the
 * ctypes binding.  Conventions are those of wn_b200.h: plain C types, 0 = OK / negative error code + wnt_last_error(),
 * no C++ exception crosses the ABI, "dev" pointers are caller-owned CUDA device memory, calls on one handle are
 * stream-ordered by the caller.  There is no CPU fallback.
 *
 * Parameters, gradients and optimizer state live in FLAT fp32 device buffers owned by the caller (torch tensors), so
 * that the data-parallel gradient all-reduce is one NCCL call on one buffer; the library keeps a compute-dtype copy of
 * the parameters, activations and workspaces.  The flat layout is internal (filter|gate kernels are stored side by side
 * for one GEMM); wnt_set_tensor / wnt_get_tensor translate from / to the TF variable names and shapes of SURVEY.md
 * Appendix B.
 */
#ifndef WN_TRAIN_B200_H
#define WN_TRAIN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WNT_OK 0
#define WNT_ERR_ARG (-1)
#define WNT_ERR_STATE (-2)
#define WNT_ERR_CUDA (-3)
#define WNT_ERR_CUBLAS (-4)
#define WNT_ERR_UNSUPPORTED (-5)

#define WNT_MAX_LAYERS 64
#define WNT_MAX_UPSAMPLE 8

#define WNT_DTYPE_BF16 0   /* bf16 activations / weights copy, fp32 accumulation, fp32 master weights (BASELINE configs[3]) */
#define WNT_DTYPE_FP32 1   /* everything fp32 (validation against the fp32 oracle) */

/* WaveNetModel(...) constructor arguments (wavenet/model.py:8-10) + the crop length of the data feeder. */
typedef struct wnt_config {
    int32_t batch_size;
    int32_t n_layers, dilations[WNT_MAX_LAYERS];
    int32_t residual_channels, dilation_channels, skip_channels;   /* powers of two, 8..512 */
    int32_t out_channels, quantization_channels;
    int32_t use_biases, scalar_input, initial_filter_width;        /* scalar_input 1: MoL head; 0: mu-law one-hot input + softmax head */
    int32_t gc_channels, gc_cardinality;                           /* 0 = no global conditioning */
    int32_t lc_channels;                                           /* 0 = no local conditioning; else multiple of 8 */
    int32_t n_upsample, upsample_factor[WNT_MAX_UPSAMPLE];
    int32_t sample_size;                                           /* samples per crop (hparams.sample_size / max_time_steps) */
    int32_t dtype;                                                 /* WNT_DTYPE_* */
} wnt_config;

typedef struct wnt_info {
    int64_t n_params;             /* length of the flat buffers (floats), including alignment padding */
    int64_t n_weights;            /* the first n_weights floats are kernels (L2-regularised), the rest biases */
    int64_t n_trainable;          /* number of real trainable scalars (= sum of TF variable sizes) */
    int64_t workspace_bytes;
    int32_t receptive_field, output_width, rows_per_crop;
    int32_t mel_frames;           /* local-condition frames per crop = sample_size / prod(upsample_factor) */
    int64_t gemm_launches, kernel_launches;   /* launched by this handle so far */
    double flops_per_step;        /* tensor-core GEMM flops of one forward+backward */
    int64_t fused_launches;       /* launches of the fused tcgen05 layer kernel (R = D = 128, bf16); 0 = cuBLASLt path only */
} wnt_info;

typedef struct wnt_handle wnt_handle;

/* WaveNetModel(train_mode=True, ...) as train_vocoder.py:100-116 constructs it (wavenet/model.py:8-30) */
int wnt_create(const wnt_config *cfg, wnt_handle **out);
void wnt_destroy(wnt_handle *h);
const char *wnt_last_error(const wnt_handle *h);       /* h may be NULL: last create error */
int wnt_get_info(const wnt_handle *h, wnt_info *info);

/* Binds the caller's flat fp32 device buffers (each info.n_params floats, zero-initialised by the caller):
 * tf.global_variables_initializer + the Adam slots + the EMA shadows (wavenet/model.py:30,325,346). */
int wnt_bind(wnt_handle *h, float *params_dev, float *grads_dev, float *adam_m_dev, float *adam_v_dev, float *ema_dev);

/* Saver.restore / Saver.save by TF variable name (utils/__init__.py:62-90).  `which`: 0 params, 1 grads, 2 EMA shadow,
 * 3 Adam m, 4 Adam v.  `host` holds n floats in the TF shape.  After writing parameters (which = 0) call
 * wnt_params_changed once to refresh the compute-dtype copy. */
int wnt_set_tensor(wnt_handle *h, int which, const char *name, const float *host, int64_t n);
int64_t wnt_get_tensor(wnt_handle *h, int which, const char *name, float *host, int64_t n);   /* returns the size */
/* Variable names, '\n'-separated, in tf.trainable_variables() order; returns the string length (copies at most n). */
int64_t wnt_variable_names(const wnt_handle *h, char *out, int64_t n);
/* Copies the EMA shadows over the parameters or back is the caller's business (plain tensor copy on the flat buffers);
 * after any direct write to params_dev call this to refresh the compute-dtype copy. */
int wnt_params_changed(wnt_handle *h, void *stream);

/* net.add_loss(input_batch, local_condition, global_condition_batch, l2_regularization_strength) (wavenet/model.py:247-312,
 * train_vocoder.py:122) evaluated with its gradients (optimizer.compute_gradients, wavenet/model.py:327):
 *   wav_dev (N, sample_size) fp32 in [-1,1]; mel_dev (N, mel_frames, lc_channels) fp32 or NULL; gc_ids_dev (N) int32 or
 *   NULL; l2_strength < 0 means None.  Writes the scalar loss to loss_dev[0] and d loss / d params to the bound grads.
 * Asynchronous on `stream`. */
int wnt_loss_and_grads(wnt_handle *h, const float *wav_dev, const float *mel_dev, const int32_t *gc_ids_dev,
                       float l2_strength, float *loss_dev, void *stream);

/* optimizer.apply_gradients + ema.apply (wavenet/model.py:325-346): `t` = 1-based number of this update,
 * learning_rate already decayed (tf.train.exponential_decay), grad_scale multiplies every gradient first (1/world
 * after a sum all-reduce), clip_norm > 0 enables tf.clip_by_global_norm(gradients, clip_norm). */
typedef struct wnt_adam {
    float learning_rate, beta1, beta2, epsilon, ema_decay, grad_scale, clip_norm;
    int32_t t;
} wnt_adam;
int wnt_apply(wnt_handle *h, const wnt_adam *a, void *stream);

/* Test hook: copies an intermediate of the LAST wnt_loss_and_grads to the host as fp32 ("raw_output" (N*output_width,
 * out_channels), "lc" (N*rows_per_crop, lc_channels), "x<l>" (N*rows_per_crop, R)); synchronises the device.  Returns
 * the number of floats available (copies at most n). */
int64_t wnt_debug_get(wnt_handle *h, const char *name, float *host, int64_t n);

#ifdef __cplusplus
}
#endif
#endif

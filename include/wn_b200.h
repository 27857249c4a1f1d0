/*
 * wn_b200.h -- C ABI of libwn_b200.so: the B200 (sm_100a) WaveNet autoregressive sample loop.
 *
 * The reference (hccho2/Tacotron-Wavenet-Vocoder-Korean) has no FFI: the path sits behind Python
 * class / CLI signatures.  Each entry point below names the reference interface it stands in for
 * (paths relative to the reference root).  INTEGRATION.md shows the ctypes binding a maintainer of
 * the reference would add.
 *
 * Conventions: plain C types only, no torch / C++ types cross the boundary.  Every call returns
 * WN_OK (0) or a negative error code; wn_last_error(h) gives the message.  No C++ exception crosses
 * the ABI.  "dev" pointers are CUDA device pointers owned by the caller (e.g. torch tensors' data_ptr())
 * and only borrowed for the duration of the stream-ordered call; the library owns its packed weight
 * images, mailboxes and dilation-queue rings.  A handle is re-entrant per stream: calls on one handle
 * must be stream-ordered by the caller; different handles are independent.  There is no CPU fallback:
 * every compute entry point fails with WN_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef WN_B200_H
#define WN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WN_OK 0
#define WN_ERR_ARG (-1)        /* bad argument / unsupported configuration */
#define WN_ERR_STATE (-2)      /* call order (e.g. generate before finalize, missing weights) */
#define WN_ERR_CUDA (-3)       /* CUDA runtime error, no device, launch failure */
#define WN_ERR_TIMEOUT (-4)    /* the persistent kernel aborted on its in-kernel watchdog */

#define WN_MAX_LAYERS 256
#define WN_MAX_UPSAMPLE 8
#define WN_MAX_BATCH 32

/* Construction arguments of wavenet.model.WaveNetModel.__init__ (wavenet/model.py:8-10), train_mode=False.
 * force_M / force_Mt = 0 lets the library pick the layer / tail split (tests pin them to cover
 * every evaluation plan). */
typedef struct wn_config {
    int32_t batch;                     /* batch_size: utterances generated together (<= WN_MAX_BATCH) */
    int32_t n_layers;                  /* len(dilations) */
    int32_t filter_width;              /* must be 2 (hparams.py:59) */
    int32_t residual_channels;
    int32_t dilation_channels;
    int32_t skip_channels;
    int32_t quantization_channels;
    int32_t out_channels;
    int32_t use_biases;
    int32_t scalar_input;
    int32_t initial_filter_width;
    int32_t gc_channels;               /* global_condition_channels or 0 */
    int32_t gc_cardinality;            /* global_condition_cardinality or 0 */
    int32_t lc_channels;               /* local_condition_channels or 0 */
    int32_t n_upsample;
    int32_t upsample_factor[WN_MAX_UPSAMPLE];
    int32_t dilations[WN_MAX_LAYERS];
    int32_t force_M;
    int32_t force_Mt;
    int32_t flags;                     /* WN_FLAG_* */
} wn_config;

#define WN_FLAG_GENERIC_KERNEL 1       /* never pick a compile-time specialised kernel instantiation */
#define WN_FLAG_NO_DIE_AWARE 2         /* single-homed mailboxes (no die calibration); also env WN_NO_DIE_AWARE */
#define WN_FLAG_NO_CLUSTER 4           /* never use the thread-block-cluster / DSMEM layer kernel; also env WN_NO_CLUSTER */
#define WN_FLAG_FAST_ACT 8             /* scalar-input (mixture-of-logistics) models on the cluster path: tanh / sigmoid through
                                          ex2.approx + rcp.approx (logits within 1e-4 of the pinned arithmetic, not bit-identical) */

/* Floating-point evaluation order implemented by the kernel (DESIGN.md "Pinned arithmetic").
 * Field meaning is identical to oracle/wn_oracle.c's orc_plan. */
typedef struct wn_plan {
    int32_t M, Mt, t_cur, t_old, t_lc, t_gc, t_dense, t_skip, t_post1, t_post2, t_causal;
} wn_plan;

typedef struct wn_info {
    int32_t grid;                      /* CTAs of the persistent kernel (1 per SM) */
    int32_t threads;                   /* threads per CTA */
    int32_t M, Mt;
    int32_t smem_bytes_layer, smem_bytes_tail, smem_bytes_sampler;
    int32_t sm_count;
    int64_t p_hot;                     /* weights touched per step, gc folded into biases */
    int64_t weights_in_smem;           /* floats resident in shared memory over the whole grid */
    int64_t weights_in_global;         /* floats that overflowed to L2/HBM */
    int64_t kernel_launches;           /* kernels launched by this handle so far */
    int32_t static_shape;              /* 0: runtime-shaped kernel; 1: cfg2 shape, 2: cfg1 shape, 3: hparams.py default shape */
    int32_t die_aware;                 /* 1: mailboxes are dual-homed (one L2 copy per die), set by wn_finalize */
    int32_t cluster_path;              /* 1: layer chain runs in 8-CTA clusters with DSMEM hops (set by wn_finalize) */
    int32_t fast_act;                  /* 1: WN_FLAG_FAST_ACT is in effect */
} wn_info;

typedef struct wn_handle wn_handle;

/* WaveNetModel(...) constructor, wavenet/model.py:8-30. */
int wn_create(const wn_config *cfg, wn_handle **out);
void wn_destroy(wn_handle *h);
const char *wn_last_error(const wn_handle *h);      /* h may be NULL: last create error */

/* tf.train.Saver.restore of the non-queue variables (generate.py:157-161, utils/__init__.py:75-90).
 * `name` is the TF variable name, `data` a HOST pointer to n floats in TF layout. */
int wn_set_weight(wn_handle *h, const char *name, const float *data, int64_t n);
/* Packs the weights into per-SM shared-memory images and uploads them.  Must follow the last
 * wn_set_weight and precede any compute call. */
int wn_finalize(wn_handle *h);

/* Topology / evaluation plan / shared-memory budget the library would choose for `cfg` on a device with
 * sm_count SMs (0 = 148, a B200).  Pure host function: no device needed. */
int wn_plan_config(const wn_config *cfg, int sm_count, wn_plan *plan, wn_info *info);

int wn_get_plan(const wn_handle *h, wn_plan *plan);
int wn_get_info(const wn_handle *h, wn_info *info);

/* WaveNetModel.calculate_receptive_field, wavenet/model.py:31-39. */
int wn_receptive_field(int filter_width, const int32_t *dilations, int n, int scalar_input,
                       int initial_filter_width);

/* WaveNetModel.create_upsample, wavenet/model.py:102-111 (evaluated once per utterance, generate.py:155,200).
 * mel_dev (rows, t_mel, lc_channels) -> out_dev (rows, t_mel*prod(upsample_factor), lc_channels). */
int wn_upsample(wn_handle *h, const float *mel_dev, int rows, int t_mel, float *out_dev, void *stream);

/* The hot loop of generate.py:202-233 over sess.run(predict_proba_incremental) (wavenet/model.py:215-245)
 * including the sample draw (wavenet/mixture.py:84-114 or generate.py:219-231), as ONE persistent kernel. */
typedef struct wn_generate_args {
    int32_t rows;                 /* utterances in this call, 1..cfg.batch */
    int32_t T;                    /* network steps per row (max over rows when T_row != NULL) */
    const int32_t *T_row;         /* HOST, optional per-row step counts (<= T) */
    int32_t n_forced;             /* >= 1: x_in(t) = forced[row][t] for t < n_forced, else the sample drawn at t-1
                                     (1 = free running from an initial sample, generate.py:184-192;
                                      len(seed) = priming, generate.py:177-180; T = teacher forcing) */
    const float *forced_dev;      /* (rows, n_forced) fp32; mu-law ids are stored as floats (generate.py:190) */
    const float *lc_dev;          /* (rows, t_lc, lc_channels) upsampled local condition, or NULL */
    int32_t t_lc;
    int32_t lc_shift;             /* step t pushes row t - lc_shift into the lc queue (zeros if out of range) */
    const int32_t *gc_ids;        /* HOST (rows) speaker ids or NULL */
    const void *uniforms_dev;     /* scalar_input: (rows, T, out_channels/3 + 1) fp32 in (1e-5, 1-1e-5)
                                     one-hot:      (rows, T) fp64 in [0,1) */
    float temperature;            /* generate.py:51; only the mu-law path uses it */
    float *out_samples_dev;       /* (rows, T) fp32 */
    float *out_logits_dev;        /* optional (rows, T, out_dim) raw conv2 output, or NULL */
    const float *mel_dev;         /* optional (rows, t_mel, lc_channels) mel frames: create_upsample (wavenet/model.py:102-111) is
                                     then evaluated inside the generation kernel, frame by frame, instead of being materialised
                                     (generate.py:155,200); lc_dev / t_lc are ignored, lc_shift keeps its meaning */
    int32_t t_mel;
} wn_generate_args;

/* Stream-ordered and asynchronous: returns after the launch. */
int wn_generate(wn_handle *h, const wn_generate_args *args, void *stream);

/* ONE step of the network with persistent device queues: what the reference's per-sample loop evaluates through
 * sess.run(predict_proba_incremental) (generate.py:202-211, wavenet/model.py:215-245).  A wn_state holds the causal,
 * local-condition and dilation queues of model.py:49-64 for `rows` utterances, zeroed like queue_initializer
 * (generate.py:163); every wn_step advances them by one position, so a call costs O(1) in the number of steps taken.
 * A loop of wn_step calls fed with its own draws reproduces wn_generate bit for bit (same evaluation plan). */
typedef struct wn_state wn_state;
int wn_state_create(wn_handle *h, int rows, wn_state **out);
int wn_state_reset(wn_handle *h, wn_state *s, void *stream);          /* queue_initializer */
void wn_state_destroy(wn_state *s);
typedef struct wn_step_args {
    int32_t rows;                 /* must equal the state's rows */
    const float *x_in_dev;        /* (rows) network input: previous sample, or mu-law id stored as float (generate.py:204) */
    const float *lc_row_dev;      /* (rows, lc_channels) upsampled[:, step, :] (generate.py:211) or NULL */
    const int32_t *gc_ids;        /* HOST (rows) or NULL */
    const void *uniforms_dev;     /* optional: (rows, out_channels/3 + 1) fp32 [scalar input] or (rows) fp64 [one-hot]: also draw */
    float temperature;            /* generate.py:51, one-hot models, used with uniforms_dev */
    float *out_logits_dev;        /* optional (rows, out_dim) raw conv2 output */
    float *out_probs_dev;         /* one-hot models: optional (rows, Q) float32(softmax(float64(logits))), the node's return value */
    float *out_sample_dev;        /* optional (rows): the drawn sample / id when uniforms_dev is given */
} wn_step_args;
int wn_step(wn_handle *h, wn_state *s, const wn_step_args *args, void *stream);

/* Synchronises `stream` and reports WN_ERR_TIMEOUT if the kernel's watchdog aborted the last launch. */
int wn_sync_check(wn_handle *h, void *stream);

/* Same, through HOST buffers: mel in, waveform out, host<->device copies inside (the call a
 * generate.py user effectively makes: np.load(mel) ... save_wav).  mel_host (rows, t_mel, lc_channels) is
 * upsampled on the device (args->lc_dev is then ignored), or NULL.  Pointer fields of `args` named
 * *_dev are HOST pointers here.  Synchronous; includes the watchdog check. */
int wn_generate_host(wn_handle *h, const wn_generate_args *args, const float *mel_host, int t_mel);

/* wavenet.ops.mu_law_encode / mu_law_decode, wavenet/ops.py:22-47, element-wise on device buffers. */
int wn_mu_law_encode(const float *audio_dev, int64_t n, int quantization_channels, int32_t *out_dev, void *stream);
int wn_mu_law_decode(const float *in_dev, int64_t n, int quantization_channels, int quantization,
                     float *out_dev, void *stream);

/* wavenet/mixture.py:84-114 sample_from_discretized_mix_logistic(y, log_scale_min) on device tensors: y_dev (rows, 3*nr_mix)
 * network outputs [logit | mean | log_scale], uniforms_dev (rows, nr_mix + 1) in (1e-5, 1 - 1e-5) in place of TF's unseeded
 * tf.random_uniform (nr_mix for the Gumbel-max, one for the logistic) -> out_dev (rows) in [-1, 1].  Same pinned arithmetic
 * as the draw inside wn_generate / wn_step (bit-identical to the oracle). */
int wn_mol_sample(const float *y_dev, const float *uniforms_dev, int64_t rows, int nr_mix, float log_scale_min,
                  float *out_dev, void *stream);
/* wavenet/mixture.py:27-81 discretized_mix_logistic_loss(y_hat, y, num_class, log_scale_min, reduce), forward value only (the
 * training step evaluates it fused with its gradient, wn_train_b200.h): y_hat_dev (rows, 3*nr_mix), y_dev (rows) targets in
 * [-1, 1] -> loss_out_dev (rows) per-step losses (reduce=False) and / or sum_out_dev (one double: reduce=True); either may be NULL. */
int wn_mol_loss(const float *y_hat_dev, const float *y_dev, int64_t rows, int nr_mix, int num_class, float log_scale_min,
                float *loss_out_dev, double *sum_out_dev, void *stream);

/* utils/audio.py:69-75 melspectrogram(wav, hparams) ("next-2"): pre-emphasis -> librosa.stft
 * (center, reflect pad, periodic Hann zero-padded to fft_size) -> |D| -> Slaney mel basis -> 20*log10(max(1e-5, .))
 * - ref_level_db -> symmetric normalisation + clip.  Fields are the reference's hparams.py:18-34.
 * wav_dev (rows, n) fp32 -> out_dev (rows, 1 + n / hop_size, num_mels) fp32 (the reference returns the
 * transpose, (num_mels, frames), for one row).  On error wn_last_error(NULL) has the message. */
typedef struct wn_mel_config {
    int32_t sample_rate, fft_size, hop_size, win_size, num_mels;
    int32_t preemphasize;
    float preemphasis, min_level_db, ref_level_db, max_abs_value;
} wn_mel_config;
int wn_melspectrogram(const float *wav_dev, int rows, int64_t n, const wn_mel_config *mc, float *out_dev, void *stream);

#ifdef __cplusplus
}
#endif
#endif

/*
 * taco_b200.h -- C ABI of libtaco_b200.so: the B200 (sm_100a) Tacotron text->mel inference path
 * ("next-1").
 *
 * The reference (hccho2/Tacotron-Wavenet-Vocoder-Korean) has no FFI: this path sits behind
 * tacotron.Tacotron(hparams).initialize(...) (tacotron/tacotron.py:31-235) and synthesizer.Synthesizer
 * (synthesizer.py:30-200).  Each entry point names the reference interface it stands in for (paths relative to
 * the reference root); INTEGRATION.md shows the ctypes binding a maintainer of the reference would add.
 *
 * Conventions are those of wn_b200.h: plain C types only, 0 = OK / negative error code, taco_last_error() for
 * the message, no C++ exception crosses the ABI, "dev" pointers are caller-owned CUDA device memory borrowed
 * for the stream-ordered call, the library owns its packed weights and workspaces.  Calls on one handle must be
 * stream-ordered by the caller.  There is no CPU fallback.
 */
#ifndef TACO_B200_H
#define TACO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TACO_OK 0
#define TACO_ERR_ARG (-1)
#define TACO_ERR_STATE (-2)
#define TACO_ERR_CUDA (-3)
#define TACO_ERR_TIMEOUT (-4)    /* the persistent decoder kernel aborted on its barrier watchdog */

#define TACO_MAX_PRENET 4
#define TACO_MAX_PROJ 4
#define TACO_MAX_DEC_LAYERS 4

/* hparams.attention_type (tacotron/tacotron.py:126-145); the other TF attention classes are not built. */
#define TACO_ATT_BAH_MON 0        /* tf.contrib.seq2seq.BahdanauMonotonicAttention(normalize=False) */
#define TACO_ATT_BAH_MON_NORM 1   /* ... normalize=True: the reference default (hparams.py:140) */
#define TACO_ATT_LOC_SEN 2        /* rnn_wrappers.LocationSensitiveAttention (rnn_wrappers.py:581-726) */

/* The Tacotron fields of hparams.py:124-158 plus what Tacotron.initialize receives. */
typedef struct taco_config {
    int32_t num_symbols;                 /* len(text.symbols.symbols) = 80 */
    int32_t embedding_size;
    int32_t num_speakers;                /* > 1: 'deepvoice' speaker states (tacotron.py:76-84) */
    int32_t speaker_embedding_size;
    int32_t n_enc_prenet, enc_prenet_sizes[TACO_MAX_PRENET];
    int32_t enc_bank_size, enc_bank_channel_size, enc_highway_depth, enc_rnn_size;
    int32_t n_enc_proj, enc_proj_sizes[TACO_MAX_PROJ], enc_proj_width;
    int32_t attention_type, attention_size, attention_state_size;
    int32_t dec_layer_num, dec_rnn_size;
    int32_t n_dec_prenet, dec_prenet_sizes[TACO_MAX_PRENET];
    int32_t post_bank_size, post_bank_channel_size, post_highway_depth, post_rnn_size;
    int32_t n_post_proj, post_proj_sizes[TACO_MAX_PROJ], post_proj_width;
    int32_t reduction_factor, max_iters, num_mels, num_freq;
} taco_config;

typedef struct taco_info {
    int32_t sm_count;
    int32_t dec_grid, dec_threads, dec_smem_bytes;   /* persistent decoder kernel */
    int32_t dec_phases_per_step;                      /* grid-wide barriers per decoder step */
    int32_t rnn_weights_in_smem;                      /* 1: bi-GRU recurrent weights resident in shared memory */
    int64_t n_params;
    int64_t kernel_launches;                          /* kernels launched by this handle so far */
    int64_t workspace_bytes;
    int64_t tc_gemm_launches;                         /* of kernel_launches: tcgen05 (3xTF32) conv / dense GEMMs of the CBHG stacks */
} taco_info;

typedef struct taco_handle taco_handle;

/* Tacotron(hparams) + initialize(..., rnn_decoder_test_mode=True): tacotron/tacotron.py:31-44, synthesizer.py:52-56. */
int taco_create(const taco_config *cfg, taco_handle **out);
void taco_destroy(taco_handle *h);
const char *taco_last_error(const taco_handle *h);      /* h may be NULL: last create error */

/* tf.train.Saver.restore (synthesizer.py:68-70).  `name` is the variable name (tacotron_..._b200/synth.py
 * taco_weight_shapes), `data` a HOST pointer to n floats in TF layout. */
int taco_set_weight(taco_handle *h, const char *name, const float *data, int64_t n);
/* Packs (GRU input/recurrent split, highway H/T interleave, batch-norm scale/shift, per-CTA decoder images)
 * and uploads.  Must follow the last taco_set_weight and precede taco_synthesize. */
int taco_finalize(taco_handle *h);
int taco_get_info(const taco_handle *h, taco_info *info);

/* One sess.run([linear_outputs, alignments, mel_outputs]) of synthesizer.py:129-160: encoder (embedding, prenet,
 * CBHG), the dynamic_decode loop (tacotron.py:200-201) as ONE persistent kernel, post CBHG and the final Dense. */
typedef struct taco_synth_args {
    int32_t N;                        /* sentences in the batch */
    int32_t T_in;                     /* padded token length */
    const int32_t *ids_dev;           /* (N, T_in) token ids, 0 = pad, 1 = EOS (text/symbols.py) */
    const int32_t *lengths;           /* HOST (N): argmax(ids == EOS) + 1 (synthesizer.py:126) */
    const int32_t *speaker_ids;       /* HOST (N) or NULL (= 0) */
    int32_t n_steps;                  /* decoder iterations; 0 = cfg.max_iters (helpers.py:38: the all-zero stop
                                         condition is evaluated by the caller on the outputs) */
    const float *manual_alignments_dev; /* optional (N, n_steps, T_in): is_manual_attention (rnn_wrappers.py:374) */
    float *mel_dev;                   /* out (N, n_steps*reduction_factor, num_mels) */
    float *linear_dev;                /* out (N, n_steps*reduction_factor, num_freq), or NULL to skip post-processing */
    float *alignments_dev;            /* out (N, T_in, n_steps) */
} taco_synth_args;

/* Stream-ordered and asynchronous: returns after the last launch. */
int taco_synthesize(taco_handle *h, const taco_synth_args *args, void *stream);
/* Synchronises `stream` and reports TACO_ERR_TIMEOUT if the decoder's watchdog aborted the last launch. */
int taco_sync_check(taco_handle *h, void *stream);
/* Same through HOST buffers (ids in, mel/linear/alignments out), copies inside, synchronous. */
int taco_synthesize_host(taco_handle *h, const taco_synth_args *args);

/* Test hook: copies an intermediate of the LAST taco_synthesize to the host ("enc_prenet", "enc_bank",
 * "enc_highway_in", "enc_rnn_in", "encoder_out", "keys", "post_bank", "post_highway_in", "post_rnn_in",
 * "post_out"); synchronises the device.  Returns the number of floats available (copies at most n). */
int64_t taco_debug_get(taco_handle *h, const char *name, float *host_out, int64_t n);

#ifdef __cplusplus
}
#endif
#endif

/* ttsb200 — C ABI of the B200-native text->mel->waveform inference path.
 *
 * The reference (nipponjo/tts-arabic-pytorch) has no FFI: its operator boundary for this path is
 * three Python callables. Each entry point below replaces one of them and is what
 * a Python/ctypes (or any other FFI) binding of the reference would call instead:
 *
 *   ttsb_hifigan_*      replaces vocoder.hifigan.models.Generator.forward      (vocoder/hifigan/models.py:111-127)
 *                       and the loader vocoder.load_hifigan                     (vocoder/__init__.py:3-20)
 *   ttsb_fastpitch_*    replaces models.fastpitch.fastpitch.model.FastPitch.infer (models/fastpitch/fastpitch/model.py:351-409)
 *   ttsb_conv1d_*       op-level entry used by the parity tests: one nn.Conv1d / nn.Linear /
 *                       nn.ConvTranspose1d site                                 (vocoder/hifigan/models.py:26-44,98-100)
 *
 * Conventions
 *   - plain C types only; every pointer named d_* is a DEVICE pointer owned by the caller,
 *     h_* is a HOST pointer. `stream` is a cudaStream_t passed as void*.
 *   - the library never allocates on the hot path: callers pass a workspace sized by
 *     *_workspace_bytes(); weights are copied/packed once at *_create().
 *   - every function returns 0 on success; on failure a non-zero code, and ttsb_last_error()
 *     returns a thread-local message. Nothing here falls back to the CPU: without a CUDA device
 *     *_create() fails.
 */
#ifndef TTSB200_H
#define TTSB200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ttsb_hifigan ttsb_hifigan_t;
typedef struct ttsb_fastpitch ttsb_fastpitch_t;
typedef struct ttsb_conv1d ttsb_conv1d_t;
typedef struct ttsb_tacotron2 ttsb_tacotron2_t;

/* A named host fp32 tensor in the reference's own state_dict layout (after weight-norm has been
 * folded, i.e. what remove_weight_norm() leaves: vocoder/__init__.py:19). */
typedef struct {
    const char* name;
    const float* h_data;
    int ndim;
    int64_t shape[4];
} ttsb_tensor_t;

const char* ttsb_last_error(void);
int ttsb_version(void);

/* Runtime switches (also readable from the environment at load: TTSB_CONV_IMPL=tc|simt,
 * TTSB_DESC_MODE=0..3). impl 0 = tcgen05 kernel, 1 = SIMT check kernel. */
int ttsb_set_conv_impl(int impl);
int ttsb_set_desc_mode(int mode);
/* tcgen05 kernel generation: 2 = persistent conv_tc2 (default), 1 = one-tile-per-CTA conv_tc (cross-check). */
int ttsb_set_tc_version(int v);
int ttsb_get_conv_impl(void);
int ttsb_get_desc_mode(void);
/* Number of kernels launched by this library since load (all streams). */
int64_t ttsb_launch_count(void);
/* Debug: device buffer of 256 x 64 int64 that conv_tc_kernel fills with clock64() stamps per CTA
 * (NULL disables). Slot meaning in csrc/conv_tc.cu (tl_mark). Not for production use. */
int ttsb_debug_set_timeline(void* d_buf);
/* Stage profiler (bench.py's per-kernel rooflines): while enabled, the model entry points record a CUDA event on the
 * caller's stream before each stage; ttsb_prof_collect synchronises the device and returns the summed device time per
 * stage tag in milliseconds (tags: csrc/common.cuh ProfTag, mirrored in bench.py; index 0 = untagged gaps). Not thread-safe;
 * meant for a dedicated measurement pass, never for the timed region of a throughput number. */
int ttsb_prof_enable(int on);
int ttsb_prof_n_tags(void);
int ttsb_prof_collect(double* h_ms_by_tag, int n_tags);
/* Device-side error flag raised by bounded mbarrier waits (0 = none). Synchronises the device. */
int ttsb_device_error_flag(int* h_flag);

/* ---------------------------------------------------------------------------------------------
 * HiFi-GAN V1 generator (config.json: upsample_rates/kernel_sizes, resblock kernel sizes and
 * dilations, upsample_initial_channel; resblock type "1" only).
 * --------------------------------------------------------------------------------------------- */
typedef struct {
    int num_mels;                 /* 80 */
    int upsample_initial_channel; /* 512 */
    int num_upsamples;            /* 4 */
    int upsample_rates[8];
    int upsample_kernel_sizes[8];
    int num_kernels;              /* 3 */
    int resblock_kernel_sizes[8];
    int resblock_dilations[8][3];
} ttsb_hifigan_config_t;

int ttsb_hifigan_create(const ttsb_hifigan_config_t* cfg, const ttsb_tensor_t* weights, int n_weights,
                        int device, ttsb_hifigan_t** out);
void ttsb_hifigan_destroy(ttsb_hifigan_t* h);
int ttsb_hifigan_hop(const ttsb_hifigan_t* h); /* samples per mel frame (256) */
size_t ttsb_hifigan_workspace_bytes(const ttsb_hifigan_t* h, int B, int T);
/* mel -> waveform for a padded batch. Exactly one of d_mel_f32 ([B,num_mels,T] fp32, the
 * reference layout) or d_mel_cl ([B,T,128] fp16 channel-last, rows >= len zero) is non-NULL.
 * d_lens [B] int32 frames per utterance (NULL = all T): every layer treats frames beyond an
 * utterance's length as the zero padding the reference's per-utterance call would see
 * (models/fastpitch/networks.py:340-345). d_wav: [B, T*hop] fp32, zero beyond len*hop.
 * h_lens (optional): a HOST copy of d_lens. With it the batch is processed in chunks that each run at their own
 * longest utterance instead of T, so the padding of a mixed-length batch is not computed (results are identical). */
int ttsb_hifigan_forward(ttsb_hifigan_t* h, const float* d_mel_f32, const void* d_mel_cl,
                         const int32_t* d_lens, const int32_t* h_lens, int B, int T, float* d_wav, void* d_workspace,
                         size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * FastPitch (net_config of models/fastpitch/__init__.py:3-41; 1 attention head of 64).
 * The call is split where the reference has data-dependent host decisions:
 *   encode     ids -> encoder -> duration & pitch predictors        (model.py:364-371)
 *   condition  (transformed) pitch -> pitch/energy embedding -> durations -> frame counts
 *                                                                    (model.py:382-403, 68-79)
 *   decode     length regulation -> decoder -> mel projection        (model.py:81-86, 405-409)
 * Between condition and decode the caller reads max(dec_lens) — the same host sync the reference
 * has at model.py:76.
 * --------------------------------------------------------------------------------------------- */
typedef struct {
    int n_mel_channels;  /* 80 */
    int n_symbols;
    int d_model;         /* 384 */
    int n_layers_enc, n_layers_dec;
    int d_head;          /* 64, one head */
    int d_inner;         /* 1536 */
    int conv_kernel;     /* 3 */
    int pred_filter;     /* 256 */
    int pred_kernel;     /* 3 */
    int energy_conditioning;
    int n_speakers;
    float speaker_emb_weight;
} ttsb_fastpitch_config_t;

int ttsb_fastpitch_create(const ttsb_fastpitch_config_t* cfg, const ttsb_tensor_t* weights, int n_weights,
                          int device, ttsb_fastpitch_t** out);
void ttsb_fastpitch_destroy(ttsb_fastpitch_t* h);
/* d_state: persistent between encode -> condition -> decode of one batch (token lengths,
 * conditioned encoder output, cumulative durations). d_workspace: scratch, sized for
 * max(L, T) rows; encode/condition need T = 0, decode needs the real T. */
size_t ttsb_fastpitch_state_bytes(const ttsb_fastpitch_t* h, int B, int L);
size_t ttsb_fastpitch_workspace_bytes(const ttsb_fastpitch_t* h, int B, int L, int T);

/* d_ids [B,L] int64 (0 = padding, trailing only). Outputs: d_log_dur [B,L] fp32 (masked head
 * output, before exp), d_pitch [B,L] fp32. Speaker conditioning (model.py:358-362, multi-speaker checkpoints only):
 * d_speaker_ids [B] int64 per utterance, or NULL and one `speaker` for the whole batch; speaker < 0 with NULL ids = none.
 * Ids are validated ON THE DEVICE (no host sync here): out-of-range token / speaker ids and non-trailing padding are
 * reported through ttsb_fastpitch_condition's d_summary[1]; the gathers themselves are clamped. */
int ttsb_fastpitch_encode(ttsb_fastpitch_t* h, const int64_t* d_ids, int B, int L, int speaker,
                          const int64_t* d_speaker_ids, float* d_log_dur, float* d_pitch, void* d_state,
                          void* d_workspace, size_t workspace_bytes, void* stream);
/* Tap for parity tests: after encode (and before condition, which adds the pitch / energy embeddings in place) the state
 * holds the encoder output enc_out [B,L,d_model] as fp16 (model.py:364, `enc_out, enc_mask = self.encoder(...)`). */
int ttsb_fastpitch_read_enc_out(ttsb_fastpitch_t* h, int B, int L, const void* d_state, void* d_enc_out, void* stream);
/* d_pitch_in [B,L]: pitch track to embed (predicted, transformed or target). d_energy_tgt [B,L] or
 * NULL (predict). d_dur_tgt [B,L] or NULL (use exp(log_dur)-1). Outputs: d_dur_pred [B,L],
 * d_energy_pred [B,L] (untouched if no energy conditioning), d_dec_lens [B] int64, and d_summary (int32[2 + B], may be
 * NULL): [2..] = dec_lens as int32 (so that ONE device -> host read gives the caller every frame count), [0] = max(dec_lens) — the one value the caller has to read on the host before decode (the reference's own
 * sync, model.py:76) — and [1] = input status bits from encode: 1 token id outside [0, n_symbols), 2 padding that is
 * not trailing or an empty utterance, 4 speaker id outside [0, n_speakers). */
int ttsb_fastpitch_condition(ttsb_fastpitch_t* h, int B, int L, const float* d_log_dur,
                             const float* d_pitch_in, const float* d_energy_tgt, const float* d_dur_tgt,
                             float pace, float max_duration, float* d_dur_pred, float* d_energy_pred,
                             int64_t* d_dec_lens, int32_t* d_summary, void* d_state, void* d_workspace,
                             size_t workspace_bytes, void* stream);
/* T = max(dec_lens) read by the caller. d_mel [B,n_mel,T] fp32 (reference layout; frames beyond
 * an utterance hold proj.bias exactly like the reference). d_mel_cl optional [B,T,128] fp16
 * channel-last copy for ttsb_hifigan_forward (zero beyond each utterance), may be NULL. */
int ttsb_fastpitch_decode(ttsb_fastpitch_t* h, int B, int L, int T, float* d_mel, void* d_mel_cl,
                          void* d_state, void* d_workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Tacotron2 (multi-speaker), replaces models.tacotron2.tacotron2_ms.Tacotron2MS.infer
 * (models/tacotron2/tacotron2_ms.py:278-332; arithmetic of torchaudio 2.11 _Encoder/_Decoder/_Postnet).
 * Weights: the module's state_dict (BatchNorm running stats included; folded at create()).
 *   encode   tokens -> encoder memory, decoder state reset
 *   decode   n autoregressive steps; the prenet's always-on dropout (torchaudio:283-285) takes its
 *            keep-masks from the caller (d_masks [n_steps, 2, B, 256] bytes, 1 = keep)
 *   finish   postnet + residual, lengths, alignments
 * The caller polls `h_done_step` every few steps instead of the reference's per-step host sync.
 * --------------------------------------------------------------------------------------------- */
int ttsb_tacotron2_create(const ttsb_tensor_t* weights, int n_weights, int device, ttsb_tacotron2_t** out);
void ttsb_tacotron2_destroy(ttsb_tacotron2_t* h);
size_t ttsb_tacotron2_state_bytes(const ttsb_tacotron2_t* h, int B, int L, int max_steps);
size_t ttsb_tacotron2_workspace_bytes(const ttsb_tacotron2_t* h, int B, int L, int T);
int ttsb_tacotron2_encode(ttsb_tacotron2_t* h, const int64_t* d_tokens, const int32_t* d_lengths,
                          const int64_t* d_speaker_ids, int B, int L, int max_steps, void* d_state,
                          void* d_workspace, size_t workspace_bytes, void* stream);
/* Runs decoder steps [step0, step0 + n_steps) as ONE cooperative launch (persistent kernel, grid barriers between the
 * phases of a step). early_stop != 0: the chunk ends after the first step at which every utterance has raised its stop
 * gate (torchaudio:850, decided on the device). h_done_step (host int[2], optional): after the call — which then
 * synchronises the stream — [0] = that step or -1, [1] = input status bits raised by encode (1: a token id outside
 * [0, n_symbol), 4: a speaker id outside [0, num_speakers); the gathers themselves were clamped). */
int ttsb_tacotron2_decode(ttsb_tacotron2_t* h, int B, int L, int max_steps, int step0, int n_steps,
                          const uint8_t* d_masks, float gate_threshold, int early_stop, void* d_state,
                          int* h_done_step, void* stream);
int ttsb_tacotron2_finish(ttsb_tacotron2_t* h, int B, int L, int max_steps, int T, float* d_mel,
                          int32_t* d_mel_lengths, float* d_alignments, void* d_mel_cl, void* d_state,
                          void* d_workspace, size_t workspace_bytes, void* stream);

/* Wrapper post-processing of a whole batch in one launch (replaces the per-utterance Python loop of
 * models/tacotron2/networks.py:192-206 with its host sync per utterance): for utterance b with d_cols[b] >= 0 the mel is
 * cut where alignments[b, :, d_cols[b]] first reaches 80 % of its maximum and its last kept frame is repeated three times
 * (truncate_mel, :44-49); then, if rate != 1, it is resized in time to int(len / rate) frames with torch's bicubic kernel
 * (resize_mel, :52-67). d_mel [B,n_mel,T] fp32, d_align [B,T,L] fp32 (may be NULL when d_cols is NULL), d_out
 * [B,n_mel,T_out] fp32 (frames beyond d_out_lens[b] are zero), d_out_lens [B] int32 (clipped to T_out). */
int ttsb_tacotron2_postprocess(const float* d_mel, const int32_t* d_mel_lens, const float* d_align, const int32_t* d_cols,
                               double rate, int B, int n_mel, int T, int L, int T_out, float* d_out, int32_t* d_out_lens,
                               void* stream);

/* ---------------------------------------------------------------------------------------------
 * Single conv site, for parity tests of the dense-contraction kernel in isolation.
 *   kind 0: nn.Conv1d     weight [Cout,Cin,K], stride 1, padding = dilation*(K-1)/2
 *   kind 1: nn.ConvTranspose1d weight [Cin,Cout,K=2*stride], padding = stride/2
 * --------------------------------------------------------------------------------------------- */
int ttsb_conv1d_create(int kind, int cin, int cout, int ksize, int dilation, int stride,
                       const float* h_weight, const float* h_bias, int device, ttsb_conv1d_t** out);
void ttsb_conv1d_destroy(ttsb_conv1d_t* h);
/* d_in: [B,T,cin_pad] fp16 channel-last (cin_pad = cin rounded up to 32/64 as reported by
 * ttsb_conv1d_cin_pad). d_out: [B,T*stride,cout] fp16. Optional fused terms: d_residual (same shape
 * as d_out), act_slope (<0: none; else leaky-relu slope applied to the output), d_lens. */
int ttsb_conv1d_cin_pad(const ttsb_conv1d_t* h);
int ttsb_conv1d_forward(ttsb_conv1d_t* h, const void* d_in, int B, int T, const void* d_residual,
                        float act_slope, const int32_t* d_lens, void* d_out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * HiFi-GAN bias denoiser on a padded batch (replaces Denoiser.forward, vocoder/hifigan/denoiser.py:66-72, called
 * once per utterance by models/fastpitch/networks.py:343-344): STFT(1024, hop 256, periodic hann, center/reflect)
 * -> max(|X| - strength * bias_spec, 0) * exp(j arg X) -> ISTFT, each utterance at its OWN length
 * d_n_samples[b] (multiples of 256; samples beyond it are written as zeros).
 * d_wav / d_out: [B, n_max] fp32; d_bias_spec: [513] fp32 (Denoiser.bias_spec); workspace: frame buffer.
 * --------------------------------------------------------------------------------------------- */
size_t ttsb_denoiser_workspace_bytes(int B, int n_max);
int ttsb_denoiser_forward(const float* d_wav, const int32_t* d_n_samples, int B, int n_max, const float* d_bias_spec,
                          float strength, float* d_out, void* d_workspace, size_t workspace_bytes, void* stream);

/* Op-level entry for one fused ResBlock1 step (vocoder/hifigan/models.py:46-53, one (c1, c2) iteration):
 *   out = x + conv2(lrelu(conv1(lrelu(x)) + b1)) + b2,  conv1 = Conv1d(C, C, k, dilation=d), conv2 = Conv1d(C, C, k)
 * as ONE kernel launch (csrc/conv_pair.cu). w1/w2: [C, C, k] fp32 host, b1/b2: [C]. d_x/d_out: [B, T, C] fp16
 * channel-last, rows >= d_lens[b] (optional) are zero padding on both sides of the step.
 * ttsb_convpair_plan fills {ok, rows per tile, x slots, t slots, w2 resident, w2 ring stages, TMEM columns, smem bytes};
 * ok = 0 means the shape has no fused plan (the generator then runs the two convs as separate launches). */
typedef struct ttsb_convpair ttsb_convpair_t;
int ttsb_convpair_create(int channels, int ksize, int dilation, const float* h_w1, const float* h_b1,
                         const float* h_w2, const float* h_b2, int device, ttsb_convpair_t** out);
void ttsb_convpair_destroy(ttsb_convpair_t* h);
int ttsb_convpair_plan(const ttsb_convpair_t* h, int* out8);
int ttsb_convpair_forward(ttsb_convpair_t* h, const void* d_x, int B, int T, const int32_t* d_lens, float slope,
                          void* d_out, void* stream);
/* Same step on ACTIVATED tensors, the form the generator keeps in HBM: d_lx = lrelu(x, slope), d_lout = lrelu(out, slope)
 * (the residual add inverts the activation; csrc/hifigan.cu). */
int ttsb_convpair_forward_act(ttsb_convpair_t* h, const void* d_lx, int B, int T, const int32_t* d_lens, float slope,
                              void* d_lout, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TTSB200_H */

// HBM-bound / small kernels of the hot path (everything that is not a dense contraction).
// Each launcher returns 0 or sets the thread-local error and returns non-zero.
#pragma once
#include "common.cuh"

namespace ttsb {

// [B,80,T] fp32 (reference mel layout, hifigan/models.py:111) -> channel-last fp16 [B,T,ld]
// (ld >= C, channels [C,ld) zero), rows >= lens[b] zeroed (per-utterance zero padding).
// t_stride (0 = T): frames per utterance of `mel` in memory when only the first T are packed
int launch_pack_mel(const float* mel, const int* lens, int B, int C, int T, __half* out, int ld,
                    cudaStream_t s, int t_stride = 0);

// conv_post (Conv1d 32->1 k7 p3) + tanh on the already leaky-relu'd stage-4 activations
// (hifigan/models.py:123-125). x: [B,N,32] fp16, w: [7][32] fp32 (tap-major), wav: [B,N] fp32.
// n_stride (0 = N): samples per utterance of `wav` in memory when only the first N are written
int launch_conv_post_tanh(const __half* x, const float* w, float bias, const int* lens, int len_mul,
                          int B, int N, float* wav, cudaStream_t s, int n_stride = 0);

// word embedding gather + sinusoidal positional embedding * mask + conditioning
// (transformer.py:212-219, 34-48). ids int64 [B,L]; out fp16 [B,L,D].
// speaker conditioning (model.py:358-362): spk_table [n_speakers, D] pre-scaled, row chosen per utterance by spk_ids
// (int64 [B]) or by spk_scalar; ids / speaker ids are clamped (reported by launch_ids_to_lens).
int launch_embed(const int64_t* ids, const float* emb, const float* spk_table, const int64_t* spk_ids, int spk_scalar,
                 int n_speakers, const float* inv_freq, int B, int L, int D, int n_symbols, __half* out, cudaStream_t s);

// count of non-pad ids per row (mask = ids != padding_idx, transformer.py:214) + input validation: *status |= 1 (token
// id out of range), 2 (padding not trailing / empty utterance), 4 (speaker id out of range)
int launch_ids_to_lens(const int64_t* ids, int B, int L, int n_symbols, const int64_t* spk_ids, int n_speakers, int* lens,
                       int* status, cudaStream_t s);

// single-head attention with key-padding mask (transformer.py:131-146), d_head = 64.
// qkv: [B,S,192] fp16 (q|k|v), out: [B,S,64] fp16.
int launch_attention(const __half* qkv, const int* lens, int B, int S, float scale, __half* out,
                     cudaStream_t s, __half* vt_scratch = nullptr, int* err_flag = nullptr);
// the tcgen05 implementation (attention_tc.cu; the default): needs attention_tc_scratch_bytes(B, S) of scratch for V^T
size_t attention_tc_scratch_bytes(int B, int S);
int launch_attention_tc(const __half* qkv, const int* lens, int B, int S, float scale, __half* out, __half* vt_scratch,
                        int* err_flag, cudaStream_t s);

// x[b,l,:] = (x[b,l,:] + bias + sum_k w[:,k] * p[b,l+k-1]) * mask   (model.py:382-386,393-397:
// Conv1d(1->D,k3,p1) of the per-token pitch / energy track added onto enc_out)
int launch_scalar_embed_add(__half* x, const float* p, const float* w, const float* bias,
                            const int* lens, int B, int L, int D, int apply_mask, cudaStream_t s);

// durations -> repeats -> cumulative frame offsets (model.py:368, 68-79).
// log_dur (masked head output) or dur_tgt; writes dur_pred (fp32), cum [B,L+1] int32, dec_lens.
int launch_durations(const float* log_dur, const float* dur_tgt, float pace, float max_duration,
                     int B, int L, float* dur_pred, int* cum, int* dec_lens, int64_t* dec_lens64,
                     int* max_dec_len, cudaStream_t s);

// length regulator as a row gather (model.py:81-86) fused with the decoder's positional
// embedding (transformer.py:216-219): out[b,t,:] = enc[b,tok(t),:] + posemb(t) for t < dec_len, else 0.
int launch_regulate(const __half* enc, const int* cum, const int* dec_lens, const float* inv_freq,
                    int B, int L, int T, int D, __half* out, cudaStream_t s);

}  // namespace ttsb

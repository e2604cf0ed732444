"""Stated floating-point tolerances of the CUDA path against the fp32 oracle.

The CUDA path stores activations in fp16 and accumulates in fp32 (tcgen05 kind::f16 / fp32 FMA),
i.e. the precision class of the reference run with `.half()`. Calibration in BASELINE.md §2: the
reference's own fp16-vs-fp32 gap is mel L_inf 1.7e-2 and waveform relative RMS 3.2e-4.

  MEL_LINF        max |mel_cuda - mel_oracle| with identical durations (teacher-forced or const-4)
  WAV_REL_RMS     rms(wav_cuda - wav_oracle) / rms(wav_oracle), identical mel input
  WAV_LINF        max |wav_cuda - wav_oracle| (waveforms are tanh outputs in [-1, 1])
  SCALAR_ABS      per-token predictor outputs (log-duration, pitch, energy), O(1) values
  E2E_*           text ids -> waveform through both models (mel error feeds the vocoder)
"""
MEL_LINF = 2e-2
WAV_REL_RMS = 2e-3
WAV_LINF = 1e-2
SCALAR_ABS = 1e-2
E2E_WAV_REL_RMS = 2e-2
E2E_WAV_LINF = 5e-2

# Tacotron2 (fp16 weights, fp32 recurrent state, identical injected prenet masks; the decoder feeds its own
# output back for every step, so the bound is looser than for the feed-forward FastPitch):
T2_MEL_LINF = 5e-2
T2_ALIGN_ABS = 2e-2

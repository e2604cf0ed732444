"""Deterministic stand-in for Tacotron2MS.infer, shared by oracle/make_golden_r2.py (which drives the REAL reference
wrapper, models/tacotron2/networks.py:123-208, with it) and tests/test_t2_wrapper.py (which drives this package's wrapper
with it): test infrastructure only."""
import torch


def stub_infer_outputs(tokens, lengths, seed):
    """Deterministic stand-in for Tacotron2MS.infer used on BOTH sides of the wrapper fixture (shared through this
    module): mel [B,8,T], mel lengths, alignments [B,T,L] whose separator column peaks late in the utterance."""
    g = torch.Generator().manual_seed(seed)
    B, L = tokens.shape
    lengths = torch.as_tensor(lengths).cpu() if lengths is not None else torch.full((B,), L)
    mel_lens = (lengths * 5 + 7).to(torch.int32)
    T = int(mel_lens.max())
    mel = torch.randn(B, 8, T, generator=g)      # 8 rows: the wrapper is agnostic to the number of mel bins, the fixture stays small
    align = torch.rand(B, T, L, generator=g) * 0.05
    for b in range(B):
        n, t = int(lengths[b]), int(mel_lens[b])
        peak = max(1, int(0.8 * t))
        ramp = torch.linspace(0, 1, peak)
        align[b, :peak, max(n - 3, 0)] += ramp ** 4          # the separator column (n - n_eos - 1 with n_eos = 2)
        align[b, peak:t, max(n - 3, 0)] += 1.0
    return mel, mel_lens, align



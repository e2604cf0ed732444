"""Mints the golden fixtures under tests/golden/ by running the REAL reference code
(/root/reference, imported in-process; works only in the build container) on the seeded synthetic
checkpoints of tts_arabic_pytorch_b200.utils.synth. The fixtures pin the oracle (tests/test_oracle.py)
and, through it, the CUDA path.  Usage: python oracle/make_golden.py
"""
import json
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference'
sys.path.insert(0, REPO)
sys.path.insert(0, REF)
os.chdir(REF)   # configs/basic.yaml is opened relative to CWD (utils/__init__.py:31-32)

from tts_arabic_pytorch_b200.utils import synth  # noqa: E402

OUT = os.path.join(REPO, 'tests', 'golden')


def ref_hifigan():
    from vocoder.hifigan.env import AttrDict
    from vocoder.hifigan.models import Generator
    g = Generator(AttrDict(synth.HIFIGAN_CONFIG))
    g.load_state_dict(synth.hifigan_state_dict(1235))
    g.eval()
    g.remove_weight_norm()
    return g


def ref_fastpitch(dur_mode):
    from models.fastpitch.fastpitch.model import FastPitch
    m = FastPitch(**synth.FASTPITCH_CONFIG)
    m.load_state_dict(synth.fastpitch_state_dict(1234, dur_mode=dur_mode))
    m.eval()
    return m


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    gen = torch.Generator().manual_seed(0)

    # ---- HiFi-GAN: config-2 style log-mel input, short so the fixture stays small --------------
    g = ref_hifigan()
    mel = torch.clamp(torch.randn(2, 80, 24, generator=gen) * 2 - 5, -11.5129, 2.0)
    with torch.no_grad():
        wav_b = g(mel)                       # batched [2,1,6144]
        wav_u = g(mel[1, :, :17])            # unbatched, shorter utterance [1, 4352]
    np.savez_compressed(os.path.join(OUT, 'hifigan_small.npz'), mel=mel.numpy(), wav_batched=wav_b.numpy(),
                        wav_unbatched_len17=wav_u.numpy())

    # ---- FastPitch: padded batch, free-running const-4 durations ------------------------------
    m = ref_fastpitch('const4')
    ids = torch.randint(1, 40, (3, 20), generator=gen)
    ids[1, 13:] = 0
    ids[2, 7:] = 0
    with torch.no_grad():
        mel_o, dec_lens, dur, pitch, energy = m.infer(ids)
        enc_out, _ = m.encoder(ids)
    np.savez_compressed(os.path.join(OUT, 'fastpitch_const4.npz'), ids=ids.numpy(), mel=mel_o.numpy(),
                        dec_lens=dec_lens.numpy(), dur_pred=dur.numpy(), pitch_pred=pitch.numpy(),
                        energy_pred=energy.numpy(), enc_out=enc_out.numpy())

    # ---- FastPitch: random duration head, pace != 1, pitch transform, plus teacher forcing -----
    m = ref_fastpitch('random')
    with torch.no_grad():
        trf = lambda p, n, mean, std: 1.1 * p + 0.2     # noqa: E731  (networks.py:38-42 shape)
        mel_o, dec_lens, dur, pitch, energy = m.infer(ids, pace=0.9, pitch_transform=trf)
        dur_tgt = torch.randint(0, 6, (3, 20), generator=gen).float() * (ids != 0)
        mel_t, dec_lens_t, *_ = m.infer(ids, dur_tgt=dur_tgt)
    np.savez_compressed(os.path.join(OUT, 'fastpitch_random.npz'), ids=ids.numpy(), mel=mel_o.numpy(),
                        dec_lens=dec_lens.numpy(), dur_pred=dur.numpy(), pitch_pred=pitch.numpy(),
                        energy_pred=energy.numpy(), dur_tgt=dur_tgt.numpy(), mel_tf=mel_t.numpy(),
                        dec_lens_tf=dec_lens_t.numpy())

    # ---- config 1: line 0 of data/infer_text.txt through the reference front-end ---------------
    import text
    line = open(os.path.join(REF, 'data', 'infer_text.txt'), encoding='utf-8').readline().strip()
    tokens = text.buckwalter_to_tokens(line, append_space=False)
    tok_ids = text.tokens_to_ids(tokens)
    m = ref_fastpitch('const4')
    with torch.no_grad():
        mel1, dl1, *_ = m.infer(torch.LongTensor(tok_ids)[None])
    np.savez_compressed(os.path.join(OUT, 'config1_line0.npz'), ids=np.array(tok_ids), mel=mel1.numpy(),
                        dec_lens=dl1.numpy())
    with open(os.path.join(OUT, 'config1_line0.json'), 'w', encoding='utf-8') as f:
        json.dump({'line': line, 'tokens': tokens, 'ids': tok_ids}, f, ensure_ascii=False)

    # ---- end-to-end: FastPitch2Wave-equivalent on a tiny batch --------------------------------
    ids2 = torch.randint(1, 40, (2, 10), generator=gen)
    ids2[1, 6:] = 0
    with torch.no_grad():
        mel2, dl2, *_ = m.infer(ids2)
        wavs = [g(mel2[b, :, :dl2[b]])[0].numpy() for b in range(2)]   # networks.py:340-345
    np.savez_compressed(os.path.join(OUT, 'e2e_small.npz'), ids=ids2.numpy(), mel=mel2.numpy(), dec_lens=dl2.numpy(),
                        wav0=wavs[0], wav1=wavs[1])
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == '__main__':
    main()

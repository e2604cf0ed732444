"""Runs the CUDA path once under the kernel-variant switches of the calling environment (TTSB_PAIR, TTSB_TMA_OUT,
TTSB_ATTENTION, TTSB_CLUSTER ... are read once per process) and stores mel / waveform for the variant-agreement
tests.   python tests/helpers/run_variant.py out.npz"""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)


def main():
    from tts_arabic_pytorch_b200.models.fastpitch.fastpitch.model import FastPitch
    from tts_arabic_pytorch_b200.utils import synth
    from tts_arabic_pytorch_b200.vocoder.hifigan.env import AttrDict
    from tts_arabic_pytorch_b200.vocoder.hifigan.models import Generator
    dev = torch.device('cuda:0')
    fp = FastPitch(**synth.FASTPITCH_CONFIG)
    fp.load_state_dict(synth.fastpitch_state_dict(1234))
    fp = fp.eval().to(dev)
    voc = Generator(AttrDict(synth.HIFIGAN_CONFIG))
    voc.load_state_dict(synth.hifigan_state_dict(1235))
    voc.remove_weight_norm()
    voc = voc.eval().to(dev)
    gen = torch.Generator().manual_seed(11)
    lens = [37, 29, 29, 5, 1]                      # ragged, duplicates, a single-token utterance
    ids = torch.zeros(len(lens), max(lens), dtype=torch.long)
    for b, n in enumerate(lens):
        ids[b, :n] = torch.randint(1, 40, (n,), generator=gen)
    mel, dec_lens, _, _, _, mel_cl = fp.infer(ids, return_channel_last=True)
    wav = voc.run(mel_cl=mel_cl, lens=dec_lens)
    torch.cuda.synchronize()
    np.savez(sys.argv[1], mel=mel.cpu().numpy(), wav=wav.cpu().numpy(), dec_lens=dec_lens.cpu().numpy())


if __name__ == '__main__':
    main()

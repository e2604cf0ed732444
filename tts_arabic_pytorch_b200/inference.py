"""Command line synthesis: list file in, wav files + index.html out — the flags of the reference's inference.py
(inference.py:21-107: --list --model --checkpoint --vocoder_sd --vocoder_config --out_dir --speed --denoise --batch_size
--cpu), over this package's FastPitch2Wave / Tacotron2Wave.

    python -m tts_arabic_pytorch_b200.inference --list data/infer_text.txt --checkpoint fp.pth \\
        --vocoder_sd hifigan.pth --vocoder_config config.json --out_dir samples/results --batch_size 32

Differences: there is no CPU path (`--cpu` is rejected with the reason), wav files are written by a 20-line RIFF writer
instead of torchaudio.save (32-bit float PCM at 22 050 Hz, what torchaudio writes for a float tensor), and --gpus N > 1
shards every batch over N processes with parallel.synthesize when launched under torchrun.
"""
import argparse
import html as _html
import os
import struct
import sys

import torch


def write_wav_f32(path, wav, sample_rate=22050):
    """1-D float32 CPU tensor -> IEEE-float WAV (RIFF, format tag 3, mono)."""
    data = wav.detach().to(torch.float32).contiguous().cpu().numpy().astype('<f4').tobytes()
    with open(path, 'wb') as f:
        f.write(b'RIFF' + struct.pack('<I', 4 + 26 + 12 + 8 + len(data)) + b'WAVE')
        f.write(b'fmt ' + struct.pack('<IHHIIHHH', 18, 3, 1, sample_rate, sample_rate * 4, 4, 32, 0))
        f.write(b'fact' + struct.pack('<II', 4, len(data) // 4))
        f.write(b'data' + struct.pack('<I', len(data)))
        f.write(data)


def read_wav_f32(path):
    """Inverse of write_wav_f32 (tests)."""
    import numpy as np
    raw = open(path, 'rb').read()
    assert raw[:4] == b'RIFF' and raw[8:12] == b'WAVE'
    pos, rate, data = 12, None, None
    while pos + 8 <= len(raw):
        tag, n = raw[pos:pos + 4], struct.unpack('<I', raw[pos + 4:pos + 8])[0]
        body = raw[pos + 8:pos + 8 + n]
        if tag == b'fmt ':
            fmt, ch, rate = struct.unpack('<HHI', body[:8])
            assert fmt == 3 and ch == 1
        elif tag == b'data':
            data = np.frombuffer(body, dtype='<f4').copy()
        pos += 8 + n + (n & 1)
    return torch.from_numpy(data), rate


def _index_entry(idx, wav_rel, text_arabic, phonemes):
    return ('<div class="sample"><div class="text-arabic" dir="rtl">%s</div><div>%s</div>'
            '<audio controls preload="none" src="%s"></audio></div>\n' % (
                _html.escape(text_arabic), _html.escape('%d) %s' % (idx, phonemes)), _html.escape(wav_rel)))


def infer(args):
    from . import text
    from .utils import read_lines_from_file
    if args.cpu or not torch.cuda.is_available():
        raise RuntimeError('this build synthesizes on a CUDA device only (sm_100a kernels behind a C ABI; no CPU path)')
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    device = torch.device('cuda', int(os.environ.get('LOCAL_RANK', '0')))
    torch.cuda.set_device(device)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=device)
    if args.model == 'fastpitch':
        from .models.fastpitch import FastPitch2Wave
        model = FastPitch2Wave(args.checkpoint, vocoder_sd=args.vocoder_sd, vocoder_config=args.vocoder_config)
    elif args.model == 'tacotron2':
        from .models.tacotron2 import Tacotron2Wave
        model = Tacotron2Wave(args.checkpoint, vocoder_sd=args.vocoder_sd, vocoder_config=args.vocoder_config)
    else:
        raise ValueError('model type not supported: %r' % args.model)
    model = model.to(device).eval()
    lines = read_lines_from_file(args.list)
    batches = [lines[k:k + args.batch_size] for k in range(0, len(lines), args.batch_size)]
    if rank == 0:
        os.makedirs(os.path.join(args.out_dir, 'wavs'), exist_ok=True)
    idx = 0
    out_html = None
    if rank == 0:
        out_html = open(os.path.join(args.out_dir, 'index.html'), 'w', encoding='utf-8')
        out_html.write('<!DOCTYPE html>\n<html lang="en"><head><meta charset="UTF-8"><title>%s</title></head><body>\n'
                       % _html.escape(os.path.basename(args.out_dir)))
    try:
        for batch in batches:
            if world > 1 and args.model == 'fastpitch':
                from . import parallel
                ids = [model.model._ids(line, None) for line in batch]
                wav_list = parallel.synthesize(model, ids, speed=args.speed, denoise=args.denoise, deliver='nccl_host')
            else:
                wav_list = model.tts(batch, batch_size=args.batch_size, denoise=args.denoise, speed=args.speed)
            if rank != 0:
                continue
            for text_line, wav in zip(batch, wav_list):
                write_wav_f32(os.path.join(args.out_dir, 'wavs', 'static%d.wav' % idx), wav)
                buckw = text.arabic_to_buckwalter(text_line)
                phon = text.simplify_phonemes(text.buckwalter_to_phonemes(buckw).replace(' ', '').replace('+', ' '))
                out_html.write(_index_entry(idx, 'wavs/static%d.wav' % idx, text.buckwalter_to_arabic(buckw), phon))
                idx += 1
    finally:
        if out_html is not None:
            out_html.write('</body></html>\n')
            out_html.close()
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()
    if rank == 0:
        print('Saved files to: %s' % args.out_dir)
    return idx


def main(argv=None):
    p = argparse.ArgumentParser(description=__doc__.split('\n')[0])
    p.add_argument('--list', type=str, default='./data/infer_text.txt')
    p.add_argument('--model', type=str, default='fastpitch')
    p.add_argument('--checkpoint', type=str, default='pretrained/fastpitch_ar_adv.pth')
    p.add_argument('--vocoder_sd', type=str, default=None)
    p.add_argument('--vocoder_config', type=str, default=None)
    p.add_argument('--out_dir', type=str, default='samples/results')
    p.add_argument('--speed', type=float, default=1.0)
    p.add_argument('--denoise', type=float, default=0)
    p.add_argument('--batch_size', type=int, default=2)
    p.add_argument('--cpu', action='store_true')
    return infer(p.parse_args(argv))


if __name__ == '__main__':
    main()
    sys.exit(0)

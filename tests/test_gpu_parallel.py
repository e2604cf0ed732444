"""BASELINE config 5 semantics on one GPU: a batch sharded the way `parallel.synthesize` shards it over N ranks (snake
deal + the global padded-position condition in the token AND the frame domain) must give, utterance by utterance, exactly
the result of the reference's single padded batch (models/fastpitch/networks.py:140-195) — FastPitch is not
batch-invariant (transformer.py:83-85, model.py:129-133), so naive per-shard batches do NOT (the last test shows it)."""
import numpy as np
import pytest
import torch

from tts_arabic_pytorch_b200 import parallel
from tts_arabic_pytorch_b200.utils import synth

pytestmark = pytest.mark.gpu


def _dev():
    if not torch.cuda.is_available():
        pytest.fail('GPU tests need a CUDA device (they never fall back to the CPU)')
    return torch.device('cuda:0')


@pytest.fixture(scope='module')
def tts(tmp_path_factory):
    _dev()
    from tts_arabic_pytorch_b200.models.fastpitch import FastPitch2Wave
    d = tmp_path_factory.mktemp('ckpt_par')
    fp, hg, cj = synth.write_checkpoints(str(d), seed=1234)
    return FastPitch2Wave(fp, vocoder_sd=hg, vocoder_config=cj, arabic_in=False).cuda()


def _utterances(n, lo, hi, seed):
    g = torch.Generator().manual_seed(seed)
    lens = torch.randint(lo, hi + 1, (n,), generator=g).tolist()
    return [torch.randint(1, 40, (k,), generator=g) for k in lens]


def _single_batch(tts, ids):
    (mel, dec_lens, *_), inverse = tts.model._infer_ids(ids, 1.0, 0, None, None, None, None, 75)
    lens = dec_lens.tolist()
    return [mel[r, :, :lens[r]].clone() for r in inverse.tolist()]


def _sharded(tts, ids, world, keep_global_padding=True):
    """What parallel.synthesize does, with the ranks run one after another on this GPU (two passes: the frame-domain
    all-reduce(max) needs every shard's frame count first)."""
    lengths = [int(x.numel()) for x in ids]
    shards, pad_to = parallel.plan_shards(lengths, world)
    t_sub = []
    for r in range(world):
        (mel, dec_lens, *_), _ = tts.model._infer_ids([ids[i] for i in shards[r]], 1.0, 0, None, None, None, None, 75,
                                                      pad_to=pad_to[r] if keep_global_padding else 0)
        t_sub.append(int(dec_lens.max()))
    t_global = max(t_sub)
    per_rank = []
    for r in range(world):
        hook = (lambda t: t + 1 if t < t_global else t) if keep_global_padding else None
        (mel, dec_lens, *_), inverse = tts.model._infer_ids([ids[i] for i in shards[r]], 1.0, 0, None, None, None, None,
                                                            75, pad_to=pad_to[r] if keep_global_padding else 0,
                                                            frame_len_hook=hook)
        lens = dec_lens.tolist()
        per_rank.append([mel[row, :, :lens[row]].clone() for row in inverse.tolist()])
    return parallel.unshard(per_rank, shards)


@pytest.mark.parametrize('world', [2, 8])
def test_sharded_batch_equals_single_padded_batch_bit_for_bit(tts, world):
    ids = _utterances(48, 64, 256, seed=0)
    ref = _single_batch(tts, ids)
    got = _sharded(tts, ids, world)
    for i, (a, b) in enumerate(zip(got, ref)):
        assert a.shape == b.shape, i
        assert torch.equal(a, b), (i, float((a - b).abs().max()))


def test_naive_shards_differ_which_is_why_the_padding_condition_travels(tts):
    ids = _utterances(24, 32, 96, seed=1)
    ref = _single_batch(tts, ids)
    naive = _sharded(tts, ids, 4, keep_global_padding=False)
    worst = max(float((a - b).abs().max()) for a, b in zip(naive, ref) if a.shape == b.shape)
    assert worst > 1e-3          # the longest utterance of each shard lost its padded neighbour


def test_synthesize_single_process_returns_waveforms_in_input_order(tts):
    """world = 1 (no process group): the product entry point degenerates to synthesize_ids."""
    ids = _utterances(5, 8, 24, seed=2)
    res = parallel.synthesize(tts, ids, deliver='nccl_host')
    ref, _ = tts.synthesize_ids(ids)
    assert len(res) == 5
    for a, b in zip(res, ref):
        assert a.device.type == 'cpu' and a.shape == b.shape
        assert float(np.abs(a.numpy() - b.numpy()).max()) == 0.0

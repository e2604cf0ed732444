"""The text front-end is host-side Python but must yield identical id tensors:
token-for-token comparison with fixtures produced by the reference front-end on its own corpus
(data/infer_text.txt, train/test transcripts), on Arabic-script input, and on fuzzed strings
(oracle/make_text_golden.py)."""
import gzip
import json
import os

import pytest
import torch

from tts_arabic_pytorch_b200 import text


@pytest.fixture(scope='module')
def rec(golden_dir):
    with gzip.open(os.path.join(golden_dir, 'text_frontend.json.gz'), 'rt', encoding='utf-8') as f:
        return json.load(f)


def test_symbol_table(rec):
    assert text.symbols == rec['symbols'] and len(text.symbols) == 40
    assert text.phon_to_id_['_pad_'] == 0


def test_reference_corpus_buckwalter(rec):
    assert len(rec['buckwalter']) > 2000
    for line, phonemes, tokens in rec['buckwalter']:
        assert text.buckwalter_to_phonemes(line) == phonemes, line
        assert text.buckwalter_to_tokens(line, append_space=False) == tokens, line
    # every corpus line maps onto the 40-symbol table (ids feed the embedding)
    ids = text.tokens_to_ids(rec['buckwalter'][0][2])
    assert min(ids) >= 1 and max(ids) < 40


def test_arabic_script_input(rec):
    for line, buckw, tokens in rec['arabic']:
        assert text.arabic_to_buckwalter(line) == buckw
        assert text.arabic_to_tokens(line) == tokens
    assert text.buckwalter_to_arabic(rec['buckwalter'][0][0]) == rec['round_trip']
    assert text.arabic_to_buckwalter(text.buckwalter_to_arabic('>als~alAmu Ealaykum')) == '>als~alAmu Ealaykum'


def test_fuzzed_strings(rec):
    n = 0
    for line, phonemes, tokens in rec['fuzz']:
        if phonemes is None:
            with pytest.raises(Exception):
                text.buckwalter_to_tokens(line)
            continue
        assert text.buckwalter_to_phonemes(line) == phonemes, repr(line)
        assert text.buckwalter_to_tokens(line) == tokens, repr(line)
        n += 1
    assert n > 3000


def test_unknown_phoneme_raises_keyerror():
    # punctuation survives phonetisation but is not in the symbol table: KeyError, like the reference
    with pytest.raises(KeyError):
        text.tokens_to_ids(text.buckwalter_to_tokens('qAl.', append_space=False))


def test_collate_sorts_pads_and_inverts():
    from tts_arabic_pytorch_b200.models.fastpitch.networks import text_collate_fn
    batch = [torch.LongTensor([5, 6, 7]), torch.LongTensor([9]), torch.LongTensor([1, 2, 3, 4, 5]), torch.LongTensor([8, 8])]
    padded, lens, inverse = text_collate_fn(batch)
    assert lens.tolist() == [5, 3, 2, 1]
    assert padded.shape == (4, 5) and padded[3].tolist() == [9, 0, 0, 0, 0]
    for i, src in enumerate(batch):
        row = padded[inverse[i]]
        assert row[:src.numel()].tolist() == src.tolist() and int(row[src.numel():].sum()) == 0


def test_word_cache_is_transparent():
    """word phonetisation is memoised: results must not depend on cache state and a caller
    editing a returned list must not poison later calls."""
    from tts_arabic_pytorch_b200 import text
    from tts_arabic_pytorch_b200.text import phonetiser
    line = ">als~alAmu Ealaykum yA Sadiyqiy , >als~alAmu Ealaykum"
    phonetiser._word_to_phones_cached.cache_clear()
    cold = text.buckwalter_to_tokens(line)
    warm = text.buckwalter_to_tokens(line)
    assert cold == warm
    ph = phonetiser.word_to_phones(">als~alAmu")
    assert isinstance(ph, list)
    ph.append('XX')
    assert phonetiser.word_to_phones(">als~alAmu") == ph[:-1]
    assert text.buckwalter_to_tokens(line) == cold
    assert phonetiser._word_to_phones_cached.cache_info().hits > 0


def test_vowelizer_plugin_is_applied_where_the_reference_applies_it():
    """networks.py:75-85: the vowelizer sees the Arabic-script sentence before tokenisation; registered objects are
    memoised per sentence; unregistered names fail loudly."""
    import pytest
    from tts_arabic_pytorch_b200.models.fastpitch import networks as nw

    calls = []

    class Upper:
        def predict(self, s):
            calls.append(s)
            return s

    nw.register_vowelizer('unit_test_vowelizer', Upper())
    v = nw._load_vowelizer('unit_test_vowelizer', {})
    assert v.predict('abc') == 'abc' and v.predict('abc') == 'abc'
    assert calls == ['abc']
    nw.register_vowelizer('unit_test_factory', lambda cfg: Upper())
    assert nw._load_vowelizer('unit_test_factory', {}).predict('x') == 'x'
    with pytest.raises(NotImplementedError):
        nw._load_vowelizer('shakkala', {})

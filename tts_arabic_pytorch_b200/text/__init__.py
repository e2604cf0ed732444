"""Text front-end: Arabic / Buckwalter text -> phoneme tokens -> ids
(drop-in for the reference `text` package: text/__init__.py:24-72)."""
from .phonetiser import arabic_to_buckwalter, buckwalter_to_arabic  # noqa: F401
from .phonetiser import utterance_to_phoneme_string as process_utterance
from .symbols import DOUBLING_TOKEN, EOS_TOKEN, SEPARATOR_TOKEN, symbols

_STRESSED = {'aa': ('aa', 'AA'), 'uu': ('uu0', 'uu1', 'UU0', 'UU1'), 'ii': ('ii0', 'ii1', 'II0', 'II1'),
             'a': ('a', 'A'), 'u': ('u0', 'u1', 'U0', 'U1'), 'i': ('i0', 'i1', 'I0', 'I1')}
vowel_map = {variant: plain for plain, variants in _STRESSED.items() for variant in variants}
vowels = list(vowel_map)
phon_to_id_ = {phon: i for i, phon in enumerate(symbols)}
_TOKEN_MEMO = {}


def tokens_to_ids(phonemes, phon_to_id=None):
    table = phon_to_id_ if phon_to_id is None else phon_to_id
    return [table[p] for p in phonemes]     # KeyError on unknown phonemes, like the reference


def ids_to_tokens(ids):
    return [symbols[i] for i in ids]


def arabic_to_phonemes(arabic):
    return process_utterance(arabic_to_buckwalter(arabic))


def buckwalter_to_phonemes(buckw):
    return process_utterance(buckw)


def phonemes_to_tokens(phonemes: str, append_space=True):
    """phoneme string -> model tokens: word separators, geminates as consonant + doubling token,
    stress/emphasis variants of vowels merged (text/__init__.py:43-60)."""
    toks = phonemes.replace('sil', '').replace('+', SEPARATOR_TOKEN).split()
    out = []
    for t in toks:
        m = _TOKEN_MEMO.get(t)
        if m is None:       # the per-phoneme rule, evaluated once per distinct phoneme string
            if len(t) == 2 and t not in vowel_map and t[0] == t[1]:
                m = (t[0], DOUBLING_TOKEN)
            else:
                m = (vowel_map.get(t, t),)
            _TOKEN_MEMO[t] = m
        out.extend(m)
    if append_space:
        out.append(SEPARATOR_TOKEN)
    out.append(EOS_TOKEN)
    return out


def buckwalter_to_tokens(buckw, append_space=True):
    return phonemes_to_tokens(buckwalter_to_phonemes(buckw), append_space=append_space)


def arabic_to_tokens(arabic, append_space=True):
    return buckwalter_to_tokens(arabic_to_buckwalter(arabic), append_space=append_space)


def simplify_phonemes(phonemes):
    for k, v in vowel_map.items():
        phonemes = phonemes.replace(k, v)
    return phonemes

"""Rule-based Buckwalter -> phoneme conversion for vocalised Modern Standard Arabic.

Behavioural equivalent of the reference front-end's `process_utterance`
(text/phonetise_buckwalter.py:164-400, itself adapted from Halabi's Arabic-Phonetiser) for the one
output the TTS models consume: the FIRST pronunciation of every word. The reference enumerates all
pronunciation variants of a word and then keeps variant 0; this implementation walks the word once
and emits that variant directly. `tests/test_text.py` checks it token-for-token against fixtures
produced by the reference on its own corpus plus fuzzed strings.
"""
import functools
import re

# ---- script conversion -----------------------------------------------------------------------
_BUCKWALTER = "'|>&<}AbptvjHxd*rzs$SDTZEg_fqklmnhwYyFNKaui~o"   # U+0621 .. U+0652 in code-point order
_AR2BW = {}
_cp = 0x0621
for _ch in _BUCKWALTER:
    if _cp == 0x063B:       # U+063B..U+063F are not Arabic letters used here
        _cp = 0x0640
    _AR2BW[chr(_cp)] = _ch
    _cp += 1
_AR2BW[chr(0x062B)] = '^'   # thaa' is '^' in this dialect of Buckwalter (not 'v')
del _AR2BW[chr(0x0640)]     # tatweel is stripped later, not transliterated
_BW2AR = {v: k for k, v in _AR2BW.items()}


_AR2BW_TABLE = {ord(k): v for k, v in _AR2BW.items()}
_BW2AR_TABLE = {ord(k): v for k, v in _BW2AR.items()}


def arabic_to_buckwalter(s):
    return s.translate(_AR2BW_TABLE)


def buckwalter_to_arabic(s):
    return s.translate(_BW2AR_TABLE)


# ---- letter classes ----------------------------------------------------------------------------
HAMZAS = "><}&'"
PLAIN = set("b*Tmtr" "Zn^zEh" "jsgHqf" "xS$dDk")          # consonants that are their own phoneme
SIMPLE_CONS = PLAIN | set(HAMZAS)                        # every hamza form is the phoneme '<'
SHORT_MARKS = set("oauiFNK")                              # diacritics except shadda
MARKS = SHORT_MARKS | {'~'}
VOWEL_LETTERS = set("AYwyaui")
EMPHATIC = set("DSTZgxq")
FORWARD_ONLY = set("gx")                                  # do not spread emphasis backwards
CONS = SIMPLE_CONS | set("lmn") | {'|'}                  # "consonant" for context tests
CONS = CONS | set("lmnh")
PUNCT = ('.', ',', '?', '!')
LONG = {'w': ('uu0', 'UU0'), 'y': ('ii0', 'II0')}
SHORT = {'u': (('u0', 'u1'), ('U0', 'U1')), 'i': (('i0', 'i1'), ('I0', 'I1'))}

# irregular words, keyed by their skeleton over the letters h*Ahn'>wl}kmyTtfd
_IRREGULAR = {
    "h*A": ["h aa * aa", "h aa * a"],
    "h*h": ["h aa * i0 h i0", "h aa * i1 h"],
    "h*An": ["h aa * aa n i0", "h aa * aa n"],
    "h&lA'": ["h aa < u0 l aa < i0", "h aa < u0 l aa <"],
    "*lk": ["* aa l i0 k a", "* aa l i0 k"],
    "k*lk": ["k a * aa l i0 k a", "k a * aa l i1 k"],
    "*lkm": "* aa l i0 k u1 m",
    ">wl}k": ["< u0 l aa < i0 k a", "< u0 l aa < i1 k"],
    "Th": "T aa h a",
    "lkn": ["l aa k i0 nn a", "l aa k i1 n"],
    "lknh": "l aa k i0 nn a h u0",
    "lknhm": "l aa k i0 nn a h u1 m",
    "lknk": ["l aa k i0 nn a k a", "l aa k i0 nn a k i0"],
    "lknkm": "l aa k i0 nn a k u1 m",
    "lknkmA": "l aa k i0 nn a k u0 m aa",
    "lknnA": "l aa k i0 nn a n aa",
    "AlrHmn": ["rr a H m aa n i0", "rr a H m aa n"],
    "Allh": ["ll aa h i0", "ll aa h", "ll AA h u0", "ll AA h a", "ll AA h", "ll A"],
    "h*yn": ["h aa * a y n i0", "h aa * a y n"],
    "nt": "n i1 t",
    "fydyw": "v i0 d y uu1",
    "lndn": "l A n d u1 n",
}
_SKELETON = re.compile("[^h*Ahn'>wl}kmyTtfd]")


def _irregular(word):
    """First listed pronunciation whose final phoneme is compatible with the word's last letter."""
    entry = _IRREGULAR.get(_SKELETON.sub('', word))
    if entry is None:
        return None
    if isinstance(entry, str):
        return entry.split(' ')
    last = word[-1:] if word else ''
    if last == 'a':
        ends = ('a', 'A')
    elif last == 'A':
        ends = ('aa',)
    elif last == 'u':
        ends = ('u0',)
    elif last == 'i':
        ends = ('i0',)
    elif last in SIMPLE_CONS:
        ends = ('<' if last in HAMZAS else last,)
    else:
        ends = tuple(last)          # substring test on the raw letter, as the reference does
        for cand in entry:
            if cand.split(' ')[-1] in last:
                return cand.split(' ')
        return None
    for cand in entry:
        if cand.split(' ')[-1] in ends:
            return cand.split(' ')
    return None


def normalise(utterance):
    """Orthographic normalisation + tokenisation into words (reference preprocess_utterance)."""
    u = utterance
    for a, b in (('AF', 'F'), ('ـ', ''), ('o', ''), ('aA', 'A'), ('aY', 'Y'), (' A', ' '), ('F', 'an'), ('N', 'un'),
                 ('K', 'in'), ('|', '>A'), ('i~', '~i'), ('a~', '~a'), ('u~', '~u'), ('Ai', '<i'), ('Aa', '>a'),
                 ('Au', '>u')):
        u = u.replace(a, b)
    u = re.sub(r'^>([^auAw])', r'>a\1', u)
    u = re.sub(r' >([^auAw ])', r' >a\1', u)
    u = re.sub(r'<([^i])', r'<i\1', u)
    u = re.sub(r'(\S)(\.|\?|,|!)', r'\1 \2', u)
    return u.split(' ')


class _Out(list):
    """Phone slots of one word. A slot is [text, fixed]; `fixed` slots came from a multi-variant
    site in the reference (a list), which a following shadda does not double in variant 0."""

    def emit(self, text, fixed=False):
        self.append([text, fixed])

    def geminate(self):
        if self and not self[-1][1]:
            self[-1][0] += self[-1][0]


def _walk(word):
    w = 'bb' + word + 'ee'
    out = _Out()
    emph = False
    for i in range(2, len(w) - 2):
        c, n1, n2, p1, p2 = w[i], w[i + 1], w[i + 2], w[i - 1], w[i - 2]
        # emphasis spreading
        # every non-emphatic consonant (ra' included: the reference's exception list is inert) resets it
        if (c in CONS or c in 'wy') and c not in EMPHATIC:
            emph = False
        if c in EMPHATIC:
            emph = True
        if n1 in EMPHATIC and n1 not in FORWARD_ONLY:
            emph = True
        E = 1 if emph else 0

        if c in SIMPLE_CONS:
            out.emit('<' if c in HAMZAS else c)
        if c == 'l':
            # assimilated lam of the article: next letter is a bare consonant carrying a shadda
            sun = n1 not in MARKS and n1 not in VOWEL_LETTERS and n2 == '~'
            out.emit('' if sun else 'l')
        if c == '~' and p1 not in 'wy':
            out.geminate()
        if c == '|':
            out.emit('<', fixed=True)
        if c == 'p':
            out.emit('t' if n1 in MARKS else '')
        if c in 'wy':
            glide = c
            before_vowel = n1 in SHORT_MARKS or n1 in 'AY'
            before_glide = n1 in 'wy' and not (n2 in MARKS or n2 in 'Awy')
            closes_syllable = p1 in SHORT_MARKS and (n1 in CONS or n1 == 'e')
            if before_vowel or before_glide or closes_syllable:
                homorganic = 'u' if c == 'w' else 'i'
                blockers = 'aiAY' if c == 'w' else 'auAY'
                if p1 == homorganic and n1 not in blockers:
                    out.emit(LONG[c][E])
                elif c == 'w' and n1 == 'A' and n2 == 'e':
                    out.emit(glide, fixed=True)
                else:
                    out.emit(glide)
            elif n1 == '~':
                if p1 == 'a' or (c == 'w' and p1 in 'iy') or (c == 'y' and p1 in 'wu'):
                    out.emit(glide)
                    out.emit(glide)
                else:
                    out.emit(LONG[c][0])
                    out.emit(glide)
            else:
                word_final_after_cons = (p1 in CONS or p1 in 'ui') and n1 == 'e'
                out.emit(LONG[c][E], fixed=word_final_after_cons)
        if c in 'ui':
            weak = (n1 in SIMPLE_CONS or n1 == 'l') and n2 == 'e' and len(w) > 7
            out.emit(SHORT[c][E][1 if weak else 0])
        if c in 'aAY':
            long_v = ('aa', 'AA')[E]
            if c == 'A' and p1 in 'wk' and p2 == 'b':
                out.emit('a', fixed=True)
            elif c == 'A' and p1 in 'ui':
                pass
            elif c == 'A' and p1 == 'w' and n1 == 'e':
                out.emit('aa', fixed=True)
            elif c in 'AY' and n1 == 'e':
                out.emit(long_v, fixed=True)
            elif c == 'a':
                out.emit(('a', 'A')[E])
            else:
                out.emit(long_v)
    return [t for t, _ in out if t != '']


_LONGS = ('aa', 'uu0', 'ii0', 'AA', 'UU0', 'II0')


def _tidy(phones):
    """Merges a short vowel into a following identical long vowel, collapses repeated u0/i0 and
    fuses doubled glides (same passes, same order as the reference's house-keeping loop)."""
    drop = []
    prev = ''
    for i in range(len(phones)):
        cur = phones[i]
        if cur in _LONGS and prev.lower() == cur[1:].lower():
            drop.append(i - 1)
            phones[i] = phones[i - 1][0] + phones[i - 1]
        if cur in ('u0', 'i0') and prev.lower() == cur.lower():
            drop.append(i - 1)
            phones[i] = phones[i - 1]
        if cur in ('y', 'w') and prev == cur:
            phones[i - 1] += phones[i - 1]
            drop.append(i)
        prev = cur
    for j in reversed(drop):
        del phones[j]
    return phones


@functools.lru_cache(maxsize=1 << 16)
def _word_to_phones_cached(word):
    """The phonetisation of a word depends on the word alone (the reference's process_word takes nothing else), and
    running text repeats its words: once the GPU path synthesises ~10^4 x real time the rule walk is what a text batch
    waits for. Tuples, so cached results cannot be edited by a caller."""
    if word in PUNCT:
        return word
    fixed = _irregular(word)
    return tuple(_tidy(list(fixed) if fixed is not None else _walk(word)))


def word_to_phones(word):
    ph = _word_to_phones_cached(word)
    return ph if isinstance(ph, str) else list(ph)


def utterance_to_phoneme_string(utterance):
    groups = []
    for word in normalise(utterance):
        if word in ('-', 'sil'):
            groups.append(['sil'])
            continue
        ph = _word_to_phones_cached(word)
        if isinstance(ph, str) and groups:       # punctuation attaches to the previous word
            groups[-1] = list(groups[-1]) + [ph]
        else:
            groups.append(ph)
    return ' + '.join(' '.join(g) for g in groups)

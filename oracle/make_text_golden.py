"""Mints tests/golden/text_frontend.json.gz: inputs and the reference front-end's outputs
(phoneme string + tokens) for the reference's own corpus (data/*.txt) and for fuzzed Buckwalter
strings. Runs only in the build container (imports /root/reference)."""
import gzip
import json
import os
import random
import re
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference'
sys.path.insert(0, REF)
os.chdir(REF)
import text as ref_text  # noqa: E402


def corpus():
    lines = [l.strip() for l in open('data/infer_text.txt', encoding='utf-8') if l.strip()]
    for f in ['data/train_buckw.txt', 'data/test_buckw.txt']:
        for l in open(f, encoding='utf-8'):
            m = re.findall(r'"([^"]*)"', l)
            if len(m) >= 2:
                lines.append(m[1])
    arab = []
    for f in ['data/train_arab.txt', 'data/test_arab.txt']:
        for l in open(f, encoding='utf-8'):
            m = re.findall(r'"([^"]*)"', l)
            if len(m) >= 2:
                arab.append(m[1])
    return lines, arab[:300]


def fuzz(n, seed=0):
    rnd = random.Random(seed)
    cons = list("btjHxd*rzs$SDTZEgfqklmnhwy><}&'^")
    marks = list("aui~oFNK")
    specials = ['A', 'Y', 'p', '|', 'Al', 'All', 'wA', 'uw', 'iy', 'aw', 'ay', '~a', '~i', '~u', 'w~', 'y~', 'l~', ' ',
                ' ', ' ', '.', ',', '?', '!', '-', 'sil', 'h*A', 'Allh', 'lkn', '*lk', 'AlrHmn']
    out = []
    for _ in range(n):
        k = rnd.randint(1, 14)
        s = ''
        for _ in range(k):
            r = rnd.random()
            if r < 0.45:
                s += rnd.choice(cons) + (rnd.choice(marks) if rnd.random() < 0.7 else '')
            elif r < 0.75:
                s += rnd.choice(specials)
            else:
                s += rnd.choice(cons) + rnd.choice(marks) + rnd.choice(marks)
        out.append(s)
    return out


def main():
    lines, arab = corpus()
    fz = fuzz(4000)
    rec = {'buckwalter': [], 'arabic': [], 'fuzz': []}
    for l in lines:
        rec['buckwalter'].append([l, ref_text.buckwalter_to_phonemes(l), ref_text.buckwalter_to_tokens(l, append_space=False)])
    for l in arab:
        rec['arabic'].append([l, ref_text.arabic_to_buckwalter(l), ref_text.arabic_to_tokens(l, append_space=True)])
    for l in fz:
        try:
            rec['fuzz'].append([l, ref_text.buckwalter_to_phonemes(l), ref_text.buckwalter_to_tokens(l)])
        except Exception as e:   # the reference itself may raise on garbage; record that
            rec['fuzz'].append([l, None, repr(type(e).__name__)])
    rec['symbols'] = ref_text.symbols
    rec['round_trip'] = ref_text.buckwalter_to_arabic(lines[0])
    path = os.path.join(REPO, 'tests', 'golden', 'text_frontend.json.gz')
    with gzip.open(path, 'wt', encoding='utf-8') as f:
        json.dump(rec, f, ensure_ascii=False)
    print(path, os.path.getsize(path), len(lines), len(arab), len(fz))


if __name__ == '__main__':
    main()

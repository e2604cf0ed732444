"""Recipe for oracle/_ref/: the UNMODIFIED reference modules of the hot path, copied file by file from /root/reference
(read-only there; only in the build container) so that the reference itself — not only its restatement — can be timed on
the GPU box's host cores (`bench.py --impl reference`, `cpu_baseline.kind = "reference"`) and on its GPU through PyTorch
eager (`gpu_eager_baseline`). TEST / MEASUREMENT INFRASTRUCTURE ONLY: oracle/_ref/ is git-ignored (no reference source
enters the history) but not gpurun-ignored, so it travels with the snapshot like the built .so files. The files are
byte-for-byte copies (sha256 recorded in oracle/_ref/MANIFEST.json); the package __init__.py files are generated EMPTY so
that importing the model classes does not drag in the reference's text front-end, config loader or diacritizers.

    python oracle/make_ref.py          (also run by __graft_entry__.build() when /root/reference exists)
"""
import hashlib
import json
import os
import shutil
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference'
DST = os.path.join(REPO, 'oracle', '_ref')

# (source relative to /root/reference, destination relative to oracle/_ref)
FILES = [
    ('models/fastpitch/fastpitch/model.py', 'fastpitch/model.py'),                # FastPitch, TemporalPredictor, regulate_len
    ('models/fastpitch/fastpitch/transformer.py', 'fastpitch/transformer.py'),    # FFTransformer, MultiHeadAttn, PositionwiseConvFF
    ('models/fastpitch/fastpitch/attention.py', 'fastpitch/attention.py'),        # imported by model.py (training-time aligner)
    ('models/fastpitch/fastpitch/alignment.py', 'fastpitch/alignment.py'),        # imported by model.py (numba MAS)
    ('models/fastpitch/fastpitch/LICENSE', 'fastpitch/LICENSE'),
    ('vocoder/hifigan/models.py', 'hifigan/models.py'),                           # Generator, ResBlock1
    ('vocoder/hifigan/env.py', 'hifigan/env.py'),                                 # AttrDict
    ('vocoder/hifigan/LICENSE', 'hifigan/LICENSE'),
]


def build(verbose=True):
    if not os.path.isdir(REF):
        return False
    manifest = {}
    for src, dst in FILES:
        s, d = os.path.join(REF, src), os.path.join(DST, dst)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(s, d)
        with open(d, 'rb') as fh:
            manifest[dst] = {'source': src, 'sha256': hashlib.sha256(fh.read()).hexdigest()}
    for pkg in ('', 'fastpitch', 'hifigan'):
        open(os.path.join(DST, pkg, '__init__.py'), 'w').close()
    with open(os.path.join(DST, 'MANIFEST.json'), 'w') as fh:
        json.dump(manifest, fh, indent=1, sort_keys=True)
    if verbose:
        print('oracle/_ref: %d reference files copied unmodified from %s' % (len(FILES), REF))
    return True


if __name__ == '__main__':
    sys.exit(0 if build() else 1)

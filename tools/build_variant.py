"""Builds libttsb200_<tag>.so from the csrc/ + include/ of a git ref (default HEAD) next to the working-tree library, for
same-box A/B runs:  python tools/build_variant.py base [ref]   then   TTSB_LIB=tts_arabic_pytorch_b200/libttsb200_base.so python ...
The variant must export the same C ABI as the working tree's _lib.py (same header), or loading fails loudly."""
import os
import shutil
import subprocess
import sys
import tempfile
from concurrent.futures import ThreadPoolExecutor

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from tts_arabic_pytorch_b200 import build as b  # noqa: E402


def main():
    tag = sys.argv[1]
    ref = sys.argv[2] if len(sys.argv) > 2 else 'HEAD'
    out = os.path.join(REPO, 'tts_arabic_pytorch_b200', 'libttsb200_%s.so' % tag)
    tmp = tempfile.mkdtemp(prefix='ttsb_variant_')
    try:
        subprocess.check_call('git archive %s tts_arabic_pytorch_b200/csrc include | tar -x -C %s' % (ref, tmp), shell=True, cwd=REPO)
        csrc = os.path.join(tmp, 'tts_arabic_pytorch_b200', 'csrc')

        def one(src):
            obj = os.path.join(tmp, src.replace('.cu', '.o'))
            subprocess.check_call([b._nvcc()] + b.NVCC_FLAGS + ['-c', os.path.join(csrc, src), '-o', obj])
            return obj

        with ThreadPoolExecutor(max_workers=6) as ex:
            objs = list(ex.map(one, b.SOURCES))
        subprocess.check_call([b._nvcc(), '-shared', '-o', out] + objs)
        print(out)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == '__main__':
    main()

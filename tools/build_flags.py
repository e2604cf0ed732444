"""Builds libttsb200_<tag>.so from the WORKING TREE with extra nvcc flags (e.g. -DTTSB_EPI_DEBUG) for same-box experiments:
  python tools/build_flags.py dbg -DTTSB_EPI_DEBUG     then     TTSB_LIB=tts_arabic_pytorch_b200/libttsb200_dbg.so python ..."""
import os
import subprocess
import sys
import tempfile
from concurrent.futures import ThreadPoolExecutor

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from tts_arabic_pytorch_b200 import build as b  # noqa: E402


def main():
    tag, extra = sys.argv[1], sys.argv[2:]
    out = os.path.join(REPO, 'tts_arabic_pytorch_b200', 'libttsb200_%s.so' % tag)
    with tempfile.TemporaryDirectory(prefix='ttsb_flags_') as tmp:
        def one(src):
            obj = os.path.join(tmp, src.replace('.cu', '.o'))
            subprocess.check_call([b._nvcc()] + b.NVCC_FLAGS + extra + ['-c', os.path.join(b.CSRC, src), '-o', obj])
            return obj
        with ThreadPoolExecutor(max_workers=6) as ex:
            objs = list(ex.map(one, b.SOURCES))
        subprocess.check_call([b._nvcc(), '-shared', '-o', out] + objs)
    print(out)


if __name__ == '__main__':
    main()

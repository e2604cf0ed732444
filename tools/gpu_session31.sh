#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
exec > >(tee gpurun_out/session31.log) 2>&1
echo "=== pytest denoiser"; timeout 600 python -m pytest tests/test_gpu_denoiser.py tests/test_gpu_api.py -q -m gpu 2>&1 | tail -25
echo "=== denoiser timing"
timeout 300 python - <<'PY'
import torch, time
from tts_arabic_pytorch_b200.utils import synth
from tts_arabic_pytorch_b200.vocoder.hifigan.denoiser import Denoiser
from tts_arabic_pytorch_b200.vocoder.hifigan.env import AttrDict
from tts_arabic_pytorch_b200.vocoder.hifigan.models import Generator
g = Generator(AttrDict(synth.HIFIGAN_CONFIG)); g.load_state_dict(synth.hifigan_state_dict(1235)); g.eval(); g.remove_weight_norm(); g = g.cuda()
d = Denoiser(g).cuda()
B, N = 256, 131072
wav = torch.tanh(torch.randn(B, N, device='cuda') * 0.3)
n = torch.full((B,), N, dtype=torch.int32, device='cuda')
for _ in range(3): out = d.denoise_batch(wav, n, 0.005)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): out = d.denoise_batch(wav, n, 0.005)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print('cuda batched denoiser: %.2f ms per %d x %d samples = %.2e samples/s, %.0f GB/s algorithmic (40 B/sample)' % (ms, B, N, B * N / ms * 1e3, 40.0 * B * N / ms / 1e6))
t0 = time.perf_counter()
for b in range(32): r = d._forward_torch(wav[b:b+1], 0.005)
torch.cuda.synchronize()
print('torch per-utterance loop (reference formulation): %.2f ms per utterance' % ((time.perf_counter() - t0) / 32 * 1e3))
PY
echo "=== done"

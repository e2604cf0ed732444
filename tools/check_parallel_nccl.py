"""Multi-GPU check of the product path, one process per GPU over NCCL:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/check_parallel_nccl.py
Every rank calls parallel.synthesize on the SAME mixed-length utterance list (BASELINE config 5 shape, scaled down) with
each delivery mode; rank 0 then synthesizes the whole list as the reference's single padded batch
(models/fastpitch/networks.py:140-195) on its own GPU and requires every waveform to be BIT-identical.
Prints one line per mode and 'parallel nccl check: PASS' (exit code 0) or raises."""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    import tempfile
    import torch
    import torch.distributed as dist
    from tts_arabic_pytorch_b200 import parallel
    from tts_arabic_pytorch_b200.models.fastpitch import FastPitch2Wave
    from tts_arabic_pytorch_b200.utils import synth

    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    local = int(os.environ.get('LOCAL_RANK', rank))
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    n_utt = int(os.environ.get('TTSB_CHECK_UTTS', '48'))
    with tempfile.TemporaryDirectory(prefix='ttsb_par_%d_' % rank) as d:
        fp, hg, cj = synth.write_checkpoints(d, seed=1234)
        tts = FastPitch2Wave(fp, vocoder_sd=hg, vocoder_config=cj, arabic_in=False).cuda()
    g = torch.Generator().manual_seed(7)
    lens = torch.randint(16, 65, (n_utt,), generator=g).tolist()
    ids = [torch.randint(1, 40, (k,), generator=g) for k in lens]
    ref = None
    if rank == 0:
        ref, _ = tts.synthesize_ids(ids)
        ref = [w.clone() for w in ref]
    for mode in ('nccl', 'nccl_host', 'host_shm', 'host_shm', 'host_shm_fallback'):
        # host_shm twice: the second call reuses a persistent segment; then the fallback when no segment can be created
        if mode == 'host_shm_fallback':
            os.environ['TTSB_SHM_DISABLE'] = '1'
            parallel._close_shm_pool()
        out, stats = parallel.synthesize(tts, ids, deliver='host_shm' if mode.startswith('host_shm') else mode, return_stats=True)
        fr = torch.tensor([stats['frames'], stats['utterances']], dtype=torch.int64, device='cuda')
        allfr = [torch.empty_like(fr) for _ in range(world)]
        dist.all_gather(allfr, fr)
        if rank == 0:
            assert len(out) == n_utt
            worst = 0.0
            for k, (a, b) in enumerate(zip(out, ref)):
                a = a.cpu()
                assert a.shape == b.shape, (mode, k, a.shape, b.shape)
                worst = max(worst, float((a - b).abs().max()))
            assert worst == 0.0, 'mode %s: sharded result differs from the single padded batch by %g' % (mode, worst)
            print('deliver=%-9s world=%d: %d utterances bit-identical to the single padded batch; frames per rank %s, '
                  'utterances per rank %s' % (mode, world, n_utt, [int(x[0]) for x in allfr], [int(x[1]) for x in allfr]),
                  flush=True)
        else:
            assert out is None
    dist.barrier()
    if rank == 0:
        print('parallel nccl check: PASS', flush=True)
    dist.destroy_process_group()


if __name__ == '__main__':
    main()

"""The multi-GPU job path on real devices: NCCL world-2 run of dist.generate_job (broadcast of the weights, scatter of the
padded mels, per-rank generation through the persistent kernels, gather of the padded waveforms) against the same job on
one GPU.  Utterances are independent (wavenet/model.py:112-167 is row-wise), so the two must agree bit for bit.
Needs >= 2 GPUs (`gpurun --gpus 2 -- python -m pytest tests/test_dist_gpu.py -m gpu`); skipped otherwise."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _job():
    from tacotron_wavenet_vocoder_korean_b200 import synth
    kw = synth.cfg2(4)
    rs = np.random.RandomState(11)
    frames = [3, 1, 4, 2, 2, 5, 1, 3, 2]
    mels = [np.clip(rs.randn(f, 80) * 1.5, -4, 4).astype(np.float32) for f in frames]
    gcs = [i % 2 for i in range(len(frames))]
    return kw, mels, gcs


def _make_group_fn(kw, dev):
    from tacotron_wavenet_vocoder_korean_b200.wavenet import WaveNetModel
    cache = {}

    def generate_group(state, gmels, ggc, gidx):
        if 'net' not in cache:
            net = WaveNetModel(train_mode=False, device=dev, **kw)
            net.load_state_dict(state)
            cache['net'] = net
        net = cache['net']
        rows = len(gmels)
        fmax = max(m.shape[0] for m in gmels)
        mel = np.zeros((rows, fmax, 80), np.float32)
        for r, m in enumerate(gmels):
            mel[r, :m.shape[0]] = m
        T = fmax * 300
        # per-utterance noise keyed by the job index: the result must not depend on which rank / group ran it
        uni = np.stack([np.random.RandomState(500 + int(i)).uniform(1e-5, 1 - 1e-5, (T, 11)).astype(np.float32) for i in gidx])
        x0 = np.zeros((rows, 1), np.float32)
        T_row = [int(m.shape[0]) * 300 for m in gmels]
        wav = net.generate(T, x0, uni, mel=mel, gc_ids=ggc, T_row=T_row).cpu().numpy()
        return [wav[r, :T_row[r]] for r in range(rows)]
    return generate_group


def _worker(rank, world, port, ret):
    from tacotron_wavenet_vocoder_korean_b200 import synth, dist as wdist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    try:
        kw, mels, gcs = _job()
        state = synth.make_weights(**kw) if rank == 0 else None
        out = wdist.generate_job(_make_group_fn(kw, dev), state, mels if rank == 0 else None, gcs if rank == 0 else None,
                                 batch=kw['batch_size'], hop=300, src=0, device=dev)
        if rank == 0:
            ret.put([np.asarray(o) for o in out])
        else:
            assert out is None
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs (gpurun --gpus 2)')
def test_generate_job_nccl_world2_equals_single_gpu():
    from tacotron_wavenet_vocoder_korean_b200 import synth
    ctx = mp.get_context('spawn')
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    got = ret.get(timeout=600)
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    # the same job on one GPU, no process group: every utterance alone through the same kernels
    kw, mels, gcs = _job()
    fn = _make_group_fn(kw, torch.device('cuda', 0))
    state = synth.make_weights(**kw)
    for i, (m, g) in enumerate(zip(mels, gcs)):
        exp = fn(state, [m], [g], [i])[0]
        assert len(got[i]) == m.shape[0] * 300 and np.array_equal(got[i], exp), 'utterance %d differs' % i

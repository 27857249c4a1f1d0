"""Host-side logic of the multi-GPU driver, exercised with gloo on CPU (world_size 2)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tacotron_wavenet_vocoder_korean_b200 import dist as wdist


def test_lpt_assignment_balances_and_covers():
    lengths = [160, 40, 300, 10, 10, 220, 90, 90, 5]
    for world in (1, 2, 4, 8):
        a = wdist.lpt_assign(lengths, world)
        assert sorted(i for r in a for i in r) == list(range(len(lengths)))
        loads = [sum(lengths[i] for i in r) for r in a]
        assert max(loads) - min(loads) <= max(lengths)
    assert wdist.lpt_assign([], 2) == [[], []]
    g = wdist.make_groups([0, 1, 2, 3, 4], [5, 9, 1, 7, 3], 2)
    assert g == [[1, 3], [0, 4], [2]]


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_group(state, mels, gcs, idx):
    # stand-in for the device work: a deterministic function of (weights, mel, speaker id)
    k = float(state['w'].sum())
    return [np.repeat(m.mean(axis=1), 3) * k + (g if gcs is not None else 0) for m, g in zip(mels, gcs or [0] * len(mels))]


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        rs = np.random.RandomState(0)
        lens = [7, 3, 12, 5, 9]
        mels = [rs.randn(t, 4).astype(np.float32) for t in lens] if rank == 0 else None
        gcs = [0, 1, 1, 0, 1] if rank == 0 else None
        state = {'w': np.arange(6, dtype=np.float32).reshape(2, 3), 'b': np.ones(2, np.float32)} if rank == 0 else None
        out = wdist.generate_job(_fake_group, state, mels, gcs, batch=2, hop=3)
        if rank == 0:
            exp = _fake_group({'w': np.arange(6, dtype=np.float32)}, mels, gcs, None)
            ok = all(np.allclose(a, b) for a, b in zip(out, exp)) and [len(o) for o in out] == [3 * t for t in lens]
            ret.put(bool(ok))
        else:
            assert out is None
    finally:
        dist.destroy_process_group()


def test_scatter_generate_gather_world2():
    ctx = mp.get_context('spawn')
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret.get(timeout=5) is True


# ---- end-to-end text -> wav job (pipeline.py) -----------------------------------------------------------
class _FakeTTS(object):
    def synthesize(self, texts, speaker_ids, **kw):
        return [np.full(3 * len(t) + s, float(len(t)), np.float32) for t, s in zip(texts, speaker_ids)]


def _job_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from tacotron_wavenet_vocoder_korean_b200 import pipeline
    texts = ['가' * n for n in (5, 1, 9, 3, 7)] if rank == 0 else None
    spk = [0, 1, 0, 1, 1] if rank == 0 else None
    out = pipeline.run_job(_FakeTTS, texts, spk, src=0)
    if rank == 0:
        ret.put([o.tolist() for o in out])
    dist.destroy_process_group()


def test_text_job_shards_and_gathers_in_order_world2():
    from tacotron_wavenet_vocoder_korean_b200 import pipeline
    a = pipeline.shard_sentences(['a' * n for n in (5, 1, 9, 3, 7)], 2)
    assert sorted(i for r in a for i in r) == [0, 1, 2, 3, 4]
    ctx = mp.get_context('spawn')
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_job_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    out = ret.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    exp = [[float(n)] * (3 * n + s) for n, s in zip((5, 1, 9, 3, 7), (0, 1, 0, 1, 1))]
    assert out == exp


# ---- data-parallel training exchange (dist.allreduce_mean_) -------------------------------------------------------------
def _dp_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import torch
    from tacotron_wavenet_vocoder_korean_b200 import dist as wdist
    g = torch.arange(6, dtype=torch.float32) * (rank + 1)          # rank 0: [0..5], rank 1: 2*[0..5]
    loss = torch.tensor([float(rank + 1)])
    scale, mean_loss = wdist.allreduce_mean_(g, loss)
    p = torch.full((4,), float(rank))
    wdist.broadcast_flat_(p, src=1)
    if rank == 0:
        ret.put((scale, g.tolist(), float(mean_loss), float(loss), p.tolist()))
    dist.destroy_process_group()


def test_data_parallel_gradient_exchange_world2():
    import torch
    from tacotron_wavenet_vocoder_korean_b200 import dist as wdist
    g = torch.ones(3)
    assert wdist.allreduce_mean_(g, None) == (1.0, None) and g.tolist() == [1.0, 1.0, 1.0]     # no process group: no-op
    ctx = mp.get_context('spawn')
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    scale, g, mean_loss, own_loss, p0 = ret.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert scale == 0.5 and g == [0.0, 3.0, 6.0, 9.0, 12.0, 15.0]     # SUM in the buffer, 1/world folded into Adam
    assert mean_loss == 1.5 and own_loss == 1.0 and p0 == [1.0] * 4

"""End-to-end text -> 24 kHz waveform on N GPUs (BASELINE configs[4], "cfg-5"): every rank synthesises its own
share of the sentences (weak scaling: --per_gpu sentences per GPU), Tacotron -> mel (kept in HBM) -> WaveNet.
Launch:  python scripts/bench_e2e.py            (1 GPU)
         python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_e2e.py
Prints one JSON line on rank 0: aggregate audio samples/s and real-time factor (max over ranks of the device time).
Mels are capped at --frames (default 160 = 2 s) so the run stays short; weights are seeded synthetic."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--per_gpu', type=int, default=32)
    ap.add_argument('--frames', type=int, default=160)
    ap.add_argument('--wn_batch', type=int, default=16)
    ap.add_argument('--iters', type=int, default=2)
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    from bench_taco import make_texts
    from tacotron_wavenet_vocoder_korean_b200 import pipeline, synth
    from tests.taco_helpers import Bag
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl')
    hp = dict(synth.TACO_HP)
    kw = synth.cfg2(a.wn_batch)
    tts = pipeline.TextToSpeech(Bag(hp), synth.make_taco_weights(hp, 2), 2, kw, synth.make_weights(**kw))
    texts = make_texts(a.per_gpu * world)[rank::world]
    spk = [(i % 2) for i in range(len(texts))]
    n_samples = 0
    ms = []
    for it in range(a.iters + 1):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        wavs = tts.synthesize(texts, spk, attention_trim=False, max_mel_frames=a.frames, seed=it)
        e1.record()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        n_samples = sum(len(w) for w in wavs)
        if it > 0:
            ms.append(max(e0.elapsed_time(e1), 0.0))
            last_wall = wall
    t = torch.tensor([float(np.mean(ms)), float(n_samples)], device='cuda')
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        step_ms, total = float(tmax[0]), float(tsum[1])
    else:
        step_ms, total = float(t[0]), float(t[1])
    if rank == 0:
        print(json.dumps({
            "metric": "end-to-end text->wav audio samples/s (Tacotron -> mel -> WaveNet 24 kHz), %d sentences per GPU, %d mel frames each" % (a.per_gpu, a.frames),
            "value": total / (step_ms * 1e-3), "unit": "samples/s", "rtf": total / (step_ms * 1e-3) / 24000.0, "n_gpus": world,
            "ms_per_job": step_ms, "wall_ms_rank0": last_wall, "sentences": a.per_gpu * world, "scaling": "weak", "wavenet_rows_in_flight": a.wn_batch,
            "data": "synthetic weights, own Korean sentences", "dtype": "f32"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

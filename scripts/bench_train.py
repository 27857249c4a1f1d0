"""WaveNet training-step throughput (BASELINE configs[3], "cfg-4"): 30-layer R=D=128 S=512 MoL-10 mel+speaker conditioned
WaveNet, bf16 tensor-core GEMMs / fp32 master weights, batch 64 crops x 7500 samples per GPU (the config says 7680; the
reference's feeder keeps crops hop(300)-aligned, datafeeder_wavenet.py:41-47 -> 7500), MoL loss,
Adam + exponential decay + EMA; data-parallel over N GPUs with ONE NCCL all-reduce of the flat fp32 gradient buffer.
Launch:  python scripts/bench_train.py [--batch 64 --samples 7500 --steps 5 --warmup 3]
         python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_train.py
Prints one JSON line on rank 0 (device time, max over ranks): audio samples trained per second, ms/step, the split
loss+grads / all-reduce / apply, achieved GEMM TFLOP/s against MEASURED_PEAKS.json's bf16 peak."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=64)
    ap.add_argument('--samples', type=int, default=7500)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--dtype', default='bf16')
    ap.add_argument('--one_hot', action='store_true', help='scalar_input=False: mu-law one-hot input + softmax cross-entropy head')
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    from tacotron_wavenet_vocoder_korean_b200 import synth, dist as wdist
    from tacotron_wavenet_vocoder_korean_b200.wavenet.train import WaveNetTrainer
    from tacotron_wavenet_vocoder_korean_b200.hparams import hparams
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl')
    kw = synth.cfg2(a.batch)
    if a.one_hot:
        kw = dict(kw, scalar_input=False)
    tr = WaveNetTrainer(a.samples, dtype=a.dtype, **kw)
    tr.load_state_dict(synth.make_weights(**kw))
    tr.sync_params(0)
    rs = np.random.RandomState(100 + rank)
    t = np.arange(a.samples)[None, :]
    wav = np.clip(0.5 * np.sin(2 * np.pi * t * rs.uniform(0.005, 0.05, (a.batch, 1))) + 0.1 * rs.randn(a.batch, a.samples), -1, 1).astype(np.float32)
    mel = np.clip(rs.randn(a.batch, a.samples // 300, 80) * 1.5, -4, 4).astype(np.float32)
    gc = (np.arange(a.batch) % 2).astype(np.int32)
    wav_d, mel_d, gc_d = torch.from_numpy(wav).cuda(), torch.from_numpy(mel).cuda(), torch.from_numpy(gc).cuda()
    hp = hparams

    def ev():
        return torch.cuda.Event(enable_timing=True)
    losses, seg = [], []
    for it in range(a.warmup + a.steps):
        if it == a.warmup:
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e_start = ev()
            e_start.record()
        e0, e1, e2, e3 = ev(), ev(), ev(), ev()
        e0.record()
        loss = tr.loss_and_grads(wav_d, mel_d, gc_d)
        e1.record()
        scale, _ = wdist.allreduce_mean_(tr.grads)
        e2.record()
        from tacotron_wavenet_vocoder_korean_b200.wavenet.train import learning_rate_at
        tr.apply(learning_rate_at(hp, tr.global_step), grad_scale=scale)
        e3.record()
        if it >= a.warmup:
            seg.append((e0, e1, e2, e3))
        losses.append(loss.clone())
    e_end = ev()
    e_end.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    total_ms = e_start.elapsed_time(e_end)
    parts = np.array([[x[0].elapsed_time(x[1]), x[1].elapsed_time(x[2]), x[2].elapsed_time(x[3])] for x in seg]).mean(0)
    tt = torch.tensor([total_ms], device='cuda')
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    step_ms = float(tt[0]) / a.steps
    if rank == 0:
        info = tr.info()
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        tf = info['flops_per_step'] / (parts[0] * 1e-3) / 1e12
        peak = None
        for k in ('bf16_tflops_sustained', 'bf16_tflops', 'bf16_dense_tflops', 'tensor_tflops'):
            if k in peaks:
                peak = float(peaks[k])
                break
        print(json.dumps({
            "metric": "WaveNet training throughput, audio samples/s (cfg-4: 30-layer R=D=128 S=512 %s, batch %d x %d samples per GPU, %s, Adam+EMA)" % ("mu-law one-hot input + softmax-256" if a.one_hot else "MoL-10", a.batch, a.samples, a.dtype),
            "value": world * a.batch * a.samples / (step_ms * 1e-3), "unit": "samples/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": step_ms, "split_ms": {"loss_and_grads": float(parts[0]), "grad_allreduce": float(parts[1]), "adam_ema": float(parts[2])},
            "gemm_tflops_achieved": tf, "gemm_tflops_peak": peak, "gemm_flops_per_step": info['flops_per_step'],
            "trained_outputs_per_step": world * a.batch * info['output_width'], "scaling": "weak", "dtype": a.dtype,
            "loss_first_last": [float(losses[0].item()), float(losses[-1].item())],
            "gemm_launches_per_step": info['gemm_launches'] // (a.warmup + a.steps), "fused_tcgen05_launches_per_step": info['fused_launches'] // (a.warmup + a.steps), "kernel_launches_per_step": info['kernel_launches'] // (a.warmup + a.steps),
            "workspace_gb": info['workspace_bytes'] / 2 ** 30, "n_trainable": info['n_trainable'], "data": "synthetic"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

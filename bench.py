#!/usr/bin/env python
"""bench.py -- WaveNet autoregressive generation throughput on B200 (BASELINE.json metric).

A "step" is one pass of the hot path over one batch: cfg-2 (BASELINE configs[1]: 30-layer 3x10 WaveNet,
R=D=128, S=512, MoL-10, mel + speaker conditioned) generating 8 utterances x 48 000 samples (2 s @ 24 kHz)
per GPU -- upsample network + ONE persistent-kernel launch.  Synthetic seeded mel / uniforms / weights.

  value      samples/s over all GPUs, inputs resident in HBM, CUDA events around K steps, max over ranks
  e2e        same metric through WaveNetModel.generate_host (wn_generate_host C ABI) with pinned HOST buffers:
             H2D of mel/uniforms/initial sample and D2H of the waveform inside the timed region
  roofline   HBM-bandwidth roofline of the per-sample matvec chain: algorithmic bytes per launch /
             measured duration of the persistent kernel, against MEASURED_PEAKS.json
  cpu_baseline  the CPU oracle (a port: the TF reference cannot run here) on a bounded sample, host cores

`--impl reference` times the CPU restatement alone (all host threads) on the same workload definition.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

T_STEPS = 48000          # 2 s @ 24 kHz
ROWS = 8
T_MEL = 160
SAMPLE_RATE = 24000
WORKLOAD = ("cfg2: 30-layer (3x10 dilation cycles) WaveNet, R=D=128, S=512, MoL-10 head, scalar input (ifw 32), "
            "mel local conditioning (80 ch, upsample 5x5x12) + speaker embedding, batch=8 utterances x 48000 samples "
            "(2 s @ 24 kHz) per GPU, free-running generation")


def make_job(rank):
    from tacotron_wavenet_vocoder_korean_b200 import synth
    kw = synth.cfg2(ROWS)
    w = synth.make_weights(**kw)
    rs = np.random.RandomState(100 + rank)
    mel = np.clip(rs.randn(ROWS, T_MEL, 80) * 1.5, -4, 4).astype(np.float32)
    uniforms = rs.uniform(1e-5, 1 - 1e-5, (ROWS, T_STEPS, 11)).astype(np.float32)
    x0 = (2 * rs.rand(ROWS, 1) - 1).astype(np.float32)
    gc = [0, 0, 0, 0, 1, 1, 1, 1]
    return kw, w, mel, uniforms, x0, gc


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons while the timed region runs (B200_PROFILING.md clocks line)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.sm_max = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {getattr(nv, 'nvmlClocksThrottleReasonHwSlowdown', 0x8): 'hw_slowdown',
                 getattr(nv, 'nvmlClocksThrottleReasonHwThermalSlowdown', 0x40): 'hw_thermal_slowdown',
                 getattr(nv, 'nvmlClocksThrottleReasonSwThermalSlowdown', 0x20): 'sw_thermal_slowdown',
                 getattr(nv, 'nvmlClocksThrottleReasonSwPowerCap', 0x4): 'sw_power_cap'}
        while not self.stop_flag:
            try:
                self.sm.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.sm_max,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


def cpu_model_name():
    """the host CPU the baselines ran on."""
    try:
        with open('/proc/cpuinfo') as f:
            for line in f:
                if line.lower().startswith('model name'):
                    return line.split(':', 1)[1].strip()
    except Exception:
        pass
    return None


def cpu_port_throughput(steps, threads):
    """The CPU oracle (plain-C restatement of the reference loop) on `ROWS` rows x `steps` steps of the
    workload; rows are spread over host threads.  Returns samples/s."""
    import oracle
    from tacotron_wavenet_vocoder_korean_b200 import synth
    kw, w, mel, uniforms, x0, gc = make_job(0)
    om = oracle.OracleModel(**kw)
    om.set_weights(w)
    t_mel = (steps + 299) // 300
    lc = om.upsample(mel[:, :t_mel])
    oracle.lib().orc_set_threads(threads)
    t0 = time.perf_counter()
    om.generate(steps, x0, uniforms[:, :steps], lc_up=lc, gc_ids=np.asarray(gc, np.int32))
    dt = time.perf_counter() - t0
    oracle.lib().orc_set_threads(1)
    return ROWS * steps / dt


def cpu_best_effort_throughput(steps, threads):
    """BASELINE.md section 3 "B-cpu": oracle/wn_cpu_best.h -- the same arithmetic as the port (bit-identical output), organised
    like the GPU path: every thread keeps a packed column slice of all weights in its own cache, all rows advance together,
    two spin barriers per layer.  Returns samples/s."""
    import oracle
    kw, w, mel, uniforms, x0, gc = make_job(0)
    om = oracle.OracleModel(**kw)
    om.set_weights(w)
    lc = om.upsample(mel[:, :(steps + 299) // 300])
    t0 = time.perf_counter()
    om.generate(steps, x0, uniforms[:, :steps], lc_up=lc, gc_ids=np.asarray(gc, np.int32), best_effort_threads=threads)
    return ROWS * steps / (time.perf_counter() - t0)


def best_effort_thread_counts():
    """Column slices are whole 8-float vectors of the 128 gated channels: at most 16 threads, powers of two."""
    cores = os.cpu_count() or 1
    return [n for n in (16, 8, 4, 2, 1) if n <= cores][:2] or [1]


def cpu_best_effort(steps):
    """Best thread count of a short trial, then the bounded sample.  Returns (samples/s, threads)."""
    trials = {n: cpu_best_effort_throughput(60, n) for n in best_effort_thread_counts()}
    n = max(trials, key=trials.get)
    return cpu_best_effort_throughput(steps, n), n


def cpu_reference_structure_throughput(steps):
    """SURVEY.md 8(d) "reference-structure" baseline: the numpy restatement driven by the same per-sample Python loop as
    generate.py:202-233, every queue shifted by a full copy per step (model.py:122,125,145), numpy's BLAS threads as they
    come.  Returns samples/s over ROWS rows."""
    from oracle import np_oracle
    kw, w, mel, uniforms, x0, gc = make_job(0)
    net = np_oracle.NumpyWaveNet(**kw)
    net.set_weights(w)
    lc = net.create_upsample(mel[:, :(steps + 299) // 300])
    np_oracle.generate(net, 10, np.asarray(x0).reshape(ROWS, -1)[:, :1], uniforms[:, :10], lc_up=lc, gc_ids=np.asarray(gc))
    t0 = time.perf_counter()
    np_oracle.generate(net, steps, np.asarray(x0).reshape(ROWS, -1)[:, :1], uniforms[:, :steps], lc_up=lc, gc_ids=np.asarray(gc))
    return ROWS * steps / (time.perf_counter() - t0)


CPU_SAMPLE_STEPS = 600     # the cpu_baseline leg and the --impl reference arm time the same bounded sample


def run_reference(args, rank, world):
    """The reference's own CPU path cannot run (TensorFlow 1.x absent): time its CPU
    restatement (oracle port) with all host threads; each step is a bounded sample of the workload."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    threads = min(cores, ROWS)
    sample_steps = CPU_SAMPLE_STEPS
    # two organisations of the same arithmetic (bit-identical outputs): rows over threads (the plain port) and weight-stationary
    # threads (wn_cpu_best.h, BASELINE.md "B-cpu"); the warm-up picks the faster one on this host, the timed steps run it
    trial = {("port", threads): 0.0}
    for n in best_effort_thread_counts():
        trial[("best_effort", n)] = 0.0
    for _ in range(max(1, args.warmup)):
        for (kind, n) in trial:
            v = cpu_port_throughput(60, n) if kind == "port" else cpu_best_effort_throughput(60, n)
            trial[(kind, n)] = max(trial[(kind, n)], v)
    (impl_kind, threads) = max(trial, key=trial.get)
    run = (lambda st: cpu_port_throughput(st, threads)) if impl_kind == "port" else (lambda st: cpu_best_effort_throughput(st, threads))
    t0 = time.perf_counter()
    vals = [run(sample_steps) for _ in range(args.steps)]
    dt = time.perf_counter() - t0
    # each step's throughput is timed around the generation loop only (as generate.py:199 does); the
    # wall time per step additionally contains building the oracle model and the upsampling
    v = ROWS * sample_steps * args.steps / sum(ROWS * sample_steps / x for x in vals)
    line = {"impl": "reference", "metric": "wavenet_generation_samples_per_sec", "value": v, "unit": "samples/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "rows_per_gpu": ROWS, "steps_per_row": T_STEPS},
            "cpu_baseline": {"value": v, "unit": "samples/s", "cores": threads, "kind": "port", "cpu_model": cpu_model_name(),
                             "host_cores": cores, "organisation": impl_kind,
                             "sample": "%d rows x %d steps of the workload per bench step on %d host threads: %s "
                                       "(the TF 1.x reference is not installable)" % (
                                           ROWS, sample_steps, threads,
                                           "oracle/wn_cpu_best.h, weight-stationary threads, all rows together (BASELINE.md B-cpu)"
                                           if impl_kind == "best_effort" else "oracle/wn_oracle.c, rows over threads"),
                             "warmup_trials_samples_per_sec": {"%s/%d threads" % k: x for k, x in trial.items()},
                             "note": "a stated baseline, not the target: the faster of the two CPU organisations of the oracle's arithmetic on this "
                                     "host (both bit-identical to the oracle, built with -ffp-contract=off, AVX2 + FMA); a GPU/CPU ratio says "
                                     "nothing about kernel quality -- roofline.frac does"},
            "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "per_step_values": vals}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------------
# extra cells of BASELINE.md section 5 / BASELINE.json configs, timed by the same run (each guarded: a failure is
# recorded as a string and never costs the headline line)
def guarded(name, fn, out):
    try:
        out[name] = fn()
    except Exception as e:       # noqa: BLE001
        out[name] = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}


def ev_pair():
    import torch
    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def extra_cfg1(dev, peak):
    """BASELINE configs[0]: 10-layer mu-law 256 WaveNet, 0.5 s @ 16 kHz = 8000 steps, unconditioned, batch 1."""
    import torch
    from tacotron_wavenet_vocoder_korean_b200 import synth
    from tacotron_wavenet_vocoder_korean_b200.wavenet import WaveNetModel
    kw = synth.cfg1()
    net = WaveNetModel(train_mode=False, device=dev, **kw)
    net.load_state_dict(synth.make_weights(**kw))
    T = 8000
    rs = np.random.RandomState(1)
    uni = torch.from_numpy(rs.random_sample((1, T))).to(dev)
    x0 = torch.from_numpy(np.random.RandomState(0).randint(256, size=(1, 1)).astype(np.float32)).to(dev)
    for _ in range(2):
        net.generate(T, x0, uni)
    e0, e1 = ev_pair()
    e0.record()
    for _ in range(3):
        net.generate(T, x0, uni, sync=False)
    e1.record()
    torch.cuda.synchronize()
    net.sync_check()
    ms = e0.elapsed_time(e1) / 3
    info = net.info()
    L, R = len(kw['dilations']), kw['residual_channels']
    b_step = 4 * info['p_hot'] + 4 * (2 * L * R + 2)
    sps = T / (ms / 1e3)
    return {"samples_per_sec": sps, "rtf_16khz": sps / 16000.0, "us_per_step": 1e3 * ms / T, "roofline_frac": b_step * sps / 1e9 / peak,
            "algorithmic_bytes_per_step": b_step, "grid": info['grid'], "kernel": "wn_persistent_kernel_s<ShapeCfg1>" if info['static_shape'] == 2 else "runtime-shaped"}


def extra_cfg3(dev):
    """BASELINE configs[2]: Tacotron text->mel, 32 Korean sentences (2 speakers), attention decoder + CBHG, 200 decoder steps."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, 'scripts'))
    from bench_taco import make_texts
    from tacotron_wavenet_vocoder_korean_b200 import synth
    from tacotron_wavenet_vocoder_korean_b200.tacotron import Tacotron
    from tacotron_wavenet_vocoder_korean_b200.text import text_to_sequence, prepare_inputs
    from tests.taco_helpers import Bag
    hp = dict(synth.TACO_HP)
    m = Tacotron(Bag(hp))
    m.load_state_dict(synth.make_taco_weights(hp, 2))
    ids = prepare_inputs([text_to_sequence(t) for t in make_texts(32)])
    lens = np.array([int(np.argmax(s == 1)) + 1 for s in ids], np.int32)
    spk = (np.arange(32) % 2).astype(np.int32)
    for _ in range(3):
        m.initialize(ids, lens, 2, spk, rnn_decoder_test_mode=True, n_steps=200, want_linear=True)
    torch.cuda.synchronize()
    e0, e1 = ev_pair()
    e0.record()
    for _ in range(5):
        m.initialize(ids, lens, 2, spk, rnn_decoder_test_mode=True, n_steps=200, want_linear=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    return {"sentences_per_sec": 32 / (ms / 1e3), "ms_per_32_sentences": ms, "decoder_steps": 200, "mel_frames_per_sentence": 200 * hp['reduction_factor']}


def extra_cfg4(world):
    """BASELINE configs[3]: WaveNet training step bf16, 30 layers, batch 64 x 7500 samples per GPU, MoL loss, Adam + EMA,
    data-parallel with ONE NCCL all-reduce of the flat fp32 gradient buffer."""
    import torch
    import torch.distributed as dist
    from tacotron_wavenet_vocoder_korean_b200 import synth, dist as wdist
    from tacotron_wavenet_vocoder_korean_b200.wavenet.train import WaveNetTrainer, learning_rate_at
    from tacotron_wavenet_vocoder_korean_b200.hparams import hparams
    B, S = 64, 7500
    kw = synth.cfg2(B)
    tr = WaveNetTrainer(S, dtype='bf16', **kw)
    tr.load_state_dict(synth.make_weights(**kw))
    tr.sync_params(0)
    rs = np.random.RandomState(100 + int(os.environ.get("RANK", "0")))
    t = np.arange(S)[None, :]
    wav = np.clip(0.5 * np.sin(2 * np.pi * t * rs.uniform(0.005, 0.05, (B, 1))) + 0.1 * rs.randn(B, S), -1, 1).astype(np.float32)
    mel = np.clip(rs.randn(B, S // 300, 80) * 1.5, -4, 4).astype(np.float32)
    wav_d, mel_d = torch.from_numpy(wav).cuda(), torch.from_numpy(mel).cuda()
    gc_d = torch.from_numpy((np.arange(B) % 2).astype(np.int32)).cuda()

    def one():
        loss = tr.loss_and_grads(wav_d, mel_d, gc_d)
        scale, _ = wdist.allreduce_mean_(tr.grads)
        tr.apply(learning_rate_at(hparams, tr.global_step), grad_scale=scale)
        return loss
    for _ in range(3):
        one()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = ev_pair()
    e0.record()
    for _ in range(3):
        loss = one()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    tt = torch.tensor([e0.elapsed_time(e1) / 3], device='cuda')
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms = float(tt[0])
    info = tr.info()
    del tr
    torch.cuda.empty_cache()
    return {"ms_per_step": ms, "samples_per_sec": world * B * S / (ms / 1e3), "n_gpus": world, "batch_per_gpu": B, "samples_per_crop": S,
            "gemm_tflops": info['flops_per_step'] / (ms * 1e-3) / 1e12, "loss": float(loss.item()),
            "collective": ("one NCCL all_reduce of the %d-float gradient buffer per step" % int(info.get('n_trainable', 0))) if world > 1 else "none (1 GPU)"}


def extra_cfg5(world, rank, dev):
    """BASELINE configs[4]: end to end, 32 sentences per GPU (256 at 8 GPUs): Tacotron -> mel (stays in HBM) -> WaveNet 24 kHz."""
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, 'scripts'))
    from bench_taco import make_texts
    from tacotron_wavenet_vocoder_korean_b200 import pipeline, synth
    from tests.taco_helpers import Bag
    hp = dict(synth.TACO_HP)
    kw = synth.cfg2(16)
    tts = pipeline.TextToSpeech(Bag(hp), synth.make_taco_weights(hp, 2), 2, kw, synth.make_weights(**kw))
    texts = make_texts(32 * world)[rank::world]
    spk = [(i % 2) for i in range(len(texts))]
    ms, n = None, 0
    for it in range(2):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = ev_pair()
        e0.record()
        wavs = tts.synthesize(texts, spk, attention_trim=False, max_mel_frames=160, seed=it)
        e1.record()
        torch.cuda.synchronize()
        ms, n = e0.elapsed_time(e1), sum(len(w) for w in wavs)
    t = torch.tensor([ms, float(n)], device=dev)
    if world > 1:
        tm = t.clone()
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        ms, n = float(tm[0]), float(t[1])
    del tts
    torch.cuda.empty_cache()
    return {"sentences": 32 * world, "samples_per_sec": n / (ms / 1e3), "rtf": n / (ms / 1e3) / SAMPLE_RATE, "ms_per_job": ms, "n_gpus": world,
            "mel_frames_per_sentence": 160, "wavenet_rows_in_flight": 16}


JOB_ROWS = 16      # rows per launch of the 64-utterance job


def extra_job64(world, rank, dev, kw, w, fast_act):
    """Strong scaling on the REAL multi-GPU job path (dist.generate_job): a fixed job of 64 ragged utterances
    (60..120 mel frames) on rank 0 -> NCCL broadcast of the weights, NCCL scatter of the padded mels (LPT shares), groups of
    <= 16 rows through the persistent kernels (16 rows in flight give 1.1x the sample rate of 8), NCCL gather of the padded
    waveforms back to rank 0."""
    import torch
    import torch.distributed as dist
    from tacotron_wavenet_vocoder_korean_b200 import dist as wdist
    from tacotron_wavenet_vocoder_korean_b200.wavenet import WaveNetModel
    rs = np.random.RandomState(7)
    frames = rs.randint(60, 121, size=64)
    mels = [np.clip(rs.randn(int(f), 80) * 1.5, -4, 4).astype(np.float32) for f in frames] if rank == 0 else None
    gcs = [int(i % 2) for i in range(64)] if rank == 0 else None
    nets = {}

    def generate_group(state, gmels, ggc, gidx):
        if 'net' not in nets:
            net = WaveNetModel(train_mode=False, device=dev, fast_act=fast_act, **dict(kw, batch_size=JOB_ROWS))
            net.load_state_dict(state)
            nets['net'] = net
        net = nets['net']
        rows = len(gmels)
        fmax = max(m.shape[0] for m in gmels)
        mel = np.zeros((rows, fmax, 80), np.float32)
        for r, m in enumerate(gmels):
            mel[r, :m.shape[0]] = m
        T = fmax * 300
        g = torch.Generator(device=dev)
        g.manual_seed(1000 + int(gidx[0]))
        uni = torch.empty((rows, T, 11), dtype=torch.float32, device=dev).uniform_(1e-5, 1 - 1e-5, generator=g)
        x0 = torch.zeros((rows, 1), dtype=torch.float32, device=dev)
        T_row = [int(m.shape[0]) * 300 for m in gmels]
        wav = net.generate(T, x0, uni, mel=torch.from_numpy(mel).to(dev), gc_ids=ggc, T_row=T_row).cpu().numpy()
        return [wav[r, :T_row[r]] for r in range(rows)]
    times = []
    for it in range(2):
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = wdist.generate_job(generate_group, w if rank == 0 else None, mels, gcs, JOB_ROWS, 300, src=0, device=dev)
        torch.cuda.synchronize()
        dist.barrier()
        times.append(time.perf_counter() - t0)
    t = torch.tensor([times[-1]], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    n = int(frames.sum()) * 300
    if rank == 0:
        assert out is not None and len(out) == 64 and all(len(o) == int(f) * 300 for o, f in zip(out, frames))
    return {"utterances": 64, "rows_in_flight": JOB_ROWS, "samples": n, "seconds": float(t[0]), "samples_per_sec": n / float(t[0]), "n_gpus": world, "scaling": "strong",
            "collectives": "broadcast(weights 22 MB) + scatter(padded mels) + gather(padded waveforms), NCCL" if dist.get_backend() == "nccl" else dist.get_backend(),
            "first_call_seconds_incl_model_setup": times[0]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="headline line only (skip the cfg-1/3/4/5, batch-1 and 64-utterance job cells)")
    ap.add_argument("--fast-activation", action="store_true",
                    help="ex2/rcp.approx gate (WN_FLAG_FAST_ACT: MoL logits within 3e-6 of the pinned arithmetic, tolerance 1e-4) for the "
                         "headline instead of the default pinned exp32 + IEEE-divide gate, which is bit-identical to the CPU oracle; "
                         "the other one is always reported as an extra key (measured: < 5 %% apart)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from tacotron_wavenet_vocoder_korean_b200.wavenet import WaveNetModel
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU port")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", str(29500 + (os.getpid() % 2000)))
    # one process per GPU; at N = 1 a single-rank group, so that the job path below is the same code
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    W = max(args.warmup, 3)
    fast = args.fast_activation

    kw, w, mel, uniforms, x0, gc = make_job(rank)
    net = WaveNetModel(train_mode=False, device=dev, fast_act=fast, **kw)
    net.load_state_dict(w)
    info = net.info()
    mel_d = torch.from_numpy(mel).to(dev)
    uni_d = torch.from_numpy(uniforms).to(dev)
    x0_d = torch.from_numpy(x0).to(dev)

    def step():
        # mel frames in, waveform out: create_upsample is evaluated inside the kernels (frames staged by TMA)
        return net.generate(T_STEPS, x0_d, uni_d, mel=mel_d, gc_ids=gc, sync=False)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(W):
        step()
    net.sync_check()
    # ---- kernel-only timing of the generation kernels (roofline numerator) ----------------------------------
    torch.cuda.synchronize()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kms = []
    for _ in range(2):
        k0.record()
        step()
        k1.record()
        torch.cuda.synchronize()
        kms.append(k0.elapsed_time(k1))
    kernel_ms = float(np.mean(kms))

    # ---- timed region: K steps ----------------------------------------------------------------------------
    launches0 = net.info()['kernel_launches']
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        out = step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    sampler.stop_flag = True
    net.sync_check()
    launches = net.info()['kernel_launches'] - launches0
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * ROWS * T_STEPS * args.steps / (ms / 1e3)

    # ---- e2e: host buffers through the C-ABI host entry point ------------------------------------------------
    pin = lambda a: torch.from_numpy(a).pin_memory().numpy()
    mel_h, uni_h, x0_h = pin(mel), pin(uniforms), pin(x0)
    net.generate_host(T_STEPS, x0_h, uni_h, mel=mel_h, gc_ids=gc)          # warm (allocations)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        wave = net.generate_host(T_STEPS, x0_h, uni_h, mel=mel_h, gc_ids=gc)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = world * ROWS * T_STEPS * args.steps / e2e_s
    assert np.array_equal(wave, out.cpu().numpy()), "host entry point and device entry point disagree"
    h2d = mel_h.nbytes + uni_h.nbytes + x0_h.nbytes
    d2h = wave.nbytes

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"

    # ---- cells that need every rank (collectives inside) -------------------------------------------------------
    extras = {}
    if not args.no_extras:
        guarded("job64_strong_scaling", lambda: extra_job64(world, rank, dev, kw, w, fast), extras)
        del net
        torch.cuda.empty_cache()
        guarded("cfg5_end_to_end", lambda: extra_cfg5(world, rank, dev), extras)
        guarded("cfg4_training_step", lambda: extra_cfg4(world), extras)
    if rank != 0:
        dist.destroy_process_group()
        return

    # ---- rank-0 cells -------------------------------------------------------------------------------------------
    def rate(rows, steps, fast_act, host=False):
        kwb = dict(kw, batch_size=rows)
        netb = WaveNetModel(train_mode=False, device=dev, fast_act=fast_act, **kwb)
        netb.load_state_dict(w)
        reps = (rows + ROWS - 1) // ROWS
        melb = mel_d.repeat(reps, 1, 1)[:rows, :(steps + 299) // 300].contiguous()
        unib = uni_d[:, :steps].repeat(reps, 1, 1)[:rows].contiguous()
        x0b = x0_d.repeat(reps, 1)[:rows].contiguous()
        gcb = (gc * reps)[:rows]
        if host:
            mh, uh, xh = melb.cpu().numpy(), unib.cpu().numpy(), x0b.cpu().numpy()
            netb.generate_host(steps, xh, uh, mel=mh, gc_ids=gcb)
            t0 = time.perf_counter()
            netb.generate_host(steps, xh, uh, mel=mh, gc_ids=gcb)
            return rows * steps / (time.perf_counter() - t0)
        netb.generate(steps, x0b, unib, mel=melb, gc_ids=gcb)
        b0, b1 = ev_pair()
        b0.record()
        netb.generate(steps, x0b, unib, mel=melb, gc_ids=gcb, sync=False)
        b1.record()
        torch.cuda.synchronize()
        netb.sync_check()
        return rows * steps / (b0.elapsed_time(b1) / 1e3)

    if not args.no_extras:
        guarded("batch1_samples_per_sec", lambda: rate(1, 12000, fast), extras)
        guarded("batch1_e2e_host_samples_per_sec", lambda: rate(1, 12000, fast, host=True), extras)
        guarded("batch8_%s_activation_samples_per_sec" % ("exact" if fast else "fast"), lambda: rate(8, 12000, not fast), extras)
        guarded("batch1_%s_activation_samples_per_sec" % ("exact" if fast else "fast"), lambda: rate(1, 12000, not fast), extras)
        guarded("batch16_samples_per_sec", lambda: rate(16, 6000, fast), extras)
        guarded("batch32_samples_per_sec", lambda: rate(32, 6000, fast), extras)
        guarded("cfg1", lambda: extra_cfg1(dev, peak), extras)
        guarded("cfg3_tacotron", lambda: extra_cfg3(dev), extras)

    # ---- roofline -------------------------------------------------------------------------------------------
    L, R, C = len(kw['dilations']), kw['residual_channels'], kw['local_condition_channels']
    b_step = 4 * info['p_hot'] + ROWS * 4 * (2 * L * R + C + 2)
    achieved = b_step * T_STEPS / (kernel_ms / 1e3) / 1e9
    traffic = None
    for name in ("r02_traffic.json", "r01_traffic.json"):
        tp = os.path.join(ROOT, "profiles", name)
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get("dram_bytes_per_launch")
                break
            except Exception:
                traffic = None
    cluster = info.get("cluster_path", 0) >= 1
    shape = {1: "15 clusters of 8 CTAs", 2: "7 clusters of 16 CTAs + 1 cluster of 8 (two launches)"}.get(info.get("cluster_path", 0), "")
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src,
                "kernel": ("wn_layers_kernel_v2<ShapeCfg2> (%s) + wn_tail_kernel_v2<ShapeCfg2> (16 CTAs), concurrent" % shape if cluster
                           else "wn_persistent_kernel_s<ShapeCfg2>"),
                "kernel_ms": kernel_ms, "algorithmic_bytes_per_step": b_step, "steps_per_launch": T_STEPS,
                "traffic_source": "ncu --replay-mode range over one 8 x 48000 launch (profiles/r02_traffic.json): the concurrent kernels cannot be "
                                  "profiled by kernel replay, which serialises launches",
                "note": "weights (21.4 MB) are resident in registers / shared memory across the grid, so DRAM traffic (53 MB per launch, rings pinned in L2) is far "
                        "below the algorithmic bytes (1.04 TB per launch); the binding limit is the dependent chain per sample: 30 layers x "
                        "(~950 cycles of compute + a 280-cycle DSMEM or 650-cycle L2 hop) + tail (DESIGN.md latency model, profiles/r02_*)",
                "hbm_roofline_samples_per_sec": ROWS * peak * 1e9 / b_step}

    cpu = None
    if not args.no_cpu_baseline and world == 1:      # rank 0 at N = 1 only
        threads = min(os.cpu_count() or 1, ROWS)
        v = cpu_port_throughput(CPU_SAMPLE_STEPS, threads)
        cpu = {"value": v, "unit": "samples/s", "cores": threads, "kind": "port", "cpu_model": cpu_model_name(),
               "host_cores": os.cpu_count(), "sample": "%d rows x %d steps of the same workload, rows over %d host threads (oracle/wn_oracle.c, plain-C "
                         "restatement; the TF 1.x reference cannot be installed)" % (ROWS, CPU_SAMPLE_STEPS, threads),
               "note": "a stated baseline, not the target (rows over threads: every row re-streams the 21 MB of weights; built with "
                       "-ffp-contract=off for bit-comparability); best_effort is the B-cpu organisation of BASELINE.md section 3",
               "best_effort": (lambda r: {"value": r[0], "unit": "samples/s", "cores": r[1], "sample": "%d rows x %d steps, oracle/wn_cpu_best.h: "
                               "weight-stationary threads with packed column slices, all rows together, 2 spin barriers per layer; "
                               "output bit-identical to the port" % (ROWS, CPU_SAMPLE_STEPS)})(cpu_best_effort(CPU_SAMPLE_STEPS)),
               "reference_structure": {"value": cpu_reference_structure_throughput(300), "unit": "samples/s",
                                       "sample": "%d rows x 300 steps, numpy restatement in the per-sample Python loop of "
                                                 "generate.py:202-233 with full queue copies per step (oracle/np_oracle.py)" % ROWS}}

    line = {"metric": "wavenet_generation_samples_per_sec", "value": value, "unit": "samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": W, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "rows_per_gpu": ROWS, "steps_per_row": T_STEPS, "sample_rate": SAMPLE_RATE,
                       "grid": info['grid'], "M": info['M'], "Mt": info['Mt'], "cluster_path": info.get("cluster_path"),
                       "activation": ("ex2.approx/rcp.approx gate: MoL logits within 3e-6 of the pinned arithmetic (tolerance 1e-4), "
                                      "tests/test_gpu_parity.py::test_fast_activation_within_north_star_tolerance" if info.get("fast_act")
                                      else "pinned exp32 + IEEE divide (bit-identical to the CPU oracle)"),
                       "local_condition": "mel frames (8 x 160 x 80) staged by TMA, create_upsample folded into the kernel",
                       "l2": "uniforms (17 MB per step) + rings / mailboxes (~60 MB, rewritten every step) against a 126 MB L2; no explicit flush: "
                             "the kernel is latency-bound on resident weights, not on cached inputs"},
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(launches), "clocks": sampler.summary(), "roofline": roofline, "cpu_baseline": cpu,
            "realtime_factor_per_utterance": value / world / ROWS / SAMPLE_RATE,
            "die_aware_mailboxes": info.get("die_aware"), **extras}
    print(json.dumps(line))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

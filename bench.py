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


def run_reference(args, rank, world):
    """The reference's own CPU path cannot run (TensorFlow 1.x absent): time its CPU
    restatement (oracle port) with all host threads; each step is a bounded sample of the workload."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    threads = min(cores, ROWS)
    sample_steps = 300
    for _ in range(args.warmup):
        cpu_port_throughput(60, threads)
    t0 = time.perf_counter()
    vals = [cpu_port_throughput(sample_steps, threads) for _ in range(args.steps)]
    dt = time.perf_counter() - t0
    # each step's throughput is timed around the generation loop only (as generate.py:199 does); the
    # wall time per step additionally contains building the oracle model and the upsampling
    v = ROWS * sample_steps * args.steps / sum(ROWS * sample_steps / x for x in vals)
    line = {"impl": "reference", "metric": "wavenet_generation_samples_per_sec", "value": v, "unit": "samples/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "rows_per_gpu": ROWS, "steps_per_row": T_STEPS},
            "cpu_baseline": {"value": v, "unit": "samples/s", "cores": threads, "kind": "port", "cpu_model": cpu_model_name(),
                             "host_cores": cores, "sample": "%d rows x %d steps of the workload per bench step, rows over %d host threads "
                                       "(oracle/wn_oracle.c; the TF 1.x reference is not installable)" % (ROWS, sample_steps, threads)},
            "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "per_step_values": vals}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    W = max(args.warmup, 3)

    kw, w, mel, uniforms, x0, gc = make_job(rank)
    net = WaveNetModel(train_mode=False, device=dev, **kw)
    net.load_state_dict(w)
    info = net.info()
    mel_d = torch.from_numpy(mel).to(dev)
    uni_d = torch.from_numpy(uniforms).to(dev)
    x0_d = torch.from_numpy(x0).to(dev)

    def step():
        lc = net.create_upsample(mel_d)
        return net.generate(T_STEPS, x0_d, uni_d, lc_up=lc, gc_ids=gc, sync=False)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(W):
        step()
    net.sync_check()
    # ---- kernel-only timing of the persistent kernel (roofline numerator) -------------------------------
    lc = net.create_upsample(mel_d)
    torch.cuda.synchronize()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kms = []
    for _ in range(2):
        k0.record()
        net.generate(T_STEPS, x0_d, uni_d, lc_up=lc, gc_ids=gc, sync=False)
        k1.record()
        torch.cuda.synchronize()
        kms.append(k0.elapsed_time(k1))
    kernel_ms = float(np.mean(kms))
    del lc

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

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- batch=1 latency figure (north_star asks for batch 1 and batch 8) -----------------------------------
    kw1 = dict(kw, batch_size=1)
    net1 = WaveNetModel(train_mode=False, device=dev, **kw1)
    net1.load_state_dict(w)
    lc1 = net1.create_upsample(mel_d[:1])
    net1.generate(12000, x0_d[:1], uni_d[:1, :12000], lc_up=lc1, gc_ids=gc[:1])
    b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    b0.record()
    net1.generate(12000, x0_d[:1], uni_d[:1, :12000], lc_up=lc1, gc_ids=gc[:1], sync=False)
    b1.record()
    torch.cuda.synchronize()
    batch1 = 12000 / (b0.elapsed_time(b1) / 1e3)

    # ---- more rows in flight (the e2e scenario of BASELINE configs[4] has 32 sentences per GPU) ---------------
    more = {}
    for nb in (16, 32):
        kwb = dict(kw, batch_size=nb)
        netb = WaveNetModel(train_mode=False, device=dev, **kwb)
        netb.load_state_dict(w)
        reps = nb // ROWS
        melb = mel_d.repeat(reps, 1, 1)
        unib = uni_d[:, :6000].repeat(reps, 1, 1)
        lcb = netb.create_upsample(melb)
        gcb = gc * reps
        netb.generate(6000, x0_d.repeat(reps, 1), unib, lc_up=lcb, gc_ids=gcb)
        b0.record()
        netb.generate(6000, x0_d.repeat(reps, 1), unib, lc_up=lcb, gc_ids=gcb, sync=False)
        b1.record()
        torch.cuda.synchronize()
        more["batch%d_samples_per_sec" % nb] = nb * 6000 / (b0.elapsed_time(b1) / 1e3)
        del netb, lcb, melb, unib

    # ---- roofline -------------------------------------------------------------------------------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    L, R, C = len(kw['dilations']), kw['residual_channels'], kw['local_condition_channels']
    b_step = 4 * info['p_hot'] + ROWS * 4 * (2 * L * R + C + 2)
    achieved = b_step * T_STEPS / (kernel_ms / 1e3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "kernel": "wn_persistent_kernel_s<ShapeCfg2>" if info.get("static_shape") == 1 else "wn_persistent_kernel",
                "kernel_ms": kernel_ms, "algorithmic_bytes_per_step": b_step, "steps_per_launch": T_STEPS,
                "note": "weights (21.4 MB) are resident in shared memory across the grid, so DRAM traffic is far below the "
                        "algorithmic bytes; the binding limit is the 34-stage dependent chain per sample (DESIGN.md latency model)",
                "hbm_roofline_samples_per_sec": ROWS * peak * 1e9 / b_step}

    cpu = None
    if not args.no_cpu_baseline and world == 1:      # rank 0 at N = 1 only: the other ranks' hosts are idle at this point
        threads = min(os.cpu_count() or 1, ROWS)
        steps = 1500
        v = cpu_port_throughput(steps, threads)
        cpu = {"value": v, "unit": "samples/s", "cores": threads, "kind": "port", "cpu_model": cpu_model_name(),
               "host_cores": os.cpu_count(), "sample": "%d rows x %d steps of the same workload, rows over %d host threads (oracle/wn_oracle.c, plain-C "
                         "restatement; the TF 1.x reference cannot be installed)" % (ROWS, steps, threads),
               "reference_structure": {"value": cpu_reference_structure_throughput(300), "unit": "samples/s",
                                       "sample": "%d rows x 300 steps, numpy restatement in the per-sample Python loop of "
                                                 "generate.py:202-233 with full queue copies per step (oracle/np_oracle.py)" % ROWS}}

    line = {"metric": "wavenet_generation_samples_per_sec", "value": value, "unit": "samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": W, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "rows_per_gpu": ROWS, "steps_per_row": T_STEPS, "sample_rate": SAMPLE_RATE,
                       "grid": info['grid'], "M": info['M'], "Mt": info['Mt'],
                       "l2": "per-step inputs (upsampled mel 123 MB + uniforms 17 MB) exceed the 126 MB L2; no explicit flush"},
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(launches), "clocks": sampler.summary(), "roofline": roofline, "cpu_baseline": cpu,
            "realtime_factor_per_utterance": value / world / ROWS / SAMPLE_RATE,
            "batch1_samples_per_sec": batch1, "die_aware_mailboxes": info.get("die_aware"), **more}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

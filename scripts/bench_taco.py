"""Tacotron text->mel throughput on one B200 (BASELINE configs[2], "cfg-3": 32 Korean sentences, two speakers,
attention decoder + CBHG, 200 decoder steps = 1000 mel frames = 12.5 s of audio per sentence).

Prints one JSON line: device-resident time of one `taco_synthesize` (CUDA events), the same through the HOST API
(`taco_synthesize_host`: ids in, mel + linear + alignments out), per-kernel-family split (separate timed runs with
linear output off / fewer steps), and the numpy oracle on a bounded sample of the same workload as CPU baseline.
Weights are seeded synthetic (no checkpoint ships with the reference).

  python scripts/bench_taco.py [--sentences 32] [--steps 200] [--iters 5] [--no-cpu]
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

SENTENCES = [
    "오늘은 날씨가 맑고 바람이 선선해서 산책하기에 아주 좋은 날입니다",
    "새로운 음성 합성 시스템은 문장을 자연스러운 목소리로 읽어 줍니다",
    "내일 오전 9시에 회의가 있으니 자료를 미리 준비해 주시기 바랍니다",
    "이 열차는 잠시 후 서울역에 도착합니다. 내리실 문은 왼쪽입니다",
    "도서관에서 빌린 책 3권을 이번 주 금요일까지 반납해야 합니다",
    "커피 한 잔과 따뜻한 빵으로 아침을 시작하면 하루가 즐겁습니다",
    "연구팀은 지난 12개월 동안 수집한 자료를 분석해 결과를 발표했습니다",
    "멀리서 들려오는 파도 소리가 마음을 편안하게 만들어 주었습니다",
]


def make_texts(n):
    tails = ["", " 감사합니다", " 다시 한 번 말씀드립니다", " 잘 들어 주세요"]
    return [SENTENCES[i % len(SENTENCES)] + tails[(i // len(SENTENCES)) % len(tails)] for i in range(n)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--sentences', type=int, default=32)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--iters', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--no-cpu', action='store_true')
    a = ap.parse_args()
    import torch
    from tacotron_wavenet_vocoder_korean_b200 import _taco_lib, synth
    from tacotron_wavenet_vocoder_korean_b200.tacotron import Tacotron
    from tacotron_wavenet_vocoder_korean_b200.text import text_to_sequence, prepare_inputs
    from tests.taco_helpers import Bag
    hp = dict(synth.TACO_HP)
    w = synth.make_taco_weights(hp, 2)
    texts = make_texts(a.sentences)
    ids = prepare_inputs([text_to_sequence(t) for t in texts])
    lens = np.array([int(np.argmax(s == 1)) + 1 for s in ids], np.int32)
    spk = (np.arange(a.sentences) % 2).astype(np.int32)
    N, T_in = ids.shape
    S, r, nm, nf = a.steps, hp['reduction_factor'], hp['num_mels'], hp['num_freq']
    m = Tacotron(Bag(hp))
    m.load_state_dict(w)

    def run(steps, want_linear):
        m.initialize(ids, lens, 2, spk, rnn_decoder_test_mode=True, n_steps=steps, want_linear=want_linear)

    def timed(steps, want_linear):
        for _ in range(a.warmup):
            run(steps, want_linear)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = m.info()['kernel_launches']
        e0.record()
        for _ in range(a.iters):
            run(steps, want_linear)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / a.iters, (m.info()['kernel_launches'] - l0) // a.iters

    ms_full, launches = timed(S, True)
    ms_nolin, _ = timed(S, False)
    ms_enc, _ = timed(1, False)
    # host API: ids in, everything out, copies inside
    L = _taco_lib.lib()
    mel_h = np.empty((N, S * r, nm), np.float32)
    lin_h = np.empty((N, S * r, nf), np.float32)
    al_h = np.empty((N, T_in, S), np.float32)
    ids_c = np.ascontiguousarray(ids, np.int32)
    args = _taco_lib.TacoSynthArgs()
    args.N, args.T_in, args.n_steps = N, T_in, S
    args.ids_dev = ids_c.ctypes.data
    args.lengths = lens.ctypes.data_as(C.POINTER(C.c_int32))
    args.speaker_ids = spk.ctypes.data_as(C.POINTER(C.c_int32))
    args.mel_dev, args.linear_dev, args.alignments_dev = mel_h.ctypes.data, lin_h.ctypes.data, al_h.ctypes.data
    for _ in range(2):
        assert L.taco_synthesize_host(m._h, C.byref(args)) == 0, L.taco_last_error(m._h)
    t0 = time.perf_counter()
    for _ in range(a.iters):
        assert L.taco_synthesize_host(m._h, C.byref(args)) == 0
    ms_host = (time.perf_counter() - t0) / a.iters * 1e3
    frames = N * S * r
    audio_s = frames * 300 / 24000.0
    out = {
        "metric": "Tacotron text->mel sentences/sec (cfg-3: %d sentences, T_in=%d jamo, %d decoder steps, r=%d)" % (N, T_in, S, r),
        "value": N / (ms_full * 1e-3), "unit": "sentences/s", "ms_per_batch": ms_full,
        "mel_frames_per_s": frames / (ms_full * 1e-3), "audio_seconds_per_s": audio_s / (ms_full * 1e-3),
        "split_ms": {"encoder(+1 decoder step)": ms_enc, "decoder_loop": ms_nolin - ms_enc, "post_cbhg+linear": ms_full - ms_nolin},
        "us_per_decoder_step": (ms_nolin - ms_enc) / max(S - 1, 1) * 1e3,
        "e2e_host_api": {"value": N / (ms_host * 1e-3), "unit": "sentences/s", "ms_per_batch": ms_host,
                         "h2d_bytes": int(ids_c.nbytes), "d2h_bytes": int(mel_h.nbytes + lin_h.nbytes + al_h.nbytes)},
        "gpu_launches_per_batch": int(launches), "dtype": "f32", "data": "synthetic weights, own Korean sentences", "info": m.info(),
    }
    # per-op CUDA-event timings of one call (TACO_TIME_OPS=1 makes taco_synthesize bracket every op of its launch list)
    os.environ['TACO_TIME_OPS'] = '1'
    run(S, True)
    torch.cuda.synchronize()
    del os.environ['TACO_TIME_OPS']
    try:
        ops = m.debug_tensor('op_times', (-1, 6))
    except Exception:
        ops = None
    if ops is not None and len(ops):
        ops = [[float(x) for x in o] for o in ops]
        kinds = {0: 'other', 1: 'fp32 SIMT GEMM', 2: 'tcgen05 3xTF32 GEMM (+ operand split)', 3: 'bidirectional GRU', 4: 'decoder loop'}
        by_kind = {}
        for ms, kind, fl, M, Nn, K in ops:
            by_kind.setdefault(kinds[int(kind)], [0.0, 0.0, 0])
            by_kind[kinds[int(kind)]][0] += float(ms); by_kind[kinds[int(kind)]][1] += float(fl); by_kind[kinds[int(kind)]][2] += 1
        out["ops_ms"] = {k: {"ms": v[0], "launch_groups": v[2], "useful_gflop": v[1] / 1e9} for k, v in by_kind.items()}
        tc = [o for o in ops if int(o[1]) == 2]
        if tc:
            peaks = {}
            try:
                peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')))
            except Exception:
                pass
            bf16 = float(peaks.get('bf16_tflops', 0) or 0)
            peak_tf32, src = (bf16 / 2, "MEASURED_PEAKS.json bf16_tflops / 2 (kind::tf32 runs at half the bf16 rate)") if bf16 else \
                (1125.0, "fallback: nominal dense 2250 TFLOP/s bf16 / 2 (B200_PROFILING.md)")
            big = max(tc, key=lambda o: o[2])
            tot_ms, tot_fl = sum(o[0] for o in tc), sum(o[2] for o in tc)
            out["roofline"] = {
                "bound": "tensor", "unit": "TFLOP/s", "peak": peak_tf32, "peak_source": src,
                "kernel": "tc::gemm_tc_kernel (largest launch: M=%d N=%d sum K=%d, operand split included in its time)" % (int(big[3]), int(big[4]), int(big[5])),
                "achieved": 3 * big[2] / (big[0] * 1e-3) / 1e12, "frac": 3 * big[2] / (big[0] * 1e-3) / 1e12 / peak_tf32,
                "useful_fp32_tflops": big[2] / (big[0] * 1e-3) / 1e12, "kernel_ms": float(big[0]),
                "all_tc_launches": {"ms": float(tot_ms), "useful_gflop": float(tot_fl / 1e9), "executed_tf32_tflops": float(3 * tot_fl / (tot_ms * 1e-3) / 1e12)},
                "note": "achieved counts the three TF32 products per fp32 multiply-add that the split executes; the path as a whole is bound by the decoder loop "
                        "(13 grid barriers per step), see ops_ms"}
    if not a.no_cpu:
        from oracle.taco_oracle import TacotronOracle
        n_cpu, s_cpu = min(N, 8), min(S, 40)
        o = TacotronOracle(hp, w, 2)
        t0 = time.perf_counter()
        o.synthesize(ids[:n_cpu], lens[:n_cpu], spk[:n_cpu], max_iters=s_cpu)
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": n_cpu * s_cpu * r / dt, "unit": "mel frames/s", "kind": "port",
                               "cores": os.cpu_count(), "sample": "%d sentences x %d decoder steps, numpy (BLAS threads = host default)" % (n_cpu, s_cpu),
                               "seconds": dt}
    print(json.dumps(out))


if __name__ == "__main__":
    main()

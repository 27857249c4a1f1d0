# coding: utf-8
"""End-to-end text -> waveform job (BASELINE configs[4]): the reference chains
`synthesizer.py` (text -> mel .npy, synthesizer.py:279-280) into `generate.py --mel` (mel -> 24 kHz wav,
generate.py:151-262) through files; here the mel stays in HBM between the two models.

Sharding: sentences are independent, so a job splits by sentence -- one process per GPU, LPT on the token count,
no collective on the data path; torch.distributed (NCCL on the GPU box, gloo in the CPU tests) only broadcasts
the job description and gathers the padded waveforms.
"""
import numpy as np

from . import dist as wdist


class TextToSpeech(object):
    """Tacotron (libtaco_b200.so) + WaveNet vocoder (libwn_b200.so) on one GPU."""

    def __init__(self, taco_hparams, taco_weights, num_speakers, wavenet_kwargs, wavenet_weights, hop_size=300):
        from .tacotron import Tacotron
        from .wavenet import WaveNetModel
        self.hp = taco_hparams
        self.num_speakers = num_speakers
        self.taco = Tacotron(taco_hparams)
        self.taco.load_state_dict(taco_weights)
        self.wn = WaveNetModel(train_mode=False, **wavenet_kwargs)
        self.wn.load_state_dict(wavenet_weights)
        self.hop = hop_size
        up = int(np.prod(wavenet_kwargs['upsample_factor']))
        if up != hop_size:
            raise ValueError("prod(upsample_factor) = %d must equal hop_size = %d (generate.py:152)" % (up, hop_size))

    def _get(self, k):
        return self.hp[k] if isinstance(self.hp, dict) else getattr(self.hp, k)

    def text_to_mel(self, texts, speaker_ids, attention_trim=True, max_mel_frames=None):
        """-> list of (frames_i, num_mels) CUDA tensors (synthesizer.py:72-200 without the files)."""
        import torch
        from .synthesizer import attention_trim_index
        from .text import text_to_sequence, prepare_inputs
        seqs = prepare_inputs([text_to_sequence(t) for t in texts])
        lens = [int(np.argmax(s == 1)) + 1 for s in seqs]
        self.taco.initialize(seqs, lens, self.num_speakers, speaker_ids, rnn_decoder_test_mode=True, want_linear=False)
        mel = self.taco.mel_outputs
        frames = [mel.shape[1]] * len(texts)
        if attention_trim:
            al = self.taco.alignments.cpu().numpy()
            frames = [min(mel.shape[1], attention_trim_index(al[i], len(seqs[i]), self._get('reduction_factor'))) for i in range(len(texts))]
        if max_mel_frames:
            frames = [min(f, max_mel_frames) for f in frames]
        return [mel[i, :frames[i]] for i in range(len(texts))]

    def mel_to_wav(self, mels, speaker_ids, seed=0):
        """generate.py:151-256 for a list of mels of different lengths: groups of <= batch_size rows, one persistent
        kernel launch per group, per-row step counts.  -> list of 1-D float32 numpy waveforms."""
        import torch
        rs = np.random.RandomState(seed)
        order = wdist.make_groups(list(range(len(mels))), [int(m.shape[0]) for m in mels], self.wn.batch_size)
        out = [None] * len(mels)
        nr1 = self.wn.out_channels // 3 + 1
        for grp in order:
            rows = len(grp)
            fmax = max(int(mels[i].shape[0]) for i in grp)
            T = fmax * self.hop
            lc = torch.zeros((rows, fmax, mels[grp[0]].shape[1]), dtype=torch.float32, device=mels[grp[0]].device)
            for r, i in enumerate(grp):
                lc[r, :mels[i].shape[0]] = mels[i]
            if self.wn.scalar_input:
                x0 = (2 * rs.rand(rows, 1) - 1).astype(np.float32)                               # generate.py:186-188
                uni = torch.empty((rows, T, nr1), dtype=torch.float32, device=lc.device).uniform_(1e-5, 1 - 1e-5)
            else:
                x0 = rs.randint(self.wn.quantization_channels, size=(rows, 1)).astype(np.float32)  # generate.py:190-192
                uni = torch.rand((rows, T), dtype=torch.float64, device=lc.device)
            gc = [int(speaker_ids[i]) for i in grp] if self.wn.global_condition_channels else None
            wav = self.wn.generate(T, x0, uni, mel=lc, lc_shift=0, gc_ids=gc,
                                   T_row=[int(mels[i].shape[0]) * self.hop for i in grp])
            wav = wav.cpu().numpy()
            for r, i in enumerate(grp):
                out[i] = wav[r, :int(mels[i].shape[0]) * self.hop].copy()
        return out

    def synthesize(self, texts, speaker_ids, attention_trim=True, max_mel_frames=None, seed=0, taco_batch=32):
        mels = []
        for k in range(0, len(texts), taco_batch):
            mels += self.text_to_mel(texts[k:k + taco_batch], speaker_ids[k:k + taco_batch], attention_trim, max_mel_frames)
        return self.mel_to_wav(mels, speaker_ids, seed=seed)


def shard_sentences(texts, world):
    """LPT on the character count (a proxy for the synthesis cost before the decoder has run)."""
    return wdist.lpt_assign([len(t) for t in texts], world)


def gather_varlen(waves, assign, dst=0, device=None):
    """Every rank contributes its list of 1-D arrays (job order given by assign[rank]); rank `dst` returns the
    full list in job order.  One all_gather_object of the lengths and one padded gather."""
    import torch
    import torch.distributed as dist
    device = wdist._backend_device(device)
    world, rank = dist.get_world_size(), dist.get_rank()
    lens = [None] * world
    dist.all_gather_object(lens, [int(len(w)) for w in waves])
    n_max = max(max((len(l) for l in lens), default=0), 1)
    t_max = max(max((max(l) if l else 0) for l in lens), 1)
    buf = np.zeros((n_max, t_max), np.float32)
    for j, w in enumerate(waves):
        buf[j, :len(w)] = w
    send = torch.from_numpy(buf).to(device)
    bins = [torch.zeros((n_max, t_max), dtype=torch.float32, device=device) for _ in range(world)] if rank == dst else None
    dist.gather(send, bins, dst=dst)
    if rank != dst:
        return None
    total = sum(len(a) for a in assign)
    out = [None] * total
    for r, a in enumerate(assign):
        got = bins[r].cpu().numpy()
        for j, i in enumerate(a):
            out[i] = got[j, :lens[r][j]].copy()
    return out


def run_job(make_tts, texts, speaker_ids, src=0, device=None, **synth_kwargs):
    """Whole multi-GPU job.  `make_tts()` builds this rank's TextToSpeech (weights are seeded or loaded per rank);
    `texts` / `speaker_ids` need to be valid on rank `src` only.  Returns the waveforms on `src`."""
    import torch.distributed as dist
    job = [None]
    if dist.get_rank() == src:
        job[0] = (list(texts), [int(s) for s in speaker_ids])
    dist.broadcast_object_list(job, src=src)
    texts, speaker_ids = job[0]
    assign = shard_sentences(texts, dist.get_world_size())
    mine = assign[dist.get_rank()]
    tts = make_tts()
    waves = tts.synthesize([texts[i] for i in mine], [speaker_ids[i] for i in mine], **synth_kwargs) if mine else []
    return gather_varlen(waves, assign, dst=src, device=device)

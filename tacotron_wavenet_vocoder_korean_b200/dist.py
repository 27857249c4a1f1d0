"""Multi-GPU driver for WaveNet generation: utterances are independent (every op of
wavenet/model.py:112-167 is row-wise), so the job shards by utterance -- one process per GPU,
no collective on the data path.  torch.distributed (NCCL over NVLink on the GPU box, gloo in the
CPU tests) is used exactly three times per job: broadcast the weights, scatter the mel inputs,
gather the waveforms.  A single utterance never spans GPUs: its 64-stage
serial chain would pay an NVLink hop per stage ("replicas only" inside an utterance).
"""
import numpy as np
import torch
import torch.distributed as dist


def lpt_assign(lengths, world):
    """Longest-processing-time-first: sort utterances by length (descending) and deal each to the
    least-loaded rank.  Returns a list (per rank) of utterance indices."""
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    load = [0] * world
    out = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += int(lengths[i])
    return out


def make_groups(indices, lengths, batch):
    """Split one rank's utterances into groups of <= batch rows, longest first so a group's rows
    have similar lengths (the kernel runs max(T_row) steps for the group)."""
    idx = sorted(indices, key=lambda i: (-int(lengths[i]), i))
    return [idx[k:k + batch] for k in range(0, len(idx), batch)]


def _backend_device(device=None):
    if device is not None:
        return torch.device(device)
    if dist.is_initialized() and dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def broadcast_state(state, src=0, device=None):
    """Broadcast {tf_variable_name: array} from `src` as ONE flat fp32 tensor (<= 22 MB)."""
    device = _backend_device(device)
    meta = [None]
    if dist.get_rank() == src:
        meta[0] = [(k, tuple(np.shape(v))) for k, v in state.items()]
    dist.broadcast_object_list(meta, src=src)
    meta = meta[0]
    total = int(sum(int(np.prod(s)) for _, s in meta))
    flat = torch.empty(total, dtype=torch.float32, device=device)
    if dist.get_rank() == src:
        flat.copy_(torch.from_numpy(np.concatenate([np.asarray(state[k], np.float32).ravel() for k, _ in meta])))
    dist.broadcast(flat, src=src)
    flat = flat.cpu().numpy()
    out, o = {}, 0
    for k, s in meta:
        n = int(np.prod(s))
        out[k] = flat[o:o + n].reshape(s).copy()
        o += n
    return out


def scatter_utterances(mels, gc_ids, src=0, device=None):
    """Rank `src` holds the job: list of (T_mel_i, C) mel arrays + speaker ids.  Every rank receives
    its LPT share as (indices, [mel], [gc_id]); mels travel as one padded (n_r, T_max, C) tensor."""
    device = _backend_device(device)
    world, rank = dist.get_world_size(), dist.get_rank()
    meta = [None]
    if rank == src:
        lengths = [int(m.shape[0]) for m in mels]
        assign = lpt_assign(lengths, world)
        C = int(mels[0].shape[1]) if mels else 0
        n_max = max((len(a) for a in assign), default=0)
        t_max = max(lengths, default=0)
        meta[0] = dict(assign=assign, lengths=lengths, C=C, n_max=n_max, t_max=t_max,
                       gc=[int(g) for g in gc_ids] if gc_ids is not None else None)
    dist.broadcast_object_list(meta, src=src)
    m = meta[0]
    shape = (max(m['n_max'], 1), max(m['t_max'], 1), max(m['C'], 1))
    recv = torch.zeros(shape, dtype=torch.float32, device=device)
    chunks = None
    if rank == src:
        chunks = []
        for a in m['assign']:
            buf = np.zeros(shape, np.float32)
            for j, i in enumerate(a):
                buf[j, :m['lengths'][i], :] = mels[i]
            chunks.append(torch.from_numpy(buf).to(device))
    dist.scatter(recv, chunks, src=src)
    mine = m['assign'][rank]
    recv = recv.cpu().numpy()
    my_mels = [recv[j, :m['lengths'][i], :].copy() for j, i in enumerate(mine)]
    my_gc = [m['gc'][i] for i in mine] if m['gc'] is not None else None
    return mine, my_mels, my_gc, m


def gather_waveforms(indices, waves, meta, hop, dst=0, device=None):
    """Inverse of scatter_utterances: every rank contributes its waveforms (list of 1-D arrays);
    rank `dst` returns the full list in job order, the others None."""
    device = _backend_device(device)
    world, rank = dist.get_world_size(), dist.get_rank()
    shape = (max(meta['n_max'], 1), max(meta['t_max'], 1) * hop)
    buf = np.zeros(shape, np.float32)
    for j, w in enumerate(waves):
        buf[j, :len(w)] = w
    send = torch.from_numpy(buf).to(device)
    bins = [torch.zeros(shape, dtype=torch.float32, device=device) for _ in range(world)] if rank == dst else None
    dist.gather(send, bins, dst=dst)
    if rank != dst:
        return None
    out = [None] * len(meta['lengths'])
    for r, a in enumerate(meta['assign']):
        got = bins[r].cpu().numpy()
        for j, i in enumerate(a):
            out[i] = got[j, :meta['lengths'][i] * hop].copy()
    return out


def generate_job(generate_group, state, mels, gc_ids, batch, hop, src=0, device=None):
    """Whole job: broadcast weights, scatter inputs, run `generate_group(state, [mel], [gc], [job_index])
    -> [waveform]` on groups of <= batch utterances, gather.  `generate_group` is the only device work."""
    state = broadcast_state(state if dist.get_rank() == src else None, src=src, device=device)
    mine, my_mels, my_gc, meta = scatter_utterances(mels, gc_ids, src=src, device=device)
    lengths = [m.shape[0] for m in my_mels]
    waves = [None] * len(mine)
    for grp in make_groups(list(range(len(mine))), lengths, batch):
        res = generate_group(state, [my_mels[k] for k in grp], [my_gc[k] for k in grp] if my_gc is not None else None,
                             [mine[k] for k in grp])
        for k, w in zip(grp, res):
            waves[k] = np.asarray(w, np.float32)
    return gather_waveforms(mine, waves, meta, hop, dst=src, device=device)


# ---- data-parallel training ------------------------------------------------------------
def allreduce_mean_(flat_grads, loss=None):
    """Data-parallel gradient exchange of the training step: ONE sum all-reduce of the flat fp32 gradient buffer (NCCL over
    NVLink on the GPU box, gloo in the CPU tests); returns (grad_scale, mean loss).  The caller folds grad_scale = 1/world
    into the Adam kernel instead of a separate pass over the buffer.  No-op without an initialised process group."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return 1.0, loss
    world = dist.get_world_size()
    dist.all_reduce(flat_grads)
    if loss is not None:
        loss = loss.clone()
        dist.all_reduce(loss)
        loss /= world
    return 1.0 / world, loss


def broadcast_flat_(flat_params, src=0):
    """Every rank starts from rank `src`'s parameters (tf.train.Saver.restore on one worker + broadcast)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(flat_params, src=src)
    return flat_params

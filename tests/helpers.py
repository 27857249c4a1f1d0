"""Shared input builders for the parity tests: the same seeded inputs go to the CPU oracle
(oracle/) and to the CUDA path (through the C ABI)."""
import numpy as np

import oracle
from tacotron_wavenet_vocoder_korean_b200 import synth


def make_inputs(kw, T, t_mel=None, seed=1):
    """Seeded inputs after SURVEY.md section 8(d): uniforms RandomState(seed), initial sample RandomState(0),
    mel clip(N(0,1.5),-4,4) RandomState(2)."""
    N = kw['batch_size']
    rs = np.random.RandomState(seed)
    inp = {}
    if kw['scalar_input']:
        nr1 = kw['out_channels'] // 3 + 1
        inp['uniforms'] = rs.uniform(1e-5, 1 - 1e-5, (N, T, nr1)).astype(np.float32)
        inp['x0'] = (2 * np.random.RandomState(0).rand(N, 1) - 1).astype(np.float32)       # generate.py:188
        inp['forced_full'] = (2 * np.random.RandomState(3).rand(N, T) - 1).astype(np.float32) * 0.5
    else:
        inp['uniforms'] = rs.random_sample((N, T))
        Q = kw['quantization_channels']
        inp['x0'] = np.random.RandomState(0).randint(Q, size=(N, 1)).astype(np.float32)      # generate.py:192
        inp['forced_full'] = np.random.RandomState(3).randint(Q, size=(N, T)).astype(np.float32)
    if kw.get('local_condition_channels'):
        hop = int(np.prod(kw['upsample_factor']))
        t_mel = t_mel or (T + hop - 1) // hop
        inp['mel'] = np.clip(np.random.RandomState(2).randn(N, t_mel, kw['local_condition_channels']) * 1.5, -4, 4).astype(np.float32)
    if kw.get('global_condition_channels'):
        card = kw['global_condition_cardinality']
        inp['gc_ids'] = np.array([(i * 2 // max(N, 1)) % card if N > 1 else 0 for i in range(N)], np.int32)
    else:
        inp['gc_ids'] = None
    return inp


def oracle_model(kw, weights):
    om = oracle.OracleModel(**kw)
    om.set_weights(weights)
    return om


def plan_from_dict(d):
    return oracle.OrcPlan.from_dict(d)

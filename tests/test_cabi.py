"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol the header
declares, validates arguments, plans on the host, and FAILS LOUDLY without a GPU (no fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from tacotron_wavenet_vocoder_korean_b200 import _lib, synth
from tacotron_wavenet_vocoder_korean_b200.wavenet import WaveNetModel
from tacotron_wavenet_vocoder_korean_b200.wavenet.model import plan_config, _make_cfg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, 'include', 'wn_b200.h')).read()
    return sorted(set(re.findall(r'^(?:int|void|const char \*)\s*\*?(wn_[a-z_0-9]+)\s*\(', src, flags=re.M)))


def test_library_exports_every_declared_symbol():
    names = header_functions()
    assert len(names) >= 14, names
    lib = _lib.lib()
    for n in names:
        assert hasattr(lib, n), 'libwn_b200.so does not export %s' % n
    assert sorted(_lib.EXPORTS) == names


def test_config_struct_layout():
    assert C.sizeof(_lib.WnConfig) == 4 * (15 + _lib.WN_MAX_UPSAMPLE + _lib.WN_MAX_LAYERS + 3)
    assert C.sizeof(_lib.WnPlan) == 44


def test_create_validates_arguments():
    lib = _lib.lib()
    bad = [dict(filter_width=3), dict(batch_size=0), dict(batch_size=33), dict(quantization_channels=100, scalar_input=False),
           dict(out_channels=31, scalar_input=True), dict(residual_channels=512), dict(dilations=[1, 0, 2])]
    for override in bad:
        kw = synth.tiny_mulaw() if 'quantization_channels' in override else synth.tiny_mol()
        kw.update(override)
        cfg = _make_cfg(**kw)
        h = C.c_void_p()
        assert lib.wn_create(C.byref(cfg), C.byref(h)) == -1, override
        assert not h.value
        assert len(lib.wn_last_error(None)) > 0
        with pytest.raises(ValueError):
            WaveNetModel(train_mode=False, **kw)


def test_receptive_field_through_abi():
    assert WaveNetModel.calculate_receptive_field(2, [1, 2, 4, 8, 16, 32, 64, 128, 256, 512] * 5, False, 32) == 5117
    assert WaveNetModel.calculate_receptive_field(2, [1, 2, 4, 8, 16, 32, 64, 128, 256, 512] * 5, True, 32) == 5147


def test_host_side_planning():
    plan, info = plan_config(**synth.cfg2())
    assert (info['grid'], info['M'], info['Mt'], info['threads']) == (137, 4, 16, 256)
    assert info['p_hot'] == 5347102 and info['weights_in_global'] == 0
    assert info['smem_bytes_layer'] <= 227 * 1024
    assert plan['M'] == 4 and plan['Mt'] == 16 and info['static_shape'] == 1
    plan, info = plan_config(**synth.cfg1())
    assert info['p_hot'] == 615168 and info['grid'] == 27 and info['static_shape'] == 2
    plan, info = plan_config(**synth.cfg_hparams_default())
    assert info['p_hot'] == 1640670 and info['M'] == 1 and info['grid'] == 67 and info['static_shape'] == 3
    _, info = plan_config(generic_kernel=True, **synth.cfg2())
    assert info['static_shape'] == 0
    _, info = plan_config(**synth.tiny_mol())
    assert info['static_shape'] == 0                    # unknown shapes run the runtime-shaped kernel
    # fewer SMs: the split shrinks instead of failing; too few SMs is an error
    _, small = plan_config(sm_count=80, **synth.cfg2())
    assert small['grid'] <= 80
    with pytest.raises(ValueError):
        plan_config(sm_count=20, **synth.cfg2())
    # a model whose layer slice overflows shared memory spills to global memory instead of failing
    big = synth.cfg2()
    big.update(residual_channels=256, dilation_channels=256, skip_channels=1024)
    _, inf = plan_config(**big)
    assert inf['weights_in_global'] > 0 and inf['smem_bytes_layer'] <= 227 * 1024


def test_packing_is_a_permutation():
    # evaluation-plan invariants the kernel relies on
    for fac in (synth.cfg1, synth.cfg2, synth.cfg_hparams_default, synth.tiny_mol, synth.tiny_mulaw):
        kw = fac()
        plan, _ = plan_config(**kw)
        D, R, S = kw['dilation_channels'], kw['residual_channels'], kw['skip_channels']
        assert D % plan['M'] == 0 and S % plan['Mt'] == 0
        assert R % plan['t_cur'] == 0 and plan['t_cur'] == plan['t_old'] and plan['t_cur'] <= 64
        assert (D // plan['M']) % plan['t_dense'] == 0 and D % plan['t_skip'] == 0
        assert S % plan['t_post1'] == 0 and (S // plan['Mt']) % plan['t_post2'] == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-GPU failure mode')
def test_no_gpu_means_loud_failure_not_fallback():
    kw = synth.tiny_mol()
    net = WaveNetModel(train_mode=False, **kw)
    with pytest.raises(RuntimeError):
        net.load_state_dict(synth.make_weights(**kw))
    with pytest.raises(RuntimeError):
        net.generate(4, np.zeros((2, 1), np.float32), np.full((2, 4, 11), 0.5, np.float32))
    with pytest.raises(RuntimeError):
        net.create_upsample(np.zeros((2, 3, 20), np.float32))
    lib = _lib.lib()
    cfg = _make_cfg(**kw)
    h = C.c_void_p()
    assert lib.wn_create(C.byref(cfg), C.byref(h)) == 0
    a = np.zeros(8, np.float32)
    assert lib.wn_set_weight(h, b'wavenet/conv1d/kernel', a.ctypes.data_as(C.c_void_p), a.size) == 0
    assert lib.wn_finalize(h) == -3                       # WN_ERR_CUDA
    assert b'cuda' in lib.wn_last_error(h).lower()
    args = _lib.WnGenerateArgs()
    assert lib.wn_generate(h, C.byref(args), None) == -2  # WN_ERR_STATE: not finalized
    lib.wn_destroy(h)


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, 'tacotron_wavenet_vocoder_korean_b200')
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle', src, flags=re.M), f
                assert 'liborc' not in src and 'wn_oracle' not in src, f
    for f in ('generate.py',):
        p = os.path.join(ROOT, f)
        if os.path.exists(p):
            assert not re.search(r'^\s*(from|import)\s+oracle', open(p).read(), flags=re.M)


def test_training_entry_points_need_train_mode_and_a_loss_first():
    net = WaveNetModel(train_mode=False, **synth.tiny_mol())
    with pytest.raises(RuntimeError):
        net.add_loss(None)                      # generation-mode model (wavenet/model.py:26 train_mode)
    with pytest.raises(RuntimeError):
        net.add_optimizer(None, None)           # "Supposes that initialize function has already been called" (:315)


def test_integration_md_stub_matches_the_shipped_binding():
    """The ctypes structures a maintainer of the reference is told to add must have the layout of the
    binding this package ships (and therefore of include/wn_b200.h)."""
    import ctypes as C
    import re
    from tacotron_wavenet_vocoder_korean_b200 import _lib
    text = open(os.path.join(ROOT, 'INTEGRATION.md'), encoding='utf-8').read()
    block = re.search(r"```python\n# wavenet/b200.py(.*?)```", text, re.S).group(1)
    classes = block[block.index('class WnConfig'):block.index('def b200_generate')]
    ns = {'C': C}
    exec(classes, ns)
    for name in ('WnConfig', 'WnGenerateArgs'):
        doc, real = ns[name], getattr(_lib, name)
        assert C.sizeof(doc) == C.sizeof(real), name
        assert [(n, getattr(doc, n).offset, getattr(doc, n).size) for n, _ in doc._fields_] == \
               [(n, getattr(real, n).offset, getattr(real, n).size) for n, _ in real._fields_], name
    for fn in re.findall(r"_lib\.(wn_\w+)\(", block):
        assert fn in _lib.EXPORTS, fn

# coding: utf-8
"""Generates tests/golden/ref_utils.json by running THE REFERENCE'S OWN utils/__init__.py helpers that sit on the drop-in boundary:
`load_json` (euc-kr, tolerant of trailing commas) and `load_hparams` (params.json overrides onto the hparams singleton, unknown keys
skipped).

    python tests/golden/make_reference_utils_golden.py        (build container only)
"""
import contextlib
import importlib.util
import io
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get('REFERENCE_ROOT', '/root/reference')
sys.path.insert(0, HERE)
import tf_numpy_shim as tf      # noqa: E402

tf.install()

PARAMS_TEXT = '{\n "sample_rate": 16000,\n "num_mels": 40,\n "dilations": [1, 2, 4,],\n "name": "한국어 모델",\n "not_a_hparam": 3,\n "upsample_factor": [4, 4, 10],\n}\n'


class HP(object):
    def __init__(self, **kw):
        self.__dict__.update(kw)


def main():
    spec = importlib.util.spec_from_file_location('ref_utils', os.path.join(REF, 'utils', '__init__.py'))
    ru = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ru)
    out = {'params_text': PARAMS_TEXT}
    with tempfile.TemporaryDirectory() as d:
        with open(os.path.join(d, 'params.json'), 'w', encoding='euc-kr') as f:
            f.write(PARAMS_TEXT)
        out['load_json'] = ru.load_json(os.path.join(d, 'params.json'))
        hp = HP(sample_rate=24000, num_mels=80, dilations=[1, 2], name='x', upsample_factor=[5, 5, 12], hop_size=300)
        with contextlib.redirect_stdout(io.StringIO()):
            ru.load_hparams(hp, d)
        out['load_hparams'] = dict(vars(hp))
        # utils.get_most_recent_checkpoint itself cannot run: it uses glob without importing it
    json.dump(out, open(os.path.join(HERE, 'ref_utils.json'), 'w', encoding='utf-8'), ensure_ascii=False, indent=1)
    print(out)


if __name__ == '__main__':
    main()

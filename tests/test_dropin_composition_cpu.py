"""`cutie/` (this repo's drop-in shim) composed with a reference checkout on sys.path: modules the shim provides -- the
hot-path surface -- win; everything else (dataset readers, palette, ...) resolves to the reference's own files, so
eval_vos.py-style imports work unchanged with this repo placed AHEAD of the reference on PYTHONPATH.  Needs the
read-only reference checkout of the build container (skipped elsewhere)."""
import os
import subprocess
import sys

import pytest

from tests.conftest import ROOT

REF = '/root/reference'

PROBE = r'''
import importlib, json
out = {}
for mod in ['cutie.inference.inference_core', 'cutie.inference.memory_manager', 'cutie.inference.kv_memory_store',
            'cutie.inference.object_manager', 'cutie.model.cutie', 'cutie.utils.get_default_model',
            'cutie.inference.data.video_reader', 'cutie.inference.data.vos_test_dataset', 'cutie.utils.palette']:
    out[mod] = importlib.import_module(mod).__file__
from cutie.inference.inference_core import InferenceCore
out['InferenceCore'] = InferenceCore.__module__
print(json.dumps(out))
'''


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'cutie')), reason='reference checkout not present')
def test_shim_wins_for_the_hot_path_and_defers_to_the_reference_elsewhere(tmp_path):
    import json
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([ROOT, REF]))
    r = subprocess.run([sys.executable, '-c', PROBE], capture_output=True, text=True, cwd=str(tmp_path), env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    got = json.loads(r.stdout.strip().splitlines()[-1])
    ours = [m for m, f in got.items() if f.startswith(ROOT + os.sep)]
    theirs = [m for m, f in got.items() if f.startswith(REF + os.sep)]
    assert set(ours) == {'cutie.inference.inference_core', 'cutie.inference.memory_manager', 'cutie.inference.kv_memory_store',
                         'cutie.inference.object_manager', 'cutie.model.cutie', 'cutie.utils.get_default_model'}
    assert set(theirs) == {'cutie.inference.data.video_reader', 'cutie.inference.data.vos_test_dataset', 'cutie.utils.palette'}
    assert got['InferenceCore'] == 'cutie_b200.inference.inference_core'


def test_shim_alone_still_imports():
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, '-c', 'from cutie.inference.inference_core import InferenceCore; '
                        'from cutie.utils.get_default_model import get_default_model; print(InferenceCore.__module__)'],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0 and 'cutie_b200.inference.inference_core' in r.stdout, r.stderr[-2000:]

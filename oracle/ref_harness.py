"""Loads the UNMODIFIED reference (hkchengrex/Cutie, read-only at /root/reference) on CPU.

TEST / BASELINE INFRASTRUCTURE: used by tests/golden/make_golden.py to emit the committed fixtures that pin oracle/
and the CUDA path, by tests that re-validate the oracle against the live reference, by tests/ref_runner.py (the
reference run in eager fp32 on the GPU box, from baseline/_ref/) and by bench.py --impl reference.

Accommodations:
  * `omegaconf` is not installed: a stand-in module with DictConfig/OmegaConf/open_dict is injected
    into sys.modules (the reference uses DictConfig for annotations and duck-typed access only).
  * `resnet18/50(pretrained=True)` would hit the network (cutie/model/utils/resnet.py:168-179):
    wrapped to pretrained=False.
  * cfg = the reference's own YAML (cutie/config/eval_config.yaml + model/base.yaml) with the three
    ${...} interpolations resolved by hand.
"""
import contextlib
import os
import sys
import types

import yaml

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_root() -> str:
    """CUTIE_REFERENCE_ROOT, else the read-only mount of the build container, else the install under baseline/_ref/
    (git-ignored, shipped to the GPU box by gpurun; baseline/install_reference.py)."""
    env = os.environ.get('CUTIE_REFERENCE_ROOT')
    if env:
        return env
    for cand in ('/root/reference', os.path.join(os.path.dirname(_HERE), 'baseline', '_ref')):
        if os.path.isdir(os.path.join(cand, 'cutie', 'inference')):
            return cand
    return '/root/reference'


REF_ROOT = _find_root()


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, 'cutie'))


class DictConfig(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v


def to_cfg(o):
    if isinstance(o, dict):
        return DictConfig({k: to_cfg(v) for k, v in o.items()})
    return o


def _install_omegaconf_standin():
    if 'omegaconf' in sys.modules:
        return
    m = types.ModuleType('omegaconf')
    m.DictConfig = DictConfig

    class OmegaConf:
        create = staticmethod(to_cfg)

    m.OmegaConf = OmegaConf
    m.open_dict = lambda cfg: contextlib.nullcontext()
    sys.modules['omegaconf'] = m


def reference_cfg(**overrides) -> DictConfig:
    cfg_dir = os.path.join(REF_ROOT, 'cutie', 'config')
    with open(os.path.join(cfg_dir, 'model', 'base.yaml')) as f:
        model = yaml.safe_load(f)
    model['object_transformer']['embed_dim'] = model['embed_dim']
    model['object_summarizer']['embed_dim'] = model['embed_dim']
    model['object_summarizer']['num_summaries'] = model['object_transformer']['num_queries']
    with open(os.path.join(cfg_dir, 'eval_config.yaml')) as f:
        ev = yaml.safe_load(f)
    for k in ('defaults', 'hydra', 'datasets'):
        ev.pop(k, None)
    ev['model'] = model
    ev.update(mem_every=5, use_long_term=False)  # what get_dataset_cfg escalates for d17-val
    for k, v in overrides.items():
        if k == 'long_term':
            ev['long_term'].update(v)
        else:
            ev[k] = v
    return to_cfg(ev)


_loaded = None


def load_reference():
    """Returns a namespace with the reference's modules (imported from REF_ROOT, never copied)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f'reference tree not found at {REF_ROOT}')
    _install_omegaconf_standin()
    for name in list(sys.modules):
        if name == 'cutie' or name.startswith('cutie.'):
            raise RuntimeError('a different `cutie` package is already imported; load the reference in a '
                               'fresh process (tests run it through a subprocess)')
    # this repo ships a `cutie/` drop-in shim (a regular package, which would shadow -- and, composed with the reference,
    # override -- the reference's namespace package): keep every path entry that holds it out of sight while the
    # reference's modules are imported; they stay cached in sys.modules afterwards
    saved_path = list(sys.path)
    sys.path[:] = [REF_ROOT] + [q for q in sys.path
                                if not os.path.isfile(os.path.join(q or os.getcwd(), 'cutie', '__init__.py'))]
    try:
        import cutie.model.utils.resnet as R
        r18, r50 = R.resnet18, R.resnet50
        R.resnet18 = lambda pretrained=True, extra_dim=0, model_dir=None: r18(False, extra_dim, model_dir)
        R.resnet50 = lambda pretrained=True, extra_dim=0, model_dir=None: r50(False, extra_dim, model_dir)
        import cutie.model.utils.memory_utils as memory_utils
        import cutie.inference.kv_memory_store as kv_memory_store
        import cutie.inference.memory_manager as memory_manager
        import cutie.inference.inference_core as inference_core
        import cutie.model.cutie as cutie_model
        import cutie.model.transformer.object_transformer as object_transformer
        import cutie.model.transformer.transformer_layers as transformer_layers
        import cutie.model.transformer.positional_encoding as positional_encoding
        import cutie.model.transformer.object_summarizer as object_summarizer
        import cutie.utils.tensor_utils as tensor_utils
    finally:
        sys.path[:] = saved_path
    ns = types.SimpleNamespace(memory_utils=memory_utils, kv_memory_store=kv_memory_store,
                               memory_manager=memory_manager, inference_core=inference_core,
                               cutie_model=cutie_model, object_transformer=object_transformer,
                               transformer_layers=transformer_layers,
                               positional_encoding=positional_encoding,
                               object_summarizer=object_summarizer, tensor_utils=tensor_utils,
                               CUTIE=cutie_model.CUTIE, InferenceCore=inference_core.InferenceCore)
    _loaded = ns
    return ns


def build_reference_model(cfg, seed: int = 0):
    import torch
    from oracle.synth import synthetic_state_dict
    ref = load_reference()
    torch.manual_seed(0)
    net = ref.CUTIE(cfg).eval()
    net.load_state_dict(synthetic_state_dict(net.state_dict(), seed), strict=True)
    return net

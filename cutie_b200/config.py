"""Configuration without Hydra/OmegaConf (neither is needed at run time): a dict with attribute access
holding the reference's defaults (cutie/config/eval_config.yaml:13-51, cutie/config/model/base.yaml).
Any object offering the same attribute *and* item access (e.g. an OmegaConf DictConfig) is accepted by
InferenceCore / CUTIE unchanged."""
import copy


class Config(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v

    def __deepcopy__(self, memo):
        return Config({k: copy.deepcopy(v, memo) for k, v in self.items()})


def to_config(o):
    if isinstance(o, dict):
        return Config({k: to_config(v) for k, v in o.items()})
    return o


_PE_QKV = [True, True, False]

MODEL_BASE = {
    'pixel_mean': [0.485, 0.456, 0.406],
    'pixel_std': [0.229, 0.224, 0.225],
    'pixel_dim': 256, 'key_dim': 64, 'value_dim': 256, 'sensory_dim': 256, 'embed_dim': 256,
    'pixel_encoder': {'type': 'resnet50', 'ms_dims': [1024, 512, 256]},
    'mask_encoder': {'type': 'resnet18', 'final_dim': 256},
    'pixel_pe_scale': 32, 'pixel_pe_temperature': 128,
    'object_transformer': {
        'embed_dim': 256, 'ff_dim': 2048, 'num_heads': 8, 'num_blocks': 3, 'num_queries': 16,
        'read_from_pixel': {'input_norm': False, 'input_add_pe': False, 'add_pe_to_qkv': _PE_QKV},
        'read_from_past': {'add_pe_to_qkv': _PE_QKV},
        'read_from_memory': {'add_pe_to_qkv': _PE_QKV},
        'read_from_query': {'add_pe_to_qkv': _PE_QKV, 'output_norm': False},
        'query_self_attention': {'add_pe_to_qkv': _PE_QKV},
        'pixel_self_attention': {'add_pe_to_qkv': _PE_QKV},
    },
    'object_summarizer': {'embed_dim': 256, 'num_summaries': 16, 'add_pe': True},
    'aux_loss': {'sensory': {'enabled': True, 'weight': 0.01}, 'query': {'enabled': True, 'weight': 0.01}},
    'mask_decoder': {'up_dims': [256, 128, 128]},
}

MODEL_SMALL = dict(MODEL_BASE, pixel_encoder={'type': 'resnet18', 'ms_dims': [256, 128, 64]})

EVAL_DEFAULTS = {
    'amp': False, 'flip_aug': False, 'max_internal_size': -1,
    'use_long_term': False, 'mem_every': 5,           # what get_dataset_cfg escalates for d17-val
    'max_mem_frames': 5,
    'long_term': {'count_usage': True, 'max_mem_frames': 10, 'min_mem_frames': 5, 'num_prototypes': 128,
                  'max_num_tokens': 10000, 'buffer_tokens': 2000},
    'top_k': 30, 'stagger_updates': 5, 'chunk_size': -1, 'save_scores': False, 'save_aux': False,
    'visualize': False,
}


def default_config(model: str = 'base', **overrides) -> Config:
    """eval_config.yaml + model/{base,small}.yaml with keyword overrides; `long_term` merges."""
    cfg = copy.deepcopy(EVAL_DEFAULTS)
    cfg['model'] = copy.deepcopy(MODEL_BASE if model == 'base' else MODEL_SMALL)
    for k, v in overrides.items():
        if k == 'long_term':
            cfg['long_term'].update(v)
        else:
            cfg[k] = v
    return to_config(cfg)

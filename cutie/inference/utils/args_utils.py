"""get_dataset_cfg (cutie/inference/utils/args_utils.py:7-30): escalate per-dataset overrides to the top level."""
import logging

log = logging.getLogger()
_KEYS = ['image_directory', 'mask_directory', 'json_directory', 'size', 'save_all', 'use_all_masks',
         'use_long_term', 'mem_every']


def get_dataset_cfg(cfg):
    data_cfg = cfg.datasets[cfg.dataset]
    for k in _KEYS:
        if cfg.get(k) is not None:
            log.info(f'Overriding config {k} from {data_cfg.get(k)} to {cfg[k]}')
            data_cfg[k] = cfg[k]
        if k in data_cfg:
            cfg[k] = data_cfg[k]
    return data_cfg

"""get_default_model (cutie/utils/get_default_model.py:14-28) without Hydra or network access: default eval
config + cutie-base architecture on the GPU; loads ./weights/cutie-base-mega.pth (or $CUTIE_WEIGHTS) when the
file exists -- the reference downloads it from GitHub, which is impossible offline."""
import logging
import os

import torch

from cutie_b200.config import default_config
from cutie_b200.model.cutie import CUTIE

log = logging.getLogger()


def get_default_model() -> CUTIE:
    cfg = default_config()
    weights = os.environ.get('CUTIE_WEIGHTS', os.path.join('weights', 'cutie-base-mega.pth'))
    cfg['weights'] = weights
    cutie = CUTIE(cfg).cuda().eval()
    if os.path.exists(weights):
        cutie.load_weights(torch.load(weights, map_location='cuda'))
    else:
        log.warning(f'{weights} not found: running with randomly initialised weights')
    return cutie

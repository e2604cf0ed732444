"""`vocoder.load_hifigan` drop-in (vocoder/__init__.py:3-20)."""
import json

import torch


def load_hifigan(state_dict_path, config_file):
    from .hifigan.env import AttrDict
    from .hifigan.models import Generator

    with open(config_file) as f:
        h = AttrDict(json.load(f))
    generator = Generator(h)
    ckpt = torch.load(state_dict_path, map_location='cpu')
    generator.load_state_dict(ckpt['generator'])
    generator.eval()
    generator.remove_weight_norm()
    return generator

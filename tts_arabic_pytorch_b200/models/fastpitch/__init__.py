"""`models.fastpitch` drop-in: FastPitch, FastPitch2Wave, net_config (models/fastpitch/__init__.py:1-41)."""
from ...utils.synth import FASTPITCH_CONFIG as _DEFAULTS
from .networks import FastPitch, FastPitch2Wave  # noqa: F401

# default hyper-parameters used when a checkpoint carries no 'config' entry
net_config = dict(_DEFAULTS)

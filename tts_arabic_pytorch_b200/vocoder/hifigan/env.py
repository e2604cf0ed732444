class AttrDict(dict):
    """dict whose keys are also attributes (what vocoder.hifigan.env.AttrDict provides:
    vocoder/hifigan/env.py:5-8)."""

    def __init__(self, *args, **kwargs):
        dict.__init__(self, *args, **kwargs)
        self.__dict__ = self

"""`models.tacotron2` drop-in: Tacotron2, Tacotron2Wave (models/tacotron2/__init__.py:1)."""
from .networks import Tacotron2, Tacotron2Wave  # noqa: F401

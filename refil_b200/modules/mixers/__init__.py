"""Mixers (reference: src/modules/mixers/{flex_qmix,vdn}.py) -- one class parameterised by args.mixer."""
from ..nets import Mixer

__all__ = ["Mixer"]

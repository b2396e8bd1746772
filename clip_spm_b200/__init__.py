"""clip_spm_b200 -- B200-native (sm_100a) implementation of the CLIP-SPM episode-evaluation hot path
(reference: models/model_clipspm.py::CNN.forward), behind a C-ABI shared library."""
from . import _lib  # noqa: F401
from .model import CNN  # noqa: F401
from .model_clipfsar import CNN_OTAM_CLIPFSAR, CNN_STEN  # noqa: F401
from .model_cpm2c import CLIP_CPMMC_FSAR  # noqa: F401
from .text import TextTower  # noqa: F401
from .tokenizer import ClipTokenizer  # noqa: F401

__all__ = ["CNN", "CNN_OTAM_CLIPFSAR", "CNN_STEN", "CLIP_CPMMC_FSAR", "TextTower", "ClipTokenizer", "_lib"]

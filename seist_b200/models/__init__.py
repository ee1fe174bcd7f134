"""`models` surface of the reference (models/__init__.py:1-13) for the SeisT path."""
from . import seist  # noqa: F401  (registers the 15 seist_* creators)
from .loss import (BCELoss, BinaryFocalLoss, CELoss, CombinationLoss, FocalLoss, HuberLoss,  # noqa: F401
                   MousaviLoss, MSELoss)
from ._factory import (create_model, get_model_list, load_checkpoint, register_model,  # noqa: F401
                       save_checkpoint)

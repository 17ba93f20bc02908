"""B200-native (sm_100a) engine for the patch-based super-resolution path of 4DFlowNet.

The directory name starts with a digit (the project's name), so import it with
``importlib.import_module("4dflownet_b200")``.
"""
from . import _lib  # noqa: F401
from .engine import Engine, Sr4dError  # noqa: F401
from .Network.SR4DFlowNet import SR4DFlowNet, SR4DFlowModel  # noqa: F401
from .Network.PatchGenerator import PatchGenerator  # noqa: F401
from .predictor import prepare_network  # noqa: F401

__all__ = ["Engine", "Sr4dError", "SR4DFlowNet", "SR4DFlowModel", "PatchGenerator", "prepare_network"]

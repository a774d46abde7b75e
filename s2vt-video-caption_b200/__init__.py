"""s2vt-video-caption_b200: B200-native (sm_100a) drop-in for the S2VT encoder-decoder hot path of
Kamino666/S2VT-video-caption.  Import through the `s2vt_b200` shim at the repo root."""
import os as _os

# the wave-front schedule keeps several streams busy at once (two sweeps, the products between them, weight gradients, NCCL): give each
# its own hardware queue so that a stream memory wait never holds up an unrelated stream (read by the driver at context creation)
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

from .att_model import ATT_PARAM_ORDER, Att_Baseline
from .criterion import MaskCriterion
from .data import DeviceFeatureStore, ids_to_sentence, predictions_to_dict
from .lib import LIB_PATH, S2VTLibraryError, launch_count, load
from .model import PARAM_ORDER, S2VT, S2VTModel
from .optim import FusedAdam
from .dp import DataParallelTrainer, GraphedLoopBody
from .training import EarlyStopping, fit, validate

BF16_TRAIN_READY = True     # bench.py: the tensor-core training path is the default for supported shapes

__all__ = ["BF16_TRAIN_READY", "S2VT", "S2VTModel", "Att_Baseline", "ATT_PARAM_ORDER", "MaskCriterion", "FusedAdam", "DeviceFeatureStore", "ids_to_sentence", "predictions_to_dict", "EarlyStopping", "fit", "validate", "PARAM_ORDER", "load", "launch_count", "LIB_PATH",
           "S2VTLibraryError", "DataParallelTrainer", "GraphedLoopBody"]

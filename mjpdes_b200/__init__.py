"""mjpdes_b200 -- B200-native batched run loop for usnistgov/MJPdes line models.

Public surface (mirrors gov.nist.MJPdes.core):
    model.map_to_Model / map_to_ExpoMachine / map_to_Buffer / map_to_JobType, model.read_model,
    model.preprocess_model, model.main_loop, model.calc_basics, model.calc_bneck
    engine.Handle  -- the C-ABI call that replaces core/main-loop-loop for a batch of instances
    scada          -- SCADA record stream -> the reference's log lines
The compute path is libmjpdes_b200.so (CUDA, sm_100a); there is no CPU fallback.
"""
from . import abi, edn  # noqa: F401

__all__ = ["abi", "edn", "engine", "model", "scada", "philox", "dist"]

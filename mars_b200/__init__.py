"""mars_b200 -- B200-native (sm_100a) groupby-aggregation hot path for Mars.

Layout (only what the path needs):
  csrc/ + libmarsgb.so   hand-written CUDA kernels behind the C ABI of include/mgb.h
  _lib.py, kernels.py    ctypes binding / tensor-level wrappers (PyTorch = allocation + streams only)
  device.py              HBM-resident frames exchanged between executors
  executors.py           `fn(ctx, op)` executors for DataFrameGroupByAgg / DataFrameGroupByOperand /
                         DataFrameConcat / DataFrameToGPU (drop-in via Operand.register_executor)
  core.py, dataframe.py, reduction.py, session.py   host-side mirror of the reference interface
  parallel.py            multi-GPU shuffle exchange (grouped ncclSend/ncclRecv)
  plugin.py              registration on the real `mars` package when it is importable
"""
from .core import OperandStage, ExecutionError, execute  # noqa: F401
from .dataframe import (DataFrame, DataFrameConcat, DataFrameFilter, DataFrameSetitem, DataFrameGroupByAgg, DataFrameGroupByOperand,  # noqa: F401
                        DataFrameGroupbyConcatPivot, DataFrameGroupbySortShuffle, DataFrameMergeAlign, DataFramePSRSGroupbySample,
                        DataFrameShuffleProxy, DataFrameToGPU, options)
from .executors import register_gpu_executors  # noqa: F401
from .session import Session, new_session, get_default_session  # noqa: F401
from ._lib import MarsGBError  # noqa: F401

__version__ = "0.1.0"

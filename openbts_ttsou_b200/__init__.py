"""openbts_ttsou_b200 -- B200-native burst DSP for the OpenBTS software transceiver hot path.

The product is libbtsdsp.so (CUDA kernels for sm_100a behind the C ABI in include/btsdsp.h, built by
openbts_ttsou_b200.build); `BtsDsp` is its ctypes face, `shard` the per-GPU work split.  The C++ shim that keeps
the reference's sigProcLib.h signatures lives in openbts_ttsou_b200/host/.
"""
from .api import (BtsDsp, BtsDspError, EXPORTS, LIB_PATH, load_library,  # noqa: F401
                  FULL_SPAN, OVERLAP_ONLY, START_ONLY, WITH_TAIL, NO_DELAY,
                  T_COS, T_SIN, T_ROT, T_REVROT, T_PULSE, T_MID_SEQ, T_MID_META, T_RACH_SEQ, T_RACH_META,
                  T_LPF_RX, T_LPF_TX)
from . import shard  # noqa: F401

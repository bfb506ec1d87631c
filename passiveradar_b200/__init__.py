"""passiveradar_b200 -- B200-native (sm_100a) core for the passive-radar hot path.

Drop-in operators with the reference's call signatures (Max-Manning/passiveRadar):

    from passiveradar_b200 import fast_xambg, LS_Filter, NLMS_filter, block_NLMS

and a device-resident frame pipeline for streams of CPI frames
(``passiveradar_b200.frames.FramePipeline``).  All arithmetic runs in
``libprcore.so`` (hand-written CUDA, C ABI in ``include/prcore.h``); importing
this package never imports the CPU oracle and never falls back to the CPU.
"""
from .range_doppler_processing import fast_xambg            # noqa: F401
from .clutter_removal import LS_Filter, NLMS_filter, block_NLMS   # noqa: F401

__all__ = ["fast_xambg", "LS_Filter", "NLMS_filter", "block_NLMS"]
__version__ = "0.1.0"

"""hulk_b200 -- B200-native (sm_100a) implementation of the `hulk sketch` hot path of will-rowe/hulk v1.0.0.

The product is libhulk_b200.so (hand-written CUDA kernels behind the C ABI in include/hulk_b200.h)
plus the `hulk` command-line front end; this package is the Python mirror of the reference's
pipeline interface over that ABI.  There is no CPU fallback: every numeric operation fails loudly
if the CUDA library is missing or no GPU is present.
"""
import os as _os0

# One hardware queue per CUDA stream, read when the device's context is created: the asynchronous host path (feeder thread)
# is only taken when every stream of a context has a queue of its own (hulk_b200_create).  A default, not an override.
_os0.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

from ._native import load, LIB_PATH, EXPORTS  # noqa: F401,E402
from .sketch import (GroupSketch, HistoSketch, HulkError, load_sketch, md5_mins, new_cws, pack_bases, pack_reads, sketch_json,  # noqa: F401
                     sketch_reads, smash, spectrum_size)
from .distributed import ShardedSketch, chunk_range, sketch_reads_sharded, slot_range  # noqa: F401
from .seqio import NativeReader, read_fastq, read_fasta, synthetic_reads  # noqa: F401

import os as _os

CLI_PATH = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "bin", "hulk")   # the `hulk` front end

__version__ = "1.0.0"

"""svscope_b200 — B200-native localGraph hot path of SVScope (POA + edit distance + sequence
mixture model) behind the reference's Python call signatures.  See DESIGN.md."""
__version__ = "0.1.0"

import os as _os

# Rounds of the alignment scheduler run on many streams; they only overlap on the device when
# the streams map to different hardware queues.  Must be set before the CUDA context exists.
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

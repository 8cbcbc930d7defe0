"""B200-native H.264 picture-reconstruction engine (hot path of jfu222/h264_video_decoder_demo).

The product is the CUDA library behind include/h264_recon_b200.h; this package holds its
sources (csrc/), the ctypes mirror of the ABI, the host-side mirror of the reference's
decoder interface and the replay-file reader.  Nothing here imports oracle/.
"""
from . import abi, replay, sharding  # noqa: F401

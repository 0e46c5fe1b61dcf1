"""ltx_video_swift_mlx_b200 -- B200-native (sm_100a) implementation of the LTX-2 denoise hot path.

The arithmetic lives in libltxcuda.so (hand-written CUDA, C ABI in include/ltxcuda.h).  This package is the host-side
mirror of the reference's Swift interface for that path (LTXTransformer, LTXScheduler, LatentUtils, decodeVideo and the
generateVideo step loop), implemented over the C ABI with ctypes.  There is no CPU fallback."""
__version__ = "0.1.0"

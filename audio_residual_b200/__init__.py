"""audio_residual_b200: B200-native (sm_100a CUDA) implementation of Audio-ResiDual's HTSAT + ResiDual hot path."""

"""sound_event_detection_transformer_b200: the B200-native SEDT hot path.

    from sound_event_detection_transformer_b200.sedt import build_model

mirrors the reference's `from sedt import build_model` (sedt/__init__.py:8).
The forward and the Hungarian matcher run in hand-written sm_100a CUDA kernels
(csrc/) behind a C ABI (include/sedt_b200.h); there is no CPU fallback.
"""
__version__ = "0.1.0"

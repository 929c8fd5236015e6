"""B200-native Lift-Splat (camera -> BEV) hot path for Multimodal-XAD."""
__version__ = "0.1.0"

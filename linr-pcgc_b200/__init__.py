"""linr-pcgc_b200: B200-native overfit + coding hot path of LINR-PCGC."""
__version__ = "0.1.0"

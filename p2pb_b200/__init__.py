"""p2pb_b200 -- B200-native implementation of P2P-Bridge's denoising hot path."""
__version__ = "0.1.0"

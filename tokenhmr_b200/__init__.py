"""tokenhmr_b200 — B200-native inference engine for TokenHMR's per-image forward path."""
__version__ = "0.1.0"

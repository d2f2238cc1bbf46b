"""tscode_b200 — B200-native conformer-ensemble hot path for TSCoDe."""
__version__ = "0.1.0"

"""distar_b200 — B200-native AlphaStar policy hot path for DI-star."""

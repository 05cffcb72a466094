"""jtk_b200: B200-native per-chunk pair-HMM path of ban-m/jtk."""

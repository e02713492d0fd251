"""pam-b200: B200-native per-frame geometric hot path of Part-Aware Measurement."""

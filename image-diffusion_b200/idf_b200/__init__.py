"""idf_b200 — host side of the B200-native latent-diffusion hot path (ctypes binding, kernel wrappers, engine)."""

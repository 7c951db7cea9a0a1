"""Drop-in module tree of the reference project (ArgoHA/custom_d_fine): ``src.d_fine`` and ``src.dl.train`` keep
the reference's import paths, names and signatures for the training hot path and forward everything to the
B200-native implementation in ``custom_d_fine_b200``."""

"""``src.dl`` of the reference: only the training entry point (train.py) is on the hot path."""

"""Device-backed mirrors of the reference's `datasets/` helpers that feed the hot path."""

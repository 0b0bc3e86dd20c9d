"""metadamage_b200 — B200-native (sm_100a) implementation of metadamage's per-TaxID
damage-fitting hot path (counts.py -> fits.py), behind the reference's own Python seams."""
__version__ = "0.1.0"

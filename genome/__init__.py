"""Namespace package: `genome.distance_b200` is the B200-native k-mer set distance engine."""

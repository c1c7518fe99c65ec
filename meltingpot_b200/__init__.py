"""meltingpot_b200: a B200-native batched Melting Pot substrate engine."""

__version__ = '0.1.0'

"""normalisr_b200: B200-native (sm_100a) implementation of Normalisr's linear
association-testing hot path (co-expression and differential expression)."""
__version__ = "0.1.0"
__all__ = ["association", "coex", "de", "normalisr", "engine"]

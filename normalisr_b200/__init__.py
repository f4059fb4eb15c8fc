"""normalisr_b200: B200-native (sm_100a) implementation of Normalisr's linear
association-testing hot path (co-expression and differential expression) and the steps on either
side of it (lcpm, compute_var, normvar upstream; binnet downstream; text I/O)."""
__version__ = "0.2.0"
__all__ = ["association", "binnet", "coex", "de", "engine", "io", "lcpm", "norm", "normalisr", "parallel"]

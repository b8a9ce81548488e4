"""svscope_b200 — B200-native localGraph hot path of SVScope (POA + edit distance + sequence
mixture model) behind the reference's Python call signatures.  See DESIGN.md."""
__version__ = "0.1.0"

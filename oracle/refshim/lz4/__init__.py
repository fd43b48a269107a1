"""Identity stand-in for `lz4` (import-time only; see refshim/__init__.py)."""

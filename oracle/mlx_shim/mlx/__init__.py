"""Oracle-only stand-in package so `import mlx.core as mx` resolves (see core.py)."""

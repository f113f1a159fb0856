"""Import shim (test infrastructure)."""
def use(*a, **k): pass

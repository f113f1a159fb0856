"""Import shim (test infrastructure)."""

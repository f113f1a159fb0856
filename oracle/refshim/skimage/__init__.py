"""Import shim (test infrastructure)."""
from . import measure

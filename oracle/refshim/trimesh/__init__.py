"""Import shim (test infrastructure)."""
class Trimesh: pass

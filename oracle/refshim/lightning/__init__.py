"""Import shim (test infrastructure): the oracle never launches Fabric."""
class Fabric:
    def __init__(self, *a, **k):
        raise RuntimeError("lightning shim: Fabric is not available in the oracle harness")

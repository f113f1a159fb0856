def __getattr__(name):
    raise RuntimeError("matplotlib shim")

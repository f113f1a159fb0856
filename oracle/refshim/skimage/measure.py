def marching_cubes(*a, **k):
    raise RuntimeError("skimage shim")

class PeakSignalNoiseRatio:
    def __init__(self, *a, **k): pass
class StructuralSimilarityIndexMeasure:
    def __init__(self, *a, **k): pass

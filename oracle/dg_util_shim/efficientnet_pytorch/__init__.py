"""Import stub: backbone_models.py:2 imports EfficientNet unconditionally; never used on the hot path."""


class EfficientNet:
    @staticmethod
    def from_name(name):
        raise NotImplementedError("efficientnet_pytorch stub (not installed; out of scope)")

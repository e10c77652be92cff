"""mmf_b200: B200-native generation hot path of dfaroughy/Multimodal-flows (drop-in Python surface)."""
from .param_spec import make_config
from .tensorclass import DataCoupling, TensorMultiModal

__all__ = ["make_config", "DataCoupling", "TensorMultiModal"]

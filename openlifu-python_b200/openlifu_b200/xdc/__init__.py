from .element import Element
from .transducer import Transducer, TransformedTransducer

__all__ = ["Element", "Transducer", "TransformedTransducer"]

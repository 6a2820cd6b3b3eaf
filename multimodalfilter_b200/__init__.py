"""B200-native (sm_100a) differentiable filtering recursion behind the torchfilter / crossmodal API."""
__version__ = "0.1.0"

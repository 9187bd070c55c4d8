"""B200-native segmented cumprod / cumsum compositing path (drop-in for the reference's
`grouped_cumprod` torch extension).  Importing this package never touches the GPU; the ops
load libgcp_b200.so on first use and fail loudly if it is missing."""
from .ops import (grouped_cumprod_backward, grouped_cumprod_forward, grouped_cumsum_forward,  # noqa: F401
                  validate_segments)
from .autograd import GroupedCumprod, grouped_cumprod  # noqa: F401

__all__ = ["grouped_cumprod_forward", "grouped_cumsum_forward", "grouped_cumprod_backward",
           "validate_segments", "GroupedCumprod", "grouped_cumprod"]

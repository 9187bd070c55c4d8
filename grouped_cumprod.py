"""Drop-in module with the reference extension's name.

`import grouped_cumprod` (gs_model.py:8, cuda_test.py:6 of the reference) resolves to this file
when the repository root is on sys.path; it exposes exactly the three callables the reference's
PYBIND11_MODULE registers (cuda_kernel/cuda_kernel.cpp:17-22).
"""
from simplegaussiansplat_tk71_b200.ops import (grouped_cumprod_backward,  # noqa: F401
                                               grouped_cumprod_forward, grouped_cumsum_forward)

__all__ = ["grouped_cumprod_forward", "grouped_cumprod_backward", "grouped_cumsum_forward"]

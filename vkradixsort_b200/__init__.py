"""vkradixsort_b200 -- B200-native (sm_100a) LSD radix sort behind VkRadixSort's pass surface.

The product is the C-ABI shared library built from vkradixsort_b200/csrc (declared in
include/vkradixsort_b200.h, C++ facade in include/vkradixsort_b200.hpp).  The Python modules here
are the test / bench drivers over that ABI:

  capi    ctypes binding, one method per C entry point
  passes  mirror of the reference's MultiRadixSortPass / SingleRadixSortPass / MultiRadixSort /
          SingleRadixSort host interface
  dist    one-process-per-GPU bucket exchange (torch.distributed) around the local sort
"""
from . import capi  # noqa: F401
from .capi import Handle, MultiPushConstants, SinglePushConstants, VkrsError, multi_push_constants  # noqa: F401
from .passes import (GPUContext, MultiRadixSort, MultiRadixSortPass, SingleRadixSort,  # noqa: F401
                     SingleRadixSortPass)

__all__ = ["capi", "Handle", "MultiPushConstants", "SinglePushConstants", "VkrsError", "multi_push_constants",
           "GPUContext", "MultiRadixSort", "MultiRadixSortPass", "SingleRadixSort", "SingleRadixSortPass"]

"""raw_image_pipeline_b200 -- B200-native (sm_100a) implementation of the per-frame RAW chain of
leggedrobotics/raw_image_pipeline behind the reference's own ``RawImagePipeline`` API.

The pixel work is done by hand-written CUDA kernels in ``librip_b200.so`` (C ABI declared in
``include/rip_b200.h``); this package is the Python face of that library, mirroring the
reference's pybind module ``py_raw_image_pipeline``.
"""
from .multi_gpu import MultiGpuPipeline  # noqa: F401
from .pipeline import RawImagePipeline, RawImagePipelineError  # noqa: F401

__all__ = ["RawImagePipeline", "RawImagePipelineError", "MultiGpuPipeline"]

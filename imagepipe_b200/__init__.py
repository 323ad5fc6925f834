"""imagepipe-b200: the raw->sRGB OpBuffer hot path of pedrocr/imagepipe as sm_100a CUDA kernels.

The package is a thin host mirror of the reference's Pipeline / ImageOp surface (pipeline.py) over the
C-ABI library libipb200.so (include/ipb200.h, sources in csrc/).  Nothing here computes pixels on the CPU.
"""
from ._capi import IpbError, LIB_PATH, lib  # noqa: F401
from .pipeline import (Context, DeviceArray, ImageSource, OpBaseCurve, OpBuffer, OpDemosaic, OpFromLab,  # noqa: F401
                       OpGamma, OpGoFloat, OpRotateCrop, OpToLab, OpTransform, Pipeline, PipelineCache, PipelineGlobals,
                       PipelineOps, PipelineSettings, Rotation, SplineFunc, SRGBImage, SRGBImage16,
                       calculate_scale, default_context, lanczos_resize, rotate_buffer, scale_down_srgb, scaling_size,
                       synth_cfa_u16)

__version__ = "0.1.0"

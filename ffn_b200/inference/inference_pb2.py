"""Drop-in for ffn/inference/inference_pb2.py (runtime-built descriptors, see protos.py)."""
from .protos import (AlignmentOptions, CoordinateExpressionOptions, CounterValue, DecoratedVolume,  # noqa: F401
                     ImageMaskOptions, InferenceOptions, InferenceRequest, MaskChannelConfig,
                     MaskConfig, MaskConfigs, ResegmentationPoint, ResegmentationRequest,
                     SegmentationSource, TaskCounters, VolumeMaskOptions)

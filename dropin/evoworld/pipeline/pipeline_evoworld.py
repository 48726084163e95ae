"""Shadows evoworld/pipeline/pipeline_evoworld.py: the denoise loop runs on evoworld_b200."""
from evoworld_b200.pipeline import StableVideoDiffusionPipeline, StableVideoDiffusionPipelineOutput, _append_dims  # noqa: F401

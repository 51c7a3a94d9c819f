"""reference signerf/signerf_pipeline.py"""
from signerf_b200.plugin.pipeline import SIGNeRFPipeline, SIGNeRFPipelineConfig  # noqa: F401

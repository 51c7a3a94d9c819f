"""reference signerf/renderer/renderer.py -> the CUDA proxy-mesh rasteriser behind the same classes."""
from signerf_b200.plugin.renderer import NERFSTUDIO_BLENDER_SCALE_RATIO, Renderer, RendererConfig  # noqa: F401

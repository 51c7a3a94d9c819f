"""reference signerf/diffuser/diffuser.py -> the in-process SDXL + ControlNet denoiser behind `mode="custom"`."""
from signerf_b200.plugin.diffuser import Diffuser, DiffuserConfig, InProcessSDXL  # noqa: F401

"""Host-side mirror of the SIGNeRF plugin surface for the reference-sheet hot path.

Same class / method / field names as the reference modules they replace:
    signerf/renderer/renderer.py                 -> plugin.renderer      (RendererConfig, Renderer)
    signerf/diffuser/diffuser.py                 -> plugin.diffuser      (DiffuserConfig, Diffuser; mode="custom" built)
    signerf/datasetgenerator/datasetgenerator.py -> plugin.datasetgenerator (DatasetGeneratorConfig, DatasetGenerator:
                                                    render_camera / generate_reference_sheet / generate_with_reference_sheet)
    SIGNeRFModel eval hook (signerf/signerf.py)  -> plugin.model.FusedNerfactoGraph (get_outputs_for_camera_ray_bundle)
INTEGRATION.md shows how a maintainer wires them into `ns-train signerf`.  nerfstudio is imported lazily and only for
its config base class; everything here works without it (it is not installable in the build container)."""
from .base import CameraBatch  # noqa: F401
from .datasetgenerator import DatasetGenerator, DatasetGeneratorConfig  # noqa: F401
from .diffuser import Diffuser, DiffuserConfig, InProcessSDXL  # noqa: F401
from .model import FusedNerfactoGraph  # noqa: F401
from .pipeline import SIGNeRFPipeline, SIGNeRFPipelineConfig  # noqa: F401
from .renderer import Renderer, RendererConfig  # noqa: F401
from . import base  # noqa: F401

"""`signerf` — the reference's package name, provided by signerf_b200 for its hot path.

The modules of the reference-sheet path live here under the reference's own import paths, so that
`ns-train signerf` (entry points `signerf.signerf_config:signerf_method`, `signerf.signerf_nerfacto_config:
signerf_nerfacto_method`, reference pyproject.toml:44-46) resolves them without a single edit:

    signerf.renderer.renderer                 RendererConfig, Renderer                 (reference signerf/renderer/renderer.py)
    signerf.diffuser.diffuser                 DiffuserConfig, Diffuser                 (signerf/diffuser/diffuser.py)
    signerf.datasetgenerator.datasetgenerator DatasetGeneratorConfig, DatasetGenerator (signerf/datasetgenerator/datasetgenerator.py)
    signerf.signerf_pipeline                  SIGNeRFPipelineConfig, SIGNeRFPipeline   (signerf/signerf_pipeline.py)

Everything else of the reference (trainer, data stack, viewer / interface, model losses, method configs, utils) is
out of scope for this repo and is NOT re-implemented: this package is an OVERLAY.  When a checkout / install of
cgtuebingen/SIGNeRF is reachable — `SIGNERF_REFERENCE_DIR=<path to its signerf/ directory>`, or another `signerf`
directory further down `sys.path` — its directory is appended to this package's `__path__`, so
`signerf.signerf_trainer`, `signerf.data.*`, `signerf.interface.*`, `signerf.signerf`, `signerf.signerf_config`, ...
import from there while the four modules above shadow their reference namesakes.  The reference model stays as it is:
`DatasetGenerator.render_camera` attaches the fused renderer behind `graph.get_outputs_for_camera_ray_bundle`
(plugin/datasetgenerator.py `_fused_graph`)."""
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))


def _reference_package_dirs():
    env = os.environ.get("SIGNERF_REFERENCE_DIR")
    if env:
        yield env
    for entry in sys.path:
        cand = os.path.join(entry or ".", "signerf")
        if os.path.isdir(cand) and os.path.isfile(os.path.join(cand, "signerf_trainer.py")):
            yield cand


for _d in _reference_package_dirs():
    _d = os.path.abspath(_d)
    if _d != _here and _d not in __path__:
        __path__.append(_d)   # AFTER our own directory: our modules win, the rest falls through to the reference
        break

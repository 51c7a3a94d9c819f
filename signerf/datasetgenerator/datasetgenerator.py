"""reference signerf/datasetgenerator/datasetgenerator.py -> the fused reference-sheet generator."""
from signerf_b200.plugin.datasetgenerator import DatasetGenerator, DatasetGeneratorConfig  # noqa: F401

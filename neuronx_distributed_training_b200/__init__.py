"""B200-native 3D-parallel LLM training framework (capabilities of neuronx-distributed-training)."""
from ._version import __version__  # noqa: F401

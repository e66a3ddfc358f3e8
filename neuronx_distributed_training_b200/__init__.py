"""B200-native 3D-parallel LLM training framework (capabilities of neuronx-distributed-training)."""
__version__ = "0.1.0"

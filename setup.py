"""Packaging for the B200-native training framework.

``pip install -e .`` installs the pure-Python package; the native pieces (sm_100a kernels ``ops/_C.so``, the Megatron
dataset index builder and the checkpoint writer) are compiled **in-tree** by ``python setup.py build_native`` (or
``install_setup.sh`` / ``python -c "import __graft_entry__ as g; g.build()"``) so that they sit next to their sources,
ship with a repository snapshot and are visible to tooling that records which extensions a process loaded.
"""
import os
import sys

from setuptools import Command, find_packages, setup

ROOT = os.path.dirname(os.path.abspath(__file__))


class BuildNative(Command):
    description = "compile the CUDA (sm_100a) and C++ extensions in-tree"
    user_options = []

    def initialize_options(self):
        pass

    def finalize_options(self):
        pass

    def run(self):
        sys.path.insert(0, ROOT)
        import __graft_entry__ as entry
        entry.build()


def _version():
    ns = {}
    with open(os.path.join(ROOT, "neuronx_distributed_training_b200", "_version.py")) as f:
        exec(f.read(), ns)
    return ns["__version__"]


setup(
    name="neuronx-distributed-training-b200",
    version=_version(),
    description="B200-native (sm_100a) 3D-parallel LLM training framework with the capabilities of "
                "neuronx-distributed-training",
    packages=find_packages(include=["neuronx_distributed_training_b200", "neuronx_distributed_training_b200.*"]),
    package_data={"neuronx_distributed_training_b200": ["models/configs/*.json", "ops/csrc/*", "data/csrc/*",
                                                        "checkpoint/csrc/*", "ops/*.so", "data/*.so", "checkpoint/*.so"]},
    python_requires=">=3.10",
    install_requires=["torch>=2.4", "numpy", "pyyaml"],
    extras_require={"hf": ["transformers", "datasets"], "eval": ["transformers"], "s3": ["boto3"],
                    "logging": ["tensorboard", "wandb", "mlflow"]},
    cmdclass={"build_native": BuildNative},
)

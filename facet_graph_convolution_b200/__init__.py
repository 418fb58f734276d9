"""facet_graph_convolution_b200 -- B200-native facet-graph convolution hot path.

Drop-in for the FeaStNet-style layer and the multi-scale normal-denoising network of
Elensil/Facet_Graph_Convolution (reference Code/model.py), built from scratch as hand-written
sm_100a CUDA kernels behind a C ABI (include/facetconv_b200.h).  ``mesh`` (host-side synthetic
generators), ``mesh_io`` (OBJ files) and ``checkpoint`` (Saver-file interchange) import without the CUDA library; everything else needs libfacetconv_b200.so and a
CUDA device and fails loudly otherwise.
"""
from . import mesh  # noqa: F401  (NumPy only)

__all__ = ["mesh", "mesh_io", "checkpoint", "ops", "model", "autograd", "torch_ops", "build_library"]


def build_library(force=False):
    from .build import build
    return build(force=force)


def __getattr__(name):
    if name in ("ops", "model", "autograd", "patches", "train", "checkpoint", "mesh_io", "coarsening", "torch_ops"):
        import importlib
        return importlib.import_module("." + name, __name__)
    raise AttributeError(name)

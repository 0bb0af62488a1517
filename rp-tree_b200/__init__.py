"""rp-tree_b200: B200-native random-projection-forest engine behind the Data.RPTree API surface.

Only the hot path of ocramz/rp-tree lives here: csrc/ (hand-written sm_100a CUDA + the C ABI of
include/rpforest.h) and the host-side mirror of the reference's interface for that path (api.py),
plus tree-sharded multi-GPU plumbing (dist.py).  The directory name has a hyphen; import it as
`rp_tree_b200` (see the shim rp_tree_b200.py at the repo root).
"""
from ._lib import RPForestError, lib, SO_PATH, SIGNATURES
from .api import (RPForest, SparseRows, RPTreeConfig, metricL2, rpTreeCfg, sampleHyperplanes, topologyPlan, slice_hyperplanes,
                  forestBatch, treeBatch, forest, tree, chunksOf, knn, knnPQ, knnH, candidates, recallWith, levels, leafSizes,
                  treeSize, points, serialiseRPForest, deserialiseRPForest)
from . import _build
from . import dist
from . import idx

__all__ = ["RPForest", "SparseRows", "RPForestError", "RPTreeConfig", "metricL2", "rpTreeCfg", "sampleHyperplanes", "topologyPlan",
           "slice_hyperplanes", "forestBatch", "treeBatch", "forest", "tree", "chunksOf", "knn", "knnPQ", "knnH", "candidates", "recallWith",
           "levels", "leafSizes", "treeSize", "points", "serialiseRPForest", "deserialiseRPForest", "lib", "SO_PATH", "SIGNATURES"]

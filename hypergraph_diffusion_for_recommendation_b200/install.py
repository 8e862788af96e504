"""Plug this package into an UNMODIFIED checkout of the reference (``HD_SELFRec/``): ``install()``.

The reference's entry chain is ``main.py:125-138`` -> ``SELFRec.execute`` (SELFRec.py:37-42) -> ``<Model>(GraphRecommender)``
(base/graph_recommender.py:18-45) -> ``model/graph/<Model>.py``.  Those files stay as they are.  What changes is what their
imports resolve to (seams 1-6):

=================================  =======================================================================================
``base.torch_interface``            ``TorchGraphInterface.convert_sparse_mat_to_tensor`` returns a ``DeviceCSR`` (seam 2); because
                                    ``DeviceCSR`` implements ``__torch_function__``, the encoders' own
                                    ``torch.sparse.mm(self.sparse_norm_adj, x)`` lines run on ``hgr_spmm_f32`` (seam 3)
``util.loss_torch``                 ``bpr_loss / l2_reg_loss / contrastLoss / InfoNCE`` on libhgr.so (seam 4)
``util.sampler``                    ``next_batch_pairwise`` on the device sampler
``data.ui_graph`` / ``data.loader`` ``Interaction`` / ``FileIO.load_data_set``: the array-backed facade, matrices from the device builder (seam 1)
``GraphRecommender.test``           one ``hgr_fullrank_topk_f32`` call for all test users, same ``rec_list`` dict (seam 5)
``util.evaluation.ranking_evaluation``  same strings, metric sums on the device
=================================  =======================================================================================

A name this package does not provide (``util.loss_torch.kl_divergence``, ``FileIO.load_kg_data``, the other samplers ...)
falls through to the reference's own module, loaded under a private name, so every other model of the reference keeps importing.

    import hypergraph_diffusion_for_recommendation_b200 as hgr
    hgr.install("/path/to/HD_SELFRec")      # before `import SELFRec` / `from model.graph.LightGCN import LightGCN`
"""
from __future__ import annotations

import importlib
import importlib.util
import os
import sys
import types

SEAMS = ("base.torch_interface", "util.loss_torch", "util.sampler", "data.ui_graph", "data.loader")
_state = {"installed": False, "root": None}


class _SeamModule(types.ModuleType):
    """A module whose attributes come from this package first and from the reference's original module otherwise."""

    def __init__(self, name, provided: dict, original_path: str | None):
        super().__init__(name)
        self.__dict__.update(provided)
        self.__dict__["_hgr_original_path"] = original_path
        self.__dict__["_hgr_original"] = None
        self.__dict__["__hgr_seam__"] = True

    def __getattr__(self, attr):
        if attr.startswith("__") and attr.endswith("__"):
            raise AttributeError(attr)
        orig = self.__dict__.get("_hgr_original")
        path = self.__dict__.get("_hgr_original_path")
        if orig is None and path and os.path.exists(path):
            spec = importlib.util.spec_from_file_location("_hgr_reference_" + self.__name__.replace(".", "_"), path)
            orig = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(orig)
            self.__dict__["_hgr_original"] = orig
        if orig is not None and hasattr(orig, attr):
            return getattr(orig, attr)
        raise AttributeError("module %r (hgr seam) has no attribute %r" % (self.__name__, attr))


def _fileio_class(root):
    """``data.loader.FileIO`` with ``load_data_set`` from the facade and every other static method from the reference's class."""
    from . import data as hdata

    base = object
    path = os.path.join(root, "data", "loader.py") if root else None
    if path and os.path.exists(path):
        spec = importlib.util.spec_from_file_location("_hgr_reference_data_loader", path)
        orig = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(orig)
        base = orig.FileIO

    class FileIO(base):  # noqa: D401 - same name as the reference's class on purpose
        load_data_set = staticmethod(hdata.FileIO.load_data_set)

    return FileIO


def _test(self):
    """``GraphRecommender.test`` (base/graph_recommender.py:61-92): ``rec_list[user] = [(item, score), ...]`` of length
    ``max_N`` for every user of ``data.test_set`` from ``self.user_emb / self.item_emb`` -- one fused call instead of a python
    loop over users (``predict`` + mask + ``find_k_largest``); ``refquirk`` mode reproduces the reference's lists exactly."""
    from . import evaluation

    return evaluation.test(self, self.user_emb, self.item_emb, mode="refquirk")


def install(reference_root: str | None = None, patch_test: bool = True) -> dict:
    """Register the seam modules in ``sys.modules`` and patch ``GraphRecommender.test``.  Call it before the reference's model
    modules are imported (modules that already bound the old names keep them).  Returns ``{seam name: module}``."""
    from . import data as hdata
    from . import encoders, evaluation, loss_torch, sampler

    root = reference_root or _state["root"]
    if root:
        root = os.path.abspath(root)
        if root not in sys.path:
            sys.path.insert(0, root)
    _state["root"] = root

    def orig(rel):
        return os.path.join(root, *rel.split("/")) if root else None

    mods = {
        "base.torch_interface": _SeamModule("base.torch_interface", {"TorchGraphInterface": encoders.TorchGraphInterface},
                                            orig("base/torch_interface.py")),
        "util.loss_torch": _SeamModule("util.loss_torch", {k: getattr(loss_torch, k) for k in
                                                           ("bpr_loss", "l2_reg_loss", "contrastLoss", "InfoNCE", "bpr_l2_from_tables")},
                                       orig("util/loss_torch.py")),
        "util.sampler": _SeamModule("util.sampler", {"next_batch_pairwise": sampler.next_batch_pairwise}, orig("util/sampler.py")),
        "data.ui_graph": _SeamModule("data.ui_graph", {"Interaction": hdata.Interaction}, orig("data/ui_graph.py")),
        "data.loader": _SeamModule("data.loader", {"FileIO": _fileio_class(root)}, orig("data/loader.py")),
    }
    for name, mod in mods.items():
        sys.modules[name] = mod
        parent = sys.modules.get(name.split(".")[0])
        if parent is not None:  # `import util.sampler` style access through the already imported package object
            setattr(parent, name.split(".")[1], mod)
    if patch_test:
        try:
            ev = importlib.import_module("util.evaluation")
            ev.ranking_evaluation = evaluation.ranking_evaluation
            gr = importlib.import_module("base.graph_recommender")
            gr.GraphRecommender.test = _test
            gr.ranking_evaluation = evaluation.ranking_evaluation
            mods["base.graph_recommender.GraphRecommender.test"] = gr.GraphRecommender.test
        except ImportError:
            if root:
                raise
    _state["installed"] = True
    return mods


def installed() -> bool:
    return bool(_state["installed"])

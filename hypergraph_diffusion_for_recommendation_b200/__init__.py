"""B200-native embedding-propagation / loss / full-rank-eval path of the SELFRec-based hypergraph-diffusion recommender.

``install(reference_root)`` plugs the package into an unmodified checkout of the reference (see install.py)."""


def install(reference_root=None, patch_test=True):
    from .install import install as _install

    return _install(reference_root, patch_test)

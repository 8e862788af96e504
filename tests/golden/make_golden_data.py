#!/usr/bin/env python
"""Golden vectors for the data façade (tests/test_data_facade.py): run the UNMODIFIED reference's
``FileIO.load_data_set`` (data/loader.py:24-38) and ``Interaction`` (data/ui_graph.py) on small messy files and dump
what they hold -- id maps, dict-of-dict sets in their insertion order, sizes, accessors, and the five scipy matrices.

    python tests/golden/make_golden_data.py      # needs /root/reference; rewrites tests/golden/data_facade.{json,npz}

Nothing is copied from the reference: it is imported, called, and only its outputs are stored.
"""
import json
import os
import sys
import warnings

import numpy as np

REF = os.environ.get("HGR_REFERENCE", "/root/reference/HD_SELFRec")
HERE = os.path.dirname(os.path.abspath(__file__))
warnings.filterwarnings("ignore")

TRAIN_TAB = """user\titem
10\t100
7\t103
10\t101
3\t100
10\t100
7\t100
12\t109
3\t105
3\t103
7\t103
12\t100
5\t101
"""
TEST_TAB = """user\titem
10\t103
99\t100
7\t555
7\t101
10\t105
10\t103
3\t109
41\t42
"""
TRAIN_COMMA = "u,i,r\n4,9,5\n2,9,3\n4,8,1\n2,7,2\n"
TRAIN_MIXED = "u\ti\n4\t9\n2,9\n4\t8\n"


def csr(m):
    m = m.tocsr().copy()
    m.sum_duplicates()
    m.sort_indices()
    return m.indptr.astype(np.int64), m.indices.astype(np.int32), m.data.astype(np.float32)


def main():
    sys.path.insert(0, REF)
    from data.loader import FileIO
    from data.ui_graph import Interaction

    files = {"train_tab.txt": TRAIN_TAB, "test_tab.txt": TEST_TAB, "train_comma.txt": TRAIN_COMMA, "train_mixed.txt": TRAIN_MIXED}
    tmp = os.path.join(HERE, "_tmp_data")
    os.makedirs(tmp, exist_ok=True)
    for name, text in files.items():
        with open(os.path.join(tmp, name), "w") as f:
            f.write(text)
    loaded = {name: FileIO.load_data_set(os.path.join(tmp, name)) for name in files}
    for name in files:
        os.remove(os.path.join(tmp, name))
    os.rmdir(tmp)

    # a second, weighted case that never goes through the loader: ratings other than 1 and a rewritten pair; 2 users x 3 items, so
    # normalize_graph_mat takes its rectangular branch (the tab case has 5 users and 5 items: the SQUARE branch, data/graph.py:14-19)
    weighted_train = [[1, 5, 1.0], [2, 5, 0.5], [1, 6, 2.0], [1, 5, 3.0], [2, 6, 1.0], [1, 6, 1.0], [2, 7, 1.0]]
    weighted_test = [[2, 5, 1.0], [1, 9, 4.0], [1, 9, 2.0]]
    cases = {"tab": (loaded["train_tab.txt"], loaded["test_tab.txt"]), "weighted": (weighted_train, weighted_test)}
    out, arrays = {"files": files, "loaded": loaded, "cases": {}}, {}
    for cname, (train, test) in cases.items():
        d = Interaction(None, [list(e) for e in train], [list(e) for e in test])
        users, items = list(d.user.keys()), list(d.item.keys())
        c = {
            "train": train, "test": test,
            "user": [[k, v] for k, v in d.user.items()], "item": [[k, v] for k, v in d.item.items()],
            "id2user": [[k, v] for k, v in d.id2user.items()], "id2item": [[k, v] for k, v in d.id2item.items()],
            "training_set_u": [[u, [[k, v] for k, v in d.training_set_u[u].items()]] for u in d.training_set_u],
            "training_set_i": [[i, [[k, v] for k, v in d.training_set_i[i].items()]] for i in d.training_set_i],
            "test_set": [[u, [[k, v] for k, v in d.test_set[u].items()]] for u in d.test_set],
            "user_history_dict": [[u, list(v)] for u, v in d.user_history_dict.items()],
            "test_set_item": sorted(d.test_set_item),
            "n_users": d.n_users, "n_items": d.n_items, "n_cf_train": d.n_cf_train, "n_cf_test": d.n_cf_test,
            "training_size": list(d.training_size()), "test_size": list(d.test_size()),
            "user_rated": [[u, [list(x) for x in d.user_rated(u)]] for u in users],
            "item_rated": [[i, [list(x) for x in d.item_rated(i)]] for i in items],
            "get_user_id": [[u, d.get_user_id(u)] for u in users + [424242]],
            "get_item_id": [[i, d.get_item_id(i)] for i in items + [424242]],
            "contain": [[u, i, d.contain(u, i)] for u in users + [424242] for i in items + [424242]],
            "contain_user": [[u, d.contain_user(u)] for u in users + [424242]],
            "contain_item": [[i, d.contain_item(i)] for i in items + [424242]],
            "edge_index": d.edge_index.tolist(), "edge_index_t": d.edge_index_t.tolist(),
        }
        out["cases"][cname] = c
        arrays[cname + "_row"] = np.stack([d.row(k) for k in range(d.n_users)])
        arrays[cname + "_col"] = np.stack([d.col(k) for k in range(d.n_items)])
        arrays[cname + "_matrix"] = d.matrix()
        for m in ("ui_adj", "norm_adj", "interaction_mat", "inv_interaction_mat", "norm_interaction_mat", "norm_inv_interaction_mat"):
            ip, ix, dv = csr(getattr(d, m))
            arrays["%s_%s_indptr" % (cname, m)], arrays["%s_%s_indices" % (cname, m)], arrays["%s_%s_values" % (cname, m)] = ip, ix, dv
        if cname == "tab":
            # convert_to_laplacian_mat of the interaction matrix with two entries knocked out (what SGL's edge dropout hands it)
            pert = d.interaction_mat.tolil()
            pert[0, 0] = 0
            pert[2, 1] = 0
            ip, ix, dv = csr(d.convert_to_laplacian_mat(pert.tocsr()))
            arrays["tab_laplacian_indptr"], arrays["tab_laplacian_indices"], arrays["tab_laplacian_values"] = ip, ix, dv
            arrays["tab_laplacian_dropped"] = np.array([[0, 0], [2, 1]])
    with open(os.path.join(HERE, "data_facade.json"), "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)
    np.savez_compressed(os.path.join(HERE, "data_facade.npz"), **arrays)
    print("wrote data_facade.json / data_facade.npz:", {k: v.shape for k, v in arrays.items() if k.endswith("values")})


if __name__ == "__main__":
    main()

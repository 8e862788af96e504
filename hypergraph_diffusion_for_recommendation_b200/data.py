"""Data layer façade: ``FileIO.load_data_set`` and ``Interaction`` of the reference, array-backed.

Reference: ``data/loader.py:24-38`` reads ``train.txt`` / ``test.txt`` into a python list of ``[user, item, 1.0]``;
``data/ui_graph.py:12-112`` (``Interaction``) then walks that list in python to number users and items in order of first
appearance, fills dict-of-dict training / test sets, and builds five scipy matrices with list comprehensions and dict
lookups -- 5 s at 1 M interactions, 22 s at 3 M, impossible at 1 B.

Here the interaction list is three numpy columns (``InteractionList``; it still iterates, indexes, shuffles and
``len()``s like the list it replaces), the id numbering is one ``np.unique`` per side, the dict-of-dict sets are
read-only mapping VIEWS over CSR arrays that keep python-dict semantics (key order = first insertion, a repeated pair
keeps its first position and its last rating, raw ids as python ints), and every matrix comes out of the device builder
(``graph.build_norm_adj`` / ``build_interaction_csr``) as a ``DeviceCSR`` on first use.  ``GraphRecommender``, the
samplers and the metric code of the reference read the same attributes and get the same answers
(``tests/test_data_facade.py`` against vectors dumped from the reference's own classes).

There is no CPU fallback for the matrices: touching ``norm_adj`` & co. without a CUDA device raises ``HgrError``.
"""
from __future__ import annotations

from collections.abc import Mapping
from re import split

import numpy as np

from . import _lib


# ------------------------------------------------------------------------------------------------ interaction list
class InteractionList:
    """``[[user, item, weight], ...]`` held as three columns.  Quacks like the list ``FileIO.load_data_set`` returns:
    ``len``, iteration, integer / slice indexing, item assignment (``random.shuffle(training_data)`` in
    util/sampler.py:239 swaps entries in place), ``np.array(data)`` (util/sampler.py:9)."""

    def __init__(self, users, items, weights=None):
        self.users = np.ascontiguousarray(users, dtype=np.int64)
        self.items = np.ascontiguousarray(items, dtype=np.int64)
        if self.users.shape != self.items.shape or self.users.ndim != 1:
            raise ValueError("users and items must be 1-D arrays of the same length")
        self.weights = (np.ones(self.users.size, dtype=np.float64) if weights is None
                        else np.ascontiguousarray(weights, dtype=np.float64))
        if self.weights.shape != self.users.shape:
            raise ValueError("weights must match users")

    @classmethod
    def from_entries(cls, entries) -> "InteractionList":
        if isinstance(entries, cls):
            return entries
        n = len(entries)
        if n == 0:
            return cls(np.zeros(0, np.int64), np.zeros(0, np.int64))
        arr = np.asarray(entries, dtype=np.float64)  # the reference truncates with int(user), int(item)
        if arr.ndim != 2 or arr.shape[1] < 3:
            raise ValueError("interaction entries must be [user, item, rating]")
        return cls(arr[:, 0].astype(np.int64), arr[:, 1].astype(np.int64), arr[:, 2])

    def __len__(self):
        return int(self.users.size)

    def __getitem__(self, k):
        if isinstance(k, slice):
            return InteractionList(self.users[k], self.items[k], self.weights[k])
        return [int(self.users[k]), int(self.items[k]), float(self.weights[k])]

    def __setitem__(self, k, entry):
        self.users[k], self.items[k], self.weights[k] = int(entry[0]), int(entry[1]), float(entry[2])

    def __iter__(self):
        for u, i, w in zip(self.users.tolist(), self.items.tolist(), self.weights.tolist()):
            yield [u, i, w]

    def __array__(self, dtype=None, copy=None):
        out = np.stack([self.users.astype(np.float64), self.items.astype(np.float64), self.weights], axis=1)
        return out if dtype is None else out.astype(dtype)

    def __eq__(self, other):
        if isinstance(other, InteractionList):
            return (np.array_equal(self.users, other.users) and np.array_equal(self.items, other.items)
                    and np.array_equal(self.weights, other.weights))
        try:
            return len(other) == len(self) and all(list(a) == b for a, b in zip(other, self))
        except TypeError:
            return NotImplemented


class FileIO(object):
    """``data/loader.py`` -- only the loader the graph recommenders use."""

    @staticmethod
    def load_data_set(file, rec_type='graph') -> InteractionList:
        """``FileIO.load_data_set`` (data/loader.py:24-38): skip the header line; a line is split on tabs if it has one,
        on commas otherwise; the first two fields are ``int`` ids; the weight is always 1.  Parsed with pandas' C reader
        when the file uses one separator throughout, line by line (the reference's loop) otherwise."""
        import pandas as pd

        with open(file) as f:
            f.readline()
            probe = f.readline()
        if probe == "":
            return InteractionList(np.zeros(0, np.int64), np.zeros(0, np.int64))
        sep = '\t' if '\t' in probe else ','
        try:
            # a line that does not fit the probe's separator leaves a non-integer field behind and sends us to the loop below
            df = pd.read_csv(file, sep=sep, header=None, skiprows=1, usecols=[0, 1], dtype=np.int64, engine='c',
                             skip_blank_lines=False)
            return InteractionList(df[0].to_numpy(), df[1].to_numpy())
        except Exception:
            users, items = [], []
            with open(file) as f:
                next(f)
                for line in f:
                    parts = split(',', line.strip()) if '\t' not in line else split('\t', line.strip())
                    users.append(int(parts[0]))
                    items.append(int(parts[1]))
            return InteractionList(users, items)


# ------------------------------------------------------------------------------------------------ mapping views
class IdMap(Mapping):
    """raw id -> dense id, dense ids numbered in order of first appearance (``self.user`` / ``self.item`` of
    ``Interaction``, data/ui_graph.py:47-52).  Iteration and ``keys()`` give the raw ids in dense order, as the dict does."""

    def __init__(self, raw_by_dense: np.ndarray):
        self.raw_by_dense = np.ascontiguousarray(raw_by_dense, dtype=np.int64)
        order = np.argsort(self.raw_by_dense, kind="stable")
        self._sorted_raw = self.raw_by_dense[order]
        self._dense_of_sorted = order.astype(np.int64)

    def lookup(self, raw, default: int = -1) -> np.ndarray:
        """Vectorised ``[self.get(r, default) for r in raw]``."""
        raw = np.asarray(raw, dtype=np.int64)
        if self._sorted_raw.size == 0:
            return np.full(raw.shape, default, dtype=np.int64)
        pos = np.searchsorted(self._sorted_raw, raw)
        pos[pos >= self._sorted_raw.size] = 0
        hit = self._sorted_raw[pos] == raw
        return np.where(hit, self._dense_of_sorted[pos], default)

    def __getitem__(self, raw):
        try:
            d = int(self.lookup(np.asarray([raw]))[0]) if float(raw) == int(raw) else -1
        except (TypeError, ValueError):
            d = -1
        if d < 0:
            raise KeyError(raw)
        return d

    def __contains__(self, raw):
        try:
            return self[raw] >= 0
        except KeyError:
            return False

    def __iter__(self):
        return iter(self.raw_by_dense.tolist())

    def __len__(self):
        return int(self.raw_by_dense.size)


class DenseToRaw(Mapping):
    """dense id -> raw id (``id2user`` / ``id2item``)."""

    def __init__(self, raw_by_dense: np.ndarray):
        self.raw_by_dense = raw_by_dense

    def __getitem__(self, dense):
        d = int(dense)
        if d != dense or not 0 <= d < self.raw_by_dense.size:
            raise KeyError(dense)
        return int(self.raw_by_dense[d])

    def __iter__(self):
        return iter(range(self.raw_by_dense.size))

    def __len__(self):
        return int(self.raw_by_dense.size)


class _Row(Mapping):
    """One inner dict ``{raw column id: rating}`` in insertion order."""

    __slots__ = ("_cols", "_vals", "_dict")

    def __init__(self, cols: np.ndarray, vals: np.ndarray):
        self._cols, self._vals, self._dict = cols, vals, None

    def _d(self):
        if self._dict is None:
            self._dict = dict(zip(self._cols.tolist(), self._vals.tolist()))
        return self._dict

    def __getitem__(self, k):
        return self._d()[k]

    def __contains__(self, k):
        return k in self._d()

    def __iter__(self):
        return iter(self._cols.tolist())

    def __len__(self):
        return int(self._cols.size)

    def __repr__(self):
        return repr(self._d())


_EMPTY_I = np.zeros(0, np.int64)
_EMPTY_F = np.zeros(0, np.float64)


class RowsView(Mapping):
    """``defaultdict(dict)`` view over CSR arrays: ``view[raw_row][raw_col] -> rating``.  A missing row reads as an empty
    dict, like the defaultdict (which would also INSERT it; nothing in the reference depends on that)."""

    def __init__(self, keys: IdMap, indptr: np.ndarray, cols: np.ndarray, vals: np.ndarray):
        self.key_map, self.indptr, self.cols, self.vals = keys, indptr, cols, vals

    def __getitem__(self, raw):
        try:
            r = self.key_map[raw]
        except KeyError:
            return _Row(_EMPTY_I, _EMPTY_F)
        a, b = self.indptr[r], self.indptr[r + 1]
        return _Row(self.cols[a:b], self.vals[a:b])

    def __contains__(self, raw):
        return raw in self.key_map

    def __iter__(self):
        return iter(self.key_map)

    def __len__(self):
        return len(self.key_map)


class _ListRows(RowsView):
    """``user_history_dict``: raw user -> python list of raw items (duplicates kept)."""

    def __getitem__(self, raw):
        try:
            r = self.key_map[raw]
        except KeyError:
            return {}
        return self.cols[self.indptr[r]:self.indptr[r + 1]].tolist()


_PACK_LIMIT = np.int64(1) << np.int64(62)


def _group_sorted(key: np.ndarray, want_inverse: bool = False):
    """Distinct values of the non-negative int64 ``key`` (ascending) with the position of the first and of the last entry
    of each.  When ``max(key) * n`` fits 62 bits the position is packed under the key and ONE plain ``np.sort`` does it
    (numpy's vectorised quicksort: 25 x faster than the stable argsort ``np.unique(return_index=True)`` runs); otherwise
    ``np.unique`` + a reverse fancy assignment (repeated indices keep the last value written)."""
    n = key.size
    pos = np.arange(n, dtype=np.int64)
    kmax = int(key.max()) if n else 0
    if n and int(key.min()) >= 0 and (kmax + 1) * n < int(_PACK_LIMIT):
        packed = np.sort(key * np.int64(n) + pos)
        k = packed // np.int64(n)
        head = np.ones(n, dtype=bool)
        head[1:] = k[1:] != k[:-1]
        starts = np.nonzero(head)[0]
        ends = np.append(starts[1:], n) - 1
        out = (k[starts], packed[starts] % np.int64(n), packed[ends] % np.int64(n))
        if want_inverse:  # group of every entry: scatter the running group number back to the entry's position
            inv = np.empty(n, dtype=np.int64)
            inv[packed % np.int64(n)] = np.cumsum(head) - 1
            out += (inv,)
        return out
    uk, inv = np.unique(key, return_inverse=True)
    inv = inv.reshape(-1)
    first = np.empty(uk.size, dtype=np.int64)
    last = np.empty(uk.size, dtype=np.int64)
    first[inv[::-1]] = pos[::-1]
    last[inv] = pos
    return (uk, first, last, inv.astype(np.int64)) if want_inverse else (uk, first, last)


def _first_seen(x: np.ndarray):
    """Dense ids in order of first appearance: ``(dense_of_entry, raw_by_dense)``."""
    lo = int(x.min()) if x.size else 0
    uniq, first, _, inv = _group_sorted(x - np.int64(lo), want_inverse=True)
    order = np.argsort(first)  # distinct values, one per distinct id
    raw_by_dense = (uniq[order] + np.int64(lo)).astype(np.int64)
    rank = np.empty_like(order)
    rank[order] = np.arange(order.size)
    return rank[inv].astype(np.int64), raw_by_dense


def _dict_rows(row: np.ndarray, col_key: np.ndarray, col_out: np.ndarray, val: np.ndarray, n_rows: int, n_col_keys: int):
    """CSR of ``d[row][col] = val`` executed entry by entry: inside a row the columns stand in order of first insertion and
    carry the LAST value written.  ``col_key`` are dense column keys (< n_col_keys) used to find repeats, ``col_out`` what
    the view shows."""
    n = row.size
    if n == 0:
        return np.zeros(n_rows + 1, np.int64), _EMPTY_I, _EMPTY_F
    ncol = np.int64(max(n_col_keys, 1))
    uk, first_pos, last_pos = _group_sorted(row * ncol + col_key)  # distinct (row, col) pairs, row-major
    pair_row = uk // ncol
    # rows ascending, insertion order inside a row: sort the distinct keys row * n + first_pos, read the position back out
    first_sorted = np.sort(pair_row * np.int64(n) + first_pos) % np.int64(n)
    last_of_first = np.empty(n, dtype=np.int64)
    last_of_first[first_pos] = last_pos
    indptr = np.zeros(n_rows + 1, dtype=np.int64)
    np.cumsum(np.bincount(pair_row, minlength=n_rows), out=indptr[1:])
    return indptr, col_out[first_sorted], val[last_of_first[first_sorted]]


# ------------------------------------------------------------------------------------------------ Interaction
class Data(object):
    """data/data.py"""

    def __init__(self, conf, training, test):
        self.config = conf
        self.training_data = training
        self.test_data = test


class Interaction(Data):
    """``data/ui_graph.py:12-178`` with the same attributes and methods; see the module docstring for what is underneath."""

    def __init__(self, conf, training, test, device="cuda"):
        self.conf = conf
        training, test = InteractionList.from_entries(training), InteractionList.from_entries(test)
        Data.__init__(self, conf, training, test)
        self.device = device
        # __generate_set (data/ui_graph.py:43-68)
        self.dense_u, id2user = _first_seen(training.users)
        self.dense_i, id2item = _first_seen(training.items)
        self.user, self.item = IdMap(id2user), IdMap(id2item)
        self.id2user, self.id2item = DenseToRaw(id2user), DenseToRaw(id2item)
        nu, ni = id2user.size, id2item.size
        ip, cols, vals = _dict_rows(self.dense_u, self.dense_i, training.items, training.weights, nu, ni)
        self.training_set_u = RowsView(self.user, ip, cols, vals)
        ip, cols, vals = _dict_rows(self.dense_i, self.dense_u, training.users, training.weights, ni, nu)
        self.training_set_i = RowsView(self.item, ip, cols, vals)
        ones = training.weights == 1.0  # user_history_dict: every rating-1 entry, repeats included (:54-57)
        hu, hist_users = _first_seen(training.users[ones]) if ones.any() else (_EMPTY_I, _EMPTY_I)
        o = np.argsort(hu, kind="stable")
        hp = np.zeros(hist_users.size + 1, dtype=np.int64)
        np.cumsum(np.bincount(hu, minlength=hist_users.size), out=hp[1:])
        self.user_history_dict = _ListRows(IdMap(hist_users), hp, training.items[ones][o], _EMPTY_F)
        known = self.user.lookup(test.users) >= 0  # test entries of users never seen in training are skipped (:62-66)
        tu, ti, tw = test.users[known], test.items[known], test.weights[known]
        if tu.size:
            td, test_users = _first_seen(tu)
            tk, test_items = _first_seen(ti)
        else:
            td, test_users, tk, test_items = _EMPTY_I, _EMPTY_I, _EMPTY_I, _EMPTY_I
        ip, cols, vals = _dict_rows(td, tk, ti, tw, test_users.size, test_items.size)
        self.test_set = RowsView(IdMap(test_users), ip, cols, vals)
        self.test_set_item = set(test_items.tolist())
        self.n_users, self.n_items = int(nu), int(ni)
        self.n_cf_train, self.n_cf_test = len(training), len(test)
        self._mats = {}

    # ---- matrices: built on the device on first use (data/ui_graph.py:36-41,70-112; data/graph.py:11-25) ------------
    def _mat(self, name):
        m = self._mats.get(name)
        if m is None:
            import torch

            from . import graph

            if not torch.cuda.is_available():
                raise _lib.HgrError("Interaction.%s is built on the GPU (no CPU path)" % name)
            u, i, nu, ni, dev = self.dense_u, self.dense_i, self.n_users, self.n_items, self.device
            if name == "ui_adj":
                m = graph.build_norm_adj(u, i, nu, ni, device=dev, normalize=False)
            elif name == "norm_adj":
                m = graph.build_norm_adj(u, i, nu, ni, device=dev)
            elif name.startswith("norm_"):
                # normalize_graph_mat picks its branch by SHAPE (data/graph.py:14-24): with as many users as items the
                # interaction matrix is "square" and gets D^-1/2 R D^-1/2 with the ROW sums on both sides, else D^-1 R
                m = self.normalize_graph_mat(self._mat(name[5:]))
            else:
                m = graph.build_interaction_csr(u, i, nu, ni, device=dev, transpose="inv" in name)
            self._mats[name] = m
        return m

    ui_adj = property(lambda self: self._mat("ui_adj"))
    norm_adj = property(lambda self: self._mat("norm_adj"))
    interaction_mat = property(lambda self: self._mat("interaction_mat"))
    inv_interaction_mat = property(lambda self: self._mat("inv_interaction_mat"))
    norm_interaction_mat = property(lambda self: self._mat("norm_interaction_mat"))
    norm_inv_interaction_mat = property(lambda self: self._mat("norm_inv_interaction_mat"))

    @property
    def edge_index(self):
        import torch

        return torch.from_numpy(np.stack([self.dense_u, self.dense_i]))

    @property
    def edge_index_t(self):
        import torch

        return torch.from_numpy(np.stack([self.dense_i, self.dense_u]))

    def normalize_graph_mat(self, adj_mat):
        """``Graph.normalize_graph_mat`` (data/graph.py:11-25) for a matrix that is already a ``DeviceCSR``: square ->
        ``D^-1/2 A D^-1/2``, rectangular -> ``D^-1 A`` with D the row sums."""
        import torch

        from . import graph

        rows = torch.repeat_interleave(torch.arange(adj_mat.shape[0], device=adj_mat.device), adj_mat.indptr[1:] - adj_mat.indptr[:-1])
        rowsum = torch.zeros(adj_mat.shape[0], dtype=torch.float32, device=adj_mat.device).index_add_(0, rows, adj_mat.values)
        deg = rowsum.to(torch.int32)
        if not torch.equal(deg.to(torch.float32), rowsum):
            raise _lib.HgrError("normalize_graph_mat on the device needs integer row sums (unit-weight interactions)")
        vals = adj_mat.values.clone()
        if adj_mat.shape[0] == adj_mat.shape[1]:
            d = graph._degree_scale(deg, -0.5)
            graph._scale(adj_mat.indptr, adj_mat.indices, vals, adj_mat.shape[0], d, d)
        else:
            graph._scale(adj_mat.indptr, adj_mat.indices, vals, adj_mat.shape[0], graph._degree_scale(deg, -1.0), None)
        return graph.DeviceCSR(adj_mat.indptr, adj_mat.indices, vals, adj_mat.shape, symmetric=adj_mat.symmetric,
                               chunk_nnz=adj_mat.chunk_nnz, split=adj_mat.split)

    def convert_to_laplacian_mat(self, adj_mat):
        """data/ui_graph.py:86-93: the normalised ``(U+I)^2`` adjacency of a (perturbed) ``[U, I]`` interaction matrix, whose
        VALUES are carried over (a pair that was listed twice in training weighs 2).  Entries with value 0 stand for
        dropped edges (``DeviceCSR.with_values``) and vanish, as they do from scipy's ``.nonzero()``."""
        import torch

        from . import graph

        n_u, n_i = adj_mat.shape
        rows = torch.repeat_interleave(torch.arange(n_u, device=adj_mat.device, dtype=torch.int32), adj_mat.indptr[1:] - adj_mat.indptr[:-1])
        keep = adj_mat.values != 0
        r, c, v = rows[keep], adj_mat.indices[keep], adj_mat.values[keep]
        pattern = graph.build_norm_adj(r, c, n_u, n_i, device=adj_mat.device, normalize=False)  # unit weights, canonical CSR
        # user rows keep the order of adj_mat's rows; item rows list the same entries sorted by (item, user)
        vals = torch.cat([v, v[torch.sort(c, stable=True).indices]])
        return self.normalize_graph_mat(graph.DeviceCSR(pattern.indptr, pattern.indices, vals, pattern.shape, symmetric=True,
                                                        chunk_nnz=pattern.chunk_nnz, split=pattern.split))

    # ---- accessors (data/ui_graph.py:114-178) -----------------------------------------------------------------------
    def get_user_id(self, u):
        if u in self.user:
            return self.user[u]

    def get_item_id(self, i):
        if i in self.item:
            return self.item[i]

    def training_size(self):
        return len(self.user), len(self.item), len(self.training_data)

    def test_size(self):
        return len(self.test_set), len(self.test_set_item), len(self.test_data)

    def contain(self, u, i):
        'whether user u rated item i'
        return bool(u in self.user and i in self.training_set_u[u])

    def contain_user(self, u):
        'whether user is in training set'
        return u in self.user

    def contain_item(self, i):
        """whether item is in training set"""
        return i in self.item

    def user_rated(self, u):
        r = self.training_set_u[u]
        return list(r.keys()), list(r.values())

    def item_rated(self, i):
        r = self.training_set_i[i]
        return list(r.keys()), list(r.values())

    def row(self, u):
        k, v = self.user_rated(self.id2user[u])
        vec = np.zeros(len(self.item))
        vec[self.item.lookup(k)] = v
        return vec

    def col(self, i):
        k, v = self.item_rated(self.id2item[i])
        vec = np.zeros(len(self.user))
        vec[self.user.lookup(k)] = v
        return vec

    def matrix(self):
        m = np.zeros((len(self.user), len(self.item)))
        v = self.training_set_u
        rows = np.repeat(np.arange(self.n_users), np.diff(v.indptr))
        m[rows, self.item.lookup(v.cols)] = v.vals
        return m

    # ---- what the device-side consumers of this package read (no python loop over interactions) --------------------
    def dense_training_pairs(self):
        """``(user_idx, item_idx)`` of every training entry in file order (``self.user[pair[0]]``, ``self.item[pair[1]]``)."""
        return self.dense_u, self.dense_i

    def eval_arrays(self):
        """For ``evaluation.EvalData``: test users (dense, ``test_set`` order), their raw ids, the ground truth as CSR of
        dense item ids (-1: item never seen in training), and the dense -> raw item table."""
        t = self.test_set
        raw_users = t.key_map.raw_by_dense
        return (self.user.lookup(raw_users), raw_users.tolist(), t.indptr.copy(), self.item.lookup(t.cols), self.id2item.raw_by_dense)

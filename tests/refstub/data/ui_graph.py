import numpy as np
import scipy.sparse as sp


class Interaction(object):
    def __init__(self, conf, training, test):
        self.training_data, self.test_data = training, test
        self.user, self.item, self.id2user, self.id2item = {}, {}, {}, {}
        self.training_set_u, self.test_set = {}, {}
        for u, i, w in training:
            if u not in self.user:
                self.user[u] = len(self.user)
                self.id2user[self.user[u]] = u
            if i not in self.item:
                self.item[i] = len(self.item)
                self.id2item[self.item[i]] = i
            self.training_set_u.setdefault(u, {})[i] = w
        for u, i, w in test:
            if u in self.user:
                self.test_set.setdefault(u, {})[i] = w
        self.n_users, self.n_items = len(self.user), len(self.item)
        r = np.array([self.user[p[0]] for p in training])
        c = np.array([self.item[p[1]] for p in training]) + self.n_users
        n = self.n_users + self.n_items
        a = sp.csr_matrix((np.ones(len(r), np.float32), (r, c)), shape=(n, n))
        a = a + a.T
        d = np.power(np.asarray(a.sum(1)).ravel(), -0.5)
        d[np.isinf(d)] = 0
        self.norm_adj = sp.diags(d).dot(a).dot(sp.diags(d)).tocsr()

    def get_user_id(self, u):
        return self.user.get(u)

    def user_rated(self, u):
        return list(self.training_set_u[u].keys()), list(self.training_set_u[u].values())

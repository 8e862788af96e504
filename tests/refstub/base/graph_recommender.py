import numpy as np
import torch

from data.loader import FileIO  # noqa: F401  (bound at import time, like the reference does)
from data.ui_graph import Interaction
from util.evaluation import ranking_evaluation


class GraphRecommender(object):
    def __init__(self, conf, training_set, test_set, **kwargs):
        self.config = conf
        self.data = Interaction(conf, training_set, test_set)
        self.topN = [int(n) for n in kwargs.get("item_ranking", "10,20").split(",")]
        self.max_N = max(self.topN)
        self.batch_size = int(kwargs.get("batch_size", 256))
        self.emb_size = int(kwargs.get("embedding_size", 64))
        self.user_emb = self.item_emb = None

    def predict(self, u):
        return torch.matmul(self.user_emb[self.data.get_user_id(u)], self.item_emb.transpose(0, 1)).cpu().numpy()

    def test(self):
        """python loop over the test users: scores, training items pushed down, arg-sort (what install() replaces)"""
        rec_list = {}
        for user in self.data.test_set:
            cand = self.predict(user)
            for item in self.data.user_rated(user)[0]:
                cand[self.data.item[item]] = -10e8
            ids = np.argsort(-cand, kind="stable")[:self.max_N]
            rec_list[user] = [(self.data.id2item[int(i)], float(cand[i])) for i in ids]
        return rec_list

    def fast_evaluation(self, epoch):
        rec_list = self.test()
        return ranking_evaluation(self.data.test_set, rec_list, self.topN), rec_list

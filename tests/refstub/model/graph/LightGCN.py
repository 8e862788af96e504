import torch
import torch.nn as nn

from base.graph_recommender import GraphRecommender
from base.torch_interface import TorchGraphInterface
from util.loss_torch import bpr_loss, l2_reg_loss
from util.sampler import next_batch_pairwise


class LightGCN(GraphRecommender):
    def __init__(self, conf, training_set, test_set, **kwargs):
        super(LightGCN, self).__init__(conf, training_set, test_set, **kwargs)
        self.n_layers = int(kwargs.get("n_layers", 2))
        self.reg, self.lRate, self.maxEpoch = float(kwargs.get("reg", 1e-4)), float(kwargs.get("lrate", 0.01)), int(kwargs.get("max_epoch", 2))
        self.model = LGCN_Encoder(self.data, self.emb_size, self.n_layers)
        self.history = []

    def train(self):
        model = self.model.cuda()
        optimizer = torch.optim.Adam(model.parameters(), lr=self.lRate)
        for epoch in range(self.maxEpoch):
            losses = []
            for user_idx, pos_idx, neg_idx in next_batch_pairwise(self.data, self.batch_size):
                rec_user_emb, rec_item_emb = model()
                user_emb, pos_item_emb, neg_item_emb = rec_user_emb[user_idx], rec_item_emb[pos_idx], rec_item_emb[neg_idx]
                batch_loss = bpr_loss(user_emb, pos_item_emb, neg_item_emb) + l2_reg_loss(self.reg, user_emb, pos_item_emb, neg_item_emb) / self.batch_size
                optimizer.zero_grad()
                batch_loss.backward()
                optimizer.step()
                losses.append(batch_loss.item())
            with torch.no_grad():
                self.user_emb, self.item_emb = model()
                measures, _ = self.fast_evaluation(epoch)
            self.history.append((sum(losses) / len(losses), measures))


class LGCN_Encoder(nn.Module):
    def __init__(self, data, emb_size, n_layers):
        super(LGCN_Encoder, self).__init__()
        self.data, self.layers = data, n_layers
        init = nn.init.xavier_uniform_
        self.embedding_dict = nn.ParameterDict({
            "user_emb": nn.Parameter(init(torch.empty(data.n_users, emb_size))),
            "item_emb": nn.Parameter(init(torch.empty(data.n_items, emb_size))),
        })
        self.sparse_norm_adj = TorchGraphInterface.convert_sparse_mat_to_tensor(data.norm_adj).cuda()

    def forward(self):
        ego = torch.cat([self.embedding_dict["user_emb"], self.embedding_dict["item_emb"]], 0)
        layers = [ego]
        for _ in range(self.layers):
            ego = torch.sparse.mm(self.sparse_norm_adj, ego)
            layers.append(ego)
        out = torch.mean(torch.stack(layers, dim=1), dim=1)
        return out[:self.data.n_users], out[self.data.n_users:]

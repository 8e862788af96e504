from random import choice, shuffle

import torch


def next_batch_pairwise(data, batch_size, n_negs=1, device=None):
    pairs = list(data.training_data)
    shuffle(pairs)
    items = list(data.item.keys())
    for s in range(0, len(pairs), batch_size):
        u_idx, i_idx, j_idx = [], [], []
        for user, item, _ in pairs[s:s + batch_size]:
            u_idx.append(data.user[user])
            i_idx.append(data.item[item])
            neg = choice(items)
            while neg in data.training_set_u[user]:
                neg = choice(items)
            j_idx.append(data.item[neg])
        yield torch.LongTensor(u_idx), torch.LongTensor(i_idx), torch.LongTensor(j_idx)

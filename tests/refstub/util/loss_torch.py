import torch


def bpr_loss(user_emb, pos_item_emb, neg_item_emb):
    pos = torch.mul(user_emb, pos_item_emb).sum(dim=1)
    neg = torch.mul(user_emb, neg_item_emb).sum(dim=1)
    return torch.mean(-torch.log(10e-6 + torch.sigmoid(pos - neg)))


def l2_reg_loss(reg, *args):
    total = 0
    for emb in args:
        total = total + torch.norm(emb, p=2)
    return total * reg


def only_in_the_tree(x):
    """a helper the B200 package does not provide: install() must let it fall through to this module"""
    return x + 1

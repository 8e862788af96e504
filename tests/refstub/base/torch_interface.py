import numpy as np
import torch


class TorchGraphInterface(object):
    @staticmethod
    def convert_sparse_mat_to_tensor(X):
        coo = X.tocoo()
        i = torch.from_numpy(np.vstack([coo.row, coo.col]).astype(np.int64))
        return torch.sparse_coo_tensor(i, torch.from_numpy(coo.data.astype(np.float32)), coo.shape)

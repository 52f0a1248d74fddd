"""Embedding table loader, interface of reference ``index/datasets.py`` (:6-21): ``np.load`` the whole
(N, dim) array; ``__getitem__`` accepts an int or a list of ints (generate_indices.py:117 relies on
the list form) and returns a float32 tensor."""
import numpy as np
import torch
import torch.utils.data as data


class EmbDataset(data.Dataset):
    def __init__(self, data_path, mmap=False):
        self.data_path = data_path
        self.embeddings = np.load(data_path, mmap_mode="r" if mmap else None)
        self.dim = self.embeddings.shape[-1]

    def __getitem__(self, index):
        return torch.as_tensor(np.asarray(self.embeddings[index]), dtype=torch.float32)

    def __len__(self):
        return len(self.embeddings)

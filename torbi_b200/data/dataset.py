"""File dataset (restates /root/reference/torbi/data/dataset.py:10-29)."""
import torch

import torbi_b200


class Dataset(torch.utils.data.Dataset):

    def __init__(self, input_files):
        self.input_files = input_files

    def __getitem__(self, index):
        """Load the index-th observation file, chunked when MIN_CHUNK_SIZE is set"""
        input_file = self.input_files[index]
        observation = torch.load(input_file, map_location='cpu')
        if torbi_b200.MIN_CHUNK_SIZE is not None:
            observation = torbi_b200.chunk(observation)
        return observation, input_file

    def __len__(self):
        return len(self.input_files)

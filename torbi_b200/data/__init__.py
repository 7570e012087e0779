"""Batch assembly, file dataset, loader and native file I/O of the file API.
Public names match torbi.data: Dataset, collate, separate, loader."""
from . import collate as _collate_module
from . import dataset as _dataset_module
from . import loader as _loader_module

Dataset = _dataset_module.Dataset
collate = _collate_module.collate
separate = _collate_module.separate
loader = _loader_module.loader

__all__ = ['Dataset', 'collate', 'separate', 'loader']

from .collate import collate, separate
from .dataset import Dataset
from .loader import loader

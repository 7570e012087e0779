"""torbi_b200 -- B200-native drop-in for torbi's Viterbi decoding path.

    import torbi_b200 as torbi
    indices = torbi.from_probabilities(observation, ..., gpu=0)

Public surface = the hot-path subset of /root/reference/torbi/__init__.py:72-78:
decode, from_probabilities, from_file, from_file_to_file, from_files_to_files,
from_dataloader, save, save_masked, chunk, data, and the decode-relevant
configuration constants.  The decode itself lives in libtorbi_b200.so
(hand-written CUDA for sm_100a behind the C ABI of include/torbi_b200.h).
"""
import importlib.util
import pathlib

###############################################################################
# Configuration (reference: yapecs.configure, torbi/__init__.py:7-17)
###############################################################################

from .config import defaults as _defaults
from .config.defaults import *  # noqa: F401,F403


def configure(config_file):
    """Override configuration constants from a Python file, the way
    `--config file.py` does through yapecs in the reference."""
    spec = importlib.util.spec_from_file_location('torbi_b200_user_config', str(config_file))
    module = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(module)
    for name in dir(module):
        if name.isupper() and hasattr(_defaults, name):
            globals()[name] = getattr(module, name)
            setattr(_defaults, name, getattr(module, name))


###############################################################################
# Module imports
###############################################################################

from . import _ffi                                   # noqa: E402
from . import timing                                 # noqa: E402
from .viterbi import decode, decode_host, decode_packed, decode_host_packed   # noqa: E402
from .core import *                                  # noqa: E402,F401,F403
from .core import lpt_shards, size_bucketed_batches  # noqa: E402,F401
from .chunk import chunk                             # noqa: E402
from . import data                                   # noqa: E402
from . import evaluate                               # noqa: E402
from . import ops                                    # noqa: E402,F401  (registers torch.ops.torbi_b200)

LIBRARY = pathlib.Path(_ffi.LIBRARY)

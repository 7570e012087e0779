"""Decode-relevant configuration constants, same names and defaults as
/root/reference/torbi/config/defaults.py (lines cited per constant).  The
dataset download / partition bookkeeping constants of that file are out of scope
(SURVEY.md section 8); the evaluation runner takes its directories from `root`.  Override with `torbi_b200.configure(path)` or the CLI's
`--config path.py` (the reference does this through yapecs,
/root/reference/torbi/__init__.py:10-11)."""

# Configuration name (defaults.py:10)
CONFIG = 'torbi'

# When set to a positive integer, enables chunking for long sequences by
# splitting sequences at low-entropy frames (defaults.py:39-41)
MIN_CHUNK_SIZE = None

# Threshold below which to split the sequence when performing chunked decoding
# (defaults.py:44)
ENTROPY_THRESHOLD = 0.5

# Seed for all random number generators (defaults.py:71)
RANDOM_SEED = 1234

# Batch size (defaults.py:80)
BATCH_SIZE = 512

# Number of parallel CPU workers (defaults.py:83)
NUM_WORKERS = 0

# Datasets evaluated by `python -m torbi_b200.evaluate` (defaults.py:56 lists daps / vctk, which need
# downloads; the synthetic corpus stands in)
DATASETS = ['synthetic']

# Thresholds (in number of 5 cent bins) for raw pitch accuracy (defaults.py:62)
PITCH_ERROR_THRESHOLDS = [0, 1, 2]

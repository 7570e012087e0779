"""Command line: same flags as /root/reference/torbi/__main__.py:12-50, with
`--gpu` accepting one or more indices (several = files sharded across GPUs)."""
import argparse
from pathlib import Path

import torbi_b200


def parse_args(argv=None):
    """Parse command-line arguments"""
    parser = argparse.ArgumentParser(
        prog='python -m torbi_b200',
        description='Viterbi-decode time-varying categorical distribution files on B200 GPUs')
    parser.add_argument(
        '--input_files', type=Path, nargs='+', required=True,
        help='Time-varying categorical distribution files')
    parser.add_argument(
        '--output_files', type=Path, nargs='+', required=True,
        help='Files to save decoded indices')
    parser.add_argument(
        '--transition_file', type=Path,
        help='Categorical transition matrix file; defaults to uniform')
    parser.add_argument(
        '--initial_file', type=Path,
        help='Categorical initial distribution file; defaults to uniform')
    parser.add_argument(
        '--log_probs', action='store_true',
        help='Whether inputs are in (natural) log space')
    parser.add_argument(
        '--gpu', type=int, nargs='+',
        help='GPU index (or indices) to use for decoding. Required: there is no CPU path.')
    parser.add_argument(
        '--num_threads', type=int, default=1,
        help='Accepted for compatibility; ignored')
    parser.add_argument(
        '--config', type=Path,
        help='Python file overriding configuration constants (BATCH_SIZE, MIN_CHUNK_SIZE, ...)')
    # unknown flags are ignored like the reference's parse_known_args (__main__.py:50)
    args = vars(parser.parse_known_args(argv)[0])
    config = args.pop('config')
    if config is not None:
        torbi_b200.configure(config)
    if args['gpu'] is not None and len(args['gpu']) == 1:
        args['gpu'] = args['gpu'][0]
    return args


def main(argv=None):
    args = parse_args(argv)
    if len(args['input_files']) != len(args['output_files']):
        raise ValueError('--input_files and --output_files must have the same length')
    torbi_b200.from_files_to_files(**args)


if __name__ == '__main__':
    main()

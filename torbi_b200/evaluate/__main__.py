"""python -m torbi_b200.evaluate: /root/reference/torbi/evaluate/__main__.py:13-34 (argparse instead of yapecs)."""
import argparse
import json

import torbi_b200


def parse_args(argv=None):
    """Parse command-line arguments"""
    parser = argparse.ArgumentParser(description='Perform evaluation')
    parser.add_argument('--config', help='Python file overriding configuration constants')
    parser.add_argument('--datasets', nargs='+', default=None, help='The datasets to evaluate')
    parser.add_argument('--gpu', type=int, nargs='+', help='The index (or indices) of the gpu to use for evaluation')
    parser.add_argument('--num_threads', type=int, default=1, help='Ignored (CPU threads in the reference)')
    parser.add_argument('--root', help='Directory holding data/cache, assets and eval (default: TORBI_B200_EVAL_ROOT or ./torbi_eval)')
    parser.add_argument('--files', type=int, default=64, help='Files of a synthetic dataset that does not exist yet')
    args = parser.parse_args(argv)
    if args.config:
        torbi_b200.configure(args.config)
    gpu = args.gpu[0] if args.gpu and len(args.gpu) == 1 else args.gpu
    return dict(datasets=args.datasets or torbi_b200.DATASETS, gpu=gpu, num_threads=args.num_threads, root=args.root,
                synthetic_files=args.files)


if __name__ == '__main__':
    print(json.dumps(torbi_b200.evaluate.datasets(**parse_args()), indent=1))

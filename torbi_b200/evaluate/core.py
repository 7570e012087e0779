"""Evaluation runner: /root/reference/torbi/evaluate/core.py:14-123 on a synthetic corpus."""
import json
import os
import pathlib
import time
from typing import Dict, List, Optional, Sequence

import torch

import torbi_b200
from . import metrics as _metrics


def frames_to_seconds(frames: int) -> float:
    """penn.convert.frames_to_seconds with penn's constants (evaluate/core.py:115)"""
    return frames * _metrics.HOPSIZE / _metrics.SAMPLE_RATE


def _root(root=None) -> pathlib.Path:
    return pathlib.Path(root or os.environ.get('TORBI_B200_EVAL_ROOT') or 'torbi_eval')


def synthetic_dataset(name: str, root=None, files: int = 64, states: int = _metrics.PITCH_BINS,
                      median_frames: int = 430, seed: Optional[int] = None, dtype=torch.float32) -> List[str]:
    """Write a corpus of synthetic pitch posteriorgrams the way the reference's preprocessing leaves
    them (one (frames, states) log-probability tensor per stem under data/cache/<dataset>/, the stems
    listed in assets/partitions/<dataset>.json; evaluate/core.py:43-48): a pitch contour that drifts a
    few bins per frame, a peaked posterior around it in voiced stretches, a flat noisy one in unvoiced
    stretches.  Frame counts are log-normal around `median_frames`.  The true
    contour is kept beside each file (<stem>.truth.pt) for an accuracy-against-truth figure."""
    root = _root(root)
    cache = root / 'data' / 'cache' / name
    cache.mkdir(parents=True, exist_ok=True)
    (root / 'assets' / 'partitions').mkdir(parents=True, exist_ok=True)
    g = torch.Generator().manual_seed(torbi_b200.RANDOM_SEED if seed is None else seed)
    stems = []
    bins = torch.arange(states, dtype=torch.float32)
    for i in range(files):
        frames = int(torch.clamp(torch.exp(torch.randn((), generator=g) * 0.55) * median_frames, 50, 5000))
        steps = torch.randn(frames, generator=g) * 2.0
        contour = torch.cumsum(steps, 0) + float(torch.randint(states // 4, 3 * states // 4, (), generator=g))
        contour = contour.clamp(5, states - 6)
        voiced = (torch.rand(frames, generator=g) < 0.97).float()
        voiced = (torch.nn.functional.avg_pool1d(voiced[None, None], 9, 1, 4)[0, 0] > 0.8)   # unvoiced in stretches
        sharp = torch.where(voiced, 0.5, 1e-4)[:, None]
        logits = -sharp * (bins[None] - contour[:, None]) ** 2 + torch.randn(frames, states, generator=g)
        stem = f'{name}-{i:05d}'
        torch.save(torch.log_softmax(logits, dim=-1).to(dtype), cache / f'{stem}.pt')
        torch.save(contour.round().to(torch.int32), cache / f'{stem}.truth.pt')
        stems.append(stem)
    with open(root / 'assets' / 'partitions' / f'{name}.json', 'w') as file:
        json.dump(stems, file)
    return stems


def datasets(datasets: Optional[Sequence[str]] = None, gpu=None, num_threads: int = 1, root=None,
             synthetic_files: int = 64) -> Dict:
    """Evaluate Viterbi decoding methods (evaluate/core.py:14-123): per dataset, decode every stem of the
    partition with from_files_to_files under the current configuration (torbi_b200.CONFIG: chunking,
    batch size, ...), compare with the reference decode -- the un-chunked `torbi` configuration, i.e. the
    reference's COMPARE_WITH_REFERENCE = False branch (its librosa runner is not available offline) --
    and report raw pitch accuracy at PITCH_ERROR_THRESHOLDS plus the real-time factor (seconds of audio
    decoded per second).  A dataset whose partition file does not exist is synthesised first."""
    del num_threads
    root = _root(root)
    datasets = list(datasets or torbi_b200.DATASETS)
    transition_file = root / 'assets' / 'stats' / 'transition.pt'
    if not transition_file.exists():                          # evaluate/core.py:23-34
        transition_file.parent.mkdir(parents=True, exist_ok=True)
        torch.save(_metrics.pitch_transition(), transition_file)

    results = {}
    for dataset in datasets:
        partition = root / 'assets' / 'partitions' / f'{dataset}.json'
        if not partition.exists():
            synthetic_dataset(dataset, root, files=synthetic_files)
        with open(partition) as file:
            stems = json.load(file)
        input_files = [root / 'data' / 'cache' / dataset / f'{stem}.pt' for stem in stems]
        eval_dir = root / 'eval' / dataset

        def decode(config_name, chunk_size):
            output_files = [eval_dir / config_name / f'{stem}.pt' for stem in stems]
            (eval_dir / config_name).mkdir(parents=True, exist_ok=True)
            saved = torbi_b200.MIN_CHUNK_SIZE
            torbi_b200.MIN_CHUNK_SIZE = chunk_size
            torbi_b200.timing.reset()
            try:
                start = time.perf_counter()
                torbi_b200.from_files_to_files(
                    input_files, output_files, transition_file=transition_file, log_probs=True, gpu=gpu)
                wall = time.perf_counter() - start
            finally:
                torbi_b200.MIN_CHUNK_SIZE = saved
            return output_files, wall, torbi_b200.timing.results()

        # reference outputs: the un-chunked decode (evaluate/core.py:77-83)
        reference_files = [eval_dir / 'torbi' / f'{stem}.pt' for stem in stems]
        if torbi_b200.CONFIG != 'torbi' and not all(file.exists() for file in reference_files):
            decode('torbi', None)
        output_files, wall, timers = decode(torbi_b200.CONFIG, torbi_b200.MIN_CHUNK_SIZE)

        measured = _metrics.Metrics()
        truth = _metrics.Metrics()
        for stem, predicted_file, target_file in zip(stems, output_files, reference_files):
            predicted = torch.load(predicted_file, map_location='cpu')
            measured.update(predicted, torch.load(target_file, map_location='cpu'))
            truth_file = root / 'data' / 'cache' / dataset / f'{stem}.truth.pt'
            if truth_file.exists():
                truth.update(predicted, torch.load(truth_file))

        # speed as real-time factor, i.e. seconds decoded per second (evaluate/core.py:114-118)
        seconds = frames_to_seconds(measured.rpas[0].count)
        rtf = {key: seconds / value for key, value in timers.items() if value > 0}
        rtf['end_to_end'] = seconds / wall
        results[dataset] = measured() | {'rtf': rtf, 'rpa_vs_truth': truth()['rpa'], 'frames': measured.rpas[0].count,
                                         'files': len(stems)}
    (root / 'eval').mkdir(parents=True, exist_ok=True)
    with open(root / 'eval' / f'{torbi_b200.CONFIG}.json', 'w') as file:
        json.dump(results, file)
    return results

"""GPU (-m gpu): parity of the pieces added in round 2, all through the C ABI and against the CPU
oracle -- the cluster split-K kernel, fp16 / bf16 observations, forward-score checkpoints and the
segment-by-segment verification of a long 4096-state decode (config 5), the blocked backtrace,
config 3 at full size, the advisory cases of round 1 (banded hint on a very wide state space,
streamed host entry with S % 4 != 0) and a long-running stress of the copy rings."""
import numpy as np
import pytest
import torch

import cases
import torbi_b200
from torbi_b200 import _ffi

from test_gpu_parity import bits, cuda, gpu_decode, random_problem, full_size_problem, replay_path_score

pytestmark = pytest.mark.gpu


def last_kernel():
    return _ffi.library().torbi_b200_last_kernel().decode()


def oracle_segment(oracle, obs_segment, start_scores, A):
    """Run the CPU restatement over obs_segment[1:] starting from the forward scores `start_scores`
    (= P at the time of obs_segment[0], whose own values are not used): returns (scores at the end,
    backpointers (len, S) with row 0 unused).  P_0 = start + 0 exactly, then viterbi.cpp:78-108."""
    seg = np.array(obs_segment, dtype=np.float32, copy=True)
    seg[0] = start_scores
    S = seg.shape[1]
    _, post, trellis = oracle.c_decode(seg[None], [seg.shape[0]], A, np.zeros(S, np.float32), exact_scan=False,
                                       want_posterior=True, want_trellis=True)
    return post[0], trellis[0]


# ---------------------------------------------------------------------------
# cluster split-K kernel
# ---------------------------------------------------------------------------

@pytest.mark.parametrize('shape', [
    # (B, T, S, quant, transition): several interleaved batches, several groups, a group with more
    # batches than it interleaves at once, ragged tails that leave one batch running alone
    (64, 40, 1440, None, 'triangular'),
    (200, 30, 1440, 1, 'dense'),
    (33, 25, 1441, 2, 'dense'),
    (70, 60, 640, 2, 'dense'),
    (900, 6, 96, None, 'dense'),
    (800, 5, 1440, None, 'triangular'),
    (5, 90, 1000, 1, 'dense'),
    (40, 12, 20, None, 'dense'),
], ids=lambda s: f'B{s[0]}-T{s[1]}-S{s[2]}')
def test_cluster_kernel_matches_oracle(shape, oracle):
    B, T, S, quant, transition = shape
    obs, frames, A, init = random_problem(B * 77 + S, B, T, S, quant=quant, transition=transition)
    want_idx, want_post = oracle.c_decode(obs, frames, A, init, exact_scan=False, want_posterior=True)
    indices, posterior = gpu_decode(obs, frames, A, init, flags=_ffi.FORCE_CLUSTER)
    assert last_kernel() == 'forward_cluster'
    assert np.array_equal(indices, want_idx)
    assert np.array_equal(bits(posterior), bits(want_post))


def test_cluster_kernel_rejects_wide_state_spaces():
    obs, frames, A, init = random_problem(1, 2, 3, 1500, ragged=False)
    with pytest.raises(ValueError, match='cluster kernel does not support'):
        gpu_decode(obs, frames, A, init, flags=_ffi.FORCE_CLUSTER)


def test_strong_scaling_share_of_config2(oracle):
    """64 sequences x 2000 frames (the per-GPU share of config 2 on 8 GPUs) through whatever the
    selector picks and through the cluster kernel: identical, and two items equal the oracle."""
    B, T, S = 64, 2000, 1440
    obs, A, init = full_size_problem(B, T, S, seed=99)
    frames = torch.full((B,), T, dtype=torch.int32)
    auto, auto_post = torbi_b200.decode(obs, frames, A, init, flags=_ffi.EPS_FLOOR, return_posterior=True)
    kernel = last_kernel()
    cluster, cluster_post = torbi_b200.decode(obs, frames, A, init, flags=_ffi.EPS_FLOOR | _ffi.FORCE_CLUSTER,
                                              return_posterior=True)
    assert torch.equal(auto, cluster), kernel
    assert torch.equal(auto_post.view(torch.int32), cluster_post.view(torch.int32))
    A_host, init_host = A.cpu().numpy(), init.cpu().numpy()
    for b in (5, 63):
        want, want_post = oracle.c_decode(obs[b:b + 1].cpu().numpy(), [T], A_host, init_host, exact_scan=False,
                                          want_posterior=True)
        assert np.array_equal(cluster[b].cpu().numpy(), want[0])
        assert np.array_equal(bits(cluster_post[b].cpu().numpy()), bits(want_post[0]))


# ---------------------------------------------------------------------------
# fp16 / bf16 observations
# ---------------------------------------------------------------------------

@pytest.mark.parametrize('dtype', [torch.float16, torch.bfloat16])
@pytest.mark.parametrize('kernel,flags', [
    ('auto', 0), ('generic', _ffi.FORCE_GENERIC), ('wide', _ffi.FORCE_WIDE), ('tiled64', _ffi.TILE_ROWS_64),
    ('tiled16bp', _ffi.TILE_ROWS_16 | _ffi.BACKPOINTERS), ('banded', _ffi.BANDED), ('split', _ffi.FORCE_SPLIT),
    ('split_streamed', _ffi.FORCE_SPLIT | _ffi.SPLIT_STREAMED), ('cluster', _ffi.FORCE_CLUSTER), ('flow', _ffi.FORCE_FLOW)])
@pytest.mark.parametrize('states', [1440, 333])
def test_half_precision_observations_are_upcast_exactly(states, kernel, flags, dtype, oracle):
    """The reference casts the observation to fp32 on the host (core.py:191); here the 2-byte
    elements reach the kernel and are upcast in its observation load: same decode as the oracle
    on the upcast values.  S = 333 takes the scalar load path, S = 1440 the 4-element one."""
    obs, frames, A, init = random_problem(states + 5, 70, 14, states, quant=None)
    half = torch.from_numpy(obs).to(dtype)
    upcast = half.to(torch.float32).numpy()
    want_idx, want_post = oracle.c_decode(upcast, frames, A, init, exact_scan=False, want_posterior=True)
    indices, posterior = torbi_b200.decode(half.cuda(), torch.from_numpy(frames), cuda(A), cuda(init),
                                           flags=flags | _ffi.EPS_FLOOR, return_posterior=True)
    # (eps floor: the identity on these values, like in the fp32 tests)
    assert np.array_equal(indices.cpu().numpy(), want_idx)
    assert np.array_equal(bits(posterior.cpu().numpy()), bits(want_post))


@pytest.mark.parametrize('dtype', [torch.float16, torch.bfloat16])
def test_half_precision_through_the_python_api_and_the_host_entry(dtype, oracle):
    rng = np.random.default_rng(12)
    B, T, S = 24, 1200, 1440                                     # 83 MB as 2-byte elements: streamed in
    half = torch.from_numpy(cases.log_softmax(rng.standard_normal((B, T, S)))).to(dtype)
    frames = torch.from_numpy(rng.integers(1, T + 1, size=B).astype(np.int32))
    frames[2] = T
    A = torch.from_numpy(cases.log_triangular_transition(S))
    want = torbi_b200.from_probabilities(half.to(torch.float32), frames, A, None, log_probs=True, gpu=0).cpu()
    got = torbi_b200.from_probabilities(half, frames, A, None, log_probs=True, gpu=0).cpu()
    assert torch.equal(got, want)
    init = torch.full((S,), float(np.log(1.0 / S + np.finfo(np.float32).tiny)))
    host = torbi_b200.decode_host(half.pin_memory(), frames, A, init, gpu=0, flags=_ffi.EPS_FLOOR)
    assert torch.equal(host, want)
    # one item against the oracle
    b = 2
    ref = oracle.ref_from_probabilities(half[b:b + 1].to(torch.float32), None, A, None, True)
    assert np.array_equal(got[b].numpy(), ref[0])
    with pytest.raises(ValueError, match='float32, float16 or bfloat16'):
        torbi_b200.decode(half.to(torch.float64).cuda()[:1, :4], torch.tensor([4]), A.cuda(), init.cuda())


# ---------------------------------------------------------------------------
# checkpoints, long sequences, blocked backtrace
# ---------------------------------------------------------------------------

@pytest.mark.parametrize('kernel,flags', [('wide', _ffi.FORCE_WIDE), ('generic', _ffi.FORCE_GENERIC),
                                          ('tiled', _ffi.FORCE_TILED), ('split', _ffi.FORCE_SPLIT),
                                          ('cluster', _ffi.FORCE_CLUSTER), ('flow', _ffi.FORCE_FLOW)])
def test_checkpoints_hold_the_forward_scores(kernel, flags, oracle):
    obs, frames, A, init = random_problem(91, 5, 40, 257, quant=2)
    every = 7
    indices, checkpoints = torbi_b200.decode(cuda(obs), torch.from_numpy(frames), cuda(A), cuda(init), flags=flags,
                                             checkpoint_every=every)
    assert np.array_equal(indices.cpu().numpy(), oracle.c_decode(obs, frames, A, init))
    checkpoints = checkpoints.cpu().numpy()
    assert checkpoints.shape == (5, (40 - 1) // every, 257)
    for b in range(5):
        for k in range(checkpoints.shape[1]):
            t = (k + 1) * every
            if t >= frames[b]:
                assert np.isnan(checkpoints[b, k]).all()              # untouched
                continue
            _, want = oracle.c_decode(obs[b:b + 1, :t + 1], [t + 1], A, init, want_posterior=True)
            assert np.array_equal(bits(checkpoints[b, k]), bits(want[0])), (b, k)
    with pytest.raises(ValueError, match='backpointer mode'):
        torbi_b200.decode(cuda(obs), torch.from_numpy(frames), cuda(A), cuda(init),
                          flags=_ffi.FORCE_TILED | _ffi.BACKPOINTERS, checkpoint_every=every)


@pytest.mark.parametrize('shape', [(1, 40000, 256), (3, 20000, 333), (1, 17000, 1440)],
                         ids=lambda s: f'B{s[0]}-T{s[1]}-S{s[2]}')
def test_blocked_backtrace_equals_the_plain_chase(shape, oracle):
    """Sequences of >= 16384 frames with stored backpointers take the blocked backtrace (composed
    backpointer maps per block, then all blocks in parallel): same indices as the oracle's chase."""
    B, T, S = shape
    obs, frames, A, init = random_problem(T + S, B, T, S, quant=2, transition='dense')
    want = oracle.c_decode(obs, frames, A, init, exact_scan=False)
    for flags in (_ffi.FORCE_WIDE, _ffi.FORCE_GENERIC, _ffi.FORCE_FLOW):
        indices, _ = gpu_decode(obs, frames, A, init, flags=flags)
        assert _ffi.library().torbi_b200_last_backtrace().decode() == 'backtrace_blocked'
        assert np.array_equal(indices, want)


def test_config5_long_sequence_verified_in_segments(oracle):
    """BASELINE config 5 scaled in time only: ONE 4096-state sequence, dense transition, 200 000 frames
    through the wide kernel (what `auto` selects).  The reference cannot index this shape at full
    length (t * states overflows int at t >= 524 288, viterbi.cpp:99) and a CPU decode of the whole
    run would take hours, so the run is verified the way SURVEY H7 lays out: the GPU hands out its
    forward scores every 400 frames; the CPU restatement, started from the GPU's OWN checkpoint k,
    must reproduce checkpoint k + 1 bit for bit and the decoded path must follow the restatement's
    backpointers through the segment.  Segments are independent, so a sample of them is checked
    (first, last, and some in between); plus the whole prefix against the unmodified reference op."""
    S, T, every = 4096, 200_000, 400
    g = torch.Generator(device='cuda:0').manual_seed(55)
    obs = torch.empty((1, T, S), device='cuda:0')
    for start in range(0, T, 20_000):
        obs[0, start:start + 20_000] = torch.log_softmax(
            torch.randn(20_000, S, device='cuda:0', generator=g), dim=-1)
    A = torch.log_softmax(torch.randn(S, S, device='cuda:0', generator=g), dim=1)
    init = torch.full((S,), float(np.log(1.0 / S + np.finfo(np.float32).tiny)), device='cuda:0')
    frames = torch.tensor([T], dtype=torch.int32)
    indices, posterior, checkpoints = torbi_b200.decode(obs, frames, A, init, return_posterior=True,
                                                        checkpoint_every=every)
    assert last_kernel() in ('forward_wide', 'forward_flow')
    assert _ffi.library().torbi_b200_last_backtrace().decode() == 'backtrace_blocked'
    path = indices[0].cpu().numpy()
    A_host = A.cpu().numpy()
    rows = checkpoints.shape[1]
    assert rows == (T - 1) // every and not torch.isnan(checkpoints).any()
    assert int(path[-1]) == int(posterior[0].argmax())            # first-occurrence final argmax (viterbi.cpp:218)

    def check_segment(t0, t1, start, end_want):
        seg = obs[0, t0:t1 + 1].cpu().numpy()
        end, bp = oracle_segment(oracle, seg, start, A_host)
        assert np.array_equal(bits(end), bits(end_want)), (t0, t1)
        for t in range(t1, t0, -1):                               # viterbi.cpp:140-160 along the GPU's path
            assert path[t - 1] == bp[t - t0, path[t]], (t0, t1, t)

    # prefix: frames 0 .. every from the initial distribution
    first = obs[0, :every + 1].cpu().numpy()
    _, want = oracle.c_decode(first[None], [every + 1], A_host, init.cpu().numpy(), exact_scan=False, want_posterior=True)
    assert np.array_equal(bits(checkpoints[0, 0].cpu().numpy()), bits(want[0]))
    for k in (0, 123, 250, 377, rows - 2):                        # checkpoint k is at t = (k + 1) * every
        check_segment((k + 1) * every, (k + 2) * every, checkpoints[0, k].cpu().numpy(),
                      checkpoints[0, k + 1].cpu().numpy())
    # the tail: last checkpoint to the final scores
    check_segment(rows * every, T - 1, checkpoints[0, rows - 1].cpu().numpy(), posterior[0].cpu().numpy())
    # the unmodified reference operator on a prefix it can index (its own ties, its own backtrace):
    # the prefix of the long path agrees wherever the prefix's own best path has merged with it
    if oracle.ref_available():
        n = 600
        ref = oracle.ref_decode(obs[:, :n].cpu().numpy(), [n], A_host, init.cpu().numpy())
        assert np.array_equal(ref[0][:n // 4], path[:n // 4])     # (best paths coalesce going backwards)


# ---------------------------------------------------------------------------
# config 3 at full size
# ---------------------------------------------------------------------------

_CONFIG3_ORACLE = {}


@pytest.mark.parametrize('banded', [False, True])
def test_config3_full_size(banded, oracle):
    """BASELINE config 3 as specified: S=1440, batch 256, lengths log-uniform in [200, 20000] with both
    ends present; whatever kernels `auto` selects (dense) and the banded path.  Per-item oracle check
    of six items including the longest and the shortest (the reference cannot index the padded
    batch: B*T*S >= 2^31)."""
    B, S = 256, 1440
    rng = np.random.default_rng(1234)
    frames = np.exp(rng.uniform(np.log(200), np.log(20000), size=B)).round().astype(np.int32)
    frames[0], frames[1] = 20000, 200
    T = int(frames.max())
    g = torch.Generator(device='cuda:0').manual_seed(77)
    obs = torch.empty((B, T, S), device='cuda:0')
    for start in range(0, B, 16):
        obs[start:start + 16] = torch.log_softmax(torch.randn(16, T, S, device='cuda:0', generator=g), dim=-1)
    A = cuda(cases.log_triangular_transition(S))
    init = torch.full((S,), float(np.log(1.0 / S + np.finfo(np.float32).tiny)), device='cuda:0')
    indices, posterior = torbi_b200.decode(obs, torch.from_numpy(frames), A, init,
                                           flags=_ffi.BANDED if banded else 0, return_posterior=True)
    kernel = last_kernel()
    assert ('banded' in kernel) == banded
    idx = indices.cpu().numpy()
    A_host, init_host = A.cpu().numpy(), init.cpu().numpy()
    for b in range(B):
        assert (idx[b, frames[b] - 1:] == idx[b, frames[b] - 1]).all()       # tail fill (viterbi.cpp:219-221)
    order = np.argsort(frames)
    for b in (0, 1, int(order[B // 2]), int(order[B // 4]), int(order[3]), int(order[-3])):
        f = int(frames[b])
        if b not in _CONFIG3_ORACLE:                              # (same seeded inputs for both parametrisations)
            _CONFIG3_ORACLE[b] = oracle.c_decode(obs[b:b + 1, :f].cpu().numpy(), [f], A_host, init_host,
                                                 exact_scan=False, want_posterior=True)
        want, want_post = _CONFIG3_ORACLE[b]
        assert np.array_equal(idx[b, :f], want[0]), (b, f, kernel)
        assert np.array_equal(bits(posterior[b].cpu().numpy()), bits(want_post[0])), (b, f, kernel)
    del obs, indices
    torch.cuda.empty_cache()


# ---------------------------------------------------------------------------
# advisory cases of round 1
# ---------------------------------------------------------------------------

def test_banded_hint_on_a_state_space_too_wide_for_the_tiled_kernel(oracle):
    """transition=None (all one constant -> the banded hint) with batch >= 32 and S > 80 x #SM: the
    hint is dropped and the wide / generic kernels decode it (was: ValueError)."""
    rng = np.random.default_rng(3)
    B, T, S = 32, 3, 12000
    obs = torch.from_numpy(cases.log_softmax(rng.standard_normal((B, T, S))))
    got = torbi_b200.from_probabilities(obs, None, None, None, log_probs=True, gpu=0)
    assert last_kernel() in ('forward_wide', 'forward_generic', 'forward_flow')
    want = oracle.ref_from_probabilities(obs, None, None, None, True)
    assert np.array_equal(got.cpu().numpy(), want)
    lib = _ffi.library()
    frames = (np.full(B, T, np.int32))
    assert lib.torbi_b200_plan(frames.ctypes.data, B, T, S, _ffi.BANDED, 148) == lib.torbi_b200_plan(
        frames.ctypes.data, B, T, S, 0, 148)


@pytest.mark.parametrize('states', [50, 333])
def test_streamed_host_entry_with_unaligned_rows(states, oracle):
    """Pinned batch >= 64 MB with S % 4 != 0 (and S <= 80: one CTA reads across frame boundaries):
    the observation loads must be coherent with the copy still in flight."""
    rng = np.random.default_rng(states)
    B = 200 if states == 50 else 60
    T = 2000 if states == 50 else 1000
    obs = torch.from_numpy(cases.log_softmax(rng.standard_normal((B, T, states)))).pin_memory()
    assert obs.numel() * 4 >= 64 << 20
    frames = torch.from_numpy(rng.integers(T // 2, T + 1, size=B).astype(np.int32))
    A = torch.from_numpy(cases.log_softmax(rng.standard_normal((states, states))))
    init = torch.from_numpy(cases.log_softmax(rng.standard_normal((states,))))
    want = oracle.c_decode(obs.numpy(), frames.numpy(), A.numpy(), init.numpy(), exact_scan=False)
    # a first decode of different data leaves stale observations in the arena
    torbi_b200.decode_host((obs * 0.5).pin_memory(), frames, A, init, gpu=0)
    for _ in range(3):
        got = torbi_b200.decode_host(obs, frames, A, init, gpu=0)
        assert np.array_equal(got.numpy(), want)


def test_failed_host_call_leaves_the_arena_usable(oracle):
    obs, frames, A, init = random_problem(5, 6, 20, 12000 // 100, ragged=True)
    t = [torch.from_numpy(x) for x in (obs, frames, A, init)]
    wide = torch.from_numpy(cases.log_softmax(np.random.default_rng(1).standard_normal((2, 3, 12000))))
    with pytest.raises(ValueError, match='tiled kernel does not support'):
        torbi_b200.decode_host(wide.pin_memory(), torch.tensor([3, 3]), torch.zeros(12000, 12000),
                               torch.zeros(12000), gpu=0, flags=_ffi.FORCE_TILED)
    got = torbi_b200.decode_host(t[0], t[1], t[2], t[3], gpu=0)
    assert np.array_equal(got.numpy(), oracle.c_decode(obs, frames, A, init))


# ---------------------------------------------------------------------------
# long-running stress of the cp.async / mbarrier rings
# ---------------------------------------------------------------------------

@pytest.mark.parametrize('kernel,flags', [
    ('tiled16', _ffi.TILE_ROWS_16), ('tiled32', _ffi.TILE_ROWS_32), ('tiled16bp', _ffi.TILE_ROWS_16 | _ffi.BACKPOINTERS),
    ('banded16', _ffi.TILE_ROWS_16 | _ffi.BANDED), ('split', _ffi.FORCE_SPLIT),
    ('split_streamed', _ffi.FORCE_SPLIT | _ffi.SPLIT_STREAMED), ('cluster', _ffi.FORCE_CLUSTER), ('wide', _ffi.FORCE_WIDE),
    ('flow', _ffi.FORCE_FLOW)])
def test_ring_stress_many_steps_small_tiles(kernel, flags, oracle, ring_stress_problem):
    """6000 timesteps x small tiles x quantised inputs (exact ties everywhere): tens of thousands of
    ring stages with the warps free to drift apart; any lost or early copy changes a score or a tie."""
    obs, frames, A, init, want_idx, want_post = ring_stress_problem
    indices, posterior = gpu_decode(obs, frames, A, init, flags=flags)
    assert np.array_equal(indices, want_idx)
    assert np.array_equal(bits(posterior), bits(want_post))


@pytest.fixture(scope='module')
def ring_stress_problem(oracle):
    obs, frames, A, init = random_problem(4242, 32, 6000, 480, quant=2, transition='dense')
    frames[5] = 5999
    want_idx, want_post = oracle.c_decode(obs, frames, A, init, exact_scan=False, want_posterior=True)
    return obs, frames, A, init, want_idx, want_post


# ---------------------------------------------------------------------------
# packed (unpadded) layout
# ---------------------------------------------------------------------------

def _packed_problem(seed, lengths, S, gap_rows=(0, 3, 0, 1), transition='dense', quant=2):
    """Items of the given lengths laid out in one (rows, S) buffer in a shuffled order with gaps."""
    B, T = len(lengths), int(max(lengths))
    obs, frames, A, init = random_problem(seed, B, T, S, quant=quant, transition=transition)
    frames = np.asarray(lengths, np.int32)
    rng = np.random.default_rng(seed)
    placement = rng.permutation(B)
    item_row = np.zeros(B, np.int64)
    row = 0
    for k, b in enumerate(placement):
        row += gap_rows[k % len(gap_rows)]
        item_row[b] = row
        row += int(frames[b])
    rows = row + 2
    packed = np.full((rows, S), np.nan, np.float32)              # gaps hold NaN: they must never be read
    for b in range(B):
        packed[item_row[b]:item_row[b] + frames[b]] = obs[b, :frames[b]]
    return obs, frames, A, init, packed, item_row, rows


@pytest.mark.parametrize('kernel,flags', [
    ('auto', 0), ('generic', _ffi.FORCE_GENERIC), ('wide', _ffi.FORCE_WIDE), ('tiled64', _ffi.TILE_ROWS_64),
    ('tiled16bp', _ffi.TILE_ROWS_16 | _ffi.BACKPOINTERS), ('banded', _ffi.BANDED | _ffi.TILE_ROWS_32),
    ('split', _ffi.FORCE_SPLIT), ('split_streamed', _ffi.FORCE_SPLIT | _ffi.SPLIT_STREAMED), ('cluster', _ffi.FORCE_CLUSTER),
    ('flow', _ffi.FORCE_FLOW)])
@pytest.mark.parametrize('states', [1440, 333])
def test_packed_layout_matches_oracle(states, kernel, flags, oracle):
    """torbi_b200_decode_packed: ragged items anywhere in one (rows, S) buffer, every kernel; the
    states of item b come back in its own rows, rows between items are left alone."""
    rng = np.random.default_rng(states)
    lengths = rng.integers(1, 40, size=70)
    lengths[3], lengths[11] = 1, 39
    obs, frames, A, init, packed, item_row, rows = _packed_problem(states + 1, lengths, states)
    want, want_post = oracle.c_decode(obs, frames, A, init, exact_scan=False, want_posterior=True)
    out = torch.full((rows,), -7, dtype=torch.int32, device='cuda:0')
    got, post = torbi_b200.decode_packed(cuda(packed), torch.from_numpy(item_row), torch.from_numpy(frames), cuda(A), cuda(init),
                                         flags=flags, return_posterior=True, out=out)
    got = got.cpu().numpy()
    covered = np.zeros(rows, bool)
    for b in range(len(lengths)):
        lo, hi = item_row[b], item_row[b] + frames[b]
        assert np.array_equal(got[lo:hi], want[b, :frames[b]]), (kernel, b)
        covered[lo:hi] = True
    assert (got[~covered] == -7).all()
    assert np.array_equal(bits(post.cpu().numpy()), bits(want_post))


def test_packed_layout_long_item_uses_the_blocked_backtrace(oracle):
    lengths = [17000, 300, 5]
    obs, frames, A, init, packed, item_row, rows = _packed_problem(8, lengths, 256, transition='dense')
    want = oracle.c_decode(obs, frames, A, init, exact_scan=False)
    got = torbi_b200.decode_packed(cuda(packed), torch.from_numpy(item_row), torch.from_numpy(frames), cuda(A), cuda(init),
                                   flags=_ffi.FORCE_WIDE).cpu().numpy()
    assert _ffi.library().torbi_b200_last_backtrace().decode() == 'backtrace_blocked'
    for b in range(3):
        assert np.array_equal(got[item_row[b]:item_row[b] + frames[b]], want[b, :frames[b]])


def test_packed_host_entry_and_validation(oracle):
    lengths = [30, 12, 1, 25, 30, 7]
    obs, frames, A, init, packed, item_row, rows = _packed_problem(21, lengths, 97, quant=None)
    want = oracle.c_decode(obs, frames, A, init, exact_scan=False)
    t = lambda x: torch.from_numpy(x)
    out = torch.full((rows,), -3, dtype=torch.int32)
    got = torbi_b200.decode_host_packed(t(np.nan_to_num(packed)), t(item_row), t(frames), t(A), t(init), gpu=0, out=out).numpy()
    for b in range(len(lengths)):
        assert np.array_equal(got[item_row[b]:item_row[b] + frames[b]], want[b, :frames[b]])
    bad = item_row.copy()
    bad[0] = bad[1]                                              # two items on the same rows
    with pytest.raises(ValueError, match='overlap'):
        torbi_b200.decode_packed(cuda(packed), t(bad), t(frames), cuda(A), cuda(init))
    bad = item_row.copy()
    bad[int(np.argmax(item_row))] = rows                         # runs off the end
    with pytest.raises(ValueError, match='leave the buffer'):
        torbi_b200.decode_packed(cuda(packed), t(bad), t(frames), cuda(A), cuda(init))


# ---------------------------------------------------------------------------
# flow kernel (tagged 8-byte score exchange through L2, no grid barrier)
# ---------------------------------------------------------------------------

@pytest.mark.parametrize('shape', [
    # (B, T, S, quant, transition): every register / shared-memory / streamed residency split of the
    # transition rows, S % 4 != 0, fewer states than CTAs, several sequences back to back (buffer sets
    # alternate, tags keep counting), one-frame items between long ones
    (1, 300, 1440, 1, 'triangular'), (1, 120, 1441, 2, 'dense'), (3, 90, 4096, 2, 'dense'), (1, 60, 4099, None, 'dense'),
    (5, 70, 333, 2, 'dense'), (4, 50, 7, 1, 'dense'), (2, 40, 96, None, 'dense'), (2, 25, 6000, 2, 'dense'),
    (1, 12, 9400, 1, 'dense'), (6, 200, 2500, 2, 'dense'), (1, 2000, 640, 2, 'dense'),
], ids=lambda s: f'B{s[0]}-T{s[1]}-S{s[2]}')
def test_flow_kernel_matches_oracle(shape, oracle):
    B, T, S, quant, transition = shape
    obs, frames, A, init = random_problem(B * 31 + S, B, T, S, quant=quant, transition=transition)
    if B >= 3:
        frames[1] = 1                                            # no recursion step at all
        frames[2] = T
    want_idx, want_post = oracle.c_decode(obs, frames, A, init, exact_scan=False, want_posterior=True)
    indices, posterior = gpu_decode(obs, frames, A, init, flags=_ffi.FORCE_FLOW)
    assert last_kernel() == 'forward_flow'
    assert np.array_equal(indices, want_idx)
    assert np.array_equal(bits(posterior), bits(want_post))
    # twice more on the same workspace pattern: stale tags of an earlier call must never match
    for _ in range(2):
        again, _ = gpu_decode(obs, frames, A, init, flags=_ffi.FORCE_FLOW)
        assert np.array_equal(again, want_idx)


def test_flow_kernel_runs_of_one_frame_sequences(oracle):
    """Sequences follow each other without a grid barrier and two buffer sets alternate.  Runs of
    one-frame items (no recursion step, hence no CTA barrier of the step loop) between longer ones are
    the case where a fast CTA could reuse a set while a slow one still reads the final scores in it."""
    B, T, S = 60, 9, 1440
    obs, frames, A, init = random_problem(777, B, T, S, quant=2)
    frames[:] = 1
    frames[::7] = [9, 3, 2, 8, 1, 5, 2, 3, 1][:len(frames[::7])]
    want_idx, want_post = oracle.c_decode(obs, frames, A, init, exact_scan=False, want_posterior=True)
    for _ in range(5):
        indices, posterior = gpu_decode(obs, frames, A, init, flags=_ffi.FORCE_FLOW)
        assert np.array_equal(indices, want_idx)
        assert np.array_equal(bits(posterior), bits(want_post))


def test_flow_kernel_is_what_single_sequences_get(oracle):
    """BASELINE config 1 shape (S=1440, B=1, T=2000, uniform transition / initial) through the selector."""
    rng = np.random.default_rng(1)
    obs = torch.from_numpy(cases.log_softmax(rng.standard_normal((1, 2000, 1440))))
    got = torbi_b200.from_probabilities(obs, None, None, None, log_probs=True, gpu=0)
    assert last_kernel() == 'forward_flow'
    want = oracle.ref_from_probabilities(obs, None, None, None, True)
    assert np.array_equal(got.cpu().numpy(), want)


def test_flow_kernel_rejects_very_wide_state_spaces():
    obs, frames, A, init = random_problem(1, 1, 2, 12000, ragged=False)
    with pytest.raises(ValueError, match='flow kernel does not support'):
        gpu_decode(obs, frames, A, init, flags=_ffi.FORCE_FLOW)

"""CPU statements of the three post-filter algorithms the short-list kernels rest on (csrc/cldet_detect.cu), each checked
against the plain definition it replaces.  They mirror the kernels step by step (same invariants, same order of updates), so
the equivalences DESIGN.md claims -- "this IS the pairwise rank", "identical to the sequential scan" -- are pinned without a GPU;
the -m gpu tests then compare the kernels themselves with torchvision / the oracle.

  * bucket rank            rank_sort_kernel, `n <= bucket_n` branch
  * parallel-round chain   nms_fused_kernel / nms_resolve_stream_kernel, one 64-box chunk
  * pipelined resolve      fold of the next column block by the chain warp + absorb one chunk behind
"""
import numpy as np
import pytest

MASK64 = (1 << 64) - 1


def bucket_rank(keys, buckets=2048):
    """rank_i = #{l : key_l > key_i} through the bucket pass: bucket = (key - min) >> shift, shift so that the range spans
    `buckets` buckets; rank = keys in higher buckets + larger keys of the own bucket."""
    ks = [int(k) for k in keys]
    lo, hi = min(ks), max(ks)
    rng = hi - lo
    shift = 0 if rng < buckets else rng.bit_length() - 11
    b = [(k - lo) >> shift for k in ks]
    assert max(b) < buckets
    cnt = np.bincount(b, minlength=buckets)
    start = np.concatenate([np.cumsum(cnt[::-1])[::-1][1:], [0]])          # keys in HIGHER buckets
    members = {}
    for i, bi in enumerate(b):
        members.setdefault(bi, []).append(ks[i])
    return np.array([start[b[i]] + sum(1 for x in members[b[i]] if x > ks[i]) for i in range(len(ks))])


@pytest.mark.parametrize('n', [1, 2, 5, 100, 1270, 2048])
@pytest.mark.parametrize('kind', ['spread', 'all_equal_scores', 'few_scores', 'one_outlier'])
def test_bucket_rank_is_the_pairwise_rank(n, kind):
    rng = np.random.default_rng(n + len(kind))
    if kind == 'spread':
        hi = rng.integers(0xBD4CCCCD, 0xBF800000, n, dtype=np.uint64)
    elif kind == 'all_equal_scores':
        hi = np.full(n, 0xBF400000, np.uint64)
    elif kind == 'few_scores':
        hi = rng.choice(np.array([0xBD000000, 0xBE000000, 0xBF000000], np.uint64), n)
    else:
        hi = np.full(n, 0xBD800000, np.uint64)
        hi[0] = 0xBF7FFFFF
    low = np.uint64(0xFFFFFFFF) - rng.permutation(300000)[:n].astype(np.uint64)          # ~anchor: distinct per image
    keys = (hi << np.uint64(32)) | low
    want = np.argsort(np.argsort(np.uint64(MASK64) - keys, kind='stable'), kind='stable')
    assert np.array_equal(bucket_rank(keys), want)


def greedy_chunk(diag, alive):
    """The sequential scan over one 64-box chunk: box b survives iff its bit is clear when the scan reaches it."""
    kept, rem = 0, ~alive & MASK64
    for b in range(64):
        if not (rem >> b) & 1:
            kept |= 1 << b
            rem |= diag[b]
    return kept


def rounds_chunk(diag, alive, max_rounds=8):
    """The chain as shipped: an undecided box that NO undecided box could suppress is kept; the boxes the newly kept ones
    suppress leave the undecided set; box by box after max_rounds.  Returns (kept, rounds used, fell back)."""
    und, kept, r = alive, 0, 0
    while und:
        r += 1
        if r > max_rounds:
            while und:
                b = (und & -und).bit_length() - 1
                kept |= 1 << b
                und &= ~((1 << b) | diag[b])
            return kept, r, True
        threat = 0
        for b in range(64):
            if (und >> b) & 1:
                threat |= diag[b]
        if not threat & und:
            return kept | und, r, False
        fresh = und & ~threat
        assert fresh                                   # the first undecided box is never threatened
        gone = 0
        for b in range(64):
            if (fresh >> b) & 1:
                gone |= diag[b]
        kept |= fresh
        und &= ~(fresh | gone)
    return kept, r, False


def _diag_from(mask64):
    return [sum(1 << int(c) for c in np.nonzero(mask64[b])[0]) for b in range(64)]


@pytest.mark.parametrize('density', [0.0, 0.005, 0.03, 0.1, 0.5, 0.9])
def test_parallel_round_chain_equals_the_sequential_scan(density):
    rng = np.random.default_rng(int(density * 1000))
    for _ in range(200):
        diag = _diag_from(np.triu(rng.random((64, 64)) < density, 1))
        alive = int(rng.integers(0, 1 << 63)) * 2 + int(rng.integers(0, 2))
        assert rounds_chunk(diag, alive)[0] == greedy_chunk(diag, alive)


def test_parallel_round_chain_falls_back_on_a_long_dependency_chain():
    diag = [(1 << (b + 1)) if b < 63 else 0 for b in range(64)]          # box b suppresses only box b+1: depth 64
    kept, _, fell_back = rounds_chunk(diag, MASK64)
    assert fell_back and kept == greedy_chunk(diag, MASK64) == int('01' * 32, 2)


def greedy_nms(m, n):
    removed, keep = np.zeros(n, bool), []
    for i in range(n):
        if not removed[i]:
            keep.append(i)
            removed |= m[i]
    return keep


def pipelined_resolve(m, n):
    """The chunk loop of the resolve kernels: the chain folds chunk c's kept rows into column word c+1 itself, the absorb of
    chunk c's kept rows into words c+2.. runs during chunk c+1; removed[c] must be complete when the chain reads it."""
    cb = (n + 63) // 64
    removed = [0] * (cb + 1)
    kept_s = [0, 0]
    keep = []

    def word(row, w):
        c0 = w * 64
        return sum(1 << int(t) for t in np.nonzero(m[row, c0:min(n, c0 + 64)])[0])

    for c in range(cb):
        buf = c & 1
        late = []
        if c >= 1 and c + 1 < cb:                                        # absorbers: chunk c-1 into words c+1..
            for wd in range(c + 1, cb):
                v = 0
                for b in range(64):
                    if (kept_s[buf ^ 1] >> b) & 1:
                        v |= word((c - 1) * 64 + b, wd)
                late.append((wd, v))
        rows = min(64, n - c * 64)
        alive = ~(removed[c] | (MASK64 << rows)) & MASK64
        diag = [word(c * 64 + b, c) if b < rows else 0 for b in range(64)]
        kept = rounds_chunk(diag, alive)[0]
        if c + 1 < cb:
            for b in range(64):
                if (kept >> b) & 1:
                    removed[c + 1] |= word(c * 64 + b, c + 1)
        kept_s[buf] = kept
        keep += [c * 64 + b for b in range(64) if (kept >> b) & 1]
        for wd, v in late:                                               # lands before the barrier that ends iteration c
            removed[wd] |= v
    return keep


@pytest.mark.parametrize('n', [1, 63, 64, 65, 129, 500, 1024])
@pytest.mark.parametrize('density', [0.0005, 0.005, 0.05])
def test_pipelined_resolve_equals_greedy_nms(n, density):
    rng = np.random.default_rng(n)
    m = np.triu(rng.random((n, n)) < density, 1)
    assert pipelined_resolve(m, n) == greedy_nms(m, n)

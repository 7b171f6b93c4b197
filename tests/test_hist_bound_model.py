"""Host-side model of the per-row histogram threshold sharing (ccr_tc.cu:
``seed_tau_kernel`` bucket parameters, ``hist_count`` on the best item of every 8-column group that
has a hit, ``hist_bound`` = lower edge of the highest bucket with >= k items counted at or above it).

The model replays a row's score stream through several concurrent "streams" in tile order and checks
the two properties the kernel relies on, for every evaluation point:
  * validity  -- the bound never exceeds the row's true k-th best score (so no top-k item is ever
    filtered out), including with masked items counted in include mode (k + h);
  * usefulness -- once a few tiles have been seen the bound is far tighter than the seed.
No GPU involved: this pins the arithmetic (ord32 bucketing, clamping, the k + h rule)."""
import numpy as np
import pytest

BINS = 128


def ord32(x):
    b = np.asarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    return np.where(b & 0x80000000, (~b) & 0xFFFFFFFF, b | 0x80000000).astype(np.uint64)


def unord32(o):
    bits = (o & 0x7FFFFFFF) if (o & 0x80000000) else (~o & 0xFFFFFFFF)
    return np.array([bits], dtype=np.uint32).view(np.float32)[0]


def seed_params(sample, kth):
    """seed_tau_kernel: base = kth best sampled ord, buckets of width 2^shift up to max + 1/8 span."""
    o = np.sort(ord32(sample))[::-1]
    base, top = int(o[kth - 1]), int(o[0])
    span = top - base
    want = span + (span >> 3) + 1
    shift = 0
    while (BINS << shift) < want:
        shift += 1
    return base, shift


def bound_from_hist(hist, need, base, shift):
    acc = 0
    for b in range(BINS - 1, -1, -1):
        acc += int(hist[b])
        if acc >= need:
            return base + (b << shift)
    return 0


@pytest.mark.parametrize("dist,k,h", [("normal", 100, 0), ("normal", 1001, 0), ("normal", 10, 37),
                                      ("ties", 50, 0), ("heavy_tail", 100, 5)])
def test_histogram_bound_is_valid_and_tightens(dist, k, h):
    rs = np.random.RandomState(k + h)
    n_streams, tiles, cols = 6, 400, 128          # a row seen by 6 concurrent streams of 400 tiles x 128 columns
    n = n_streams * tiles * cols
    if dist == "normal":
        scores = (rs.standard_normal(n) * 27.7).astype(np.float32)
    elif dist == "ties":
        scores = rs.randint(0, 200, size=n).astype(np.float32)
    else:
        scores = (rs.standard_t(3, size=n) * 10).astype(np.float32)
    need = k + h                                   # include mode: k plus the row's mask entries
    kth_true = np.sort(scores)[::-1][need - 1]     # every bound must stay <= the (k+h)-th best ...
    sample = scores[:: max(1, n // 4096)]
    base, shift = seed_params(sample, min(need, len(sample)))
    tau = np.float32(np.sort(sample)[::-1][min(need, len(sample)) - 1])   # seed threshold (value)
    assert ord32(tau) == base
    hist = np.zeros(BINS, dtype=np.int64)
    streams = scores.reshape(n_streams, tiles, cols)
    seed_rate = np.mean(scores >= tau)
    final_bound = 0
    for t in range(tiles):                          # streams advance in lock step, like concurrent CTAs
        for s in range(n_streams):
            groups = streams[s, t].reshape(-1, 8).max(axis=1)      # best item of every 8-column group
            for m in groups[groups >= tau]:
                b = min(int((int(ord32(m)) - base) >> shift), BINS - 1)
                hist[b] += 1
        ti = t + 1
        if ti >= 4 and ((ti & (ti - 1)) == 0 or (ti & 255) == 0):
            bound = bound_from_hist(hist, need, base, shift)
            if bound:
                assert bound <= int(ord32(kth_true)), (dist, ti, bound, int(ord32(kth_true)))   # validity
                if bound > int(ord32(tau)):
                    tau = unord32(bound)            # threads adopt the tighter threshold
                final_bound = max(final_bound, bound)
    assert final_bound > 0
    final_rate = np.mean(ord32(scores) >= final_bound)
    if dist != "ties":
        assert final_rate < seed_rate / 3 or final_rate <= 4.0 * need / n, (seed_rate, final_rate)   # usefulness
    # nothing of the true top-(k+h) was ever below a threshold in force
    assert (ord32(np.sort(scores)[::-1][:need]) >= final_bound).all()

"""Host-side segment logic for `-S <secs>` mode and its multi-GPU sharding.

Mirrors the reference's model-free split search (find_split_point, qwen_asr.c:617-643, and the
split loop of qwen_transcribe_audio, qwen_asr.c:941-970): segments are independent units
(default --past-text no), so ranks take disjoint blocks of segments with no data-path
collective; rank 0 only gathers the per-segment token ids in segment order.
"""
import numpy as np

SAMPLE_RATE = 16000
ENERGY_WINDOW_MS = 100
REFERENCE_MAX_SPLITS = 127  # qwen_asr.c:958,968 (splits[128], `if (n_splits >= 127) break`)


def find_split_point(samples, target_sample, search_sec):
    """Centre of the lowest-energy 100 ms window within +-search_sec of target (qwen_asr.c:617-643)."""
    n = len(samples)
    half = int(search_sec * SAMPLE_RATE)
    lo = max(0, target_sample - half)
    hi = min(n, target_sample + half)
    win = (ENERGY_WINDOW_MS * SAMPLE_RATE) // 1000  # 1600
    if hi - lo < win:
        return target_sample
    # windows start at lo, lo + win/2, ... while pos + win <= hi; the first window with the strictly lowest mean energy wins
    sq = samples[lo:hi].astype(np.float32) ** 2
    starts = np.arange(0, hi - lo - win + 1, win // 2)
    view = np.lib.stride_tricks.sliding_window_view(sq, win)[starts]
    energy = view.sum(axis=1, dtype=np.float32) / np.float32(win)
    best = int(np.argmin(energy))
    if not energy[best] < np.float32(1e30):
        return target_sample
    return lo + int(starts[best]) + win // 2


def split_segments(samples, segment_sec, search_sec=3.0, max_splits=REFERENCE_MAX_SPLITS):
    """[(start, end)] sample ranges exactly as the reference cuts them (qwen_asr.c:941-970).

    max_splits=None lifts the reference's 127-split cap (documented divergence for hour-long audio).
    """
    n = len(samples)
    search = min(search_sec, segment_sec / 2.0)
    target = int(segment_sec * SAMPLE_RATE)
    margin = int(search * SAMPLE_RATE)
    if segment_sec <= 0 or n <= target + margin:
        return [(0, n)]
    splits = [0]
    pos = 0
    while pos + target + margin < n:
        sp = find_split_point(samples, pos + target, search)
        splits.append(sp)
        pos = sp
        if max_splits is not None and len(splits) >= max_splits:
            break
    splits.append(n)
    return [(splits[i], splits[i + 1]) for i in range(len(splits) - 1)]


def pad_short(segment, min_samples=SAMPLE_RATE // 2):
    """Segments under 0.5 s are zero-padded (qwen_asr.c:1003-1011)."""
    if len(segment) >= min_samples:
        return segment
    out = np.zeros(min_samples, np.float32)
    out[:len(segment)] = segment
    return out


def shard_range(n_items, rank, world):
    """Contiguous block [lo, hi) of n_items for `rank` of `world` (sizes differ by at most 1)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def tokens_cap(n_samples):
    """Per-segment new-token cap used when weights are random-init and never emit EOS
    (SURVEY.md 8d, config 3): ceil(4 * seconds) + 8."""
    return int(np.ceil(4.0 * n_samples / SAMPLE_RATE)) + 8


def transcribe_segments(engine, samples, ranges, cap_fn=tokens_cap):
    """Run `engine.transcribe_ids` over sample ranges; returns [(ids, info)] in order."""
    out = []
    for (a, b) in ranges:
        seg = pad_short(np.ascontiguousarray(samples[a:b], np.float32))
        out.append(engine.transcribe_ids(seg, cap_fn(b - a)))
    return out


def gather_in_order(local_results, n_items, rank, world, dist=None):
    """All ranks' per-item results concatenated in item order on every rank (object all_gather).
    With world == 1 (or no process group) this is the identity."""
    if world == 1 or dist is None:
        return list(local_results)
    bucket = [None] * world
    dist.all_gather_object(bucket, list(local_results))
    merged = []
    for r in range(world):
        lo, hi = shard_range(n_items, r, world)
        assert len(bucket[r]) == hi - lo, "rank returned a different number of segments than its shard"
        merged.extend(bucket[r])
    return merged

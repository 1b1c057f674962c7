"""Host-side work partitioning for multi-GPU runs (one process + one cucd handle per GPU).

The path has no cross-shard reduction: All-Intra pictures and closed intra periods are
independent, so ranks only need to agree on WHO takes WHICH pictures.  The fork's online
train/verify/test schedule is keyed on POC % 60 and carries its decision switches from the verify
picture to the following 57 (tools_YS.cpp:1237-1242, TEncTop.cpp:395-398, 550-552), so a shard that
must reproduce the serial encoder's bitstream is a whole number of 60-picture periods.
"""

FORK_PERIOD = 60  # g_iP, tools_YS.cpp:42-44


def shard_pictures(n_pictures, rank, world, period=FORK_PERIOD):
    """Contiguous, period-aligned block of picture indices for `rank`: range(start, stop).

    Periods are dealt out as evenly as possible (the first `n_periods % world` ranks get one more);
    a trailing partial period stays with the rank that owns the last full one's successor slot."""
    if world < 1 or not (0 <= rank < world) or n_pictures < 0 or period < 1:
        raise ValueError("bad shard request")
    n_periods = (n_pictures + period - 1) // period
    base, extra = divmod(n_periods, world)
    first = rank * base + min(rank, extra)
    count = base + (1 if rank < extra else 0)
    start = min(first * period, n_pictures)
    stop = min((first + count) * period, n_pictures)
    return range(start, stop)


def shard_independent(n_items, rank, world):
    """Round-robin split for units with no schedule coupling (e.g. the benchmark's synthetic pictures)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad shard request")
    return range(rank, n_items, world)

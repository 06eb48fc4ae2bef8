"""Frame sharding for the multi-GPU front end: frames are independent (no cross-frame state in
ORBextractor::operator() or CAPE::process), so a sequence is cut into contiguous blocks, one per
rank, with no collective on the data path.  Host logic only."""


def frame_block(total_frames, rank, world):
    """[first, last) of the contiguous block rank owns; blocks differ by at most one frame."""
    if not (0 <= rank < world) or total_frames < 0:
        raise ValueError("bad rank/world/total")
    base, extra = divmod(total_frames, world)
    first = rank * base + min(rank, extra)
    return first, first + base + (1 if rank < extra else 0)


def weak_block(frames_per_rank, rank):
    """Weak-scaling layout of bench.py: every rank processes its own frames_per_rank frames."""
    return rank * frames_per_rank, (rank + 1) * frames_per_rank


def owner_of(frame, total_frames, world):
    for r in range(world):
        a, b = frame_block(total_frames, r, world)
        if a <= frame < b:
            return r
    raise ValueError("frame out of range")

"""Work partitioning across the GPUs of one box.  The path shards by independent
units -- encode streams, or within one stream the (b, ref) search fields -- so there is no data-path
collective: every rank computes its own units and returns per-frame cost/MV arrays to the host."""


def assign_streams(n_streams, world, rank):
    """configs[4]: encode streams round-robin over ranks (64 streams / 8 GPUs -> 8 each)."""
    return [s for s in range(n_streams) if s % world == rank]


def assign_fields(fields, world):
    """configs[3]: one stream's search fields (b, ref, list) spread over the GPUs by (b + ref) mod world;
    the evaluation (p0, p1, b) then runs where b's fields live.  Returns {rank: [fields]}."""
    out = {r: [] for r in range(world)}
    for (b, ref, lst) in fields:
        out[(b + ref) % world].append((b, ref, lst))
    return out


def stream_seed(global_stream_id):
    """synthetic content seed of a stream; depends on the global id only, so any world size sees the same streams."""
    return 7 * global_stream_id + 1

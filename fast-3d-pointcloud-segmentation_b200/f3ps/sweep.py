"""Directory-sweep sharding (-d of the reference CLI, src/supervoxel_clustering.cpp:209-227, 303):
files are independent (fresh SupervoxelClustering + Clustering per file, :348,408), so they are dealt
round-robin to the ranks, one GPU + one stream per rank, and no data-path collective exists.  Only the
final timing / score reduction crosses ranks."""


def shard_frames(n_frames, rank, world):
    """Indices of the frames rank `rank` of `world` processes (file order is kept inside a rank)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return list(range(rank, n_frames, world))


def aggregate_throughput(points_per_rank, seconds_per_rank):
    """Whole-job Mpoints/s: all points / slowest rank's time."""
    return sum(points_per_rank) / max(seconds_per_rank) / 1e6


def reduce_sweep(local_points, local_seconds, dist=None):
    """Cross-rank reduction of a sweep with torch.distributed (sum of points, max of time)."""
    if dist is None or not dist.is_initialized():
        return local_points, local_seconds
    import torch
    t = torch.tensor([float(local_points)], dtype=torch.float64)
    m = torch.tensor([float(local_seconds)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    dist.all_reduce(m, op=dist.ReduceOp.MAX)
    return float(t[0]), float(m[0])

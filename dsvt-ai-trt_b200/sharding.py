"""Frame-level data parallelism: frames are independent, so frame f goes to rank f mod N, no
data-path collective exists, and the only exchanges are the final gather of the [500,9] boxes + count per frame
and the max-reduction of the per-rank device time.  Works on NCCL (GPU) and gloo (CPU tests)."""
import torch
import torch.distributed as dist


def frames_for_rank(n_frames: int, rank: int, world: int):
    """Global frame ids processed by `rank` (round-robin, the reference's one-frame-per-stream idea across GPUs)."""
    return list(range(rank, n_frames, world))


def global_frame_id(rank: int, world: int, local_index: int) -> int:
    return local_index * world + rank


def reduce_max(value: float, device="cpu") -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def gather_results(boxes: torch.Tensor, valid: torch.Tensor, dst: int = 0):
    """boxes [F,K,9] f32, valid [F] i32 of this rank's F frames -> on `dst`: ([N*F,K,9], [N*F]) in GLOBAL frame order
    (frame f = local index f // N on rank f % N); None elsewhere."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return boxes, valid
    world, rank = dist.get_world_size(), dist.get_rank()
    bl = [torch.empty_like(boxes) for _ in range(world)] if rank == dst else None
    vl = [torch.empty_like(valid) for _ in range(world)] if rank == dst else None
    dist.gather(boxes, bl, dst=dst)
    dist.gather(valid, vl, dst=dst)
    if rank != dst:
        return None
    F = boxes.shape[0]
    all_b = torch.stack(bl, dim=1).reshape(world * F, *boxes.shape[1:])     # [F, N, ...] -> frame-major interleave
    all_v = torch.stack(vl, dim=1).reshape(world * F)
    return all_b, all_v


def gather_packed(packed: torch.Tensor, dst: int = 0):
    """packed [F_local, W] (one row per frame: boxes + count) -> on `dst` [N * F_local, W] in GLOBAL frame order; None on
    the other ranks.  ONE batched collective (64 x 18 004 B = 1.15 MB); enqueued on the current stream."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return packed
    world, rank = dist.get_world_size(), dist.get_rank()
    parts = [torch.empty_like(packed) for _ in range(world)] if rank == dst else None
    dist.gather(packed, parts, dst=dst)
    if rank != dst:
        return None
    return torch.stack(parts, dim=1).reshape(world * packed.shape[0], packed.shape[1])   # frame f = local f // N on rank f % N

"""Multi-GPU plumbing: one process per GPU, one all-gather of partial points.

Two decompositions, both ending in the same gather + sum of G Jacobian partials:
  * "points"  - rank g owns a contiguous shard of the points and its own table shard (memory / G); the bucket
                reduction is repeated on every rank, so latency stops scaling once it dominates;
  * "buckets" - every rank holds all points and tables and handles 1/G of the bucket-reduction chunks
                (MsmContext.set_bucket_shard): accumulation AND reduction scale 1/G; the scalars are all-gathered
                over NVLink when they arrive sharded from the host.

MSM is a sum over points, so rank g owns points [g*n/G, (g+1)*n/G), builds its own table shard from its own
P_i (no communication), reduces its scalars to ONE partial Jacobian point (144 B G1 / 288 B G2) on its GPU, and
the only exchange step is an all-gather of G partials (NCCL over NVLink on the GPU box, gloo in CPU tests),
followed by G-1 additions and one to_affine. torch.distributed is plumbing only.
"""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous shard [lo, hi) of n points for `rank`; sizes differ by at most one."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


# Measured on B200 (tests/gpu_cfg_sweep_dev.py): configuration 16 (q = 2^19, h = 14) has an 8-bit top digit that piles
# 256 extra entries on each of the lowest 256 bucket values, so a 2^16-point G2 shard is faster under configuration 17
# (2.43 vs 2.92 ms); a 2^20-point G1 shard is faster with q = 2^20, h = 13 (6.00 vs 6.19 ms: the reduction shrinks more
# than the accumulation grows).
_SHARD_CONFIG_OVERRIDE = {(2, 16): "17", (1, 20): "19"}


def shard_config_name(n_shard, group=1):
    """Reference configuration tuned for the shard size (legal: the output encoding is canonical)."""
    exp = max(8, min(21, (max(1, n_shard) - 1).bit_length()))
    return _SHARD_CONFIG_OVERRIDE.get((group, exp), str(exp))


def all_gather_partials(partial, world=None):
    """partial: uint8 tensor of one Jacobian point on this rank's device -> (world, nbytes) tensor on every rank."""
    world = world or dist.get_world_size()
    out = torch.empty((world,) + tuple(partial.shape), dtype=partial.dtype, device=partial.device)
    if partial.is_cuda:
        dist.all_gather_into_tensor(out, partial)
    else:  # gloo (CPU tests)
        parts = [torch.empty_like(partial) for _ in range(world)]
        dist.all_gather(parts, partial)
        out = torch.stack(parts)
    return out


def all_gather_scalars(shard, full):
    """NCCL all-gather of the per-rank scalar shards (uint8 CUDA tensors) into `full` (world * len(shard) bytes)."""
    dist.all_gather_into_tensor(full, shard)
    return full


def _bind_stream(ctx, tensor):
    """All three legs (MSM kernels, the collective, the combine) are ordered on ONE stream: the context is bound to
    torch's current stream of the tensor's device, which is also the stream NCCL orders its collective after."""
    if tensor.is_cuda:
        cur = torch.cuda.current_stream(tensor.device).cuda_stream
        if getattr(ctx, "_bound_stream", None) != cur:
            ctx.set_stream(cur)
            ctx._bound_stream = cur


def msm_sharded(ctx, method, scalars_dev, partial_buf=None):
    """Run this rank's shard on its GPU and combine: returns the affine result (numpy uint8) on every rank.

    ctx: msm_blst_b200.MsmContext for this rank's shard (same configuration on every rank); scalars_dev: torch uint8 CUDA
    tensor (n_shard x 32). Every rank stops after its bucket reduction and contributes its per-bit XYZZ sums (a few KB);
    after ONE all-gather the entries are summed over the ranks and a single Horner pass + inversion finishes the job
    (msmb200_msm_bits_device / msmb200_combine_bits_device). Contexts that cannot hand out per-bit sums fall back to
    Jacobian partials (partial_buf: torch uint8 CUDA tensor of JAC_BYTES, allocated on demand).
    Stream ordering: the context is bound to torch's current stream, so the MSM, the NCCL collective and the combine are
    ordered without host synchronisation.
    """
    from .api import JAC_BYTES, XYZZ_BYTES
    _bind_stream(ctx, scalars_dev)
    cache = ctx.__dict__.setdefault("_bits_cache", {})
    if method not in cache:
        lay = ctx.msm_bits_layout(method)
        buf = torch.zeros(lay[0] * lay[1] * XYZZ_BYTES[ctx.group], dtype=torch.uint8, device=scalars_dev.device) if lay else None
        cache[method] = (lay, buf)
    lay, buf = cache[method]
    if lay is not None:
        ctx.msm_bits_device(method, scalars_dev.data_ptr(), buf.data_ptr())
        gathered = all_gather_partials(buf)
        return ctx.combine_bits_device(gathered.data_ptr(), gathered.shape[0], lay)
    if partial_buf is None:
        partial_buf = torch.zeros(JAC_BYTES[ctx.group], dtype=torch.uint8, device=scalars_dev.device)
    ctx.msm_partial_device(method, scalars_dev.data_ptr(), partial_buf.data_ptr())
    gathered = all_gather_partials(partial_buf)
    return ctx.sum_partials_device(gathered.data_ptr(), gathered.shape[0])

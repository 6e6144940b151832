"""Multi-GPU sharding of the smoothing path: one process per GPU, sessions are the shard unit.

Independent (session, keypoint) sequences never exchange data on the recursion, so the data
path needs no collective: each rank smooths its own sessions.  torch.distributed is used only to agree on
the partition and to gather the small per-keypoint results (s, iteration counts).
"""

from __future__ import annotations

import os

import torch
import torch.distributed as dist


def _parse_cpulist(text: str) -> list[int]:
    cpus = []
    for part in text.strip().split(','):
        if not part:
            continue
        if '-' in part:
            lo, hi = part.split('-')
            cpus.extend(range(int(lo), int(hi) + 1))
        else:
            cpus.append(int(part))
    return cpus


def bind_to_gpu_numa_node(device_index: int) -> dict:
    """Pin this process (one rank per GPU) to the CPUs of the NUMA node its GPU hangs off, BEFORE it allocates pinned
    host buffers: page-locked memory is placed on the node of the allocating thread, and a host -> device copy that
    crosses the inter-socket link is what limited the 8-GPU end-to-end number of round 1 (eight ranks, every pinned
    buffer on node 0: 14.9 GB/s per rank against 48.5 GB/s for one rank).  Linux sysfs only; returns what was done
    ({'node': n, 'cpus': count} or {'node': None, 'why': ...}) and never raises."""
    try:
        props = torch.cuda.get_device_properties(device_index)
        bus = f'{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0'
        with open(f'/sys/bus/pci/devices/{bus}/numa_node') as f:
            node = int(f.read().strip())
        if node < 0:
            return {'node': None, 'why': 'the platform reports no NUMA affinity for this GPU'}
        with open(f'/sys/devices/system/node/node{node}/cpulist') as f:
            cpus = _parse_cpulist(f.read())
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return {'node': node, 'why': 'none of the node\'s CPUs is in this process\'s cpuset'}
        os.sched_setaffinity(0, allowed)
        return {'node': node, 'cpus': len(allowed)}
    except Exception as e:   # not Linux, no sysfs, restricted container ...
        return {'node': None, 'why': f'{type(e).__name__}: {e}'}


def shard_indices(n_items: int, rank: int, world_size: int) -> list[int]:
    """Contiguous, balanced partition of range(n_items): the first (n_items % world) ranks get one more."""
    if not (0 <= rank < world_size):
        raise ValueError(f'rank {rank} outside world of size {world_size}')
    base, extra = divmod(n_items, world_size)
    start = rank * base + min(rank, extra)
    return list(range(start, start + base + (1 if rank < extra else 0)))


def gather_session_results(local: torch.Tensor, n_items: int, group=None) -> torch.Tensor:
    """All-gather per-session rows (n_local, ...) from every rank into (n_items, ...) in session order.

    Shards may be ragged (n_items not divisible by the world size): rows are padded to the largest shard
    for the collective and trimmed afterwards."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    counts = [len(shard_indices(n_items, r, world)) for r in range(world)]
    assert local.shape[0] == counts[rank], 'local rows do not match this rank\'s shard'
    width = max(counts)
    pad = torch.zeros((width, *local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return torch.cat([o[:c] for o, c in zip(out, counts)], dim=0)


def smooth_sessions_sharded(load_session, n_sessions: int, smooth_fn, group=None):
    """Run `smooth_fn(raw)` on this rank's sessions (`raw = load_session(i)`), gather s_finals.

    Returns (local_results: list, s_all (n_sessions, K) tensor)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    mine = shard_indices(n_sessions, rank, world)
    results = [smooth_fn(load_session(i)) for i in mine]
    s_local = torch.stack([r.s_finals.reshape(-1) for r in results]) if results else None
    if world == 1:
        return results, s_local
    if s_local is None:  # more ranks than sessions
        raise ValueError('every rank needs at least one session')
    return results, gather_session_results(s_local, n_sessions, group)

"""Host-side logic of the multi-GPU path: contiguous object slices, one process
per GPU, no data-path collective.  Pure Python + torch.distributed plumbing, CPU-testable with
the gloo backend (tests/test_sharding_gloo.py).

* ``shard_range``        - slice of rank r: starts on a multiple of 1024 objects so every slice
                           covers whole 128-byte lines of the bitset (and whole u32 / u64 words).
* ``word_offset``        - where a slice's words go inside the full bitset.
* ``merge_changed``      - per-rank changed lists (local indices, ascending) -> one global
                           ascending list: concatenation in rank order is already sorted.
* ``tree_shard``         - the transform hierarchy of one rank (SURVEY.md 8e "Transform propagation"): levels
                           0..L-2 replicated (every GPU propagates them redundantly, no communication), the
                           leaf level sharded with the objects.
* ``exchange_ipc``       - all-gather of cudaIpc handles so that every rank can hand the cull
                           kernel its peers' full-bitset buffers (dpcuCullResultSetPeerBits):
                           the bitset all-gather then happens as NVLink stores inside the
                           kernel's epilogue.
* ``allgather_words``    - the library collective, used ONLY to verify the fused path (and as
                           the stated alternative when peer access is unavailable).
"""
from __future__ import annotations

import numpy as np

SLICE_ALIGN = 1024


def shard_range(n_total: int, world: int, rank: int):
    """(first, count) of rank's contiguous slice; all but the last slice are multiples of 1024."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    blocks = (n_total + SLICE_ALIGN - 1) // SLICE_ALIGN
    per = (blocks + world - 1) // world
    first = min(rank * per * SLICE_ALIGN, n_total)
    last = min((rank + 1) * per * SLICE_ALIGN, n_total)
    return first, last - first


def word_offset(first: int) -> int:
    assert first % 32 == 0
    return first // 32


def total_words(n_total: int) -> int:
    return (n_total + 31) // 32


def merge_changed(per_rank_lists, firsts):
    """Concatenate per-rank changed lists (local ascending indices) into the global ascending list."""
    out = [np.asarray(l, dtype=np.uint64) + np.uint64(f) for l, f in zip(per_rank_lists, firsts)]
    return np.concatenate(out) if out else np.zeros(0, np.uint64)


def place_words(full: np.ndarray, local_words: np.ndarray, first: int):
    """What the kernel epilogue does with peer stores, on the host (for tests)."""
    o = word_offset(first)
    full[o:o + len(local_words)] = local_words
    return full


def exchange_ipc(dist, handle_bytes: bytes):
    """All ranks' handles, indexed by rank."""
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, handle_bytes)
    return out


def allgather_words(dist, local_words, words_per_rank, device=None):
    """Library all-gather of equally sized word slices (torch tensor in, list of tensors out)."""
    import torch
    t = local_words if isinstance(local_words, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(local_words).view(np.int32))
    if device is not None:
        t = t.to(device)
    outs = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(outs, t)
    return outs


def tree_shard(levels, world: int, rank: int):
    """Topology of rank `rank`'s transform tree for a uniform-fan-out hierarchy `levels` (nodes per level, e.g.
    scenes.hierarchy_topology's): the upper levels whole, and of the last level only the leaves of this rank's
    object slice [first, first + count) (object i <-> leaf i).

    Returns (entries (E,2) uint32, level_offsets (L+1,) uint32, n_nodes, first_leaf_local, first, count,
    global_node_of_local) where global_node_of_local maps a node index of the shard tree to its index in the
    whole tree (local matrices are a function of the global node index)."""
    from . import scenes
    levels = tuple(int(x) for x in levels)
    n_leaves = levels[-1]
    first, count = shard_range(n_leaves, world, rank)
    up_entries, up_offsets, n_upper = scenes.hierarchy_topology(levels[:-1])        # incl. the virtual root 0
    fan = n_leaves // levels[-2]
    first_parent = n_upper - levels[-2]                                              # first node of level L-2
    k = np.arange(first, first + count, dtype=np.uint64)
    parent = (np.uint64(first_parent) + k // np.uint64(fan)).astype(np.uint32)
    node = (np.uint64(n_upper) + (k - np.uint64(first))).astype(np.uint32)
    entries = np.concatenate([up_entries, np.stack([parent, node], axis=1)]).astype(np.uint32)
    offsets = np.concatenate([up_offsets, [up_offsets[-1] + count]]).astype(np.uint32)
    n_nodes = n_upper + count
    global_of_local = np.concatenate([np.arange(n_upper, dtype=np.uint64), np.uint64(n_upper) + k])
    return entries, offsets, n_nodes, n_upper, first, count, global_of_local

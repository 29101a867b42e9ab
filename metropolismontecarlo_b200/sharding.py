"""Host-side plumbing of the sharded full-energy path: one process per GPU,
replicated coordinates, pair work units and ρ(k) sites split per rank, ONE all-reduce of the
partial-sum vector, every rank finalises.

Partial-sum vector layout (mmc_partial_count doubles, see csrc/mmc_common.cuh MMC_NSCAL):
  [0] Σ lj_pot  [1] Σ lj_vir  [2] Σ coulomb (un-scaled)  [3] #overlapped molecules
  [4] E_recip (filled by finalize)  [5] #pairs in cutoff  [6], [7] internal
  [8 + 2k], [9 + 2k]  Re, Im of this rank's ρ(k) partial
"""
from __future__ import annotations

NSCAL = 8


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous share [lo, hi) of n items for `rank` — the same integer formula the library uses
    for cell layers and for sites (mmc_eval.cu: n * rank / world)."""
    return n * rank // world, n * (rank + 1) // world


def partial_len(nkvecs: int) -> int:
    return NSCAL + 2 * max(nkvecs, 1)


def sharded_potential(eng, style, vec, world: int, group=None):
    """partial → all-reduce (NCCL over NVLink when vec is a CUDA tensor) → finalize.
    `vec` must live on the stream the engine was created with."""
    eng.potential_partial(style, vec.data_ptr())
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(vec, group=group)
    return eng.potential_finalize(style, vec.data_ptr())


def setup_peer_exchange(eng, world: int, group=None):
    """One process per GPU on one node: every rank exports the CUDA IPC handle of its exchange buffer, the handles
    travel through torch.distributed (any transport would do), every rank maps every peer's buffer.  After this
    eng.potential_sharded(style) needs no collective library: partial sums go peer to peer over NVLink."""
    import torch.distributed as dist
    try:
        mine = eng.peer_export()
    except Exception:
        mine = None
    handles = [None] * world
    dist.all_gather_object(handles, mine, group=group)      # every rank takes part even if its export failed
    if any(hd is None for hd in handles):
        raise RuntimeError("a rank could not export its exchange buffer (CUDA IPC unavailable)")
    for r, hd in enumerate(handles):
        eng.peer_import(r, hd)

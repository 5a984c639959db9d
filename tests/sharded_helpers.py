"""Helpers shared by tests/test_sharded_gpu.py (G shards searched by G threads of one process on one GPU) and
tests/multigpu_check.py (one rank per GPU under torchrun)."""
import threading

import torch

from scaling_retriever_b200 import _lib, ops


class LocalExchange:
    """Stand-in for shard.TauExchange when the G shards of a sharded search live in ONE process: every shard is searched by its
    own thread on its own stream and the per-round MIN "all-reduce" of the published bounds (include/b200ret.h (3b)) is a
    thread barrier + an element-wise minimum.  Same struct, same hook protocol, no process group."""

    class Shared:
        def __init__(self, world):
            self.barrier = threading.Barrier(world, timeout=120)
            self.members = []

    def __init__(self, shared, n_exchanges, growth, world, device):
        self.shared, self.n_exchanges, self.growth, self.world, self.device = shared, int(n_exchanges), int(growth), world, device
        self.aux = None
        self.rounds_seen = 0
        self._hook = _lib.RoundExchange.HOOK(self._on_round)
        shared.members.append(self)

    def _on_round(self, _user):
        try:
            stream = torch.cuda.current_stream(self.device)
            stream.synchronize()                      # this shard's bound is in self.aux
            self.shared.barrier.wait()
            low = torch.stack([m.aux for m in self.shared.members]).amin(dim=0)
            stream.synchronize()
            self.shared.barrier.wait()                # everyone has read every aux before anyone overwrites its own
            self.aux.copy_(low)
            self.rounds_seen += 1
            return 0
        except Exception:                             # a broken barrier (another shard failed): the C side returns an error
            return 1

    def struct(self, n_queries, k):
        import ctypes
        if self.aux is None or self.aux.numel() != n_queries:
            self.aux = torch.empty(n_queries, dtype=torch.float32, device=self.device)
        aux_rank = (int(k) + self.world - 1) // self.world
        self._struct = _lib.RoundExchange(aux_rank, self.n_exchanges, self.growth, 0, self.aux.data_ptr(), self._hook, None)
        return ctypes.byref(self._struct)


def search_shards_threaded(search_one, world, device):
    """Run `search_one(g)` for g in range(world) on `world` threads, each under its own CUDA stream; returns the results."""
    out, errors = [None] * world, []

    def work(g):
        try:
            torch.cuda.set_device(device)
            stream = torch.cuda.Stream(device)
            stream.wait_stream(torch.cuda.default_stream(device))
            with torch.cuda.stream(stream):
                out[g] = search_one(g)
            stream.synchronize()
        except Exception as exc:        # noqa: BLE001 - reported by the caller
            errors.append((g, exc))

    threads = [threading.Thread(target=work, args=(g,)) for g in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0][1]
    return out


def rising_score_shard(device, k):
    """One shard (32 doc blocks, 64 terms) whose scores RISE with the doc id, so every later document passes tau:

    * term 0 ("hot"): 0.7 x the candidate list's head-room of docs spread over blocks [2, 8) and as many over [8, 32), weights
      increasing — each round of the plain geometric schedule (blocks [0,2) [2,8) [8,32)) fits its list, both together do not;
    * terms 1..50: one per remaining doc, weight 0.5 (a low floor of matches);
    * term 51: every doc, weight increasing — any round larger than the first overflows.
    Returns (rows, cols, vals) with LOCAL doc ids (int32, int32, fp32), the shard's doc count, and the two query sets
    ((q_offsets, q_terms, q_weights) over term 0 + three floor terms / over term 51 only)."""
    bd, round0 = ops.block_docs(), 2
    room = max(round0 * bd, 5 * k)                       # candidate capacity = k + room (sparse_search.cu search_cap)
    n_hot = int(0.7 * room)
    n_shard = 32 * bd
    j = torch.arange(n_shard, device=device)
    hot = torch.cat([2 * bd + (torch.arange(n_hot, device=device) * (6 * bd)) // n_hot,
                     8 * bd + (torch.arange(n_hot, device=device) * (24 * bd)) // n_hot])
    is_hot = torch.zeros(n_shard, dtype=torch.bool, device=device)
    is_hot[hot] = True
    cold = j[~is_hot]
    rows = torch.cat([hot, cold, j]).to(torch.int32)
    cols = torch.cat([torch.zeros_like(hot), 1 + cold % 50, torch.full_like(j, 51)]).to(torch.int32)
    vals = torch.cat([1.0 + torch.arange(hot.numel(), device=device) * 1e-4, torch.full((cold.numel(),), 0.5, device=device),
                      1.0 + j * 1e-5]).float()
    nq = 48
    q = torch.arange(nq, device=device)
    fill = torch.sort(1 + (q[:, None] * 3 + torch.arange(3, device=device)[None, :]) % 50, dim=1).values
    qa = ((torch.arange(nq + 1, device=device) * 4).to(torch.int32),
          torch.cat([torch.zeros(nq, 1, dtype=torch.int64, device=device), fill], dim=1).reshape(-1).to(torch.int32),
          torch.cat([1.0 + 0.01 * q[:, None], torch.full((nq, 3), 0.5, device=device)], dim=1).reshape(-1).float())
    qb = (torch.arange(nq + 1, device=device).to(torch.int32), torch.full((nq,), 51, dtype=torch.int32, device=device),
          (1.0 + 0.01 * q).float())
    return (rows, cols, vals), n_shard, 64, qa, qb


# score-kernel launches of a forced-overflow search of one rising-score shard: 2 rounds of the forced schedule, 3 rounds of the
# middle tier (plain schedule, the shard's own bounds), 16 rounds of the safe schedule (32 blocks in rounds of 2)
LAUNCHES_MIDDLE_TIER = 2 + 3
LAUNCHES_SAFE_TIER = 2 + 3 + 16

"""Batched entry points: millions of independent games per call, results as torch tensors.

This is the new surface the B200 build adds on top of the reference's per-object API.  Every tensor
returned lives on the GPU (export with ``torch.utils.dlpack.to_dlpack`` / ``tensor.__dlpack__()``),
is allocated by torch and only *filled* by ``libbgs_b200.so``; the ``*_host`` helpers copy the small
per-game results into pinned host memory.

The loop being replaced is the reference's README.md:49-72::

    state = config.sample_initial_state()
    while not state.has_ended:
        action = random.choice(state.actions)
        state = action.sample_next_state()
    reward = state.reward
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Any

from . import _native as N


def _hwk(config) -> tuple[int, int, int]:
    if isinstance(config, (tuple, list)):
        h, w, k = config
    else:
        h, w, k = config.height, config.width, config.count
    return int(h), int(w), int(k)


@dataclass
class RolloutResult:
    """Outputs of a batched rollout. ``stats`` is the int64[256] vector that is summed across GPUs."""

    n_games: int
    game_id0: int
    seed: int
    stats: Any
    length: Any = None
    winner: Any = None
    actions: Any = None  # Connect: uint8[n, H*W] columns; Bounce: uint8[n, max_plies, 2] cells
    final_grid: Any = None
    reward: Any = None
    extra: dict = field(default_factory=dict)

    def stats_dict(self) -> dict[str, int]:
        s = self.stats.cpu()
        return {
            "games": int(s[N.STAT_GAMES]),
            "wins0": int(s[N.STAT_WIN0]),
            "wins1": int(s[N.STAT_WIN1]),
            "draws": int(s[N.STAT_DRAWS]),
            "steps": int(s[N.STAT_STEPS]),
            "truncated": int(s[N.STAT_TRUNCATED]),
        }

    def length_histogram(self):
        return self.stats[N.STAT_HIST0:].clone()

    def to_json(self, config) -> list[dict]:
        """One dict per game in the reference's wire format (tests/test_connect.py:123-145,
        tests/test_bounce.py:365-410): ``{"state": State.to_json() of the final position (needs
        ``final_grid``), "actions": [Action.to_json() per ply played]}`` -- what
        ``[a.to_json() for a in trajectory]`` / ``state.to_json()`` give through the object API."""
        n = self.n_games
        length = self.length.cpu().numpy().astype("int64")
        winner = self.winner.cpu().numpy()
        out = []
        if self.actions is not None and self.actions.dim() == 3:  # Bounce: (source cell, target cell) per ply
            W = int(self.final_grid.shape[2]) if self.final_grid is not None else int(_bounce_grid(config).shape[1])
            mv = self.actions.cpu().numpy()
            acts = [[{"source": [int(s % W), int(s // W)], "target": [int(t % W), int(t // W)]} for s, t in mv[i, : length[i]]]
                    for i in range(n)]
        elif self.actions is not None:
            a = self.actions.cpu().numpy()
            acts = [[{"column": int(c)} for c in a[i, : length[i]]] for i in range(n)]
        else:
            acts = [None] * n
        grids = self.final_grid.cpu().numpy() if self.final_grid is not None else None
        first = self.extra.get("first_player")
        first = None if first is None else first.cpu().numpy().astype("int64")
        for i in range(n):
            d = {}
            if grids is not None:
                w = int(winner[i])
                d["state"] = {"grid": grids[i].tolist(),
                              "player": int((length[i] + (0 if first is None else first[i])) & 1), "winner": w if w >= 0 else -1}
            if acts[i] is not None:
                d["actions"] = acts[i]
            out.append(d)
        return out


def all_reduce_stats(stats):
    """Sum the statistics vector over all ranks (NCCL over NVLink; the path's only collective)."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    return stats


def shard_range(n_total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous shard [start, start+count) of game ids owned by ``rank``."""
    base, rem = divmod(int(n_total), int(world))
    start = rank * base + min(rank, rem)
    return start, base + (1 if rank < rem else 0)


# -------------------------------------------------------------------------------------------------
# Connect-k
# -------------------------------------------------------------------------------------------------
def connect_rollout(
    config,
    n_games: int,
    seed: int = 0,
    game_id0: int = 0,
    *,
    per_game: bool = True,
    actions: bool = False,
    final_grid: bool = False,
    reward: bool = False,
    stats=None,
    out: RolloutResult | None = None,
    start: "ConnectBatch | ConnectPacked | None" = None,
) -> RolloutResult:
    """Play ``n_games`` uniform-random Connect-k games to the end on the current CUDA device.

    With ``start`` (a :class:`ConnectBatch`, or its 17-bytes-per-position :class:`ConnectPacked` form, of ``n_games`` positions) the games continue from those
    positions instead of the empty board -- the leaf evaluation of a tree search: ``length`` and
    ``actions`` then cover only the plies played by the rollout, and game ``i`` draws from the stream of
    global id ``game_id0 + i`` starting at draw 0.

    ``per_game`` returns ``length`` uint8[n] and ``winner`` int8[n] (0 / 1 / -1 draw); ``actions``
    the trajectories uint8[n, H*W] (0xFF padded); ``final_grid`` int8[n,H,W]; ``reward``
    float32[n,2].  ``stats`` (int64[256], device) is accumulated into when given.  Pass a previous
    result as ``out`` to reuse its buffers.  Asynchronous: nothing is synchronised here.
    """
    torch = N.require_cuda()
    L = N.lib()
    H, W, K = _hwk(config)
    if not L.bgs_connect_supported(H, W, K):
        raise RuntimeError(f"Connect {H}x{W} k={K} is not supported by the CUDA kernels")
    n = int(n_games)
    dev = torch.device("cuda", torch.cuda.current_device())

    def buf(name, shape, dtype, want):
        if not want:
            return None
        t = getattr(out, name, None) if out is not None else None
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype or t.device != dev:
            t = torch.empty(shape, dtype=dtype, device=dev)
        return t

    want_winner = per_game or reward
    res = RolloutResult(n_games=n, game_id0=int(game_id0), seed=int(seed), stats=None)
    res.length = buf("length", (n,), torch.uint8, per_game or actions)
    res.winner = buf("winner", (n,), torch.int8, want_winner)
    res.actions = buf("actions", (n, H * W), torch.uint8, actions)
    res.final_grid = buf("final_grid", (n, H, W), torch.int8, final_grid)
    res.reward = buf("reward", (n, 2), torch.float32, reward)
    if stats is None:
        stats = torch.zeros(N.STATS_LEN, dtype=torch.int64, device=dev)
    res.stats = stats
    st = N.stream_ptr(torch)
    seed64 = int(seed) & 0xFFFFFFFFFFFFFFFF
    if start is None:
        # one call: every output already in the reference's layouts; a single pass (the rollout kernel
        # writes trajectory rows, final grids and rewards itself) on the boards that have a fused kernel
        N.check(
            L.bgs_connect_rollout_export(
                H, W, K, n, int(game_id0), seed64, N.ptr(res.actions), N.ptr(res.length), N.ptr(res.winner),
                N.ptr(res.final_grid), N.ptr(res.reward), N.ptr(stats), st,
            )
        )
        return res
    if start.n != n or (not isinstance(start, ConnectPacked) and tuple(start.grid.shape[1:]) != (H, W)):
        raise ValueError("start must hold n_games positions of this configuration")
    packed = None
    if final_grid:
        pw = L.bgs_connect_packed_words(H, W)
        packed = out.extra.get("packed") if out is not None else None
        if packed is None or tuple(packed.shape) != (n, pw) or packed.device != dev:
            packed = torch.empty((n, pw), dtype=torch.int64, device=dev)
        res.extra["packed"] = packed
    work = out.extra.get("workspace") if out is not None else None
    sw = L.bgs_connect_start_words(H, W)
    if work is None or tuple(work.shape) != (n, sw) or work.device != dev:
        work = torch.empty((n, sw), dtype=torch.int64, device=dev)
    if isinstance(start, ConnectPacked):  # 17 / 33 bytes per position instead of H*W + 2
        N.check(
            L.bgs_connect_rollout_from_packed(
                H, W, K, n, int(game_id0), seed64, N.ptr(start.packed), N.ptr(start.meta), N.ptr(work),
                N.ptr(res.actions), N.ptr(res.length), N.ptr(res.winner), N.ptr(packed), N.ptr(stats), st,
            )
        )
    else:
        grid0 = start.grid.contiguous()
        N.check(
            L.bgs_connect_rollout_from(
                H, W, K, n, int(game_id0), seed64,
                N.ptr(grid0), N.ptr(start.player.contiguous()), N.ptr(start.winner.contiguous()), N.ptr(work),
                N.ptr(res.actions), N.ptr(res.length), N.ptr(res.winner), N.ptr(packed), N.ptr(stats), st,
            )
        )
    res.extra["workspace"] = work
    if final_grid or reward:
        N.check(
            L.bgs_connect_export(H, W, n, N.ptr(packed), N.ptr(res.winner), N.ptr(res.final_grid), N.ptr(res.reward), st)
        )
    return res


def connect_trajectory_grids(config, actions, length, out=None):
    """Observation tensors for a learner: ``int8[n, H*W+1, H, W]`` with entry ``t`` = the position
    after ``t`` plies of each recorded game (``actions`` uint8[n, H*W], ``length`` uint8[n] from
    :func:`connect_rollout`); entries past the end of a game repeat its final position."""
    torch = N.require_cuda()
    H, W, _ = _hwk(config)
    n = int(actions.shape[0])
    if tuple(actions.shape) != (n, H * W) or tuple(length.shape) != (n,):
        raise ValueError("actions must be uint8[n, H*W] and length uint8[n]")
    if out is None:
        out = torch.empty((n, H * W + 1, H, W), dtype=torch.int8, device=actions.device)
    N.check(
        N.lib().bgs_connect_trajectory_grids(
            H, W, n, N.ptr(actions.contiguous()), N.ptr(length.contiguous()), N.ptr(out), N.stream_ptr(torch)
        )
    )
    return out


def gpu_local_cpus(device_index: int):
    """The CPUs NVML reports as local to GPU ``device_index`` (same NUMA node / PCIe root), or None."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        return cpus or None
    except Exception:
        return None


class _numa_local:
    """Context manager: run the enclosed allocations on CPUs local to the GPU, so that pinned host pages
    (allocated where the calling thread runs) land on the GPU's NUMA node; the caller's affinity is restored."""

    def __init__(self, device_index: int, enabled: bool = True):
        self.cpus = gpu_local_cpus(device_index) if enabled else None
        self.saved = None

    def __enter__(self):
        if self.cpus:
            try:
                self.saved = os.sched_getaffinity(0)
                os.sched_setaffinity(0, self.cpus)
            except OSError:
                self.saved = None
        return self

    def __exit__(self, *exc):
        if self.saved is not None:
            try:
                os.sched_setaffinity(0, self.saved)
            except OSError:
                pass
        return False


class HostRollout:
    """End-to-end rollouts with results in pinned host memory (what bench.py's ``e2e`` times).

    A rollout from the empty board has no tensor input -- its inputs are the scalars (config, n_games,
    seed, game_id0), which travel in the kernel launch parameters -- so the host->device side is 0
    bytes; the per-game results and the statistics vector come back device->host into pinned memory for
    EVERY batch, as ONE copy: the device-side record of a batch is ``[per-game results | int64[256]
    statistics]`` in one buffer (the rollout kernel accumulates its statistics straight into the tail).

    ``game="connect"`` (``config`` = (H, W, K) or a Config) or ``game="bounce"`` (``config`` = the start grid).
    ``packed=True`` shrinks the per-game results on the device before they cross PCIe:
      Connect, at most 63 cells: 1 byte ``length | (winner + 1) << 6``; 64..127 cells: 1 byte
      ``length | draw << 7`` (a decided game's winner is the parity of its length); Bounce: 2 bytes
      ``length | (winner + 2) << 14``.  ``HostRollout.unpack`` / ``unpack_results`` recover (length, winner).
    ``packed="dense"`` (Connect): several games per 16-bit word in base S = number of possible outcomes
      (6x7x4: 3 games per word, 5.33 bits per game) -- for hosts whose D2H bandwidth is the limit.
    ``packed=False``: length and winner arrays as they are (Connect 2, Bounce 3 bytes per game).

    ``run`` is the synchronous call (kernel, then copy).  ``stream`` is the pipelined iterator: the
    device->host copy of batch i runs on a copy stream while the kernel of batch i+1 runs on the
    compute stream (``depth`` buffer sets), which hides the PCIe time behind the kernel.  The pinned
    buffers are allocated on CPUs local to the GPU (NUMA node of its PCIe root; ``numa_local=False`` disables).
    """

    def __init__(self, config, n_games: int, depth: int = 2, packed: bool = False, game: str = "connect",
                 max_plies: int = 512, rules: int = 0, numa_local: bool = True):
        torch = N.require_cuda()
        self.torch = torch
        self.config = config
        self.game = game
        self.n = n = int(n_games)
        self.depth = int(depth)
        self.packed = bool(packed)
        self.max_plies, self.rules = int(max_plies), int(rules)
        dev = torch.device("cuda", torch.cuda.current_device())
        if game == "connect":
            H, W, K = _hwk(config)
            if packed == "dense":
                import ctypes as C

                lmin, sym = C.c_int(0), C.c_int(0)
                self.dense = (N.lib().bgs_connect_dense_results(H, W, K, C.byref(lmin), C.byref(sym)), lmin.value, sym.value)
                if self.dense[0] == 0:
                    raise ValueError("no dense result code for this board")
                self.mode = "dense"
            else:
                if packed and H * W > 127:
                    raise ValueError("packed per-game results need a board of at most 127 cells")
                self.mode = ("u8" if H * W <= 63 else "u8wide") if packed else "connect2"
        elif game == "bounce":
            if packed and self.max_plies > 16383:
                raise ValueError("packed Bounce results need max_plies < 16384")
            self.mode = "u16" if packed else "bounce3"
        else:
            raise ValueError("game must be 'connect' or 'bounce'")
        n16 = (n + 15) // 16 * 16
        #: byte offsets of the pieces of one batch record
        if self.mode in ("u8", "u8wide"):
            self.off = {"result": 0}
            body = n16
        elif self.mode == "dense":
            self.off = {"result": 0}
            self.nwords = (n + self.dense[0] - 1) // self.dense[0]
            body = (2 * self.nwords + 15) // 16 * 16
        elif self.mode == "connect2":
            self.off = {"length": 0, "winner": n16}
            body = 2 * n16
        elif self.mode == "u16":
            self.off = {"result": 0}
            body = 2 * n16
        else:
            self.off = {"length": 0, "winner": 2 * n16}
            body = 3 * n16
        self.off["stats"] = body
        self.nbytes = body + N.STATS_LEN * 8
        self.sets = []
        with _numa_local(dev.index, numa_local) as pin:
            self.numa_cpus = len(pin.cpus) if pin.cpus else None
            for _ in range(self.depth):
                rec_dev = torch.zeros(self.nbytes, dtype=torch.uint8, device=dev)
                rec_host = torch.zeros(self.nbytes, dtype=torch.uint8).pin_memory()
                self.sets.append({
                    "rec_dev": rec_dev, "rec_host": rec_host,
                    "stats_dev": rec_dev[body:].view(torch.int64),
                    "res": None,
                    "computed": torch.cuda.Event(),
                    "copied": torch.cuda.Event(),
                })
        self.copy_stream = torch.cuda.Stream()
        # Bounce: every buffer set launches on a compute stream of its own.  One ~400-ply game among millions ends a
        # launch with a ~1 ms tail in which the GPU is nearly idle; with the batches alternating over `depth` streams
        # the next batch's CTAs fill the SMs while the stragglers of the previous one finish (8.4 -> 7.8 ms per batch
        # of 4 Mi default games).  A Connect launch has no such tail (a game lasts at most H*W plies).
        for s in self.sets:
            s["cs"] = torch.cuda.Stream() if (game == "bounce" and self.depth > 1) else None
        self.h2d_bytes = 0
        self.d2h_bytes = self.nbytes

    # -- views of one record ---------------------------------------------------------------------------------
    def _view(self, rec, name, dtype, count):
        torch = self.torch
        o = self.off[name]
        return rec[o: o + count * torch.empty(0, dtype=dtype).element_size()].view(dtype)

    @staticmethod
    def unpack(result):
        """(length uint8, winner int8) from 1-byte packed Connect results of a board of at most 63 cells."""
        import torch

        return result & 63, (result >> 6).to(torch.int8) - 1

    def unpack_results(self, result):
        """(length, winner) from this object's packed per-game results (host or device tensor)."""
        torch = self.torch
        if self.mode == "u8":
            return self.unpack(result)
        if self.mode == "u8wide":
            length = result & 127
            decided = torch.where((length & 1) == 1, 0, 1).to(torch.int8)
            return length, torch.where(result >= 128, torch.full_like(decided, -1), decided)
        if self.mode == "u16":
            r = result.to(torch.int32) & 0xFFFF
            return (r & 0x3FFF).to(torch.int16), ((r >> 14) - 2).to(torch.int8)
        if self.mode == "dense":
            G, lmin, S = self.dense
            r = result.to(torch.int32) & 0xFFFF
            syms = []
            for _ in range(G):
                syms.append(r % S)
                r = r // S
            sym = torch.stack(syms, dim=1).reshape(-1)[: self.n]
            draw = sym == S - 1
            length = torch.where(draw, torch.full_like(sym, _hwk(self.config)[0] * _hwk(self.config)[1]), sym + lmin)
            winner = torch.where(draw, torch.full_like(sym, -1), 1 - (length & 1))
            return length.to(torch.uint8), winner.to(torch.int8)
        raise ValueError("results are not packed")

    def _launch(self, s, seed, game_id0):
        torch = self.torch
        if s["cs"] is not None and torch.cuda.current_stream() != s["cs"]:
            s["cs"].wait_stream(torch.cuda.current_stream())  # ordered after whatever the caller queued before
            with torch.cuda.stream(s["cs"]):
                return self._launch(s, seed, game_id0)
        n = self.n
        s["stats_dev"].zero_()
        rec = s["rec_dev"]
        L = N.lib()
        if self.game == "connect":
            out = s["res"]
            if out is None and self.mode == "connect2":  # the kernel writes straight into the record
                out = RolloutResult(n_games=n, game_id0=0, seed=0, stats=None,
                                    length=self._view(rec, "length", torch.uint8, n),
                                    winner=self._view(rec, "winner", torch.int8, n))
            s["res"] = connect_rollout(self.config, n, seed, game_id0, per_game=True, stats=s["stats_dev"], out=out)
            if self.mode in ("u8", "u8wide"):
                fn = L.bgs_connect_pack_results if self.mode == "u8" else L.bgs_connect_pack_results_wide
                N.check(fn(n, N.ptr(s["res"].length), N.ptr(s["res"].winner), N.ptr(rec), N.stream_ptr(torch)))
            elif self.mode == "dense":
                H, W, K = _hwk(self.config)
                N.check(L.bgs_connect_pack_results_dense(H, W, K, n, N.ptr(s["res"].length), N.ptr(s["res"].winner),
                                                         N.ptr(rec), N.stream_ptr(torch)))
        else:
            s["res"] = bounce_rollout(self.config, n, seed, game_id0, max_plies=self.max_plies, rules=self.rules,
                                      per_game=True, stats=s["stats_dev"])
            if self.mode == "u16":
                N.check(L.bgs_bounce_pack_results(n, N.ptr(s["res"].length), N.ptr(s["res"].winner), N.ptr(rec),
                                                  N.stream_ptr(torch)))
            else:
                self._view(rec, "length", torch.int16, n).copy_(s["res"].length)
                self._view(rec, "winner", torch.int8, n).copy_(s["res"].winner)
        s["computed"].record()
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(s["computed"])
            s["rec_host"].copy_(rec, non_blocking=True)  # ONE device->host copy per batch
            s["copied"].record()

    def _out(self, s):
        torch = self.torch
        h = s["rec_host"]
        stats = h[self.off["stats"]:].view(torch.int64)
        if self.mode in ("u8", "u8wide"):
            return stats, self._view(h, "result", torch.uint8, self.n)
        if self.mode == "u16":
            return stats, self._view(h, "result", torch.int16, self.n)
        if self.mode == "dense":
            return stats, self._view(h, "result", torch.int16, self.nwords)
        ldt = torch.uint8 if self.game == "connect" else torch.int16
        return stats, self._view(h, "length", ldt, self.n), self._view(h, "winner", torch.int8, self.n)

    def run(self, seed: int, game_id0: int = 0):
        """One end-to-end rollout; returns ``(stats, length, winner)`` -- or ``(stats, result)`` when packed -- as
        views of the pinned host record (synchronised)."""
        s = self.sets[0]
        self._launch(s, seed, game_id0)
        s["copied"].synchronize()
        return self._out(s)

    def stream(self, seed: int, game_id0: int, n_batches: int):
        """Yields the host tensors of ``run`` for ``n_batches`` consecutive batches of ``n_games`` games (global
        ids ``game_id0 + i*n_games ...``).  A yielded set is valid until the next-but-one ``next()``."""
        torch = self.torch
        pending = []
        for i in range(n_batches):
            s = self.sets[i % self.depth]
            # the kernel about to overwrite this set's device buffers must wait for its last copy
            (s["cs"] or torch.cuda.current_stream()).wait_event(s["copied"])
            self._launch(s, seed, game_id0 + i * self.n)
            pending.append(s)
            if len(pending) == self.depth:
                done = pending.pop(0)
                done["copied"].synchronize()
                yield self._out(done)
        for done in pending:
            done["copied"].synchronize()
            yield self._out(done)


class HostLeafRollout:
    """End-to-end leaf evaluation: positions in pinned HOST memory -> rollouts -> per-game results in pinned
    host memory, pipelined (``State.from_json`` positions + the README.md:49-72 loop, for a tree search).

    Positions travel as :class:`ConnectPacked` records (two bitboards + a meta byte: 17 bytes per 6x7 position
    instead of 44); results come back as one packed byte per game + the statistics vector.  ``submit`` enqueues
    one batch (host->device copy, rollout, device->host copy on three streams) and returns a ticket;
    ``result(ticket)`` waits for it.  With ``depth`` tickets in flight the H2D copy of batch i+1, the kernel of
    batch i and the D2H copy of batch i-1 overlap."""

    def __init__(self, config, n_positions: int, depth: int = 3, numa_local: bool = True):
        torch = N.require_cuda()
        self.torch, self.config, self.n, self.depth = torch, config, int(n_positions), int(depth)
        H, W, _ = _hwk(config)
        if H * W > 63:
            raise ValueError("HostLeafRollout packs results in one byte: boards of at most 63 cells")
        n = self.n
        dev = torch.device("cuda", torch.cuda.current_device())
        pw = N.lib().bgs_connect_packed_words(H, W)
        n16 = (n + 15) // 16 * 16
        self.in_bytes = n * pw * 8 + n16
        self.out_off_stats = n16
        self.out_bytes = n16 + N.STATS_LEN * 8
        self.sets = []
        with _numa_local(dev.index, numa_local):
            for _ in range(self.depth):
                in_dev = torch.empty(self.in_bytes, dtype=torch.uint8, device=dev)
                out_dev = torch.zeros(self.out_bytes, dtype=torch.uint8, device=dev)
                self.sets.append({
                    "in_host": torch.empty(self.in_bytes, dtype=torch.uint8).pin_memory(),
                    "out_host": torch.zeros(self.out_bytes, dtype=torch.uint8).pin_memory(),
                    "in_dev": in_dev, "out_dev": out_dev,
                    "packed": in_dev[: n * pw * 8].view(torch.int64).view(n, pw),
                    "meta": in_dev[n * pw * 8: n * pw * 8 + n],
                    "stats_dev": out_dev[n16:].view(torch.int64),
                    "res": None, "loaded": torch.cuda.Event(), "computed": torch.cuda.Event(), "copied": torch.cuda.Event(),
                })
        self.h2d_stream, self.d2h_stream = torch.cuda.Stream(), torch.cuda.Stream()
        self.h2d_bytes, self.d2h_bytes = self.in_bytes, self.out_bytes
        self._next = 0

    def host_inputs(self, ticket_slot: int):
        """(packed int64[n, words], meta uint8[n]) views of the pinned INPUT record of a slot: fill them (e.g.
        from ``ConnectBatch.pack()`` copied to the host) before ``submit``."""
        s = self.sets[ticket_slot]
        n = self.n
        pw = s["packed"].shape[1]
        h = s["in_host"]
        return h[: n * pw * 8].view(self.torch.int64).view(n, pw), h[n * pw * 8: n * pw * 8 + n]

    def submit(self, seed: int, game_id0: int, slot: int | None = None) -> int:
        torch = self.torch
        if slot is None:
            slot = self._next
            self._next = (self._next + 1) % self.depth
        s = self.sets[slot]
        with torch.cuda.stream(self.h2d_stream):
            self.h2d_stream.wait_event(s["computed"])  # the previous rollout that read in_dev is done
            s["in_dev"].copy_(s["in_host"], non_blocking=True)
            s["loaded"].record()
        cur = torch.cuda.current_stream()
        cur.wait_event(s["loaded"])
        cur.wait_event(s["copied"])  # the previous results of this slot have left out_dev
        s["stats_dev"].zero_()
        start = ConnectPacked(self.config, s["packed"], s["meta"])
        s["res"] = connect_rollout(self.config, self.n, seed, game_id0, per_game=True, stats=s["stats_dev"], out=s["res"],
                                   start=start)
        N.check(N.lib().bgs_connect_pack_results(self.n, N.ptr(s["res"].length), N.ptr(s["res"].winner), N.ptr(s["out_dev"]),
                                                  N.stream_ptr(torch)))
        s["computed"].record()
        with torch.cuda.stream(self.d2h_stream):
            self.d2h_stream.wait_event(s["computed"])
            s["out_host"].copy_(s["out_dev"], non_blocking=True)
            s["copied"].record()
        return slot

    def result(self, slot: int):
        """(stats int64[256], result uint8[n]) of a submitted batch, as views of its pinned host record."""
        s = self.sets[slot]
        s["copied"].synchronize()
        h = s["out_host"]
        return h[self.out_off_stats:].view(self.torch.int64), h[: self.n]


@dataclass
class ConnectPacked:
    """n Connect-k positions as two bitboards + one meta byte each (``ConnectBatch.pack()``): ``packed``
    int64[n, 2 or 4] in the packed-board format of ``include/bgs_b200.h``, ``meta`` uint8[n] (bit 0 = side to
    move, bits 1..2 = winner + 1).  ``connect_rollout(start=...)`` accepts it; 17 bytes per 6x7 position cross
    PCIe instead of 44."""

    config: Any
    packed: Any
    meta: Any

    @property
    def n(self) -> int:
        return int(self.meta.shape[0])


def _keys_unique(keys):
    """(unique keys int64[m,2], index of the first state with each key int64[m], inverse int64[n]) -- the rows
    in ascending (signed) lexicographic order of the two key words."""
    import torch

    uniq, inverse = torch.unique(keys, dim=0, return_inverse=True)
    n = keys.shape[0]
    first = torch.full((uniq.shape[0],), n, dtype=torch.int64, device=keys.device)
    first.scatter_reduce_(0, inverse, torch.arange(n, device=keys.device), reduce="amin")
    return uniq, first, inverse


class ConnectBatch:
    """n Connect-k states as tensors: the batched equivalent of the reference's State objects.

    ``grid`` int8[n,H,W] (row 0 bottom; -1 / 0 / 1), ``player`` int8[n], ``winner`` int8[n] (-1 none),
    ``has_ended`` uint8[n], ``legal`` uint32[n] (bit c = column c playable), ``reward`` float32[n,2].
    """

    def __init__(self, config, grid, player, winner, has_ended=None, legal=None, reward=None):
        self.config = config
        self.grid, self.player, self.winner = grid, player, winner
        self.has_ended, self.legal, self.reward = has_ended, legal, reward
        if has_ended is None:
            self._query()

    @property
    def n(self) -> int:
        return self.grid.shape[0]

    @staticmethod
    def initial(config, n: int) -> "ConnectBatch":
        """n copies of ``config.sample_initial_state()`` (reference connect.cpp:32)."""
        torch = N.require_cuda()
        H, W, _ = _hwk(config)
        grid = torch.full((n, H, W), -1, dtype=torch.int8, device="cuda")
        player = torch.zeros(n, dtype=torch.int8, device="cuda")
        winner = torch.full((n,), -1, dtype=torch.int8, device="cuda")
        return ConnectBatch(config, grid, player, winner)

    def _query(self):
        torch = N.require_cuda()
        H, W, _ = _hwk(self.config)
        n = self.n
        self.has_ended = torch.empty(n, dtype=torch.uint8, device=self.grid.device)
        self.legal = torch.empty(n, dtype=torch.int32, device=self.grid.device)
        self.reward = torch.empty((n, 2), dtype=torch.float32, device=self.grid.device)
        N.check(
            N.lib().bgs_connect_query(
                H, W, n, N.ptr(self.grid), N.ptr(self.winner), N.ptr(self.has_ended), N.ptr(self.legal),
                N.ptr(self.reward), N.stream_ptr(torch),
            )
        )

    # -- equality / ordering / hashing of whole batches (reference helper.hpp:10-25) ----------------------
    def key(self):
        """int64[n, 2]: one canonical 128-bit key per state, equal iff the reference's ``==`` holds (same
        grid, player, winner).  Exact and invertible for boards of at most 62 cells, a hash beyond."""
        torch = N.require_cuda()
        H, W, _ = _hwk(self.config)
        keys = torch.empty((self.n, 2), dtype=torch.int64, device=self.grid.device)
        N.check(N.lib().bgs_connect_keys(H, W, self.n, N.ptr(self.grid.contiguous()), N.ptr(self.player.contiguous()),
                                         N.ptr(self.winner.contiguous()), N.ptr(keys), N.stream_ptr(torch)))
        return keys

    def equal(self, other: "ConnectBatch"):
        """bool[n]: ``state_i == other_i`` for every i (also ``batch == other``)."""
        if _hwk(self.config) != _hwk(other.config) or self.n != other.n:
            raise ValueError("batches of different configurations / sizes")
        return (self.key() == other.key()).all(dim=1)

    __eq__ = equal
    __hash__ = None

    def unique(self):
        """Deduplicate: ``(batch of the distinct states, first int64[m], inverse int64[n])`` with
        ``batch[inverse[i]] == self[i]`` -- the set a transposition table would hold."""
        _, first, inverse = _keys_unique(self.key())
        return self.select(first), first, inverse

    def select(self, index) -> "ConnectBatch":
        """The states at ``index`` (an int64 tensor) as a new batch."""
        pick = lambda t: None if t is None else t.index_select(0, index)
        return ConnectBatch(self.config, pick(self.grid), pick(self.player), pick(self.winner), pick(self.has_ended),
                            pick(self.legal), pick(self.reward))

    def pack(self) -> ConnectPacked:
        """Two bitboards + a meta byte per state (17 bytes for 6x7 instead of 44)."""
        torch = N.require_cuda()
        L = N.lib()
        H, W, _ = _hwk(self.config)
        packed = torch.empty((self.n, L.bgs_connect_packed_words(H, W)), dtype=torch.int64, device=self.grid.device)
        meta = torch.empty(self.n, dtype=torch.uint8, device=self.grid.device)
        N.check(L.bgs_connect_pack(H, W, self.n, N.ptr(self.grid.contiguous()), N.ptr(self.player.contiguous()),
                                   N.ptr(self.winner.contiguous()), N.ptr(packed), N.ptr(meta), N.stream_ptr(torch)))
        return ConnectPacked(self.config, packed, meta)

    # -- the reference's JSON wire format for a whole batch (tests/test_connect.py:130-139) ----------------
    def to_json(self) -> list[dict]:
        """``[state.to_json() for state in batch]``: ``{"grid": rows bottom-up, "player", "winner"}``."""
        g = self.grid.cpu().numpy().tolist()
        p = self.player.cpu().numpy().tolist()
        w = self.winner.cpu().numpy().tolist()
        return [{"grid": g[i], "player": int(p[i]), "winner": int(w[i])} for i in range(self.n)]

    @staticmethod
    def from_json(values: list[dict], config) -> "ConnectBatch":
        """``[State.from_json(v, config) for v in values]`` as one batch on the current CUDA device."""
        import numpy as np

        torch = N.require_cuda()
        H, W, _ = _hwk(config)
        grid = np.array([v["grid"] for v in values], dtype=np.int8).reshape(len(values), H, W)
        player = np.array([v["player"] for v in values], dtype=np.int8)
        winner = np.array([v["winner"] for v in values], dtype=np.int8)
        return ConnectBatch(config, torch.from_numpy(grid).cuda(), torch.from_numpy(player).cuda(),
                            torch.from_numpy(winner).cuda())

    def sample_step(self, probs, seed: int, game_id0: int = 0, game_ids=None, draw_index=None):
        """``random.choices(state.actions, weights)`` then ``sample_next_state()`` for every state in ONE kernel
        (reference textual/examples/arena.py:64-68): ``probs`` float32[n, W] are per-column weights (columns
        that are not playable are ignored); the draw is the Philox draw of global game id ``game_id0 + i`` (or
        ``game_ids[i]``) at index ``draw_index[i]`` (default: the number of stones on the board), so equal weights
        replay ``connect_rollout``'s games ply by ply.  Returns ``(next_batch, action int32[n], status)``;
        ``action`` is -1 and ``status`` 1 where the game had already ended."""
        torch = N.require_cuda()
        H, W, K = _hwk(self.config)
        n, dev = self.n, self.grid.device
        probs = probs.to(device=dev, dtype=torch.float32).contiguous()
        if tuple(probs.shape) != (n, W):
            raise ValueError("probs must be float32[n, W]")
        gids = None if game_ids is None else game_ids.to(device=dev, dtype=torch.int64).contiguous()
        didx = None if draw_index is None else draw_index.to(device=dev, dtype=torch.int32).contiguous()
        grid = torch.empty_like(self.grid)
        player = torch.empty_like(self.player)
        winner = torch.empty_like(self.winner)
        ended = torch.empty(n, dtype=torch.uint8, device=dev)
        legal = torch.empty(n, dtype=torch.int32, device=dev)
        reward = torch.empty((n, 2), dtype=torch.float32, device=dev)
        action = torch.empty(n, dtype=torch.int32, device=dev)
        status = torch.empty(n, dtype=torch.int32, device=dev)
        N.check(
            N.lib().bgs_connect_sample_step(
                H, W, K, n, N.ptr(self.grid), N.ptr(self.player), N.ptr(self.winner), N.ptr(probs),
                int(seed) & 0xFFFFFFFFFFFFFFFF, int(game_id0), N.ptr(gids), N.ptr(didx), N.ptr(grid), N.ptr(player),
                N.ptr(winner), N.ptr(ended), N.ptr(reward), N.ptr(legal), N.ptr(action), N.ptr(status), N.stream_ptr(torch),
            )
        )
        return ConnectBatch(self.config, grid, player, winner, ended, legal, reward), action, status

    def legal_mask(self):
        """bool[n, W]: which columns ``state.actions`` would list (reference connect.cpp:43)."""
        torch = N.require_cuda()
        W = _hwk(self.config)[1]
        bits = torch.arange(W, device=self.legal.device, dtype=torch.int32)
        return ((self.legal.unsqueeze(1) >> bits) & 1).bool()

    def step(self, actions):
        """``action_at(col).sample_next_state()`` for every state (reference connect.cpp:44,52).

        Returns ``(next_batch, status)``; ``status`` int32[n] is 0, or 1 where the column was illegal
        or the game had ended (that state is returned unchanged -- the object API raises instead).
        """
        torch = N.require_cuda()
        H, W, K = _hwk(self.config)
        n = self.n
        dev = self.grid.device
        actions = actions.to(device=dev, dtype=torch.int32).contiguous()
        grid = torch.empty_like(self.grid)
        player = torch.empty_like(self.player)
        winner = torch.empty_like(self.winner)
        ended = torch.empty(n, dtype=torch.uint8, device=dev)
        legal = torch.empty(n, dtype=torch.int32, device=dev)
        reward = torch.empty((n, 2), dtype=torch.float32, device=dev)
        status = torch.empty(n, dtype=torch.int32, device=dev)
        N.check(
            N.lib().bgs_connect_step(
                H, W, K, n, N.ptr(self.grid), N.ptr(self.player), N.ptr(self.winner), N.ptr(actions),
                N.ptr(grid), N.ptr(player), N.ptr(winner), N.ptr(ended), N.ptr(reward), N.ptr(legal),
                N.ptr(status), N.stream_ptr(torch),
            )
        )
        return ConnectBatch(self.config, grid, player, winner, ended, legal, reward), status


# -------------------------------------------------------------------------------------------------
# Bounce
# -------------------------------------------------------------------------------------------------
def _bounce_grid(config):
    import numpy as np

    g = config if not hasattr(config, "_grid") else config._grid
    g = np.ascontiguousarray(np.asarray(g), dtype=np.int8)
    if g.ndim != 2:
        raise TypeError("Bounce grid must be a 2-D array")
    return g


def _bounce_check(L, g):
    H, W = g.shape
    if g.min() < 0 or not L.bgs_bounce_supported(H, W, int(g.max())):
        raise RuntimeError(
            f"Bounce {H}x{W} with values up to {int(g.max())} is not supported by the CUDA kernels "
            "(need H*W <= 128, W <= 16, values 0..15)"
        )


def bounce_rollout(
    config,
    n_games: int,
    seed: int = 0,
    game_id0: int = 0,
    *,
    max_plies: int = 512,
    rules: int = 0,
    per_game: bool = True,
    moves: bool = False,
    final_grid: bool = False,
    reward: bool = False,
    stats=None,
    start: "BounceBatch | None" = None,
) -> RolloutResult:
    """Play ``n_games`` uniform-random Bounce games from ``config`` (a ``bounce.Config`` or an int8
    grid) on the current CUDA device -- or, with ``start`` (a :class:`BounceBatch` of ``n_games``
    positions; ``config`` may then be None), from those positions.  Games still running after ``max_plies`` plies are reported
    with ``winner == -2`` and counted in ``stats[5]``.  ``moves`` returns uint8[n, max_plies, 2] =
    (source cell, target cell) with cell = y*W + x."""
    torch = N.require_cuda()
    L = N.lib()
    n = int(n_games)
    if not 0 <= int(max_plies) <= 32767:
        raise ValueError("max_plies must be in 0..32767 (lengths are returned as int16)")
    if start is None:
        g = _bounce_grid(config)
        _bounce_check(L, g)
        H, W = g.shape
    else:
        if start.n != n:
            raise ValueError("start must hold n_games positions")
        H, W = int(start.grid.shape[1]), int(start.grid.shape[2])
        if not L.bgs_bounce_supported(H, W, 0):
            raise RuntimeError(f"Bounce {H}x{W} is not supported by the CUDA kernels (need H*W <= 128, W <= 16)")
    dev = torch.device("cuda", torch.cuda.current_device())
    res = RolloutResult(n_games=n, game_id0=int(game_id0), seed=int(seed), stats=None)
    res.length = torch.empty(n, dtype=torch.int16, device=dev) if per_game else None
    res.winner = torch.empty(n, dtype=torch.int8, device=dev) if per_game else None
    res.actions = torch.empty((n, max_plies, 2), dtype=torch.uint8, device=dev) if moves else None
    res.final_grid = torch.empty((n, H, W), dtype=torch.int8, device=dev) if final_grid else None
    res.reward = torch.empty((n, 2), dtype=torch.float32, device=dev) if reward else None
    if stats is None:
        stats = torch.zeros(N.STATS_LEN, dtype=torch.int64, device=dev)
    res.stats = stats
    res.extra["max_plies"] = int(max_plies)
    if start is None:
        N.check(
            L.bgs_bounce_rollout(
                g.ctypes.data, H, W, int(rules), int(max_plies), n, int(game_id0), int(seed) & 0xFFFFFFFFFFFFFFFF,
                N.ptr(res.actions), N.ptr(res.length), N.ptr(res.winner), N.ptr(res.final_grid), N.ptr(res.reward),
                N.ptr(stats), N.stream_ptr(torch),
            )
        )
    else:
        N.check(
            L.bgs_bounce_rollout_from(
                H, W, int(rules), int(max_plies), n, int(game_id0), int(seed) & 0xFFFFFFFFFFFFFFFF,
                N.ptr(start.grid.contiguous()), N.ptr(start.player.contiguous()), N.ptr(start.winner.contiguous()),
                N.ptr(start.has_ended.contiguous()), N.ptr(res.actions), N.ptr(res.length), N.ptr(res.winner),
                N.ptr(res.final_grid), N.ptr(res.reward), N.ptr(stats), N.stream_ptr(torch),
            )
        )
    return res


class BounceBatch:
    """n Bounce states as tensors: ``grid`` int8[n,H,W], ``player`` int8[n], ``winner`` int8[n],
    ``has_ended`` uint8[n].  ``moves()`` gives the legal actions, ``step()`` the transition."""

    def __init__(self, grid, player, winner, has_ended, rules: int = 0, reward=None, ply=None):
        self.grid, self.player, self.winner, self.has_ended = grid, player, winner, has_ended
        self.rules = int(rules)
        self.reward = reward
        #: int32[n] plies played so far (the draw index of ``sample_step``); None = 0 everywhere
        self.ply = ply

    def key(self):
        """int64[n, 2]: a 128-bit hash key per state, equal iff grid, player and winner are equal
        (reference helper.hpp:10-25; values unpinned)."""
        torch = N.require_cuda()
        n, H, W = self.grid.shape
        keys = torch.empty((n, 2), dtype=torch.int64, device=self.grid.device)
        N.check(N.lib().bgs_bounce_keys(H, W, n, N.ptr(self.grid.contiguous()), N.ptr(self.player.contiguous()),
                                        N.ptr(self.winner.contiguous()), N.ptr(keys), N.stream_ptr(torch)))
        return keys

    def equal(self, other: "BounceBatch"):
        if tuple(self.grid.shape) != tuple(other.grid.shape):
            raise ValueError("batches of different shapes")
        return (self.key() == other.key()).all(dim=1)

    __eq__ = equal
    __hash__ = None

    def select(self, index) -> "BounceBatch":
        pick = lambda t: None if t is None else t.index_select(0, index)
        return BounceBatch(pick(self.grid), pick(self.player), pick(self.winner), pick(self.has_ended), self.rules,
                           pick(self.reward), pick(self.ply))

    def unique(self):
        """``(batch of the distinct states, first int64[m], inverse int64[n])``."""
        _, first, inverse = _keys_unique(self.key())
        return self.select(first), first, inverse

    def to_json(self) -> list[dict]:
        """``[state.to_json() for state in batch]`` (tests/test_bounce.py:385-399)."""
        g = self.grid.cpu().numpy().tolist()
        p = self.player.cpu().numpy().tolist()
        w = self.winner.cpu().numpy().tolist()
        return [{"grid": g[i], "player": int(p[i]), "winner": int(w[i])} for i in range(self.n)]

    @staticmethod
    def from_json(values: list[dict], config=None, rules: int = 0) -> "BounceBatch":
        """``[State.from_json(v, config) for v in values]`` as one batch.  ``has_ended`` is recovered as the
        object API does: a winner, or (for a drawn game, winner -1) a side to move without any action."""
        import numpy as np

        torch = N.require_cuda()
        grid = torch.from_numpy(np.array([v["grid"] for v in values], dtype=np.int8)).cuda()
        player = torch.from_numpy(np.array([v["player"] for v in values], dtype=np.int8)).cuda()
        winner = torch.from_numpy(np.array([v["winner"] for v in values], dtype=np.int8)).cuda()
        b = BounceBatch(grid, player, winner, (winner >= 0).to(torch.uint8), rules)
        _, _, count = b.moves()
        b.has_ended = ((winner >= 0) | (count == 0)).to(torch.uint8)
        return b

    def sample_step(self, probs, seed: int, game_id0: int = 0, game_ids=None, draw_index=None):
        """``random.choices(state.actions, weights)`` + ``sample_next_state()`` in one kernel (reference
        textual/examples/arena.py:64-68).  ``probs`` float32[n, W, H*W]: ``probs[i, sx, ty*W + tx]`` is the weight of
        moving the piece in column ``sx`` of the mover's source row to ``(tx, ty)``.  Draw index = ``draw_index``
        or ``self.ply`` (0 if neither): equal weights replay ``bounce_rollout``.  Returns ``(next_batch,
        move int32[n,4], status)``."""
        torch = N.require_cuda()
        n, H, W = self.grid.shape
        dev = self.grid.device
        probs = probs.to(device=dev, dtype=torch.float32).contiguous()
        if tuple(probs.shape) != (n, W, H * W):
            raise ValueError("probs must be float32[n, W, H*W]")
        didx = draw_index if draw_index is not None else self.ply
        didx = None if didx is None else didx.to(device=dev, dtype=torch.int32).contiguous()
        gids = None if game_ids is None else game_ids.to(device=dev, dtype=torch.int64).contiguous()
        grid = torch.empty_like(self.grid)
        player = torch.empty_like(self.player)
        winner = torch.empty_like(self.winner)
        ended = torch.empty(n, dtype=torch.uint8, device=dev)
        reward = torch.empty((n, 2), dtype=torch.float32, device=dev)
        move = torch.empty((n, 4), dtype=torch.int32, device=dev)
        status = torch.empty(n, dtype=torch.int32, device=dev)
        N.check(
            N.lib().bgs_bounce_sample_step(
                H, W, self.rules, n, N.ptr(self.grid), N.ptr(self.player), N.ptr(self.winner), N.ptr(self.has_ended),
                N.ptr(probs), int(seed) & 0xFFFFFFFFFFFFFFFF, int(game_id0), N.ptr(gids), N.ptr(didx), N.ptr(grid),
                N.ptr(player), N.ptr(winner), N.ptr(ended), N.ptr(reward), N.ptr(move), N.ptr(status), N.stream_ptr(torch),
            )
        )
        base = didx if didx is not None else torch.zeros(n, dtype=torch.int32, device=dev)
        return BounceBatch(grid, player, winner, ended, self.rules, reward, base + (status == 0).to(torch.int32)), move, status

    @property
    def n(self) -> int:
        return self.grid.shape[0]

    @staticmethod
    def initial(config, n: int, rules: int = 0) -> "BounceBatch":
        torch = N.require_cuda()
        g = _bounce_grid(config)
        _bounce_check(N.lib(), g)
        grid = torch.from_numpy(g).cuda().unsqueeze(0).repeat(n, 1, 1).contiguous()
        z = torch.zeros(n, dtype=torch.int8, device="cuda")
        return BounceBatch(grid, z, torch.full_like(z, -1), torch.zeros(n, dtype=torch.uint8, device="cuda"), rules)

    def moves(self):
        """``(source_row int8[n], targets int64[n,W], count int32[n])``: bit ``y*W+x`` of
        ``targets[i, sx]`` is set iff ``(sx, source_row[i]) -> (x, y)`` is legal (reference
        bounce.cpp:40-41).  ``count`` is -1 for a grid with values outside 0..15.  Boards of more than
        64 cells or more than 8 columns use two words per mask: ``targets int64[n,W,2]`` (low, high)."""
        torch = N.require_cuda()
        n, H, W = self.grid.shape
        dev = self.grid.device
        row = torch.empty(n, dtype=torch.int8, device=dev)
        wide = W > 8 or H * W > 64
        targets = torch.empty((n, W, 2) if wide else (n, W), dtype=torch.int64, device=dev)
        count = torch.empty(n, dtype=torch.int32, device=dev)
        N.check(
            N.lib().bgs_bounce_moves(
                H, W, self.rules, n, N.ptr(self.grid), N.ptr(self.player), N.ptr(self.has_ended), N.ptr(row),
                N.ptr(targets), N.ptr(count), N.stream_ptr(torch),
            )
        )
        return row, targets, count

    def step(self, move):
        """``action_at(source, target).sample_next_state()`` for every state; ``move`` is
        int32[n,4] = (sx, sy, tx, ty).  Returns ``(next_batch, status)`` with status 1 = illegal."""
        torch = N.require_cuda()
        n, H, W = self.grid.shape
        dev = self.grid.device
        move = move.to(device=dev, dtype=torch.int32).contiguous()
        grid = torch.empty_like(self.grid)
        player = torch.empty_like(self.player)
        winner = torch.empty_like(self.winner)
        ended = torch.empty(n, dtype=torch.uint8, device=dev)
        reward = torch.empty((n, 2), dtype=torch.float32, device=dev)
        status = torch.empty(n, dtype=torch.int32, device=dev)
        N.check(
            N.lib().bgs_bounce_step(
                H, W, self.rules, n, N.ptr(self.grid), N.ptr(self.player), N.ptr(self.winner), N.ptr(self.has_ended),
                N.ptr(move),
                N.ptr(grid), N.ptr(player), N.ptr(winner), N.ptr(ended), N.ptr(reward), N.ptr(status),
                N.stream_ptr(torch),
            )
        )
        ply = None if self.ply is None else self.ply + (status == 0).to(torch.int32)
        return BounceBatch(grid, player, winner, ended, self.rules, reward, ply), status

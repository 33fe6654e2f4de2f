"""One process per GPU: sharding + the join-key shuffle.

* triples are sharded by `kb_shard_of(subject, world)` = (subject >> 10) % world — block-cyclic on the dense dictionary id, the same
  function the device's `kb_partition` uses; it balances like a hash and keeps each shard's key domain dense (direct join tables
  stay as small on N GPUs as on one: `Context.set_sharding`);
* a star join on the subject needs no communication: every rank joins its shard, counts are summed, times are max-reduced;
* a join on a non-subject key shuffles the rows once. Two implementations of that exchange:
  - `PeerShuffle` (GPU box): ONE kernel partitions and transfers — `kb_shuffle_scatter` writes every row straight into the receive
    buffer of the rank that owns its key, peer memory over NVLink (buffers from torch symmetric memory, which is only plumbing
    here: allocation, address exchange, barrier). Only the world x world matrix of row counts travels through a collective.
  - `shuffle_relation`: `kb_partition` (device) splits a relation into `world` contiguous ranges, `all_to_all_relation` exchanges
    the range sizes and then the columns with `torch.distributed.all_to_all_single` — NCCL on the GPU box, gloo in the CPU tests.

* Datalog over a sharded store: the broadcast plan (`datalog_fixpoint_sharded`) and, for rules that join two sharded predicates, the
  super-step scheme at the end of this file (`exchange_plan`, `ShardedFixpoint`, `run_sharded_fixpoint`): closures and dedup on the
  device (`kb_datalog_fixpoint`, `kb_datalog_fixpoint_seed`), the derived facts of a super-step routed to their home ranks here.

Plumbing lives here; the data-touching work is in libkolibrie_b200.so — with one exception, said where it stands: the routing of a
super-step's derived facts (which rank is home to which row) is numpy on the host, the rows travel as host arrays.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import torch
import torch.distributed as dist

from . import datagen


def shard_of(keys: np.ndarray, world: int) -> np.ndarray:
    """rank owning each key (== kb_shard_of): block-cyclic on the dense dictionary id"""
    return datagen.shard_of_np(keys, world)


def shard_triples(s: np.ndarray, p: np.ndarray, o: np.ndarray, rank: int, world: int):
    """this rank's triples under hash(subject) sharding, store order preserved"""
    if world == 1:
        return s, p, o
    keep = shard_of(s, world) == rank
    return s[keep], p[keep], o[keep]


def all_to_all_relation(cols: Sequence[torch.Tensor], part_offsets: Sequence[int], group=None) -> List[torch.Tensor]:
    """`cols` are equally long 1-D tensors already partitioned into `world` contiguous ranges [part_offsets[r], part_offsets[r+1]).
    Range r goes to rank r. Returns the received columns (ranges concatenated in source-rank order)."""
    world = dist.get_world_size(group)
    assert len(part_offsets) == world + 1
    dev = cols[0].device if cols else torch.device("cpu")
    send_counts = torch.tensor([part_offsets[r + 1] - part_offsets[r] for r in range(world)], dtype=torch.int64, device=dev)
    recv_counts = torch.empty(world, dtype=torch.int64, device=dev)
    dist.all_to_all_single(recv_counts, send_counts, group=group)
    sc = [int(x) for x in send_counts.tolist()]
    rc = [int(x) for x in recv_counts.tolist()]
    out = []
    for c in cols:
        recv = torch.empty(sum(rc), dtype=c.dtype, device=dev)
        dist.all_to_all_single(recv, c.contiguous(), output_split_sizes=rc, input_split_sizes=sc, group=group)
        out.append(recv)
    return out


class _DevPtr:
    """wraps a raw device pointer of the library as a CUDA array (no copy)"""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i4", "data": (ptr, False), "version": 2}


def shuffle_relation(ctx, rel, key_slot: int, group=None):
    """GPU only: re-shard `rel` so that every row lives on rank kb_shard_of(row[key_slot], world). Returns a new Relation."""
    world = dist.get_world_size(group)
    n, slots = rel.info()
    part, offs = ctx.partition(rel, key_slot, world)
    dev = torch.device("cuda", ctx.device)
    cols = [torch.as_tensor(_DevPtr(part.device_ptr(i), max(n, 1)), device=dev)[:n] for i in range(len(slots))]
    recv = all_to_all_relation(cols, offs, group)
    torch.cuda.synchronize(dev)
    m = int(recv[0].numel()) if recv else 0
    out = ctx.rel_from_device(slots, [int(t.data_ptr()) for t in recv], m)
    return out


def shuffle_plan(counts_matrix: np.ndarray, rank: int):
    """counts_matrix[src][dst] = rows rank src sends to rank dst. Returns (base, recv_total): base[d] = first row of this rank's range
    in rank d's receive buffer (ranges in source-rank order), recv_total = rows this rank receives."""
    m = np.asarray(counts_matrix, dtype=np.int64)
    base = [int(m[:rank, d].sum()) for d in range(m.shape[1])]
    return base, int(m[:, rank].sum())


class PeerShuffle:
    """Receive buffers + receive cursor in peer-mapped memory, allocated and rendezvoused once, reused for every shuffle of up to
    `capacity_rows` rows x `n_cols` columns per rank. One exchange = ONE kernel per rank between two barriers: the kernel reserves
    its ranges on the receivers' cursors itself (kb_shuffle_push), so no row counts are computed or exchanged beforehand."""

    def __init__(self, ctx, n_cols: int, capacity_rows: int, group=None):
        import torch.distributed._symmetric_memory as symm_mem

        self.ctx, self.n_cols = ctx, n_cols
        self.cap = (int(capacity_rows) + 63) // 64 * 64  # columns start 256-byte aligned
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.dev = torch.device("cuda", ctx.device)
        # [n_cols columns of cap rows][64 words: word 0 = receive cursor][64 words of slack for the 256-byte over-read of the last column]
        self.buf = symm_mem.empty(n_cols * self.cap + 128, dtype=torch.int32, device=self.dev)
        self.hdl = symm_mem.rendezvous(self.buf, self.group)
        self.ptrs = [int(x) for x in self.hdl.buffer_ptrs]
        self.cursor = self.buf[n_cols * self.cap: n_cols * self.cap + 64]
        self.peer_cols = [self.col_ptr(d, c_) for d in range(self.world) for c_ in range(self.n_cols)]
        self.peer_cursors = [self.ptrs[d] + 4 * self.n_cols * self.cap for d in range(self.world)]
        self.last_bytes_sent = 0

    def col_ptr(self, peer: int, col: int) -> int:
        return self.ptrs[peer] + 4 * col * self.cap

    def _barrier(self):
        self.hdl.barrier()                      # device-side barrier on torch's current stream ...
        torch.cuda.current_stream(self.dev).synchronize()  # ... the library launches on its own stream: order through the host

    def shuffle(self, rel, key_slot: int, copy: bool = False):
        """re-shard `rel` so that every row lives on rank kb_shard_of(row[key_slot], world). Returns a Relation over the receive buffer
        (a VIEW, overwritten by the next shuffle; copy=True for an owned relation)."""
        n, slots = rel.info()
        assert len(slots) == self.n_cols
        self.cursor.zero_()
        self._barrier()  # every cursor is zero and every rank is done reading what the previous shuffle left in its buffer
        self.ctx.shuffle_push(rel, key_slot, self.world, self.peer_cols, self.peer_cursors, self.cap)  # returns after its stores are fenced
        self._barrier()  # every rank's stores have landed
        recv_total = int(self.cursor[0].item())
        ptrs = [self.col_ptr(self.rank, c_) for c_ in range(self.n_cols)]
        if copy:
            return self.ctx.rel_from_device(slots, ptrs, recv_total)
        return self.ctx.rel_wrap_device(slots, ptrs, recv_total)

    def shuffle_planned(self, rel, key_slot: int):
        """the two-pass variant (kb_partition_counts -> all_gather of the count matrix -> kb_shuffle_scatter with fixed ranges):
        deterministic placement, rows in source-rank order; kept for A/B and as the reference for the push variant's tests"""
        n, slots = rel.info()
        mine = torch.tensor(self.ctx.partition_counts(rel, key_slot, self.world), dtype=torch.int64, device=self.dev)
        allc = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(allc, mine, group=self.group)
        base, recv_total = shuffle_plan(torch.stack(allc).cpu().numpy(), self.rank)
        if recv_total > self.cap:
            raise ValueError(f"rank {self.rank} would receive {recv_total} rows, capacity {self.cap}")
        self._barrier()
        self.ctx.shuffle_scatter(rel, key_slot, self.world, self.peer_cols, base, self.cap)
        self._barrier()
        return self.ctx.rel_from_device(slots, [self.col_ptr(self.rank, c_) for c_ in range(self.n_cols)], recv_total)


# ---------------------------------------------------------------------------------------------------------------------
# GROUP BY across ranks: every rank aggregates its shard; the packed partials (kb_groups_pack) are all-gathered and folded by
# kb_groups_merge. One collective: the buffers travel in fixed-size slots whose first 8 bytes carry the payload length.
GROUPS_SLOT_BYTES = 56 + 88 * 4096 + 8  # header + 4096 groups (the capacity of a prepared GROUP BY) + the length prefix


def allgather_bytes(buf: np.ndarray, device=None, group=None, slot_bytes: int = GROUPS_SLOT_BYTES) -> List[np.ndarray]:
    """all-gather of one variable-length byte string per rank (NCCL on `device`, gloo on the CPU). Payloads longer than the slot
    take a second collective sized by the longest one."""
    world = dist.get_world_size(group)
    dev = device or torch.device("cpu")
    buf = np.ascontiguousarray(buf, dtype=np.uint8)

    def gather(slot):
        mine = np.zeros(slot, dtype=np.uint8)
        mine[:8] = np.frombuffer(np.int64(buf.nbytes).tobytes(), dtype=np.uint8)
        m = min(buf.nbytes, slot - 8)
        mine[8:8 + m] = buf[:m]
        t = torch.from_numpy(mine).to(dev)
        out = torch.empty(world * slot, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(out, t, group=group)
        return out.cpu().numpy().reshape(world, slot)

    got = gather(slot_bytes)
    sizes = [int(np.frombuffer(got[r, :8].tobytes(), dtype=np.int64)[0]) for r in range(world)]
    if max(sizes) + 8 > slot_bytes:
        got = gather((max(sizes) + 8 + 255) // 256 * 256)
    return [got[r, 8:8 + sizes[r]].copy() for r in range(world)]


GROUPS_MAGIC = 0x4B4247524F555031  # "KBGROUP1": layout of kb_groups_pack (include/kolibrie_b200.h)
_REC = np.dtype([("keys", "<u4", 4), ("count", "<u8"), ("raw", "<f8", 8)])
assert _REC.itemsize == 88


def pack_groups_np(keys: Sequence[np.ndarray], counts: np.ndarray, raw: Sequence[np.ndarray], kinds: Sequence[int]) -> np.ndarray:
    """the kb_groups_pack byte layout built on the host (gloo tests, tooling): header {magic u64, n_groups u64, n_group_cols u32,
    n_aggs u32, kinds u32[8]} then one 88-byte record {keys u32[4], count u64, raw f64[8]} per group; raw = the accumulator (AVG: the sum)"""
    n = len(counts)
    head = np.zeros(56, dtype=np.uint8)
    head[:16] = np.frombuffer(np.array([GROUPS_MAGIC, n], dtype="<u8").tobytes(), dtype=np.uint8)
    kk = np.zeros(8, dtype="<u4")
    kk[: len(kinds)] = kinds
    head[16:24] = np.frombuffer(np.array([len(keys), len(raw)], dtype="<u4").tobytes(), dtype=np.uint8)
    head[24:56] = np.frombuffer(kk.tobytes(), dtype=np.uint8)
    rec = np.zeros(n, dtype=_REC)
    for c, k in enumerate(keys):
        rec["keys"][:, c] = k
    rec["count"] = counts
    for a, v in enumerate(raw):
        rec["raw"][:, a] = v
    return np.concatenate([head, np.frombuffer(rec.tobytes(), dtype=np.uint8)])


def unpack_groups_np(buf: np.ndarray):
    """inverse of pack_groups_np / kb_groups_pack: (keys [n, n_group_cols], counts [n], raw [n, n_aggs], kinds)"""
    buf = np.ascontiguousarray(buf, dtype=np.uint8)
    magic, n = np.frombuffer(buf[:16].tobytes(), dtype="<u8")
    assert int(magic) == GROUPS_MAGIC, "not a kb_groups_pack buffer"
    ng, na = (int(x) for x in np.frombuffer(buf[16:24].tobytes(), dtype="<u4"))
    kinds = [int(x) for x in np.frombuffer(buf[24:56].tobytes(), dtype="<u4")[:na]]
    rec = np.frombuffer(buf[56:56 + int(n) * 88].tobytes(), dtype=_REC)
    return rec["keys"][:, :ng].copy(), rec["count"].copy(), rec["raw"][:, :na].copy(), kinds


def attach_group_plan(plan, group=None):
    """Give a grouped prepared plan its peers: scratch in torch symmetric memory (plumbing: allocation + address exchange), after which
    plan.submit() merges the GROUP BY across ranks on the device (barrier over peer-memory flags + one merge kernel reading the peers'
    partial tables over NVLink) and plan.collect_groups() returns the global groups on every rank."""
    import torch.distributed._symmetric_memory as symm_mem

    grp = group if group is not None else dist.group.WORLD
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n = (plan.peer_scratch_bytes() + 3) // 4
    buf = symm_mem.empty(n, dtype=torch.int32, device=torch.device("cuda", plan.ctx.device))
    buf.zero_()
    hdl = symm_mem.rendezvous(buf, grp)
    torch.cuda.synchronize()
    hdl.barrier()
    torch.cuda.synchronize()
    plan.attach_peers(rank, world, [int(x) for x in hdl.buffer_ptrs], keep=(buf, hdl))
    return plan


def allgather_groups(packed: np.ndarray, device=None, group=None) -> List[np.ndarray]:
    """the packed partial GROUP BY results of all ranks, in rank order (input of Context.groups_merge)"""
    return allgather_bytes(packed, device, group)


def sum_over_ranks(value: int, device=None, group=None) -> int:
    t = torch.tensor([int(value)], dtype=torch.int64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return int(t[0])


def max_over_ranks(value: float, device=None, group=None) -> float:
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t[0])


# ---------------------------------------------------------------------------------------------------------------------
# Datalog across ranks (config 4 shape): broadcast the small relations, keep the big ones sharded by subject.
def _allgather_rows(rows: np.ndarray, device=None, group=None) -> np.ndarray:
    """concatenation over ranks of an [n, k] uint32 array (n differs per rank)"""
    world = dist.get_world_size(group)
    dev = device or torch.device("cpu")
    n = torch.tensor([rows.shape[0]], dtype=torch.int64, device=dev)
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(x[0]) for x in sizes]
    mx = max(sizes + [1])
    pad = torch.zeros((mx, rows.shape[1]), dtype=torch.int32, device=dev)
    if rows.shape[0]:
        pad[: rows.shape[0]] = torch.from_numpy(rows.astype(np.int32)).to(dev)
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return np.concatenate([o[:m].cpu().numpy().astype(np.uint32) for o, m in zip(out, sizes)], axis=0)


def check_broadcast_plan(rules: Sequence[dict], replicated_preds: Sequence[int]):
    """A rule set can run shard-locally after broadcasting `replicated_preds` when every rule has at most one premise over a
    sharded predicate and, if it has one, its heads keep that premise's subject variable and a sharded head predicate (derived facts
    stay on the rank that owns the subject); rules over replicated predicates only must derive replicated predicates (they are
    evaluated redundantly and identically on every rank). Raises ValueError otherwise."""
    rep = set(int(x) for x in replicated_preds)
    for ri, r in enumerate(rules):
        sharded = [p for p in r["premise"] if p.p.is_var or int(p.p.value) not in rep]
        if len(sharded) > 1:
            raise ValueError(f"rule {ri}: two premises over sharded predicates need a join-key shuffle")
        for h in r["conclusion"]:
            if h.p.is_var:
                raise ValueError(f"rule {ri}: variable head predicate")
            head_rep = int(h.p.value) in rep
            if sharded:
                sp = sharded[0]
                if head_rep or not (h.s.is_var and sp.s.is_var and h.s.value == sp.s.value):
                    raise ValueError(f"rule {ri}: head must keep the sharded premise's subject and a sharded predicate")
            elif not head_rep:
                raise ValueError(f"rule {ri}: a rule over replicated predicates must derive a replicated predicate")


def datalog_fixpoint_sharded(local_fixpoint, s, p, o, rules, replicated_preds, device=None, group=None):
    """`s,p,o`: this rank's shard (hash(subject) sharding). `local_fixpoint(s,p,o) -> [n,3] inferred` runs the fixpoint on one rank
    (the device library on the GPU box, the oracle in the CPU tests). Returns this rank's inferred facts: sharded predicates as
    derived locally, replicated predicates on rank 0 only (every rank derives the same ones)."""
    check_broadcast_plan(rules, replicated_preds)
    rank = dist.get_rank(group)
    rep = np.isin(p, np.asarray(list(replicated_preds), dtype=np.uint32))
    rep_rows = _allgather_rows(np.stack([s[rep], p[rep], o[rep]], axis=1), device, group)
    rep_rows = np.unique(rep_rows, axis=0) if len(rep_rows) else rep_rows
    S = np.concatenate([rep_rows[:, 0], s[~rep]])
    P = np.concatenate([rep_rows[:, 1], p[~rep]])
    O = np.concatenate([rep_rows[:, 2], o[~rep]])
    inferred = np.asarray(local_fixpoint(S, P, O), dtype=np.uint32).reshape(-1, 3)
    is_rep = np.isin(inferred[:, 1], np.asarray(list(replicated_preds), dtype=np.uint32))
    return inferred[~is_rep] if rank != 0 else inferred


# ---------------------------------------------------------------------------------------------------------------------
# Datalog across ranks, general case (SURVEY.md §8e "Datalog fixpoint": local delta joins, derived facts routed to their owners, the
# owner dedups, termination by an all-reduce of the number of facts sent). Rules may join TWO premises over sharded predicates.
#
# Placement. A fact (s, P, o) lives on owner(s) = kb_shard_of(s, world) — its home, the rank that reports it. When a rule joins two
# sharded premises on a variable v that sits in the OBJECT position of a premise over P, facts of P are kept on owner(o) as well
# ("object-homed"): every fact that mentions the value v in a join position is then on owner(v), so the join for that value is
# complete there and nowhere else has to see both sides. Replicated predicates (small ones: the TBox) are on every rank.
# Super-steps. (0) base facts travel to their second homes; (1) every rank closes its local facts under the rules
# (kb_datalog_fixpoint) and sends each derived fact to its homes; (k >= 2) the facts that arrived are the SEED of an incremental
# closure (kb_datalog_fixpoint_seed: the known-fact set tells which of them are new, those alone are the first delta, the store is
# OLD); what that derives is sent on. The loop ends when no rank sent anything. A rank that holds a fact it does not own may derive
# from it too: sound (every stored fact is in the closure), redundant at worst — the owners' known-fact sets drop the repeats.
def exchange_plan(rules: Sequence[dict], replicated_preds: Sequence[int] = ()):
    """Returns (object_homed, replicated): the predicates whose facts need a second home on owner(object) for `rules` (compiled
    rules: kolibrie_b200.engine.compile_rule), and the replicated ones. Raises ValueError for a rule the placement cannot serve:
    more than two premises over sharded predicates, two that share no variable, or a sharded premise under a replicated head."""
    rep = set(int(x) for x in replicated_preds)
    homed = set()
    for ri, r in enumerate(rules):
        if any(p.p.is_var for p in r["premise"]):
            continue  # a variable predicate never matches (join_algorithm.rs:515-521): the rule derives nothing
        sharded = [p for p in r["premise"] if int(p.p.value) not in rep]
        for h in r["conclusion"]:
            if h.p.is_var:
                raise ValueError(f"rule {ri}: variable head predicate")
            if sharded and int(h.p.value) in rep:
                raise ValueError(f"rule {ri}: a rule with a sharded premise must not derive a replicated predicate")
        if len(sharded) > 2:
            raise ValueError(f"rule {ri}: {len(sharded)} premises over sharded predicates (at most two: their join needs one exchange per extra premise)")
        if len(sharded) == 2:
            a, b = sharded
            best = None
            for pos_a, ta in (("s", a.s), ("o", a.o)):
                for pos_b, tb in (("s", b.s), ("o", b.o)):
                    if ta.is_var and tb.is_var and ta.value == tb.value:
                        need = set()
                        if pos_a == "o":
                            need.add(int(a.p.value))
                        if pos_b == "o":
                            need.add(int(b.p.value))
                        if best is None or len(need - homed) < len(best - homed):
                            best = need
            if best is None:
                raise ValueError(f"rule {ri}: its two sharded premises share no variable")
            homed |= best
    return homed, rep


class ShardedFixpoint:
    """One rank of the exchange scheme above. `engine` does the data work on this rank's facts:
    `load(rows)`, `closure() -> inferred rows`, `closure_seed(rows) -> (accepted rows, inferred rows)` — `DeviceFixpointEngine` on a
    GPU, an oracle-backed stand-in in the CPU tests. Rows are [n, 3] uint32 arrays (s, p, o). Drive it with `run_sharded_fixpoint`."""

    def __init__(self, engine, rank: int, world: int, rules: Sequence[dict], replicated_preds: Sequence[int] = ()):
        self.engine, self.rank, self.world = engine, int(rank), int(world)
        self.homed, self.rep = exchange_plan(rules, replicated_preds)
        self.mine: List[np.ndarray] = []  # the inferred facts this rank reports
        self.steps = 0
        self.sent_rows = 0
        self._loaded = False
        self._base: Optional[np.ndarray] = None

    def _isin(self, preds: np.ndarray, which: set) -> np.ndarray:
        return np.isin(preds, np.fromiter(which, dtype=np.uint32, count=len(which))) if which else np.zeros(len(preds), dtype=bool)

    def _route(self, rows: np.ndarray, base: bool) -> List[np.ndarray]:
        """per destination rank, the rows to send there (never to this rank itself)"""
        parts = [np.empty((0, 3), np.uint32) for _ in range(self.world)]
        if len(rows) == 0 or self.world == 1:
            return parts
        is_rep = self._isin(rows[:, 1], self.rep)
        home_s = shard_of(rows[:, 0], self.world)
        home_o = np.where(self._isin(rows[:, 1], self.homed), shard_of(rows[:, 2], self.world), home_s)
        for d in range(self.world):
            if d == self.rank:
                continue
            # base facts: this rank is their subject home, so only the second home and the replicas are missing;
            # derived facts of replicated predicates are derived on every rank alike (exchange_plan) and never travel
            to_d = ((home_s == d) | (home_o == d)) & ~is_rep
            if base:
                to_d |= is_rep
            parts[d] = np.ascontiguousarray(rows[to_d])
        return parts

    def _report(self, rows: np.ndarray):
        if len(rows) == 0:
            return
        is_rep = self._isin(rows[:, 1], self.rep)
        keep = np.where(is_rep, self.rank == 0, shard_of(rows[:, 0], self.world) == self.rank)
        if keep.any():
            self.mine.append(rows[keep])

    def start(self, s: np.ndarray, p: np.ndarray, o: np.ndarray) -> List[np.ndarray]:
        """super-step 0: `s, p, o` = this rank's subject shard of the base facts; returns what the other ranks need of it"""
        rows = np.stack([np.asarray(s, np.uint32), np.asarray(p, np.uint32), np.asarray(o, np.uint32)], axis=1)
        if self.world > 1 and not (shard_of(rows[:, 0], self.world) == self.rank).all():
            raise ValueError("the base facts of a rank must be its subject shard (dist.shard_triples)")
        self._base = rows
        parts = self._route(rows, base=True)
        self.sent_rows += sum(len(x) for x in parts)
        return parts

    def step(self, received: np.ndarray) -> List[np.ndarray]:
        """one super-step: takes what arrived, runs the engine, returns the derived facts per destination"""
        received = np.asarray(received, np.uint32).reshape(-1, 3)
        self.steps += 1
        if not self._loaded:
            self.engine.load(np.concatenate([self._base, received], axis=0))
            self._base = None
            self._loaded = True
            inferred = self.engine.closure()
        else:
            accepted, inferred = self.engine.closure_seed(received)
            self._report(accepted)
        self._report(inferred)
        parts = self._route(inferred, base=False)
        self.sent_rows += sum(len(x) for x in parts)
        return parts

    def result(self) -> np.ndarray:
        """the inferred facts this rank reports: those whose subject it owns; replicated predicates on rank 0 only"""
        return np.concatenate(self.mine, axis=0) if self.mine else np.empty((0, 3), np.uint32)


class DeviceFixpointEngine:
    """ShardedFixpoint's engine on one GPU: the closures are kb_datalog_fixpoint / kb_datalog_fixpoint_seed over the context's store."""

    def __init__(self, ctx, rules: Sequence[dict], strategy: int = 0):
        self.ctx, self.rules, self.strategy = ctx, list(rules), strategy
        self.device_ms = 0.0

    def load(self, rows: np.ndarray):
        self.ctx.store_load(rows[:, 0], rows[:, 1], rows[:, 2])

    def closure(self) -> np.ndarray:
        rel, st = self.ctx.datalog_fixpoint(self.rules, self.strategy)
        self.device_ms += st.device_ms
        out = rel.to_numpy([0, 1, 2])
        rel.free()
        return out

    def closure_seed(self, rows: np.ndarray):
        if len(rows) == 0:
            e = np.empty((0, 3), np.uint32)
            return e, e
        seed = self.ctx.rel_from_host([0, 1, 2], [np.ascontiguousarray(rows[:, k]) for k in range(3)])
        rel, n_new, st = self.ctx.datalog_fixpoint_seed(self.rules, seed, self.strategy)
        self.device_ms += st.device_ms
        out = rel.to_numpy([0, 1, 2])
        rel.free()
        seed.free()
        return out[:n_new], out[n_new:]


def exchange_rows(parts: Sequence[np.ndarray], device=None, group=None) -> np.ndarray:
    """variable-size all-to-all of [n, 3] uint32 row blocks: parts[d] goes to rank d; returns what arrived (source-rank order)"""
    world = dist.get_world_size(group)
    dev = device or torch.device("cpu")
    offs = [0]
    for r in range(world):
        offs.append(offs[-1] + len(parts[r]))
    flat = np.concatenate([np.asarray(x, np.uint32).reshape(-1, 3) for x in parts], axis=0) if offs[-1] else np.empty((0, 3), np.uint32)
    cols = [torch.from_numpy(np.ascontiguousarray(flat[:, k]).view(np.int32)).to(dev) for k in range(3)]
    recv = all_to_all_relation(cols, offs, group)
    return np.stack([t.cpu().numpy().view(np.uint32) for t in recv], axis=1) if len(recv[0]) else np.empty((0, 3), np.uint32)


def run_sharded_fixpoint(node: ShardedFixpoint, s, p, o, device=None, group=None) -> np.ndarray:
    """Drives one rank's `ShardedFixpoint` with torch.distributed (NCCL on the GPU box, gloo in the CPU tests): exchange, step, and an
    all-reduce of the rows sent per super-step as the termination test. Returns this rank's share of the inferred facts."""
    parts = node.start(s, p, o)
    parts = node.step(exchange_rows(parts, device, group))
    while sum_over_ranks(sum(len(x) for x in parts), device, group) > 0:
        parts = node.step(exchange_rows(parts, device, group))
    return node.result()


def run_sharded_fixpoint_local(nodes: Sequence[ShardedFixpoint], shards: Sequence[tuple]) -> List[np.ndarray]:
    """The same loop for `world` nodes living in ONE process (one context per node on the same GPU, or oracle engines): the exchange is
    a transpose of the outgoing lists. Used by the single-GPU tests of the scheme; `shards[r]` = (s, p, o) of rank r."""
    world = len(nodes)

    def transpose(out):
        return [np.concatenate([out[src][dst] for src in range(world)], axis=0) for dst in range(world)]

    out = [nodes[r].start(*shards[r]) for r in range(world)]
    inbox = transpose(out)
    out = [nodes[r].step(inbox[r]) for r in range(world)]
    while sum(len(x) for parts in out for x in parts) > 0:
        inbox = transpose(out)
        out = [nodes[r].step(inbox[r]) for r in range(world)]
    return [n.result() for n in nodes]

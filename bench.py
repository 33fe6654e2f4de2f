#!/usr/bin/env python
"""bench.py — joined bindings/s of the 3-pattern BGP hot path (BASELINE.json metric) on B200.

A "step" = one pass of the hot path over the resident store: ONE fused TMA scan of all three patterns (FILTER pushed into
the salary pattern) -> two direct hash builds -> one fused multiway probe that emits the joined bindings.
Workload (config.workload): the BASELINE configs[1] query (`?e foaf:title ?t . ?e ds:annual_salary ?s . ?e foaf:name ?n
FILTER(?s > 100000)`) on the employee shape scaled to the size the metric is quoted on: 16 666 667 employees = 100 000 002
dictionary-encoded triples PER GPU (weak scaling: with N GPUs the global dataset has N x that, sharded by kb_shard_of(subject, N);
a subject-star join needs no exchange).

  value      bindings/s with the store already resident in HBM (device path only), whole job over all ranks
  e2e        same metric through the one-shot C-ABI call with HOST (pinned) buffers: upload of the triple columns, the join,
             and the download of the binding columns are all inside the timed region
  roofline   algorithmic bytes (SURVEY.md §8d formulas) / CUDA-event time of the dominant kernel family, vs the measured HBM peak
  cpu_baseline / --impl reference: the oracle's restatement of the reference's own algorithm, timed on the host cores
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "joined bindings/sec on 3-pattern BGP over 100M triples"
UNIT = "bindings/s"
DEFAULT_EMPLOYEES = 16_666_667  # x6 = 100 000 002 triples


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--employees", type=int, default=DEFAULT_EMPLOYEES, help="employees per GPU (6 triples each)")
    ap.add_argument("--query", default="cfg2", choices=["cfg2", "star3", "cfg3", "cfg1"])
    ap.add_argument("--cpu-sample", type=int, default=300_000, help="employees in the bounded CPU sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--numa", action="store_true", help="bind the rank to its GPU's NUMA node (helps the e2e leg at 8 ranks: 28 vs 33 ms per "
                    "step; off by default: the one 8-rank run with it on also showed a 3x slower host side of the resident step)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-index", action="store_true", help="headline on the store-scanning path (no predicate-partitioned index)")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """SM clocks / throttle reasons of the job's GPUs during the timed region (B200_PROFILING.md recipe). Rank 0 samples every GPU of
    the job through NVML in-process; one `nvidia-smi` subprocess per rank every 0.2 s (the first version) initialises NVML for all
    eight GPUs each time and takes driver locks next to the ranks' launches — at 8 ranks that alone tripled the host side of a
    0.1 ms step. Falls back to nvidia-smi when pynvml is missing."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    BITS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, indices):
        super().__init__(daemon=True)
        self.indices = list(indices)
        self.samples = []  # (sm_mhz, max_mhz, [reasons])
        self.stop_flag = threading.Event()
        self.nvml = None
        try:
            if not self.indices:
                raise RuntimeError("nothing to sample")
            import pynvml

            pynvml.nvmlInit()
            self.handles = [self._handle(pynvml, i) for i in self.indices]
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    @staticmethod
    def _handle(n, cuda_index):
        """NVML handle of CUDA device `cuda_index`: by PCI bus id, so that a CUDA_VISIBLE_DEVICES remapping cannot make the sampler
        watch somebody else's (idle) GPU; by index when torch does not expose the bus id"""
        try:
            import torch

            pr = torch.cuda.get_device_properties(cuda_index)
            bus = "%08X:%02X:%02X.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
            return n.nvmlDeviceGetHandleByPciBusId(bus.encode())
        except Exception:
            return n.nvmlDeviceGetHandleByIndex(cuda_index)

    def _sample_nvml(self):
        n = self.nvml
        for h in self.handles:
            sm = n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)
            mx = n.nvmlDeviceGetMaxClockInfo(h, n.NVML_CLOCK_SM)
            get = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(n, "nvmlDeviceGetCurrentClocksThrottleReasons")
            mask = int(get(h))
            self.samples.append((float(sm), float(mx), [k for k, b in self.BITS.items() if mask & b]))

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", ",".join(str(i) for i in self.indices)],
                             stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=5).stdout.strip()
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            self.samples.append((float(f[0]), float(f[1]), [k for k, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7])
                                                                 if v.lower().startswith("active")]))

    def run(self):
        while not self.stop_flag.is_set():
            try:
                self._sample_nvml() if self.nvml else self._sample_smi()
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        sm = [s[0] for s in self.samples]
        mx = [s[1] for s in self.samples]
        reasons = sorted({r for s in self.samples for r in s[2]})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm),
                "gpus_sampled": len(self.indices), "via": "nvml" if self.nvml else "nvidia-smi"}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def traffic_from_profiles(family):
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(family)
    except Exception:
        return None


# ---------------------------------------------------------------------------------------------------------------------
def cpu_reference_run(args, steps, warmup):
    """The reference's own algorithm for this query, restated (oracle 'faithful' mode): StarJoin plan (optimizer.rs:84-152) =
    index scan of the first pattern, then per binding one bound index lookup per remaining pattern in the reference's SEQUENTIAL
    mode (engine.rs:621-655: results > 10 000), rows of maps, then the FILTER. Indexes are built before timing, as the reference's
    harnesses do (n_triple_10M.rs:79-95). Result caps (quirk Q1) off. Also times the oracle's columnar OpenMP mode as the strong
    CPU competitor."""
    from kolibrie_b200 import datagen
    from tests import oracle_api as O

    E = min(args.cpu_sample, args.employees)
    d = datagen.employee_dataset(E)
    js, pats, filt = datagen.employee_queries(d)[args.query]
    db = O.Db(d.s, d.p, d.o, d.num_or0, d.is_num)
    db.build_index()
    O.set_threads(O.usable_cpus())  # the CPU arm gets every host thread this process may use (affinity mask and cgroup quota)
    cores = O.num_threads()
    rows = 0
    for _ in range(max(1, min(warmup, 2))):
        rows = db.bgp(pats, filt, mode=1).n_rows
    t0 = time.perf_counter()
    n_done = 0
    for _ in range(steps):
        rows = db.bgp(pats, filt, mode=1).n_rows
        n_done += 1
        if time.perf_counter() - t0 > 120:  # bounded: never more than ~2 minutes of CPU work
            break
    dt = (time.perf_counter() - t0) / n_done
    t1 = time.perf_counter()
    reps = 0
    while reps < 3:
        rows_c = db.bgp(pats, filt, mode=0).n_rows
        reps += 1
    dt_c = (time.perf_counter() - t1) / reps
    assert rows_c == rows
    return {
        "value": rows / dt, "unit": UNIT, "cores": cores, "kind": "port",
        "sample": f"{E} employees = {6 * E} triples, query {args.query}, {n_done} steps; oracle faithful mode (reference StarJoin, sequential mode as engine.rs:621 "
                  f"dictates above 10 000 rows => 1 worker thread; FILTER stage on {cores} threads); indexes prebuilt",
        "ms_per_step": dt * 1e3, "rows_per_step": int(rows),
        "columnar_openmp": {"value": rows / dt_c, "unit": UNIT, "cores": cores, "note": "oracle columnar mode (OpenMP scan + hash joins on u32 columns), same sample"},
    }, n_done


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base, n_done = cpu_reference_run(args, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": n_done, "warmup": args.warmup,
        "ms_per_step": base["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": workload_name(args), "sample": base["sample"]},
        "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "columnar_openmp": base["columnar_openmp"],
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def bind_near_gpu(local):
    """Run this rank (and first-touch the pinned buffers it allocates) on the NUMA node its GPU hangs off: host<->device copies of the
    e2e leg then stay on one socket. Returns (previous affinity, node or None); placement only, no effect on results."""
    try:
        prev = os.sched_getaffinity(0)
    except (AttributeError, OSError):
        return None, None
    try:
        import torch

        pr = torch.cuda.get_device_properties(local)
        bus = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
    except Exception:
        try:
            import pynvml

            pynvml.nvmlInit()
            bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(local)).busId
            bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()[-12:]
        except Exception:
            return prev, None
    try:
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return prev, None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= prev
        if not cpus:
            return prev, None
        os.sched_setaffinity(0, cpus)
        return prev, node
    except (OSError, ValueError):
        return prev, None


def workload_name(args):
    q = {"cfg2": "3-pattern star BGP (?e title ?t . ?e annual_salary ?s . ?e name ?n) + FILTER(?s > 100000)",
         "star3": "3-pattern star BGP (?e title ?t . ?e annual_salary ?s . ?e name ?n), no FILTER",
         "cfg3": "4-pattern star BGP", "cfg1": "2-pattern BGP (?p workplaceHomepage ?c . ?p name ?n)"}[args.query]
    return f"employee shape, {args.employees} employees = {6 * args.employees} triples per GPU, {q}"


# ---------------------------------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    import torch
    import torch.distributed as dist

    from kolibrie_b200 import capi as c
    from kolibrie_b200 import datagen

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: kolibrie_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    prev_affinity, numa_node = bind_near_gpu(local) if args.numa else (None, None)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize()

    # ---- setup (untimed): this rank's shard of the global dataset, pinned on the host and resident on the device
    t_gen = time.perf_counter()
    d = datagen.employee_shard(args.employees * world, rank, world)
    n = d.n_triples
    hs, hp, ho = (torch.from_numpy(x).pin_memory() for x in (d.s, d.p, d.o))
    t_gen = time.perf_counter() - t_gen
    ctx = c.Context(local)
    ctx.set_sharding(rank, world)
    ctx.dict_numeric_load(d.num_or0, d.is_num)
    ctx.store_load(d.s, d.p, d.o)
    # SparqlDatabase::build_all_indexes, once, outside the timed region (the reference's harnesses do the same, n_triple_10M.rs:91-95)
    n_pred, index_ms = (0, 0.0) if args.no_index else ctx.build_index()
    js, pats, filt = datagen.employee_queries(d)[args.query]

    def step_resident():
        r = ctx.star_join(js, pats, filt)
        rows = r.n_rows
        r.free()
        return rows

    # rank 0 samples all GPUs of the job through warm-up, the timed region and the e2e leg (all of it is load)
    sampler = ClockSampler(range(world) if rank == 0 else [])
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        rows_step = step_resident()
    ctx.get_stats(reset=True)
    ctx.set_timing(True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        rows_step = step_resident()
    ctx.synchronize()
    barrier()
    dt = time.perf_counter() - t0
    st = ctx.get_stats(reset=True)
    # the same K steps on the store-SCANNING path (index switched off): the K_scan / K_build / K_probe numbers of SURVEY.md §8(d)
    scan_leg = None
    if not args.no_index:
        ctx.set_use_index(False)
        for _ in range(3):
            step_resident()
        ctx.get_stats(reset=True)
        barrier()
        t0s = time.perf_counter()
        for _ in range(args.steps):
            rows_scan = step_resident()
        ctx.synchronize()
        barrier()
        dts = time.perf_counter() - t0s
        st_scan = ctx.get_stats(reset=True)
        assert rows_scan == rows_step
        scan_leg = (dts, st_scan)
        ctx.set_use_index(True)
    ctx.set_timing(False)

    # ---- e2e: host (pinned) buffers in, host (pinned) buffers out, through kb_star_join_host
    e2e = None
    if not args.no_e2e:
        n_out_cols = len({t.value for pt in pats for t in (pt.s, pt.p, pt.o) if t.is_var})
        outs = [torch.empty(max(rows_step, 1) + 16, dtype=torch.int32).pin_memory() for _ in range(n_out_cols)]

        def step_e2e():
            return c.star_join_host_raw(ctx, hs.data_ptr(), hp.data_ptr(), ho.data_ptr(), n, js, pats, filt, [o.data_ptr() for o in outs], outs[0].numel())

        for _ in range(max(1, min(args.warmup, 3))):
            rows_e, slots_e = step_e2e()
        barrier()
        t1 = time.perf_counter()
        for _ in range(args.steps):
            rows_e, slots_e = step_e2e()
        ctx.synchronize()
        barrier()
        dt_e = time.perf_counter() - t1
        assert rows_e == rows_step, (rows_e, rows_step)
        e2e = {"dt": dt_e, "h2d": 3 * 4 * n, "d2h": len(slots_e) * 4 * rows_e}
        ctx.get_stats(reset=True)
    if args.no_e2e or args.steps * 0.03 < 1.0:  # keep the GPU under the same load until nvidia-smi has a few samples
        t_s = time.perf_counter()
        while time.perf_counter() - t_s < 1.2:
            step_resident() if args.no_e2e else step_e2e()
    sampler.stop_flag.set()
    if rank == 0:
        sampler.join(timeout=2)

    # ---- reduce over ranks: max time, sum of rows
    if world > 1:
        t = torch.tensor([dt, e2e["dt"] if e2e else 0.0, scan_leg[0] if scan_leg else 0.0], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        r = torch.tensor([rows_step, n], dtype=torch.int64, device=dev)
        dist.all_reduce(r, op=dist.ReduceOp.SUM)
        dt, dte = float(t[0]), float(t[1])
        if scan_leg:
            scan_leg = (float(t[2]), scan_leg[1])
        rows_all, n_all = int(r[0]), int(r[1])
    else:
        dte = e2e["dt"] if e2e else 0.0
        rows_all, n_all = rows_step, n
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    ms_step = dt / args.steps * 1e3
    value = rows_all / (dt / args.steps)
    peak, peak_src = measured_peak()

    # ---- roofline of each kernel family (rank 0's launches): algorithmic bytes per SURVEY.md §8(d)
    K = args.steps
    E_loc = d.n_employees
    n_pat = len(pats)
    m_rows = []
    for k in range(n_pat):
        filtered = bool(filt) and k == 1  # cfg2: the salary pattern carries the FILTER
        m_rows.append(rows_step if filtered else E_loc)
    probe_k = max((k for k in range(n_pat) if not (bool(filt) and k == 1)), key=lambda k: m_rows[k])
    builds = [k for k in range(n_pat) if k != probe_k]
    T = len(builds)
    b_probe = 4 * 2 * m_rows[probe_k] + 8 * T * m_rows[probe_k] + 4 * (n_pat + 1) * rows_step
    kernel_names = {"scan": "kb::scan_kernel<K>", "scan+build": "kb::scan_kernel<K> (fused scan + direct-table build) + table memsets",
                    "build": "kb::build_pairs_filtered_kernel / build_direct_pairs_kernel + table memsets", "probe": "kb::probe_fast_kernel<T>"}

    def families(stx, indexed):
        if indexed:
            # build sides are read from their predicate slices: 8 B per slice row in, (filtered) rows into the table: 8*M_slice + 8*M_b
            b_build = sum(8 * E_loc + 8 * m_rows[k] for k in builds)
            fam = {"probe": {"alg_bytes": b_probe, "ms": stx["probe_ms"] / K, "launches_per_step": stx["probe_launches"] / K,
                             "note": "probe rows = one predicate slice of the index (zero copy); lookups go to direct tables"}}
            if stx["build_launches"] > 0:  # a pattern without a persistent table in the index is built per query
                fam["build"] = {"alg_bytes": b_build, "ms": stx["build_ms"] / K, "launches_per_step": stx["build_launches"] / K,
                                "note": "index path: K_build reads the predicate slice (8 B/row) and evaluates the pushed-down FILTER itself"}
            else:
                fam["probe"]["note"] += "; all build sides are persistent per-predicate tables of the index (the reference's spo[s][P] lookup): no per-query build" 
        else:
            b_scan = 12 * n + sum(4 * 2 * m for m in m_rows)
            b_build = sum(16 * m_rows[k] for k in builds)
            if stx.get("fused_scan_builds", 0) > 0:
                fam = {"scan+build": {"alg_bytes": b_scan + b_build, "ms": (stx["scan_ms"] + stx["build_ms"]) / K, "launches_per_step": stx["scan_launches"] / K,
                                      "note": "one kernel does K_scan and K_build of SURVEY.md 8(d): build-side patterns insert into their direct tables"},
                       "probe": {"alg_bytes": b_probe, "ms": stx["probe_ms"] / K, "launches_per_step": stx["probe_launches"] / K}}
            else:
                fam = {"scan": {"alg_bytes": b_scan, "ms": stx["scan_ms"] / K, "launches_per_step": stx["scan_launches"] / K},
                       "build": {"alg_bytes": b_build, "ms": stx["build_ms"] / K, "launches_per_step": stx["build_launches"] / K},
                       "probe": {"alg_bytes": b_probe, "ms": stx["probe_ms"] / K, "launches_per_step": stx["probe_launches"] / K}}
        for v in fam.values():
            v["achieved_gbs"] = v["alg_bytes"] / (v["ms"] * 1e-3) / 1e9 if v["ms"] > 0 else None
            v["frac"] = v["achieved_gbs"] / peak if v["achieved_gbs"] else None
        return fam

    def roof(fam, stx):
        dom = max(fam, key=lambda k: fam[k]["ms"])
        name = "kb::probe_index_kernel<T,PRE>" if (dom == "probe" and stx.get("index_joins", 0) and stx["build_launches"] == 0) else kernel_names[dom]
        return {"bound": "hbm", "kernel": name, "achieved": fam[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s", "frac": fam[dom]["frac"],
                "traffic": traffic_from_profiles(dom.split("+")[0] + ("_index" if stx.get("index_joins", 0) else "")), "peak_source": peak_src,
                "alg_bytes_per_launch": fam[dom]["alg_bytes"], "ms_per_launch": fam[dom]["ms"], "families": fam, "device_ms_per_step": stx["total_ms"] / K}

    indexed = st.get("index_joins", 0) > 0
    roofline = roof(families(st, indexed), st)
    scan_path = None
    if scan_leg is not None:
        dts, st_scan = scan_leg
        scan_path = {"value": rows_all / (dts / K), "unit": UNIT, "ms_per_step": dts / K * 1e3, "gpu_launches": int(st_scan["kernel_launches"]),
                     "roofline": roof(families(st_scan, False), st_scan),
                     "note": "same K steps with the index switched off: every step scans the 12-byte/triple store (the e2e leg always does)"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": workload_name(args), "triples_total": n_all, "bindings_per_step": rows_all, "sharding": "kb_shard_of(subject) = (id >> 10) % n_gpus (block-cyclic on dense ids), no data-path collective",
                   "l2": "inputs per step (index path: 0.4 GB of predicate slices + 0.13 GB of tables; scan path: 1.2 GB of triple columns) exceed the 126 MB L2; no explicit flush",
                   "store": ("predicate-partitioned index built ONCE at load by kb_store_build_index (= SparqlDatabase::build_all_indexes), %d predicates, %.1f ms, outside the timed region"
                             % (n_pred, index_ms)) if not args.no_index else "unindexed: every step scans the store",
                   "datagen_s": round(t_gen, 1), "host_numa_node": numa_node,
                   "timing": "wall clock around K steps between barrier+synchronize, max over ranks; every step ends with a stream sync inside the library"},
        "roofline": roofline,
        "gpu_launches": int(st["kernel_launches"]),
        "clocks": sampler.summary(),
    }
    if scan_path:
        line["scan_path"] = scan_path
    if e2e:
        line["e2e"] = {"value": rows_all / (dte / args.steps), "unit": UNIT, "h2d_bytes_per_step": e2e["h2d"], "d2h_bytes_per_step": e2e["d2h"],
                       "ms_per_step": dte / args.steps * 1e3, "api": "kb_star_join_host (pinned host columns in, pinned host binding columns out; chunked upload overlapped with the scan)"}
    if world == 1 and not args.no_cpu:
        if prev_affinity:
            os.sched_setaffinity(0, prev_affinity)  # the CPU arm gets every host thread back
        base, _ = cpu_reference_run(args, steps=5, warmup=1)
        line["cpu_baseline"] = {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")}
        line["cpu_columnar_openmp"] = base["columnar_openmp"]
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Headline benchmark: top-1000 QPS of the retrieval hot path on synthetic MS-MARCO-shaped data (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload both|sparse|dense] [--dim D]

A "step" is one pass of the hot path over the whole query batch (6,980 queries, top-1000) against the index resident
in HBM.  The default run measures BOTH halves of the metric ("sparse & dense") and prints ONE JSON line on rank 0:

  * top level = BASELINE.json configs[1]: sparse inverted-index retrieval over 8,841,823 synthetic docs x 128,256 terms;
  * "dense"   = configs[2]: flat inner-product search over 8,841,823 x 2048 bf16 (tcgen05 GEMM + fused top-k), a complete
    sub-record with its own value / ms_per_step / e2e / roofline / clocks / cpu_baseline;
  * "dense_4096" = configs[3] (8,841,823 x 4096, corpus sharded over the GPUs) when N >= 8.

For N > 1 (torchrun, one rank per GPU) the corpus is sharded by doc-id range, every rank searches its shard and the
per-shard top-k rows are merged after an NCCL exchange of packed candidate keys ("strong" scaling: total work fixed).
Timing: CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks; the index / corpus
(GBs) is far larger than L2 (126 MB), so no explicit L2 flush is needed between steps.

Per workload the line reports
  value     device-resident QPS (inputs already in HBM),
  e2e       the same search through the class-API call with HOST buffers (pinned h2d of the queries + d2h of the rows),
  e2e_api   the REFERENCE-FACING method itself: SparseRetrieval._sparse_retrieve_multithreaded(query vecs, qids, ...) /
            DenseFlatIndexer.search_knn(query_reps, k) — what eval_sparse.py / eval_dense.py call — wall clock, plus the
            time retrieve() additionally spends writing run.json,
  result_digest  sha256 over the merged ids + score bits: equal digests at N = 1/2/4/8 prove sharded == unsharded.

`--impl reference` times the CPU port of the reference's path (sparse: oracle/sparse_oracle.c, the numba scorer +
argpartition restated in C + OpenMP; dense: the fp32 IndexFlatIP restatement on torch/MKL) with ALL host cores on a
bounded sample of the same workload; only that leg and the cpu_baseline leg execute anything under oracle/.
"""
import argparse
import hashlib
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# torch.distributed.run exports OMP_NUM_THREADS=1 to every rank.  The CPU legs (reference arm, cpu_baseline) must use the box's
# cores, so they pass thread counts EXPLICITLY (omp_set_num_threads through the oracle's n_threads argument,
# torch.set_num_threads, numba.set_num_threads) and never read the environment; the GPU arm keeps the 1-thread default (with
# N ranks x all-cores OpenMP teams spinning after every small host copy, the ranks' launch threads were starved: measured
# -30 % end-to-end QPS at N = 2).
HOST_CORES = host_cores()
os.environ.setdefault("NUMBA_NUM_THREADS", str(HOST_CORES))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from scaling_retriever_b200 import synth  # noqa: E402

K_TOP = 1000
METRIC = "top1000_qps"
UNIT = "queries/s"


def env_int(name, default):
    return int(os.environ.get(name, default))


# rank 0's stdout carries exactly one JSON line: keep NCCL's "NCCL version ..." banner (NCCL_DEBUG=VERSION) off it
if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": float(p["hbm_gbs"]), "bf16_tflops": float(p["bf16_tflops"]),
                "bf16_tflops_sustained": float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def ncu_traffic(kernel, launches_per_step, block_docs=None):
    """DRAM bytes per launch of `kernel` from the committed ncu launch list of THIS kernel shape (profiles/r02_traffic.json, made
    by tools/traffic_from_launches.py from profiles/r02_launches.csv), or None when the recorded shape / launch count per step
    no longer matches what just ran (then the number would be stale).  bench.py itself cannot measure DRAM traffic."""
    path = os.path.join(ROOT, "profiles", "r02_traffic.json")
    try:
        with open(path) as f:
            t = json.load(f)
        k = t["kernels"][kernel]
        if abs(k["launches_per_step"] - launches_per_step) > 1e-9 or (block_docs is not None and t["sparse_block_docs"] != block_docs):
            return None, None
        return k["dram_bytes_per_launch"], f"profiles/r02_traffic.json <- {t['source']} ({k['launches']} launches under ncu)"
    except (OSError, KeyError, ValueError):
        return None, None


class ClockSampler:
    """SM clocks / throttle reasons sampled DURING the timed region (the fields of the B200_PROFILING.md nvidia-smi recipe).
    Sampled in-process through NVML (nvidia-ml-py) from a background thread: a polling `nvidia-smi -lms` subprocess was measured
    to stretch the timed steps by 10-20 % (it re-attaches to the driver on every poll); NVML queries on an open handle do not.
    Falls back to the nvidia-smi subprocess when the NVML binding is missing."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    REASON_BITS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, gpu_index, period_s=0.02):
        self.proc = None
        self.thread = None
        self.samples = []          # (sm_mhz, reasons bitmask)
        self.sm_max = None
        try:
            import threading
            import pynvml
            pynvml.nvmlInit()
            uuid = str(torch.cuda.get_device_properties(gpu_index).uuid)
            try:
                handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
            except Exception:
                handle = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(handle, pynvml.NVML_CLOCK_SM))
            get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
            self._stop = threading.Event()

            def loop():
                while not self._stop.is_set():
                    try:
                        self.samples.append((float(pynvml.nvmlDeviceGetClockInfo(handle, pynvml.NVML_CLOCK_SM)), int(get_reasons(handle))))
                    except Exception:
                        pass
                    self._stop.wait(period_s)
            self.thread = threading.Thread(target=loop, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        self.path = tempfile.mktemp(prefix="clocks_", suffix=".csv")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        if self.thread is not None:
            self._stop.set()
            self.thread.join(timeout=2)
            sm = [s for s, _ in self.samples]
            if not sm:
                return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": [], "samples": 0, "via": "nvml"}
            bits = 0
            for _, r in self.samples:
                bits |= r
            return {"sm_mhz": statistics.median(sm), "sm_max_mhz": self.sm_max,
                    "reasons": sorted(n for n, b in self.REASON_BITS.items() if bits & b), "samples": len(sm), "via": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        with open(self.path) as f:
            for line in f:
                parts = [x.strip() for x in line.split(",")]
                if len(parts) < 8:
                    continue
                try:
                    sm.append(float(parts[1]))
                    mx.append(float(parts[2]))
                except ValueError:
                    continue
                for n, v in zip(names, parts[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
        os.unlink(self.path)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        busy = [c for c in sm if c > 0.5 * max(sm)] or sm
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------- workload descriptions
# The `config` dicts are built from the command line only, so both arms (`--impl b200` / `--impl reference`) of the same
# command print IDENTICAL dicts; everything measured goes elsewhere in the line.

def sparse_sizes(args):
    n_docs = args.n_docs or synth.MSMARCO_DOCS
    n_queries = args.n_queries or synth.MSMARCO_DEV_QUERIES
    return n_docs, n_queries, synth.LLAMA3_VOCAB


def parallelism(n_gpus):
    return (f"doc-range shards x{n_gpus}, per-round MIN all-reduce of the shards' bounds, packed-key all-to-all + per-slice merge + "
            "all-gather over NCCL; queries replicated" if n_gpus > 1 else "1 GPU")


L2_NOTE = "inputs larger than L2 (index / corpus of several GB vs 126 MB L2), no flush"


def sparse_config(args):
    n_docs, n_queries, n_terms = sparse_sizes(args)
    return {"workload": (f"sparse inverted-index top-{K_TOP}, synthetic {n_docs:,} docs x {n_terms:,} terms (~200 nnz/doc), "
                         f"{n_queries:,} queries (~40 nnz) [BASELINE.json configs[1]]"),
            "n_docs": n_docs, "n_queries": n_queries, "k": K_TOP, "n_terms": n_terms,
            "parallelism": parallelism(args.gpus), "l2": L2_NOTE}


def dense_config(args, dim):
    n_docs = args.n_docs or synth.MSMARCO_DOCS
    n_queries = args.n_queries or synth.MSMARCO_DEV_QUERIES
    cfg = "configs[2]" if dim == 2048 else ("configs[3]" if dim == 4096 else "configs[4] sweep point")
    return {"workload": (f"dense flat inner-product top-{K_TOP}, synthetic {n_docs:,} x {dim} bf16 corpus (L2-normalised "
                         f"Gaussian rows), {n_queries:,} queries [BASELINE.json {cfg}]"),
            "n_docs": n_docs, "n_queries": n_queries, "k": K_TOP, "dim": dim,
            "parallelism": parallelism(args.gpus), "l2": L2_NOTE}


def dense_dims(args):
    if args.workload == "sparse":
        return []
    if args.workload == "dense":
        return [args.dim]
    return [2048] + ([4096] if args.gpus >= 8 else [])


def result_digest(scores, ids):
    """sha256 over the merged rows (ids + fp32 score bits) — identical at every N iff sharded == unsharded."""
    h = hashlib.sha256()
    h.update(ids.cpu().numpy().tobytes())
    h.update(scores.cpu().numpy().view(np.uint32).tobytes())
    return h.hexdigest()


class Ctx:
    """Process-wide state of the GPU arm: rank / world / device, the process group, measured peaks."""

    def __init__(self, args):
        import torch.distributed as dist
        self.dist = dist
        self.rank, self.world, self.local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device (the product path has no CPU fallback)")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        assert self.world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={self.world} (launch N>1 with torchrun)"
        self.peaks = load_peaks()

    def barrier(self):
        torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, x):
        t = torch.tensor([float(x)], device=self.dev, dtype=torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def time_device(self, step, steps):
        """K steps between CUDA events, barrier + synchronize on both sides, max over ranks -> (ms per step, last output)."""
        self.barrier()
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        out = None
        for _ in range(steps):
            out = step()
        end.record()
        self.barrier()
        return self.max_over_ranks(start.elapsed_time(end)) / steps, out

    def time_wall(self, fn, steps, warmup=1):
        """Host wall clock per call of `fn` (the end-to-end legs), barrier on both sides, max over ranks."""
        out = None
        for _ in range(warmup):
            out = fn()
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            out = fn()
        self.barrier()
        return self.max_over_ranks((time.perf_counter() - t0) / steps), out

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------------------- sparse (GPU)

def bench_sparse(args, ctx):
    from scaling_retriever_b200 import ops, shard
    from scaling_retriever_b200.indexer import SparseRetrieval
    dev, rank, world = ctx.dev, ctx.rank, ctx.world
    n_docs, n_queries, n_terms = sparse_sizes(args)
    lo, hi = shard.ShardPlan(n_docs, world).bounds(rank)

    # ---- synthetic corpus shard -> CSR index + skip table in HBM (build timed for information) --------------------
    t0 = time.perf_counter()
    rows, cols, vals = synth.gen_sparse_docs(n_docs, n_terms=n_terms, device=dev, doc_lo=lo, doc_hi=hi)
    rows -= lo
    torch.cuda.synchronize()
    gen_s = time.perf_counter() - t0
    nnz = rows.numel()
    ops.profile_enable(True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    term_offsets, doc_ids, weights = ops.csr_build(rows, cols, vals, n_terms, hi - lo, sort_docs=False)
    ev1.record()
    torch.cuda.synchronize()
    build_ms = ev0.elapsed_time(ev1)
    sort_ms, _, _ = ops.profile_read(ops.PROF_CSR_SORT)
    del rows, cols, vals
    torch.cuda.empty_cache()
    t0 = time.perf_counter()
    # skip table + bank-ordered posting array ({doc id, fp32 weight}; --sparse-weights fp16: the opt-in compressed 4-byte format)
    index = ops.SparseDeviceIndex.from_csr(term_offsets, doc_ids, weights, hi - lo, weight_format=args.sparse_weights)
    torch.cuda.synchronize()
    table_s = time.perf_counter() - t0
    index.release_canonical()      # the search only streams index.postings; drop the 8 B/posting canonical copy
    del doc_ids, weights
    torch.cuda.empty_cache()

    q_off, q_terms, q_w = synth.gen_sparse_queries(n_queries, n_terms=n_terms, device=dev)
    # the step's inputs wait in PINNED host memory (the e2e contract): numpy views of pinned tensors
    h_pins = [t.cpu().pin_memory() for t in (q_off, q_terms, q_w)]
    h_off, h_terms, h_w = (t.numpy() for t in h_pins)
    # (4 + w) bytes per streamed posting, w = 4 (fp32 weights, the parity and default mode) or 2 (fp16)
    algo_bytes, postings = synth.sparse_algorithmic_bytes(term_offsets, q_terms, n_queries, K_TOP,
                                                          weight_bytes=4 if args.sparse_weights == "fp32" else 2)

    # N > 1: the shards exchange their bounds between the rounds (one 28 KB MIN all-reduce per round, shard.TauExchange)
    exchange = shard.TauExchange("sparse", n_docs, dev) if (world > 1 and args.sparse_weights == "fp32" and not args.no_tau_exchange) else None

    def step():
        s, i, c = ops.sparse_search(index, q_off, q_terms, q_w, K_TOP, 0.0, doc_id_base=lo, exchange=exchange)
        if world > 1:
            s, i, c = shard.merge_shards(s, i, K_TOP, n_docs_total=n_docs)
        return s, i, c

    # The clock sampler is started BEFORE the warm-up: attaching to the driver stalls the GPU for tens of ms, which must not
    # land in the timed region; the samples it takes during warm-up + timed steps are all under the same load.
    sampler = ClockSampler(ctx.local_rank) if rank == 0 else None
    for _ in range(args.warmup):
        out = step()   # held like in the timed loop: the allocator's second set of output blocks is created here, not there
    ops.profile_read(ops.PROF_SPARSE_SCORE)     # drop warm-up records and launch counts
    ops.profile_read(ops.PROF_SPARSE_SELECT)

    # ---- device-resident throughput (`value`) ----------------------------------------------------------------
    ms_per_step, out = ctx.time_device(step, args.steps)
    clocks = sampler.stop() if sampler else None
    score_ms, score_launches, all_launches = ops.profile_read(ops.PROF_SPARSE_SCORE)
    select_ms, _, _ = ops.profile_read(ops.PROF_SPARSE_SELECT)
    ops.profile_enable(False)
    value = n_queries / (ms_per_step / 1e3)

    # ---- end to end with HOST buffers (`e2e`): pinned h2d of the packed queries, search (+ merge), d2h of the rows --------
    retriever = SparseRetrieval.from_device_index(index, doc_ids=range(n_docs), doc_id_base=lo, size_collection=n_docs)
    # N > 1: every rank copies its queries in and searches its shard; the merged result is read back by rank 0 (the rank that
    # writes run.json in retrieve()), host_ranks="first"
    e2e_s, e_out = ctx.time_wall(lambda: retriever.search_arrays(h_off, h_terms, h_w, K_TOP, 0.0, host_ranks="first"),
                                 args.steps, warmup=2)
    h2d = h_off.nbytes + h_terms.nbytes + h_w.nbytes
    if rank == 0:
        d2h = sum(x.nbytes for x in e_out)
        assert np.array_equal(e_out[1], out[1].cpu().numpy())    # the host-buffer path returns the same rows

    # ---- the reference-facing method (`e2e_api`): list of per-query (col, values) arrays + qids -> run mapping + stats ----
    vecs = synth.queries_to_vecs(q_off, q_terms, q_w)
    qids = list(range(n_queries))
    api_steps = max(1, min(args.steps, 5))
    api_s, api_out = ctx.time_wall(lambda: retriever._sparse_retrieve_multithreaded(vecs, qids, threshold=0.0, topk=K_TOP),
                                   api_steps, warmup=1)
    run_json = None
    if rank == 0:
        res, _ = api_out
        t0 = time.perf_counter()
        eager = res.to_dict()                           # what a caller that walks every (query, doc) pair pays on top
        to_dict_s = time.perf_counter() - t0
        assert len(eager) == len(res)
        del eager
        assert len(res) == int((out[2] > 0).sum().item())
        first = next(iter(res))
        assert res[first][str(int(out[1][int(first), 0]))] == float(out[0][int(first), 0])
        with tempfile.TemporaryDirectory() as tmp:
            res.write_json(os.path.join(tmp, "warm.json"))                       # external-id table built once, like a retriever
            t0 = time.perf_counter()
            how = res.write_json(os.path.join(tmp, "run.json"))
            write_s = time.perf_counter() - t0
            run_json = {"writer": how, "seconds": write_s, "bytes": os.path.getsize(os.path.join(tmp, "run.json"))}

    digest = result_digest(out[0], out[1]) if rank == 0 else None
    line = None
    if rank == 0:
        per_launch_ms = score_ms / max(score_launches, 1)
        achieved = (algo_bytes * args.steps) / (score_ms / 1e3) / 1e9 if score_ms > 0 else 0.0
        peak = ctx.peaks["hbm_gbs"]
        build_gbs = nnz * 20 / (build_ms / 1e3) / 1e9
        traffic, traffic_src = ncu_traffic("sparse_score_kernel", score_launches / args.steps, ops.block_docs()) if world == 1 else (None, None)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": sparse_config(args),
            "e2e": {"value": n_queries / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "call": "SparseRetrieval.search_arrays(host CSR-packed queries) -> host rows (rank 0)"},
            "e2e_api": {"value": n_queries / api_s, "unit": UNIT, "steps": api_steps,
                        "call": "SparseRetrieval._sparse_retrieve_multithreaded(sparse_query_vecs, qids, threshold=0.0, topk=1000) "
                                "[reference indexer.py:405-474] -> (lazy run mapping, stats)",
                        "run_json": run_json,
                        "with_run_json_value": n_queries / (api_s + run_json["seconds"]),
                        "materialised_value": n_queries / (api_s + to_dict_s),
                        "note": "with_run_json_value adds retrieve()'s run.json write (native formatter) to every call; "
                                "materialised_value adds res.to_dict() — the eager dict of 7 M (docid, score) Python pairs the "
                                "reference builds in its insert loop — for callers that walk every pair"},
            "result_digest": digest,
            "gpu_launches": int(all_launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "sparse_score_kernel", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": ctx.peaks["source"], "launches": int(score_launches), "avg_launch_ms": per_launch_ms,
                         "algorithmic_bytes_per_step": algo_bytes, "score_kernel_share_of_step": (score_ms / args.steps) / ms_per_step,
                         "select_kernels_ms_per_step": select_ms / args.steps,
                         "note": "algorithmic bytes = 8 B per (query, term) posting as the reference streams them; the kernel "
                                 "re-serves postings shared between queries from L2 (DRAM traffic ~ the index once per step), so "
                                 "the limiter is the L1/shared data pipe (89 % under ncu), not DRAM; traffic = DRAM bytes per "
                                 "launch from the committed ncu launch list of this kernel shape (null when it does not match)"},
            "build": {"kernel": "csr_build (radix sort)", "csr_build_ms": build_ms, "radix_sort_ms": sort_ms, "skip_table_s": table_s,
                      "synth_gen_s": gen_s, "postings": nnz, "algorithmic_bytes": nnz * 20, "achieved": build_gbs, "unit": "GB/s",
                      "peak": peak, "frac": build_gbs / peak},
            "index_postings_this_rank": nnz, "postings_scored_per_query": postings / n_queries,
        }
        if args.sparse_weights != "fp32":
            line["config"]["sparse_weights"] = args.sparse_weights + " (opt-in compressed postings; NOT the parity mode)"
        if world == 1 and not args.no_cpu_baseline and args.sparse_weights == "fp32":
            # the oracle gets the same posting lists the GPU searched (slices in bank order; order inside a list is irrelevant)
            host = index.postings.cpu()
            h_ids, h_wts = host[:, 0].contiguous().numpy(), host[:, 1].contiguous().view(torch.float32).numpy()
            h_toff = term_offsets.cpu().numpy()
            del host
            line["cpu_baseline"] = sparse_cpu_baseline(h_toff, h_ids, h_wts, n_docs, h_off, h_terms, h_w, out)
            ref = reference_cpu_baseline(h_toff, h_ids, h_wts, n_docs, n_terms, h_off, h_terms, h_w, out)
            if ref is not None:
                line["cpu_baseline_reference"] = ref
    del index, retriever, out
    torch.cuda.empty_cache()
    return line


def sparse_cpu_baseline(off, ids, w, n_docs, h_off, h_terms, h_w, gpu_out, target_s=12.0):
    """The oracle port (C + OpenMP restatement of numba_score_float + select_topk) on this box's host cores, on a bounded
    query sample of the same index; also cross-checks the GPU rows of the sampled queries (ids + scores identical)."""
    from oracle import c_oracle
    threads = HOST_CORES
    probe = min(len(h_off) - 1, threads)
    t0 = time.perf_counter()
    c_oracle.sparse_search(off, ids, w, n_docs, h_off[:probe + 1], h_terms, h_w, K_TOP, n_threads=threads)
    per_round = max(time.perf_counter() - t0, 1e-3)
    sample = int(min(len(h_off) - 1, max(probe, probe * round(target_s / per_round))))
    t0 = time.perf_counter()
    o_scores, o_ids, o_counts = c_oracle.sparse_search(off, ids, w, n_docs, h_off[:sample + 1], h_terms, h_w, K_TOP, n_threads=threads)
    dt = time.perf_counter() - t0
    match = bool(np.array_equal(o_ids, gpu_out[1][:sample].cpu().numpy()) and
                 np.array_equal(o_scores.view(np.uint32), gpu_out[0][:sample].cpu().numpy().view(np.uint32)))
    return {"value": sample / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"first {sample} of the {len(h_off) - 1} queries over the full index, {dt:.1f} s, "
                      f"oracle/sparse_oracle.c with {threads} OpenMP threads", "gpu_rows_identical": match}


def reference_cpu_baseline(off, ids, w, n_docs, n_terms, h_off, h_terms, h_w, gpu_out, target_s=12.0, max_queries=200):
    """The REFERENCE'S OWN code (SparseRetrieval._sparse_retrieve_multithreaded: 4 Python threads x numba prange,
    indexer.py:405-474), unmodified, from the run-time copy under baseline/_ref/ (made by __graft_entry__.build() in the build
    container; git-ignored), on a bounded query sample of the same index.  None when the copy or numba is absent."""
    try:
        from oracle import ref_runner
        runner = ref_runner.load(os.path.join(ROOT, "baseline", "_ref"), threads=HOST_CORES)
    except Exception as exc:   # copy absent / numba missing: the port above stands alone
        return {"unavailable": f"{type(exc).__name__}: {exc}"[:200]}
    try:
        retr = runner.make_retriever(off, ids, w, n_docs, n_terms)
        vecs = [(h_terms[h_off[i]:h_off[i + 1]], h_w[h_off[i]:h_off[i + 1]]) for i in range(min(len(h_off) - 1, max_queries))]
        t0 = time.perf_counter()
        runner.retrieve(retr, vecs[:8], list(range(8)), K_TOP)          # JIT + first touch
        warm_s = time.perf_counter() - t0
        t0 = time.perf_counter()
        runner.retrieve(retr, vecs[:8], list(range(8)), K_TOP)
        per8 = max(time.perf_counter() - t0, 1e-3)
        sample = int(min(len(vecs), max(8, 8 * round(target_s / per8))))
        t0 = time.perf_counter()
        res, _ = runner.retrieve(retr, vecs[:sample], list(range(sample)), K_TOP)
        dt = time.perf_counter() - t0
        # cross-check: the reference's own result for a sampled query == the GPU row (same doc set, same scores)
        g_ids, g_scores = gpu_out[1][:sample].cpu().numpy(), gpu_out[0][:sample].cpu().numpy()
        same = True
        for qi in (0, sample - 1):
            ours = {str(int(d)): float(s) for d, s in zip(g_ids[qi], g_scores[qi]) if d >= 0}
            theirs = res.get(str(qi), {})
            kth = min(ours.values()) if ours else 0.0
            same = same and {d for d, s in ours.items() if s > kth} == {d for d, s in theirs.items() if s > kth} \
                and all(theirs[d] == s for d, s in ours.items() if d in theirs)
        return {"value": sample / dt, "unit": UNIT, "cores": HOST_CORES, "kind": "reference",
                "sample": f"first {sample} of the {len(h_off) - 1} queries over the full index in {dt:.1f} s through the reference's "
                          f"own SparseRetrieval._sparse_retrieve_multithreaded (4 threads x numba prange, "
                          f"{runner.numba_threads()} numba threads; index dicts are views of the same CSR; warm-up {warm_s:.1f} s)",
                "gpu_rows_agree": bool(same)}
    except Exception as exc:
        return {"unavailable": f"{type(exc).__name__}: {exc}"[:200]}


# ------------------------------------------------------------------------------------------------------------ dense (GPU)

def bench_dense(args, ctx, dim):
    """Dense workload (BASELINE.json configs[2]/[3]): bf16 tcgen05 GEMM + fused top-k over a doc-range shard per GPU."""
    from scaling_retriever_b200 import ops, shard
    from scaling_retriever_b200.indexer import DenseFlatIndexer
    dev, rank, world = ctx.dev, ctx.rank, ctx.world
    n_docs = args.n_docs or synth.MSMARCO_DOCS
    n_queries = args.n_queries or synth.MSMARCO_DEV_QUERIES
    lo, hi = shard.ShardPlan(n_docs, world).bounds(rank)
    corpus = synth.gen_dense(n_docs, dim, seed=1234, device=dev, dtype=torch.bfloat16, row_lo=lo, row_hi=hi)
    q32 = synth.gen_dense(n_queries, dim, seed=4321, device=dev)
    q16 = ops.f32_to_bf16(q32)
    h_q_pin = q32.cpu().pin_memory()         # the step's input waits in PINNED host memory (the e2e contract)
    h_q = h_q_pin.numpy()
    flops = 2.0 * n_queries * (hi - lo) * dim

    exchange = shard.TauExchange("dense", n_docs, dev) if (world > 1 and not args.no_tau_exchange) else None

    def step():
        s, i, c = ops.dense_search(corpus, q16, K_TOP, doc_id_base=lo, exchange=exchange)
        if world > 1:
            s, i, c = shard.merge_shards(s, i, K_TOP, n_docs_total=n_docs)
        return s, i, c

    ops.profile_enable(True)
    sampler = ClockSampler(ctx.local_rank) if rank == 0 else None
    for _ in range(args.warmup):
        out = step()   # held like in the timed loop (see bench_sparse)
    ops.profile_read(ops.PROF_DENSE_GEMM)
    ops.profile_read(ops.PROF_SPARSE_SELECT)
    ms_per_step, out = ctx.time_device(step, args.steps)
    clocks = sampler.stop() if sampler else None
    gemm_ms, gemm_launches, all_launches = ops.profile_read(ops.PROF_DENSE_GEMM)
    select_ms, _, _ = ops.profile_read(ops.PROF_SPARSE_SELECT)
    ops.profile_enable(False)
    value = n_queries / (ms_per_step / 1e3)

    # end to end with HOST buffers: host fp32 queries -> pinned -> device cast -> search (-> merge) -> host rows (rank 0)
    index = DenseFlatIndexer(device=dev)
    index.init_index(dim)
    index.index, index._row_lo, index._n_total = corpus, lo, n_docs
    index.index_id_to_db_id = range(n_docs)
    e2e_s, e_out = ctx.time_wall(lambda: index.search_arrays(h_q, K_TOP, host_ranks="first"), args.steps, warmup=2)
    if rank == 0:
        assert np.array_equal(e_out[1], out[1].cpu().numpy())
        d2h = int(e_out[0].nbytes + e_out[1].nbytes)

    # the reference-facing method: search_knn(query_reps fp32 [Q, d], top_docs) -> (rows of db ids, fp32 [Q, k]) on the first worker
    api_steps = max(1, min(args.steps, 3))
    api_s, api_out = ctx.time_wall(lambda: index.search_knn(h_q, K_TOP), api_steps, warmup=1)
    if rank == 0:
        assert api_out[0][0][0] == int(out[1][0, 0]) and np.array_equal(api_out[1], out[0].cpu().numpy())
        t0 = time.perf_counter()
        rows = api_out[0].tolist()                      # every row as Python lists (what iterating all rows costs in total)
        rows_s = time.perf_counter() - t0
        assert len(rows) == n_queries and len(rows[0]) == K_TOP
        del rows
    digest = result_digest(out[0], out[1]) if rank == 0 else None

    line = None
    if rank == 0:
        achieved = flops * args.steps / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
        peak = ctx.peaks["bf16_tflops"]
        traffic, traffic_src = ncu_traffic("dense_search_kernel", gemm_launches / args.steps) if (world == 1 and dim == 2048) else (None, None)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic", "config": dense_config(args, dim),
            "e2e": {"value": n_queries / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h_q.nbytes), "d2h_bytes_per_step": d2h,
                    "call": "DenseFlatIndexer.search_arrays(host fp32 queries) -> host rows (rank 0)"},
            "e2e_api": {"value": n_queries / api_s, "unit": UNIT, "steps": api_steps,
                        "call": "DenseFlatIndexer.search_knn(query_reps, 1000) [reference indexer.py:210-214] -> (rows of db ids "
                                "gathered on access, scores)",
                        "materialised_value": n_queries / (api_s + rows_s),
                        "note": "materialised_value adds turning all 6,980 x 1000 labels into Python lists of ids (CPython object "
                                "creation), which a caller that iterates every row pays in total"},
            "result_digest": digest,
            "gpu_launches": int(all_launches),
            "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": "dense_search_kernel", "achieved": achieved, "peak": peak,
                         "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": ctx.peaks["source"], "frac_of_sustained_peak": achieved / ctx.peaks["bf16_tflops_sustained"], "launches": int(gemm_launches),
                         "avg_launch_ms": gemm_ms / max(gemm_launches, 1), "algorithmic_flops_per_step": flops,
                         "gemm_kernel_share_of_step": (gemm_ms / args.steps) / ms_per_step,
                         "select_kernels_ms_per_step": select_ms / args.steps,
                         "note": "peak = measured cuBLAS bf16 burst; the timed region runs the tensor cores for seconds, "
                                 "so frac_of_sustained_peak (measured back-to-back cuBLAS) is the like-for-like figure"},
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = dense_cpu_baseline(corpus, h_q, n_docs)
    del corpus, index, out
    torch.cuda.empty_cache()
    return line


def dense_cpu_baseline(corpus, h_q, n_docs, sample_docs=200_000, sample_queries=512):
    """fp32 restatement of faiss.IndexFlatIP.search (oracle/dense_oracle.py, torch/MKL sgemm + top-k; faiss-cpu itself is not
    installable here) on a bounded slice of the same corpus; QPS is scaled linearly in N to the full corpus."""
    from oracle import dense_oracle
    torch.set_num_threads(HOST_CORES)
    nd, nq = min(sample_docs, corpus.shape[0]), min(sample_queries, len(h_q))
    docs = corpus[:nd].float().cpu().numpy()
    qs = torch.from_numpy(h_q[:nq]).to(torch.bfloat16).float().numpy()
    dense_oracle.flat_ip_search(docs[:20000], qs, K_TOP)     # MKL warm-up
    t0 = time.perf_counter()
    dense_oracle.flat_ip_search(docs, qs, K_TOP)
    dt = time.perf_counter() - t0
    return {"value": nq / (dt * n_docs / nd), "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{nq} queries x first {nd} docs in {dt:.1f} s (fp32 restatement of IndexFlatIP, not faiss), scaled "
                      f"linearly in N to {n_docs} docs"}


def run_b200(args):
    ctx = Ctx(args)
    records = []
    if args.workload in ("both", "sparse"):
        records.append(("sparse", bench_sparse(args, ctx)))
    for dim in dense_dims(args):
        records.append(("dense" if dim == 2048 or args.workload == "dense" else f"dense_{dim}", bench_dense(args, ctx, dim)))
    if ctx.rank == 0:
        line = records[0][1]
        for name, rec in records[1:]:
            line[name] = rec
        print(json.dumps(line), flush=True)
    ctx.close()


# ------------------------------------------------------------------------------------------------------- reference arm (CPU)

def reference_dense(args, dim):
    from oracle import dense_oracle
    torch.set_num_threads(HOST_CORES)
    n_docs = args.n_docs or synth.MSMARCO_DOCS
    n_queries = args.n_queries or synth.MSMARCO_DEV_QUERIES
    nd, nq = min(200_000, n_docs), min(512, n_queries)
    docs = synth.gen_dense(n_docs, dim, seed=1234, row_lo=0, row_hi=nd).numpy()
    qs = synth.gen_dense(n_queries, dim, seed=4321, row_lo=0, row_hi=nq).numpy()
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        dense_oracle.flat_ip_search(docs, qs, K_TOP)
        if i >= args.warmup:
            times.append(time.perf_counter() - t0)
    s_per_step = sum(times) / len(times)
    value = nq / (s_per_step * n_docs / nd)
    sample = (f"{nq} queries x first {nd} docs per step (fp32 restatement of faiss IndexFlatIP: torch/MKL sgemm + top-k, "
              f"{torch.get_num_threads()} threads), QPS scaled linearly in N to {n_docs} docs")
    return {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": s_per_step * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": dense_config(args, dim),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def reference_sparse(args):
    """The CPU port of the reference's sparse retrieval path, all host cores, bounded sample per step."""
    from oracle import c_oracle
    n_docs, n_queries, n_terms = sparse_sizes(args)
    dev = "cuda" if torch.cuda.is_available() else "cpu"     # torch RNG only generates the synthetic corpus
    rows, cols, vals = synth.gen_sparse_docs(n_docs, n_terms=n_terms, device=dev)
    rows, cols, vals = rows.cpu().numpy(), cols.cpu().numpy(), vals.cpu().numpy()
    off, ids, w = c_oracle.build_csr(rows, cols, vals, n_terms)
    del rows, cols, vals
    q_off, q_terms, q_w = (x.cpu().numpy() for x in synth.gen_sparse_queries(n_queries, n_terms=n_terms, device=dev))
    threads = HOST_CORES          # explicit: torchrun exports OMP_NUM_THREADS=1
    sample = min(n_queries, max(threads * 4, 16))
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        c_oracle.sparse_search(off, ids, w, n_docs, q_off[:sample + 1], q_terms, q_w, K_TOP, n_threads=threads)
        if i >= args.warmup:
            times.append(time.perf_counter() - t0)
    s_per_step = sum(times) / len(times)
    value = sample / s_per_step
    return {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": s_per_step * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": sparse_config(args),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{sample} queries per step over the full index; oracle/sparse_oracle.c (C + OpenMP "
                                       f"restatement of numba_score_float + select_topk), {threads} threads"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def run_reference(args):
    if env_int("RANK", 0) != 0:
        return               # under torchrun rank 0 alone runs the CPU arm; the other ranks exit 0 without work
    records = []
    if args.workload in ("both", "sparse"):
        records.append(("sparse", reference_sparse(args)))
    for dim in dense_dims(args):
        records.append(("dense" if dim == 2048 or args.workload == "dense" else f"dense_{dim}", reference_dense(args, dim)))
    line = records[0][1]
    for name, rec in records[1:]:
        line[name] = rec
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n-docs", type=int, default=0, help="override the corpus size (debug; the headline is 8,841,823)")
    ap.add_argument("--n-queries", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-tau-exchange", action="store_true", help="(N > 1, A/B only) shards do not exchange their bounds between rounds")
    ap.add_argument("--workload", default="both", choices=["both", "sparse", "dense"],
                    help="both (default) = sparse configs[1] at the top level + dense configs[2] (and configs[3] at N >= 8) as "
                         "sub-records; sparse / dense = that workload alone")
    ap.add_argument("--sparse-weights", default="fp32", choices=["fp32", "fp16"],
                    help="posting weights: fp32 = parity mode (default, the headline); fp16 = opt-in compressed 4-byte postings")
    ap.add_argument("--dim", type=int, default=2048, help="dense row width for --workload dense (2048 = Lion-DS-1B, 4096 = Lion-DS-8B)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()

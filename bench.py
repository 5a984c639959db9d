#!/usr/bin/env python
"""Headline benchmark: top-1000 QPS of the retrieval hot path on synthetic MS-MARCO-shaped data (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload sparse|dense]

A "step" is one pass of the hot path over the whole query batch (6,980 queries, top-1000) against the index resident
in HBM.  Default workload = BASELINE.json configs[1]: sparse inverted-index retrieval over 8,841,823 synthetic docs x
128,256 terms; for N > 1 (torchrun, one rank per GPU) the corpus is sharded by doc-id range, every rank searches its
shard and the per-shard top-k rows are merged after an NCCL all-gather ("strong" scaling: total work fixed).
Rank 0 prints ONE JSON line.  Timing: CUDA events on the launching stream, barrier + synchronize on both sides, max
over ranks; the index (14 GB) is far larger than L2 (126 MB), so no explicit L2 flush is needed between steps.
`--impl reference` times the CPU port of the reference's numba path (oracle/sparse_oracle.c, all host threads) on a
bounded sample of the same workload; only that leg and the cpu_baseline leg execute anything under oracle/.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

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


# ----------------------------------------------------------------------------------------------------- workloads

def sparse_sizes(args):
    n_docs = args.n_docs or synth.MSMARCO_DOCS
    n_queries = args.n_queries or synth.MSMARCO_DEV_QUERIES
    return n_docs, n_queries, synth.LLAMA3_VOCAB


def workload_name(args):
    n_docs, n_queries, n_terms = sparse_sizes(args)
    return (f"sparse inverted-index top-{K_TOP}, synthetic {n_docs:,} docs x {n_terms:,} terms (~200 nnz/doc), "
            f"{n_queries:,} queries (~40 nnz) [BASELINE.json configs[1]]")


def run_b200(args):
    import torch.distributed as dist
    from scaling_retriever_b200 import ops, shard
    from scaling_retriever_b200.indexer import SparseRetrieval

    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch N>1 with torchrun)"
    peaks = load_peaks()

    n_docs, n_queries, n_terms = sparse_sizes(args)
    plan = shard.ShardPlan(n_docs, world)
    lo, hi = plan.bounds(rank)

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
    index = ops.SparseDeviceIndex.from_csr(term_offsets, doc_ids, weights, hi - lo)   # skip table + bank-ordered posting array
    torch.cuda.synchronize()
    table_s = time.perf_counter() - t0
    index.release_canonical()      # the search only streams index.postings; drop the 8 B/posting canonical copy
    del doc_ids, weights
    torch.cuda.empty_cache()

    q_off, q_terms, q_w = synth.gen_sparse_queries(n_queries, n_terms=n_terms, device=dev)
    h_off, h_terms, h_w = q_off.cpu().numpy(), q_terms.cpu().numpy(), q_w.cpu().numpy()
    algo_bytes, postings = synth.sparse_algorithmic_bytes(term_offsets, q_terms, n_queries, K_TOP)

    def step():
        s, i, c = ops.sparse_search(index, q_off, q_terms, q_w, K_TOP, 0.0, doc_id_base=lo)
        if world > 1:
            s, i, c = shard.merge_shards(s, i, K_TOP)
        return s, i, c

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # nvidia-smi is started BEFORE the warm-up: attaching to the driver stalls the GPU for tens of ms, which must not land
    # in the timed region; the samples it takes during warm-up + timed steps are all under the same load.
    sampler = ClockSampler(local_rank) if rank == 0 else None
    for _ in range(args.warmup):
        out = step()   # held like in the timed loop: the allocator's second set of output blocks is created here, not there
    ops.profile_read(ops.PROF_SPARSE_SCORE)     # drop warm-up records and launch counts
    ops.profile_read(ops.PROF_SPARSE_SELECT)

    # ---- device-resident throughput (`value`) ----------------------------------------------------------------
    barrier()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    marks = []
    start.record()
    for _ in range(args.steps):
        out = step()
        if os.environ.get("B200RET_BENCH_DEBUG"):
            marks.append(torch.cuda.Event(enable_timing=True))
            marks[-1].record()
    end.record()
    barrier()
    if marks:
        print("per-step ms:", [round(a.elapsed_time(b), 2) for a, b in zip([start] + marks[:-1], marks)], file=sys.stderr)
    clocks = sampler.stop() if sampler else None
    ms_total = torch.tensor([start.elapsed_time(end)], device=dev)
    score_ms, score_launches, all_launches = ops.profile_read(ops.PROF_SPARSE_SCORE)
    select_ms, _, _ = ops.profile_read(ops.PROF_SPARSE_SELECT)
    ops.profile_enable(False)
    stats = torch.tensor([score_ms, float(algo_bytes), float(postings)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms_total, op=dist.ReduceOp.MAX)
        gathered = [torch.zeros_like(stats) for _ in range(world)]
        dist.all_gather(gathered, stats)
    else:
        gathered = [stats]
    ms_per_step = float(ms_total.item()) / args.steps
    value = n_queries / (ms_per_step / 1e3)

    # ---- end to end through the class API with HOST buffers (`e2e`) --------------------------------------------
    retriever = SparseRetrieval.from_device_index(index, doc_id_base=lo, size_collection=n_docs)
    # N > 1: every rank copies its queries in and searches its shard; the merged result is read back by rank 0 (the rank that
    # writes run.json in retrieve()), host_ranks="first"
    for _ in range(2):
        retriever.search_arrays(h_off, h_terms, h_w, K_TOP, 0.0, host_ranks="first")
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e_scores, e_ids, e_counts = retriever.search_arrays(h_off, h_terms, h_w, K_TOP, 0.0, host_ranks="first")
    barrier()
    e2e_s = torch.tensor([(time.perf_counter() - t0) / args.steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    h2d = h_off.nbytes + h_terms.nbytes + h_w.nbytes
    if rank == 0:
        d2h = e_scores.nbytes + e_ids.nbytes + e_counts.nbytes
        assert np.array_equal(e_ids, out[1].cpu().numpy())    # the host-buffer path returns the same rows

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # roofline of the dominant kernel (sparse_score_kernel) on rank 0's shard: algorithmic bytes / CUDA-event time
    r_score_ms, r_bytes, _ = gathered[0].tolist()
    per_launch_ms = r_score_ms / max(score_launches, 1)
    achieved = (r_bytes * args.steps) / (r_score_ms / 1e3) / 1e9 if r_score_ms > 0 else 0.0
    traffic = None
    prof = os.path.join(ROOT, "profiles", "sparse_score_traffic.json")
    if os.path.exists(prof):
        with open(prof) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args), "n_docs": n_docs, "n_queries": n_queries, "k": K_TOP, "n_terms": n_terms,
                   "parallelism": (f"doc-range shards x{world} + NCCL all-gather merge; e2e: queries copied in on every rank, merged "
                                   "result read back by rank 0") if world > 1 else "1 GPU",
                   "l2": "inputs larger than L2 (index %.1f GB vs 126 MB L2), no flush" % (nnz * 8 / 1e9),
                   "index_postings_this_rank": nnz, "postings_scored_per_query": postings / n_queries},
        "e2e": {"value": n_queries / float(e2e_s.item()), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(all_launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "sparse_score_kernel", "achieved": achieved, "peak": peaks["hbm_gbs"],
                     "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"], "traffic": traffic, "peak_source": peaks["source"],
                     "launches": int(score_launches), "avg_launch_ms": per_launch_ms,
                     "algorithmic_bytes_per_step": r_bytes, "score_kernel_share_of_step": (r_score_ms / args.steps) / ms_per_step,
                     "select_kernels_ms_per_step": select_ms / args.steps,
                     "note": "algorithmic bytes = 8 B per (query, term) posting as the reference streams them; the kernel "
                             "re-serves postings shared between queries from L2, so achieved may exceed DRAM traffic"},
        "build": {"csr_build_ms": build_ms, "radix_sort_ms": sort_ms, "skip_table_s": table_s, "synth_gen_s": gen_s,
                  "postings": nnz, "algorithmic_gbs": nnz * 20 / (build_ms / 1e3) / 1e9},
    }
    if world == 1 and not args.no_cpu_baseline:
        # the oracle gets the same posting lists the GPU searched (slices in bank order; order inside a list is irrelevant)
        host = index.postings.cpu()
        line["cpu_baseline"] = cpu_baseline(term_offsets, host[:, 0].contiguous(), host[:, 1].contiguous().view(torch.float32),
                                            n_docs, h_off, h_terms, h_w, out)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(term_offsets, doc_ids, weights, n_docs, h_off, h_terms, h_w, gpu_out, target_s=15.0):
    """The oracle port (C + OpenMP restatement of numba_score_float + select_topk) on this box's host cores, on a bounded
    query sample of the same index; also cross-checks the GPU rows of the sampled queries (ids + scores identical)."""
    from oracle import c_oracle
    off, ids, w = term_offsets.cpu().numpy(), doc_ids.cpu().numpy(), weights.cpu().numpy()
    threads = c_oracle.max_threads()
    probe = min(len(h_off) - 1, threads)
    t0 = time.perf_counter()
    c_oracle.sparse_search(off, ids, w, n_docs, h_off[:probe + 1], h_terms, h_w, K_TOP)
    per_round = max(time.perf_counter() - t0, 1e-3)
    sample = int(min(len(h_off) - 1, max(probe, probe * round(target_s / per_round))))
    t0 = time.perf_counter()
    o_scores, o_ids, o_counts = c_oracle.sparse_search(off, ids, w, n_docs, h_off[:sample + 1], h_terms, h_w, K_TOP)
    dt = time.perf_counter() - t0
    match = bool(np.array_equal(o_ids, gpu_out[1][:sample].cpu().numpy()) and
                 np.array_equal(o_scores.view(np.uint32), gpu_out[0][:sample].cpu().numpy().view(np.uint32)))
    return {"value": sample / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"first {sample} of the {len(h_off) - 1} queries over the full index, {dt:.1f} s, "
                      f"oracle/sparse_oracle.c with {threads} OpenMP threads", "gpu_rows_identical": match}


def dense_workload_name(n_docs, n_queries, dim):
    cfg = "configs[2]" if dim == 2048 else ("configs[3]" if dim == 4096 else "configs[4] sweep point")
    return (f"dense flat inner-product top-{K_TOP}, synthetic {n_docs:,} x {dim} bf16 corpus (L2-normalised Gaussian rows), "
            f"{n_queries:,} queries [BASELINE.json {cfg}]")


def run_b200_dense(args):
    """Dense workload (BASELINE.json configs[2]/[3]): bf16 tcgen05 GEMM + fused top-k over a doc-range shard per GPU."""
    import torch.distributed as dist
    from scaling_retriever_b200 import ops, shard
    from scaling_retriever_b200.indexer import DenseFlatIndexer

    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch N>1 with torchrun)"
    peaks = load_peaks()
    n_docs = args.n_docs or synth.MSMARCO_DOCS
    n_queries = args.n_queries or synth.MSMARCO_DEV_QUERIES
    dim = args.dim
    lo, hi = shard.ShardPlan(n_docs, world).bounds(rank)
    corpus = synth.gen_dense(n_docs, dim, seed=1234, device=dev, dtype=torch.bfloat16, row_lo=lo, row_hi=hi)
    q32 = synth.gen_dense(n_queries, dim, seed=4321, device=dev)
    q16 = ops.f32_to_bf16(q32)
    h_q = q32.cpu().numpy()
    flops = 2.0 * n_queries * (hi - lo) * dim

    def step():
        s, i, c = ops.dense_search(corpus, q16, K_TOP, doc_id_base=lo)
        if world > 1:
            s, i, c = shard.merge_shards(s, i, K_TOP)
        return s, i, c

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ops.profile_enable(True)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    for _ in range(args.warmup):
        out = step()   # held like in the timed loop (see run_b200)
    ops.profile_read(ops.PROF_DENSE_GEMM)
    ops.profile_read(ops.PROF_SPARSE_SELECT)
    barrier()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(args.steps):
        out = step()
    end.record()
    barrier()
    clocks = sampler.stop() if sampler else None
    ms_total = torch.tensor([start.elapsed_time(end)], device=dev)
    gemm_ms, gemm_launches, all_launches = ops.profile_read(ops.PROF_DENSE_GEMM)
    select_ms, _, _ = ops.profile_read(ops.PROF_SPARSE_SELECT)
    ops.profile_enable(False)
    if world > 1:
        dist.all_reduce(ms_total, op=dist.ReduceOp.MAX)
    ms_per_step = float(ms_total.item()) / args.steps
    value = n_queries / (ms_per_step / 1e3)

    # end to end through the class API: host fp32 queries -> pinned -> device cast -> search (-> merge) -> host rows
    index = DenseFlatIndexer(device=dev)
    index.init_index(dim)
    index.index = corpus
    index._row_lo = lo
    for _ in range(2):
        index.search_arrays(h_q, K_TOP, host_ranks="first")
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e_scores, e_ids = index.search_arrays(h_q, K_TOP, host_ranks="first")   # N > 1: result read back by rank 0
    barrier()
    e2e_s = torch.tensor([(time.perf_counter() - t0) / args.steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    if rank == 0:
        assert np.array_equal(e_ids, out[1].cpu().numpy())
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    achieved = flops * args.steps / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": dense_workload_name(n_docs, n_queries, dim), "n_docs": n_docs, "n_queries": n_queries, "k": K_TOP,
                   "dim": dim, "parallelism": (f"doc-range shards x{world} + NCCL all-gather merge; e2e: queries copied in on every rank, "
                                               "merged result read back by rank 0") if world > 1 else "1 GPU",
                   "l2": "inputs larger than L2 (corpus shard %.1f GB vs 126 MB L2), no flush" % ((hi - lo) * dim * 2 / 1e9)},
        "e2e": {"value": n_queries / float(e2e_s.item()), "unit": UNIT, "h2d_bytes_per_step": int(h_q.nbytes),
                "d2h_bytes_per_step": int(e_scores.nbytes + e_ids.nbytes)},
        "gpu_launches": int(all_launches),
        "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "dense_search_kernel", "achieved": achieved, "peak": peaks["bf16_tflops"],
                     "unit": "TFLOP/s", "frac": achieved / peaks["bf16_tflops"], "traffic": None, "peak_source": peaks["source"],
                     "frac_of_sustained_peak": achieved / peaks["bf16_tflops_sustained"], "launches": int(gemm_launches),
                     "avg_launch_ms": gemm_ms / max(gemm_launches, 1), "algorithmic_flops_per_step": flops,
                     "gemm_kernel_share_of_step": (gemm_ms / args.steps) / ms_per_step,
                     "select_kernels_ms_per_step": select_ms / args.steps},
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = dense_cpu_baseline(corpus, h_q, out, n_docs)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def dense_cpu_baseline(corpus, h_q, gpu_out, n_docs, sample_docs=200_000, sample_queries=512):
    """fp32 restatement of faiss.IndexFlatIP.search (oracle/dense_oracle.py, torch/MKL sgemm + top-k; faiss-cpu itself is not
    installable here) on a bounded slice of the same corpus; QPS is scaled linearly in N to the full corpus."""
    from oracle import dense_oracle
    nd, nq = min(sample_docs, corpus.shape[0]), min(sample_queries, len(h_q))
    docs = corpus[:nd].float().cpu().numpy()
    qs = torch.from_numpy(h_q[:nq]).to(torch.bfloat16).float().numpy()
    t0 = time.perf_counter()
    o_scores, o_ids = dense_oracle.flat_ip_search(docs, qs, K_TOP)
    dt = time.perf_counter() - t0
    return {"value": nq / (dt * n_docs / nd), "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{nq} queries x first {nd} docs in {dt:.1f} s (fp32 restatement of IndexFlatIP, not faiss), scaled "
                      f"linearly in N to {n_docs} docs"}


def run_reference_dense(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    from oracle import dense_oracle
    n_docs = args.n_docs or synth.MSMARCO_DOCS
    n_queries = args.n_queries or synth.MSMARCO_DEV_QUERIES
    nd, nq = min(200_000, n_docs), min(512, n_queries)
    docs = synth.gen_dense(n_docs, args.dim, seed=1234, row_lo=0, row_hi=nd).numpy()
    qs = synth.gen_dense(n_queries, args.dim, seed=4321, row_lo=0, row_hi=nq).numpy()
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        dense_oracle.flat_ip_search(docs, qs, K_TOP)
        if i >= args.warmup:
            times.append(time.perf_counter() - t0)
    s_per_step = sum(times) / len(times)
    value = nq / (s_per_step * n_docs / nd)
    sample = (f"{nq} queries x first {nd} docs per step (fp32 restatement of faiss IndexFlatIP: torch/MKL sgemm + top-k), QPS "
              f"scaled linearly in N to {n_docs} docs")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": s_per_step * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": dense_workload_name(n_docs, n_queries, args.dim), "n_docs": n_docs, "n_queries": n_queries,
                   "k": K_TOP, "dim": args.dim},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)


def run_reference(args):
    """Reference arm: the CPU port of the reference's sparse retrieval path, all host threads, bounded sample per step."""
    rank, world = env_int("RANK", 0), env_int("WORLD_SIZE", 1)
    if rank != 0:
        return
    from oracle import c_oracle
    n_docs, n_queries, n_terms = sparse_sizes(args)
    dev = "cuda" if torch.cuda.is_available() else "cpu"     # torch RNG only generates the synthetic corpus
    rows, cols, vals = synth.gen_sparse_docs(n_docs, n_terms=n_terms, device=dev)
    rows, cols, vals = rows.cpu().numpy(), cols.cpu().numpy(), vals.cpu().numpy()
    off, ids, w = c_oracle.build_csr(rows, cols, vals, n_terms)
    del rows, cols, vals
    q_off, q_terms, q_w = (x.cpu().numpy() for x in synth.gen_sparse_queries(n_queries, n_terms=n_terms, device=dev))
    threads = c_oracle.max_threads()
    sample = min(n_queries, max(threads * 4, 16))
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        c_oracle.sparse_search(off, ids, w, n_docs, q_off[:sample + 1], q_terms, q_w, K_TOP)
        if i >= args.warmup:
            times.append(time.perf_counter() - t0)
    s_per_step = sum(times) / len(times)
    value = sample / s_per_step
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": s_per_step * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args), "n_docs": n_docs, "n_queries": n_queries, "k": K_TOP, "n_terms": n_terms},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{sample} queries per step over the full index; oracle/sparse_oracle.c (C + OpenMP "
                                   f"restatement of numba_score_float + select_topk), {threads} threads"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
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
    ap.add_argument("--workload", default="sparse", choices=["sparse", "dense"],
                    help="sparse = BASELINE.json configs[1] (default, the headline); dense = configs[2]/[3]")
    ap.add_argument("--dim", type=int, default=2048, help="dense row width (2048 = Lion-DS-1B, 4096 = Lion-DS-8B)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        (run_reference_dense if args.workload == "dense" else run_reference)(args)
    else:
        (run_b200_dense if args.workload == "dense" else run_b200)(args)


if __name__ == "__main__":
    main()

"""dev helper: time / profile the CSR build on a synthetic shard (run under gpurun, optionally under ncu)."""
import sys, time, torch
sys.path.insert(0, ".")
from scaling_retriever_b200 import ops, synth
n_docs = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
dev = torch.device("cuda", 0)
rows, cols, vals = synth.gen_sparse_docs(n_docs, device=dev)
torch.cuda.synchronize()
for it in range(3):
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    off, ids, w = ops.csr_build(rows, cols, vals, synth.LLAMA3_VOCAB, n_docs)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    print(f"csr_build {rows.numel()/1e6:.1f} M postings: {ms:.2f} ms  {rows.numel()*20/ms/1e6:.1f} GB/s algorithmic")
t0 = time.perf_counter()
index = ops.SparseDeviceIndex.from_csr(off, ids, w, n_docs)
torch.cuda.synchronize()
print(f"table+layout {time.perf_counter()-t0:.3f} s")

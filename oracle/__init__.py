"""CPU oracle for the B200 retrieval engine — TEST INFRASTRUCTURE ONLY.

Restates the reference's algorithm (numpy in sparse_oracle.py / dense_oracle.py, C + OpenMP in sparse_oracle.c).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this package;
the product package (scaling_retriever_b200, scaling_retriever) never does.
"""

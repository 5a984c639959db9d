"""Drop-in for scaling_retriever/utils/inverted_index.py (IndexDictOfArray, merge_indexes, CLI)."""
from scaling_retriever_b200.inverted_index import IndexDictOfArray, merge_indexes  # noqa: F401

if __name__ == "__main__":
    import runpy
    runpy.run_module("scaling_retriever_b200.inverted_index", run_name="__main__")

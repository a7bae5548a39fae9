"""rag-project-icd10_b200 -- B200-native engines behind the retrieval hot path of
yilane/rag-project-icd10 (embedding of ICD-10 texts + exact cosine top-k with level re-rank).

The directory name carries a hyphen, so import it with
    pkg = importlib.import_module("rag-project-icd10_b200")
or through the repo-root alias module ``icd10_b200``.

Layout
    csrc/       CUDA kernels (sm_100a) + the C ABI (include/icdrag.h) -> csrc/libicdrag.so
    _native.py  ctypes binding of the C ABI (fails loudly when the library or a GPU is missing)
    engine/     host side above the ABI: vector store, encoder engine, tokeniser, sharding
    services/   drop-in mirrors of the reference's services/embedding_service.py,
                services/milvus_service.py, services/hierarchical_similarity_service.py
    tools/      drop-in mirror of the reference's tools/build_database.py
"""
__version__ = "0.1.0"

from . import _native  # noqa: F401  (does not load the library until first use)

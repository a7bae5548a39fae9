"""Drop-in mirrors of the reference's services on the retrieval hot path.

    embedding_service.EmbeddingService            <- /root/reference/services/embedding_service.py
    milvus_service.MilvusService                  <- /root/reference/services/milvus_service.py
    hierarchical_similarity_service.*             <- /root/reference/services/hierarchical_similarity_service.py
    uncertainty_diagnosis_service.*               <- /root/reference/services/uncertainty_diagnosis_service.py
Same class names, method names, argument meaning, return shapes and error behaviour; the
engines behind them are libicdrag.so kernels instead of sentence-transformers / Milvus Lite.
"""

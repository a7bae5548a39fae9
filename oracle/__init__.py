"""CPU oracle for the ICD-10 retrieval hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and there only as the checker or the timed CPU
baseline.  The product package (``rag-project-icd10_b200``) never imports it and raises
when its CUDA library is missing.

What is restated (all citations into /root/reference/):
  * text preparation and the CSV -> record rules      (oracle/text.py)
      services/embedding_service.py:68-73,117-120; tools/build_database.py:62-192
  * the encoder arithmetic of SentenceTransformer.encode (oracle/encoder.py)
      third party: sentence-transformers>=4.1.0 (requirements.txt:7, not installable here)
      -> HF BertModel fp32 + masked mean + L2 normalise, call sites embedding_service.py:81,97,120
  * Milvus FLAT/IP search + the post-top-k level re-rank (oracle/search.py)
      third party: pymilvus==2.5.10 (requirements.txt:35, not installable here)
      -> exact fp32 inner product, call site services/milvus_service.py:271-320,550-558

PARITY PINNING.  The reference ships no tests, golden vectors or fixtures for this path
(SURVEY.md section 4), and its two engines (sentence-transformers, pymilvus) cannot be
installed offline, so the encoder and search arithmetic are "parity unpinned": they follow
the published algorithm of the third-party engines and the reference's call sites.  What
*is* pinned against the reference run in the build container (tests/golden/make_golden.py):
the CSV->record rules (tools/build_database.py imported with stubbed service modules) and
the scoring services (services/hierarchical_similarity_service.py and
services/uncertainty_diagnosis_service.py imported unmodified).
"""

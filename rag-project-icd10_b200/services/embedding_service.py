"""EmbeddingService -- drop-in mirror of /root/reference/services/embedding_service.py.

Same public surface (encode_single / encode_batch / encode_query / encode_icd_record /
get_model_info / test_embedding, config and device attributes, error behaviour); the engine
behind ``self.model`` is engine.encoder.EncoderEngine (libicdrag.so: tcgen05 GEMMs, fused
epilogues, attention, pooling) instead of sentence_transformers.SentenceTransformer.
Text preparation is kept verbatim in behaviour: "passage: " unless the text already starts
with query:/passage: (reference :68-73), unconditional "query: " in encode_query (:117-120).
"""
from __future__ import annotations

import os
from typing import Any, Dict, List

import numpy as np

try:
    from loguru import logger
except Exception:  # pragma: no cover
    import logging
    logger = logging.getLogger("icd10_b200")

try:  # the reference loads .env at import time (embedding_service.py:9-10)
    from dotenv import load_dotenv
    load_dotenv()
except Exception:  # pragma: no cover
    pass

from ..engine.encoder import EncoderEngine

_DEFAULT_MODEL = "intfloat/multilingual-e5-large-instruct"  # the reference's code default (:26)


class EmbeddingService:
    engine_factory = EncoderEngine  # tests may substitute a pre-built engine factory

    def __init__(self):
        self.config = self._load_config()
        self.model = None
        self.device = self._get_device()
        self._load_model()

    def _load_config(self) -> Dict[str, Any]:
        return {"embedding": {
            "model_name": os.getenv("EMBEDDING_MODEL_NAME", _DEFAULT_MODEL),
            "max_length": 512,
            "batch_size": 32,
            "device": os.getenv("EMBEDDING_DEVICE", "auto"),
        }}

    def _get_device(self) -> str:
        wanted = self.config.get("embedding", {}).get("device", "auto")
        if wanted != "auto":
            return wanted
        import torch
        if torch.cuda.is_available():
            return "cuda"
        if hasattr(torch.backends, "mps") and torch.backends.mps.is_available():
            return "mps"
        return "cpu"

    def _load_model(self):
        name = self.config.get("embedding", {}).get("model_name", _DEFAULT_MODEL)
        try:
            logger.info(f"正在加载嵌入模型: {name}")
            logger.info(f"目标设备: {self.device}")
            if not name:
                raise ValueError("模型名称不能为空")
            # raises on non-CUDA devices: this build has no CPU path
            self.model = type(self).engine_factory(name, device=self.device)
            logger.info(f"模型加载成功，使用设备: {self.device}")
        except Exception as e:
            logger.error(f"模型加载失败: {e}")
            logger.error(f"配置信息: {self.config}")
            raise

    def _prepare_text_for_embedding(self, text: str) -> str:
        if not text.startswith(("query:", "passage:")):
            text = f"passage: {text}"
        return text

    def encode_single(self, text: str) -> np.ndarray:
        if not self.model:
            raise RuntimeError("嵌入模型未加载")
        return self.model.encode(self._prepare_text_for_embedding(text), normalize_embeddings=True)

    def encode_batch(self, texts: List[str], show_progress: bool = True) -> List[np.ndarray]:
        if not self.model:
            raise RuntimeError("嵌入模型未加载")
        if not texts:
            return []
        prepared = [self._prepare_text_for_embedding(t) for t in texts]
        batch_size = self.config.get("embedding", {}).get("batch_size", 32)
        vectors = self.model.encode(prepared, batch_size=batch_size, show_progress_bar=show_progress,
                                    normalize_embeddings=True)
        return vectors.tolist()

    def encode_icd_record(self, icd_record: Dict[str, Any]) -> np.ndarray:
        title = icd_record.get("preferred_zh", "")
        if not title.strip():
            title = f"ICD代码 {icd_record.get('code', 'unknown')}"
        return self.encode_single(title)

    def encode_query(self, query: str) -> np.ndarray:
        # like the reference, no None-check here (embedding_service.py:117-120)
        return self.model.encode(f"query: {query}", normalize_embeddings=True)

    # extension used by tools/build_database.py and the batched serve path: many queries, one launch
    def encode_queries(self, queries: List[str]) -> np.ndarray:
        if not self.model:
            raise RuntimeError("嵌入模型未加载")
        return self.model.encode([f"query: {q}" for q in queries], normalize_embeddings=True)

    # extension for the sharded build: the embeddings stay on this rank's GPU (a float32 CUDA tensor)
    def encode_queries_device(self, queries: List[str]):
        if not self.model:
            raise RuntimeError("嵌入模型未加载")
        return self.model.encode([f"query: {q}" for q in queries], normalize_embeddings=True, convert_to_tensor=True)

    def get_model_info(self) -> Dict[str, Any]:
        if not self.model:
            return {"loaded": False}
        return {
            "loaded": True,
            "model_name": self.config.get("embedding", {}).get("model_name"),
            "device": self.device,
            "max_seq_length": getattr(self.model, "max_seq_length", None),
            "embedding_dimension": self.model.get_sentence_embedding_dimension(),
        }

    def test_embedding(self, test_text: str = "测试文本") -> Dict[str, Any]:
        try:
            vec = self.encode_single(test_text)
            return {"success": True, "embedding_shape": vec.shape, "embedding_type": str(type(vec)),
                    "sample_values": vec[:5].tolist()}
        except Exception as e:
            return {"success": False, "error": str(e)}

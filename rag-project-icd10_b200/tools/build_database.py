#!/usr/bin/env python3
"""tools/build_database.py -- drop-in mirror of /root/reference/tools/build_database.py.

Same CLI (--input, --rebuild, --verify-only; exit code 0/1), same DatabaseBuilder methods, same
CSV -> record rules (hierarchy parse :128-154, semantic text :156-171, insert batch size
:183-192), same text per record ("query: " + semantic_text through encode_query's rule) and the
same insertion order and insert batch boundaries.  The one deliberate change:
the reference encodes ONE text per call, 40 474 serial batch-1 forwards; here the texts of many
insert batches go through the GPU encoder together, length-bucketed, and are then inserted
batch by batch in the original order.
"""
from __future__ import annotations

import csv
import os
import sys
from typing import Any, Dict, List

try:
    from loguru import logger
except Exception:  # pragma: no cover
    import logging
    logger = logging.getLogger("icd10_b200")

_HERE = os.path.dirname(os.path.abspath(__file__))
if __package__ in (None, ""):
    # executed as a script: make the hyphenated package importable
    import importlib
    sys.path.insert(0, os.path.dirname(os.path.dirname(_HERE)))
    _pkg = importlib.import_module("rag-project-icd10_b200")
    EmbeddingService = importlib.import_module("rag-project-icd10_b200.services.embedding_service").EmbeddingService
    MilvusService = importlib.import_module("rag-project-icd10_b200.services.milvus_service").MilvusService
else:
    from ..services.embedding_service import EmbeddingService
    from ..services.milvus_service import MilvusService

ENCODE_CHUNK = 8192  # records encoded per GPU call (a multiple of every insert batch size)


class DatabaseBuilder:
    def __init__(self):
        self.embedding_service = None
        self.milvus_service = None
        try:
            logger.add("logs/database_build.log", rotation="50 MB", level="INFO")
        except Exception:
            pass

    # reference :30-60
    def initialize_services(self):
        logger.info("初始化服务...")
        try:
            self.embedding_service = EmbeddingService()
            probe = self.embedding_service.test_embedding("测试")
            if not probe.get("success"):
                raise Exception(f"向量化服务测试失败: {probe.get('error')}")
            logger.info(f"向量化模型加载成功: {self.embedding_service.get_model_info()}")
            self.milvus_service = MilvusService(embedding_service=self.embedding_service)
            conn = self.milvus_service.test_connection()
            if not conn.get("connected"):
                raise Exception(f"Milvus连接失败: {conn.get('error')}")
            logger.info(f"向量库连接成功: {conn}")
            logger.info(f"向量维度: {self.milvus_service.dimension}")
        except Exception as e:
            logger.error(f"服务初始化失败: {e}")
            raise

    # reference :62-126 (pandas.read_csv(encoding='utf-8') keeps the BOM out of the header the same
    # way utf-8-sig does; rows whose code/disease are empty or the string 'nan' are skipped)
    def load_csv_data(self, input_file: str) -> List[Dict]:
        logger.info(f"开始加载数据: {input_file}")
        try:
            records: List[Dict[str, Any]] = []
            titles_seen: Dict[str, str] = {}
            with open(input_file, "r", encoding="utf-8-sig", newline="") as fh:
                rows = list(csv.DictReader(fh))
            logger.info(f"成功读取 {len(rows)} 条记录")
            for row in rows:
                code = str(row.get("code", "") if row.get("code") is not None else "nan").strip()
                disease = str(row.get("disease", "") if row.get("disease") is not None else "nan").strip()
                if not code or not disease or code == "nan" or disease == "nan":
                    continue
                main_code, secondary_code, has_complication = code, "", False
                if "+" in code and "*" in code:
                    pieces = code.split("+")
                    if len(pieces) == 2:
                        main_code = pieces[0].strip()
                        secondary_code = pieces[1].replace("*", "").strip()
                        has_complication = True
                level, parent_code, category_path = self._parse_hierarchy(code, titles_seen)
                records.append({
                    "code": code,
                    "preferred_zh": disease,
                    "main_code": main_code,
                    "secondary_code": secondary_code,
                    "has_complication": has_complication,
                    "level": level,
                    "parent_code": parent_code,
                    "category_path": category_path,
                    "semantic_text": self._build_semantic_text(code, disease, category_path, titles_seen),
                })
                titles_seen[code] = disease
            logger.info(f"转换完成，获得 {len(records)} 条有效记录")
            self._log_hierarchy_stats(records)
            return records
        except Exception as e:
            logger.error(f"数据加载失败: {e}")
            raise

    # reference :128-154
    def _parse_hierarchy(self, code: str, parent_info: Dict[str, str]) -> tuple:
        if "." not in code:
            return 1, "", code
        head = code.split(".")[0]
        after = code.split(".")[1]
        if code.count(".") == 1 and len(after) <= 1:
            return 2, head, f"{head} > {code}"
        if len(after) >= 3:
            mid = f"{head}.{after[0]}"
            return 3, mid, f"{head} > {mid} > {code}"
        return 3, head, f"{head} > {code}"

    # reference :156-171
    def _build_semantic_text(self, code: str, disease: str, category_path: str, parent_info: Dict[str, str]) -> str:
        parts = [disease]
        for ancestor in category_path.split(" > ")[:-1]:
            title = parent_info.get(ancestor)
            if title is not None and title not in parts:
                parts.append(title)
        parts.append(f"ICD-10: {code}")
        return " | ".join(parts)

    def _log_hierarchy_stats(self, records: List[Dict]):
        counts = {1: 0, 2: 0, 3: 0}
        for r in records:
            if r.get("level", 0) in counts:
                counts[r["level"]] += 1
        logger.info(f"层级统计 - 主类: {counts[1]}, 亚类: {counts[2]}, 细分类: {counts[3]}")

    # reference :183-192
    def _calculate_optimal_batch_size(self, total_records: int) -> int:
        if total_records < 1000:
            return 32
        if total_records < 10000:
            return 64
        if total_records < 50000:
            return 128
        return 256

    # reference :194-260
    def vectorize_and_index(self, records: List[Dict]) -> bool:
        logger.info(f"开始向量化 {len(records)} 条记录")
        try:
            batch_size = self._calculate_optimal_batch_size(len(records))
            total_batches = (len(records) + batch_size - 1) // batch_size
            logger.info(f"开始批量向量化，每批 {batch_size} 条，共 {total_batches} 批")
            chunk = max(batch_size, (ENCODE_CHUNK // batch_size) * batch_size)
            batch_idx = 0
            for lo in range(0, len(records), chunk):
                part = records[lo:lo + chunk]
                texts = [r.get("semantic_text", r.get("preferred_zh", "")) for r in part]
                try:
                    vectors = self.embedding_service.encode_queries(texts)     # one GPU pass
                    failed = False
                except Exception as e:
                    logger.error(f"记录 {part[0].get('code')}.. 向量化失败: {e}")
                    failed = True
                for blo in range(0, len(part), batch_size):
                    batch_records = part[blo:blo + batch_size]
                    if failed:
                        # the reference substitutes plain-list zero vectors (:231-232), which
                        # insert_records then rejects (no .tolist()) -> the build aborts
                        batch_embeddings = [[0.0] * self.milvus_service.dimension for _ in batch_records]
                    else:
                        batch_embeddings = [vectors[blo + i] for i in range(len(batch_records))]
                    batch_idx += 1
                    if not self.milvus_service.insert_records(batch_records, batch_embeddings):
                        logger.error(f"批次 {batch_idx} 插入失败...")
                        return False
                logger.info(f"✅ 已处理 {min(lo + chunk, len(records))}/{len(records)} 条记录")
            logger.info("向量化和索引建立完成")
            if not self.milvus_service.load_collection():
                logger.warning("集合加载失败，但数据插入成功")
            return True
        except Exception as e:
            logger.error(f"向量化和索引失败: {e}")
            return False

    # reference :262-295
    def verify_database(self) -> Dict[str, Any]:
        logger.info("验证数据库状态...")
        try:
            stats = self.milvus_service.get_collection_stats()
            if not self.milvus_service.load_collection():
                logger.warning("集合加载失败，可能影响搜索结果")
            probe = self.embedding_service.encode_query("急性胃肠炎")
            hits = self.milvus_service.search(probe, top_k=5)
            result = {"database_stats": stats,
                      "search_test": {"query": "急性胃肠炎", "results_count": len(hits),
                                      "top_results": hits[:3] if hits else []}}
            logger.info(f"数据库验证完成: {result}")
            return result
        except Exception as e:
            logger.error(f"数据库验证失败: {e}")
            return {"error": str(e)}

    # reference :297-337
    def build_full_database(self, input_file: str = "data/ICD_10v601.csv", rebuild: bool = False) -> bool:
        logger.info("开始完整构建ICD诊断数据库")
        try:
            self.initialize_services()
            if rebuild:
                logger.info("重建模式：清空现有数据")
                self.milvus_service.clear_collection()
            else:
                logger.info("增量模式：基于现有数据")
            records = self.load_csv_data(input_file)
            if not self.vectorize_and_index(records):
                logger.error("向量化失败")
                return False
            verification = self.verify_database()
            if "error" in verification:
                logger.error(f"数据库验证失败: {verification['error']}")
                return False
            logger.info("数据库构建完成!")
            logger.info(f"最终统计: {verification['database_stats']}")
            return True
        except Exception as e:
            logger.error(f"数据库构建失败: {e}")
            return False


# reference :340-385
def main(argv=None):
    import argparse
    parser = argparse.ArgumentParser(description="ICD数据库构建工具（简化版）")
    parser.add_argument("--input", default="data/ICD_10v601.csv", help="输入CSV文件路径")
    parser.add_argument("--rebuild", action="store_true", help="重建数据库（清空现有数据）")
    parser.add_argument("--verify-only", action="store_true", help="仅验证现有数据库")
    args = parser.parse_args(argv)
    builder = DatabaseBuilder()
    try:
        if args.verify_only:
            builder.initialize_services()
            verification = builder.verify_database()
            if "error" not in verification:
                print("数据库状态正常")
                return True
            logger.error(f"数据库验证失败: {verification['error']}")
            return False
        if builder.build_full_database(args.input, rebuild=args.rebuild):
            print("数据库构建完成")
            return True
        logger.error("数据库构建失败")
        return False
    except KeyboardInterrupt:
        print("操作已中断")
        return False
    except Exception as e:
        logger.error(f"运行出错: {e}")
        print(f"错误: {e}")
        return False


if __name__ == "__main__":
    sys.exit(0 if main() else 1)

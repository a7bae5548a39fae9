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

ENCODE_CHUNK = 65536  # records encoded per GPU call (a multiple of every insert batch size): the whole ICD table in one pass


class DatabaseBuilder:
    def __init__(self, rank: int = 0, world: int = 1, dist=None):
        """rank / world / dist: one process per GPU (torchrun); the default is the reference's single process."""
        self.embedding_service = None
        self.milvus_service = None
        self.rank, self.world, self.dist = int(rank), int(world), dist
        self.shard_group = None
        try:
            logger.add("logs/database_build.log", rotation="50 MB", level="INFO")
        except Exception:
            pass

    # reference :30-60
    def initialize_services(self):
        logger.info("初始化服务...")
        try:
            if self.world > 1:
                # one encoder replica and one row shard of the table per GPU
                local = int(os.environ.get("LOCAL_RANK", self.rank))
                os.environ["EMBEDDING_DEVICE"] = f"cuda:{local}"
                MilvusService.client_kwargs = {"device": local, "shard": (self.rank, self.world)}
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

    def _barrier(self):
        if self.dist is not None and self.world > 1:
            self.dist.barrier()

    # SURVEY 8e (encoder row): the same build with the records split over the GPUs of one box.  Shard r encodes exactly
    # the rows it will hold -- rows shard_rows(n, r, world), in CSV order, so global row ids equal the single-process
    # build's -- appends them to ITS device table straight from the encoder's output buffer, and writes its slice of the
    # flat column files; rank 0 writes the scalar columns and commits the header.  No collective on the data path.
    def vectorize_and_index_sharded(self, records: List[Dict]) -> bool:
        import importlib
        import time
        store = self.milvus_service.client
        name = self.milvus_service.collection_name
        shard = importlib.import_module(__package__.rsplit(".", 1)[0] + ".engine.shard" if __package__ else
                                        "rag-project-icd10_b200.engine.shard")
        lo, hi = shard.shard_bounds(len(records), self.rank, self.world)
        stats = {"records": len(records), "rows_local": hi - lo, "encode_s": 0.0, "insert_s": 0.0, "load_s": 0.0}
        self.last_build_stats = stats
        try:
            logger.info(f"分片向量化: rank {self.rank}/{self.world} 处理记录 [{lo}, {hi})")
            if self.rank == 0:
                rows_all = [self.milvus_service._row_of_record(r) for r in records]
                store.sharded_append_prepare(name, rows_all)
            self._barrier()
            part = records[lo:hi]
            t0 = time.perf_counter()
            texts = [r.get("semantic_text", r.get("preferred_zh", "")) for r in part]
            vectors = self.embedding_service.encode_queries_device(texts)       # stays on this GPU
            t1 = time.perf_counter()
            store.sharded_append_slice(name, lo, [self.milvus_service._row_of_record(r) for r in part], vectors)
            stats["encode_s"], stats["insert_s"] = t1 - t0, time.perf_counter() - t1
            self._barrier()
            if self.rank == 0:
                store.sharded_append_commit(name, len(records), is_writer=True)
            self._barrier()
            if self.rank != 0:
                store.sharded_append_commit(name, len(records), is_writer=False)
            # serve: merged top-k over all shards (engine/shard.py: local scan, peer-store exchange, merge)
            self.shard_group = shard.ShardGroup(store.cols[name].index, row_offset=lo, rank=self.rank, world=self.world)
            store.attach_group(name, self.shard_group)
            logger.info(f"分片构建完成: 向量化 {stats['encode_s']:.2f}s, 插入 {stats['insert_s']:.2f}s")
            return True
        except Exception as e:
            logger.error(f"向量化和索引失败: {e}")
            return False

    # reference :194-260
    def vectorize_and_index(self, records: List[Dict]) -> bool:
        if self.world > 1:
            return self.vectorize_and_index_sharded(records)
        logger.info(f"开始向量化 {len(records)} 条记录")
        import time
        stats = {"records": len(records), "encode_s": 0.0, "insert_s": 0.0, "load_s": 0.0}
        self.last_build_stats = stats
        try:
            batch_size = self._calculate_optimal_batch_size(len(records))
            total_batches = (len(records) + batch_size - 1) // batch_size
            logger.info(f"开始批量向量化，每批 {batch_size} 条，共 {total_batches} 批")
            chunk = max(batch_size, (ENCODE_CHUNK // batch_size) * batch_size)
            batch_idx = 0
            for lo in range(0, len(records), chunk):
                part = records[lo:lo + chunk]
                texts = [r.get("semantic_text", r.get("preferred_zh", "")) for r in part]
                t0 = time.perf_counter()
                try:
                    vectors = self.embedding_service.encode_queries(texts)     # one GPU pass
                    failed = False
                except Exception as e:
                    logger.error(f"记录 {part[0].get('code')}.. 向量化失败: {e}")
                    failed = True
                t1 = time.perf_counter()
                stats["encode_s"] += t1 - t0
                if failed:
                    # the reference substitutes plain-list zero vectors (:231-232), which insert_records then
                    # rejects (no .tolist()) -> the build aborts
                    for blo in range(0, len(part), batch_size):
                        batch_records = part[blo:blo + batch_size]
                        batch_idx += 1
                        zeros = [[0.0] * self.milvus_service.dimension for _ in batch_records]
                        if not self.milvus_service.insert_records(batch_records, zeros):
                            logger.error(f"批次 {batch_idx} 插入失败...")
                            return False
                else:
                    # the records of this chunk in their original order, as ONE array: the rows the reference inserts
                    # batch by batch (:246-252) end up in the same order with the same auto ids
                    batch_idx += (len(part) + batch_size - 1) // batch_size
                    if not self.milvus_service.insert_records_array(part, vectors):
                        logger.error(f"批次 {batch_idx} 插入失败...")
                        return False
                stats["insert_s"] += time.perf_counter() - t1
                logger.info(f"✅ 已处理 {min(lo + chunk, len(records))}/{len(records)} 条记录")
            logger.info("向量化和索引建立完成")
            t2 = time.perf_counter()
            if not self.milvus_service.load_collection():
                logger.warning("集合加载失败，但数据插入成功")
            stats["load_s"] = time.perf_counter() - t2
            logger.info(f"构建耗时: 向量化 {stats['encode_s']:.2f}s, 插入 {stats['insert_s']:.2f}s, 加载 {stats['load_s']:.2f}s")
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
            if self.world > 1 and self.rank != 0:
                self._barrier()            # rank 0 creates the collection first; the others then map it
            self.initialize_services()
            if self.world > 1 and self.rank == 0:
                self._barrier()
            if rebuild:
                logger.info("重建模式：清空现有数据")
                if self.rank == 0:
                    self.milvus_service.clear_collection()
                self._barrier()
                if self.rank != 0:     # map the collection rank 0 has just re-created
                    self.milvus_service._connect()
                    self.milvus_service._setup_collection()
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
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dist = None
    if world > 1:
        # launched with torch.distributed.run, one process per GPU: data-parallel encode, row-sharded table
        import torch
        import torch.distributed as dist
        if not dist.is_initialized():
            local = int(os.environ.get("LOCAL_RANK", "0"))
            if torch.cuda.is_available():
                torch.cuda.set_device(local)
                dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            else:
                dist.init_process_group("gloo")
    builder = DatabaseBuilder(rank=int(os.environ.get("RANK", "0")), world=world, dist=dist)
    own_group = world > 1
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
    finally:
        if own_group and dist is not None and dist.is_initialized():
            try:
                if builder.shard_group is not None:
                    builder.shard_group.close()
                dist.barrier()
                dist.destroy_process_group()
            except Exception:
                pass


if __name__ == "__main__":
    sys.exit(0 if main() else 1)

"""Oracle: host-side text rules of the reference (test infrastructure only).

Follows /root/reference/services/embedding_service.py:68-73 (passage prefix),
:117-120 (query prefix) and /root/reference/tools/build_database.py:62-192
(CSV -> records, hierarchy parse, semantic text, insert batch size).
Checked against the reference's own DatabaseBuilder by tests/golden/make_golden.py.
"""
from __future__ import annotations

import csv
from typing import Dict, List, Tuple


def passage_text(text: str) -> str:
    # embedding_service.py:68-73
    if text.startswith("query:") or text.startswith("passage:"):
        return text
    return "passage: " + text


def query_text(text: str) -> str:
    # embedding_service.py:119 -- unconditional prefix
    return "query: " + text


def level_weight(level) -> float:
    # milvus_service.py:550-558
    return {1: 1.2, 2: 1.0, 3: 0.8}.get(level, 1.0)


def parse_hierarchy(code: str) -> Tuple[int, str, str]:
    # build_database.py:128-154
    if "." not in code:
        return 1, "", code
    head, _, tail = code.partition(".")
    if code.count(".") == 1 and len(tail) <= 1:
        return 2, head, head + " > " + code
    second = code.split(".")[1]
    if len(second) >= 3:
        parent = head + "." + second[0]
        return 3, parent, head + " > " + parent + " > " + code
    return 3, head, head + " > " + code


def semantic_text(code: str, disease: str, category_path: str, seen: Dict[str, str]) -> str:
    # build_database.py:156-171
    parts = [disease]
    for anc in category_path.split(" > ")[:-1]:
        title = seen.get(anc)
        if title is not None and title not in parts:
            parts.append(title)
    parts.append("ICD-10: " + code)
    return " | ".join(parts)


def load_records(csv_path: str) -> List[dict]:
    """build_database.py:62-126 without pandas: utf-8 with BOM, columns code,disease."""
    out: List[dict] = []
    seen: Dict[str, str] = {}
    with open(csv_path, "r", encoding="utf-8-sig", newline="") as fh:
        for row in csv.DictReader(fh):
            code = str(row.get("code", "") or "").strip()
            disease = str(row.get("disease", "") or "").strip()
            if not code or not disease or code == "nan" or disease == "nan":
                continue
            main_code, secondary, combo = code, "", False
            if "+" in code and "*" in code:
                pieces = code.split("+")
                if len(pieces) == 2:
                    main_code = pieces[0].strip()
                    secondary = pieces[1].replace("*", "").strip()
                    combo = True
            level, parent, path = parse_hierarchy(code)
            out.append({
                "code": code,
                "preferred_zh": disease,
                "main_code": main_code,
                "secondary_code": secondary,
                "has_complication": combo,
                "level": level,
                "parent_code": parent,
                "category_path": path,
                "semantic_text": semantic_text(code, disease, path, seen),
            })
            seen[code] = disease
    return out


def insert_batch_size(total: int) -> int:
    # build_database.py:183-192
    if total < 1000:
        return 32
    if total < 10000:
        return 64
    if total < 50000:
        return 128
    return 256

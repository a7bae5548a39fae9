"""Multi-GPU row-sharded search (NCCL all-gather and fused peer-store exchange) == single table."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_sharded_search_equals_single_table():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    world = 8 if n >= 8 else (4 if n >= 4 else 2)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "shard_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "SHARD_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]


@pytest.mark.gpu
def test_sharded_build_equals_single_gpu_build(tmp_path):
    """tools/build_database.py under torchrun: every rank encodes the rows it will hold and keeps them on its GPU; the
    column files, the search results (through the shard group) and a reload equal the single-GPU build."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    world = 8 if n >= 8 else (4 if n >= 4 else 2)
    sys.path.insert(0, ROOT)
    from oracle import encoder as oenc
    from oracle import text as otext
    rows = 3000
    src = open(os.path.join(ROOT, "data", "ICD_10v601.csv"), encoding="utf-8-sig").read().splitlines()
    (tmp_path / "subset.csv").write_text("﻿" + "\n".join(src[:rows + 1]) + "\n", encoding="utf-8")
    recs = otext.load_records(str(tmp_path / "subset.csv"))
    texts = [otext.query_text(r["semantic_text"]) for r in recs] + [otext.query_text("急性胃肠炎 发热 结核性脑膜炎 伤寒")]
    vocab = oenc.make_vocab(texts)
    oenc.save_hf_dir(str(tmp_path / "model"), oenc.synthetic_state_dict(seed=4, num_layers=4, vocab_size=len(vocab)), vocab, 4)
    env = dict(os.environ, ICD_TEST_WORKDIR=str(tmp_path), ICD_TEST_ROWS=str(len(recs)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "tests", "build_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert out.returncode == 0 and "BUILD_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]

#!/usr/bin/env python3
"""bench.py -- headline benchmark of the retrieval hot path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            our arm (libicdrag.so on B200)
  python bench.py --impl reference --gpus N --steps K ...  the reference's CPU path (oracle)

Workload (config.workload): exact top-10 inner-product search of a batch of 1024 synthetic
768-d bf16 queries over a synthetic 100 M x 768 bf16 corpus -- BASELINE.json configs[4].  The
corpus is row-sharded over the N GPUs (strong scaling: total rows fixed, 100 M / N per GPU; at
N = 1 the whole 153.6 GB table sits in one B200's 180 GB).  One "step" = one batch through the
scan.  `value` is queries/s with queries and corpus resident in HBM; `e2e` is the same metric
through the C ABI with HOST query/result buffers (H2D + D2H inside the timed region).

The corpus (15x-1200x the 126 MB L2) is far larger than L2, so every step streams it from HBM:
no L2 flush is needed between iterations (config.l2 says so).
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

DIM = 768
K_TOP = 10


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            d = json.load(fh)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def _ncu_traffic(kernel, rows_local, batch):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture of this
    exact workload (profiles/ncu_traffic.json), or (None, None) when no capture matches."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as fh:
            ent = json.load(fh).get(f"{kernel}:{rows_local}:{batch}")
        return (ent["bytes"], ent["capture"]) if ent else (None, None)
    except Exception:
        return None, None


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []       # (arrival time, csv line)
        self.t0 = self.t1 = None

    def begin(self):
        """Start of the load window (the process was started earlier: nvidia-smi takes ~0.2 s to come up)."""
        self.t0 = time.time()

    def end(self):
        self.t1 = time.time()

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        t0 = self.t0 if self.t0 is not None else 0.0
        t1 = self.t1 if self.t1 is not None else time.time()
        for ts, ln in self.lines:
            if ts < t0 + 0.05 or ts > t1 + 0.02:   # a sample describes the period before it arrives
                continue
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def make_corpus(torch, rows, device, seed):
    """i.i.d. N(0,1) -> fp32 L2-normalise -> bf16 (the bf16 values ARE the corpus);
    level bytes ~ Categorical(0.1243, 0.2991, 0.5766).  Built in 1 M-row chunks."""
    g = torch.Generator(device=device).manual_seed(seed)
    table = torch.empty((rows, DIM), dtype=torch.bfloat16, device=device)
    chunk = 1 << 20
    for lo in range(0, rows, chunk):
        n = min(chunk, rows - lo)
        x = torch.randn((n, DIM), generator=g, device=device, dtype=torch.float32)
        x = torch.nn.functional.normalize(x, dim=1)
        table[lo:lo + n] = x.to(torch.bfloat16)
        del x
    u = torch.rand((rows,), generator=g, device=device)
    levels = torch.full((rows,), 3, dtype=torch.uint8, device=device)
    levels[u < 0.1243 + 0.2991] = 2
    levels[u < 0.1243] = 1
    return table, levels


def make_queries(torch, batch, device, seed=999):
    g = torch.Generator(device=device).manual_seed(seed)
    q = torch.randn((batch, DIM), generator=g, device=device, dtype=torch.float32)
    return torch.nn.functional.normalize(q, dim=1).to(torch.bfloat16)


def _use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm is meant to use every host core, so undo
    that before numpy / BLAS load (and through threadpoolctl if they already have)."""
    cores = os.cpu_count() or 1
    for var in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[var] = str(cores)
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=cores)
    except Exception:
        pass
    return cores


def cpu_sample(rows, batch, reps=1):
    """The reference's CPU search path (Milvus FLAT/IP restated: oracle.search.exact_topk, fp32
    numpy over all host threads) on a bounded sample of the workload."""
    _use_all_host_threads()
    import numpy as np
    from oracle import search as osearch
    rng = np.random.default_rng(1234)
    corpus = rng.standard_normal((rows, DIM), dtype=np.float32)
    corpus /= np.linalg.norm(corpus, axis=1, keepdims=True)
    corpus = osearch.bf16_round(corpus)
    q = rng.standard_normal((batch, DIM), dtype=np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    osearch.fast_topk(corpus[: rows // 8], q[:8], K_TOP)  # warm BLAS threads
    t0 = time.perf_counter()
    for _ in range(reps):
        osearch.fast_topk(corpus, q, K_TOP)
    dt = (time.perf_counter() - t0) / reps
    return dt


def cpu_encoder_sample(n_sent=256, seq=64, batch=32):
    """The reference's CPU encoder path (SentenceTransformer.encode restated: HF BertModel fp32 + masked mean +
    L2 normalise, oracle/encoder.py) on a bounded sample: n_sent synthetic sentences of `seq` tokens in batches of
    32 (encode_batch's batch_size, embedding_service.py:97-102), all host threads."""
    cores = _use_all_host_threads()
    import torch
    torch.set_num_threads(cores)
    from transformers import BertModel
    from oracle import encoder as oenc
    torch.manual_seed(0)
    model = BertModel(oenc.bert_config(12, 21128, 512), add_pooling_layer=False).eval()
    g = torch.Generator().manual_seed(7)
    ids = torch.randint(1000, 21128, (n_sent, seq), generator=g)
    ids[:, 0], ids[:, -1] = 101, 102
    mask = torch.ones_like(ids)

    def run():
        with torch.no_grad():
            for lo in range(0, n_sent, batch):
                h = model(input_ids=ids[lo:lo + batch], attention_mask=mask[lo:lo + batch]).last_hidden_state
                m = mask[lo:lo + batch].unsqueeze(-1).float()
                torch.nn.functional.normalize((h * m).sum(1) / m.sum(1).clamp(min=1e-9), p=2, dim=1)
    run()   # warm-up (thread pools, allocator)
    t0 = time.perf_counter()
    run()
    dt = time.perf_counter() - t0
    # batch 1, what encode_query does once per ICD row in the reference's build loop (build_database.py:221)
    n1 = 32
    with torch.no_grad():
        model(input_ids=ids[:1], attention_mask=mask[:1])
        t1 = time.perf_counter()
        for i in range(n1):
            h = model(input_ids=ids[i:i + 1], attention_mask=mask[i:i + 1]).last_hidden_state
            torch.nn.functional.normalize(h.mean(1), p=2, dim=1)
        dt1 = time.perf_counter() - t1
    return {"value": n_sent / dt, "unit": "sentences/s", "cores": cores, "kind": "port",
            "batch1_sentences_per_s": n1 / dt1,
            "sample": f"{n_sent} synthetic sentences x {seq} tokens, batch {batch} (and {n1} at batch 1), 12-layer HF "
                      f"BertModel fp32 (oracle/encoder.py arithmetic), torch {torch.get_num_threads()} threads"}


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path (restated: numpy
    fp32 exact IP + top-k; pymilvus/milvus-lite are not installable offline), host cores only."""
    if rank != 0:
        return
    cores = _use_all_host_threads()
    import torch
    torch.set_num_threads(cores)
    total_rows = args.rows
    sample_rows, sample_batch = args.cpu_rows, args.cpu_batch
    times = []
    for i in range(args.warmup + args.steps):
        dt = cpu_sample(sample_rows, sample_batch)
        if i >= args.warmup:
            times.append(dt)
    dt = sum(times) / len(times)
    # one step of the real workload = batch x total_rows; the sample is batch' x rows' of the same
    # arithmetic, cost linear in rows x batch (exact scan): extrapolate linearly and say so.
    qps = sample_batch / dt * (sample_rows / total_rows)
    line = {
        "impl": "reference", "metric": "top-10 cosine QPS (100M x 768)", "value": qps, "unit": "queries/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3 * (total_rows / sample_rows) * (args.batch / sample_batch),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"exact top-{K_TOP} IP, {total_rows} x {DIM} corpus, batch {args.batch}",
                   "rows": total_rows, "dim": DIM, "batch": args.batch, "k": K_TOP},
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port",
                         "torch_threads": torch.get_num_threads(),
                         "sample": f"{sample_batch} queries x {sample_rows} rows per step (numpy fp32 GEMM + argpartition "
                                   f"top-k, oracle/search.py::fast_topk), extrapolated linearly in rows to {total_rows}"},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if not args.no_encoder:
        try:
            cb = cpu_encoder_sample()
            line["encoder"] = {"metric": "text2vec sentences/sec", "value": cb["value"], "unit": "sentences/s",
                               "batch": 32, "seq_len": 64, "layers": 12, "dtype": "f32", "cpu_baseline": cb}
        except Exception as e:
            line["encoder"] = {"error": repr(e)[:200]}
    emit(line)


_REAL_STDOUT = None


def quiet_stdout():
    """Libraries (NCCL prints its version banner) write to fd 1; the contract is ONE JSON line on
    stdout, so fd 1 is pointed at stderr for the run and the line goes to the saved descriptor."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=100_000_000, help="total corpus rows over all GPUs")
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--path", type=int, default=0, help="0 auto, 1 stream, 2 tensor")
    ap.add_argument("--exchange", type=int, default=1, help="multi-GPU candidate exchange: 0 NCCL all-gather, 1 peer stores")
    ap.add_argument("--cpu-rows", type=int, default=2_000_000)
    ap.add_argument("--cpu-batch", type=int, default=256)
    ap.add_argument("--tune", default="", help="comma list of icd_tune knobs, e.g. scan_sample=0,scan_drift=0")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-encoder", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    quiet_stdout()
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    native = importlib.import_module("rag-project-icd10_b200._native")
    VectorIndex = importlib.import_module("rag-project-icd10_b200.engine.index").VectorIndex
    native.require_gpu()
    for kv in filter(None, args.tune.split(",")):
        key, _, val = kv.partition("=")
        native.tune(**{key.strip(): int(val)})
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run"

    peaks = _peaks()
    rows_total = args.rows
    lo = rows_total * rank // world
    hi = rows_total * (rank + 1) // world
    rows_local = hi - lo
    workload = f"exact top-{K_TOP} IP, {rows_total} x {DIM} bf16 corpus row-sharded over {world} GPU(s), batch {args.batch}"
    try:
        table, levels = make_corpus(torch, rows_local, dev, 1234 + rank)
    except torch.OutOfMemoryError:
        if world > 1:
            raise
        rows_total = rows_local = 10_000_000
        lo, hi = 0, rows_total
        workload = (f"exact top-{K_TOP} IP, {rows_total} x {DIM} bf16 corpus (configs[3]: 100 M rows did not fit this "
                    f"GPU), batch {args.batch}")
        torch.cuda.empty_cache()
        table, levels = make_corpus(torch, rows_local, dev, 1234)
    q_dev = make_queries(torch, args.batch, dev)
    q_host = q_dev.cpu().pin_memory()

    idx = VectorIndex(DIM, device=local_rank)
    idx.adopt(table, levels)
    idx.set_timing(True)
    group = None
    if world > 1:
        ShardGroup = importlib.import_module("rag-project-icd10_b200.engine.shard").ShardGroup
        group = ShardGroup(idx, row_offset=lo, rank=rank, world=world)

    B, k = args.batch, K_TOP
    o_score = torch.empty((B, k), dtype=torch.float32, device=dev)
    o_raw = torch.empty((B, k), dtype=torch.float32, device=dev)
    o_id = torch.empty((B, k), dtype=torch.int64, device=dev)
    h_score = torch.empty((B, k), dtype=torch.float32).pin_memory()
    h_raw = torch.empty((B, k), dtype=torch.float32).pin_memory()
    h_id = torch.empty((B, k), dtype=torch.int64).pin_memory()
    stream = torch.cuda.current_stream(dev)

    def step_device():
        if group is not None:
            group.search(q_dev, k, out=(o_score, o_raw, o_id), path=args.path, exchange=args.exchange,
                         stream=stream.cuda_stream, sync=False)
        else:
            idx.search(q_dev, k, weight_mode=native.WEIGHT_RERANK, path=args.path, out=(o_score, o_raw, o_id),
                       stream=stream.cuda_stream, sync=False)

    def step_host():
        if group is not None:
            group.search(q_host, k, out=(h_score, h_raw, h_id), path=args.path, exchange=args.exchange,
                         stream=stream.cuda_stream, sync=True)
        else:
            idx.search(q_host, k, weight_mode=native.WEIGHT_RERANK, path=args.path, out=(h_score, h_raw, h_id),
                       stream=stream.cuda_stream, sync=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ------------------------------------------------ device-resident timing (value)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step_device()
    barrier()
    launches0 = native.lib().icd_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    scan_us = []
    barrier()
    idx.set_timing(True)   # restart the library's per-search event ring: the timed steps are what gets averaged
    sampler.begin()
    e0.record(stream)
    for _ in range(args.steps):
        step_device()
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = native.lib().icd_launch_count() - launches0
    # scan-kernel time averaged over the timed steps (events recorded by the library on the launch stream)
    tm = idx.mean_timing()
    tm["launches"] = idx.last_timing()["launches"]
    t = torch.tensor([ms_total, tm["scan_us"]], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, scan_us_max = float(t[0]), float(t[1])
    # nvidia-smi samples every 100 ms: a timed region shorter than ~1 s is followed by the SAME steps, untimed, until
    # the load window is 1 s long, so the clocks line always describes this workload under load (every rank runs the
    # same number of extra steps: the count comes from the max-over-ranks time)
    extra = 0
    if ms_total < 1000.0:
        extra = min(int((1000.0 - ms_total) / max(ms_total / args.steps, 1e-3)) + 1, 100000)
        for _ in range(extra):
            step_device()
        barrier()
    sampler.end()
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["window"] = ("timed region" if extra == 0 else
                            "timed region + %d identical untimed steps (1 s of load)" % extra)
    ms_step = ms_total / args.steps
    qps = B / (ms_step * 1e-3)

    # ------------------------------------------------ e2e through the C ABI with host buffers
    for _ in range(2):
        step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_host()
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_qps = B / (float(t[0]) / args.steps)

    # sanity: device and host paths return the same ids
    step_device()
    torch.cuda.synchronize(dev)
    same = bool(torch.equal(o_id.cpu(), h_id))

    # ------------------------------------------------ encoder (BASELINE configs[2]), data-parallel replicas
    enc = None
    if not args.no_encoder:
        try:
            enc_bench = importlib.import_module("rag-project-icd10_b200.engine.encoder").bench_encoder
            enc = enc_bench(dev, peaks, barrier=barrier)
            ok = 1.0
        except Exception as e:  # encoder line is auxiliary; the headline must still print
            enc = {"error": repr(e)[:200]}
            ok = 0.0
        t = torch.tensor([enc.get("ms_per_batch", 0.0), -ok], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if float(t[1]) == -1.0 and "error" not in enc:
            ms = float(t[0])  # max over ranks; every rank encodes its own batch (weak scaling, no exchange)
            enc.update({"ms_per_batch": ms, "value": world * enc["batch"] / (ms * 1e-3), "n_gpus": world,
                        "tflops": world * enc["flops_per_batch"] / (ms * 1e-3) / 1e12,
                        "frac_of_bf16_sustained": enc["flops_per_batch"] / (ms * 1e-3) / 1e12 / peaks["bf16_tflops_sustained"],
                        "scaling": "weak (one batch per GPU, replicas, no collective)"})
            if "roofline" in enc:   # per-GPU figure from the max-over-ranks time, like frac_of_bf16_sustained
                enc["roofline"]["achieved"] = enc["flops_per_batch"] / (ms * 1e-3) / 1e12
                enc["roofline"]["frac"] = enc["roofline"]["achieved"] / enc["roofline"]["peak"]

    if rank == 0:
        flops = 2.0 * B * rows_local * DIM                     # per scan launch (per GPU)
        bytes_alg = rows_local * DIM * 2 + B * DIM * 2 + B * k * 16
        tensor_bound = flops / (peaks["bf16_tflops_sustained"] * 1e12) > bytes_alg / (peaks["hbm_gbs"] * 1e9)
        scan_s = scan_us_max * 1e-6
        if tensor_bound:
            roof = {"bound": "tensor", "achieved": flops / scan_s / 1e12, "peak": peaks["bf16_tflops_sustained"],
                    "unit": "TFLOP/s"}
        else:
            roof = {"bound": "hbm", "achieved": bytes_alg / scan_s / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s"}
        roof["frac"] = roof["achieved"] / roof["peak"]
        roof["kernel"] = "scan_tc_kernel" if tm["launches"] and B > 4 else "scan_stream_kernel"
        roof["traffic"], roof["traffic_source"] = _ncu_traffic(roof["kernel"], rows_local, B)
        roof["algorithmic_bytes"] = bytes_alg
        roof["kernel_us"] = scan_us_max
        roof["kernel_us_over"] = "mean of %d timed launches (max over ranks)" % tm["calls"]
        roof["peak_source"] = peaks["source"] + (" (sustained)" if tensor_bound else "")
        roof["hbm_gbs_of_scan"] = bytes_alg / scan_s / 1e9
        line = {
            "metric": "top-10 cosine QPS (100M x 768)", "value": qps, "unit": "queries/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": workload, "rows": rows_total, "rows_per_gpu": rows_local, "dim": DIM, "batch": B,
                       "k": k, "weight_mode": "rerank", "l2": "corpus >> L2 (no flush needed)", "tune": args.tune or None,
                       "exchange": ("peer-store" if args.exchange else "nccl-allgather") if world > 1 else None},
            "e2e": {"value": e2e_qps, "unit": "queries/s", "h2d_bytes_per_step": B * DIM * 2,
                    "d2h_bytes_per_step": B * k * 16},
            "gpu_launches": int(launches), "launches_per_step": tm["launches"],
            "roofline": roof, "clocks": clocks, "ids_match_host_device": same,
        }
        if not args.no_cpu_baseline and world == 1:
            dt = cpu_sample(args.cpu_rows, args.cpu_batch)
            line["cpu_baseline"] = {
                "value": args.cpu_batch / dt * (args.cpu_rows / rows_total), "unit": "queries/s",
                "cores": os.cpu_count(), "kind": "port",
                "sample": f"{args.cpu_batch} queries x {args.cpu_rows} rows (numpy fp32 GEMM + argpartition top-k, "
                          f"oracle/search.py::fast_topk), extrapolated linearly in rows to {rows_total}"}
        if enc is not None:
            if not args.no_cpu_baseline and world == 1 and "error" not in enc:
                try:
                    enc["cpu_baseline"] = cpu_encoder_sample()
                except Exception as e:
                    enc["cpu_baseline"] = {"error": repr(e)[:200]}
            line["encoder"] = enc
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

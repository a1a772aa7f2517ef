#!/usr/bin/env python
"""bench.py -- Merkle-committed LDE throughput of the polynomial-batch commitment (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

One "step" = one whole PolynomialBatch::from_values of the wide batch (BASELINE configs[2]:
2^20 rows x 256 columns, rate_bits 3, Poseidon Merkle tree, cap_height 4):
iNTT -> coset LDE -> leaf-ordered transpose -> leaf hashing -> tree -> cap.
    value  Gelem/s = ncols * n * 2^rate_bits (LDE elements) per second, inputs resident in HBM,
           leaves/digests left in HBM (whole-job aggregate, max over ranks)
    e2e    the same through the host-buffer C ABI (mp2gpu_commit_from_values): pinned host
           columns in, coefficients + leaves + digests + cap out, copies inside the timed region
N > 1: the batch is column-sharded, exchanged once into row shards (by default inside the LDE kernel's
stores over NVLink, --exchange nccl for an all-to-all), hashed per rank, caps all-gathered ("scaling": "strong" -- the total work is fixed).
--impl reference times the CPU restatement of the reference path (oracle/, OpenMP, all host
threads) on a bounded row sample of the same batch; the reference itself is Rust with un-vendored
dependencies and cannot be built in this image.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

P = 0xFFFFFFFF00000001
PERM_MADS = 6700  # 32-bit multiply-adds credited per Poseidon permutation


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--hash", default="poseidon", choices=["poseidon", "poseidon2"])
    ap.add_argument("--n-log", type=int, default=20)
    ap.add_argument("--ncols", type=int, default=256)
    ap.add_argument("--rate-bits", type=int, default=3)
    ap.add_argument("--cap-height", type=int, default=4)
    ap.add_argument("--cpu-sample-log", type=int, default=16,
                    help="rows (log2) of the CPU sample; 20 = the whole wide batch (one step takes minutes on 16 cores)")
    ap.add_argument("--workload", default="wide", choices=["wide", "trace"],
                    help="wide: one 2^20 x 256 batch (BASELINE configs[2], the contract line); trace: mp2 leaf-proof "
                         "commitment traces, independent proofs per GPU (configs[1]/[4])")
    ap.add_argument("--trace", default="leaf", choices=["leaf", "aggregation"],
                    help="leaf: mp2-v1 values-extraction leaf proof (configs[1]); aggregation: 2-proof universal-verifier "
                         "aggregation (configs[3])")
    ap.add_argument("--proofs", type=int, default=64, help="trace workload: proofs per GPU per step")
    ap.add_argument("--streams", type=int, default=8, help="trace workload: proofs in flight per GPU")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N > 1: peer = LDE kernel stores into the peers' buffers over NVLink; nccl = all_to_all_single")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-map-stage", action="store_true")
    ap.add_argument("--no-whole-prover", action="store_true", help="skip the whole-prover replay inside map_stage")
    ap.add_argument("--no-parity-check", action="store_true")
    ap.add_argument("--map-proofs", type=int, default=32, help="map_stage sub-record: proofs per GPU per step")
    return ap.parse_args()


def workload_name(a):
    label = "wide batch" if (a.n_log, a.ncols) == (20, 256) else "config 1 batch" if (a.n_log, a.ncols) == (14, 135) \
        else "polynomial batch"
    return "%s: 2^%d rows x %d columns coset LDE rate_bits=%d + %s Merkle tree cap_height=%d (from_values)" % (
        label, a.n_log, a.ncols, a.rate_bits, "Poseidon" if a.hash == "poseidon" else "Poseidon2", a.cap_height)


def global_columns(torch, a, c0, count, n):
    """Columns [c0, c0 + count) of the synthetic global batch, on the current CUDA device."""
    gen = torch.Generator(device="cuda")
    out = torch.empty((count, n), dtype=torch.int64, device="cuda")
    for j in range(count):
        gen.manual_seed((0x6D7033 << 20) + c0 + j)
        out[j] = torch.randint(0, 1 << 62, (n,), dtype=torch.int64, device="cuda", generator=gen)
    return out


def perms_per_commit(ncols, N, cap_height):
    leaf = N * ((ncols + 7) // 8) if ncols > 4 else 0
    return leaf, N - (1 << cap_height)


def synthetic_columns(seed, shape):
    """Uniform field elements (SplitMix64-derived, rejection below p) -- same recipe as tests/util.py."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from util import field_elems

    return field_elems(seed, shape)


# ------------------------------------------------------------------------------------------------
# CPU restatement (the reference arm and the cpu_baseline leg): oracle/ may only be executed here
# ------------------------------------------------------------------------------------------------
def host_threads():
    """All host cores this process may use.  (torchrun exports OMP_NUM_THREADS=1 to its workers; the oracle's
    OpenMP regions take an explicit thread count, so the CPU arm still uses the whole box.)"""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_commit_time(a, sample_log, repeats=1, warm=0):
    import oracle as O

    O.build()
    threads = host_threads()
    kind = 0 if a.hash == "poseidon" else 1
    cols = synthetic_columns(0x6D7033, (a.ncols, 1 << sample_log))
    times = []
    for i in range(warm + repeats):
        t0 = time.perf_counter()
        O.commit(cols, a.rate_bits, a.cap_height, kind, False, nthreads=threads, want_leaves=True)
        dt = time.perf_counter() - t0
        if i >= warm:
            times.append(dt)
    elems = a.ncols * ((1 << sample_log) << a.rate_bits)
    return times, elems, threads


def run_reference(a, rank):
    if rank != 0:
        return
    sample_log = min(a.cpu_sample_log, a.n_log)
    times, elems, threads = cpu_commit_time(a, sample_log, repeats=a.steps, warm=a.warmup)
    total = sum(times)
    value = elems * len(times) / total / 1e9
    sample = "2^%d of 2^%d rows x %d columns per step (same rate_bits/cap/hasher)" % (sample_log, a.n_log, a.ncols)
    line = {
        "impl": "reference", "metric": "Merkle-committed LDE Gelem/s", "value": value, "unit": "Gelem/s",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64 (Goldilocks field)",
        "data": "synthetic", "config": {"workload": workload_name(a), "sample": sample},
        "cpu_baseline": {"value": value, "unit": "Gelem/s", "cores": threads, "kind": "port", "sample": sample,
                         "same_config": sample_log == a.n_log,
                         "note": "CPU restatement of plonky2's rayon path (oracle/mp2_oracle.c, OpenMP, one task per "
                                 "column / per subtree, shared root tables, split-half MDS, lazy reductions); the "
                                 "reference is Rust with un-vendored crates and cannot be compiled here.  ESTIMATED "
                                 "(not measured) gap to plonky2's AVX2 Poseidon + packed FFT on the same cores: 3-6x "
                                 "slower, so divide any GPU/CPU ratio by that before quoting it against plonky2"},
        "e2e": {"value": value, "unit": "Gelem/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(device_index), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=self.tmp, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.tmp.flush()
        self.tmp.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.tmp.read().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.tmp.name)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(a):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus:
        if world == 1 and a.gpus > 1:
            raise SystemExit("--gpus %d needs torchrun with %d ranks" % (a.gpus, a.gpus))
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    import mapreduce_plonky2_b200 as G
    from mapreduce_plonky2_b200 import device as D
    from mapreduce_plonky2_b200 import sharded as S

    G.init(local_rank)
    kind = G.POSEIDON if a.hash == "poseidon" else G.POSEIDON2
    n, N = 1 << a.n_log, (1 << a.n_log) << a.rate_bits
    c_loc = a.ncols // world
    elems = a.ncols * N
    launches0 = G.launch_count()

    # synthetic inputs, resident in HBM before the timed region (field elements < 2^62 < p).  Column j of the
    # GLOBAL batch is seeded by j alone, so every N commits the same 2^n_log x ncols batch: cap_xor must be
    # identical at N = 1, 2, 4, 8, and rank 0 can rebuild the whole batch for the single-GPU cross-check below.
    cols = global_columns(torch, a, rank * c_loc, c_loc, n)
    engine = S.CudaEngine()
    scratch = {}
    exchange_note = None
    if world > 1 and a.exchange == "peer":
        # the fused exchange needs symmetric memory; if this box cannot provide it every rank switches to NCCL
        # together (and the line says so) instead of the job dying
        ok = torch.ones(1, dtype=torch.int32, device="cuda")
        try:
            scratch["peer_exchange"] = S.PeerExchange(world, c_loc, N // world)
        except Exception as e:  # noqa: BLE001
            ok.zero_()
            exchange_note = "symmetric memory unavailable (%s: %s): NCCL all-to-all used instead" % (type(e).__name__, e)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            a.exchange = "nccl"
            scratch.pop("peer_exchange", None)
            exchange_note = exchange_note or "symmetric memory unavailable on another rank: NCCL all-to-all used instead"
            print("[bench] " + exchange_note, file=sys.stderr)
        else:
            exchange_note = None

    def step():
        if world > 1:
            return S.commit_sharded(cols, a.ncols, a.rate_bits, a.cap_height, kind, engine, scratch=scratch,
                                    exchange=a.exchange)
        return commit_solo()

    solo_bufs = D.CommitBuffers(a.ncols, a.n_log, a.rate_bits, a.cap_height, True) if world == 1 else None

    def commit_solo():
        D.commit_resident(cols, solo_bufs, kind, False)
        return solo_bufs

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # nvidia-smi is started before the warm-up so that its start-up (it takes driver locks) cannot land
    # inside the timed region; it keeps sampling through it
    sampler = ClockSampler(local_rank) if rank == 0 else None
    for _ in range(max(a.warmup, 3)):
        step()
    barrier()

    # ---- timed region: exactly K steps, CUDA events on the launching stream, max over ranks ----
    D.profile_enable(True)
    D.profile_report()  # drop warm-up records
    # The wide batch is far larger than L2 (2.1 GB in, 17.2 GB LDE): its K steps are timed back to back.  A small batch
    # (config 1: 17.7 MB in, 141 MB LDE against a 126 MB L2) would find its inputs and tables in L2 from the previous
    # step, so its steps are timed one by one with a 256 MB write between them (outside the events).
    small = 8 * elems // world < (1 << 29)
    flush = torch.empty(1 << 28, dtype=torch.uint8, device="cuda") if small else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if not small:
        e0.record()
        for _ in range(a.steps):
            res = step()
        e1.record()
        barrier()
        elapsed = e0.elapsed_time(e1)
    else:
        pairs = []
        for _ in range(a.steps):
            flush.fill_(1)
            ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ea.record()
            res = step()
            eb.record()
            pairs.append((ea, eb))
        barrier()
        elapsed = sum(x.elapsed_time(y) for x, y in pairs)
    D.profile_enable(False)
    prof = D.profile_report()
    clocks = sampler.stop() if sampler else None
    if world > 1 and os.environ.get("MP2_SHARDED_TIMING"):  # diagnostic: per-stage ms of every call, every rank
        for i, rec in enumerate(S.timing_report(scratch)):
            print("[timing rank %d call %d] %s" % (rank, i, " ".join("%s=%.2f" % kv for kv in rec.items())),
                  file=sys.stderr, flush=True)
    ms_total = torch.tensor([elapsed], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms_total, op=dist.ReduceOp.MAX)
    ms_total = float(ms_total.item())
    ms_step = ms_total / a.steps
    value = elems / (ms_step * 1e-3) / 1e9
    launches_timed = G.launch_count() - launches0

    # cap to the host = the step's result (and a checksum the reference arm could be compared with)
    cap_host = res.cap.cpu().numpy().view(np.uint64)

    line = None
    if rank == 0:
        # ---- rooflines from the per-kernel CUDA-event times of the timed region (rank 0's share) ----
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
            os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        n_loc_leaves = N // world
        leaf_perms, node_perms = perms_per_commit(a.ncols, N, a.cap_height)

        def kern(name):
            cnt, ms = prof.get(name, (0, 0.0))
            return cnt, (ms / cnt if cnt else None)

        roof = {}
        traffic = measured_traffic(a, world)
        cnt, leaf_ms = kern("k_leaf_hash")
        ip = D.int_pipe_peak()
        if leaf_ms:
            perms_launch = leaf_perms // world
            bytes_launch = 8 * a.ncols * n_loc_leaves + 32 * n_loc_leaves  # read every leaf once, write its digest
            mads = perms_launch * PERM_MADS / (leaf_ms * 1e-3) / 1e12
            roof["roofline"] = {
                "kernel": "k_leaf_hash (Poseidon sponge over leaves)", "bound": "int_pipe",
                "achieved": mads, "peak": ip["t_imad_per_s"], "unit": "T imad/s (6700 credited per permutation)",
                "frac": mads / ip["t_imad_per_s"] if ip["t_imad_per_s"] else None,
                "peak_source": "measured live: %.1f IMAD/clk/SM x 148 SM at %.0f MHz" % (ip["imad_per_clk_per_sm"], ip["sm_clock_mhz"]),
                "perm_per_s": perms_launch / (leaf_ms * 1e-3), "ms_per_launch": leaf_ms, "launches": cnt,
                "traffic": traffic.get("k_leaf_hash"), "traffic_source": traffic.get("source"),
                "traffic_note": "DRAM bytes per launch (ncu); includes the row-major `leaves` the kernel also writes "
                                "(8*c*N), which the algorithmic figure below counts under the NTT stage's LDE write",
                "hbm": {"bound": "hbm", "achieved": bytes_launch / (leaf_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": bytes_launch / (leaf_ms * 1e-3) / 1e9 / hbm_peak, "algorithmic_bytes": bytes_launch},
            }
        # NTT stage: all transform kernels of one commitment against B_ntt = 8*c*n*(3 + 2^r)
        ntt_names = ["k_intt_single", "k_lde_single", "k_pass1", "k_pass2"]
        ntt_ms = sum(prof.get(k, (0, 0.0))[1] for k in ntt_names) / a.steps
        if ntt_ms:
            b_ntt = 8 * c_loc * n * (3 + (1 << a.rate_bits))
            roof["roofline_ntt"] = {
                "kernel": "iNTT + coset LDE kernels (" + ",".join(k for k in ntt_names if k in prof) + ")", "bound": "hbm",
                "achieved": b_ntt / (ntt_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": b_ntt / (ntt_ms * 1e-3) / 1e9 / hbm_peak, "peak_source": peak_src,
                "algorithmic_bytes": b_ntt, "ms_per_step": ntt_ms, "traffic": traffic.get("ntt_stage"),
                "traffic_source": traffic.get("source"),
                "traffic_note": "DRAM bytes per step over the transform kernels (ncu); the four-step path writes and "
                                "re-reads the 8*c*N intermediate once, which B_ntt (a single-pass figure) does not contain"}
        kernel_ms = {k: {"launches": v[0], "ms_total": round(v[1], 4)} for k, v in sorted(prof.items())}

        line = {
            "metric": "Merkle-committed LDE Gelem/s", "value": value, "unit": "Gelem/s", "n_gpus": world,
            "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u64 (Goldilocks field, integer)", "data": "synthetic",
            "config": {"workload": workload_name(a), "parallelism": "columns/%d -> %s -> rows/%d" % (
                           world, "peer stores from the LDE kernel (NVLink)" if a.exchange == "peer" else "NCCL all-to-all", world)
                       if world > 1 else "single GPU", "l2": ("inputs (%.1f GB) and LDE (%.1f GB) exceed the 126 MB L2" % (8 * a.ncols * n / 1e9, 8 * elems / 1e9))
                       if not small else "L2 flushed between timed steps (256 MB write outside the events); steps timed one by one",
                       "timing": "CUDA events on the launching stream, max over ranks"},
            "gpu_launches": int(launches_timed), "clocks": clocks, "kernels": kernel_ms,
            "cap_xor": "%016x" % int(np.bitwise_xor.reduce(cap_host.reshape(-1))),
            "perms_per_step": leaf_perms + node_perms,
        }
        if exchange_note:
            line["config"]["exchange_note"] = exchange_note
        line.update(roof)

    # ---- self-verification (outside every timed region) ----
    parity = parity_check(a, torch, D, S, G, world, rank, kind, res, cap_host)
    if rank == 0:
        line["parity_check"] = parity
    # ---- the second half of the metric: map-stage proofs/s with independent replicas at this N ----
    if not a.no_map_stage:
        ms = measure_map_stage(a, torch, dist, G, world, rank)
        if rank == 0:
            line["map_stage"] = ms
    del res
    # ---- e2e: host buffers, copies inside the timed region ----
    if not a.no_e2e:
        e2e = run_e2e(a, G, D, S, torch, dist, world, rank, kind, engine, scratch, solo_bufs)
        if rank == 0:
            line["e2e"] = e2e
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        sample_log = min(a.cpu_sample_log, a.n_log)
        times, s_elems, threads = cpu_commit_time(a, sample_log, repeats=1, warm=0)
        line["cpu_baseline"] = {
            "value": s_elems / times[0] / 1e9, "unit": "Gelem/s", "cores": threads, "kind": "port",
            "sample": "2^%d of 2^%d rows x %d columns, one commitment (%.1f s)" % (sample_log, a.n_log, a.ncols, times[0]),
            "note": "restated CPU oracle (oracle/mp2_oracle.c, OpenMP) -- not plonky2 itself; estimated 3-6x slower "
                    "than plonky2's AVX2 path on the same cores"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def parity_check(a, torch, D, S, G, world, rank, kind, res, cap_host):
    """N > 1: rank 0 rebuilds the WHOLE batch (same per-column seeds), commits it on its own GPU with the
    single-GPU path and compares cap, its digest slice and its leaf rows with what the sharded run produced.
    N = 1: the commitment of a 2^12-row slice of the same columns is compared with the CPU oracle (bit-exact)."""
    if a.no_parity_check:
        return {"skipped": True}
    n, N = 1 << a.n_log, (1 << a.n_log) << a.rate_bits
    out = {}
    if world > 1:
        if rank == 0:
            full = global_columns(torch, a, 0, a.ncols, n)
            bufs = D.CommitBuffers(a.ncols, a.n_log, a.rate_bits, a.cap_height, True)
            D.commit_resident(full, bufs, kind, False)
            torch.cuda.synchronize()
            nd = res.digests.shape[0]
            out = {"cap_equals_single_gpu": bool(torch.equal(bufs.cap, res.cap)),
                   "rank0_digests_equal_single_gpu": bool(torch.equal(bufs.digests[:nd], res.digests)),
                   "rank0_leaves_equal_single_gpu": bool(torch.equal(bufs.leaves[:N // world], res.leaves)),
                   "how": "rank 0 re-generated all %d columns, ran one single-GPU mp2gpu_dev_commit and compared" % a.ncols}
            del full, bufs
            torch.cuda.empty_cache()
        return out
    # N = 1: oracle cross-check on a slice small enough for the CPU (first 2^12 rows of every column)
    import oracle as O

    O.build()
    s_log = min(12, a.n_log)
    cols = global_columns(torch, a, 0, a.ncols, n)[:, :1 << s_log].contiguous()
    bufs = D.CommitBuffers(a.ncols, s_log, a.rate_bits, a.cap_height, True)
    D.commit_resident(cols, bufs, kind, False)
    torch.cuda.synchronize()
    ref = O.commit(cols.cpu().numpy().view(np.uint64), a.rate_bits, a.cap_height, kind, False, nthreads=host_threads(),
                   want_leaves=True)
    out = {"sample_rows_log": s_log,
           "cap_equals_oracle": bool(np.array_equal(bufs.cap.cpu().numpy().view(np.uint64), ref["cap"])),
           "digests_equal_oracle": bool(np.array_equal(bufs.digests.cpu().numpy().view(np.uint64), ref["digests"])),
           "leaves_equal_oracle": bool(np.array_equal(bufs.leaves.cpu().numpy().view(np.uint64), ref["leaves"])),
           "how": "first 2^%d rows of the bench's own columns committed on the GPU and by oracle/ (CPU restatement)" % s_log}
    return out


def measure_map_stage(a, torch, dist, G, world, rank):
    """BASELINE.json's second metric, 'mp2 leaf proofs/s at 1/2/4/8 B200': independent proof traces, one replica
    set per GPU, no collective (SURVEY.md 8(e) map stage; work unit = one opaque proof per task,
    mp2-v1/src/api.rs:154-165).  Trace replay with ASSUMED degrees; Poseidon2 = the reference's
    default hasher (mp2-common/src/lib.rs:37-40)."""
    from mapreduce_plonky2_b200 import trace as T
    from mapreduce_plonky2_b200 import device as D

    streams = 8
    runner = T.TraceRunner(T.LEAF_PROOF_DEGREES, 1, streams)
    for _ in range(2):
        runner.run(streams)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 2
    l0 = G.launch_count()
    e0.record()
    for _ in range(steps):
        runner.run(a.map_proofs)
        for st in runner.streams:
            torch.cuda.current_stream().wait_stream(st)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_step = float(ms.item()) / steps
    launches = G.launch_count() - l0
    out = None
    if rank == 0:
        ip = D.int_pipe_peak()
        perms = runner.perms_per_proof
        mads = perms * a.map_proofs * PERM_MADS / (ms_step * 1e-3) / 1e12
        out = {"metric": "mp2 leaf proofs/s (commitment + FRI-tree trace)", "value": world * a.map_proofs / (ms_step * 1e-3),
               "unit": "proofs/s", "n_gpus": world, "proofs_per_gpu_per_step": a.map_proofs, "steps": steps,
               "ms_per_step": ms_step, "scaling": "weak", "parallelism": "replicas only, no collective",
               "hash": "poseidon2", "degrees": "2^14 + 2^13 + 2^12 (ASSUMED: trace replay, degrees assumed)",
               "includes": T.TRACE_INCLUDES, "perms_per_proof": perms,
               "lde_elems_per_proof": runner.lde_elems_per_proof, "gpu_launches": int(launches),
               "int_pipe_frac": (mads / ip["t_imad_per_s"]) if ip["t_imad_per_s"] else None,
               "note": "upper bound on prover throughput: witness generation and quotient evaluation are host-side"}
    del runner
    torch.cuda.empty_cache()
    # ---- the same proofs through EVERY prove() step the library implements (trace.ProverTrace): commitments from
    # pinned host columns, quotient evaluation, openings, prove_openings / FRI with PoW and query rounds; one
    # prover per host thread, replicas per GPU
    whole = None
    if not a.no_whole_prover:
        import time as _time

        nthreads, nproofs = 8, a.map_proofs
        pt = T.ProverTrace(T.LEAF_PROOF_DEGREES, 1, nthreads)
        for _ in range(3):   # the stream-ordered pool needs a few rounds until every prover thread owns its blocks
            pt.run(2 * nthreads)
        torch.cuda.synchronize()
        samples, launches = [], 0
        for _ in range(3):   # three timed runs, the median is reported (pool growth makes single runs swing)
            if world > 1:
                dist.barrier()
            l0 = G.launch_count()
            t0 = _time.perf_counter()
            pt.run(nproofs)
            dt = torch.tensor([_time.perf_counter() - t0], dtype=torch.float64, device="cuda")
            launches = G.launch_count() - l0
            if world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            samples.append(float(dt.item()))
        dt = sorted(samples)[1]
        pt.free()
        torch.cuda.empty_cache()
        whole = {"metric": "mp2 leaf proofs/s (whole device-side prover replay)", "value": world * nproofs / dt,
                 "unit": "proofs/s", "n_gpus": world, "proofs_per_gpu": nproofs, "prover_threads_per_gpu": nthreads,
                 "ms_per_proof_per_gpu": dt / nproofs * 1e3, "timing": "host wall clock around the threads, max over ranks; median of 3 runs",
                 "runs_proofs_per_s": [world * nproofs / x for x in samples],
                 "includes": T.PROVER_INCLUDES, "gpu_launches": int(launches),
                 "api": "mp2gpu_prove (one native call per prove(): csrc/prover.cpp), proof bytes = bincode(ProofWithPublicInputs)",
                 "excluded": "witness generation (host work in the reference); gates limited to the staged subset; "
                             "degrees ASSUMED; synthetic (not satisfying) witness values -- the work does not depend on them"}
    if rank == 0 and out is not None and world == 1 and not a.no_cpu_baseline:
        # the commitments-only trace on the CPU port, one proof on all host threads (oracle/: the checker, timed as a baseline)
        threads = host_threads()
        dt_cpu = cpu_trace_time(1, threads, T.LEAF_PROOF_DEGREES)
        out["cpu_baseline"] = {"value": 1.0 / dt_cpu, "unit": "proofs/s", "cores": threads, "kind": "port",
                               "sample": "one leaf-proof commitment trace (%.1f s): the same commitments and FRI-layer trees as "
                                         "`value`, no quotient / openings / FRI arithmetic" % dt_cpu}
    if rank == 0 and out is not None and whole is not None:
        out["whole_prover"] = whole
        out["note"] = ("value = commitments + FRI-layer trees only (device-resident inputs, upper bound); whole_prover = "
                       "every prove() step this library implements, from host buffers")
    return out


def run_e2e(a, G, D, S, torch, dist, world, rank, kind, engine, scratch, solo_bufs):
    """Same metric through the public API with HOST buffers.  N = 1: the C-ABI call a patched plonky2
    makes (mp2gpu_commit_from_values: pinned columns in; coefficients, leaves, digests, cap out).
    N > 1: pinned host shard -> H2D -> sharded commit -> D2H of this rank's outputs."""
    n, N = 1 << a.n_log, (1 << a.n_log) << a.rate_bits
    c_loc = a.ncols // world
    ncap = 1 << a.cap_height
    steps = max(1, min(a.steps, 3))
    h2d = 8 * c_loc * n
    d2h = 8 * c_loc * n + 8 * a.ncols * (N // world) + 32 * 2 * (N - ncap) // world + 32 * ncap
    if world == 1:
        cols_h = torch.empty((a.ncols, n), dtype=torch.int64, pin_memory=True)
        cols_h.random_(0, 1 << 62)
        coeffs_h = torch.empty((a.ncols, n), dtype=torch.int64, pin_memory=True)
        leaves_h = torch.empty((N, a.ncols), dtype=torch.int64, pin_memory=True)
        dig_h = torch.empty((2 * (N - ncap), 4), dtype=torch.int64, pin_memory=True)
        cap_h = torch.empty((ncap, 4), dtype=torch.int64, pin_memory=True)
        import ctypes as C

        from mapreduce_plonky2_b200 import _lib
        u64p = _lib.u64p

        def ptrs(t):
            return (u64p * t.shape[0])(*[C.cast(t[i].data_ptr(), u64p) for i in range(t.shape[0])])

        def call():
            _lib.call("mp2gpu_commit_from_values", ptrs(cols_h), a.ncols, a.n_log, a.rate_bits, a.cap_height, kind,
                      ptrs(coeffs_h), C.cast(leaves_h.data_ptr(), u64p), C.cast(dig_h.data_ptr(), u64p),
                      C.cast(cap_h.data_ptr(), u64p), None)

        call()  # warm-up (allocations, tables)
        t0 = time.perf_counter()
        for _ in range(steps):
            call()
        dt = (time.perf_counter() - t0) / steps
        api = "mp2gpu_commit_from_values (C ABI, pinned host buffers)"

        # variant: the leaves stay on the device behind a batch handle (get_lde_values fetches rows on demand)
        handle = C.c_void_p(None)

        def call_resident():
            _lib.call("mp2gpu_commit_from_values", ptrs(cols_h), a.ncols, a.n_log, a.rate_bits, a.cap_height, kind,
                      ptrs(coeffs_h), None, C.cast(dig_h.data_ptr(), u64p), C.cast(cap_h.data_ptr(), u64p),
                      C.byref(handle))
            _lib.load().mp2gpu_batch_free(handle)

        call_resident()
        t0 = time.perf_counter()
        for _ in range(steps):
            call_resident()
        dt_res = (time.perf_counter() - t0) / steps

        # variant: what prover.prove() calls -- only the cap comes back, coefficients / rows / digests stay behind the handle
        def call_cap():
            _lib.call("mp2gpu_commit_from_values", ptrs(cols_h), a.ncols, a.n_log, a.rate_bits, a.cap_height, kind,
                      None, None, None, C.cast(cap_h.data_ptr(), u64p), C.byref(handle))
            _lib.load().mp2gpu_batch_free(handle)

        call_cap()
        t0 = time.perf_counter()
        for _ in range(steps):
            call_cap()
        dt_cap = (time.perf_counter() - t0) / steps
    else:
        cols_h = torch.empty((c_loc, n), dtype=torch.int64, pin_memory=True)
        cols_h.random_(0, 1 << 62)
        coeffs_h = torch.empty((c_loc, n), dtype=torch.int64, pin_memory=True)
        leaves_h = torch.empty((N // world, a.ncols), dtype=torch.int64, pin_memory=True)
        dig_h = torch.empty((2 * (N - ncap) // world, 4), dtype=torch.int64, pin_memory=True)
        cap_h = torch.empty((ncap, 4), dtype=torch.int64, pin_memory=True)
        cols_d = torch.empty((c_loc, n), dtype=torch.int64, device="cuda")

        host_out = S.HostOutputs(coeffs_h, leaves_h, dig_h, cap_h, torch.cuda.Stream())

        def call():
            # upload on the compute stream, then the sharded commitment with its outputs streamed back to the
            # pinned buffers under the compute (coefficients under the LDE, leaf blocks under the hashing)
            cols_d.copy_(cols_h, non_blocking=True)
            S.commit_sharded(cols_d, a.ncols, a.rate_bits, a.cap_height, kind, engine, scratch=scratch,
                             exchange=a.exchange, host_out=host_out)
            torch.cuda.synchronize()

        call()
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            call()
        dist.barrier()
        dt = (time.perf_counter() - t0) / steps
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
        api = "sharded.commit_sharded(host_out=...) with pinned host shards (H2D + D2H per rank, D2H overlapped with hashing)"
        # the same call with the leaf rows left in HBM (every later reader of the rows runs on the device)
        host_res = S.HostOutputs(coeffs_h, None, dig_h, cap_h, host_out.copy_stream)

        def call_resident():
            cols_d.copy_(cols_h, non_blocking=True)
            S.commit_sharded(cols_d, a.ncols, a.rate_bits, a.cap_height, kind, engine, scratch=scratch,
                             exchange=a.exchange, host_out=host_res)
            torch.cuda.synchronize()

        call_resident()
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            call_resident()
        dist.barrier()
        t = torch.tensor([(time.perf_counter() - t0) / steps], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt_res = float(t.item())
        # only the cap back (prover.prove()'s call): coefficients, rows and digests stay in HBM on their ranks
        host_cap = S.HostOutputs(None, None, None, cap_h, host_out.copy_stream)

        def call_cap():
            cols_d.copy_(cols_h, non_blocking=True)
            S.commit_sharded(cols_d, a.ncols, a.rate_bits, a.cap_height, kind, engine, scratch=scratch,
                             exchange=a.exchange, host_out=host_cap)
            torch.cuda.synchronize()

        call_cap()
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            call_cap()
        dist.barrier()
        t = torch.tensor([(time.perf_counter() - t0) / steps], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt_cap = float(t.item())
    elems = a.ncols * N
    d2h_res = d2h - 8 * a.ncols * (N // world)
    # Headline e2e: the call a patched plonky2 makes now that every later reader of the LDE rows runs on the device
    # (quotient evaluation: mp2gpu_quotient_polys; openings: mp2gpu_batch_eval; query rounds: mp2gpu_batch_open) --
    # host columns in; coefficients, digests and cap out; the rows stay in HBM behind the batch handle.  The variant
    # that also ships every row to the host (what round 1 reported as e2e) is kept beside it.
    out = {"value": elems / dt_res / 1e9, "unit": "Gelem/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h_res,
           "ms_per_step": dt_res * 1e3, "steps": steps,
           "api": api + "; leaves_out = NULL + handle_out (rows stay device-resident)",
           "outputs": "coefficients + digests + cap copied back to the host every step; LDE rows stay in HBM behind "
                      "mp2gpu_batch_fetch_rows / _open / _eval / mp2gpu_quotient_polys",
           "cap_only_to_host": {"value": elems / dt_cap / 1e9, "unit": "Gelem/s", "ms_per_step": dt_cap * 1e3,
                                "d2h_bytes_per_step": 32 * ncap,
                                "note": "the call prover.prove() makes: host columns in, the cap out; coefficients, rows and "
                                        "digests stay in HBM behind the handle (openings / quotient / query rounds read them there)"},
           "all_outputs_to_host": {"value": elems / dt / 1e9, "unit": "Gelem/s", "ms_per_step": dt * 1e3,
                                   "d2h_bytes_per_step": d2h,
                                   "note": "the same call with leaves_out set: all %0.1f GB of row-major leaves cross PCIe "
                                           "as well (round 1's e2e definition)" % (8 * a.ncols * (N // world) / 1e9)}}
    return out


# ------------------------------------------------------------------------------------------------
# proof-trace replay (map stage): independent proofs, one GPU each, no collective
# ------------------------------------------------------------------------------------------------
def measured_traffic(a, world):
    """DRAM bytes per launch for the shapes the committed ncu captures cover (profiles/traffic.json).  These are
    STATIC figures -- taken from an `ncu --set full` capture, not from this run (bench.py never runs under a
    profiler) -- so the source string carries the capture file and the commit of the kernels it profiled."""
    if world != 1:
        return {}
    key = "2^%dx%d r%d %s" % (a.n_log, a.ncols, a.rate_bits, "poseidon2" if a.hash in ("poseidon2", 1) else "poseidon")
    try:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "traffic.json")) as f:
            t = json.load(f).get(key, {})
    except (OSError, ValueError):
        return {}
    if t:
        t = dict(t)
        t["source"] = "static: ncu capture %s (kernels at commit %s), not measured in this run" % (
            t.get("source"), t.get("commit", "unknown"))
    return t


def cpu_trace_time(kind, threads, degrees):
    """One proof trace on the CPU restatement (oracle/)."""
    import oracle as O
    from mapreduce_plonky2_b200 import trace as T

    O.build()
    t0 = time.perf_counter()
    for i, op in enumerate(T.proof_ops(degrees)):
        if op.kind == "merkle":
            leaves = synthetic_columns(0x7000 + i, (1 << op.n_log, op.ncols))
            O.merkle_new(leaves, min(T.CAP_HEIGHT, op.n_log), kind, nthreads=threads)
        else:
            cols = synthetic_columns(0x7000 + i, (op.ncols, 1 << op.n_log))
            O.commit(cols, T.RATE_BITS, T.CAP_HEIGHT, kind, op.kind == "from_coeffs", nthreads=threads, want_leaves=True)
    return time.perf_counter() - t0


def run_trace(a):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    kind = 0 if a.hash == "poseidon" else 1
    from mapreduce_plonky2_b200 import trace as T0

    degrees = T0.LEAF_PROOF_DEGREES if a.trace == "leaf" else T0.AGGREGATION_DEGREES
    name = ("map stage: %d independent %s commitment traces per GPU (prove() degrees %s ASSUMED; %s; trace replay = "
            "upper bound on proofs/s)") % (a.proofs, "mp2-v1 leaf-proof" if a.trace == "leaf" else
                                           "2-proof universal-verifier aggregation",
                                           ", ".join("2^%d" % d for d in degrees), a.hash)
    if a.impl == "reference":
        if rank == 0:
            threads = host_threads()
            dt = [cpu_trace_time(kind, threads, degrees) for _ in range(max(1, min(a.steps, 3)))]
            v = len(dt) / sum(dt)
            print(json.dumps({"impl": "reference", "metric": "mp2 proofs/s (commitment trace)", "value": v,
                              "unit": "proofs/s", "n_gpus": a.gpus, "steps": len(dt), "warmup": 0,
                              "ms_per_step": 1e3 * sum(dt) / len(dt), "higher_is_better": True, "scaling": "weak",
                              "vs_baseline": None, "dtype": "u64 (Goldilocks field)", "data": "synthetic",
                              "config": {"workload": name, "sample": "one proof trace per step"},
                              "cpu_baseline": {"value": v, "unit": "proofs/s", "cores": threads, "kind": "port",
                                               "sample": "one proof trace per step"},
                              "e2e": {"value": v, "unit": "proofs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                              "gpu_launches": 0}), flush=True)
        return
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import mapreduce_plonky2_b200 as G
    from mapreduce_plonky2_b200 import trace as T

    G.init(local_rank)
    runner = T.TraceRunner(degrees, kind, a.streams)
    launches0 = G.launch_count()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    for _ in range(max(a.warmup, 3)):
        runner.run(a.streams)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = G.launch_count()
    e0.record()
    for _ in range(a.steps):
        runner.run(a.proofs)
        for st in runner.streams:  # the default stream (and e1) waits for every in-flight proof
            torch.cuda.current_stream().wait_stream(st)
    e1.record()
    barrier()
    clocks = sampler.stop() if sampler else None
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_step = float(ms.item()) / a.steps
    value = world * a.proofs / (ms_step * 1e-3)
    if rank == 0:
        perms = runner.perms_per_proof
        ip = __import__("mapreduce_plonky2_b200.device", fromlist=["x"]).int_pipe_peak()
        mads = perms * a.proofs * PERM_MADS / (ms_step * 1e-3) / 1e12
        line = {"metric": "mp2 proofs/s (commitment trace)", "value": value, "unit": "proofs/s", "n_gpus": world,
                "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u64 (Goldilocks field, integer)", "data": "synthetic",
                "config": {"workload": name, "parallelism": "replicas only (one proof stream set per GPU, no collective)",
                           "streams_per_gpu": a.streams, "l2": "working set of %d concurrent proofs exceeds L2" % a.streams,
                           "ops_per_proof": len(runner.ops), "perms_per_proof": perms,
                           "lde_elems_per_proof": runner.lde_elems_per_proof},
                "gpu_launches": int(G.launch_count() - launches0), "clocks": clocks,
                "lde_gelems_per_s": world * a.proofs * runner.lde_elems_per_proof / (ms_step * 1e-3) / 1e9,
                "roofline": {"kernel": "all Poseidon kernels of the trace (whole step)", "bound": "int_pipe",
                             "achieved": mads, "peak": ip["t_imad_per_s"], "unit": "T imad/s (6700 credited per permutation)",
                             "frac": mads / ip["t_imad_per_s"] if ip["t_imad_per_s"] else None, "traffic": None}}
        if not a.no_cpu_baseline and world == 1:
            threads = host_threads()
            dt = cpu_trace_time(kind, threads, degrees)
            line["cpu_baseline"] = {"value": 1.0 / dt, "unit": "proofs/s", "cores": threads, "kind": "port",
                                    "sample": "one proof trace (%.1f s)" % dt}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    a = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    if a.workload == "trace":
        run_trace(a)
        return
    if a.impl == "reference":
        run_reference(a, rank)
        return
    run_ours(a)


if __name__ == "__main__":
    main()

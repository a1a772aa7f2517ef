"""NCCL run of the sharded commitment on real GPUs (needs >= 2 devices; skipped otherwise):
column shards -> all-to-all -> row shards must reassemble to the oracle's single PolynomialBatch."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ncols, n_log, rate_bits, cap_height, kind, out_dir, exchange, from_coeffs=False):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from util import field_elems

    from mapreduce_plonky2_b200.sharded import CudaEngine, commit_sharded

    cols = field_elems(0xBEEF, (ncols, 1 << n_log))
    c_loc = ncols // world
    mine = torch.from_numpy(cols[rank * c_loc:(rank + 1) * c_loc].view(np.int64).copy()).cuda()
    res = commit_sharded(mine, ncols, rate_bits, cap_height, kind, CudaEngine(), exchange=exchange,
                         from_coeffs=from_coeffs)
    torch.cuda.synchronize()
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), coeffs=res.coeffs.cpu().numpy().view(np.uint64),
             leaves=res.leaves.cpu().numpy().view(np.uint64), digests=res.digests.cpu().numpy().view(np.uint64),
             cap=res.cap.cpu().numpy().view(np.uint64))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("exchange", ["nccl", "peer"])
@pytest.mark.parametrize("ncols,n_log,kind", [(16, 12, 0), (6, 15, 1)])
def test_sharded_nccl_equals_oracle(tmp_path, oracle, ncols, n_log, kind, exchange):
    import torch
    import torch.multiprocessing as mp

    world = 2
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    from util import field_elems

    port = 29600 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, ncols, n_log, 3, 4, kind, str(tmp_path), exchange), nprocs=world, join=True)
    cols = field_elems(0xBEEF, (ncols, 1 << n_log))
    ref = oracle.commit(cols, 3, 4, kind)
    parts = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % r)) for r in range(world)]
    assert np.array_equal(np.concatenate([p["coeffs"] for p in parts]), ref["coeffs"])
    assert np.array_equal(np.concatenate([p["leaves"] for p in parts]), ref["leaves"])
    assert np.array_equal(np.concatenate([p["digests"] for p in parts]), ref["digests"])
    for p in parts:
        assert np.array_equal(p["cap"], ref["cap"])


def test_sharded_from_coeffs_nccl(tmp_path, oracle):
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from util import field_elems

    port = 29700 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, 8, 10, 3, 4, 0, str(tmp_path), "nccl", True), nprocs=2, join=True)
    ref = oracle.commit(field_elems(0xBEEF, (8, 1 << 10)), 3, 4, 0, True)
    parts = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % r)) for r in range(2)]
    assert np.array_equal(np.concatenate([p["coeffs"] for p in parts]), ref["coeffs"])
    assert np.array_equal(np.concatenate([p["leaves"] for p in parts]), ref["leaves"])
    assert np.array_equal(parts[0]["cap"], ref["cap"])


# ---- the same path behind the C ABI: one process, one host thread per device, peer stores, no NCCL -------------
def _build_cpp(tmp_path, name):
    import subprocess

    exe = os.path.join(str(tmp_path), name)
    env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
    pkg = os.path.join(ROOT, "mapreduce_plonky2_b200")
    subprocess.run(["g++", "-std=c++17", "-O2", "-o", exe, os.path.join(ROOT, "tests", "cpp", name + ".cpp"),
                    "-L" + pkg, "-lmp2gpu", "-L" + os.path.join(ROOT, "oracle"), "-lmp2oracle",
                    "-Wl,-rpath," + pkg, "-Wl,-rpath," + os.path.join(ROOT, "oracle"), "-fopenmp"], check=True, env=env)
    return exe


@pytest.mark.parametrize("ndev", [1, 2])
def test_cpp_sharded_c_abi_equals_oracle(tmp_path, oracle, ndev):
    """tests/cpp/test_sharded_abi.cpp: mp2gpu_comm_init + mp2gpu_commit_from_values_sharded vs the oracle."""
    import subprocess

    import torch

    if torch.cuda.device_count() < ndev:
        pytest.skip("needs %d GPUs" % ndev)
    exe = _build_cpp(tmp_path, "test_sharded_abi")
    r = subprocess.run([exe, str(ndev)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "sharded C ABI OK on %d device(s)" % ndev in r.stdout


@pytest.mark.parametrize("ndev", [1, 2])
def test_python_mirror_from_values_sharded(oracle, ndev):
    import torch

    import mapreduce_plonky2_b200 as G
    from util import field_elems

    if torch.cuda.device_count() < ndev:
        pytest.skip("needs %d GPUs" % ndev)
    comm = G.Communicator(list(range(ndev)))
    cols = field_elems(0xC0FFEE, (12, 1 << 11))
    for kind in (G.POSEIDON, G.POSEIDON2):
        pb = G.PolynomialBatch.from_values_sharded(comm, cols, 3, False, 4, hash_kind=kind)
        ref = oracle.commit(cols, 3, 4, kind)
        assert np.array_equal(pb.polynomials, ref["coeffs"])
        assert np.array_equal(pb.merkle_tree.leaves, ref["leaves"])
        assert np.array_equal(pb.merkle_tree.digests, ref["digests"])
        assert np.array_equal(pb.merkle_tree.cap.hashes, ref["cap"])
    with pytest.raises(G.Mp2GpuError):
        G.PolynomialBatch.from_values_sharded(comm, cols, 3, True, 4)
    comm.free()

"""Scratch timing of the device-resident commitment (not the contract bench; see bench.py)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mapreduce_plonky2_b200 import device as D

def run(ncols, n_log, rate_bits=3, cap=4, kind=0, from_coeffs=False, iters=5, want_leaves=True):
    torch.cuda.set_device(0); D.bind_current_device()
    n = 1 << n_log
    cols = torch.randint(0, 2**62, (ncols, n), dtype=torch.int64, device="cuda")
    bufs = D.CommitBuffers(ncols, n_log, rate_bits, cap, want_leaves)
    t0=time.time()
    while time.time()-t0 < 0.5: D.commit_resident(cols, bufs, kind, from_coeffs); torch.cuda.synchronize()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ts = []
    for _ in range(iters):
        ev[0].record(); D.intt(cols, bufs.coeffs)
        ev[1].record(); D.coset_lde(bufs.coeffs, bufs.lde, rate_bits)
        ev[2].record(); D.merkle_colmajor(bufs.lde, cap, kind, bufs.leaves, bufs.digests, bufs.cap)
        ev[3].record(); torch.cuda.synchronize()
        ts.append([ev[i].elapsed_time(ev[i+1]) for i in range(3)])
    best = min(ts, key=sum)
    elems = ncols * (n << rate_bits)
    print("c=%d n=2^%d kind=%d: intt %.3f ms, lde %.3f ms, merkle %.3f ms, total %.3f ms -> %.2f Gelem/s" % (
        ncols, n_log, kind, best[0], best[1], best[2], sum(best), elems / sum(best) / 1e6), flush=True)

def field_checks():
    import ctypes as C
    from mapreduce_plonky2_b200 import _lib
    torch.cuda.set_device(0); D.bind_current_device()
    bad = (C.c_uint64 * 12)()
    _lib.call("mp2gpu_debug_field_selftest", bad, 12)
    print("field selftest mismatches:", list(bad), flush=True)
    out = (C.c_double * 2)()
    _lib.call("mp2gpu_debug_field_probe", out)
    print("field probe: %.3f x^7/clk/SM, %.3f dft8-elements/clk/SM" % (out[0], out[1]), flush=True)


if __name__ == "__main__":
    print("lib:", os.environ.get("MP2GPU_LIB", "default"), flush=True)
    try:
        field_checks()
    except Exception as e:  # older library variants have no self-test
        print("field checks unavailable:", e, flush=True)
    for kind in (0, 1):
        run(135, 14, kind=kind)
        run(20, 14, kind=kind)
        run(135, 12, kind=kind)
    run(256, 20, kind=0, iters=2)
    run(256, 20, kind=1, iters=2)

"""Dry run of bench.py WITHOUT a GPU (test infrastructure, not a measurement).

bench.py's GPU arm is executed end to end on the `tiny` workload against the CPU build of the
library (tests/cpu_emul/build_emul_lib.py, passed as LM_EMUL_LIB) with torch.cuda replaced by
host stand-ins (a stream handle, wall-clock events, no-op synchronize / pin_memory).  It checks
what a CPU box can check about the bench harness: every Python path of the timed `value` and
`e2e` legs runs, the C-ABI calls it makes exist and succeed, and the JSON line carries every key
of the contract.  The numbers it prints are meaningless (kernels run as fibers on the host).

    LM_EMUL_LIB=<liblm_b200_emul.so> python tools/bench_dryrun_cpu.py
"""
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

REQUIRED = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline", "parity_check"]


def main():
    assert os.environ.get("LM_EMUL_LIB"), "LM_EMUL_LIB must name the CPU build of the library"
    import conftest
    assert conftest._emulated_library()
    import torch

    keep = C.create_string_buffer(64)          # something non-NULL for the stream handle to point at

    class Stream:
        cuda_stream = C.addressof(keep)

    class Event:
        def __init__(self, enable_timing=False):
            self.t = None

        def record(self):
            self.t = time.perf_counter()

        def elapsed_time(self, other):
            return 1e3 * (other.t - self.t)

    torch.cuda.set_device = lambda d: None
    torch.cuda.Stream = Stream
    torch.cuda.set_stream = lambda s: None
    torch.cuda.synchronize = lambda *a: None
    torch.cuda.Event = Event
    torch.Tensor.pin_memory = lambda self: self

    import bench
    sys.argv = ["bench.py", "--workload", "tiny", "--steps", "3", "--warmup", "3", "--cpu-cols", "16"]
    lines = []
    bench.emit = lambda obj: lines.append(json.dumps(obj))      # bench.py writes its line to a private dup of fd 1
    bench.main()
    assert len(lines) == 1, "bench.py must emit exactly one JSON line, got %d" % len(lines)
    out = json.loads(lines[0])
    missing = [k for k in REQUIRED if k not in out]
    assert not missing, "bench line lacks %s" % missing
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in out["e2e"], k
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in out["roofline"], k
    for k in ("value", "unit", "cores", "kind", "sample"):
        assert k in out["cpu_baseline"], k
    assert out["gpu_launches"] > 0 and out["value"] > 0 and out["e2e"]["value"] > 0
    assert out["parity_check"]["max_rel"] < 1e-12, out["parity_check"]
    assert out["roofline"]["traffic"] is None
    assert out["e2e"]["h2d_bytes_per_step"] > 0 and out["e2e"]["d2h_bytes_per_step"] > 0
    sys.stderr.write("bench dry run ok: %d launches in the timed region, keys %s" % (out["gpu_launches"], sorted(out)))
    sys.stderr.flush()


if __name__ == "__main__":
    main()

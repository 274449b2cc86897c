"""Where does the end-to-end step lose time over PCIe?  `e2e` (host batch -> H2D -> step -> loss read-back) is bound by
the 1.77 GB fp32 volume batch at the ~39 GB/s this pool's hosts reached in round 1.  This probe
measures, on one GPU, what the host path can deliver under the options a data loader controls:

  * pinned memory as torch allocates it (cudaHostAlloc default) - one copy, and the same bytes split over 2 / 4 streams
  * write-combined pinned memory (cudaHostAllocWriteCombined): not snooped, often faster to read over PCIe
  * the process pinned to each NUMA node's CPUs before the buffer is allocated and first touched
  * zero-copy: a kernel reading the pinned buffer directly (ctk_copy_from_pinned), no copy engine involved
  * the same bytes as float16 (what `vit_exp_b200.data` ships for the *_fp16 datasets)

    python tools/h2d_probe.py > gpurun_out/h2d_probe.jsonl        (about 30 s)

One JSON line per variant: GB/s and the milliseconds a 1.77 GB batch would take.
"""
import ctypes
import glob
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

BYTES = 8 * 240 * 480 * 480 * 4          # one 8-volume fp32 batch


def timed_copy(fn, iters=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def report(name, ms, nbytes=BYTES, **kw):
    print(json.dumps({"variant": name, "GBps": nbytes / ms / 1e6, "ms_per_copy": ms,
                      "ms_for_fp32_batch": BYTES / (nbytes / ms), **kw}), flush=True)


def main():
    dev = torch.device("cuda:0")
    torch.cuda.set_device(0)
    dst = torch.empty(BYTES, dtype=torch.uint8, device=dev)
    try:
        q = subprocess.run(["nvidia-smi", "--query-gpu=pcie.link.gen.current,pcie.link.gen.max,pcie.link.width.current",
                            "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
    except Exception as e:                                  # noqa: BLE001
        q = repr(e)
    nodes = sorted(glob.glob("/sys/devices/system/node/node[0-9]*"))
    print(json.dumps({"pcie": q, "numa_nodes": len(nodes), "cpus": os.cpu_count(),
                      "affinity": len(os.sched_getaffinity(0))}), flush=True)

    # 1. torch pinned memory, one copy / split over streams
    src = torch.empty(BYTES, dtype=torch.uint8).pin_memory()
    src.fill_(1)
    report("pinned (torch), 1 copy", timed_copy(lambda: dst.copy_(src, non_blocking=True)))
    for k in (2, 4):
        streams = [torch.cuda.Stream() for _ in range(k)]
        step = BYTES // k

        def split():
            cur = torch.cuda.current_stream()
            for i, st in enumerate(streams):
                st.wait_stream(cur)
                with torch.cuda.stream(st):
                    dst[i * step:(i + 1) * step].copy_(src[i * step:(i + 1) * step], non_blocking=True)
            for st in streams:
                cur.wait_stream(st)
        report(f"pinned (torch), split over {k} streams", timed_copy(split))

    # 2. zero-copy kernel read of the same pinned buffer
    try:
        from vit_exp_b200 import _lib
        lib = _lib.load()
        s = torch.cuda.current_stream().cuda_stream
        report("zero-copy kernel read (ctk_copy_from_pinned)",
               timed_copy(lambda: _lib.check(lib.ctk_copy_from_pinned(dst.data_ptr(), src.data_ptr(), BYTES, s))))
    except Exception as e:                                  # noqa: BLE001
        print(json.dumps({"variant": "zero-copy kernel read", "error": repr(e)}), flush=True)

    # 3. half the bytes (float16 stored volumes)
    report("pinned (torch), float16 batch", timed_copy(lambda: dst[: BYTES // 2].copy_(src[: BYTES // 2], non_blocking=True)),
           nbytes=BYTES // 2)
    del src

    # 4. write-combined pinned memory
    try:
        from cuda.bindings import runtime as rt
        err, ptr = rt.cudaHostAlloc(BYTES, rt.cudaHostAllocWriteCombined)
        assert int(err) == 0, err
        ctypes.memset(ptr, 1, BYTES)
        st = torch.cuda.current_stream().cuda_stream

        def wc():
            (e,) = rt.cudaMemcpyAsync(dst.data_ptr(), ptr, BYTES, rt.cudaMemcpyKind.cudaMemcpyHostToDevice, st)
            assert int(e) == 0, e
        report("pinned write-combined (cudaHostAllocWriteCombined)", timed_copy(wc))
        rt.cudaFreeHost(ptr)
    except Exception as e:                                  # noqa: BLE001
        print(json.dumps({"variant": "pinned write-combined", "error": repr(e)}), flush=True)

    # 5. per NUMA node: bind the process, allocate + first-touch there, copy
    all_cpus = os.sched_getaffinity(0)
    for nd in nodes:
        try:
            cpus = set()
            for part in open(os.path.join(nd, "cpulist")).read().strip().split(","):
                a, _, b = part.partition("-")
                cpus |= set(range(int(a), int(b or a) + 1))
            cpus &= all_cpus
            if not cpus:
                continue
            os.sched_setaffinity(0, cpus)
            buf = torch.empty(BYTES, dtype=torch.uint8)
            buf.fill_(1)                                       # first touch on this node
            buf = buf.pin_memory()
            report(f"pinned (torch), process bound to {os.path.basename(nd)}", timed_copy(lambda: dst.copy_(buf, non_blocking=True)),
                   cpus=len(cpus))
            del buf
        except Exception as e:                              # noqa: BLE001
            print(json.dumps({"variant": f"numa {nd}", "error": repr(e)}), flush=True)
        finally:
            os.sched_setaffinity(0, all_cpus)


if __name__ == "__main__":
    main()

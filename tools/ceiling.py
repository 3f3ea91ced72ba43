#!/usr/bin/env python
"""Where is the ceiling?  Same 1 GiB buffers: torch copy_ (the MEASURED_PEAKS method), cudaMemcpyAsync D2D,
our TMA ring without digest (debug flag), with digest, and the read-only scan."""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from lzy_b200 import _native as N  # noqa: E402


def timeit(fn, stream, reps=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    N.init(0, 1, 16 << 20)
    n = 1 << 30
    stream = torch.cuda.Stream(0)
    torch.cuda.set_stream(stream)
    sp = int(stream.cuda_stream)
    a = torch.empty(n, dtype=torch.uint8, device="cuda:0")
    b = torch.empty(n, dtype=torch.uint8, device="cuda:0")
    N.fill_pattern(0, a.data_ptr(), n, 1)
    out = {}
    out["torch_copy_"] = 2 * n / timeit(lambda: b.copy_(a), stream) / 1e6
    a16, b16 = a.view(torch.bfloat16), b.view(torch.bfloat16)
    out["torch_copy_bf16"] = 2 * n / timeit(lambda: b16.copy_(a16), stream) / 1e6
    N.set_copy_engine(N.ENGINE_TMA)
    for cfg in (0, 1, 2):
        N.set_tma_config(cfg)
        for flags, tag in ((1, "nodigest"), (0, "digest")):
            N.lib.lzyb_set_debug_flags(flags)

            def pack():
                s = N.SlotHandle.create_output("c/o", 0, n)
                s.pack(a.data_ptr(), n, sp)
                s.commit(sp)
                s.close()
            out[f"tma_cfg{cfg}_{tag}"] = 2 * n / timeit(pack, stream) / 1e6
    N.lib.lzyb_set_debug_flags(0)
    N.set_copy_engine(N.ENGINE_LDG)
    for cps in (4, 0, -2, -4, -8):
        N.set_grid(cps)

        def pack_l():
            s = N.SlotHandle.create_output("c/o", 0, n)
            s.pack(a.data_ptr(), n, sp)
            s.commit(sp)
            s.close()
        out[f"ldg_cps{cps}_digest_copy"] = 2 * n / timeit(pack_l, stream) / 1e6
        out[f"ldg_cps{cps}_scan"] = n / timeit(lambda: N.digest_device(0, a.data_ptr(), n, sp), stream) / 1e6
    for mib in (64, 256):
        m = mib << 20
        for cps in (4, 0, -2):
            N.set_grid(cps)

            def pack_m():
                s = N.SlotHandle.create_output("c/m", 0, m)
                s.pack(a.data_ptr(), m, sp)
                s.commit(sp)
                s.close()
            out[f"ldg_cps{cps}_copy_{mib}MiB"] = 2 * m / timeit(pack_m, stream) / 1e6
        N.set_copy_engine(N.ENGINE_TMA)
        out[f"tma_copy_{mib}MiB"] = 2 * m / timeit(pack_m, stream) / 1e6
        N.set_copy_engine(N.ENGINE_LDG)
    N.set_grid(-1000)
    N.set_copy_engine(N.ENGINE_AUTO)
    out["auto_scan"] = n / timeit(lambda: N.digest_device(0, a.data_ptr(), n, sp), stream) / 1e6
    N.set_copy_engine(N.ENGINE_TMA)
    out["scan_only_read"] = n / timeit(lambda: N.digest_device(0, a.data_ptr(), n, sp), stream) / 1e6
    N.set_copy_engine(N.ENGINE_LDG)
    out["scan_only_read_ldg"] = n / timeit(lambda: N.digest_device(0, a.data_ptr(), n, sp), stream) / 1e6
    print(json.dumps({k: round(v, 1) for k, v in out.items()}))


if __name__ == "__main__":
    main()

"""Producer process of tests/test_multiprocess_gpu.py::test_producer_process_pushes_*: owns GPU `gpu`, commits a slot
holding the counter-based pattern (seed) and pushes it into the sink frame the consumer process described.
Protocol on stdin/stdout (hex lines): → our slot descriptor; ← the consumer's sink descriptor; → "pushed"."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))


def main():
    gpu, nbytes, seed = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    import torch

    from lzy_b200 import _native as N
    N.init(gpu, 1, 16 << 20)
    torch.cuda.set_device(gpu)
    t = torch.empty(nbytes, dtype=torch.uint8, device=f"cuda:{gpu}")
    N.fill_pattern(gpu, t.data_ptr(), nbytes, seed)
    torch.cuda.synchronize()
    out = N.SlotHandle.create_output("ipcpush/out", gpu, nbytes)
    out.set_meta(N.DT_U8, (nbytes,))
    out.pack(t.data_ptr(), nbytes)
    out.commit()
    print(bytes(out.export()).hex(), flush=True)
    for line in sys.stdin:
        line = line.strip()
        if line == "quit":
            break
        sink = N.IpcDesc.from_buffer_copy(bytes.fromhex(line))
        out.push(sink)  # blocking on the library stream: returns when every store has been issued and retired
        print("pushed", N.stats()["push_launches"], N.stats()["dma_copies"], flush=True)
    out.close()


if __name__ == "__main__":
    main()

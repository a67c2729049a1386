"""Buffer plumbing for the parity tests: the same checks run against the CUDA library (torch tensors on
cuda:0) and, on CPU-only machines, against the test-only emulation build (numpy arrays)."""
import numpy as np


class EmuBackend:
    name = "emu"

    def __init__(self, lib):
        self.lib = lib

    def to_dev(self, arr):
        return np.ascontiguousarray(arr, dtype=np.uint64).copy()

    def empty(self, shape, dtype=np.uint64):
        return np.zeros(shape, dtype=dtype)

    def to_host(self, x):
        return np.array(x, copy=True)

    def sync(self):
        pass


class CudaBackend:
    name = "cuda"

    def __init__(self):
        import torch
        from sunscreen_b200.lib import B200Lib
        self.torch = torch
        self.lib = B200Lib.default()
        assert torch.cuda.is_available()

    def to_dev(self, arr):
        a = np.ascontiguousarray(arr, dtype=np.uint64)
        return self.torch.from_numpy(a.view(np.int64)).to("cuda:0")

    def empty(self, shape, dtype=np.uint64):
        td = self.torch.int64 if dtype == np.uint64 else self.torch.int32
        return self.torch.zeros(shape, dtype=td, device="cuda:0")

    def to_host(self, x):
        self.torch.cuda.synchronize()
        a = x.cpu().numpy()
        return a.view(np.uint64) if a.dtype == np.int64 else a.view(np.uint32)

    def sync(self):
        self.torch.cuda.synchronize()

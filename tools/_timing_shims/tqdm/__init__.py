"""Timing stand-in for tqdm used ONLY by tools/bench_routes.py (placed in front of the real tqdm on PYTHONPATH): the training
scripts iterate `tqdm(range(...))`; this wrapper lets an unmodified script report its own steady-state rate -- the clock starts after
TNERF_TIMING_WARMUP iterations (device synchronised) and stops when the loop ends (device synchronised)."""
import os
import time


class tqdm:
    def __init__(self, iterable=None, **kwargs):
        self.iterable = iterable

    def __iter__(self):
        import torch
        warm = int(os.environ.get("TNERF_TIMING_WARMUP", "200"))
        t0, n = None, 0
        for i, x in enumerate(self.iterable):
            if i == warm:
                if torch.cuda.is_available():
                    torch.cuda.synchronize()
                t0 = time.perf_counter()
            yield x
            n = i + 1
        if t0 is not None and n > warm:
            if torch.cuda.is_available():
                torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            print(f"[timing] steps={n - warm} seconds={dt:.6f}", flush=True)

    def set_postfix(self, **kwargs):
        pass

    def update(self, n=1):
        pass

    def close(self):
        pass

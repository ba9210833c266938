"""Output path of the CLI, off the optimisation loop's critical path.

The reference saves synchronously from inside the per-iteration callback (/root/reference/style_transfer/cli.py:125-133):
at 2048^2 the PNG encode alone stalls the loop for hundreds of iterations' worth of device time.  Here the callback only
takes a device-side uint8 snapshot of the averaged iterate and starts its copy into pinned host memory; encoding and the
file write happen on a worker thread, newest snapshot wins.
"""
from __future__ import annotations

import os
import queue
import threading
from pathlib import Path

import numpy as np
from PIL import Image


class AsyncImageWriter:
    def __init__(self):
        self._q: queue.Queue = queue.Queue()
        self._errors: list[BaseException] = []
        self._thread = threading.Thread(target=self._run, name='stb-image-writer', daemon=True)
        self._thread.start()

    # ------------------------------------------------------------------ producer side (the stylize callback)
    def submit_snapshot(self, st, path):
        """Snapshot `st`'s current averaged image on the device and queue it for saving to `path`."""
        import torch
        t = st.get_image_tensor()                                   # [3,H,W] fp32 on the device, clamped
        u8 = t.mul(255).byte().permute(1, 2, 0).contiguous()        # to_pil_image semantics, HWC
        if u8.is_cuda:
            host = torch.empty(u8.shape, dtype=torch.uint8, pin_memory=True)
            host.copy_(u8, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
        else:
            host, ev = u8, None
        self._q.put((host, ev, Path(path)))

    def submit_array(self, array: np.ndarray, path):
        self._q.put((array, None, Path(path)))

    # ------------------------------------------------------------------ worker
    def _run(self):
        while True:
            item = self._q.get()
            if item is None:
                return
            while True:  # newest snapshot for the same path wins; never fall behind the loop
                try:
                    nxt = self._q.get_nowait()
                except queue.Empty:
                    break
                if nxt is None:
                    self._save(item)
                    return
                if nxt[2] != item[2]:
                    self._save(item)
                item = nxt
            self._save(item)

    def _save(self, item):
        host, ev, path = item
        try:
            if ev is not None:
                ev.synchronize()
            arr = host.numpy() if hasattr(host, 'numpy') else np.asarray(host)
            tmp = path.with_name(path.stem + '.part' + path.suffix)
            Image.fromarray(arr).save(tmp)
            os.replace(tmp, path)                                   # readers never see a half-written file
        except BaseException as err:  # noqa: BLE001 - reported by close()
            self._errors.append(err)

    def close(self):
        """Flush pending saves; re-raises the first error of the worker, if any."""
        self._q.put(None)
        self._thread.join()
        if self._errors:
            raise self._errors[0]

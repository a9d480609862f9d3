"""ORACLE INFRASTRUCTURE -- imports the staged reference (oracle/_ref/, produced by oracle/make_ref.py) and runs its own
training step on the CPU. Test / benchmark infrastructure only: nothing under camvid_b200 imports this.

Shims (applied here, never to the staged files): `np.float = float` for utils.mean_iou (utils.py:210-213
uses the alias numpy removed), `legacy/` on sys.path for `from metrics import Metrics`.
"""
import hashlib
import importlib
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
_MODULES = ("utils", "models", "models.unet", "models.segnet", "metrics", "transforms", "conf", "conf.settings")


def available():
    """True when every staged file is present and matches the digest recorded when it was copied."""
    try:
        man = json.load(open(os.path.join(REF_DIR, "MANIFEST.json")))
        for rel, digest in man["files"].items():
            if hashlib.sha256(open(os.path.join(REF_DIR, rel), "rb").read()).hexdigest() != digest:
                return False
        return True
    except Exception:
        return False


class reference_modules:
    """Context manager: the reference's `utils`, `models.*`, `metrics` (legacy) and `transforms` importable by those
    names, removed from sys.modules / sys.path again on exit (the names are generic; the product has its own)."""

    def __enter__(self):
        import numpy as np
        if not hasattr(np, "float"):
            np.float = float
        self.saved = {m: sys.modules.pop(m) for m in _MODULES if m in sys.modules}
        self.paths = [REF_DIR, os.path.join(REF_DIR, "legacy")]
        sys.path[:0] = self.paths
        self.dont_write = sys.dont_write_bytecode
        sys.dont_write_bytecode = True
        ns = type("ref", (), {})()
        ns.utils = importlib.import_module("utils")
        ns.metrics = importlib.import_module("metrics")
        ns.transforms = importlib.import_module("transforms")
        ns.settings = importlib.import_module("conf").settings
        return ns

    def __exit__(self, *exc):
        for p in self.paths:
            if p in sys.path:
                sys.path.remove(p)
        for m in _MODULES:
            sys.modules.pop(m, None)
        sys.modules.update(self.saved)
        sys.dont_write_bytecode = self.dont_write
        return False


def train_steps(model, batch, h, w, steps, warmup, threads=None, seed=0):
    """The loop body of train.py:124-134 with the reference's own objects on the host cores:
    `utils.get_model(model, 3, 12)` (utils.py:147-160), `nn.CrossEntropyLoss()` (train.py:105),
    `optim.AdamW(lr=5e-4 default of train.py:25, weight_decay=0)` (train.py:100); synthetic batch as BASELINE.md
    section 5. Returns (images/s, seconds per step, threads, last loss)."""
    import torch
    from . import camvid_oracle as O
    if threads:
        torch.set_num_threads(threads)
    with reference_modules() as ref:
        torch.manual_seed(seed)
        net = ref.utils.get_model(model, 3, 12)
        net.train()
        optimizer = torch.optim.AdamW(net.parameters(), lr=5e-4, weight_decay=0)
        loss_fn = torch.nn.CrossEntropyLoss()
        images, masks = O.synth_batch(batch, h, w, seed=seed)
        times, loss = [], None
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            optimizer.zero_grad()
            preds = net(images)
            loss = loss_fn(preds, masks)
            loss.backward()
            optimizer.step()
            last = loss.item()
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return batch / sec, sec, torch.get_num_threads(), last

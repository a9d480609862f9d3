"""ORACLE INFRASTRUCTURE -- stages the UNMODIFIED reference implementation of the hot path under oracle/_ref/.

    python oracle/make_ref.py            # run where /root/reference is mounted (the build container)

The reference (weiaicunzai/pytorch-camvid) is pure Python over torch: "building" it is placing its own source files
where they can be imported. This recipe copies, byte for byte, only the files the hot path consists of
 from /root/reference into oracle/_ref/ and writes a MANIFEST.json with their SHA-256 digests:

    models/unet.py  models/segnet.py  utils.py  legacy/metrics.py  transforms.py  conf/__init__.py  conf/settings.py

oracle/_ref/ is git-ignored (reference sources never enter this repository's history) but NOT gpurun-ignored, so it
travels to the GPU box with the snapshot like the built .so files. There it serves as
  * the checker the oracle restatement is validated against (tests/test_oracle.py), and
  * the CPU baseline of `bench.py --impl reference` (cpu_baseline.kind = "reference"): the reference's own
    `utils.get_model` modules, `nn.CrossEntropyLoss` and `optim.AdamW`, i.e. the loop body of train.py:124-134.
Only tests/, __graft_entry__ and bench.py's CPU legs may import it (through oracle/ref_runner.py); the product never does.
"""
import hashlib
import json
import os
import shutil
import sys

REF = os.environ.get("CAMVID_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
FILES = ["models/unet.py", "models/segnet.py", "utils.py", "legacy/metrics.py", "transforms.py", "conf/__init__.py",
         "conf/settings.py"]


def main():
    if not os.path.isdir(REF):
        print(f"make_ref: {REF} is not present; keeping whatever oracle/_ref/ already holds", file=sys.stderr)
        return 1
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(REF, rel), os.path.join(OUT, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
    json.dump({"source": REF, "files": manifest}, open(os.path.join(OUT, "MANIFEST.json"), "w"), indent=1)
    print(f"make_ref: staged {len(FILES)} reference files under {OUT}")
    return 0


if __name__ == "__main__":
    sys.exit(main())

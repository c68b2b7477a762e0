#!/usr/bin/env python3
"""Recipe for oracle/_ref: a verbatim, UNMODIFIED copy of the reference's Python sources.

TEST / MEASUREMENT INFRASTRUCTURE ONLY (never imported by vpd_b200/).

    python -m oracle.build_ref

The reference (jhong93/vpd) is a directory of Python scripts: no setup.py, no
pyproject.toml, nothing to compile - "building" it means making its modules importable
where /root/reference does not exist (the GPU box). This recipe copies every `*.py`
(364 KB, 49 files; no data, no pickles) from where it lies under /root/reference into
`oracle/_ref/`, byte for byte, and writes a manifest with the sha256 of every file so a
reader can verify nothing was edited. `oracle/_ref/` is git-ignored (reference sources
never enter the history) but not gpurun-ignored, so it travels to the GPU box, where
`bench.py --impl reference` and `cpu_baseline` drive the reference's own
`ModelTrainer.epoch` / `RGBF_EmbeddingModel.embed` on the host cores
(`cpu_baseline.kind = "reference"`).

The one shim the reference needs to import (`efficientnet_pytorch`, absent, never used
on the resnet path) lives in oracle/ref_shim.py, not in the copy.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, '_ref')
SRC = os.environ.get('VPD_REFERENCE_DIR', '/root/reference')


def build(force=False):
    """-> path of oracle/_ref, or None when there is no reference to copy and no copy yet."""
    if not os.path.isfile(os.path.join(SRC, 'models', 'rgb.py')):
        return DST if os.path.isfile(os.path.join(DST, 'MANIFEST.json')) else None
    if os.path.isfile(os.path.join(DST, 'MANIFEST.json')) and not force:
        return DST
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    manifest = {}
    for root, dirs, files in os.walk(SRC):
        dirs[:] = sorted(d for d in dirs if not d.startswith('.') and d != '__pycache__')
        for f in sorted(files):
            if not f.endswith('.py'):
                continue
            src = os.path.join(root, f)
            rel = os.path.relpath(src, SRC)
            dst = os.path.join(DST, rel)
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.copyfile(src, dst)
            with open(src, 'rb') as fp:
                manifest[rel] = hashlib.sha256(fp.read()).hexdigest()
    with open(os.path.join(DST, 'MANIFEST.json'), 'w') as fp:
        json.dump({'source': SRC, 'files': manifest}, fp, indent=1, sort_keys=True)
    return DST


def verify():
    """Every file of the copy still has the sha256 recorded when it was made."""
    with open(os.path.join(DST, 'MANIFEST.json')) as fp:
        manifest = json.load(fp)['files']
    for rel, digest in manifest.items():
        with open(os.path.join(DST, rel), 'rb') as fp:
            if hashlib.sha256(fp.read()).hexdigest() != digest:
                return False
    return True


if __name__ == '__main__':
    print(build(force='--force' in sys.argv))

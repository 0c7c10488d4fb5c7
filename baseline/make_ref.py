#!/usr/bin/env python
"""Stage the UNMODIFIED reference for the CPU arm of bench.py (`--impl reference`, `cpu_baseline.kind = "reference"`).

The reference (ruiminshen/yolo2-pytorch) is a script tree without setup.py / pyproject.toml, so `pip install` cannot install it;
this copies the files on the hot path -- model/, utils/, detect.py, config.ini, config/ -- from /root/reference into
`baseline/_ref/` (git-ignored: reference sources never enter this repository's history; the directory travels to the GPU box
with the gpurun snapshot).  One line is patched on the copy: `utils/__init__.py:109` names a parameter `async`, a reserved word
since Python 3.7, so nothing imports without the rename to `non_blocking`.

    python baseline/make_ref.py            # build container only (needs /root/reference); __graft_entry__.build() calls it
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get('YB_REFERENCE_DIR', '/root/reference')
DST = os.path.join(HERE, '_ref')
ITEMS = ('model', 'utils', 'config', 'config.ini', 'detect.py')


def main():
    if not os.path.isdir(REF):
        print('make_ref: %s not present (GPU box?) -- keeping whatever is in %s' % (REF, DST))
        return 0
    os.makedirs(DST, exist_ok=True)
    for item in ITEMS:
        src, dst = os.path.join(REF, item), os.path.join(DST, item)
        if os.path.isdir(src):
            shutil.rmtree(dst, ignore_errors=True)
            shutil.copytree(src, dst, ignore=shutil.ignore_patterns('__pycache__', '*.pyc'))
        else:
            shutil.copy2(src, dst)
    path = os.path.join(DST, 'utils', '__init__.py')
    text = open(path).read()
    patched = text.replace('async=False', 'non_blocking=False').replace('device_id, async)', 'device_id, non_blocking)')
    if patched == text and 'async' in text:
        raise SystemExit('make_ref: the async patch did not apply')
    open(path, 'w').write(patched)
    with open(os.path.join(DST, 'PROVENANCE.txt'), 'w') as f:
        f.write('copied from %s by baseline/make_ref.py; only change: utils/__init__.py async -> non_blocking (Python >= 3.7)\n' % REF)
    print('make_ref: staged %s' % DST)
    return 0


if __name__ == '__main__':
    sys.exit(main())

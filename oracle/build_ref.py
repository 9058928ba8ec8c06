"""Recipe that makes the UNMODIFIED reference travel to the GPU box.  TEST / BENCH INFRASTRUCTURE.

    python -m oracle.build_ref

The reference's hot path is pure Python (`lib/modeling/iodine.py` and the two small modules its import chain pulls
in), so "building" it is a byte-for-byte copy of those files from `/root/reference` into `oracle/_ref/` -- a
git-ignored output directory (never committed; it ships to the GPU box with the repo snapshot like the built `.so`).
`bench.py --impl reference` and the `cpu_baseline` leg then time the reference itself (`kind: "reference"`) instead
of the oracle port.  Nothing under `iodine_b200/` imports it.
"""
import hashlib
import os
import shutil

SRC = os.environ.get('IODINE_REFERENCE_ROOT', '/root/reference')
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref')
# the import chain of `lib.modeling.iodine`: lib/modeling/__init__.py -> build.py -> iodine.py, vae.py;
# iodine.py -> lib/utils/vis_logger.py
FILES = ['lib/modeling/__init__.py', 'lib/modeling/build.py', 'lib/modeling/iodine.py', 'lib/modeling/vae.py',
         'lib/utils/vis_logger.py']
PACKAGE_DIRS = ['lib', 'lib/modeling', 'lib/utils']


def build():
    """Copy the files listed above; returns the manifest {relative path: sha256} (also written to _ref/MANIFEST)."""
    if not os.path.isfile(os.path.join(SRC, FILES[2])):
        return None                                    # no reference tree here (the GPU box): keep what was shipped
    os.makedirs(DST, exist_ok=True)
    manifest = {}
    for d in PACKAGE_DIRS:
        os.makedirs(os.path.join(DST, d), exist_ok=True)
        init_src, init_dst = os.path.join(SRC, d, '__init__.py'), os.path.join(DST, d, '__init__.py')
        if os.path.join(d, '__init__.py') in FILES:
            continue
        if os.path.isfile(init_src):
            shutil.copyfile(init_src, init_dst)
        else:
            open(init_dst, 'w').close()                # namespace package in the reference (lib/utils)
    for f in FILES:
        shutil.copyfile(os.path.join(SRC, f), os.path.join(DST, f))
        manifest[f] = hashlib.sha256(open(os.path.join(DST, f), 'rb').read()).hexdigest()
    with open(os.path.join(DST, 'MANIFEST'), 'w') as fh:
        for f, h in sorted(manifest.items()):
            fh.write('%s  %s\n' % (h, f))
    return manifest


def available():
    return os.path.isfile(os.path.join(DST, FILES[2]))


if __name__ == '__main__':
    m = build()
    print('copied %d files into %s' % (len(m), DST) if m else 'reference tree not present at %s' % SRC)

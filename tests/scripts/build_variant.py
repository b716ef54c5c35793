"""Build a tuning variant of the library: python tests/scripts/build_variant.py <tag> -DXRT_FOO=1 ... -> build/var/libxrt_<tag>.so"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from xicsrt_b200 import build
tag, flags = sys.argv[1], sys.argv[2:]
os.makedirs(os.path.join(ROOT, 'build', 'var'), exist_ok=True)
print(build.build(force=True, extra_flags=flags, lib=os.path.join(ROOT, 'build', 'var', f'libxrt_{tag}.so')))

"""Build an experimental variant of the library next to the product one:
    python tools/build_variant.py <name> [-DFLAG ...]   ->  dronesim_b200/libdronesim_b200.<name>.so
Run with DRONESIM_B200_LIB=<path> python bench.py ...  (kernel experiments only; never shipped)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from dronesim_b200 import _lib

name, flags = sys.argv[1], sys.argv[2:]
out = os.path.join(os.path.dirname(_lib.__file__), "libdronesim_b200.%s.so" % name)
_lib.build(verbose="-v" in flags, force=True, extra_flags=[f for f in flags if f != "-v"], out_path=out)
print(out)

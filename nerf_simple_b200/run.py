"""Run the reference's CLI scripts UNCHANGED on the B200 engine.

    python -m nerf_simple_b200.run /path/to/Nerf-Simple/train.py --config_path configs/lego.yaml

The script is executed with runpy after the `utils` shim package (nerf_simple_b200/dropin) is
put first on sys.path, so its `from utils.nets import Nerf` etc. resolve to this engine.
"""
import os
import runpy
import sys


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        raise SystemExit(__doc__)
    script = argv[0]
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(here, "dropin"))
    sys.path.insert(1, os.path.dirname(here))
    for name in [m for m in sys.modules if m == "utils" or m.startswith("utils.")]:
        del sys.modules[name]
    sys.argv = [script] + argv[1:]
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()

"""``python -m tools.retro.cli <project dir>``: open an interactive shell with ``retro`` initialised on the project."""
import code
import sys

from .cli import retro

if __name__ == "__main__":
    if len(sys.argv) != 2:
        raise SystemExit("usage: python -m tools.retro.cli <retro project dir>")
    retro.init(sys.argv[1])
    retro.print_usage()
    code.interact(local={"retro": retro}, banner="")

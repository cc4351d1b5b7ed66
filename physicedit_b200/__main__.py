"""Launcher: run one of the reference's scripts, unmodified, against this package.

    python -m physicedit_b200 scripts/inference/validate.py --prompt "..." --image_path in.png --save_path out.png ...
    torchrun --nproc-per-node 8 -m physicedit_b200 scripts/train/train_physicedit.py --dataset_base_path ...
    accelerate launch --multi_gpu --num_processes 4 -m physicedit_b200 scripts/train/train_physicedit.py ...      (scripts/train/train_multigpu.sh:13)

`compat.install()` registers the `diffsynth` alias package first, so every `from diffsynth... import ...` of the script resolves to the native
implementation (the script's own `sys.path` line for DiffSynth-Studio then has nothing left to import); the script runs as `__main__` with its own
argument list."""
import runpy
import sys


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv or argv[0] in ("-h", "--help"):
        print(__doc__)
        return 0 if argv else 2
    from . import compat
    compat.install()
    sys.argv = argv                                  # the script sees itself as argv[0]
    runpy.run_path(argv[0], run_name="__main__")
    return 0


if __name__ == "__main__":
    sys.exit(main())

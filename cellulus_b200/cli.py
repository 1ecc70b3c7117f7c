"""`train <config.toml>` / `infer <config.toml>` (`cellulus/cli.py`)."""

from __future__ import annotations

import sys
import tomllib

from cellulus_b200.configs import ExperimentConfig


def _load(path):
    print(f"Reading config from {path}")
    with open(path, "rb") as f:
        return ExperimentConfig(**tomllib.load(f))


def train(argv=None):
    from cellulus_b200.train import train as run

    argv = sys.argv[1:] if argv is None else argv
    run(_load(argv[0]))


def infer(argv=None):
    from cellulus_b200.infer import infer as run

    argv = sys.argv[1:] if argv is None else argv
    run(_load(argv[0]))


if __name__ == "__main__":
    {"train": train, "infer": infer}[sys.argv[1]](sys.argv[2:])

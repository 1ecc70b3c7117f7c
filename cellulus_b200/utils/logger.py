"""CSV loss logger (`cellulus/utils/logger.py` without the matplotlib PNG: plotting is not on the path)."""

from __future__ import annotations

import csv
from typing import Dict, List


class Logger:
    def __init__(self, keys: List[str], title: str):
        self.keys = keys
        self.title = title
        self.data: Dict[str, List[float]] = {k: [] for k in keys}

    def add(self, key, value):
        assert key in self.data, "Key not in data"
        self.data[key].append(value)

    def write(self):
        with open(self.title + ".csv", "w", newline="") as fh:
            w = csv.writer(fh)
            w.writerow([""] + self.keys)
            for i, row in enumerate(zip(*[self.data[k] for k in self.keys])):
                w.writerow([i, *row])

    def plot(self):  # kept for interface compatibility; no-op
        pass


def get_logger(keys: List[str], title: str) -> Logger:
    return Logger(keys, title)

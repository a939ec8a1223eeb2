"""Stat logger with the reference's surface (src/utils/logging.py:5-81): log_stat / print_recent_stats / console_logger.
sacred and tensorboard_logger are not dependencies here; stats are kept in memory and optionally mirrored to a JSON-lines
file."""
import json
import logging
from collections import defaultdict

import numpy as np


class Logger:
    def __init__(self, console_logger, jsonl_path=None):
        self.console_logger = console_logger
        self.stats = defaultdict(list)
        self._jsonl = open(jsonl_path, "a") if jsonl_path else None

    def log_stat(self, key, value, t, to_sacred=True):
        value = float(value)
        self.stats[key].append((t, value))
        if self._jsonl:
            self._jsonl.write(json.dumps({"key": key, "value": value, "t": int(t)}) + "\n")
            self._jsonl.flush()

    def print_recent_stats(self):
        if "episode" not in self.stats:
            return
        head = "Recent Stats | t_env: {:>10} | Episode: {:>8}\n".format(self.stats["episode"][-1][0],
                                                                       int(self.stats["episode"][-1][1]))
        lines, i = [], 0
        for k, v in sorted(self.stats.items()):
            if k == "episode":
                continue
            i += 1
            window = 5 if k != "epsilon" else 1
            item = "{:.4f}".format(np.mean([x[1] for x in v[-window:]]))
            lines.append("{:<25}{:>8}".format(k + ":", item) + ("\n" if i % 4 == 0 else "\t"))
        self.console_logger.info(head + "".join(lines))


def get_logger(name="refil_b200"):
    logger = logging.getLogger(name)
    if not logger.handlers:
        ch = logging.StreamHandler()
        ch.setFormatter(logging.Formatter("[%(levelname)s %(asctime)s] %(name)s %(message)s", "%H:%M:%S"))
        logger.addHandler(ch)
    logger.setLevel(logging.INFO)
    return logger

"""`utils.logger.ExperimentLogger` — run directory with config copy, train/test CSV + JSONL rows and summary.json
(reference utils/logger.py:33-191).  Observability only; plotting needs matplotlib (absent here) and is skipped."""
from __future__ import annotations

import csv
import json
import os
import time
from typing import Any, Dict, Optional


class ExperimentLogger:
    def __init__(self, config: Dict[str, Any]):
        log = (config.get("logging", {}) or {}) if config else {}
        self.enabled = bool(log.get("enabled", False))
        self._run_dir: Optional[str] = None
        self._summary: Dict[str, Any] = {}
        self._fields: Dict[str, list] = {}
        self._flags = {k: {"csv": bool((log.get(k, {}) or {}).get("save_csv", True)),
                           "jsonl": bool((log.get(k, {}) or {}).get("save_jsonl", True))} for k in ("train", "test")}
        if not self.enabled:
            return
        task = config.get("task") or config.get("name") or "run"
        self._run_dir = os.path.join(log.get("root_dir", "runs"), str(task), time.strftime("%Y%m%d_%H%M%S"))
        os.makedirs(self._run_dir, exist_ok=True)
        if log.get("save_config_copy", True):
            with open(os.path.join(self._run_dir, "config.json"), "w") as f:
                json.dump(config, f, indent=2)

    def run_dir(self) -> Optional[str]:
        return self._run_dir

    def _log(self, kind: str, row: Dict[str, Any]) -> None:
        if not self.enabled:
            return
        if self._flags[kind]["jsonl"]:
            with open(os.path.join(self._run_dir, f"{kind}.jsonl"), "a") as f:
                f.write(json.dumps(row) + "\n")
        if self._flags[kind]["csv"]:
            path = os.path.join(self._run_dir, f"{kind}.csv")
            fields = self._fields.setdefault(kind, list(row.keys()))
            new = not os.path.exists(path)
            with open(path, "a", newline="") as f:
                w = csv.DictWriter(f, fieldnames=fields, extrasaction="ignore")
                if new:
                    w.writeheader()
                w.writerow(row)

    def log_train(self, row: Dict[str, Any]) -> None:
        self._log("train", row)

    def log_test(self, row: Dict[str, Any]) -> None:
        self._log("test", row)

    def set_summary(self, summary: Dict[str, Any]) -> None:
        self._summary.update(summary)
        if self.enabled:
            with open(os.path.join(self._run_dir, "summary.json"), "w") as f:
                json.dump(self._summary, f, indent=2)

    def generate_plots(self) -> None:
        try:
            import matplotlib  # noqa: F401
        except Exception:
            return

    def close(self) -> None:
        if self.enabled and self._summary:
            self.set_summary({})

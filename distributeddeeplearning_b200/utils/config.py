"""Project configuration (.env) loader / writer.

Parity: reference ``control/src/config.py:5-15`` (``load_config`` =
``dotenv_values(find_dotenv(raise_error_if_not_found=True))``) and the ``set_key`` calls the
task tree uses to persist choices (``tasks.py:70``, ``scripts/storage.py:78``).  No
python-dotenv in this image, so the (tiny) format is parsed here: ``KEY=VALUE`` lines,
``#`` comments, optional single/double quotes, optional ``export`` prefix.

The Azure keys of the reference's ``_dotenv_template`` map onto local ones (``DEFAULTS``):
a cluster becomes "GPUs of this box", a datastore becomes a directory.
"""
from __future__ import annotations

import os
from collections import OrderedDict
from typing import Dict, Optional

ENV_FILENAME = ".env"

# Local analogue of the reference's 15-key ``_dotenv_template`` (``_dotenv_template:1-15``).
DEFAULTS: "OrderedDict[str, str]" = OrderedDict(
    [
        ("PROJECT_NAME", "b200_ddl_project"),
        ("EXPERIMENT_NAME", "experiment"),
        ("CLUSTER_NAME", "local-8xb200"),      # reference: AmlCompute cluster name
        ("CLUSTER_MIN_NODES", "0"),
        ("CLUSTER_MAX_NODES", "8"),            # reference: max nodes -> here: max GPUs on the box
        ("GPUS_PER_NODE", "1"),                # reference: process_count_per_node=4
        ("RUNS_DIR", "runs"),                  # reference: AzureML run history
        ("DATASTORE_NAME", "datastore"),       # reference: blob datastore -> local directory
        ("DATA", "/data"),                     # reference: DATA (local ImageNet directory)
        ("LOG_CONFIG", ""),                    # reference: LOG_CONFIG (ini file) ; empty = built-in
        ("MASTER_ADDR", "127.0.0.1"),
        ("MASTER_PORT", "29511"),
    ]
)


class ConfigError(IOError):
    pass


def find_dotenv(start: Optional[str] = None, filename: str = ENV_FILENAME,
                raise_error_if_not_found: bool = False) -> str:
    """Walk up from ``start`` (cwd by default) to find ``filename``."""
    cur = os.path.abspath(start or os.getcwd())
    while True:
        cand = os.path.join(cur, filename)
        if os.path.isfile(cand):
            return cand
        parent = os.path.dirname(cur)
        if parent == cur:
            break
        cur = parent
    if raise_error_if_not_found:
        raise ConfigError(f"{filename} not found (searched upward from {start or os.getcwd()})")
    return ""


def _parse_line(line: str):
    s = line.strip()
    if not s or s.startswith("#"):
        return None
    if s.startswith("export "):
        s = s[len("export "):].lstrip()
    if "=" not in s:
        return None
    k, v = s.split("=", 1)
    k, v = k.strip(), v.strip()
    if len(v) >= 2 and v[0] == v[-1] and v[0] in "'\"":
        v = v[1:-1]
    elif " #" in v:
        v = v.split(" #", 1)[0].rstrip()
    return k, v


def dotenv_values(path: str) -> "OrderedDict[str, str]":
    out: "OrderedDict[str, str]" = OrderedDict()
    if not path:
        return out
    with open(path) as f:
        for line in f:
            kv = _parse_line(line)
            if kv:
                out[kv[0]] = kv[1]
    return out


def load_config(start: Optional[str] = None, required: bool = False) -> Dict[str, str]:
    """DEFAULTS overlaid by the nearest ``.env`` overlaid by matching process env vars."""
    cfg: Dict[str, str] = dict(DEFAULTS)
    path = find_dotenv(start, raise_error_if_not_found=required)
    cfg.update(dotenv_values(path))
    for k in list(cfg):
        if k in os.environ:
            cfg[k] = os.environ[k]
    return cfg


def set_key(path: str, key: str, value: str) -> None:
    """Insert or replace ``key`` in the dotenv file at ``path`` (created if missing)."""
    lines = []
    if os.path.isfile(path):
        with open(path) as f:
            lines = f.read().splitlines()
    done = False
    for i, line in enumerate(lines):
        kv = _parse_line(line)
        if kv and kv[0] == key:
            lines[i] = f'{key}="{value}"'
            done = True
    if not done:
        lines.append(f'{key}="{value}"')
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")


def write_env_template(path: str, overrides: Optional[Dict[str, str]] = None) -> str:
    """Render the ``.env`` template (reference ``hooks/post_gen_project.py:16-17`` moves
    ``_dotenv_template`` to ``.env`` after filling cookiecutter variables)."""
    vals = OrderedDict(DEFAULTS)
    vals.update(overrides or {})
    with open(path, "w") as f:
        for k, v in vals.items():
            f.write(f'{k}="{v}"\n')
    return path

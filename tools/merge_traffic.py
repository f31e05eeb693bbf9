#!/usr/bin/env python3
"""Merge the ncu_traffic entries a measurement pass wrote under gpurun_out/ into profiles/ncu_traffic.json
(the table bench.py reads for roofline.frac_pipe):  python tools/merge_traffic.py <new.json> [<table.json>]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    new_path = sys.argv[1]
    table_path = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "profiles", "ncu_traffic.json")
    with open(new_path) as f:
        new = json.load(f)
    for entry in new.values():
        if isinstance(entry, dict) and "source" in entry:
            entry["source"] = entry["source"].replace("gpurun_out/", "profiles/")
    try:
        with open(table_path) as f:
            table = json.load(f)
    except (OSError, ValueError):
        table = {}
    table.update(new)
    with open(table_path, "w") as f:
        json.dump(table, f, indent=1)


if __name__ == "__main__":
    main()

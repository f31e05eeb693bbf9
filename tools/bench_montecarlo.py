#!/usr/bin/env python3
"""The drop-in Monte-Carlo flow end to end on one GPU, files in and database out:

    montecarlo.LHS(..., sample_size=N).run()          (montecarlo.py:132-177 of the reference)
    montecarlo.Best.from_run(...) / Best(...)          (best.py:172-287)

with the time of each stage -- construction (file ingest, Latin Hypercube sample on the host),
run (upload, kernel, gather), database text (formatting + writing), conditioning on the device
against reading the database back.  Usage: python tools/bench_montecarlo.py [N ...] > profiles/rNN_montecarlo.json
"""
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from conftest import load_golden, EXTRA  # noqa: E402


def write_catchment(root):
    """The reference's test catchment as input files (the fixture of tests/test_host_logic.py)."""
    from smartpy_b200.timeframe import from_seconds
    raw = load_golden("catchment_raw")
    d = os.path.join(root, "in", "Catchment")
    os.makedirs(d)
    for name in ("rain", "peva", "flow"):
        with open(os.path.join(d, "Catchment." + name), "w") as f:
            f.write("DateTime,%s\n" % name)
            for t, v in zip(raw[name + "_t"], raw[name + "_v"]):
                f.write("%s,%s\n" % (from_seconds(t).strftime("%Y-%m-%d %H:%M:%S"), "" if np.isnan(v) else repr(float(v))))
    with open(os.path.join(d, "Catchment.parameters"), "w") as f:
        f.write("PAR_NAME,PAR_VALUE\n")
        for n, v in zip(['T', 'C', 'H', 'D', 'S', 'Z', 'SK', 'FK', 'GK', 'RK'], raw["parameters"]):
            f.write("%s,%r\n" % (n, float(v)))
    with open(os.path.join(d, "Catchment.sttngs"), "w") as f:
        f.write("ARGUMENT,VALUE\n")
        for k, v in zip(raw["sttngs_keys"], raw["sttngs_vals"]):
            f.write("%s,%s\n" % (k, v))


def main():
    import torch
    from smartpy_b200 import montecarlo
    from smartpy_b200.montecarlo import database
    sizes = [int(a) for a in sys.argv[1:]] or [100000, 1000000]
    out = {"gpu": torch.cuda.get_device_name(0), "host_cores": os.cpu_count(), "runs": []}
    with tempfile.TemporaryDirectory() as root:
        write_catchment(root)
        for n in sizes:
            np.random.seed(42)
            t0 = time.perf_counter()
            setup = montecarlo.LHS('Catchment', root, 'csv', 'csv', sample_size=n)
            setup.model.extra = dict(EXTRA)
            t1 = time.perf_counter()
            for _ in range(2):                       # the second run is warm (library loaded, buffers there)
                # time the database text on its own, through the module's own writer
                spent = {"text": 0.0}
                inner = database.SampleDatabase.write_rows

                def timed(self, *a, _inner=inner, _spent=spent, **k):
                    s = time.perf_counter()
                    _inner(self, *a, **k)
                    _spent["text"] += time.perf_counter() - s
                database.SampleDatabase.write_rows = timed
                t2 = time.perf_counter()
                setup.run()
                torch.cuda.synchronize()
                t3 = time.perf_counter()
                database.SampleDatabase.write_rows = inner
            size = os.path.getsize(setup.db_file)
            t4 = time.perf_counter()
            best = montecarlo.Best.from_run(setup, 'NSE', 100, constraining={'GW': ('equal', (1.0,))})
            torch.cuda.synchronize()
            t5 = time.perf_counter()
            best_file = montecarlo.Best('Catchment', root, 'csv', 'csv', target='NSE', nb_best=100,
                                        constraining={'GW': ('equal', (1.0,))})
            t6 = time.perf_counter()
            steps = n * (87672 + 8760)
            out["runs"].append({
                "sample_size": n,
                "construct_s (ingest + LHS sample on the host)": t1 - t0,
                "run_s (upload, kernel, download, database)": t3 - t2,
                "of_which_database_text_s": spent["text"],
                "database_bytes": size,
                "member_timesteps_per_s_of_run": steps / (t3 - t2),
                "best_from_run_s (top 100 on the device, score table in HBM)": t5 - t4,
                "best_from_file_s (database read back, host rules)": t6 - t5,
                "same_best_rows_up_to_float32_text": bool(np.allclose(np.sort(best.best_params[:, 0]),
                                                                      np.sort(best_file.best_params[:, 0]), rtol=1e-6)),
            })
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()

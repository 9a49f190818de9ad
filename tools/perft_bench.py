#!/usr/bin/env python
"""perft from the initial position on the GPU (BASELINE.json config 2): leaf counts for depths 1..D checked against
the known-answer table (reference Position semantics: pass = one ply, terminal = leaf) and
nodes/s from the CUDA-event time of the device part.

  python tools/perft_bench.py [--depth 12] [--mode 0|1] [--cpu]

--cpu also times the reference's own Position code (oracle/_ref/libref_shim.so, else the oracle port) on one host core
up to depth 9 as the reported CPU baseline."""
import argparse
import ctypes as C
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "othello-alphazero_b200"))
sys.path.insert(0, str(ROOT / "oracle"))

KNOWN = [4, 12, 56, 244, 1396, 8200, 55092, 390216, 3005288, 24571284, 212258800, 1939886636, 18429641748, 184042084512]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--depth", type=int, default=12)
    ap.add_argument("--mode", type=int, default=0)
    ap.add_argument("--cpu", action="store_true")
    ap.add_argument("--only", action="store_true", help="run only --depth (for profiling one launch)")
    a = ap.parse_args()
    import othello_mcts as om
    out = {"mode": a.mode, "depths": []}
    for d in range(a.depth if a.only else 1, a.depth + 1):
        nodes, ms = om.perft(d, mode=a.mode, return_ms=True)
        ok = nodes == KNOWN[d - 1]
        out["depths"].append({"depth": d, "nodes": nodes, "device_ms": ms, "nodes_per_s": nodes / (ms * 1e-3) if ms > 0 else None,
                              "matches_known_answer": ok})
        assert ok, (d, nodes, KNOWN[d - 1])
    if a.cpu:
        import pyoracle as po
        L = po.lib()
        p = po.Pos()
        L.ozo_initial_position(C.byref(p))
        t0 = time.perf_counter()
        n = L.ozo_perft(C.byref(p), 9)
        dt = time.perf_counter() - t0
        out["cpu_port_1core"] = {"depth": 9, "nodes": int(n), "seconds": dt, "nodes_per_s": n / dt}
    print(json.dumps(out))


if __name__ == "__main__":
    main()

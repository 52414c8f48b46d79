#!/usr/bin/env python3
"""Regenerates tests/golden/ from the reference's own golden recordings (run in the build container only).

The three recordings and their expected samedec stdout are the reference's integration fixtures
(/root/reference/sample/*.22050.s16le.{bin,txt}, driven by sample/test.sh:21-57).  They are DATA, not source:
the .bin files are stored gzip-compressed, the .txt files verbatim.  `/root/reference` does not exist on the GPU
box, so tests read only the copies made here.

Also writes oracle_events.json: the CPU oracle's link+transport event trace for each recording (samedec config,
EOF flush included).  The oracle is pinned by the .txt files; its trace is then the golden for event sample indices
(no reference test pins those bit-for-bit, see oracle/same_oracle.hpp header).
"""
import gzip
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/sample"
NAMES = ["long_message", "npt", "two_and_two"]


def main():
    sys.path.insert(0, os.path.join(HERE, "..", ".."))
    for n in NAMES:
        src = os.path.join(REF, f"{n}.22050.s16le.bin")
        with open(src, "rb") as f, gzip.GzipFile(os.path.join(HERE, f"{n}.22050.s16le.bin.gz"), "wb", mtime=0) as g:
            shutil.copyfileobj(f, g)
        shutil.copyfile(os.path.join(REF, f"{n}.22050.s16le.txt"), os.path.join(HERE, f"{n}.22050.s16le.txt"))
    from oracle import Oracle, load_golden_recording

    out = {}
    for n in NAMES:
        o = Oracle.samedec(22050)
        o.process_s16(load_golden_recording(n))
        o.flush_samedec()
        out[n] = [e.to_json() for e in o.events()]
    with open(os.path.join(HERE, "oracle_events.json"), "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()

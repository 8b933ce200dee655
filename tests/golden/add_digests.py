#!/usr/bin/env python
"""Adds "digest" (gbin_table_digest, hex) to every A/C/G/T-only, K >= 2M case of pins.json.

The digest is computed over the oracle's table AFTER checking that the table's sorted dump has the md5 that the unmodified
reference binary produced (the pin), so it is anchored on reference execution like the md5 itself.
    python tests/golden/add_digests.py"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle_lib as O  # noqa: E402

path = os.path.join(HERE, "pins.json")
pins = json.load(open(path))
for c in pins["cases"]:
    if not (c["acgt_only"] and c["k"] >= 2 * c["m"]):
        continue
    data = O.load_case_bytes(c)
    starts, lens = O.fgets_split(data, c["read_length_define"])
    t = O.run(data, starts, lens, c["k"], c["m"], c["cutoff"])
    assert t.md5() == c["md5"], c["name"]
    c["digest"] = "%016x" % O.table_digest(t)
    print(c["name"], c["digest"])
json.dump(pins, open(path, "w"), indent=1)

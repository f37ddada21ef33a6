"""Derives tests/golden/air_model_index.json from the reference's shipped checkpoint index
(/root/reference/model/air-model.index): every variable's name, dtype, shape, offset, size and
stored CRC, plus the SHA-256 of the original file.  The JSON (derived data, not the file) is what
the tests use, so nothing under /root/reference is read at test time.

    python tests/golden/make_index_fixture.py
"""
import hashlib
import importlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
ck = importlib.import_module("tf-attend-infer-repeat_b200.checkpoint")

SRC = "/root/reference/model/air-model.index"
entries = ck.read_index(SRC)
raw = open(SRC, "rb").read()
out = {"source": "model/air-model.index of aakhundov/tf-attend-infer-repeat", "sha256": hashlib.sha256(raw).hexdigest(),
       "bytes": len(raw),
       "entries": [{"name": k, "dtype": e.dtype, "shape": list(e.shape), "shard_id": e.shard_id, "offset": e.offset,
                    "size": e.size, "crc32c": e.crc32c} for k, e in entries.items()]}
json.dump(out, open(os.path.join(HERE, "air_model_index.json"), "w"), indent=0)
print(len(entries), "entries;", out["sha256"])

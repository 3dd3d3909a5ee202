#!/usr/bin/env python
"""Drop cubins that can no longer be hit: the cache key hashes the device sources, so every cubin older than
the newest device header is stale.  Keeps the snapshot sent to the GPU box small."""
import glob
import os

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
newest = max(os.path.getmtime(p) for p in glob.glob(os.path.join(REPO, "clode_b200", "csrc", "device", "*")))
n = 0
for p in glob.glob(os.path.join(REPO, "clode_b200", "_cubin_cache", "*")):
    if os.path.getmtime(p) < newest:
        os.remove(p)
        n += 1
print(f"removed {n} stale cache entries")

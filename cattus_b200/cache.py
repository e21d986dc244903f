"""`ValueFuncCache` (engine/src/mcts/cache.rs:31-75): position -> (probs, value) memo with FIFO eviction.
Host-side, shared by the worker threads of one evaluator; counts hits and misses like `cache.hits` / `cache.misses`."""
from __future__ import annotations

import threading
from collections import OrderedDict


class ValueFuncCache:
    def __init__(self, max_size: int):
        assert max_size > 0
        self.max_size = max_size
        self._lock = threading.Lock()
        self._map: "OrderedDict[object, object]" = OrderedDict()
        self.hits = 0
        self.misses = 0

    def get_or_compute(self, key, compute):
        with self._lock:
            if key in self._map:
                self.hits += 1
                return self._map[key]
        val = compute()  # outside the lock: many threads may evaluate at once (cache.rs:44-53)
        with self._lock:
            self.misses += 1
            if key not in self._map:  # another thread may have inserted the same position meanwhile (cache.rs:55-59)
                while len(self._map) >= self.max_size:
                    self._map.popitem(last=False)  # FIFO: oldest insertion goes first
                self._map[key] = val
        return val

    def metrics(self) -> dict:
        return {"cache.hits": self.hits, "cache.misses": self.misses}
